// ref_harness.cpp — TEST INFRASTRUCTURE.  Thin C API around the reference's own
// software rasteriser translation unit, which the Makefile beside this file
// compiles IN PLACE from /root/reference/src/core/gpu_sw_rasterizer.cpp (nothing
// of the reference is copied into this repository).  Output: oracle/_ref/libpsxref.so.
//
// What is reference code here: every pixel (DrawTriangle/DrawRectangle/DrawLine,
// FillVRAM/WriteVRAM/CopyVRAM, UpdateCLUT) and the scan-out pixel conversions
// (ConvertVRAMPixel, header-only, src/core/gpu_helpers.h:107-134).
// What is restated here (the ~120 lines of glue the reference keeps in classes that
// cannot be linked without the whole emulator): the HandleCommand switch
// (src/core/gpu_backend.cpp:383-544), GPU_SW::Draw* (src/core/gpu_sw.cpp:105-190),
// the clamped drawing area (src/core/gpu.cpp:2190-2200) and the scan-out
// row/column addressing (src/core/gpu_sw.cpp:230-356, :391-448).
#include "core/gpu.h"
#include "core/gpu_helpers.h"
#include "core/gpu_sw_rasterizer.h"
#include "core/video_thread_commands.h"

#include "common/log.h"
#include "common/string_util.h"

#include "../../include/psxgpu_wire.h"

#include <cstring>
#include <strings.h>

// ---- the five symbols the rasteriser TU leaves undefined ----
u16 g_vram[VRAM_SIZE / sizeof(u16)];
u16 g_gpu_clut[GPU_CLUT_SIZE];
namespace Log {
Level GetLogLevel()
{
  return Level::None;
}
void Write(MessageCategory, std::string_view)
{
}
} // namespace Log
int StringUtil::Strcasecmp(const char* a, const char* b)
{
  return strcasecmp(a, b);
}
extern "C" bool cpuinfo_has_x86_avx2()
{
  return false;
}

// ---- layout pins: our wire structs == the reference's ----
#define SAME_SIZE(ours, theirs) static_assert(sizeof(ours) == sizeof(theirs), #ours)
static_assert(sizeof(psx_cmd_header) == sizeof(VideoThreadCommand));
static_assert(sizeof(psx_draw_header) == sizeof(GPUBackendDrawCommand));
static_assert(offsetof(psx_draw_header, num_vertices) == offsetof(GPUBackendDrawCommand, num_vertices));
static_assert(offsetof(psx_draw_header, draw_mode) == offsetof(GPUBackendDrawCommand, draw_mode));
static_assert(offsetof(psx_draw_header, palette) == offsetof(GPUBackendDrawCommand, palette));
static_assert(offsetof(psx_draw_header, window) == offsetof(GPUBackendDrawCommand, window));
SAME_SIZE(psx_poly_vertex, GPUBackendDrawPolygonCommand::Vertex);
SAME_SIZE(psx_precise_poly_vertex, GPUBackendDrawPrecisePolygonCommand::Vertex);
static_assert(offsetof(psx_draw_precise_polygon, vertices) == sizeof(GPUBackendDrawPrecisePolygonCommand));
SAME_SIZE(psx_draw_rectangle, GPUBackendDrawRectangleCommand);
static_assert(offsetof(psx_draw_rectangle, x) == offsetof(GPUBackendDrawRectangleCommand, x));
static_assert(offsetof(psx_draw_rectangle, color) == offsetof(GPUBackendDrawRectangleCommand, color));
static_assert(offsetof(psx_draw_rectangle, texcoord) == offsetof(GPUBackendDrawRectangleCommand, texcoord));
SAME_SIZE(psx_line_vertex, GPUBackendDrawLineCommand::Vertex);
SAME_SIZE(psx_precise_line_vertex, GPUBackendDrawPreciseLineCommand::Vertex);
SAME_SIZE(psx_fill_vram, GPUBackendFillVRAMCommand);
static_assert(offsetof(psx_fill_vram, color) == offsetof(GPUBackendFillVRAMCommand, color));
static_assert(offsetof(psx_fill_vram, interlaced_rendering) == offsetof(GPUBackendFillVRAMCommand, interlaced_rendering));
static_assert(offsetof(psx_update_vram, data) == offsetof(GPUBackendUpdateVRAMCommand, data));
static_assert(offsetof(psx_update_vram, set_mask_while_drawing) ==
              offsetof(GPUBackendUpdateVRAMCommand, set_mask_while_drawing));
SAME_SIZE(psx_copy_vram, GPUBackendCopyVRAMCommand);
static_assert(offsetof(psx_copy_vram, set_mask_while_drawing) == offsetof(GPUBackendCopyVRAMCommand, set_mask_while_drawing));
SAME_SIZE(psx_set_drawing_area, GPUBackendSetDrawingAreaCommand);
SAME_SIZE(psx_update_clut, GPUBackendUpdateCLUTCommand);
static_assert(offsetof(psx_update_clut, clut_is_8bit) == offsetof(GPUBackendUpdateCLUTCommand, clut_is_8bit));
SAME_SIZE(psx_read_vram, GPUBackendReadVRAMCommand);
SAME_SIZE(psx_update_display, GPUBackendUpdateDisplayCommand);
static_assert(offsetof(psx_update_display, X) == offsetof(GPUBackendUpdateDisplayCommand, X));
static_assert(offsetof(psx_update_display, display_vram_left) == offsetof(GPUBackendUpdateDisplayCommand, display_vram_left));
SAME_SIZE(psx_load_state, GPUBackendLoadStateCommand);
static_assert(PSX_CMD_CLEAR_VRAM == (int)VideoThreadCommandType::ClearVRAM);
static_assert(PSX_CMD_UPDATE_DISPLAY == (int)VideoThreadCommandType::UpdateDisplay);
static_assert(PSX_CMD_LOAD_STATE == (int)VideoThreadCommandType::LoadState);
static_assert(PSX_CMD_READ_VRAM == (int)VideoThreadCommandType::ReadVRAM);
static_assert(PSX_CMD_FILL_VRAM == (int)VideoThreadCommandType::FillVRAM);
static_assert(PSX_CMD_UPDATE_VRAM == (int)VideoThreadCommandType::UpdateVRAM);
static_assert(PSX_CMD_COPY_VRAM == (int)VideoThreadCommandType::CopyVRAM);
static_assert(PSX_CMD_SET_DRAWING_AREA == (int)VideoThreadCommandType::SetDrawingArea);
static_assert(PSX_CMD_UPDATE_CLUT == (int)VideoThreadCommandType::UpdateCLUT);
static_assert(PSX_CMD_CLEAR_CACHE == (int)VideoThreadCommandType::ClearCache);
static_assert(PSX_CMD_DRAW_POLYGON == (int)VideoThreadCommandType::DrawPolygon);
static_assert(PSX_CMD_DRAW_PRECISE_POLYGON == (int)VideoThreadCommandType::DrawPrecisePolygon);
static_assert(PSX_CMD_DRAW_RECTANGLE == (int)VideoThreadCommandType::DrawRectangle);
static_assert(PSX_CMD_DRAW_LINE == (int)VideoThreadCommandType::DrawLine);
static_assert(PSX_CMD_DRAW_PRECISE_LINE == (int)VideoThreadCommandType::DrawPreciseLine);
static_assert(PSX_CMD_BUFFER_SWAPPED == (int)VideoThreadCommandType::BufferSwapped);
static_assert(PSX_CMD_SUBMIT_FRAME == (int)VideoThreadCommandType::SubmitFrame);
static_assert(PSX_CMD_CLEAR_DISPLAY == (int)VideoThreadCommandType::ClearDisplay);

namespace {
bool s_modulation_crop = false; // g_gpu_settings.gpu_modulation_crop (settings.h:83)
s32 s_clamped[4] = {0, 0, 0, 0}; // left, top, right(excl), bottom(excl): gpu.cpp:2190-2200

void SetClampedArea(const GPUDrawingArea& a)
{
  if (a.left > a.right || a.top > a.bottom)
  {
    s_clamped[0] = s_clamped[1] = s_clamped[2] = s_clamped[3] = 0;
    return;
  }
  s_clamped[2] = static_cast<s32>(std::min(a.right + 1, static_cast<u32>(VRAM_WIDTH)));
  s_clamped[0] = static_cast<s32>(std::min(a.left, std::min(a.right, VRAM_WIDTH - 1)));
  s_clamped[3] = static_cast<s32>(std::min(a.bottom + 1, static_cast<u32>(VRAM_HEIGHT)));
  s_clamped[1] = static_cast<s32>(std::min(a.top, std::min(a.bottom, VRAM_HEIGHT - 1)));
}

GPU_SW_Rasterizer::TextureModulationMode ModMode(const GPUBackendDrawCommand* c)
{
  return GPU_SW_Rasterizer::GetModulationMode(c->texture_enable, c->raw_texture_enable, s_modulation_crop);
}

// gpu_sw.cpp:105-116
void DrawPolygon(const GPUBackendDrawPolygonCommand* cmd)
{
  const auto fn = GPU_SW_Rasterizer::GetDrawTriangleFunction(cmd->shading_enable, ModMode(cmd), cmd->transparency_enable);
  fn(cmd, &cmd->vertices[0], &cmd->vertices[1], &cmd->vertices[2]);
  if (cmd->num_vertices > 3)
    fn(cmd, &cmd->vertices[2], &cmd->vertices[1], &cmd->vertices[3]);
}

// gpu_sw.cpp:118-139
void DrawPrecisePolygon(const GPUBackendDrawPrecisePolygonCommand* cmd)
{
  const auto fn = GPU_SW_Rasterizer::GetDrawTriangleFunction(cmd->shading_enable, ModMode(cmd), cmd->transparency_enable);
  GPUBackendDrawPolygonCommand::Vertex v[4];
  for (u32 i = 0; i < cmd->num_vertices && i < 4; i++)
  {
    const auto& s = cmd->vertices[i];
    v[i].x = s.native_x;
    v[i].y = s.native_y;
    v[i].color = s.color;
    v[i].texcoord = s.texcoord;
  }
  fn(cmd, &v[0], &v[1], &v[2]);
  if (cmd->num_vertices > 3)
    fn(cmd, &v[2], &v[1], &v[3]);
}

// gpu_sw.cpp:141-161 (rintersect + rempty on [x, y, x+w, y+h))
void DrawSprite(const GPUBackendDrawRectangleCommand* cmd)
{
  const s32 l = std::max(s_clamped[0], cmd->x), t = std::max(s_clamped[1], cmd->y);
  const s32 r = std::min(s_clamped[2], cmd->x + static_cast<s32>(cmd->width));
  const s32 b = std::min(s_clamped[3], cmd->y + static_cast<s32>(cmd->height));
  if (r <= l || b <= t)
    return;
  GPU_SW_Rasterizer::GetDrawRectangleFunction(ModMode(cmd), cmd->transparency_enable)(cmd);
}

// gpu_sw.cpp:163-190
void DrawLine(const GPUBackendDrawLineCommand* cmd)
{
  const auto fn = GPU_SW_Rasterizer::GetDrawLineFunction(cmd->shading_enable, cmd->transparency_enable);
  for (u16 i = 0; i + 1 < cmd->num_vertices; i += 2)
    fn(cmd, &cmd->vertices[i], &cmd->vertices[i + 1]);
}
void DrawPreciseLine(const GPUBackendDrawPreciseLineCommand* cmd)
{
  const auto fn = GPU_SW_Rasterizer::GetDrawLineFunction(cmd->shading_enable, cmd->transparency_enable);
  for (u32 i = 0; i + 1 < cmd->num_vertices; i += 2)
  {
    const auto& a = cmd->vertices[i];
    const auto& b = cmd->vertices[i + 1];
    GPUBackendDrawLineCommand::Vertex v[2];
    v[0].Set(a.native_x, a.native_y, a.color);
    v[1].Set(b.native_x, b.native_y, b.color);
    fn(cmd, &v[0], &v[1]);
  }
}

template<GPUTextureFormat fmt>
void CopyOut15(u32 src_x, u32 src_y, u32 width, u32 height, u32 line_skip, u8* dst, u32 dst_stride)
{
  // gpu_sw.cpp:241-285 (the scalar tail loop is used for every pixel; the vector body is the same conversion)
  if ((src_x + width) <= VRAM_WIDTH && (src_y + height) <= VRAM_HEIGHT)
  {
    const u16* src_ptr = &g_vram[src_y * VRAM_WIDTH + src_x];
    const u32 src_step = VRAM_WIDTH << line_skip;
    for (u32 row = 0; row < height; row++)
    {
      const u16* s = src_ptr;
      u8* d = dst;
      for (u32 x = 0; x < width; x++)
        ConvertVRAMPixel<fmt>(d, *(s++));
      src_ptr += src_step;
      dst += dst_stride;
    }
  }
  else
  {
    const u32 end_x = src_x + width;
    const u32 y_step = (1u << line_skip);
    for (u32 row = 0; row < height; row++)
    {
      const u16* s = &g_vram[(src_y % VRAM_HEIGHT) * VRAM_WIDTH];
      u8* d = dst;
      for (u32 col = src_x; col < end_x; col++)
        ConvertVRAMPixel<fmt>(d, s[col % VRAM_WIDTH]);
      src_y += y_step;
      dst += dst_stride;
    }
  }
}

void CopyOut24(u32 src_x, u32 src_y, u32 skip_x, u32 width, u32 height, u32 line_skip, u8* dst, u32 dst_stride)
{
  // gpu_sw.cpp:305-348
  if ((src_x + width) <= VRAM_WIDTH && (src_y + (height << line_skip)) <= VRAM_HEIGHT)
  {
    const u8* src_ptr = reinterpret_cast<const u8*>(&g_vram[src_y * VRAM_WIDTH + src_x]) + (skip_x * 3);
    const u32 src_stride = (VRAM_WIDTH << line_skip) * sizeof(u16);
    for (u32 row = 0; row < height; row++)
    {
      const u8* s = src_ptr;
      u8* d = dst;
      for (u32 col = 0; col < width; col++)
      {
        *(d++) = *(s++);
        *(d++) = *(s++);
        *(d++) = *(s++);
        *(d++) = 0xFF;
      }
      src_ptr += src_stride;
      dst += dst_stride;
    }
  }
  else
  {
    const u32 y_step = (1u << line_skip);
    for (u32 row = 0; row < height; row++)
    {
      const u16* s = &g_vram[(src_y % VRAM_HEIGHT) * VRAM_WIDTH];
      u32* d = reinterpret_cast<u32*>(dst);
      for (u32 col = 0; col < width; col++)
      {
        const u32 offset = (src_x + (((skip_x + col) * 3) / 2));
        const u16 s0 = s[offset % VRAM_WIDTH];
        const u16 s1 = s[(offset + 1) % VRAM_WIDTH];
        const u8 shift = static_cast<u8>(col & 1u) * 8;
        const u32 rgb = (((ZeroExtend32(s1) << 16) | ZeroExtend32(s0)) >> shift);
        *(d++) = rgb | 0xFF000000u;
      }
      src_y += y_step;
      dst += dst_stride;
    }
  }
}
} // namespace

extern "C" {

void psxref_init(void)
{
  GPU_SW_Rasterizer::SelectImplementation(); // sticky; honours SW_USE_ISA (gpu_sw_rasterizer.cpp:87-131)
}

void psxref_set_modulation_crop(int on)
{
  s_modulation_crop = (on != 0);
}

void psxref_reset(void)
{
  std::memset(g_vram, 0, sizeof(g_vram));
  std::memset(g_gpu_clut, 0, sizeof(g_gpu_clut));
  GPU_SW_Rasterizer::g_drawing_area = GPUDrawingArea{0, 0, 0, 0};
  SetClampedArea(GPU_SW_Rasterizer::g_drawing_area);
}

u16* psxref_vram(void)
{
  return g_vram;
}
u16* psxref_clut(void)
{
  return g_gpu_clut;
}

// One record.  gpu_backend.cpp:383-544.
static void HandleCommand(const VideoThreadCommand* cmd)
{
  switch (cmd->type)
  {
    case VideoThreadCommandType::ClearVRAM:
      std::memset(g_vram, 0, sizeof(g_vram));
      std::memset(g_gpu_clut, 0, sizeof(g_gpu_clut));
      break;
    case VideoThreadCommandType::LoadState:
    {
      const auto* c = static_cast<const GPUBackendLoadStateCommand*>(cmd);
      std::memcpy(g_vram, c->vram_data, sizeof(g_vram));
      std::memcpy(g_gpu_clut, c->clut_data, sizeof(g_gpu_clut));
    }
    break;
    case VideoThreadCommandType::FillVRAM:
    {
      const auto* c = static_cast<const GPUBackendFillVRAMCommand*>(cmd);
      GPU_SW_Rasterizer::FillVRAM(c->x, c->y, c->width, c->height, c->color, c->interlaced_rendering, c->active_line_lsb);
    }
    break;
    case VideoThreadCommandType::UpdateVRAM:
    {
      const auto* c = static_cast<const GPUBackendUpdateVRAMCommand*>(cmd);
      GPU_SW_Rasterizer::WriteVRAM(c->x, c->y, c->width, c->height, c->data, c->set_mask_while_drawing,
                                   c->check_mask_before_draw);
    }
    break;
    case VideoThreadCommandType::CopyVRAM:
    {
      const auto* c = static_cast<const GPUBackendCopyVRAMCommand*>(cmd);
      GPU_SW_Rasterizer::CopyVRAM(c->src_x, c->src_y, c->dst_x, c->dst_y, c->width, c->height, c->set_mask_while_drawing,
                                  c->check_mask_before_draw);
    }
    break;
    case VideoThreadCommandType::SetDrawingArea:
    {
      const auto* c = static_cast<const GPUBackendSetDrawingAreaCommand*>(cmd);
      GPU_SW_Rasterizer::g_drawing_area = c->new_area;
      SetClampedArea(c->new_area);
    }
    break;
    case VideoThreadCommandType::UpdateCLUT:
    {
      const auto* c = static_cast<const GPUBackendUpdateCLUTCommand*>(cmd);
      GPU_SW_Rasterizer::UpdateCLUT(c->reg, c->clut_is_8bit);
    }
    break;
    case VideoThreadCommandType::DrawPolygon:
      DrawPolygon(static_cast<const GPUBackendDrawPolygonCommand*>(cmd));
      break;
    case VideoThreadCommandType::DrawPrecisePolygon:
      DrawPrecisePolygon(static_cast<const GPUBackendDrawPrecisePolygonCommand*>(cmd));
      break;
    case VideoThreadCommandType::DrawRectangle:
      DrawSprite(static_cast<const GPUBackendDrawRectangleCommand*>(cmd));
      break;
    case VideoThreadCommandType::DrawLine:
      DrawLine(static_cast<const GPUBackendDrawLineCommand*>(cmd));
      break;
    case VideoThreadCommandType::DrawPreciseLine:
      DrawPreciseLine(static_cast<const GPUBackendDrawPreciseLineCommand*>(cmd));
      break;
    default:
      // ReadVRAM / ClearCache / BufferSwapped / UpdateDisplay / SubmitFrame: no VRAM effect on GPU_SW
      break;
  }
}

// Runs a stream of wire records.  Returns the number of records, or -1 on a malformed stream.
long psxref_exec(const void* cmds, size_t bytes)
{
  const u8* p = static_cast<const u8*>(cmds);
  const u8* end = p + bytes;
  long n = 0;
  while (p < end)
  {
    const VideoThreadCommand* c = reinterpret_cast<const VideoThreadCommand*>(p);
    if (c->size < sizeof(VideoThreadCommand) || (c->size & 3u) != 0 || p + c->size > end)
      return -1;
    HandleCommand(c);
    p += c->size;
    n++;
  }
  return n;
}

// GPU_SW::UpdateDisplay (gpu_sw.cpp:391-448) minus presentation.  fmt: 0 RGBA8, 1 RGB5A1, 2 A1BGR5, 3 RGB565.
// Returns 0 and the output size; 1 if the display is disabled (nothing written).
int psxref_scanout(const psx_update_display* ud, int fmt, int show_vram, void* dst, u32 dst_stride, u32* out_w, u32* out_h)
{
  const auto* cmd = reinterpret_cast<const GPUBackendUpdateDisplayCommand*>(ud);
  u32 src_x, src_y, skip_x, width, height, line_skip;
  bool is_24bit;
  if (show_vram)
  {
    src_x = src_y = skip_x = line_skip = 0;
    width = VRAM_WIDTH;
    height = VRAM_HEIGHT;
    is_24bit = false;
  }
  else
  {
    if (cmd->display_disabled)
      return 1;
    is_24bit = cmd->display_24bit;
    const u32 interleaved = BoolToUInt32(cmd->interlaced_display_interleaved);
    src_x = is_24bit ? cmd->X : cmd->display_vram_left;
    skip_x = is_24bit ? (cmd->display_vram_left - cmd->X) : 0;
    src_y = cmd->display_vram_top + (BoolToUInt8(cmd->interlaced_display_field) & BoolToUInt8(cmd->interlaced_display_interleaved));
    width = cmd->display_vram_width;
    height = cmd->display_vram_height;
    line_skip = cmd->interlaced_display_enabled ? interleaved : 0;
  }
  *out_w = width;
  *out_h = height;
  u8* d = static_cast<u8*>(dst);
  if (is_24bit)
  {
    CopyOut24(src_x, src_y, skip_x, width, height, line_skip, d, dst_stride);
    return 0;
  }
  switch (fmt)
  {
    case 0: CopyOut15<GPUTextureFormat::RGBA8>(src_x, src_y, width, height, line_skip, d, dst_stride); break;
    case 1: CopyOut15<GPUTextureFormat::RGB5A1>(src_x, src_y, width, height, line_skip, d, dst_stride); break;
    case 2: CopyOut15<GPUTextureFormat::A1BGR5>(src_x, src_y, width, height, line_skip, d, dst_stride); break;
    case 3: CopyOut15<GPUTextureFormat::RGB565>(src_x, src_y, width, height, line_skip, d, dst_stride); break;
    default: return -1;
  }
  return 0;
}

} // extern "C"
