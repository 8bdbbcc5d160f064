// backend.cpp — see backend.h.
#include "backend.h"

#include <cstdio>
#include <cstring>

namespace psx {

CudaGpuBackend::CudaGpuBackend() = default;

CudaGpuBackend::~CudaGpuBackend()
{
  if (m_owns && m_ctx)
    psxgpu_destroy(m_ctx);
}

bool CudaGpuBackend::Initialize(bool upload_vram, std::string* error, int device, const psxgpu_opts* opts)
{
  if (m_ctx)
    return true;
  psxgpu_opts o{};
  if (opts)
    o = *opts;
  o.struct_size = sizeof(o);
  const int rc = psxgpu_create(device, 1, &o, &m_ctx);
  if (rc != PSXGPU_OK)
  {
    if (error)
      *error = "psxgpu_create failed (no usable CUDA device; this backend has no CPU fallback)";
    m_ctx = nullptr;
    return false;
  }
  m_owns = true;
  m_instance = 0;
  if (upload_vram) // "old" VRAM is kept: push the host shadow to the device (gpu_sw.cpp:54-56)
    if (psxgpu_load_state(m_ctx, 0, host_vram(), nullptr) != PSXGPU_OK)
    {
      if (error)
        *error = psxgpu_last_error(m_ctx);
      return false;
    }
  return true;
}

bool CudaGpuBackend::Attach(psxgpu_ctx* shared, uint32_t instance, std::string* error)
{
  if (!shared)
  {
    if (error)
      *error = "null context";
    return false;
  }
  m_ctx = shared;
  m_instance = instance;
  m_owns = false;
  return true;
}

void CudaGpuBackend::fatal(const char* what)
{
  m_error = std::string(what) + ": " + (m_ctx ? psxgpu_last_error(m_ctx) : "no context");
  if (fatal_error_hook)
    fatal_error_hook(m_error.c_str());
  else
    std::fprintf(stderr, "CudaGpuBackend fatal: %s\n", m_error.c_str());
}

void CudaGpuBackend::submit(const void* rec, size_t bytes)
{
  if (!m_ctx)
    return fatal("backend used before Initialize");
  if (psxgpu_submit_wire(m_ctx, m_instance, rec, bytes) < 0)
    fatal("psxgpu_submit_wire");
}

void CudaGpuBackend::FlushRender()
{
  if (m_ctx && psxgpu_flush(m_ctx) < 0)
    fatal("psxgpu_flush");
}

void CudaGpuBackend::HandleCommand(const psx_cmd_header* cmd)
{
  switch (cmd->type)
  {
    case PSX_CMD_CLEAR_VRAM: ClearVRAM(); break;
    case PSX_CMD_LOAD_STATE: LoadState(reinterpret_cast<const psx_load_state*>(cmd)); break;
    case PSX_CMD_UPDATE_DISPLAY: UpdateDisplay(reinterpret_cast<const psx_update_display*>(cmd)); break;
    case PSX_CMD_CLEAR_CACHE: ClearCache(); break;
    case PSX_CMD_BUFFER_SWAPPED: OnBufferSwapped(); break;
    case PSX_CMD_READ_VRAM:
    {
      const psx_read_vram* c = reinterpret_cast<const psx_read_vram*>(cmd);
      m_counters.num_reads++;
      ReadVRAM(c->x, c->y, c->width, c->height);
    }
    break;
    case PSX_CMD_FILL_VRAM:
    {
      const psx_fill_vram* c = reinterpret_cast<const psx_fill_vram*>(cmd);
      FillVRAM(c->x, c->y, c->width, c->height, c->color, c->interlaced_rendering != 0, c->active_line_lsb);
    }
    break;
    case PSX_CMD_UPDATE_VRAM:
      m_counters.num_writes++;
      submit(cmd, cmd->size); // the record already carries its payload
      break;
    case PSX_CMD_COPY_VRAM:
      m_counters.num_copies++;
      submit(cmd, cmd->size);
      break;
    case PSX_CMD_SET_DRAWING_AREA:
      SetDrawingArea(reinterpret_cast<const psx_set_drawing_area*>(cmd)->new_area);
      DrawingAreaChanged();
      break;
    case PSX_CMD_UPDATE_CLUT:
    {
      const psx_update_clut* c = reinterpret_cast<const psx_update_clut*>(cmd);
      UpdateCLUT(c->reg, c->clut_is_8bit != 0);
    }
    break;
    case PSX_CMD_DRAW_POLYGON: DrawPolygon(reinterpret_cast<const psx_draw_polygon*>(cmd)); break;
    case PSX_CMD_DRAW_PRECISE_POLYGON: DrawPrecisePolygon(reinterpret_cast<const psx_draw_precise_polygon*>(cmd)); break;
    case PSX_CMD_DRAW_RECTANGLE: DrawSprite(reinterpret_cast<const psx_draw_rectangle*>(cmd)); break;
    case PSX_CMD_DRAW_LINE: DrawLine(reinterpret_cast<const psx_draw_line*>(cmd)); break;
    case PSX_CMD_DRAW_PRECISE_LINE: DrawPreciseLine(reinterpret_cast<const psx_draw_header*>(cmd)); break;
    default: break; // ClearDisplay / SubmitFrame: presentation, outside this path
  }
}

void CudaGpuBackend::ReadVRAM(uint32_t x, uint32_t y, uint32_t width, uint32_t height)
{
  psx_read_vram c{};
  c.h.size = sizeof(c);
  c.h.type = PSX_CMD_READ_VRAM;
  c.x = static_cast<uint16_t>(x);
  c.y = static_cast<uint16_t>(y);
  c.width = static_cast<uint16_t>(width);
  c.height = static_cast<uint16_t>(height);
  submit(&c, sizeof(c));
}

void CudaGpuBackend::FillVRAM(uint32_t x, uint32_t y, uint32_t width, uint32_t height, uint32_t color, bool interlaced_rendering,
                              uint8_t active_line_lsb)
{
  psx_fill_vram c{};
  c.h.size = sizeof(c) + 8; // records are multiples of 16 bytes (video_thread_commands.h:63-68)
  c.h.type = PSX_CMD_FILL_VRAM;
  c.x = static_cast<uint16_t>(x);
  c.y = static_cast<uint16_t>(y);
  c.width = static_cast<uint16_t>(width);
  c.height = static_cast<uint16_t>(height);
  c.color = color;
  c.interlaced_rendering = interlaced_rendering;
  c.active_line_lsb = active_line_lsb;
  uint8_t rec[32] = {};
  std::memcpy(rec, &c, sizeof(c));
  submit(rec, sizeof(rec));
}

void CudaGpuBackend::UpdateVRAM(uint32_t x, uint32_t y, uint32_t width, uint32_t height, const void* data, bool set_mask,
                                bool check_mask)
{
  const size_t payload = static_cast<size_t>(width) * height * 2;
  const size_t size = (PSX_UPDATE_VRAM_DATA_OFFSET + payload + 15) & ~static_cast<size_t>(15);
  m_scratch.assign(size, 0);
  psx_update_vram* c = reinterpret_cast<psx_update_vram*>(m_scratch.data());
  c->h.size = static_cast<uint32_t>(size);
  c->h.type = PSX_CMD_UPDATE_VRAM;
  c->x = static_cast<uint16_t>(x);
  c->y = static_cast<uint16_t>(y);
  c->width = static_cast<uint16_t>(width);
  c->height = static_cast<uint16_t>(height);
  c->set_mask_while_drawing = set_mask;
  c->check_mask_before_draw = check_mask;
  std::memcpy(m_scratch.data() + PSX_UPDATE_VRAM_DATA_OFFSET, data, payload);
  m_counters.num_writes++;
  submit(m_scratch.data(), size);
}

void CudaGpuBackend::CopyVRAM(uint32_t src_x, uint32_t src_y, uint32_t dst_x, uint32_t dst_y, uint32_t width, uint32_t height,
                              bool set_mask, bool check_mask)
{
  uint8_t rec[32] = {};
  psx_copy_vram* c = reinterpret_cast<psx_copy_vram*>(rec);
  c->h.size = sizeof(rec);
  c->h.type = PSX_CMD_COPY_VRAM;
  c->src_x = static_cast<uint16_t>(src_x);
  c->src_y = static_cast<uint16_t>(src_y);
  c->dst_x = static_cast<uint16_t>(dst_x);
  c->dst_y = static_cast<uint16_t>(dst_y);
  c->width = static_cast<uint16_t>(width);
  c->height = static_cast<uint16_t>(height);
  c->set_mask_while_drawing = set_mask;
  c->check_mask_before_draw = check_mask;
  m_counters.num_copies++;
  submit(rec, sizeof(rec));
}

void CudaGpuBackend::DrawPolygon(const psx_draw_polygon* cmd)
{
  m_counters.num_vertices += cmd->d.num_vertices;
  m_counters.num_primitives++;
  submit(cmd, cmd->d.h.size);
}
void CudaGpuBackend::DrawPrecisePolygon(const psx_draw_precise_polygon* cmd)
{
  m_counters.num_vertices += cmd->d.num_vertices;
  m_counters.num_primitives++;
  submit(cmd, cmd->d.h.size);
}
void CudaGpuBackend::DrawSprite(const psx_draw_rectangle* cmd)
{
  m_counters.num_vertices++;
  m_counters.num_primitives++;
  submit(cmd, cmd->d.h.size);
}
void CudaGpuBackend::DrawLine(const psx_draw_line* cmd)
{
  m_counters.num_vertices += cmd->d.num_vertices;
  m_counters.num_primitives += cmd->d.num_vertices / 2;
  submit(cmd, cmd->d.h.size);
}
void CudaGpuBackend::DrawPreciseLine(const psx_draw_header* cmd)
{
  m_counters.num_vertices += cmd->num_vertices;
  m_counters.num_primitives += cmd->num_vertices / 2;
  submit(cmd, cmd->h.size);
}

void CudaGpuBackend::SetDrawingArea(const psx_drawing_area& area)
{
  uint8_t rec[32] = {};
  psx_set_drawing_area* c = reinterpret_cast<psx_set_drawing_area*>(rec);
  c->h.size = sizeof(rec);
  c->h.type = PSX_CMD_SET_DRAWING_AREA;
  c->new_area = area;
  submit(rec, sizeof(rec));
}

void CudaGpuBackend::UpdateCLUT(uint16_t reg, bool clut_is_8bit)
{
  uint8_t rec[16] = {};
  psx_update_clut* c = reinterpret_cast<psx_update_clut*>(rec);
  c->h.size = sizeof(rec);
  c->h.type = PSX_CMD_UPDATE_CLUT;
  c->reg = reg;
  c->clut_is_8bit = clut_is_8bit;
  submit(rec, sizeof(rec));
}

void CudaGpuBackend::ClearVRAM()
{
  if (m_ctx && psxgpu_clear_vram(m_ctx, m_instance) < 0)
    fatal("psxgpu_clear_vram");
}

void CudaGpuBackend::UpdateDisplay(const psx_update_display* cmd)
{
  submit(cmd, cmd->h.size);
}

void CudaGpuBackend::LoadState(const psx_load_state* cmd)
{
  if (m_ctx && psxgpu_load_state(m_ctx, m_instance, cmd->vram_data, cmd->clut_data) < 0)
    fatal("psxgpu_load_state");
}

bool CudaGpuBackend::SaveMemoryState(void* dst, size_t size)
{
  if (!m_ctx || size < MEMORY_STATE_SIZE)
    return false;
  uint16_t* p = static_cast<uint16_t*>(dst);
  return psxgpu_save_state(m_ctx, m_instance, p, p + PSX_VRAM_PIXELS) == PSXGPU_OK;
}

bool CudaGpuBackend::LoadMemoryState(const void* src, size_t size)
{
  if (!m_ctx || size < MEMORY_STATE_SIZE)
    return false;
  const uint16_t* p = static_cast<const uint16_t*>(src);
  return psxgpu_load_state(m_ctx, m_instance, p, p + PSX_VRAM_PIXELS) == PSXGPU_OK;
}

uint16_t* CudaGpuBackend::host_vram()
{
  static uint16_t s_unattached[PSX_VRAM_PIXELS];
  return m_ctx ? psxgpu_shadow_vram(m_ctx, m_instance) : s_unattached;
}

const void* CudaGpuBackend::display(uint32_t* width, uint32_t* height, uint32_t* stride, bool* rgba8) const
{
  uint32_t is8 = 1;
  const void* p = m_ctx ? psxgpu_display(m_ctx, width, height, stride, &is8) : nullptr;
  if (rgba8)
    *rgba8 = is8 != 0;
  return p;
}

} // namespace psx

// ---- C shim so the class can be driven from tests without a C++ harness ----
extern "C" {
void* psxbackend_create(int device, int upload_vram)
{
  auto* b = new psx::CudaGpuBackend();
  std::string err;
  if (!b->Initialize(upload_vram != 0, &err, device))
  {
    std::fprintf(stderr, "psxbackend_create: %s\n", err.c_str());
    delete b;
    return nullptr;
  }
  return b;
}
void psxbackend_destroy(void* b)
{
  delete static_cast<psx::CudaGpuBackend*>(b);
}
// Feeds a wire stream record by record through CudaGpuBackend::HandleCommand.
long psxbackend_handle_stream(void* b, const void* wire, size_t bytes)
{
  auto* be = static_cast<psx::CudaGpuBackend*>(b);
  const uint8_t* p = static_cast<const uint8_t*>(wire);
  const uint8_t* end = p + bytes;
  long n = 0;
  while (p < end)
  {
    const psx_cmd_header* h = reinterpret_cast<const psx_cmd_header*>(p);
    if (h->size < sizeof(psx_cmd_header) || p + h->size > end)
      return -1;
    be->HandleCommand(h);
    p += h->size;
    n++;
  }
  be->FlushRender();
  return n;
}
// ReadVRAM of the whole surface, then returns the host shadow (the reference's g_vram contract).
const uint16_t* psxbackend_read_all(void* b)
{
  auto* be = static_cast<psx::CudaGpuBackend*>(b);
  be->ReadVRAM(0, 0, PSX_VRAM_WIDTH, PSX_VRAM_HEIGHT);
  return be->host_vram();
}
const void* psxbackend_display(void* b, uint32_t* w, uint32_t* h, uint32_t* stride)
{
  bool rgba8;
  return static_cast<psx::CudaGpuBackend*>(b)->display(w, h, stride, &rgba8);
}
void psxbackend_counters(void* b, uint32_t* out5)
{
  const auto& c = static_cast<psx::CudaGpuBackend*>(b)->counters();
  out5[0] = c.num_reads;
  out5[1] = c.num_writes;
  out5[2] = c.num_copies;
  out5[3] = c.num_vertices;
  out5[4] = c.num_primitives;
}
}
