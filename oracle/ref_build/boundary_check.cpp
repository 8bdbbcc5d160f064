// boundary_check.cpp — TEST INFRASTRUCTURE: the subclass of INTEGRATION.md §2, compiled against the reference's REAL
// headers (src/core/gpu_backend.h and what it pulls in), so that a drift in any virtual's signature, in a command
// struct, or in the C ABI this repository exports breaks `make boundary_check` (tests/test_abi.py runs it).
// Compile only (-c): a DuckStation tree would link this file with libpsxgpu.so.
//
// ONE-LINE REFERENCE PATCH this backend needs (top): `HandleCommand` is not virtual
// (gpu_backend.h:91) and its UpdateCLUT case (gpu_backend.cpp:490-495) fills the host-side cache without calling any
// virtual.  A device backend must see palette loads in stream order, so the case needs a hook:
//     case VideoThreadCommandType::UpdateCLUT: { ...; GPU_SW_Rasterizer::UpdateCLUT(...); OnCLUTUpdated(ccmd->reg, ccmd->clut_is_8bit); }
// with `virtual void OnCLUTUpdated(GPUTexturePaletteReg, bool) {}` in GPUBackend.  It is declared below as the method
// the patched base class would call; everything else overrides virtuals that exist today.
#include "core/gpu_backend.h"
#include "core/gpu_sw_rasterizer.h"
#include "core/settings.h"
#include "core/system.h"
#include "core/system_private.h"
#include "core/video_presenter.h"
#include "core/video_thread_commands.h"

#include "util/gpu_device.h"
#include "util/gpu_texture.h"
#include "util/state_wrapper.h"

#include "common/error.h"

#include "../../duckstation_b200/csrc/backend.h"
#include "../../include/psxgpu.h"
#include "../../include/psxgpu_wire.h"

#include <cstring>
#include <memory>
#include <vector>

// the wire mirrors are the reference's structs, byte for byte (the full list of checks: ref_harness.cpp:47-98)
static_assert(sizeof(psx_draw_rectangle) == sizeof(GPUBackendDrawRectangleCommand));
static_assert(sizeof(psx_update_display) == sizeof(GPUBackendUpdateDisplayCommand));
static_assert(sizeof(psx_fill_vram) == sizeof(GPUBackendFillVRAMCommand));
static_assert(sizeof(psx_copy_vram) == sizeof(GPUBackendCopyVRAMCommand));
static_assert(sizeof(psx_set_drawing_area) == sizeof(GPUBackendSetDrawingAreaCommand));
static_assert(sizeof(psx_update_clut) == sizeof(GPUBackendUpdateCLUTCommand));
static_assert(offsetof(psx_load_state, clut_data) == offsetof(GPUBackendLoadStateCommand, clut_data));

class GPU_CUDA final : public GPUBackend
{
public:
  GPU_CUDA() = default;
  ~GPU_CUDA() override
  {
    if (m_presenter)
      psxgpu_presenter_destroy(m_presenter);
    if (m_ctx)
      psxgpu_destroy(m_ctx);
  }

  bool Initialize(bool upload_vram, Error* error) override
  {
    if (!GPUBackend::Initialize(upload_vram, error))
      return false;
    return CreateContext(upload_vram, error);
  }

  // gpu_sw.cpp has no UpdateSettings of its own; the two settings this backend bakes into its context (the
  // Modulate5Bit crop and show-VRAM) need the context's options changed, which keeps VRAM: save, recreate, load.
  bool UpdateSettings(const GPUSettings& old_settings, Error* error) override
  {
    if (!GPUBackend::UpdateSettings(old_settings, error))
      return false;
    if (g_gpu_settings.gpu_modulation_crop != old_settings.gpu_modulation_crop ||
        g_gpu_settings.gpu_show_vram != old_settings.gpu_show_vram)
    {
      std::vector<u16> vram(VRAM_WIDTH * VRAM_HEIGHT), clut(GPU_CLUT_SIZE);
      if (psxgpu_save_state(m_ctx, 0, vram.data(), clut.data()) != PSXGPU_OK)
        return false;
      psxgpu_destroy(m_ctx);
      m_ctx = nullptr;
      if (!CreateContext(false, error))
        return false;
      psxgpu_load_state(m_ctx, 0, vram.data(), clut.data());
    }
    if (g_gpu_settings.display_deinterlacing_mode != old_settings.display_deinterlacing_mode ||
        g_gpu_settings.display_24bit_chroma_smoothing != old_settings.display_24bit_chroma_smoothing)
    {
      if (m_presenter)
        psxgpu_presenter_destroy(m_presenter);
      m_presenter = nullptr; // recreated by the next UpdateDisplay
    }
    return true;
  }

  u32 GetResolutionScale() const override { return 1; }
  void RestoreDeviceContext() override {}
  void FlushRender() override { psxgpu_flush(m_ctx); }

protected:
  // every draw / transfer virtual forwards its record or rebuilds it unchanged
  void DrawPolygon(const GPUBackendDrawPolygonCommand* cmd) override { Submit(cmd); }
  void DrawPrecisePolygon(const GPUBackendDrawPrecisePolygonCommand* cmd) override { Submit(cmd); }
  void DrawSprite(const GPUBackendDrawRectangleCommand* cmd) override { Submit(cmd); }
  void DrawLine(const GPUBackendDrawLineCommand* cmd) override { Submit(cmd); }
  void DrawPreciseLine(const GPUBackendDrawPreciseLineCommand* cmd) override { Submit(cmd); }
  void FillVRAM(u32 x, u32 y, u32 width, u32 height, u32 color, bool interlaced_rendering, u8 interlaced_display_field) override
  {
    m_builder.FillVRAM(x, y, width, height, color, interlaced_rendering, interlaced_display_field);
  }
  void UpdateVRAM(u32 x, u32 y, u32 width, u32 height, const void* data, bool set_mask, bool check_mask) override
  {
    m_builder.UpdateVRAM(x, y, width, height, data, set_mask, check_mask);
  }
  void CopyVRAM(u32 src_x, u32 src_y, u32 dst_x, u32 dst_y, u32 width, u32 height, bool set_mask, bool check_mask) override
  {
    m_builder.CopyVRAM(src_x, src_y, dst_x, dst_y, width, height, set_mask, check_mask);
  }
  void ClearVRAM() override { psxgpu_clear_vram(m_ctx, 0); }
  void LoadState(const GPUBackendLoadStateCommand* cmd) override { psxgpu_load_state(m_ctx, 0, cmd->vram_data, cmd->clut_data); }

  // g_vram is a host shadow here: make the rectangle current (what GPU_HW::DownloadVRAMFromGPU does, gpu_hw.cpp:3472-3515)
  void ReadVRAM(u32 x, u32 y, u32 width, u32 height) override
  {
    std::vector<u16> tmp(static_cast<size_t>(width) * height);
    if (psxgpu_read_vram(m_ctx, 0, x, y, width, height, tmp.data(), width) != PSXGPU_OK)
      return;
    for (u32 r = 0; r < height; r++)
      for (u32 c = 0; c < width; c++)
        g_vram[((y + r) % VRAM_HEIGHT) * VRAM_WIDTH + ((x + c) % VRAM_WIDTH)] = tmp[static_cast<size_t>(r) * width + c];
  }

  // SetDrawingArea is a case of the base class switch (gpu_backend.cpp:481-488) that stores the area in a host global
  // and then calls this virtual: forward what it stored.
  void DrawingAreaChanged() override
  {
    psx_set_drawing_area rec{};
    rec.h.size = sizeof(rec);
    rec.h.type = PSX_CMD_SET_DRAWING_AREA;
    static_assert(sizeof(rec.new_area) == sizeof(GPU_SW_Rasterizer::g_drawing_area));
    std::memcpy(&rec.new_area, &GPU_SW_Rasterizer::g_drawing_area, sizeof(rec.new_area));
    psxgpu_submit_wire(m_ctx, 0, &rec, sizeof(rec));
  }
  // the hook of the one-line patch described at the top of this file
  void OnCLUTUpdated(GPUTexturePaletteReg reg, bool clut_is_8bit) { m_builder.UpdateCLUT(reg.bits, clut_is_8bit); }

  void ClearCache() override {}
  void OnBufferSwapped() override {}
#ifdef BOUNDARY_DRIFT
  // negative control of tests/test_abi.py: a signature the base class does not have must not compile
  void DrawSprite(const GPUBackendDrawRectangleCommand* cmd, u32 extra) override { Submit(cmd); }
#endif

  void UpdateDisplay(const GPUBackendUpdateDisplayCommand* cmd) override
  {
    // scan-out (+ chroma smoothing + deinterlacing) on the device, one frame back: psxgpu_present
    if (!m_presenter &&
        psxgpu_presenter_create(m_ctx, static_cast<u32>(g_gpu_settings.display_deinterlacing_mode),
                                g_gpu_settings.display_24bit_chroma_smoothing ? 1u : 0u, &m_presenter) != PSXGPU_OK)
    {
      return;
    }
    m_frame.resize(static_cast<size_t>(4096) * 1024);
    u32 width = 0, height = 0;
    const int rc = psxgpu_present(m_presenter, 0, reinterpret_cast<const psx_update_display*>(cmd), m_frame.data(),
                                  m_frame.size() * sizeof(u32), &width, &height);
    if (rc != PSXGPU_OK || width == 0 || height == 0)
    {
      VideoPresenter::ClearDisplayTexture();
      return;
    }
    if (!m_texture || m_texture->GetWidth() != width || m_texture->GetHeight() != height)
    {
      m_texture = g_gpu_device->FetchTexture(width, height, 1, 1, 1, GPUTexture::Type::Texture, GPUTextureFormat::RGBA8,
                                             GPUTexture::Flags::None);
    }
    if (m_texture && m_texture->Update(0, 0, width, height, m_frame.data(), width * sizeof(u32)))
      VideoPresenter::SetDisplayTexture(m_texture.get(), GSVector4i(0, 0, static_cast<s32>(width), static_cast<s32>(height)));
  }

  bool AllocateMemorySaveState(System::MemorySaveState& mss, Error* error) override
  {
    mss.gpu_state_data.resize(VRAM_SIZE + sizeof(g_gpu_clut));
    return true;
  }
  void DoMemoryState(StateWrapper& sw, System::MemorySaveState& mss) override
  {
    u16* p = reinterpret_cast<u16*>(mss.gpu_state_data.data());
    if (sw.IsReading())
      psxgpu_load_state(m_ctx, 0, p, p + VRAM_WIDTH * VRAM_HEIGHT);
    else
      psxgpu_save_state(m_ctx, 0, p, p + VRAM_WIDTH * VRAM_HEIGHT);
  }

private:
  bool CreateContext(bool upload_vram, Error* error)
  {
    psxgpu_opts o{};
    o.struct_size = sizeof(o);
    o.modulation_crop = g_gpu_settings.gpu_modulation_crop;
    o.show_vram = g_gpu_settings.gpu_show_vram;
    o.display_format = PSXGPU_FMT_RGBA8;
    if (psxgpu_create(0, 1, &o, &m_ctx) != PSXGPU_OK)
    {
      // video_thread.cpp:857-870 then falls back to GPU_SW exactly as it does when a hardware backend fails
      Error::SetStringView(error, "No CUDA device: the CUDA rasteriser has no CPU fallback.");
      return false;
    }
    std::string attach_error;
    if (!m_builder.Attach(m_ctx, 0, &attach_error))
    {
      Error::SetStringView(error, attach_error);
      return false;
    }
    if (upload_vram)
      psxgpu_load_state(m_ctx, 0, g_vram, g_gpu_clut);
    return true;
  }

  template<typename T>
  void Submit(const T* cmd)
  {
    psxgpu_submit_wire(m_ctx, 0, cmd, cmd->size);
  }

  psxgpu_ctx* m_ctx = nullptr;
  psxgpu_presenter* m_presenter = nullptr;
  psx::CudaGpuBackend m_builder; // csrc/backend.h: builds the records of the scalar-argument virtuals
  std::unique_ptr<GPUTexture> m_texture;
  std::vector<u32> m_frame;
};

// the factory a DuckStation tree would add next to GPUBackend::CreateSoftwareBackend (gpu_sw.cpp:450-453)
Common::unique_aligned_ptr<GPUBackend> CreateCUDABackend()
{
  return Common::make_unique_aligned<GPU_CUDA>(HOST_CACHE_LINE_SIZE);
}
