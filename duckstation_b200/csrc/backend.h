// backend.h — C++ host-side mirror of the reference's backend interface for the software-renderer path.
//
// `psx::CudaGpuBackend` has the method set of `GPUBackend` that `GPU_SW` implements
// (reference: src/core/gpu_backend.h:75-171, src/core/gpu_sw.h:17-69) with the same names, argument
// meaning and error behaviour (bool + message for Initialize, void draw calls that cannot fail, a fatal
// error hook instead of exceptions).  Command structs are the wire mirrors of include/psxgpu_wire.h,
// which are layout-identical to the reference's (src/core/video_thread_commands.h), so in a DuckStation
// tree the subclass in INTEGRATION.md forwards its arguments with reinterpret_casts only.
//
// Differences forced by the device boundary:
//   * g_vram is a host shadow: ReadVRAM() makes the requested rectangle current (psxgpu_read_vram);
//   * SetDrawingArea / UpdateCLUT are handled by the base class in the reference (gpu_backend.cpp:481-495)
//     but must reach the device in order, so HandleCommand() is restated here and owns all record types.
#pragma once

#include "../../include/psxgpu.h"

#include <cstdint>
#include <string>
#include <vector>

namespace psx {

class CudaGpuBackend
{
public:
  struct Counters // GPUBackend::Counters, gpu_backend.h:123-131
  {
    uint32_t num_reads = 0, num_writes = 0, num_copies = 0, num_vertices = 0, num_primitives = 0;
  };

  CudaGpuBackend();
  ~CudaGpuBackend();
  CudaGpuBackend(const CudaGpuBackend&) = delete;
  CudaGpuBackend& operator=(const CudaGpuBackend&) = delete;

  // GPUBackend::Initialize(bool upload_vram, Error*) (gpu_sw.cpp:34-59).  upload_vram: take the initial contents
  // from host_vram() instead of clearing.  On failure returns false and fills *error.
  bool Initialize(bool upload_vram, std::string* error, int device = 0, const psxgpu_opts* opts = nullptr);
  // Attach to instance `instance` of an existing (batched) context instead of owning one.
  bool Attach(psxgpu_ctx* shared, uint32_t instance, std::string* error);

  uint32_t GetResolutionScale() const { return 1u; } // gpu_sw.cpp:29-32
  void RestoreDeviceContext() {}                     // gpu_sw.cpp:209-211
  void FlushRender();                                // gpu_backend.h:150

  /// Main command handler (GPUBackend::HandleCommand, gpu_backend.cpp:383-544).
  void HandleCommand(const psx_cmd_header* cmd);

  void ReadVRAM(uint32_t x, uint32_t y, uint32_t width, uint32_t height);
  void FillVRAM(uint32_t x, uint32_t y, uint32_t width, uint32_t height, uint32_t color, bool interlaced_rendering,
                uint8_t active_line_lsb);
  void UpdateVRAM(uint32_t x, uint32_t y, uint32_t width, uint32_t height, const void* data, bool set_mask, bool check_mask);
  void CopyVRAM(uint32_t src_x, uint32_t src_y, uint32_t dst_x, uint32_t dst_y, uint32_t width, uint32_t height, bool set_mask,
                bool check_mask);
  void DrawPolygon(const psx_draw_polygon* cmd);
  void DrawPrecisePolygon(const psx_draw_precise_polygon* cmd);
  void DrawSprite(const psx_draw_rectangle* cmd);
  void DrawLine(const psx_draw_line* cmd);
  void DrawPreciseLine(const psx_draw_header* cmd);
  void SetDrawingArea(const psx_drawing_area& area); // base-class case in the reference (gpu_backend.cpp:481-488)
  void UpdateCLUT(uint16_t reg, bool clut_is_8bit);   // base-class case in the reference (gpu_backend.cpp:490-495)
  void DrawingAreaChanged() {}
  void ClearCache() {}
  void OnBufferSwapped() {}
  void ClearVRAM();
  void UpdateDisplay(const psx_update_display* cmd);
  void LoadState(const psx_load_state* cmd);
  // AllocateMemorySaveState / DoMemoryState (gpu_sw.cpp:73-84): VRAM (1 MiB) followed by the CLUT cache (512 B).
  static constexpr size_t MEMORY_STATE_SIZE = PSX_VRAM_BYTES + PSX_CLUT_ENTRIES * 2;
  bool SaveMemoryState(void* dst, size_t size);
  bool LoadMemoryState(const void* src, size_t size);

  /// Stand-in for the reference's global g_vram (gpu.h:141): current after ReadVRAM / SaveMemoryState.
  uint16_t* host_vram();
  /// Last scanned-out frame (what GPU_SW hands to VideoPresenter::SetDisplayTexture).
  const void* display(uint32_t* width, uint32_t* height, uint32_t* stride, bool* rgba8) const;

  const Counters& counters() const { return m_counters; }
  psxgpu_ctx* context() const { return m_ctx; }
  uint32_t instance() const { return m_instance; }
  const std::string& last_error() const { return m_error; }
  // Called on unrecoverable device errors (the reference calls VideoThread::ReportFatalErrorAndShutdown).
  void (*fatal_error_hook)(const char* message) = nullptr;

private:
  void submit(const void* rec, size_t bytes);
  void fatal(const char* what);

  psxgpu_ctx* m_ctx = nullptr;
  uint32_t m_instance = 0;
  bool m_owns = false;
  Counters m_counters;
  std::string m_error;
  std::vector<uint8_t> m_scratch;
};

} // namespace psx
