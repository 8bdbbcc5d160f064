// kernels.h — host-callable launchers for kernels.cu.
#pragma once

#include "psx_types.h"

#include <cuda_runtime.h>

namespace psx {

struct ScanoutParams
{
  uint32_t src_x, src_y, skip_x;
  uint32_t width, height;
  uint32_t line_skip;
  uint32_t is24;
  uint32_t fast;       // the reference's non-wrapping fast path applies (gpu_sw.cpp:241 / :305)
  uint32_t format;     // 0 RGBA8, 1 RGB5A1, 2 A1BGR5, 3 RGB565 (15-bit source only)
  uint32_t out_stride; // bytes
};

// -DPSX_DEBUG_BOUNDS builds: extents of the chunk whose kernels come next on the stream (kernels.cu, PSX_CHECK).
struct DebugLimits
{
  uint32_t n_cmds, n_payload, n_clutops, n_items, n_instances, clut_slots;
};
cudaError_t launch_debug_limits(const DebugLimits& lim, cudaStream_t st);

// tile_cmds: (n + 31) / 32 groups x NUM_TILES words, the transposed tile masks the raster kernels bin from.
cudaError_t launch_setup(const PackedCmd* cmds, PrimSetup* setups, BinInfo* bins, uint32_t* tile_cmds, uint32_t n, cudaStream_t st);
inline size_t tile_cmds_bytes(size_t n_cmds)
{
  return ((n_cmds + 31) / 32 + 1) * NUM_TILES * sizeof(uint32_t);
}
// After setup: WaveItem::pad[0] = union of the tile masks of the item's commands (raster CTAs of other tiles exit at once).
cudaError_t launch_item_union(WaveItem* items, uint32_t n_items, const BinInfo* bins, cudaStream_t st);
cudaError_t launch_clut(const WaveItem* items, uint32_t n_items, const ClutOp* ops, const uint16_t* vram_pool,
                        uint16_t* clut_pool, uint32_t clut_slots, cudaStream_t st);
cudaError_t launch_raster(const WaveItem* items, uint32_t n_items, const PrimSetup* setups, const uint32_t* tile_cmds,
                          uint16_t* vram_pool, const uint16_t* clut_pool, const uint16_t* payload, uint32_t clut_slots,
                          cudaStream_t st, uint32_t tiled_cmds, unsigned int* loop_counter);
// tiled_cmds: commands of the wave's tiled epochs; loop_counter: one device word owned by the caller's context / stream
// (work counter of raster_loop_kernel; nullptr = one CTA per tile)
cudaError_t launch_serial(const WaveItem* items, uint32_t n_items, const PrimSetup* setups, uint16_t* vram_pool,
                          const uint16_t* clut_pool, const uint16_t* payload, uint32_t clut_slots, cudaStream_t st);
// Persistent (latency) mode: whole flush of up to persistent_max_instances() instances in one cooperative launch.
// `waves`: n_waves records of {item_begin, item_count, has_clut, pad} (uint32 each) in device memory.
int persistent_max_instances(int device);
cudaError_t launch_persistent(const WaveItem* items, const void* waves, uint32_t n_waves, const uint32_t* instances, uint32_t n_instances,
                              const PrimSetup* setups, const uint32_t* tile_cmds, const ClutOp* clut_ops, uint16_t* vram_pool,
                              uint16_t* clut_pool, const uint16_t* payload, uint32_t clut_slots, unsigned int* barrier_counter,
                              cudaStream_t st);
// Instance mode (many short epochs per instance): one CTA per instance for a whole flush, __syncthreads() between
// epochs.  `work` / `epochs`: device arrays (psx_types.h).  wide = 32 warps per CTA (a handful of instances), else 8.
// `tile_sets`: scratch, sixteen words per epoch (filled here by epoch_tiles_kernel: the tiles each epoch writes / reads).
cudaError_t launch_instances(const InstWork* work, uint32_t n_instances, const Epoch* epochs, uint32_t* tile_sets, const PrimSetup* setups,
                             const BinInfo* bins, const uint32_t* tile_cmds, const ClutOp* clut_ops, uint16_t* vram_pool, uint16_t* clut_pool,
                             const uint16_t* payload, uint32_t clut_slots, bool wide, cudaStream_t st, unsigned int* inst_counter);
cudaError_t launch_hash(const uint16_t* vram_pool, uint32_t first, uint32_t count, unsigned long long* out, cudaStream_t st);
cudaError_t launch_scanout(const uint16_t* vram, const ScanoutParams& p, uint8_t* out, cudaStream_t st);

// presenter.cu: the presentation steps after scan-out, on tightly packed RGBA8 frames in device memory.
struct PresentField
{
  const uint32_t* pixels; // nullptr: samples as zero (a field buffer that was never written / a cleared display)
  uint32_t width, height;
};
cudaError_t launch_weave(const PresentField& src, uint32_t field, uint32_t* target, uint32_t width, uint32_t full_height, cudaStream_t st);
cudaError_t launch_blend(const PresentField& cur, const PresentField& prev, uint32_t* target, uint32_t width, uint32_t height, cudaStream_t st);
// f[0] = this field, f[1..3] = one, two and three fields back
cudaError_t launch_adaptive(const PresentField f[4], uint32_t field, uint32_t* target, uint32_t width, uint32_t full_height, cudaStream_t st);
cudaError_t launch_chroma_smoothing(const PresentField& src, uint32_t* out, cudaStream_t st);

} // namespace psx
