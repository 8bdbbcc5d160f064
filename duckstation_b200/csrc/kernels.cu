// kernels.cu — sm_100a kernels of the PSX rasteriser.
//
//   setup_kernel    one thread per primitive: PackedCmd -> PrimSetup + BinInfo (all divisions happen here); one warp =
//                   32 consecutive commands also leaves the transpose of their tile masks (tile_cmds) for binning
//   clut_kernel     CLUT snapshot pass of an epoch wave (reference: UpdateCLUT, gpu_sw_rasterizer.cpp:43-66)
//   raster_kernel / raster_loop_kernel
//                   one warp (= one CTA of 32 threads) per (instance, 64x32 VRAM tile): lists the epoch's
//                   primitives for its tile in submission order from tile_cmds (one word per 32 commands), stages
//                   the tile in shared memory and walks the list in order; no block-level barrier exists.  Per
//                   primitive the 32 lanes compute the 32 row spans in one pass, a warp prefix sum compacts the 32-bit
//                   tile words the spans touch and they are shaded 32 words at a time — two pixels per lane, packed
//                   16x2 arithmetic (psx_pair.h) — so small primitives keep the lanes busy.  Tile rows move as bulk
//                   asynchronous copies counted in by a transaction barrier, the next primitive's record is
//                   prefetched into registers, palettes are staged in shared memory by cp.async.  Texels come from
//                   global memory through the read-only path.  The loop variant (waves of 16+ instances) is
//                   SMs x 28 persistent CTAs (72 registers, 60 KB of L1 left) pulling tiles from a counter.
//   raster_persistent_kernel
//                   latency mode (up to 4 instances): one cooperative launch for a whole flush, grid barrier
//                   between epochs, four warps per tile
//   raster_instance_kernel / raster_multi_kernel / raster_cluster_kernel
//                   streams cut into many short epochs, one launch for all of them: one CTA per instance with
//                   __syncthreads between epochs; several instances per CTA, each behind its own epoch counter
//                   (slot mode: the batched case); one thread-block cluster per instance (a handful of instances)
//   serial_kernel   epochs of a serial kind (self-sampling primitive, aliasing copy, mask-checked upload tail):
//                   one CTA each, in the reference's own order.
//   hash_kernel     vramhash64 per instance
//   scanout_kernel  15-bit / 24-bit display read-out (reference: CopyOut15Bit / CopyOut24Bit, gpu_sw.cpp:230-356)
//
// Integer work only; bound by memory traffic and issue rate, so no tensor-core path.
#include "kernels.h"
#include "psx_pair.h"
#include "psx_pixel.h"
#include "psx_setup.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace psx {

// ------------------------------------------------------------------------------------------------ debug bounds build
// compute-sanitizer is not available on the GPU pool, so a -DPSX_DEBUG_BOUNDS build carries its own checks: every index
// into a bin list, the transposed tile bitmaps, the setup records, the upload payload, the CLUT slots, the wave items
// and the shared tile, plus the grid barrier's generation, traps (`__trap()` -> the launch fails with an error the C
// ABI reports) instead of reading or writing out of range.  The extents of the chunk's buffers reach the kernels
// through g_dbg, written on the context's stream ahead of the chunk's launches (one context at a time in such a
// build).  A normal build compiles all of it away.  profiles/r02_debug_bounds.log is the GPU suite run under it.
#ifdef PSX_DEBUG_BOUNDS
__device__ DebugLimits g_dbg;
#define PSX_CHECK(cond)                                                                                                \
  do                                                                                                                   \
  {                                                                                                                    \
    if (!(cond))                                                                                                       \
    {                                                                                                                  \
      printf("PSX_DEBUG_BOUNDS: %s failed at %s:%d (block %u thread %u)\n", #cond, __FILE__, __LINE__, blockIdx.x, threadIdx.x); \
      __trap();                                                                                                        \
    }                                                                                                                  \
  } while (0)
cudaError_t launch_debug_limits(const DebugLimits& lim, cudaStream_t st)
{
  return cudaMemcpyToSymbolAsync(g_dbg, &lim, sizeof(lim), 0, cudaMemcpyHostToDevice, st);
}
#else
#define PSX_CHECK(cond) ((void)0)
cudaError_t launch_debug_limits(const DebugLimits&, cudaStream_t)
{
  return cudaSuccess;
}
#endif

// ------------------------------------------------------------------------------------------------ setup
#ifndef PSX_SETUP_COALESCE
#define PSX_SETUP_COALESCE 1 // (setup_kernel per 4096 instances: 1.108 -> 0.980 ms)
#endif
// Besides the PrimSetup records the kernel leaves two views of "which tiles can this command touch" (bounding box):
// BinInfo per command (column bits | row bits; item_union_kernel ORs them per epoch) and its transpose for the raster
// warps, tile_cmds[group * NUM_TILES + tile] = bit c set if command group * 32 + c touches the tile — a tile then
// finds its commands of an epoch with one 32-bit load per 32 commands instead of scanning every command's mask.
// A warp of this kernel is exactly one group of 32 commands, so the transpose is 32 ballots.
__global__ void __launch_bounds__(128) setup_kernel(const PackedCmd* __restrict__ cmds, PrimSetup* __restrict__ setups,
                                                    BinInfo* __restrict__ bins, uint32_t* __restrict__ tile_cmds, uint32_t n)
{
  constexpr uint32_t SV = sizeof(PrimSetup) / 16; // 16-byte chunks per record
#if PSX_SETUP_COALESCE
  __shared__ uint4 setup_stage[4][32 * SV]; // per warp: 32 records of 256 bytes
#endif
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t mask = 0, reject = 0, n_reject = 0;
  if (i < n)
  {
    // 64-byte record as four 128-bit loads
    PackedCmd c;
    const uint4* src = reinterpret_cast<const uint4*>(cmds + i);
    uint4* cdst = reinterpret_cast<uint4*>(&c);
#pragma unroll
    for (int k = 0; k < 4; k++)
      cdst[k] = __ldg(src + k);
    PrimSetup s;
    uint4* sw = reinterpret_cast<uint4*>(&s);
#pragma unroll
    for (int k = 0; k < static_cast<int>(sizeof(PrimSetup) / 16); k++)
      sw[k] = make_uint4(0, 0, 0, 0);
    BinInfo b;
    setup_prim(c, s, b);
#if PSX_SETUP_COALESCE
    // the record goes to the warp's staging area first (16-byte chunk k of lane L at L * 16 + (k ^ (L & 15)): four lanes per
    // group of four banks, the minimum for a 128-bit store) and leaves it below as 512 contiguous bytes per instruction;
    // written straight from here, every store instruction touched 32 different cache lines
    {
      uint4* const st = setup_stage[threadIdx.x >> 5];
#pragma unroll
      for (uint32_t k = 0; k < SV; k++)
        st[lane * SV + (k ^ (lane & 15u))] = sw[k];
    }
#else
    uint4* dst = reinterpret_cast<uint4*>(setups + i);
#pragma unroll
    for (int k = 0; k < static_cast<int>(sizeof(PrimSetup) / 16); k++)
      dst[k] = sw[k];
#endif
    b.readmask = (s.op == OP_NOP) ? 0u : read_tilemask(c);
    bins[i] = b;
    mask = b.tilemask;
    // Exact binning for triangles: the mask above is the bounding box, and a triangle that spans 2 x 2 tiles or more
    // usually misses a corner tile (13 % of the raster warps' visits found no pixel).  Per band of 32 rows the spans lie
    // between the two edges evaluated at the band's first and last row of each part (the edges are linear inside a
    // part, the same 32.32 values the raster warps evaluate), cut to the drawing area: tile columns outside that hull
    // cannot hold a pixel.  Up to four such tiles per command are remembered and their bits cleared below.
    if (s.op == OP_TRIANGLE && !(s.desc.path & LP_COMPLEX))
    {
      const uint32_t cols = mask & 0xFFFFu, rows = mask >> 16;
      const auto& T = s.tri;
      // (only where VRAM row == triangle row: a drawing area that reaches past row 511 makes the rows wrap)
      const bool rows_plain = imin(T.lo[0], T.lo[1]) >= 0 && imax(T.hi[0], T.hi[1]) <= static_cast<int32_t>(VRAM_H);
      if ((cols & (cols - 1u)) != 0 && (rows & (rows - 1u)) != 0 && rows_plain)
      {
        // (the set bits only: a bounding box spans two or three tile rows and columns, not sixteen)
        for (uint32_t rbits = rows; rbits != 0 && n_reject < 4u; rbits &= rbits - 1u)
        {
          const uint32_t ty = static_cast<uint32_t>(__ffs(static_cast<int>(rbits))) - 1u;
          int32_t h_lo = 0x7FFFFFFF, h_hi = static_cast<int32_t>(0x80000000u);
#pragma unroll
          for (int p = 0; p < 2; p++)
          {
            const int32_t ra = imax(T.lo[p], static_cast<int32_t>(ty * TILE_H)), rb = imin(T.hi[p], static_cast<int32_t>(ty * TILE_H + TILE_H)) - 1;
            if (ra > rb)
              continue;
            const TriPart& P = T.part[p];
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
              const uint64_t k = static_cast<uint64_t>(static_cast<uint32_t>((q ? rb : ra) - T.lo[p]));
              h_lo = imin(h_lo, static_cast<int32_t>((P.l_start + k * P.l_step) >> 32));
              h_hi = imax(h_hi, static_cast<int32_t>((P.r_start + k * P.r_step) >> 32));
            }
          }
          h_lo = imax(h_lo, s.clip_l);
          h_hi = imin(h_hi, s.clip_r + 1);
          for (uint32_t cbits = cols; cbits != 0 && n_reject < 4u; cbits &= cbits - 1u)
          {
            const uint32_t tx = static_cast<uint32_t>(__ffs(static_cast<int>(cbits))) - 1u;
            if (h_hi <= h_lo || h_hi <= static_cast<int32_t>(tx * TILE_W) || h_lo >= static_cast<int32_t>(tx * TILE_W + TILE_W))
            {
              reject |= (ty * TILES_X + tx) << (8u * n_reject);
              n_reject++;
            }
          }
        }
      }
    }
  }
  if ((i & ~31u) >= n)
    return; // the whole warp is past the end
#if PSX_SETUP_COALESCE
  {
    __syncwarp();
    const uint4* const st = setup_stage[threadIdx.x >> 5];
    const uint32_t wbase = i & ~31u;
    uint4* const gdst = reinterpret_cast<uint4*>(setups + wbase);
#pragma unroll
    for (uint32_t t = 0; t < SV; t++)
    {
      const uint32_t g = t * 32u + lane, r = g / SV, c = g % SV; // record of the warp, chunk of the record
      if (wbase + r < n)
        gdst[g] = st[r * SV + (c ^ (r & 15u))];
    }
  }
#endif
  // lane L writes tiles L, L + 32, ...: column L & 15, rows (L >> 4) + 2k
  uint32_t colv = 0;
#pragma unroll
  for (uint32_t x = 0; x < TILES_X; x++)
  {
    const uint32_t v = __ballot_sync(0xFFFFFFFFu, (mask >> x) & 1u);
    if ((lane & 15u) == x)
      colv = v;
  }
  uint32_t* out = tile_cmds + static_cast<size_t>(i >> 5) * NUM_TILES + lane;
#pragma unroll
  for (uint32_t y = 0; y < TILES_Y; y++)
  {
    const uint32_t v = __ballot_sync(0xFFFFFFFFu, (mask >> (16u + y)) & 1u);
    if ((y & 1u) == (lane >> 4))
      out[(y >> 1) * 32u] = v & colv;
  }
  // the corner tiles found empty above: clear the command's bit (after the warp's stores, which __syncwarp orders)
  __syncwarp();
  uint32_t* const group_words = tile_cmds + static_cast<size_t>(i >> 5) * NUM_TILES;
  if (__any_sync(0xFFFFFFFFu, n_reject != 0))
  {
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) // (predicated, not a divergent loop)
      if (k < n_reject)
        atomicAnd(group_words + ((reject >> (8u * k)) & 0xFFu), ~(1u << lane));
  }
}

// ------------------------------------------------------------------------------------------------ CLUT snapshots
__global__ void __launch_bounds__(256) clut_kernel(const WaveItem* __restrict__ items, const ClutOp* __restrict__ ops,
                                                   const uint16_t* __restrict__ vram_pool, uint16_t* __restrict__ clut_pool,
                                                   uint32_t clut_slots)
{
  const WaveItem it = items[blockIdx.x];
  const uint16_t* vram = vram_pool + static_cast<size_t>(it.instance) * VRAM_PIXELS;
  uint16_t* cluts = clut_pool + static_cast<size_t>(it.instance) * clut_slots * CLUT_ENTRIES;
  const uint32_t t = threadIdx.x;
  for (uint32_t o = it.clut_begin; o < it.clut_end; o++)
  {
    PSX_CHECK(o < g_dbg.n_clutops);
    const ClutOp op = ops[o];
    PSX_CHECK(op.dst_slot < clut_slots && op.y < VRAM_H && it.instance < g_dbg.n_instances);
    if (op.is8bit || t < 16)
      cluts[op.dst_slot * CLUT_ENTRIES + t] = vram[op.y * VRAM_W + ((op.x + t) & (VRAM_W - 1))];
  }
}

// ------------------------------------------------------------------------------------------------ raster
constexpr int SERIAL_THREADS = 256;                     // serial_kernel: one CTA per serial epoch
constexpr int SERIAL_WARPS = SERIAL_THREADS / 32;
constexpr uint32_t LIST_CAP = 160;                      // per-tile bin capacity between raster phases (5216 B per warp: 32 warps fit the 196 KB shared-memory configuration, L1 keeps 60 KB)
constexpr uint32_t SETUP_WORDS = sizeof(PrimSetup) / 4; // 60
constexpr uint32_t SETUP_VECS = sizeof(PrimSetup) / 16; // 16
#ifndef PSX_REC_MODE
#define PSX_REC_MODE 2 // (measured, draw wave of 512 instances: 0 = 1.640 ms, 1 = 1.653, 2 = 1.634; with the two options below 1.613)
#endif
#ifndef PSX_OPT_WIN
#define PSX_OPT_WIN 1
#endif
#ifndef PSX_OPT_CLUTLD
#define PSX_OPT_CLUTLD 1
#endif
#ifndef PSX_OPT_CLUTASYNC
#define PSX_OPT_CLUTASYNC 1
#endif
#ifndef PSX_SLOT_COH
#define PSX_SLOT_COH 2 // (measured, 2048 instances of the hazard variant: 16.37 -> 16.08 ms) slot mode: 1 = VRAM / CLUT reads from L2 (ld.global.cg), 2 = through L1 (the instance belongs to one SM)
#endif
#ifndef PSX_GRAB_NOPEEK
#define PSX_GRAB_NOPEEK 1
#endif
#ifndef PSX_REGION_PITCH
#define PSX_REGION_PITCH 36
#endif
constexpr uint32_t REGION_PITCH = PSX_REGION_PITCH;     // words per tile row in shared memory (+4: the same column of eight consecutive rows lands in eight different groups of four banks — narrow spans stacked over rows are what the word loop touches)
constexpr uint32_t REGION_PITCH16 = REGION_PITCH * 2;   // same, in pixels

// Per-warp shared memory.  raster_kernel: one warp == one CTA == one 64x32 tile (RH = 32), 5.5 KiB, so 32 CTAs fit an SM.
// raster_persistent_kernel: four warps per CTA, each owning a 64x8 strip of the tile (RH = 8).
template<uint32_t RH>
struct alignas(16) RasterShared
{
  uint32_t region[RH * REGION_PITCH];     // the region, 2 pixels per word (4608 B at RH = 32)
  uint32_t stage[2][SETUP_WORDS];         // the primitive being rasterised and the next one (cp.async double buffer)
  uint16_t clut[CLUT_ENTRIES];            // merged palette of the current primitive (entries 0-15 lo slot, 16-255 hi slot)
  uint16_t list[LIST_CAP];                // bin: primitive indices (relative to the epoch), submission order
  unsigned long long mbar[4];             // transaction barriers of the bulk copies (throughput kernels): stage[0], stage[1], region
};

// 16-byte asynchronous copy global -> shared (LDGSTS): no register holds the data while it is in flight.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem))), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16_l2(void* smem, const void* gmem) // .cg: from L2, never a stale L1 line
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem))), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template<int kPending>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory");
}

// ---- bulk asynchronous copies (the TMA engine, non-tensor form: UBLKCP in SASS) + transaction barriers ----
// One instruction moves a whole setup record (256 B) or a whole tile row (128 B) between global and shared memory; an
// mbarrier counts the bytes in.  No register holds data in flight, no per-16-byte address arithmetic, no LDG/STS pair
// per word.  Used by the throughput kernels (everything they read this way is constant while the kernel runs, and what
// they write back is not read again before the kernel ends).
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity)
{
  asm volatile("{\n"
               ".reg .pred p;\n"
               "MBAR_WAIT:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra MBAR_DONE;\n"
               "bra MBAR_WAIT;\n"
               "MBAR_DONE:\n"
               "}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, uint32_t bytes, unsigned long long* b)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(bytes), "r"(smem_u32(b))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem, const void* smem, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(smem_u32(smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() // the engine has read the shared memory of every committed bulk store
{
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem()
{
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr)
{
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}
// once per warp, before the first bulk copy (lane 0 initialises, the warp synchronises)
template<typename Shared>
__device__ __forceinline__ void init_bulk_barriers(Shared& sh, uint32_t lane)
{
  if (lane == 0)
  {
    mbar_init(&sh.mbar[0], 1);
    mbar_init(&sh.mbar[1], 1);
    mbar_init(&sh.mbar[2], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
}

struct PixelCtx
{
  const uint16_t* vram; // this instance's VRAM in global memory (texture / copy source reads)
  const uint16_t* clut; // the primitive's merged 256-entry palette, staged in shared memory
};

// ShadeState straight from the first words of a staged PrimSetup (see the field order in psx_types.h).
__device__ __forceinline__ ShadeState shade_state_from_words(const uint32_t* w)
{
  const uint32_t w0 = w[0], w1 = w[1];
  ShadeState st;
  st.flags = (w0 >> 8) & 0xFFFFu;                     // flags0 | flags1 << 8
  st.modes = (w0 >> 24) | (((w1 >> 5) & 3u) << 8);    // texmode | transparency mode << 8
  st.window = w[2];
  st.texbase = w[7];
  return st;
}

// kCoh: the kernel outlives an epoch (persistent kernel), so VRAM / CLUT-slot reads must not come from a stale L1 line.
// kCoh == 2: every access to this VRAM / these CLUT slots while the kernel runs comes from ONE SM (slot mode): an
// ordinary cached load is coherent with that SM's own stores (write-through L1), and repeated texels hit L1.
template<int kCoh, typename T>
__device__ __forceinline__ T ld_ro(const T* p)
{
  return kCoh == 2 ? *p : (kCoh ? __ldcg(p) : __ldg(p));
}

// One pixel through the pipeline.  bg = current value; returns the new value.
template<int kCoh>
__device__ __forceinline__ uint32_t shade_pixel(const ShadeState& st, const PixelCtx& ctx, uint32_t px, uint32_t Y, uint32_t r,
                                                uint32_t g, uint32_t b, uint32_t u, uint32_t v, uint32_t bg)
{
  uint32_t texel = 0;
  if (st.flags & F_TEXTURE)
    texel = fetch_texel<kCoh != 0, true>(st, ctx.vram, ctx.clut, ctx.clut, u, v);
  uint32_t colour;
  if (!shade_colour(st, px, Y, r, g, b, texel, colour))
    return bg;
  return blend_mask(st, bg, colour);
}

__device__ __forceinline__ uint32_t set_half(uint32_t word, uint32_t half, uint32_t value)
{
  return half ? ((word & 0x0000FFFFu) | (value << 16)) : ((word & 0xFFFF0000u) | (value & 0xFFFFu));
}

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, uint32_t)
{
  // shfl.up hands back a predicate "the source lane exists": add under it, no lane compare needed
#pragma unroll
  for (uint32_t o = 1; o < 32; o <<= 1)
    asm volatile("{ .reg .pred p; .reg .u32 t; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, t; }" : "+r"(v) : "r"(o));
  return v;
}

// Which row lane owns compacted pixel base + lane?  Row lane r starts at compacted index start_r (rows without pixels
// carry ROW_EMPTY).  Rows that start inside this batch of 32 leave a mark at their first pixel's lane (one REDUX.OR);
// rows that started at or before `base` are counted with a ballot; the sum is the pixel's ordinal among the rows that
// have pixels, and `rowmap` (lane k = the k-th such row, see row_lane_map) turns the ordinal into the row lane.
constexpr uint32_t ROW_EMPTY = 0x40000000u;
__device__ __forceinline__ uint32_t row_of_pixel(uint32_t start, uint32_t base, uint32_t rowmap, uint32_t lane)
{
  const uint32_t t = start - base;
  uint32_t mark;
  asm("shl.b32 %0, 1, %1;" : "=r"(mark) : "r"(t)); // PTX shl clamps the amount: 0 for t >= 32 (later batches, empty rows) and for "negative" t
  const uint32_t marks = __reduce_or_sync(0xFFFFFFFFu, mark);
  const uint32_t before = __popc(__ballot_sync(0xFFFFFFFFu, static_cast<int32_t>(t) < 0));
  const uint32_t k = before - 1u + __popc(marks & ((2u << lane) - 1u));
  return __shfl_sync(0xFFFFFFFFu, rowmap, k);
}
// lane k -> the row lane of the k-th set bit of `rows` (rows with pixels); usually one contiguous run.
__device__ __forceinline__ uint32_t row_lane_map(uint32_t rows, uint32_t lane)
{
  const uint32_t first = static_cast<uint32_t>(__ffs(static_cast<int>(rows))) - 1u;
  const uint32_t run = rows >> first;
  if ((run & (run + 1u)) == 0)
    return first + lane;
  uint32_t pos = 0, rem = lane;
#pragma unroll
  for (uint32_t step = 16; step >= 1; step >>= 1)
  {
    const uint32_t c = __popc(rows & (((1u << step) - 1u) << pos));
    if (rem >= c)
    {
      pos += step;
      rem -= c;
    }
  }
  return pos & 31u;
}

// Triangle or rectangle on one tile, two pixels per lane.
//
// Lane r works out the span of tile row r (one pass for all 32 rows).  The spans are cut into the 32-bit words of the
// tile they touch (a word = an even / odd column pair), a warp prefix sum compacts the words of all rows, and they are
// shaded 32 words at a time, one word per lane, with the packed pipeline of psx_pair.h: one shared-memory load and
// store, one dither / clamp per channel and one blend for both pixels.  What a word's lane needs from its row lane
// travels by shuffle: the word offset and the span's ends in one register, the per-row interpolant bases.
// Everything that depends only on the primitive (texture mode, raw / modulated, flat / Gouraud, transparency) is a
// warp-uniform branch around a short block, and the constants those blocks use come ready-made from the staged record
// (LoopDesc, written by setup_kernel): the loop keeps few registers live, so nothing is rematerialised per word.
template<uint32_t RH, int kCoh>
__device__ __forceinline__ void raster_spans(const PrimSetup& s, const uint16_t* vram, RasterShared<RH>& sh,
                                             uint32_t rx0, uint32_t ry0, uint32_t lane, bool clut_pending)
{
  const LoopDesc& D = s.desc;
  const uint32_t path = D.path;
  const auto& T = s.tri;

  // ---- spans: lane = tile row ----
  int32_t xs = 0, xe = 0;
  uint32_t rb_r = 0, rb_g = 0, rb_b = 0, rb_u = 0, rb_v = 0; // triangle: interpolant at tile column 0 of this row; rectangle: rb_v = v
  if (path & LP_TRI)
  {
    int32_t lo, hi, cy, xdelta = 0;
    bool ok;
    if (!(path & LP_COMPLEX))
    {
      // tri_row (psx_pixel.h) for a triangle whose coordinates stay inside 11 bits: no wrap of rows or columns, and
      // the drawing area's rows are already cut into the row ranges (psx_setup.h)
      cy = T.y0 + static_cast<int32_t>((ry0 + lane - static_cast<uint32_t>(static_cast<int32_t>(T.y0))) & (VRAM_H - 1));
      const uint32_t p = (cy >= T.y1) ? 1u : 0u;
      const uint32_t k = static_cast<uint32_t>(cy - T.lo[p]);
      ok = (RH == 32 || lane < RH) & (k < static_cast<uint32_t>(T.hi[p] - T.lo[p]));
      if (path & LP_INTERLACED)
        ok = ok & !interlace_skip(s, static_cast<uint32_t>(cy));
      const TriPart& P = T.part[p];
      const int32_t x_start = static_cast<int32_t>((P.l_start + static_cast<uint64_t>(k) * P.l_step) >> 32);
      const int32_t x_bound = static_cast<int32_t>((P.r_start + static_cast<uint64_t>(k) * P.r_step) >> 32);
      lo = imax(imax(x_start, s.clip_l), static_cast<int32_t>(rx0));
      hi = imin(imin(x_bound, s.clip_r + 1), static_cast<int32_t>(rx0 + TILE_W));
    }
    else
    {
      TriRow row;
      ok = (RH == 32 || lane < RH) && tri_row(s, ry0 + lane, row);
      cy = row.cy;
      xdelta = row.xdelta;
      lo = imax(row.x_lo, static_cast<int32_t>(rx0));
      hi = imin(row.x_hi, static_cast<int32_t>(rx0 + TILE_W));
    }
    if (ok & (hi > lo))
    {
      xs = lo - static_cast<int32_t>(rx0);
      xe = hi - static_cast<int32_t>(rx0);
    }
    const uint32_t ucy = static_cast<uint32_t>(cy), x0 = rx0 + static_cast<uint32_t>(xdelta);
    if (path & LP_GOURAUD)
    {
      rb_r = T.base[0] + T.ddy[0] * ucy + T.ddx[0] * x0;
      rb_g = T.base[1] + T.ddy[1] * ucy + T.ddx[1] * x0;
      rb_b = T.base[2] + T.ddy[2] * ucy + T.ddx[2] * x0;
    }
    if (path & LP_TEXTURED)
    {
      rb_u = T.base[3] + T.ddy[3] * ucy + T.ddx[3] * x0;
      rb_v = T.base[4] + T.ddy[4] * ucy + T.ddx[4] * x0;
    }
  }
  else
  {
    RectRow row;
    if ((RH == 32 || lane < RH) && rect_row(s, ry0 + lane, row))
    {
      xs = imax(row.x_lo, static_cast<int32_t>(rx0)) - static_cast<int32_t>(rx0);
      xe = imin(row.x_hi, static_cast<int32_t>(rx0 + TILE_W)) - static_cast<int32_t>(rx0);
      rb_v = row.v * 0x10001u;
    }
  }
  const uint32_t cnt = (xe > xs) ? static_cast<uint32_t>(((xe - 1) >> 1) - (xs >> 1) + 1) : 0u; // words touched
  const uint32_t rows = __ballot_sync(0xFFFFFFFFu, cnt != 0);
  if (rows == 0)
  {
    if (clut_pending) // (the palette counts as staged from here on)
    {
      cp_async_wait<0>();
      __syncwarp();
    }
    return;
  }
  const uint32_t incl = warp_inclusive_scan(cnt, lane);
  const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
  const uint32_t start = cnt ? (incl - cnt) : ROW_EMPTY;
  const uint32_t rowmap = row_lane_map(rows, lane);
  // word of compacted index j = wrel + j; the span is [xs, xe) in tile columns
  const uint32_t packed = (static_cast<uint32_t>((xs >> 1) - static_cast<int32_t>(incl - cnt)) & 0xFFFFu) |
                          (static_cast<uint32_t>(xs) << 16) | (static_cast<uint32_t>(xe) << 24);
  const uint32_t ub = (D.ubase + rx0) & 0xFFu; // rectangle: u of tile column 0 (mod 256: (ub + column) * 0x10001 cannot carry)

  // The word loop exists in three instantiations — textured flat, textured Gouraud, untextured — chosen per visit: the
  // flags fixed by the instantiation (kMask / kVal) fold at compile time, which removes their warp-uniform branches
  // and the zero-initialisations behind them from every iteration; the other flags stay run-time (uniform) tests.
  auto word_loop = [&](auto kmask_c, auto kval_c) {
    constexpr uint32_t kMask = decltype(kmask_c)::value, kVal = decltype(kval_c)::value;
    const uint32_t path = (D.path & ~kMask) | kVal;
    // The loop is software-pipelined by hand: the row lookup of the next batch (a REDUX, two POPC and two dependent
    // shuffles) is issued right after this batch's texel loads, so that it runs while they are in flight.  (A deeper
    // pipeline — the whole address / texel-load stage of batch b + 1 ahead of the shading of batch b, nine values carried
    // — was built and measured: 1.552 against 1.528 ms, and much slower in the 64-register resident kernels.)
    // warp-uniform constants of the loop, read once (the compiler keeps them in uniform registers)
    const uint32_t c_chk15 = D.chk15, c_set15 = D.set15, c_tex_shift = D.tex_shift, c_tex_bx2 = D.tex_bx2, c_trow = D.tex_rows;
    const uint32_t c_sub_mask2 = D.sub_mask2, c_sub_shift = D.sub_shift, c_im = D.idx_mask, c_fr = D.fr, c_fg = D.fg, c_fb = D.fb;
    const uint32_t c_du = T.ddx[3], c_dv = T.ddx[4];
#if PSX_OPT_CLUTLD
    const uint32_t clut_s = smem_u32(sh.clut);
#endif
    uint32_t r = row_of_pixel(start, 0, rowmap, lane);
    uint32_t pk = __shfl_sync(0xFFFFFFFFu, packed, r);
  #pragma unroll 1
    for (uint32_t base = 0; base < total; base += 32)
    {
      const uint32_t w = pk_prmt<0x9910>(pk, 0) + base + lane; // sign-extended word offset + j = word of the tile row
      // coverage of the two halves: column 2w >= xs, column 2w + 1 < xe (bit 15 of each half; the bias keeps the halves
      // apart).  Lanes past the last word of the batch land behind their row's last word: not covered.
      const uint32_t pos = w * 0x00020002u + 0x80018000u;
      const uint32_t cov15 = (pos - pk_prmt<0x4242>(pk, 0)) & ~(pos - pk_prmt<0x4343>(pk, 0)) & 0x80008000u;
      uint32_t* const wp = sh.region + r * REGION_PITCH + w;
      const uint32_t bg = *wp; // (lanes past the end read at most 31 words behind their row: still inside this warp's shared memory)

      uint32_t d16 = 0;
      if (path & LP_DITHER)
      {
        // pair_dither_entry in closed form: rows 2, 3 are rows 0, 1 with the column pairs swapped
        d16 = 0x0000FFC0u + (r & 1u) * 0xFFDF0060u + ((w ^ (r >> 1)) & 1u) * 0x00100010u;
      }

      uint32_t t0 = 0, t1 = 0, U = 0;
      if (path & LP_TEXTURED)
      {
        uint32_t V;
        if (path & LP_TRI)
        {
          const uint32_t du = c_du, dv = c_dv;
          const uint32_t u0 = __shfl_sync(0xFFFFFFFFu, rb_u, r) + du * (2u * w);
          const uint32_t v0 = __shfl_sync(0xFFFFFFFFu, rb_v, r) + dv * (2u * w);
          U = __byte_perm(u0, u0 + du, 0x7733); // top bytes; the odd bytes are cleared by the masks below
          V = __byte_perm(v0, v0 + dv, 0x7733);
        }
        else
        {
          U = (ub + 2u * w) * 0x00010001u + 0x00010000u;
          V = __shfl_sync(0xFFFFFFFFu, rb_v, r);
        }
#if PSX_OPT_WIN
        // the window constants of a primitive without a texture window are and = 0xFF, or = 0 (make_loop_desc): one
        // form for both cases, no test
        U = (U & D.and_x2) | D.or_x2;
        V = (V & D.and_y2) | D.or_y2;
#else
        if (path & LP_WINDOW)
        {
          U = (U & D.and_x2) | D.or_x2;
          V = (V & D.and_y2) | D.or_y2;
        }
        else
        {
          U &= 0x00FF00FFu;
          V &= 0x00FF00FFu;
        }
#endif
        // psx_pixel.h:texel_addr on both pixels at once; page base row (0 or 256) + v never wraps
        const uint32_t tx = ((U >> c_tex_shift) + c_tex_bx2) & 0x03FF03FFu;
        const uint32_t trow = c_trow;
        PSX_CHECK(trow + ((V & 0xFFFFu) << 10) + (tx & 0xFFFFu) < VRAM_PIXELS && trow + ((V >> 16) << 10) + (tx >> 16) < VRAM_PIXELS);
        t0 = ld_ro<kCoh>(vram + (trow + ((V & 0xFFFFu) << 10) + (tx & 0xFFFFu)));
        t1 = ld_ro<kCoh>(vram + (trow + ((V >> 16) << 10) + (tx >> 16)));
      }
      // per-pixel colours (before the row lookup moves on)
      uint32_t ar = 0, ag = 0, ab = 0;
      if (path & LP_GOURAUD)
      {
        ar = __shfl_sync(0xFFFFFFFFu, rb_r, r) + T.ddx[0] * (2u * w);
        ag = __shfl_sync(0xFFFFFFFFu, rb_g, r) + T.ddx[1] * (2u * w);
        ab = __shfl_sync(0xFFFFFFFFu, rb_b, r) + T.ddx[2] * (2u * w);
      }
      // next batch's row lookup, while the texel loads are in flight
      if (base + 32 < total)
      {
        r = row_of_pixel(start, base + 32, rowmap, lane);
        pk = __shfl_sync(0xFFFFFFFFu, packed, r);
      }

      uint32_t tex = 0, we15;
      if (path & LP_TEXTURED)
      {
        if (path & LP_PALETTED)
        {
          const uint32_t sub = (U & c_sub_mask2) << c_sub_shift; // bit offset of the index in its word
          const uint32_t im = c_im;
          PSX_CHECK(im < CLUT_ENTRIES);
#if PSX_OPT_CLUTLD
          // (explicit 32-bit shared addresses from a base computed once per visit: the generic form re-derives the
          // shared window's base inside the loop)
          t0 = lds_u16(clut_s + 2u * ((t0 >> (sub & 0xFFFFu)) & im));
          t1 = lds_u16(clut_s + 2u * ((t1 >> (sub >> 16)) & im));
#else
          t0 = sh.clut[(t0 >> (sub & 0xFFFFu)) & im];
          t1 = sh.clut[(t1 >> (sub >> 16)) & im];
#endif
        }
        tex = __byte_perm(t0, t1, 0x5410);
        we15 = cov15 & (((tex & 0x7FFF7FFFu) + 0x7FFF7FFFu) | tex) & ~(bg & c_chk15); // pair_we_textured
      }
      else
        we15 = cov15 & ~(bg & c_chk15);

      uint32_t col15;
      if (!(path & LP_MODULATED))
        col15 = tex & 0x7FFF7FFFu;
      else
      {
        uint32_t pr, pg, pb;
        if (path & LP_GOURAUD)
        {
          const uint32_t dr = T.ddx[0], dg = T.ddx[1], db = T.ddx[2];
          if (path & LP_TEXTURED)
          {
            const uint32_t cm = D.cmask;
            const uint32_t tr = tex & 0x001F001Fu, tg = (tex >> 5) & 0x001F001Fu, tb = (tex >> 10) & 0x001F001Fu;
            pr = (tr & 0xFFFFu) * ((ar >> 24) & cm) + (tr & 0xFFFF0000u) * (((ar + dr) >> 24) & cm);
            pg = (tg & 0xFFFFu) * ((ag >> 24) & cm) + (tg & 0xFFFF0000u) * (((ag + dg) >> 24) & cm);
            pb = (tb & 0xFFFFu) * ((ab >> 24) & cm) + (tb & 0xFFFF0000u) * (((ab + db) >> 24) & cm);
          }
          else
          {
            pr = ((ar >> 24) << 4) + (((ar + dr) >> 24) << 20);
            pg = ((ag >> 24) << 4) + (((ag + dg) >> 24) << 20);
            pb = ((ab >> 24) << 4) + (((ab + db) >> 24) << 20);
          }
        }
        else if (path & LP_TEXTURED)
        {
          pr = (tex & 0x001F001Fu) * c_fr;
          pg = ((tex >> 5) & 0x001F001Fu) * c_fg;
          pb = ((tex >> 10) & 0x001F001Fu) * c_fb;
        }
        else
        {
          pr = c_fr;
          pg = c_fg;
          pb = c_fb;
        }
        col15 = pair_pack15(pr, pg, pb, d16);
      }
      const uint32_t t15 = tex & 0x80008000u;
      if (path & LP_TRANSP)
      {
        const uint32_t blended = pair_blend15((path >> LP_TMODE_SHIFT) & 3u, bg & 0x7FFF7FFFu, col15);
        const uint32_t bm = (path & LP_TEXTURED) ? pk_expand15(t15) : 0xFFFFFFFFu;
        col15 = (col15 & ~bm) | (blended & bm);
      }
      if (we15)
      {
PSX_CHECK(r < RH && w < TILE_W / 2);
                const uint32_t wm = pk_expand15(we15);
        *wp = (bg & ~wm) | ((col15 | t15 | c_set15) & wm);
      }
    }
  };
  if (clut_pending) // the palette requested by raster_prim_on_region (asynchronous copy) has to be there now
  {
    cp_async_wait<0>();
    __syncwarp();
  }
  using std::integral_constant;
  if (!(D.path & LP_TEXTURED))
    word_loop(integral_constant<uint32_t, LP_TEXTURED | LP_PALETTED | LP_WINDOW | LP_MODULATED>{}, integral_constant<uint32_t, LP_MODULATED>{});
  else if (D.path & LP_GOURAUD)
    word_loop(integral_constant<uint32_t, LP_TEXTURED | LP_GOURAUD | LP_MODULATED | LP_TRI>{},
              integral_constant<uint32_t, LP_TEXTURED | LP_GOURAUD | LP_MODULATED | LP_TRI>{});
  else
    word_loop(integral_constant<uint32_t, LP_TEXTURED | LP_GOURAUD>{}, integral_constant<uint32_t, LP_TEXTURED>{});
}

// A warp applies one primitive to its region (rows ry0 .. ry0+RH-1, columns rx0 .. rx0+63), held in shared memory.
// clut_key: which palette (lo slot | hi slot << 8 | depth << 16) currently sits in clut_smem; reloaded on change.
template<uint32_t RH, int kCoh>
__device__ __forceinline__ void raster_prim_on_region(const PrimSetup& s, RasterShared<RH>& sh, uint32_t rx0, uint32_t ry0,
                                                      const uint16_t* vram, const uint16_t* clut_pool_inst,
                                                      const uint16_t* __restrict__ payload, uint32_t& clut_key,
                                                      uint32_t lane)
{
  const uint32_t op = s.op;
  PSX_CHECK(op <= OP_NOP);
  if (op <= OP_RECT) // OP_TRIANGLE, OP_RECT
  {
    bool clut_pending = false;
    // paletted: make sure the merged palette is in shared memory (one coalesced 512-byte read when it changes)
    const uint32_t key = s.desc.clut_key; // lo slot | hi slot << 8 | 8-bit << 16 | 1 << 31, 0 = no palette
    if (key != 0 && key != clut_key)
    {
      clut_key = key;
      const uint32_t lo = key & 0xFFu, hi = (key >> 8) & 0xFFu, is8 = (key >> 16) & 1u;
      PSX_CHECK(lo < g_dbg.clut_slots && hi < g_dbg.clut_slots);
      uint4* dst = reinterpret_cast<uint4*>(sh.clut);
#if PSX_OPT_CLUTASYNC
      // asynchronous copy (no register, no wait here): the palette arrives while the spans are worked out, the word
      // loop waits for it.  Entries 0-15 (lanes 0-1) come from the lo slot, the rest of an 8-bit palette from the hi slot.
      if (is8 || lane < 2)
      {
        const uint4* src = reinterpret_cast<const uint4*>(clut_pool_inst + ((lane < 2) ? lo : hi) * CLUT_ENTRIES) + lane;
        if (kCoh)
          cp_async16_l2(dst + lane, src);
        else
          cp_async16(dst + lane, src);
      }
      cp_async_commit();
      clut_pending = true;
#else
      if (is8)
      {
        dst[lane] = ld_ro<kCoh>(reinterpret_cast<const uint4*>(clut_pool_inst + hi * CLUT_ENTRIES) + lane);
        __syncwarp();
      }
      if (lane < 2 && (!is8 || lo != hi))
        dst[lane] = ld_ro<kCoh>(reinterpret_cast<const uint4*>(clut_pool_inst + lo * CLUT_ENTRIES) + lane);
      __syncwarp();
#endif
    }
    raster_spans<RH, kCoh>(s, vram, sh, rx0, ry0, lane, clut_pending);
    return;
  }
  const ShadeState st = shade_state_from_words(reinterpret_cast<const uint32_t*>(&s));
  PixelCtx ctx;
  ctx.vram = vram;
  ctx.clut = sh.clut;
  uint32_t* const region = sh.region;
  uint16_t* region16 = reinterpret_cast<uint16_t*>(region);
  const int32_t x_min = static_cast<int32_t>(rx0);

  switch (op)
  {
    case OP_LINE:
    {
      if (s.line.x_major)
      {
        // at most one pixel per column: lanes take columns lane and lane+32
#pragma unroll
        for (int h = 0; h < 2; h++)
        {
          const int32_t px = x_min + static_cast<int32_t>(lane) + 32 * h;
          LinePixel p;
          if (line_step(s, line_step_for_column(s, px), p) && p.x == px)
          {
            const uint32_t Y = static_cast<uint32_t>(p.y) & (VRAM_H - 1);
            const uint32_t r = Y - ry0;
            if (r < RH)
            {
              uint16_t* q = region16 + r * REGION_PITCH16 + (px - x_min);
              *q = static_cast<uint16_t>(shade_pixel<kCoh>(st, ctx, static_cast<uint32_t>(px), Y, p.r, p.g, p.b, 0, 0, *q));
            }
          }
        }
      }
      else
      {
        // at most one pixel per row: lane r takes tile row r
        const uint32_t Y = ry0 + lane;
        LinePixel p;
        if ((RH == 32 || lane < RH) && line_step(s, line_step_for_row(s, Y), p) && (static_cast<uint32_t>(p.y) & (VRAM_H - 1)) == Y)
        {
          const uint32_t col = static_cast<uint32_t>(p.x) - rx0;
          if (col < TILE_W)
          {
            uint16_t* q = region16 + lane * REGION_PITCH16 + col;
            *q = static_cast<uint16_t>(shade_pixel<kCoh>(st, ctx, static_cast<uint32_t>(p.x), Y, p.r, p.g, p.b, 0, 0, *q));
          }
        }
      }
    }
    break;

    // transfers: lane = one word (2 pixels) of the row, rows in a loop
    case OP_FILL:
    {
      const uint32_t px0 = rx0 + 2 * lane;
      // whole tile inside the rectangle and no interlace skip: 32 stores, nothing to test
      if (!s.fill.interlaced && ((rx0 - s.fill.x) & (VRAM_W - 1)) + TILE_W <= s.fill.w &&
          ((ry0 - s.fill.y) & (VRAM_H - 1)) + RH <= s.fill.h)
      {
        const uint32_t c2 = static_cast<uint32_t>(s.fill.color16) * 0x10001u;
#pragma unroll 8
        for (uint32_t r = 0; r < RH; r++)
          region[r * REGION_PITCH + lane] = c2;
        break;
      }
#pragma unroll 4
      for (uint32_t r = 0; r < RH; r++)
      {
        const uint32_t Y = ry0 + r;
        const bool c0 = fill_covers(s, px0, Y), c1 = fill_covers(s, px0 + 1, Y);
        if (!(c0 || c1))
          continue;
        uint32_t word = region[r * REGION_PITCH + lane];
        if (c0)
          word = set_half(word, 0, s.fill.color16);
        if (c1)
          word = set_half(word, 1, s.fill.color16);
        region[r * REGION_PITCH + lane] = word;
      }
    }
    break;

    case OP_COPY:
    {
      const uint32_t px0 = rx0 + 2 * lane;
      // whole tile inside the destination rectangle, no mask test, the 64 source columns do not wrap: a straight copy
      // (the same shape as the upload path below; source rows may wrap, they are addressed one by one)
      {
        const uint32_t cx = (rx0 - s.copy.dx) & (VRAM_W - 1), cy = (ry0 - s.copy.dy) & (VRAM_H - 1);
        const uint32_t sx = (s.copy.sx + cx) & (VRAM_W - 1);
        if (!(s.flags0 & F_CHECK_MASK) && cx + TILE_W <= s.copy.w && cy + RH <= s.copy.h && sx + TILE_W <= VRAM_W)
        {
          const uint32_t mask_or = (s.flags0 & F_SET_MASK) ? 0x80008000u : 0u;
          const uint32_t col = sx + 2 * lane;
#pragma unroll 8
          for (uint32_t r = 0; r < RH; r++)
          {
            const uint16_t* src = vram + ((s.copy.sy + cy + r) & (VRAM_H - 1)) * VRAM_W + col;
            uint32_t word;
            if (sx & 1u) // (the same for every lane and row)
              word = static_cast<uint32_t>(ld_ro<kCoh>(src)) | (static_cast<uint32_t>(ld_ro<kCoh>(src + 1)) << 16);
            else
              word = ld_ro<kCoh>(reinterpret_cast<const uint32_t*>(src));
            region[r * REGION_PITCH + lane] = word | mask_or;
          }
          break;
        }
      }
#pragma unroll 2
      for (uint32_t r = 0; r < RH; r++)
      {
        const uint32_t Y = ry0 + r;
        uint32_t s0, s1;
        const bool c0 = copy_covers(s, px0, Y, s0), c1 = copy_covers(s, px0 + 1, Y, s1);
        if (!(c0 || c1))
          continue;
        uint32_t word = region[r * REGION_PITCH + lane];
        if (c0)
          word = set_half(word, 0, transfer_write(s, static_cast<uint16_t>(word), ld_ro<kCoh>(vram + s0)));
        if (c1)
          word = set_half(word, 1, transfer_write(s, static_cast<uint16_t>(word >> 16), ld_ro<kCoh>(vram + s1)));
        region[r * REGION_PITCH + lane] = word;
      }
    }
    break;

    case OP_UPLOAD:
    {
      const uint32_t px0 = rx0 + 2 * lane;
      // whole tile inside the rectangle and no mask test: a straight copy of 32 payload rows, the loads of eight rows in
      // flight at a time (the general loop below has one dependent payload load per row and lane)
      {
        const uint32_t cx = (rx0 - s.upload.x) & (VRAM_W - 1), cy = (ry0 - s.upload.y) & (VRAM_H - 1);
        if (!(s.flags0 & F_CHECK_MASK) && cx + TILE_W <= s.upload.w && cy + RH <= s.upload.h)
        {
          const uint32_t mask_or = (s.flags0 & F_SET_MASK) ? 0x80008000u : 0u;
          const uint32_t w = s.upload.w;
          uint32_t idx = s.upload.payload + cy * w + cx + 2 * lane;
#pragma unroll 8
          for (uint32_t r = 0; r < RH; r++)
          {
            PSX_CHECK(idx + 1 < g_dbg.n_payload);
            uint32_t word;
            if (idx & 1u) // (the same for every lane of the row)
              word = static_cast<uint32_t>(__ldg(payload + idx)) | (static_cast<uint32_t>(__ldg(payload + idx + 1)) << 16);
            else
              word = __ldg(reinterpret_cast<const uint32_t*>(payload + idx));
            region[r * REGION_PITCH + lane] = word | mask_or;
            idx += w;
          }
          break;
        }
      }
#pragma unroll 2
      for (uint32_t r = 0; r < RH; r++)
      {
        const uint32_t Y = ry0 + r;
        uint32_t p0, p1;
        const bool c0 = upload_covers(s, px0, Y, p0), c1 = upload_covers(s, px0 + 1, Y, p1);
        if (!(c0 || c1))
          continue;
        PSX_CHECK((!c0 || p0 < g_dbg.n_payload) && (!c1 || p1 < g_dbg.n_payload));
        uint32_t word = region[r * REGION_PITCH + lane];
        if (c0)
          word = set_half(word, 0, transfer_write(s, static_cast<uint16_t>(word), __ldg(payload + p0)));
        if (c1)
          word = set_half(word, 1, transfer_write(s, static_cast<uint16_t>(word >> 16), __ldg(payload + p1)));
        region[r * REGION_PITCH + lane] = word;
      }
    }
    break;

    default: break;
  }
}

// ---- serial epoch kinds: one CTA, the reference's own order ----

// a textured primitive that samples pixels it writes.  The reference (4-wide vector build) gathers the
// texels of 4 pixels, then stores those 4 pixels; rows and groups are visited in walk order.
//
// The splitter sends a primitive here when its texture footprint RECTANGLE meets its write rectangle; whether a texel
// that is actually fetched is a pixel the primitive writes is only known per pixel.  So the CTA first walks the
// primitive without touching memory (all warps, rows dealt round-robin) and asks of every texel address whether it
// lies inside the write bounding box.  If none does — the usual case — no fetch can see one of the primitive's own
// stores, order is unobservable and the rows are shaded by all warps at once.  Otherwise warp 0 replays the
// reference's order: eight consecutive 4-pixel groups of a row per step, group by group where a texel of group k is
// one of the pixels that groups 0..k-1 of the step store.
template<int NT>
__device__ void serial_self_sampling(const PrimSetup& s, uint16_t* vram, const uint16_t* clut_pool_inst)
{
  constexpr uint32_t N_WARPS = NT / 32;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint16_t* clut_lo = clut_pool_inst + s.clut_lo * CLUT_ENTRIES;
  const uint16_t* clut_hi = clut_pool_inst + s.clut_hi * CLUT_ENTRIES;
  volatile uint16_t* vv = vram;
  const ShadeState st = make_shade_state(s);
  enum Mode : uint32_t { CHECK, PARALLEL, ORDERED };
  bool conflict_seen = false;

  // 32 pixels of one row: x0 .. x0+31, lane = pixel
  auto batch = [&](uint32_t mode, int32_t x0, uint32_t Y, bool active, uint32_t r, uint32_t g, uint32_t b, uint32_t u, uint32_t v) {
    uint32_t shift, idxmask;
    const uint32_t addr = texel_addr(st, u, v, shift, idxmask);
    if (mode == CHECK)
    {
      const uint32_t tcol = addr & (VRAM_W - 1), trow = addr >> 10;
      conflict_seen |= active && static_cast<int32_t>(tcol) >= s.bx0 && static_cast<int32_t>(tcol) <= s.bx1 &&
                       ((trow - s.by0) & (VRAM_H - 1)) < s.bycount;
      return;
    }
    const uint32_t first_px = Y * VRAM_W + static_cast<uint32_t>(x0);
    // the lane's own pixel is written by nobody else in this step: its old value can be fetched with the texel
    const uint16_t bg = active ? vv[first_px + lane] : uint16_t(0);
    bool serial = false;
    if (mode == ORDERED)
    {
      const bool conflict = active && (addr - first_px) < (lane & ~3u); // texel in [first_px, first_px + 4k), k = lane / 4
      serial = __any_sync(0xFFFFFFFFu, conflict);
    }
    for (uint32_t k = 0; k < (serial ? 8u : 1u); k++)
    {
      const bool mine = active && (!serial || (lane >> 2) == k);
      uint32_t texel = 0;
      if (mine)
      {
        const uint32_t w = vv[addr];
        texel = w;
        if (idxmask != 0)
        {
          const uint32_t idx = (w >> shift) & idxmask;
          texel = __ldcg((idx < 16u ? clut_lo : clut_hi) + idx); // (coherent: the kernels that stay resident rewrite CLUT slots between epochs)
        }
      }
      if (mode == ORDERED)
        __syncwarp();
      if (mine)
      {
        uint32_t colour;
        const uint32_t px = static_cast<uint32_t>(x0) + lane;
        if (shade_colour(st, px, Y, r, g, b, texel, colour))
          vv[Y * VRAM_W + px] = static_cast<uint16_t>(blend_mask(st, bg, colour));
      }
      if (mode == ORDERED)
        __syncwarp();
    }
  };

  // rows in walk order; in CHECK / PARALLEL mode warp w takes the rows whose ordinal is w mod N_WARPS
  auto walk = [&](uint32_t mode) {
    uint32_t ordinal = 0;
    auto my_row = [&]() {
      const bool take = mode == ORDERED || (ordinal % N_WARPS) == warp;
      ordinal++;
      return take;
    };
    if (s.op == OP_RECT)
    {
      for (uint32_t oy = 0; oy < s.rect.h; oy++)
      {
        if (!my_row())
          continue;
        const int32_t y = s.rect.y + static_cast<int32_t>(oy);
        const uint32_t Y = static_cast<uint32_t>(y) & (VRAM_H - 1);
        RectRow row;
        if (!rect_row(s, Y, row) || row.y != y)
          continue;
        // groups are counted from the unclipped origin (inl:777-779); 32 is a multiple of 4, so batches keep that phase
        for (uint32_t ox = 0; ox < s.rect.w; ox += 32)
        {
          const int32_t x = s.rect.x + static_cast<int32_t>(ox + lane);
          const bool active = (ox + lane) < s.rect.w && x >= row.x_lo && x < row.x_hi;
          batch(mode, s.rect.x + static_cast<int32_t>(ox), Y, active, s.rect.r, s.rect.g, s.rect.b, (s.rect.u0 + ox + lane) & 0xFFu, row.v);
        }
      }
      return;
    }
    if (s.op != OP_TRIANGLE)
      return;
    const auto& T = s.tri;
    const bool upper_up = (T.tflags & TRI_UPPER_UP) != 0, lower_up = (T.tflags & TRI_LOWER_UP) != 0;
    // part order = triparts[0], triparts[1] (inl:1557-1560); the upper part is triparts[vo]
    for (int part = 0; part < 2; part++)
    {
      const bool is_upper = (part == 0) != upper_up;
      const int32_t lo = T.lo[is_upper ? 0 : 1], hi = T.hi[is_upper ? 0 : 1];
      const bool up = is_upper ? upper_up : lower_up;
      for (int32_t k = 0; k < hi - lo; k++)
      {
        if (!my_row())
          continue;
        const int32_t cy = up ? (hi - 1 - k) : (lo + k);
        const uint32_t Y = static_cast<uint32_t>(sext11(cy)) & (VRAM_H - 1);
        TriRow row;
        if (!tri_row(s, Y, row) || row.cy != cy)
          continue;
        // groups are counted from the clipped span start (inl:1262-1268)
        for (int32_t gx = row.x_lo; gx < row.x_hi; gx += 32)
        {
          const int32_t px = gx + static_cast<int32_t>(lane);
          const bool active = px < row.x_hi;
          uint32_t r, g, b, u, v;
          tri_attrs(s, row, px, r, g, b, u, v);
          batch(mode, gx, Y, active, r, g, b, u, v);
        }
      }
    }
  };

  walk(CHECK);
  if (!__syncthreads_or(conflict_seen ? 1 : 0))
    walk(PARALLEL);
  else if (warp == 0)
    walk(ORDERED);
}

// CopyVRAM whose source and destination alias: wrap split in the reference's order, rows sequential, each row
// read completely before it is written (CopyVRAMImpl, inl:1833-1939).
template<int NT>
__device__ void serial_copy_rows(uint16_t* vram, uint32_t sx, uint32_t sy, uint32_t dx, uint32_t dy, uint32_t w, uint32_t h,
                                 uint16_t mand, uint16_t mor)
{
  // No column wrap in here (the caller splits those).  The reference walks the rows in order; what that order can be
  // observed through is limited to three things, with d = (dy - sy) mod 512:
  //   d == 0: source and destination share rows; a row is read completely before it is written (memmove), and rows do
  //           not interact  ->  one warp per row, all rows at once;
  //   row r reads what row r - d wrote (if d < h), and row r + (512 - d) overwrites what row r read (if 512 - d < h)
  //           ->  any G = min(d, 512 - d, h) consecutive rows are free of hazards and are copied as one parallel batch.
  volatile uint16_t* vv = vram;
  const uint32_t d = (dy - sy) & (VRAM_H - 1);
  if (d == 0)
  {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (uint32_t r = warp; r < h; r += NT / 32)
    {
      const uint32_t row = ((sy + r) & (VRAM_H - 1)) * VRAM_W;
      uint16_t px[VRAM_W / 32];
#pragma unroll
      for (uint32_t k = 0; k < VRAM_W / 32; k++)
      {
        const uint32_t c = k * 32 + lane;
        px[k] = (c < w) ? vv[row + sx + c] : uint16_t(0);
      }
      __syncwarp();
#pragma unroll
      for (uint32_t k = 0; k < VRAM_W / 32; k++)
      {
        const uint32_t c = k * 32 + lane;
        if (c < w)
        {
          const uint32_t a = row + dx + c;
          if ((vv[a] & mand) == 0)
            vv[a] = static_cast<uint16_t>(px[k] | mor);
        }
      }
    }
    __syncthreads();
    return;
  }
  uint32_t G = h;
  if (d < h)
    G = min(G, d);
  if (VRAM_H - d < h)
    G = min(G, VRAM_H - d);
  for (uint32_t r0 = 0; r0 < h; r0 += G)
  {
    const uint32_t rows = min(G, h - r0);
    // (row, column) pairs of the batch, NT at a time; a thread keeps its row/column by stepping instead of dividing
    uint32_t rr = threadIdx.x / w, c = threadIdx.x - rr * w;
    const uint32_t step_r = NT / w, step_c = NT - step_r * w;
    while (rr < rows)
    {
      const uint32_t y = r0 + rr;
      const uint16_t src = vv[((sy + y) & (VRAM_H - 1)) * VRAM_W + sx + c];
      const uint32_t a = ((dy + y) & (VRAM_H - 1)) * VRAM_W + dx + c;
      if ((vv[a] & mand) == 0)
        vv[a] = static_cast<uint16_t>(src | mor);
      rr += step_r;
      c += step_c;
      if (c >= w)
      {
        c -= w;
        rr++;
      }
    }
    __syncthreads();
  }
}

template<int NT>
__device__ void serial_copy(const PrimSetup& s, uint16_t* vram)
{
  const uint16_t mand = (s.flags0 & F_CHECK_MASK) ? 0x8000u : 0, mor = (s.flags0 & F_SET_MASK) ? 0x8000u : 0;
  const uint32_t sx = s.copy.sx, sy = s.copy.sy, dx = s.copy.dx, dy = s.copy.dy, w = s.copy.w, h = s.copy.h;
  if ((sx + w) <= VRAM_W && (dx + w) <= VRAM_W)
  {
    serial_copy_rows<NT>(vram, sx, sy, dx, dy, w, h, mand, mor);
    return;
  }
  uint32_t rem_rows = h, csy = sy, cdy = dy;
  while (rem_rows > 0)
  {
    const uint32_t rows = min(rem_rows, min(VRAM_H - csy, VRAM_H - cdy));
    uint32_t rem_cols = w, csx = sx, cdx = dx;
    while (rem_cols > 0)
    {
      const uint32_t cols = min(rem_cols, min(VRAM_W - csx, VRAM_W - cdx));
      serial_copy_rows<NT>(vram, csx, csy, cdx, cdy, cols, rows, mand, mor);
      csx = (csx + cols) & (VRAM_W - 1);
      cdx = (cdx + cols) & (VRAM_W - 1);
      rem_cols -= cols;
    }
    csy = (csy + rows) & (VRAM_H - 1);
    cdy = (cdy + rows) & (VRAM_H - 1);
    rem_rows -= rows;
  }
}

// mask-checked upload.  Source index of a pixel = its linear index minus the number of masked pixels
// before it (in scan order) that sit in the scalar tail of their row (columns >= aw).  Rows never revisit a pixel
// (w <= 1024, h <= 512), so the masked-tail count of every row can be taken up front, prefix-summed over the rows, and
// all rows written at once — one warp per row.  `row_count`: shared, VRAM_H + 1 words.
template<int NT>
__device__ void serial_upload_masked(const PrimSetup& s, uint16_t* vram, const uint16_t* __restrict__ payload,
                                     uint32_t* row_count)
{
  const auto& U = s.upload;
  const uint32_t aw = min(static_cast<uint32_t>(U.w), VRAM_W - U.x) & ~7u;
  const uint16_t mor = (s.flags0 & F_SET_MASK) ? 0x8000u : 0;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t h = min(static_cast<uint32_t>(U.h), VRAM_H), w = U.w;
  // pass 1: masked pixels in the tail of every row
  for (uint32_t r = warp; r < h; r += NT / 32)
  {
    const uint32_t drow = ((U.y + r) & (VRAM_H - 1)) * VRAM_W;
    uint32_t n = 0;
    for (uint32_t base = aw; base < w; base += 32)
    {
      const uint32_t c = base + lane;
      const bool masked = c < w && (vram[drow + ((U.x + c) & (VRAM_W - 1))] & 0x8000u);
      n += __popc(__ballot_sync(0xFFFFFFFFu, masked));
    }
    if (lane == 0)
      row_count[r] = n;
  }
  __syncthreads();
  // pass 2: exclusive prefix over the rows (one warp, 16 rows per lane)
  if (warp == 0)
  {
    uint32_t local[VRAM_H / 32], sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < VRAM_H / 32; k++)
    {
      const uint32_t r = lane * (VRAM_H / 32) + k;
      local[k] = sum;
      sum += (r < h) ? row_count[r] : 0u;
    }
    uint32_t incl = sum;
#pragma unroll
    for (uint32_t o = 1; o < 32; o <<= 1)
    {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o)
        incl += t;
    }
    const uint32_t before = incl - sum;
    __syncwarp();
#pragma unroll
    for (uint32_t k = 0; k < VRAM_H / 32; k++)
    {
      const uint32_t r = lane * (VRAM_H / 32) + k;
      if (r < h)
        row_count[r] = before + local[k];
    }
  }
  __syncthreads();
  // pass 3: every row on its own
  for (uint32_t r = warp; r < h; r += NT / 32)
  {
    const uint32_t drow = ((U.y + r) & (VRAM_H - 1)) * VRAM_W;
    uint32_t deficit = row_count[r];
    for (uint32_t base = 0; base < w; base += 32)
    {
      const uint32_t c = base + lane;
      const bool in = c < w;
      const uint32_t a = drow + ((U.x + c) & (VRAM_W - 1));
      const uint16_t d = in ? vram[a] : uint16_t(0);
      const bool masked = in && (d & 0x8000u);
      const uint32_t bal = __ballot_sync(0xFFFFFFFFu, masked && c >= aw);
      const uint32_t before = __popc(bal & ((1u << lane) - 1u));
      PSX_CHECK(!(in && !masked) || U.payload + r * w + c - (deficit + before) < g_dbg.n_payload);
      if (in && !masked)
        vram[a] = static_cast<uint16_t>(__ldg(payload + U.payload + r * w + c - (deficit + before)) | mor);
      deficit += __popc(bal);
    }
  }
  __syncthreads();
}

// Epochs of a serial kind: one CTA per item, everything else returns at once.
__global__ void __launch_bounds__(SERIAL_THREADS) serial_kernel(const WaveItem* __restrict__ items, const PrimSetup* __restrict__ setups,
                                                                uint16_t* vram_pool, const uint16_t* __restrict__ clut_pool,
                                                                const uint16_t* __restrict__ payload, uint32_t clut_slots)
{
  __shared__ alignas(16) uint32_t stage[SETUP_WORDS];
  __shared__ uint32_t row_count[VRAM_H + 1];
  const WaveItem it = items[blockIdx.x];
  if (it.kind == EPOCH_TILED || it.cmd_begin == it.cmd_end)
    return;
  uint16_t* vram = vram_pool + static_cast<size_t>(it.instance) * VRAM_PIXELS;
  const uint16_t* cluts = clut_pool + static_cast<size_t>(it.instance) * clut_slots * CLUT_ENTRIES;
  if (threadIdx.x < SETUP_WORDS)
    stage[threadIdx.x] = reinterpret_cast<const uint32_t*>(setups + it.cmd_begin)[threadIdx.x];
  __syncthreads();
  const PrimSetup& s = *reinterpret_cast<const PrimSetup*>(stage);
  if (s.op == OP_NOP)
    return;
  if (it.kind == EPOCH_SELF_SAMPLING)
    serial_self_sampling<SERIAL_THREADS>(s, vram, cluts);
  else if (it.kind == EPOCH_COPY_OVERLAP)
    serial_copy<SERIAL_THREADS>(s, vram);
  else
    serial_upload_masked<SERIAL_THREADS>(s, vram, payload, row_count);
}

// Binning step of a region's warp: appends the commands of [next_cmd, it.cmd_end) that can touch tile (tx, ty) to the
// shared list, in command order, until the epoch's commands are all in it or the list is (nearly) full.
__device__ __forceinline__ void bin_fill(const WaveItem& it, uint16_t* list, uint32_t tx, uint32_t ty, const uint32_t* __restrict__ tile_cmds,
                                         uint32_t& list_n, uint32_t& next_cmd, uint32_t lane)
{
  // fill the list until the epoch's commands are all in it, or it is (nearly) full
#pragma unroll 1
  while (next_cmd < it.cmd_end)
  {
    const uint32_t g = next_cmd >> 5, mine = g + lane;
    const uint32_t c0 = mine << 5; // first command of this lane's group
    uint32_t bits = 0;
    if (c0 < it.cmd_end)
    {
      PSX_CHECK(tx < TILES_X && ty < TILES_Y && mine <= (g_dbg.n_cmds + 31u) / 32u);
      bits = __ldg(tile_cmds + (static_cast<size_t>(mine) * NUM_TILES + (ty * TILES_X + tx)));
      if (c0 < next_cmd)
        bits &= 0xFFFFFFFFu << (next_cmd - c0);
      if (c0 + 32u > it.cmd_end)
        bits &= (1u << (it.cmd_end - c0)) - 1u;
    }
    if (!__any_sync(0xFFFFFFFFu, bits != 0))
    {
      next_cmd = (g + 32u) << 5;
      continue;
    }
    const uint32_t cnt = __popc(bits);
    const uint32_t incl = warp_inclusive_scan(cnt, lane);
    const bool fits = list_n + incl <= LIST_CAP;
    const uint32_t n_fit = __popc(__ballot_sync(0xFFFFFFFFu, fits)); // a prefix of the lanes: incl is monotone
    // (REDUX results are uniform for the compiler: the warp stays converged for the collectives of the raster code)
    const uint32_t added = __reduce_add_sync(0xFFFFFFFFu, fits ? cnt : 0u);
    const uint32_t rounds = __reduce_max_sync(0xFFFFFFFFu, fits ? cnt : 0u);
    uint32_t pos = list_n + incl - cnt;
    const uint32_t rel = c0 - it.cmd_begin; // (wraps for the first group: those bits are masked off)
    for (uint32_t k = 0; k < rounds; k++)
    {
      if (fits && bits)
      {
        const uint32_t c = static_cast<uint32_t>(__ffs(static_cast<int>(bits))) - 1u;
        bits &= bits - 1u;
        PSX_CHECK(pos < LIST_CAP && rel + c < it.cmd_end - it.cmd_begin);
        list[pos++] = static_cast<uint16_t>(rel + c);
      }
    }
    list_n += added;
    next_cmd = max(next_cmd, (g + n_fit) << 5);
    if (n_fit != 32u || list_n + 32u > LIST_CAP)
      break;
  }
}

// What a warp can do for its next region before the barrier that ends the current epoch opens: bin the region (the
// bins and the setup records do not change while the kernel runs) and request the first command's setup record.
struct PreBin
{
  uint32_t list_n, next_cmd;
};

// One warp runs one epoch of one instance on its region: bins the epoch's primitives (submission order), stages the
// region in shared memory on the first hit, rasterises, writes it back.
template<uint32_t RH, int kCoh>
__device__ __forceinline__ void process_epoch_region(const WaveItem& it, RasterShared<RH>& sh, uint32_t tx, uint32_t ty, uint32_t ry0,
                                                     const PrimSetup* __restrict__ setups, const uint32_t* __restrict__ tile_cmds,
                                                     uint16_t* vram, const uint16_t* cluts, const uint16_t* __restrict__ payload,
                                                     uint32_t lane, const PreBin* pre = nullptr, uint32_t* clut_key_io = nullptr,
                                                     uint32_t* mb_phase_io = nullptr)
{
  // kBulk (the throughput kernels): setup records and tile rows travel as bulk asynchronous copies counted in by
  // transaction barriers; *mb_phase_io carries the barriers' phase bits from one region of the warp to the next.
  // Otherwise (the kernels that stay resident across epochs and read what other CTAs wrote): 16-byte cp.async from L2.
  constexpr bool kBulk = !kCoh;
  // How the 256-byte setup records reach shared memory in the throughput kernels (PSX_REC_MODE): 0 = one bulk copy per
  // record + transaction barrier, 1 = 16-byte cp.async (as the resident kernels), 2 = each lane loads 8 bytes of the NEXT
  // record into two registers while the current one is rasterised and stores them when its turn comes.
  constexpr bool kBulkRec = kBulk && PSX_REC_MODE == 0;
  constexpr bool kRegRec = kBulk && PSX_REC_MODE == 2;
  uint2 rec_next = make_uint2(0u, 0u);
  PSX_CHECK(it.cmd_begin <= it.cmd_end && it.cmd_end <= g_dbg.n_cmds && it.instance < g_dbg.n_instances);
  PSX_CHECK(tx < TILES_X && ry0 + RH <= VRAM_H && (ry0 % RH) == 0);
  const uint32_t rx0 = tx * TILE_W;
  uint32_t* gword = reinterpret_cast<uint32_t*>(vram + ry0 * VRAM_W + rx0) + lane; // this lane's word of region row 0

  // Binning: tile_cmds holds, per group of 32 commands, one word per tile (setup_kernel).  A step covers 32 groups
  // (1024 commands), one load per lane; the set bits of the lanes expand into the list in command order, as many
  // lanes at a time as the list has room for.  All that survives a raster pass is next_cmd (uniform).
  bool loaded = false;
  uint32_t clut_key = clut_key_io ? *clut_key_io : 0u; // (a caller that knows the palette slots did not change keeps the staged palette)
  uint32_t list_n = pre ? pre->list_n : 0u;
  uint32_t next_cmd = pre ? pre->next_cmd : it.cmd_begin; // first command of the epoch that is not in a list yet
  bool first_requested = pre != nullptr; // pre-binned: the list's first setup record is already on its way to stage[0]
  uint32_t mb_phase = (kBulk && mb_phase_io) ? *mb_phase_io : 0u;
  if (pre && list_n == 0)
    return; // pre-binned and nothing of the epoch touches the region
  if (kCoh)
  {
    // The kernels that stay resident over many short epochs are bound by the chain of dependent round trips per tile
    // (bins -> setup record -> tile -> texel): the tile's rows are requested first, asynchronously, and arrive while the
    // bins and the first setup record are fetched.
    const uint16_t* gtile = vram + ry0 * VRAM_W + rx0;
#pragma unroll
    for (uint32_t c = lane; c < RH * (TILE_W / 8); c += 32)
      cp_async16_l2(sh.region + (c / (TILE_W / 8)) * REGION_PITCH + (c % (TILE_W / 8)) * 4, gtile + (c / (TILE_W / 8)) * VRAM_W + (c % (TILE_W / 8)) * 8);
    cp_async_commit();
  }
  const PrimSetup* const epoch_setups = setups + it.cmd_begin;
  // request the setup record of list entry `rel` into stage[buf]
  auto request_record = [&](uint32_t buf, uint32_t rel) {
    PSX_CHECK(buf < 2 && rel < it.cmd_end - it.cmd_begin);
    if (kRegRec)
      rec_next = __ldg(reinterpret_cast<const uint2*>(epoch_setups + rel) + lane); // 32 lanes x 8 bytes = one record (a streaming load, ld.global.cs, measured the same)
    else if (kBulkRec)
    {
      if (lane == 0)
      {
        mbar_expect_tx(&sh.mbar[buf], static_cast<uint32_t>(sizeof(PrimSetup)));
        bulk_g2s(sh.stage[buf], epoch_setups + rel, static_cast<uint32_t>(sizeof(PrimSetup)), &sh.mbar[buf]);
      }
    }
    else
    {
      if (lane < SETUP_VECS)
        cp_async16(reinterpret_cast<uint4*>(sh.stage[buf]) + lane, reinterpret_cast<const uint4*>(epoch_setups + rel) + lane);
      cp_async_commit();
    }
  };
  // wait until stage[buf] holds its record (newer_in_flight: one younger request may stay outstanding)
  auto await_record = [&](uint32_t buf, bool newer_in_flight) {
    if (kRegRec)
      reinterpret_cast<uint2*>(sh.stage[buf])[lane] = rec_next;
    else if (kBulkRec)
    {
      mbar_wait(&sh.mbar[buf], (mb_phase >> buf) & 1u);
      mb_phase ^= 1u << buf;
    }
    else if (newer_in_flight)
      cp_async_wait<1>();
    else
      cp_async_wait<0>();
    __syncwarp();
  };
#pragma unroll 1
  for (;;)
  {
    bin_fill(it, sh.list, tx, ty, tile_cmds, list_n, next_cmd, lane);
    if (list_n == 0)
      break;
    __syncwarp();

    // ---- raster: walk the bin in order ----
    // The 256-byte setup records travel global -> shared asynchronously, the next primitive's while the current one is
    // rasterised: no register holds a record in flight.
    if (!first_requested)
      request_record(0, sh.list[0]);
    first_requested = false;
    bool first_awaited = false;
    if (!loaded)
    {
      if (kBulk)
      {
        bulk_wait_read(); // the previous region's write-back has left the buffer (it was reading it during the binning above)
        __syncwarp();
      }
      // A FillVRAM (not interlaced) or an unmasked upload that covers the whole region overwrites every pixel of it:
      // the region's old contents need not be read (a frame clear would otherwise read all of VRAM first).
      await_record(0, false); // (kCoh: also covers the rows requested on entry)
      first_awaited = true;
      const PrimSetup& first = *reinterpret_cast<const PrimSetup*>(sh.stage[0]);
      const uint32_t op = first.op;
      bool overwritten = false;
      if (op == OP_FILL || (op == OP_UPLOAD && !(first.flags0 & F_CHECK_MASK)))
      {
        // (fill and upload share the x, y, w, h layout)
        const bool interlaced = (op == OP_FILL) && first.fill.interlaced;
        overwritten = !interlaced && (((rx0 - first.fill.x) & (VRAM_W - 1)) + TILE_W <= first.fill.w) &&
                      (((ry0 - first.fill.y) & (VRAM_H - 1)) + RH <= first.fill.h);
      }
      if (kBulk && !overwritten)
      {
        // lane r fetches row r of the region: one 128-byte bulk copy each, counted in by the region's barrier
        if (lane == 0)
          mbar_expect_tx(&sh.mbar[2], RH * TILE_W * 2u);
        __syncwarp();
        if (RH == 32 || lane < RH)
          bulk_g2s(sh.region + lane * REGION_PITCH, vram + (ry0 + lane) * VRAM_W + rx0, TILE_W * 2u, &sh.mbar[2]);
        mbar_wait(&sh.mbar[2], (mb_phase >> 2) & 1u);
        mb_phase ^= 4u;
      }
      loaded = true;
    }
#pragma unroll 1
    for (uint32_t e = 0; e < list_n; e++)
    {
      if (kRegRec)
      {
        // the registers hold record e: store it, then start loading record e + 1
        if (!(e == 0 && first_awaited))
          await_record(e & 1u, false);
        if (e + 1 < list_n)
          request_record((e + 1) & 1u, sh.list[e + 1]);
      }
      else
      {
        if (e + 1 < list_n)
          request_record((e + 1) & 1u, sh.list[e + 1]);
        if (!(e == 0 && first_awaited))
          await_record(e & 1u, e + 1 < list_n);
      }
      raster_prim_on_region<RH, kCoh>(*reinterpret_cast<const PrimSetup*>(sh.stage[e & 1u]), sh, rx0, ry0, vram, cluts, payload, clut_key, lane);
      __syncwarp();
    }
    list_n = 0;
  }
  if (kCoh && !loaded)
  {
    cp_async_wait<0>(); // nothing to do here after all: let the requested rows land before the next tile reuses the buffer
    __syncwarp();
  }
  if (loaded)
  {
    if (kBulk)
    {
      // the rows were written through the generic proxy: make them visible to the copy engine, then one bulk store per row
      fence_proxy_async_smem();
      __syncwarp();
      if (RH == 32 || lane < RH)
        bulk_s2g(vram + (ry0 + lane) * VRAM_W + rx0, sh.region + lane * REGION_PITCH, TILE_W * 2u);
      bulk_commit(); // (the region buffer may be reused once the engine has read it: bulk_wait_read() before the next region is staged, and before the CTA exits)
    }
    else
    {
#pragma unroll 8
      for (uint32_t r = 0; r < RH; r++)
        gword[r * (VRAM_W / 2)] = sh.region[r * REGION_PITCH + lane];
    }
  }
  if (clut_key_io)
    *clut_key_io = clut_key;
  if (kBulk && mb_phase_io)
    *mb_phase_io = mb_phase;
}

// The binning half of process_epoch_region, run ahead of time (see PreBin).
template<uint32_t RH>
__device__ __forceinline__ PreBin prebin_region(const WaveItem& it, RasterShared<RH>& sh, uint32_t tx, uint32_t ty,
                                                const PrimSetup* __restrict__ setups, const uint32_t* __restrict__ tile_cmds, uint32_t lane)
{
  PreBin pb;
  pb.list_n = 0;
  pb.next_cmd = it.cmd_begin;
  bin_fill(it, sh.list, tx, ty, tile_cmds, pb.list_n, pb.next_cmd, lane);
  if (pb.list_n != 0)
  {
    __syncwarp();
    const uint4* const setup_vecs = reinterpret_cast<const uint4*>(setups + it.cmd_begin);
    if (lane < SETUP_VECS)
      cp_async16(reinterpret_cast<uint4*>(sh.stage[0]) + lane, setup_vecs + static_cast<size_t>(sh.list[0]) * SETUP_VECS + lane);
    cp_async_commit();
  }
  return pb;
}

// One warp per (instance, tile).  No block-level barrier anywhere: the hardware CTA scheduler balances tiles.
#ifndef PSX_RASTER_WARPS
#define PSX_RASTER_WARPS 1 // independent tiles (warps) per CTA; no barrier between them
#endif
#ifndef PSX_RASTER_MIN_CTAS
#define PSX_RASTER_MIN_CTAS 28 // (28 CTAs per SM at 72 registers and the 196 KB shared-memory configuration (60 KB of L1): draw wave 1.606 -> 1.523 ms against 32 CTAs at 64 registers; 27: 1.544, 24 at 80 registers: 1.613, 30: 1.62)
#endif
__global__ void __launch_bounds__(32 * PSX_RASTER_WARPS, PSX_RASTER_MIN_CTAS) raster_kernel(
  const WaveItem* __restrict__ items, const PrimSetup* __restrict__ setups, const uint32_t* __restrict__ tile_cmds,
  uint16_t* __restrict__ vram_pool, const uint16_t* __restrict__ clut_pool, const uint16_t* __restrict__ payload, uint32_t clut_slots)
{
  __shared__ RasterShared<TILE_H> sh_all[PSX_RASTER_WARPS];
  const WaveItem it = items[blockIdx.y];
  if (it.kind != EPOCH_TILED || it.cmd_begin == it.cmd_end)
    return;
  const uint32_t tile = blockIdx.x * PSX_RASTER_WARPS + (threadIdx.x >> 5);
  const uint32_t tx = tile % TILES_X, ty = tile / TILES_X;
  // pad[0] = union of the epoch's tile masks (item_union_kernel): a tile that no primitive of the epoch can touch
  // leaves without scanning the bins (half of the tiles of a frame that draws into one half of VRAM)
  const uint32_t tile_bits = (1u << tx) | (1u << (16 + ty));
  if ((it.pad[0] & tile_bits) != tile_bits)
    return;
  uint16_t* vram = vram_pool + static_cast<size_t>(it.instance) * VRAM_PIXELS;
  const uint16_t* cluts = clut_pool + static_cast<size_t>(it.instance) * clut_slots * CLUT_ENTRIES;
  uint32_t mb_phase = 0;
  init_bulk_barriers(sh_all[threadIdx.x >> 5], threadIdx.x & 31u);
  process_epoch_region<TILE_H, false>(it, sh_all[threadIdx.x >> 5], tx, ty, ty * TILE_H, setups, tile_cmds, vram, cluts, payload,
                                      threadIdx.x & 31u, nullptr, nullptr, &mb_phase);
  bulk_wait_read();
}

// Same work as raster_kernel, other scheduling: exactly as many one-warp CTAs as the GPU holds (SMs x 32), each pulling
// (item, tile) pairs from a global counter until the wave is done — no CTA launch per tile, no idle slot between a
// CTA's exit and its successor's start, dynamic balance between busy and empty tiles.
// Tiles per trip to the counter (kernel argument).  Neighbouring tiles share primitives, texture pages and palettes:
// with small grabs the warps of an SM work on the same neighbourhood at the same time and find that data in L1 / L2
// (draw wave of the batched config 17.5 -> 16.5 ms at 2 instead of 8) — but a wave of trivial tiles (a clear, an
// upload: a primitive or two per epoch) is then bound by the counter itself (1.2 -> 1.4 ms at 2, 2.6 ms at 1).  The
// launcher picks by the wave's commands per epoch.  (Last session of round 2, clear + upload wave of 4096 instances:
// 8 per trip 0.932 ms, 32 per trip and no second look at the counter per trip 0.869 ms.)
constexpr uint32_t LOOP_GRAB_LIGHT = 32, LOOP_GRAB_HEAVY = 2, LOOP_LIGHT_CMDS_PER_ITEM = 8;

__global__ void __launch_bounds__(32, PSX_RASTER_MIN_CTAS) raster_loop_kernel(
  const WaveItem* __restrict__ items, uint32_t n_items, const PrimSetup* __restrict__ setups, const uint32_t* __restrict__ tile_cmds,
  uint16_t* __restrict__ vram_pool, const uint16_t* __restrict__ clut_pool, const uint16_t* __restrict__ payload, uint32_t clut_slots,
  unsigned int* __restrict__ work_counter, uint32_t want)
{
  __shared__ RasterShared<TILE_H> sh;
  uint32_t lane;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(lane)); // (opaque to the compiler: it would otherwise re-read the special register — 20 cycles each — all over the per-primitive code instead of keeping the lane in a register)
  const uint32_t total = n_items * NUM_TILES;
  uint32_t mb_phase = 0;
  init_bulk_barriers(sh, lane);
  uint32_t seen = 0; // the counter as this warp last saw it
  for (;;)
  {
    uint32_t first = 0, grab = 0;
    if (lane == 0)
    {
      // guided self-scheduling: large grabs while there is plenty left, single tiles towards the end of the wave
#if !PSX_GRAB_NOPEEK
      seen = *reinterpret_cast<volatile unsigned int*>(work_counter);
#endif
      // (PSX_GRAB_NOPEEK: the value the warp's previous grab returned stands in for a fresh look at the counter — one
      // round trip to the counter's L2 line per grab instead of two; a wave of trivial tiles is bound by that line)
      const uint32_t left = (total > seen) ? (total - seen) : 0u;
      grab = min(want, max(1u, left / (gridDim.x * 4u)));
      first = atomicAdd(work_counter, grab);
      seen = first + grab;
    }
    first = __shfl_sync(0xFFFFFFFFu, first, 0);
    grab = __shfl_sync(0xFFFFFFFFu, grab, 0);
    if (first >= total)
      break;
    const uint32_t last = min(first + grab, total);
    for (uint32_t id = first; id < last; id++)
    {
      PSX_CHECK(id / NUM_TILES < n_items);
      const WaveItem it = items[id / NUM_TILES];
      if (it.kind != EPOCH_TILED || it.cmd_begin == it.cmd_end)
        continue;
      const uint32_t tile = id % NUM_TILES;
      const uint32_t tx = tile % TILES_X, ty = tile / TILES_X;
      const uint32_t tile_bits = (1u << tx) | (1u << (16 + ty));
      if ((it.pad[0] & tile_bits) != tile_bits)
        continue;
      uint16_t* vram = vram_pool + static_cast<size_t>(it.instance) * VRAM_PIXELS;
      const uint16_t* cluts = clut_pool + static_cast<size_t>(it.instance) * clut_slots * CLUT_ENTRIES;
      process_epoch_region<TILE_H, false>(it, sh, tx, ty, ty * TILE_H, setups, tile_cmds, vram, cluts, payload, lane, nullptr, nullptr, &mb_phase);
      __syncwarp();
    }
  }
  bulk_wait_read();
}

// Union of the tile masks of every epoch (one warp per wave item), stored in WaveItem::pad[0].
__global__ void __launch_bounds__(128) item_union_kernel(WaveItem* __restrict__ items, uint32_t n_items, const BinInfo* __restrict__ bins)
{
  const uint32_t idx = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
  if (idx >= n_items)
    return;
  const uint32_t b = items[idx].cmd_begin, e = items[idx].cmd_end;
  uint32_t m = 0;
  for (uint32_t i = b + lane; i < e; i += 32)
    m |= bins[i].tilemask;
  m = __reduce_or_sync(0xFFFFFFFFu, m);
  if (lane == 0)
    items[idx].pad[0] = m;
}

// ------------------------------------------------------------------------------------------------ persistent (latency) mode
// A handful of instances cannot fill the GPU and pays a kernel launch per epoch (a render-to-texture heavy frame has
// thousands of epochs).  This cooperative kernel stays resident for a whole flush: 256 CTAs per instance, each of 4
// warps owning the four 64x8 strips of one tile (4x the parallelism of raster_kernel for a single VRAM), with a
// grid-wide barrier instead of a launch between epochs.  VRAM and CLUT-slot reads bypass L1 (kCoh) because other CTAs
// rewrite them between epochs.
constexpr uint32_t PERSIST_RH = 8;
constexpr uint32_t PERSIST_WARPS = TILE_H / PERSIST_RH; // 4
constexpr uint32_t PERSIST_THREADS = PERSIST_WARPS * 32;

struct PersistWave
{
  uint32_t item_begin, item_count;
  uint32_t has_clut;
  uint32_t pad;
};

// Grid-wide barrier for the (cooperatively launched, hence co-resident) persistent kernel: one monotonic counter, one
// atomic per CTA per barrier, thread 0 spins.  The fences make this CTA's VRAM stores visible before it arrives and
// order the next epoch's (L1-bypassing) loads after the release.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int n_ctas, unsigned int& generation)
{
  __syncthreads();
  if (threadIdx.x == 0)
  {
    generation++;
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = generation * n_ctas;
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target)
    {
    }
    PSX_CHECK(*reinterpret_cast<volatile unsigned int*>(counter) - target < n_ctas); // nobody is a whole generation ahead
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(PERSIST_THREADS, 7) raster_persistent_kernel(
  const WaveItem* __restrict__ items, const PersistWave* __restrict__ waves, uint32_t n_waves, const uint32_t* __restrict__ instances,
  const PrimSetup* __restrict__ setups, const uint32_t* __restrict__ tile_cmds, const ClutOp* __restrict__ clut_ops, uint16_t* vram_pool,
  uint16_t* clut_pool, const uint16_t* __restrict__ payload, uint32_t clut_slots, unsigned int* barrier_counter)
{
  __shared__ RasterShared<PERSIST_RH> sh_all[PERSIST_WARPS];
  __shared__ uint32_t row_count[VRAM_H + 1];
  const unsigned int n_ctas = gridDim.x * gridDim.y;
  unsigned int generation = 0;
  const uint32_t instance = instances[blockIdx.y];
  uint16_t* vram = vram_pool + static_cast<size_t>(instance) * VRAM_PIXELS;
  uint16_t* cluts = clut_pool + static_cast<size_t>(instance) * clut_slots * CLUT_ENTRIES;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t tx = blockIdx.x % TILES_X, ty = blockIdx.x / TILES_X;

  for (uint32_t k = 0; k < n_waves; k++)
  {
    const PersistWave wv = waves[k];
    // this instance's item of the wave, if it has an epoch k (a wave holds at most a few items in this mode)
    WaveItem it;
    bool mine = false;
    for (uint32_t i = 0; i < wv.item_count && !mine; i++)
    {
      it = items[wv.item_begin + i];
      mine = it.instance == instance;
    }
    if (wv.has_clut) // uniform over the grid
    {
      if (mine && blockIdx.x == 0)
      {
        for (uint32_t o = it.clut_begin; o < it.clut_end; o++)
        {
          const ClutOp op = clut_ops[o];
          for (uint32_t t = threadIdx.x; t < (op.is8bit ? 256u : 16u); t += PERSIST_THREADS)
            cluts[op.dst_slot * CLUT_ENTRIES + t] = __ldcg(vram + op.y * VRAM_W + ((op.x + t) & (VRAM_W - 1)));
        }
      }
      grid_barrier(barrier_counter, n_ctas, generation);
    }
    if (mine && it.cmd_end > it.cmd_begin)
    {
      if (it.kind == EPOCH_TILED)
      {
        process_epoch_region<PERSIST_RH, true>(it, sh_all[warp], tx, ty, ty * TILE_H + warp * PERSIST_RH, setups, tile_cmds, vram, cluts,
                                               payload, lane);
      }
      else if (blockIdx.x == 0)
      {
        __shared__ alignas(16) uint32_t stage[SETUP_WORDS];
        if (threadIdx.x < SETUP_WORDS)
          stage[threadIdx.x] = reinterpret_cast<const uint32_t*>(setups + it.cmd_begin)[threadIdx.x];
        __syncthreads();
        const PrimSetup& s = *reinterpret_cast<const PrimSetup*>(stage);
        if (s.op != OP_NOP)
        {
          if (it.kind == EPOCH_SELF_SAMPLING)
            serial_self_sampling<PERSIST_THREADS>(s, vram, cluts);
          else if (it.kind == EPOCH_COPY_OVERLAP)
            serial_copy<PERSIST_THREADS>(s, vram);
          else
            serial_upload_masked<PERSIST_THREADS>(s, vram, payload, row_count);
        }
        __syncthreads();
      }
    }
    grid_barrier(barrier_counter, n_ctas, generation);
  }
}

// ------------------------------------------------------------------------------------------------ instance mode
// Streams that the hazard splitter cuts into many short epochs (render-to-texture feedback, transfers between draws)
// pay one kernel launch — or one grid-wide barrier — per epoch in the modes above, and the work between two cuts is a
// handful of primitives.  Here ONE CTA owns one instance for a whole flush: it walks the instance's epochs in order,
// its warps pull the epoch's tiles from a shared-memory counter, and what separates two epochs is a __syncthreads()
// (which also orders the CTA's global stores of epoch k before its loads of epoch k + 1: VRAM and CLUT-slot reads go
// to L2, kCoh).  No cross-CTA synchronisation exists: instances are independent.  WARPS = 8 for batches (four CTAs
// per SM), 32 for a handful of instances (one CTA per SM, 1024 threads on the serial epochs).
constexpr uint32_t INST_EPOCH_WINDOW = 64; // epochs whose descriptors are staged in shared memory at a time
constexpr uint32_t INST_RUN_MAX = 16;      // tiled epochs that may be in flight together (tile-granular dependencies)
constexpr uint32_t EPOCH_SET_WORDS = 16;   // per epoch: 256 bits of written tiles + 256 bits of read tiles (epoch_tiles_kernel)

struct InstanceShared
{
  uint32_t row_count[VRAM_H + 1];
  alignas(16) uint32_t stage[SETUP_WORDS];
  Epoch epochs[INST_EPOCH_WINDOW];
  uint32_t tile_sets[INST_EPOCH_WINDOW][EPOCH_SET_WORDS]; // [0..7] tiles written, [8..15] tiles read
  uint32_t tile_total[INST_EPOCH_WINDOW]; // tiles written by each epoch
  uint32_t next_tile[2];
};
// raster_instance_dag_kernel only: tile-granular dependencies inside a run of tiled epochs — per (epoch of the run, tile)
// how many writer items / reader signals must have completed before the epoch's item on that tile may start, and the
// completion counters themselves
struct InstanceDagShared : InstanceShared
{
  uint8_t need_w[INST_RUN_MAX][NUM_TILES];
  uint16_t need_r[INST_RUN_MAX][NUM_TILES];
  uint32_t done_w[NUM_TILES], done_r[NUM_TILES];
  uint32_t run_item_base[INST_RUN_MAX + 1];
  uint32_t run_next;
};

// Which tiles does each epoch of each instance touch?  One warp per epoch, lane L looks after tiles 8 L .. 8 L + 7: the
// OR of their tile_cmds words over the epoch's command groups (first and last group trimmed to the epoch) — exact at the
// binning's own granularity, where the union of the commands' column and row masks would name every tile of the
// drawing area.  Eight words (256 bits) per epoch, for raster_instance_kernel.
__global__ void __launch_bounds__(128) epoch_tiles_kernel(const InstWork* __restrict__ work, const Epoch* __restrict__ epochs,
                                                          const uint32_t* __restrict__ tile_cmds, const BinInfo* __restrict__ bins,
                                                          uint32_t* __restrict__ tile_sets, uint32_t want_reads)
{
  const InstWork w = work[blockIdx.x];
  const uint32_t lane = threadIdx.x & 31u;
  for (uint32_t k = threadIdx.x >> 5; k < w.n_epochs; k += 4)
  {
    const Epoch ep = epochs[w.epoch_first + k];
    uint32_t byte = 0; // bit j: tile 8 * lane + j is touched
    if (ep.kind == EPOCH_TILED && ep.cmd_end > ep.cmd_begin)
    {
      const uint32_t b = w.cmd_base + ep.cmd_begin, e = w.cmd_base + ep.cmd_end;
      for (uint32_t g = b >> 5; g <= (e - 1) >> 5; g++)
      {
        uint32_t keep = 0xFFFFFFFFu;
        if ((g << 5) < b)
          keep &= 0xFFFFFFFFu << (b - (g << 5));
        if ((g << 5) + 32u > e)
          keep &= (1u << (e - (g << 5))) - 1u;
        const uint4* p = reinterpret_cast<const uint4*>(tile_cmds + static_cast<size_t>(g) * NUM_TILES + lane * 8);
        const uint4 a = __ldg(p), c = __ldg(p + 1);
        byte |= ((a.x & keep) ? 1u : 0u) | ((a.y & keep) ? 2u : 0u) | ((a.z & keep) ? 4u : 0u) | ((a.w & keep) ? 8u : 0u) |
                ((c.x & keep) ? 16u : 0u) | ((c.y & keep) ? 32u : 0u) | ((c.z & keep) ? 64u : 0u) | ((c.w & keep) ? 128u : 0u);
      }
    }
    // ... and the tiles the epoch's commands may READ (texture footprints, copy sources: BinInfo::readmask, a rectangle
    // of tiles per command): lane L holds tile row L / 2, columns 8 (L & 1) .. + 7
    uint32_t rbyte = 0;
    if (want_reads && ep.cmd_end > ep.cmd_begin) // (only the kernel with tile-granular dependencies uses the read sets)
    {
      for (uint32_t i = w.cmd_base + ep.cmd_begin; i < w.cmd_base + ep.cmd_end; i++)
      {
        const uint32_t rm = __ldg(&bins[i].readmask);
        if ((rm >> (16u + (lane >> 1))) & 1u)
          rbyte |= (rm >> (8u * (lane & 1u))) & 0xFFu;
      }
    }
    // four lanes make one word: eight words of written tiles, then eight words of read tiles per epoch
    uint32_t word = byte << (8u * (lane & 3u)), rword = rbyte << (8u * (lane & 3u));
    word |= __shfl_xor_sync(0xFFFFFFFFu, word, 1);
    word |= __shfl_xor_sync(0xFFFFFFFFu, word, 2);
    rword |= __shfl_xor_sync(0xFFFFFFFFu, rword, 1);
    rword |= __shfl_xor_sync(0xFFFFFFFFu, rword, 2);
    if ((lane & 3u) == 0)
    {
      tile_sets[static_cast<size_t>(w.epoch_first + k) * EPOCH_SET_WORDS + (lane >> 2)] = word;
      tile_sets[static_cast<size_t>(w.epoch_first + k) * EPOCH_SET_WORDS + 8 + (lane >> 2)] = rword;
    }
  }
}

// The ord-th tile (in tile order) of a 256-bit tile set held one word per lane in lanes 0-7 (`mine`; `before` / `cnt` =
// exclusive prefix / popcount of the lane's word).
__device__ __forceinline__ uint32_t nth_tile(uint32_t mine, uint32_t before, uint32_t cnt, uint32_t ord, uint32_t lane)
{
  const bool here = lane < 8 && ord >= before && ord < before + cnt;
  const uint32_t src = static_cast<uint32_t>(__ffs(static_cast<int>(__ballot_sync(0xFFFFFFFFu, here)))) - 1u;
  return __shfl_sync(0xFFFFFFFFu, lane * 32u + __fns(mine, 0, static_cast<int>(ord - before) + 1), src);
}
__device__ __forceinline__ uint32_t tile_set_prefix(uint32_t mine, uint32_t lane, uint32_t& before)
{
  const uint32_t cnt = __popc(mine);
  before = cnt;
#pragma unroll
  for (uint32_t o = 1; o < 8; o <<= 1)
  {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, before, o);
    if (lane >= o)
      before += t;
  }
  const uint32_t total = __shfl_sync(0xFFFFFFFFu, before, 7);
  before -= cnt;
  return total;
}

template<uint32_t WARPS, uint32_t MIN_CTAS>
__global__ void __launch_bounds__(32 * WARPS, MIN_CTAS) raster_instance_kernel(
  const InstWork* __restrict__ work, const Epoch* __restrict__ epochs, const uint32_t* __restrict__ tile_sets,
  const PrimSetup* __restrict__ setups, const uint32_t* __restrict__ tile_cmds, const ClutOp* __restrict__ clut_ops, uint16_t* vram_pool,
  uint16_t* clut_pool, const uint16_t* __restrict__ payload, uint32_t clut_slots, uint32_t /*run_max*/)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  RasterShared<TILE_H>* sh_all = reinterpret_cast<RasterShared<TILE_H>*>(dyn_smem);
  InstanceShared& cs = *reinterpret_cast<InstanceShared*>(dyn_smem + sizeof(RasterShared<TILE_H>) * WARPS);
  constexpr uint32_t NT = 32 * WARPS;
  const InstWork w = work[blockIdx.x];
  uint16_t* vram = vram_pool + static_cast<size_t>(w.instance) * VRAM_PIXELS;
  uint16_t* cluts = clut_pool + static_cast<size_t>(w.instance) * clut_slots * CLUT_ENTRIES;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  if (threadIdx.x < 2)
    cs.next_tile[threadIdx.x] = 0;

  for (uint32_t k = 0; k < w.n_epochs; k++)
  {
    const uint32_t slot = k % INST_EPOCH_WINDOW;
    if (slot == 0)
    {
      // the next window of epoch descriptors and tile-mask unions: they do not depend on anything this kernel writes
      const uint32_t n = min(INST_EPOCH_WINDOW, w.n_epochs - k);
      const uint32_t* src = reinterpret_cast<const uint32_t*>(epochs + w.epoch_first + k);
      for (uint32_t i = threadIdx.x; i < n * (sizeof(Epoch) / 4); i += NT)
        reinterpret_cast<uint32_t*>(cs.epochs)[i] = __ldg(src + i);
      for (uint32_t i = threadIdx.x; i < n * EPOCH_SET_WORDS; i += NT)
        (&cs.tile_sets[0][0])[i] = __ldg(tile_sets + static_cast<size_t>(w.epoch_first + k) * EPOCH_SET_WORDS + i);
      __syncthreads();
    }
    const Epoch ep = cs.epochs[slot];
    WaveItem it;
    it.instance = w.instance;
    it.cmd_begin = w.cmd_base + ep.cmd_begin;
    it.cmd_end = w.cmd_base + ep.cmd_end;
    it.kind = ep.kind;
    it.clut_begin = w.clut_base + ep.clut_begin;
    it.clut_end = w.clut_base + ep.clut_end;
    if (it.clut_end > it.clut_begin) // uniform over the CTA
    {
      for (uint32_t o = it.clut_begin; o < it.clut_end; o++)
      {
        const ClutOp op = clut_ops[o];
        for (uint32_t t = threadIdx.x; t < (op.is8bit ? 256u : 16u); t += NT)
          cluts[op.dst_slot * CLUT_ENTRIES + t] = __ldcg(vram + op.y * VRAM_W + ((op.x + t) & (VRAM_W - 1)));
      }
      __syncthreads();
    }
    if (it.cmd_end > it.cmd_begin)
    {
      if (it.kind == EPOCH_TILED)
      {
        // the tiles the epoch touches (epoch_tiles_kernel), pulled from a counter; the counters of even and odd epochs
        // alternate, so that resetting one needs no barrier of its own
        if (threadIdx.x == 0)
          cs.next_tile[(k + 1) & 1u] = 0;
        const uint32_t mine = (lane < 8) ? cs.tile_sets[slot][lane] : 0u; // lane L: tiles 32 L .. 32 L + 31
        const uint32_t cnt = __popc(mine);
        uint32_t before = cnt; // exclusive prefix over lanes 0..7
#pragma unroll
        for (uint32_t o = 1; o < 8; o <<= 1)
        {
          const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, before, o);
          if (lane >= o)
            before += t;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, before, 7);
        before -= cnt;
        for (;;)
        {
          uint32_t j = 0;
          if (lane == 0)
            j = atomicAdd(&cs.next_tile[k & 1u], 1u);
          j = __shfl_sync(0xFFFFFFFFu, j, 0);
          if (j >= total)
            break;
          // the j-th set bit of the 256: the lane whose range holds it answers
          const bool here = lane < 8 && j >= before && j < before + cnt;
          const uint32_t src = static_cast<uint32_t>(__ffs(static_cast<int>(__ballot_sync(0xFFFFFFFFu, here)))) - 1u;
          const uint32_t tile = __shfl_sync(0xFFFFFFFFu, lane * 32u + __fns(mine, 0, static_cast<int>(j - before) + 1), src);
          const uint32_t tx = tile % TILES_X, ty = tile / TILES_X;
          process_epoch_region<TILE_H, true>(it, sh_all[warp], tx, ty, ty * TILE_H, setups, tile_cmds, vram, cluts, payload, lane);
          __syncwarp();
        }
      }
      else
      {
        if (threadIdx.x == 0)
          cs.next_tile[(k + 1) & 1u] = 0;
        if (threadIdx.x < SETUP_WORDS)
          cs.stage[threadIdx.x] = reinterpret_cast<const uint32_t*>(setups + it.cmd_begin)[threadIdx.x];
        __syncthreads();
        const PrimSetup& s = *reinterpret_cast<const PrimSetup*>(cs.stage);
        if (s.op != OP_NOP)
        {
          if (it.kind == EPOCH_SELF_SAMPLING)
            serial_self_sampling<NT>(s, vram, cluts);
          else if (it.kind == EPOCH_COPY_OVERLAP)
            serial_copy<NT>(s, vram);
          else
            serial_upload_masked<NT>(s, vram, payload, cs.row_count);
        }
      }
    }
    else if (threadIdx.x == 0)
      cs.next_tile[(k + 1) & 1u] = 0;
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ slot mode
// raster_instance_kernel leaves most of an SM idle on streams cut into many short epochs: an epoch of a few dozen
// primitives touches a few dozen tiles, the warp that owns the busiest tile walks its bin alone while the other warps
// wait at the barrier, and one warp issues an instruction every 5-10 cycles.  More resident warps are not to be had (64
// registers, 6 KB of shared memory each), so here the CTA's 32 warps serve SEVERAL instances at a time: the CTA holds
// n_slots instance slots, every slot walks its own instance's epochs in order, and a warp that finds nothing left to
// claim in one slot's epoch claims a tile of another slot's.  What separates two epochs of a slot is that slot's own
// counter (tiles done == tiles of the epoch), not a barrier of the CTA: the warp that finishes the epoch's last tile
// runs the transition (CLUT snapshot ops, next epoch's tile set) by itself while the others keep rasterising the other
// slots.  Finished instances are replaced from a global counter.  Only the serial epoch kinds (self-sampling, aliasing
// copies, mask-checked uploads: whole-CTA code) make the CTA meet: they are flagged, every warp leaves the work loop
// after its current tile, the CTA runs the flagged slots' serial epochs with all of its threads and resumes.
// Measured on 2048 instances of the hazard variant of the batched workload (generator config 6): barrier per epoch
// 18.5 ms, 1 slot 21.7 (the claim protocol alone costs more than __syncthreads), 2 slots 16.6, 3 slots 16.5, 4 slots 17.5
// (444+ instances in flight no longer fit L2: 8 + 5 GB of DRAM traffic per 512 instances, long-scoreboard the top
// stall).  Tried and dropped: L2 prefetch of the tiles one round ahead of their claim and at the epoch's start (18-20 ms:
// the lookups and the prefetches cost more than the misses they save), the epoch's bin words staged in shared memory by
// the transition (21.6 ms: 16 KB more shared memory put the CTA into the 228 KB configuration, no L1 left).
// Ordering inside a slot: tile write-back (st.global) -> __threadfence_block -> atomic on `done` (shared) -> ... ->
// publishing store to `next` -> the claimant's atomic on `next` -> __threadfence_block -> its loads (ld.global.cg /
// cp.async.cg: L2, never a stale L1 line).  All parties are warps of one CTA, so CTA scope is enough.
constexpr uint32_t MULTI_SLOTS_MAX = 4;
constexpr uint32_t SLOT_IDLE = 0, SLOT_TILED = 1, SLOT_SERIAL = 2, SLOT_BUSY = 3;
constexpr uint32_t SLOT_CLOSED = 0x80000000u; // `next` while the slot has no tiles to hand out (a claim lands above every total)
struct MultiSlot
{
  WaveItem it;        // the current epoch (global command / CLUT-op indices)
  uint32_t tiles[8];  // its tile set
  uint32_t before[8]; // exclusive prefix of the words' popcounts
  uint32_t total;     // tiles in the set
  uint32_t next;      // next ordinal to hand out
  uint32_t done;      // tiles finished
  uint32_t state;
  uint32_t epoch, n_epochs, epoch_first, cmd_base, clut_base, instance;
  uint32_t pad[2];
};
struct MultiShared
{
  uint32_t row_count[VRAM_H + 1];
  alignas(16) uint32_t stage[SETUP_WORDS];
  MultiSlot slot[MULTI_SLOTS_MAX];
  uint32_t serial_pending; // slots waiting for the CTA to run their serial epoch
  uint32_t active;         // slots that still hold an instance
};

// One warp moves slot `s` to its next epoch with work (or to its next instance, or to rest).  On entry nobody can claim
// from the slot (next >= SLOT_CLOSED) and every tile of its previous epoch has been written back.
__device__ __noinline__ void multi_advance_slot(MultiShared& cs, uint32_t s, bool fresh, const InstWork* __restrict__ work, uint32_t n_instances,
                                                const Epoch* __restrict__ epochs, const uint32_t* __restrict__ tile_sets,
                                                const ClutOp* __restrict__ clut_ops, uint16_t* vram_pool, uint16_t* clut_pool, uint32_t clut_slots,
                                                unsigned int* inst_counter, uint32_t lane)
{
  MultiSlot& S = cs.slot[s];
  uint32_t k = fresh ? 0u : S.epoch + 1u;
  uint32_t n_epochs = fresh ? 0u : S.n_epochs;
  for (;;)
  {
    if (k >= n_epochs)
    {
      uint32_t idx = 0;
      if (lane == 0)
        idx = atomicAdd(inst_counter, 1u);
      idx = __shfl_sync(0xFFFFFFFFu, idx, 0);
      if (idx >= n_instances)
      {
        if (lane == 0)
        {
          S.state = SLOT_IDLE;
          __threadfence_block();
          atomicSub(&cs.active, 1u);
        }
        __syncwarp();
        return;
      }
      const InstWork w = work[idx];
      if (lane == 0)
      {
        S.instance = w.instance;
        S.cmd_base = w.cmd_base;
        S.clut_base = w.clut_base;
        S.epoch_first = w.epoch_first;
        S.n_epochs = w.n_epochs;
      }
      __syncwarp();
      n_epochs = w.n_epochs;
      k = 0;
      if (n_epochs == 0)
        continue;
    }
    const uint32_t epoch_first = S.epoch_first, instance = S.instance;
    PSX_CHECK(k < n_epochs && instance < g_dbg.n_instances);
    const Epoch ep = epochs[epoch_first + k];
    PSX_CHECK(ep.cmd_begin <= ep.cmd_end && S.cmd_base + ep.cmd_end <= g_dbg.n_cmds && ep.clut_begin <= ep.clut_end &&
              S.clut_base + ep.clut_end <= g_dbg.n_clutops);
    uint16_t* vram = vram_pool + static_cast<size_t>(instance) * VRAM_PIXELS;
    uint16_t* cluts = clut_pool + static_cast<size_t>(instance) * clut_slots * CLUT_ENTRIES;
    WaveItem it;
    it.instance = instance;
    it.cmd_begin = S.cmd_base + ep.cmd_begin;
    it.cmd_end = S.cmd_base + ep.cmd_end;
    it.kind = ep.kind;
    it.clut_begin = S.clut_base + ep.clut_begin;
    it.clut_end = S.clut_base + ep.clut_end;
    it.pad[0] = it.pad[1] = 0;
    for (uint32_t o = it.clut_begin; o < it.clut_end; o++)
    {
      const ClutOp op = clut_ops[o];
      for (uint32_t t = lane; t < (op.is8bit ? 256u : 16u); t += 32)
        cluts[op.dst_slot * CLUT_ENTRIES + t] = __ldcg(vram + op.y * VRAM_W + ((op.x + t) & (VRAM_W - 1)));
    }
    if (it.cmd_end > it.cmd_begin)
    {
      if (ep.kind != EPOCH_TILED)
      {
        if (lane == 0)
        {
          S.it = it;
          S.epoch = k;
          S.state = SLOT_SERIAL;
          __threadfence_block();
          atomicAdd(&cs.serial_pending, 1u);
        }
        __syncwarp();
        return;
      }
      const uint32_t mine = (lane < 8) ? __ldg(tile_sets + static_cast<size_t>(epoch_first + k) * EPOCH_SET_WORDS + lane) : 0u;
      uint32_t before;
      const uint32_t total = tile_set_prefix(mine, lane, before);
      if (total != 0)
      {
        if (lane < 8)
        {
          S.tiles[lane] = mine;
          S.before[lane] = before;
        }
        if (lane == 0)
        {
          S.it = it;
          S.epoch = k;
          S.total = total;
          S.done = 0;
        }
        __syncwarp();
        __threadfence_block(); // the descriptor and the CLUT slots written above, before the slot opens
        if (lane == 0)
        {
          S.state = SLOT_TILED;
          atomicExch(&S.next, 0u);
        }
        __syncwarp();
        return;
      }
    }
    k++;
  }
}

template<uint32_t WARPS>
__global__ void __launch_bounds__(32 * WARPS, 1) raster_multi_kernel(
  const InstWork* __restrict__ work, uint32_t n_instances, const Epoch* __restrict__ epochs, const uint32_t* __restrict__ tile_sets,
  const PrimSetup* __restrict__ setups, const uint32_t* __restrict__ tile_cmds, const ClutOp* __restrict__ clut_ops, uint16_t* vram_pool,
  uint16_t* clut_pool, const uint16_t* __restrict__ payload, uint32_t clut_slots, uint32_t n_slots, unsigned int* inst_counter)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  RasterShared<TILE_H>* sh_all = reinterpret_cast<RasterShared<TILE_H>*>(dyn_smem);
  MultiShared& cs = *reinterpret_cast<MultiShared*>(dyn_smem + sizeof(RasterShared<TILE_H>) * WARPS);
  constexpr uint32_t NT = 32 * WARPS;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  if (threadIdx.x < MULTI_SLOTS_MAX)
  {
    cs.slot[threadIdx.x].state = SLOT_BUSY;
    cs.slot[threadIdx.x].next = SLOT_CLOSED;
    cs.slot[threadIdx.x].total = 0;
    cs.slot[threadIdx.x].done = 0;
  }
  if (threadIdx.x == 0)
  {
    cs.serial_pending = 0;
    cs.active = n_slots;
  }
  __syncthreads();
  if (warp < n_slots)
    multi_advance_slot(cs, warp, true, work, n_instances, epochs, tile_sets, clut_ops, vram_pool, clut_pool, clut_slots, inst_counter, lane);
  __syncthreads();

  const uint32_t first_slot = warp % n_slots;
  volatile uint32_t* const v_pending = &cs.serial_pending;
  volatile uint32_t* const v_active = &cs.active;
  for (;;)
  {
    // ---- work loop: claim a tile of any slot's current epoch ----
    uint32_t idle_spins = 0;
    for (;;)
    {
      uint32_t claim = 0xFFFFFFFFu; // slot | ordinal << 8
      if (lane == 0)
      {
        if (*v_pending == 0 && *v_active != 0)
        {
          uint32_t s = first_slot; // (warps start at different slots: fewer collisions on one counter)
          for (uint32_t i = 0; i < n_slots; i++, s = (s + 1u == n_slots) ? 0u : s + 1u)
          {
            volatile MultiSlot& S = cs.slot[s];
            if (S.state != SLOT_TILED || S.next >= S.total)
              continue;
            const uint32_t j = atomicAdd(const_cast<uint32_t*>(&S.next), 1u);
            __threadfence_block();
            if (j < S.total) // (read after the claim: a claim that lands in the NEXT epoch of the slot is a valid claim of that epoch)
            {
              claim = s | (j << 8);
              break;
            }
          }
          if (claim == 0xFFFFFFFFu)
            claim = 0xFFFFFFFEu; // nothing to claim right now
        }
      }
      claim = __shfl_sync(0xFFFFFFFFu, claim, 0);
      __syncwarp();
      if (claim == 0xFFFFFFFFu)
        break; // a serial epoch is waiting for the CTA, or every slot is at rest
      if (claim == 0xFFFFFFFEu)
      {
        __nanosleep(idle_spins < 8 ? 64u : 256u);
        idle_spins++;
        continue;
      }
      idle_spins = 0;
      const uint32_t s = claim & 0xFFu, j = claim >> 8;
      MultiSlot& S = cs.slot[s];
      const WaveItem it = S.it;
      const uint32_t total = S.total;
      PSX_CHECK(s < n_slots && j < total && total <= NUM_TILES && S.state == SLOT_TILED && it.kind == EPOCH_TILED);
      PSX_CHECK(it.instance < g_dbg.n_instances && it.cmd_begin < it.cmd_end && it.cmd_end <= g_dbg.n_cmds);
      const uint32_t mine = (lane < 8) ? S.tiles[lane] : 0u, before = (lane < 8) ? S.before[lane] : 0u;
      const uint32_t tile = nth_tile(mine, before, __popc(mine), j, lane);
      const uint32_t tx = tile % TILES_X, ty = tile / TILES_X;
      PSX_CHECK(tile < NUM_TILES && ((mine >> (tile & 31u)) & 1u || (lane != (tile >> 5))));
      uint16_t* vram = vram_pool + static_cast<size_t>(it.instance) * VRAM_PIXELS;
      const uint16_t* cluts = clut_pool + static_cast<size_t>(it.instance) * clut_slots * CLUT_ENTRIES;
      process_epoch_region<TILE_H, PSX_SLOT_COH>(it, sh_all[warp], tx, ty, ty * TILE_H, setups, tile_cmds, vram, cluts, payload, lane);
      __syncwarp();
      __threadfence_block(); // the tile's write-back, before it counts as done
      uint32_t last = 0;
      if (lane == 0)
      {
        const uint32_t done_now = atomicAdd(&S.done, 1u) + 1u;
        PSX_CHECK(done_now <= total);
        last = (done_now == total) ? 1u : 0u;
        if (last)
        {
          S.state = SLOT_BUSY;
          atomicExch(&S.next, SLOT_CLOSED);
        }
      }
      last = __shfl_sync(0xFFFFFFFFu, last, 0);
      if (last)
      {
        __threadfence_block();
        multi_advance_slot(cs, s, false, work, n_instances, epochs, tile_sets, clut_ops, vram_pool, clut_pool, clut_slots, inst_counter, lane);
      }
    }
    __syncthreads(); // every warp is out of the work loop: no claim, no transition in flight
    if (cs.active == 0)
      break;
    // ---- serial epochs of the flagged slots, all threads ----
    for (uint32_t s = 0; s < n_slots; s++)
    {
      if (cs.slot[s].state != SLOT_SERIAL) // (uniform: nobody writes slot states between the barriers)
        continue;
      const WaveItem it = cs.slot[s].it;
      uint16_t* vram = vram_pool + static_cast<size_t>(it.instance) * VRAM_PIXELS;
      uint16_t* cluts = clut_pool + static_cast<size_t>(it.instance) * clut_slots * CLUT_ENTRIES;
      if (threadIdx.x < SETUP_WORDS)
        cs.stage[threadIdx.x] = reinterpret_cast<const uint32_t*>(setups + it.cmd_begin)[threadIdx.x];
      __syncthreads();
      const PrimSetup& ps = *reinterpret_cast<const PrimSetup*>(cs.stage);
      if (ps.op != OP_NOP)
      {
        if (it.kind == EPOCH_SELF_SAMPLING)
          serial_self_sampling<NT>(ps, vram, cluts);
        else if (it.kind == EPOCH_COPY_OVERLAP)
          serial_copy<NT>(ps, vram);
        else
          serial_upload_masked<NT>(ps, vram, payload, cs.row_count);
      }
      __syncthreads();
    }
    if (threadIdx.x == 0)
      cs.serial_pending = 0;
    __syncthreads();
    if (warp < n_slots && cs.slot[warp].state == SLOT_SERIAL)
      multi_advance_slot(cs, warp, false, work, n_instances, epochs, tile_sets, clut_ops, vram_pool, clut_pool, clut_slots, inst_counter, lane);
    __syncthreads();
  }
}

// The same, with TILE-granular dependencies inside runs of tiled epochs instead of a barrier per epoch (PSXGPU_INST_RUN=n,
// n > 1, selects it).  Measured on the hazard variant of the batched workload (512 instances): barrier stalls per issue
// 6.0 -> 4.1, issue slots 44 % -> 46 % busy, but 13 % more instructions (per-item dependency tests, signals, tile
// lookup) — 5.8 ms against 5.3 ms for the kernel above.  The runs are short there: every CLUT snapshot op and every
// serial epoch ends one (a snapshot op may overwrite a palette slot that items of earlier epochs still use, so it has
// to wait for all of them), which leaves two or three epochs per run.  Kept for streams with long runs of plain
// render-to-texture hazards, and as the measured answer to "why not per-tile dependencies".
template<uint32_t WARPS, uint32_t MIN_CTAS>
__global__ void __launch_bounds__(32 * WARPS, MIN_CTAS) raster_instance_dag_kernel(
  const InstWork* __restrict__ work, const Epoch* __restrict__ epochs, const uint32_t* __restrict__ tile_sets,
  const PrimSetup* __restrict__ setups, const uint32_t* __restrict__ tile_cmds, const ClutOp* __restrict__ clut_ops, uint16_t* vram_pool,
  uint16_t* clut_pool, const uint16_t* __restrict__ payload, uint32_t clut_slots, uint32_t run_max)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  RasterShared<TILE_H>* sh_all = reinterpret_cast<RasterShared<TILE_H>*>(dyn_smem);
  InstanceDagShared& cs = *reinterpret_cast<InstanceDagShared*>(dyn_smem + sizeof(RasterShared<TILE_H>) * WARPS);
  constexpr uint32_t NT = 32 * WARPS;
  const InstWork w = work[blockIdx.x];
  uint16_t* vram = vram_pool + static_cast<size_t>(w.instance) * VRAM_PIXELS;
  uint16_t* cluts = clut_pool + static_cast<size_t>(w.instance) * clut_slots * CLUT_ENTRIES;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  uint32_t clut_key = 0; // the palette this warp has staged; survives until the next CLUT snapshot op

  auto epoch_item = [&](const Epoch& ep) {
    WaveItem it;
    it.instance = w.instance;
    it.cmd_begin = w.cmd_base + ep.cmd_begin;
    it.cmd_end = w.cmd_base + ep.cmd_end;
    it.kind = ep.kind;
    it.clut_begin = w.clut_base + ep.clut_begin;
    it.clut_end = w.clut_base + ep.clut_end;
    return it;
  };

  uint32_t k = 0;
  while (k < w.n_epochs)
  {
    const uint32_t slot = k % INST_EPOCH_WINDOW;
    if (slot == 0)
    {
      // the next window of epoch descriptors and tile sets: they do not depend on anything this kernel writes
      const uint32_t n = min(INST_EPOCH_WINDOW, w.n_epochs - k);
      const uint32_t* src = reinterpret_cast<const uint32_t*>(epochs + w.epoch_first + k);
      for (uint32_t i = threadIdx.x; i < n * (sizeof(Epoch) / 4); i += NT)
        reinterpret_cast<uint32_t*>(cs.epochs)[i] = __ldg(src + i);
      for (uint32_t i = threadIdx.x; i < n * EPOCH_SET_WORDS; i += NT)
        (&cs.tile_sets[0][0])[i] = __ldg(tile_sets + static_cast<size_t>(w.epoch_first + k) * EPOCH_SET_WORDS + i);
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < n; i += NT)
      {
        uint32_t c = 0;
#pragma unroll
        for (uint32_t q = 0; q < 8; q++)
          c += __popc(cs.tile_sets[i][q]);
        cs.tile_total[i] = c;
      }
      __syncthreads();
    }
    const WaveItem it = epoch_item(cs.epochs[slot]);
    if (it.clut_end > it.clut_begin) // uniform over the CTA
    {
      clut_key = 0;
      for (uint32_t o = it.clut_begin; o < it.clut_end; o++)
      {
        const ClutOp op = clut_ops[o];
        for (uint32_t t = threadIdx.x; t < (op.is8bit ? 256u : 16u); t += NT)
          cluts[op.dst_slot * CLUT_ENTRIES + t] = __ldcg(vram + op.y * VRAM_W + ((op.x + t) & (VRAM_W - 1)));
      }
      __syncthreads();
    }
    if (it.cmd_end == it.cmd_begin)
    {
      k++;
      continue;
    }
    if (it.kind != EPOCH_TILED)
    {
      if (threadIdx.x < SETUP_WORDS)
        cs.stage[threadIdx.x] = reinterpret_cast<const uint32_t*>(setups + it.cmd_begin)[threadIdx.x];
      __syncthreads();
      const PrimSetup& s = *reinterpret_cast<const PrimSetup*>(cs.stage);
      if (s.op != OP_NOP)
      {
        if (it.kind == EPOCH_SELF_SAMPLING)
          serial_self_sampling<NT>(s, vram, cluts);
        else if (it.kind == EPOCH_COPY_OVERLAP)
          serial_copy<NT>(s, vram);
        else
          serial_upload_masked<NT>(s, vram, payload, cs.row_count);
      }
      __syncthreads();
      k++;
      continue;
    }

    // ---- a run of tiled epochs: this one and the tiled epochs behind it (same window, no CLUT snapshot op in front of
    // them), executed with TILE-granular dependencies instead of a barrier per epoch.  The splitter cut these epochs
    // apart because somewhere a later command reads or overwrites what an earlier one wrote or read; that concerns a
    // few tiles, and the other tiles of the later epochs need not wait.  An item = (epoch j of the run, tile t written
    // by it), pulled in (epoch, tile) order.  Item (j, t) may start when
    //   * every earlier item on t and on every tile epoch j reads is complete (RAW, WAW): done_w[s] >= need_w[j][s];
    //   * every item of an earlier epoch that reads t is complete (WAR): done_r[t] >= need_r[j][t]
    // (epoch-level read sets from epoch_tiles_kernel: coarser than the splitter's cells, never finer — a superset of the
    // true dependencies).  A finished item writes its tile back, fences, and bumps done_w[t] and done_r[s] for every s
    // its epoch reads.  Items are pulled in an order in which every dependency comes first, so the lowest item in
    // flight never waits: no deadlock.
    uint32_t n_run = 1;
    while (n_run < run_max && slot + n_run < INST_EPOCH_WINDOW && k + n_run < w.n_epochs)
    {
      const Epoch& e2 = cs.epochs[slot + n_run];
      if (e2.kind != EPOCH_TILED || e2.clut_end > e2.clut_begin)
        break;
      n_run++;
    }
    if (threadIdx.x < NUM_TILES)
    {
      const uint32_t t = threadIdx.x, word = t >> 5, bit = t & 31u;
      uint32_t cw = 0, cr = 0;
      for (uint32_t j = 0; j < n_run; j++)
      {
        cs.need_w[j][t] = static_cast<uint8_t>(cw);
        cs.need_r[j][t] = static_cast<uint16_t>(cr);
        cw += (cs.tile_sets[slot + j][word] >> bit) & 1u;
        if ((cs.tile_sets[slot + j][8 + word] >> bit) & 1u)
          cr += cs.tile_total[slot + j];
      }
      cs.done_w[t] = 0;
      cs.done_r[t] = 0;
    }
    if (threadIdx.x == 0)
    {
      uint32_t base = 0;
      for (uint32_t j = 0; j < n_run; j++)
      {
        cs.run_item_base[j] = base;
        base += cs.tile_total[slot + j];
      }
      cs.run_item_base[n_run] = base;
      cs.run_next = 0;
    }
    __syncthreads();
    const uint32_t n_items = cs.run_item_base[n_run];
    for (;;)
    {
      uint32_t g = 0;
      if (lane == 0)
        g = atomicAdd(&cs.run_next, 1u);
      g = __shfl_sync(0xFFFFFFFFu, g, 0);
      if (g >= n_items)
        break;
      uint32_t j = 0;
      while (cs.run_item_base[j + 1] <= g)
        j++;
      const uint32_t mine = (lane < 8) ? cs.tile_sets[slot + j][lane] : 0u; // lane L: tiles 32 L .. 32 L + 31
      uint32_t before;
      tile_set_prefix(mine, lane, before);
      const uint32_t tile = nth_tile(mine, before, __popc(mine), g - cs.run_item_base[j], lane);
      // dependencies: lane L looks after tiles 8 L .. 8 L + 7
      const uint32_t reads8 = (cs.tile_sets[slot + j][8 + (lane >> 2)] >> (8u * (lane & 3u))) & 0xFFu;
      const uint32_t own8 = ((tile >> 3) == lane) ? (1u << (tile & 7u)) : 0u;
      if (j != 0) // (the first epoch of a run waits for nothing)
      {
        for (;;)
        {
          bool ok = true;
          uint32_t m = reads8 | own8;
          while (m)
          {
            const uint32_t b = static_cast<uint32_t>(__ffs(static_cast<int>(m))) - 1u;
            m &= m - 1u;
            const uint32_t sidx = lane * 8u + b;
            ok = ok && *reinterpret_cast<volatile uint32_t*>(&cs.done_w[sidx]) >= cs.need_w[j][sidx];
          }
          if (own8)
            ok = ok && *reinterpret_cast<volatile uint32_t*>(&cs.done_r[tile]) >= cs.need_r[j][tile];
          if (__all_sync(0xFFFFFFFFu, ok))
            break;
          __nanosleep(40);
        }
      }
      process_epoch_region<TILE_H, true>(epoch_item(cs.epochs[slot + j]), sh_all[warp], tile % TILES_X, tile / TILES_X, (tile / TILES_X) * TILE_H,
                                         setups, tile_cmds, vram, cluts, payload, lane, nullptr, &clut_key);
      // the tile is back in global memory: publish it to the CTA, then signal
      __threadfence_block();
      __syncwarp();
      if (lane == 0)
        atomicAdd(&cs.done_w[tile], 1u);
      uint32_t m = reads8;
      while (m)
      {
        const uint32_t b = static_cast<uint32_t>(__ffs(static_cast<int>(m))) - 1u;
        m &= m - 1u;
        atomicAdd(&cs.done_r[lane * 8u + b], 1u);
      }
    }
    __syncthreads();
    k += n_run;
  }
}

// ------------------------------------------------------------------------------------------------ cluster mode
// A single VRAM (or a handful) whose stream is cut into thousands of short epochs: one CTA cannot win (one SM's issue
// rate), and a grid-wide barrier per epoch costs more than the epoch's work (≈7 µs through L2 atomics).  Here one
// thread-block CLUSTER owns an instance: 8 or 16 CTAs of 32 warps on neighbouring SMs, the hardware cluster barrier
// (barrier.cluster, release / acquire at cluster scope, ≈0.2 µs) between epochs, and every touched tile cut into four
// 64x8 strips that are dealt round-robin over the cluster's warps — strips of one tile land on different SMs, so the
// per-epoch latency is that of the busiest strip.  CLUT snapshot ops are dealt over the CTAs; serial epochs run on CTA
// 0 (1024 threads).  VRAM / CLUT-slot reads go to L2 (kCoh), like in the other kernels that outlive an epoch.
__device__ __forceinline__ uint32_t cluster_ctarank()
{
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank()
{
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive()
{
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait()
{
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_barrier()
{
  cluster_arrive();
  cluster_wait();
}

// CopyVRAM whose source and destination alias, by the whole cluster at once.  The reference copies row after row, each
// row read completely before it is written (CopyVRAMImpl, inl:1833-1939): with d = (dy - sy) mod 512, row r reads what
// row r - d wrote wherever the source pixel lies inside the destination rectangle (rows r + 512 - d come later: they
// only overwrite what was already read).  A destination pixel is written at most once, so "was it written" is a
// question about the ORIGINAL mask bit, and what a pixel ends up with is the original value found by following
// (r, c) -> (r - d, c + sx - dx) while the step stays inside the rectangle and lands on a written pixel: a closed form
// without the mask test, a short chain of mask-bit probes with it.  Every thread resolves up to four pixels from the
// original VRAM, the cluster barrier separates the reads from the writes.  Returns false (before any barrier, the same
// answer in every thread of the cluster) when the copy wraps around the columns or is too large for one pass.
template<int NT>
__device__ __forceinline__ bool cluster_copy_fast(const PrimSetup& s, uint16_t* vram, uint32_t rank, uint32_t n_ctas)
{
  const uint32_t mand = (s.flags0 & F_CHECK_MASK) ? 0x8000u : 0u, mor = (s.flags0 & F_SET_MASK) ? 0x8000u : 0u;
  const uint32_t sx = s.copy.sx, sy = s.copy.sy, dx = s.copy.dx, dy = s.copy.dy, w = s.copy.w, h = s.copy.h;
  constexpr uint32_t PER_THREAD = 4;
  const uint32_t total = w * h, n_threads = NT * n_ctas;
  if (sx + w > VRAM_W || dx + w > VRAM_W || h > VRAM_H || total > PER_THREAD * n_threads)
    return false;
  const uint32_t d = (dy - sy) & (VRAM_H - 1);
  const int32_t delta = static_cast<int32_t>(sx) - static_cast<int32_t>(dx);
  uint32_t val[PER_THREAD], at[PER_THREAD];
#pragma unroll
  for (uint32_t k = 0; k < PER_THREAD; k++)
  {
    const uint32_t i = k * n_threads + rank * NT + threadIdx.x;
    val[k] = 0;
    at[k] = 0;
    if (i < total)
    {
      const uint32_t r = i / w, c = i - r * w;
      at[k] = ((dy + r) & (VRAM_H - 1)) * VRAM_W + dx + c;
      if ((__ldcg(vram + at[k]) & mand) == 0)
      {
        uint32_t rc = r;
        int32_t cc = static_cast<int32_t>(c);
        if (d != 0 && !mand)
        {
          uint32_t n = rc / d;
          if (delta > 0)
            n = min(n, (w - 1u - c) / static_cast<uint32_t>(delta));
          else if (delta < 0)
            n = min(n, c / static_cast<uint32_t>(-delta));
          rc -= n * d;
          cc += static_cast<int32_t>(n) * delta;
        }
        else if (d != 0)
        {
          while (rc >= d)
          {
            const int32_t cn = cc + delta;
            if (cn < 0 || cn >= static_cast<int32_t>(w))
              break;
            if (__ldcg(vram + ((dy + rc - d) & (VRAM_H - 1)) * VRAM_W + dx + static_cast<uint32_t>(cn)) & 0x8000u)
              break; // masked: row rc - d left it alone, the source pixel still holds its original value
            rc -= d;
            cc = cn;
          }
        }
        val[k] = __ldcg(vram + ((sy + rc) & (VRAM_H - 1)) * VRAM_W + sx + static_cast<uint32_t>(cc)) | mor | 0x10000u;
      }
    }
  }
  cluster_barrier();
#pragma unroll
  for (uint32_t k = 0; k < PER_THREAD; k++)
    if (val[k] & 0x10000u)
      vram[at[k]] = static_cast<uint16_t>(val[k]);
  return true;
}

// RH = rows per strip (8 or 4: 4 or 8 strips per tile; a multiple of 4 keeps the dither phase of the word loop)
template<uint32_t RH, uint32_t CLUSTER_WARPS>
__global__ void __launch_bounds__(32 * CLUSTER_WARPS, 1) raster_cluster_kernel(
  const InstWork* __restrict__ work, const Epoch* __restrict__ epochs, const uint32_t* __restrict__ tile_sets,
  const PrimSetup* __restrict__ setups, const uint32_t* __restrict__ tile_cmds, const ClutOp* __restrict__ clut_ops, uint16_t* vram_pool,
  uint16_t* clut_pool, const uint16_t* __restrict__ payload, uint32_t clut_slots, uint32_t debug_skip)
{
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  RasterShared<RH>* sh_all = reinterpret_cast<RasterShared<RH>*>(dyn_smem);
  InstanceShared& cs = *reinterpret_cast<InstanceShared*>(dyn_smem + sizeof(RasterShared<RH>) * CLUSTER_WARPS);
  constexpr uint32_t NT = 32 * CLUSTER_WARPS;
  constexpr uint32_t CLUSTER_STRIPS = TILE_H / RH;
  const uint32_t n_ctas = cluster_nctarank(), rank = cluster_ctarank();
  const InstWork w = work[blockIdx.x / n_ctas];
  uint16_t* vram = vram_pool + static_cast<size_t>(w.instance) * VRAM_PIXELS;
  uint16_t* cluts = clut_pool + static_cast<size_t>(w.instance) * clut_slots * CLUT_ENTRIES;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t gw = warp * n_ctas + rank, n_gw = CLUSTER_WARPS * n_ctas; // consecutive strips -> different SMs
  uint32_t clut_key = 0; // the palette this warp has staged; survives epochs that load no CLUT snapshot
  PreBin pb;
  bool have_pre = false;

  auto epoch_item = [&](const Epoch& ep) {
    WaveItem it;
    it.instance = w.instance;
    it.cmd_begin = w.cmd_base + ep.cmd_begin;
    it.cmd_end = w.cmd_base + ep.cmd_end;
    it.kind = ep.kind;
    it.clut_begin = w.clut_base + ep.clut_begin;
    it.clut_end = w.clut_base + ep.clut_end;
    return it;
  };

  for (uint32_t k = 0; k < w.n_epochs; k++)
  {
    const uint32_t slot = k % INST_EPOCH_WINDOW;
    if (slot == 0)
    {
      // every CTA keeps its own copy of the next window of epoch descriptors and tile sets (immutable inputs)
      const uint32_t n = min(INST_EPOCH_WINDOW, w.n_epochs - k);
      const uint32_t* src = reinterpret_cast<const uint32_t*>(epochs + w.epoch_first + k);
      for (uint32_t i = threadIdx.x; i < n * (sizeof(Epoch) / 4); i += NT)
        reinterpret_cast<uint32_t*>(cs.epochs)[i] = __ldg(src + i);
      for (uint32_t i = threadIdx.x; i < n * EPOCH_SET_WORDS; i += NT)
        (&cs.tile_sets[0][0])[i] = __ldg(tile_sets + static_cast<size_t>(w.epoch_first + k) * EPOCH_SET_WORDS + i);
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < n; i += NT)
      {
        uint32_t c = 0;
#pragma unroll
        for (uint32_t q = 0; q < 8; q++)
          c += __popc(cs.tile_sets[i][q]);
        cs.tile_total[i] = c;
      }
      __syncthreads();
    }
    const WaveItem it = epoch_item(cs.epochs[slot]);
    if (it.clut_end > it.clut_begin && !(debug_skip & 4u)) // uniform over the cluster
    {
      clut_key = 0;
      for (uint32_t o = it.clut_begin + rank; o < it.clut_end; o += n_ctas)
      {
        const ClutOp op = clut_ops[o];
        for (uint32_t t = threadIdx.x; t < (op.is8bit ? 256u : 16u); t += NT)
          cluts[op.dst_slot * CLUT_ENTRIES + t] = __ldcg(vram + op.y * VRAM_W + ((op.x + t) & (VRAM_W - 1)));
      }
      cluster_barrier();
    }
    if (it.cmd_end > it.cmd_begin && !(debug_skip & (it.kind == EPOCH_TILED ? 2u : 1u))) // (debug_skip: timing experiments only)
    {
      if (it.kind == EPOCH_TILED)
      {
        // (most epochs of such a stream touch a few tiles: a warp without a strip goes straight to the barrier)
        const uint32_t total = cs.tile_total[slot] * CLUSTER_STRIPS;
        uint32_t mine = 0, before = 0;
        if (gw < total)
        {
          mine = (lane < 8) ? cs.tile_sets[slot][lane] : 0u; // lane L: tiles 32 L .. 32 L + 31
          tile_set_prefix(mine, lane, before);
        }
        for (uint32_t j = gw; j < total; j += n_gw)
        {
          const uint32_t tile = nth_tile(mine, before, __popc(mine), j / CLUSTER_STRIPS, lane), strip = j % CLUSTER_STRIPS;
          const uint32_t tx = tile % TILES_X, ty = tile / TILES_X;
          process_epoch_region<RH, true>(it, sh_all[warp], tx, ty, ty * TILE_H + strip * RH, setups, tile_cmds, vram, cluts, payload, lane,
                                         (have_pre && j == gw) ? &pb : nullptr, &clut_key);
          __syncwarp();
        }
      }
      else if (rank == 0 || it.kind == EPOCH_COPY_OVERLAP)
      {
        if (threadIdx.x < SETUP_WORDS)
          cs.stage[threadIdx.x] = reinterpret_cast<const uint32_t*>(setups + it.cmd_begin)[threadIdx.x];
        __syncthreads();
        const PrimSetup& s = *reinterpret_cast<const PrimSetup*>(cs.stage);
        if (s.op != OP_NOP)
        {
          if (it.kind == EPOCH_SELF_SAMPLING)
          {
            if (!(debug_skip & 32u))
              serial_self_sampling<NT>(s, vram, cluts);
          }
          else if (it.kind == EPOCH_COPY_OVERLAP)
          {
            // every CTA of the cluster is here with the same record
            if (!(debug_skip & 8u) && ((debug_skip & 128u) || !cluster_copy_fast<NT>(s, vram, rank, n_ctas)) && rank == 0)
              serial_copy<NT>(s, vram);
          }
          else if (!(debug_skip & 16u))
            serial_upload_masked<NT>(s, vram, payload, cs.row_count);
        }
      }
    }
    // The barrier that ends the epoch, split: between arriving and waiting the warp bins its first strip of the next
    // epoch and requests that strip's first setup record — two dependent trips to L2 that need nothing this epoch
    // writes (bins and setup records are inputs of the kernel).
    cluster_arrive();
    have_pre = false;
    if (slot + 1 < INST_EPOCH_WINDOW && k + 1 < w.n_epochs && !(debug_skip & 64u))
    {
      const Epoch& en = cs.epochs[slot + 1];
      if (en.kind == EPOCH_TILED && en.cmd_end > en.cmd_begin && gw < cs.tile_total[slot + 1] * CLUSTER_STRIPS)
      {
        const uint32_t mine = (lane < 8) ? cs.tile_sets[slot + 1][lane] : 0u;
        uint32_t before;
        tile_set_prefix(mine, lane, before);
        {
          const uint32_t tile = nth_tile(mine, before, __popc(mine), gw / CLUSTER_STRIPS, lane);
          pb = prebin_region<RH>(epoch_item(en), sh_all[warp], tile % TILES_X, tile / TILES_X, setups, tile_cmds, lane);
          have_pre = true;
        }
      }
    }
    cluster_wait();
  }
}

// ------------------------------------------------------------------------------------------------ hash
constexpr uint32_t HASH_BLOCKS_PER_INSTANCE = 32;

__global__ void __launch_bounds__(256) hash_kernel(const uint16_t* __restrict__ vram_pool, uint32_t first_instance,
                                                   unsigned long long* __restrict__ out)
{
  const uint32_t inst = first_instance + blockIdx.y;
  const uint4* v = reinterpret_cast<const uint4*>(vram_pool + static_cast<size_t>(inst) * VRAM_PIXELS);
  constexpr uint32_t VEC_PER_INSTANCE = VRAM_PIXELS * 2 / 16; // 65536
  constexpr uint32_t VEC_PER_BLOCK = VEC_PER_INSTANCE / HASH_BLOCKS_PER_INSTANCE;
  uint64_t h = 0;
  for (uint32_t k = threadIdx.x; k < VEC_PER_BLOCK; k += 256)
  {
    const uint32_t vi = blockIdx.x * VEC_PER_BLOCK + k;
    const uint4 q = __ldg(v + vi);
    const uint64_t wi = static_cast<uint64_t>(vi) * 4;
    h += mix64(((wi + 0) << 32) | q.x);
    h += mix64(((wi + 1) << 32) | q.y);
    h += mix64(((wi + 2) << 32) | q.z);
    h += mix64(((wi + 3) << 32) | q.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    h += __shfl_xor_sync(0xFFFFFFFFu, h, o);
  __shared__ uint64_t part[8];
  if ((threadIdx.x & 31) == 0)
    part[threadIdx.x >> 5] = h;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    uint64_t t = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
      t += part[k];
    atomicAdd(out + inst, static_cast<unsigned long long>(t));
  }
}

// ------------------------------------------------------------------------------------------------ scan-out
__device__ __forceinline__ uint32_t e5to8(uint32_t c)
{
  return (c * 527u + 23u) >> 6; // gpu_helpers.h:18-31
}

__global__ void scanout_kernel(const uint16_t* __restrict__ vram, ScanoutParams p, uint8_t* __restrict__ out)
{
  const uint32_t col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
  if (col >= p.width)
    return;
  if (!p.is24)
  {
    uint32_t idx;
    if (p.fast)
      idx = (p.src_y * VRAM_W + p.src_x + row * (VRAM_W << p.line_skip) + col) & (VRAM_PIXELS - 1); // gpu_sw.cpp:246-266
    else
      idx = ((p.src_y + (row << p.line_skip)) & (VRAM_H - 1)) * VRAM_W + ((p.src_x + col) & (VRAM_W - 1)); // :272-284
    const uint32_t c = vram[idx];
    if (p.format == 0)
    {
      const uint32_t v = e5to8(c & 31u) | (e5to8((c >> 5) & 31u) << 8) | (e5to8((c >> 10) & 31u) << 16) | ((c >> 15) ? 0xFF000000u : 0u);
      reinterpret_cast<uint32_t*>(out + static_cast<size_t>(row) * p.out_stride)[col] = v;
    }
    else
    {
      uint32_t o;
      if (p.format == 1)
        o = (c & 0x83E0u) | ((c >> 10) & 0x1Fu) | ((c & 0x1Fu) << 10);
      else if (p.format == 2)
        o = ((c & 0x3E0u) << 1) | ((c >> 9) & 0x3Eu) | ((c & 0x1Fu) << 11) | (c >> 15);
      else
        o = ((c & 0x3E0u) << 1) | ((c & 0x20u) << 1) | ((c >> 10) & 0x1Fu) | ((c & 0x1Fu) << 11);
      reinterpret_cast<uint16_t*>(out + static_cast<size_t>(row) * p.out_stride)[col] = static_cast<uint16_t>(o);
    }
    return;
  }
  uint32_t rgb;
  if (p.fast)
  {
    // linear byte addressing, 3 bytes per pixel (gpu_sw.cpp:307-323); wraps at the end of VRAM instead of reading past it
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(vram);
    const uint32_t a = (p.src_y * VRAM_W + p.src_x) * 2u + p.skip_x * 3u + row * ((VRAM_W << p.line_skip) * 2u) + col * 3u;
    constexpr uint32_t M = VRAM_PIXELS * 2u - 1u;
    rgb = bytes[a & M] | (bytes[(a + 1) & M] << 8) | (bytes[(a + 2) & M] << 16);
  }
  else
  {
    const uint16_t* srow = vram + ((p.src_y + (row << p.line_skip)) & (VRAM_H - 1)) * VRAM_W; // gpu_sw.cpp:331-342
    const uint32_t off = p.src_x + (((p.skip_x + col) * 3u) / 2u);
    const uint32_t s0 = srow[off & (VRAM_W - 1)], s1 = srow[(off + 1) & (VRAM_W - 1)];
    rgb = ((s1 << 16) | s0) >> ((col & 1u) * 8u);
  }
  reinterpret_cast<uint32_t*>(out + static_cast<size_t>(row) * p.out_stride)[col] = (rgb & 0x00FFFFFFu) | 0xFF000000u;
}

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t launch_setup(const PackedCmd* cmds, PrimSetup* setups, BinInfo* bins, uint32_t* tile_cmds, uint32_t n, cudaStream_t st)
{
  if (n == 0)
    return cudaSuccess;
  setup_kernel<<<(n + 127) / 128, 128, 0, st>>>(cmds, setups, bins, tile_cmds, n);
  return cudaGetLastError();
}

cudaError_t launch_clut(const WaveItem* items, uint32_t n_items, const ClutOp* ops, const uint16_t* vram_pool,
                        uint16_t* clut_pool, uint32_t clut_slots, cudaStream_t st)
{
  if (n_items == 0)
    return cudaSuccess;
  clut_kernel<<<n_items, 256, 0, st>>>(items, ops, vram_pool, clut_pool, clut_slots);
  return cudaGetLastError();
}

cudaError_t launch_raster(const WaveItem* items, uint32_t n_items, const PrimSetup* setups, const uint32_t* tile_cmds,
                          uint16_t* vram_pool, const uint16_t* clut_pool, const uint16_t* payload, uint32_t clut_slots,
                          cudaStream_t st, uint32_t tiled_cmds, unsigned int* loop_counter)
{
  // Waves with enough work use the persistent-loop kernel (PSXGPU_RASTER_LOOP=0 forces one CTA per tile everywhere).
  static const int loop_mode = []() {
    const char* e = std::getenv("PSXGPU_RASTER_LOOP");
    return e ? std::atoi(e) : 1;
  }();
  if (loop_mode > 0 && loop_counter != nullptr && n_items >= 16 && n_items * static_cast<uint64_t>(NUM_TILES) < 0xFFFF0000ull)
  {
    // The work counter belongs to the caller's context (one per stream: the reset, the wave and the next reset are
    // ordered by the stream), so contexts on other streams or host threads never share one.
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaMemsetAsync(loop_counter, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess)
      return e;
    const uint32_t total = n_items * NUM_TILES;
    static const uint32_t ctas_per_sm = []() {
      const char* e = std::getenv("PSXGPU_LOOP_CTAS"); // experiments: persistent one-warp CTAs per SM (fewer leave room for the encoder's kernels beside a wave)
      const int v = e ? std::atoi(e) : 0;
      return (v >= 1 && v <= PSX_RASTER_MIN_CTAS) ? static_cast<uint32_t>(v) : static_cast<uint32_t>(PSX_RASTER_MIN_CTAS);
    }();
    static const bool carveout_set = []() {
      const char* e = std::getenv("PSXGPU_LOOP_CARVEOUT"); // experiments: preferred shared-memory carve-out in percent (the rest is L1)
      if (e && std::atoi(e) > 0)
        cudaFuncSetAttribute(raster_loop_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(e));
      return true;
    }();
    (void)carveout_set;
    const uint32_t ctas = std::min<uint32_t>(static_cast<uint32_t>(sms) * ctas_per_sm, total / 2u + 1u);
    static const uint32_t grab_light = []() {
      const char* e = std::getenv("PSXGPU_GRAB_LIGHT"); // experiments
      return (e && std::atoi(e) > 0) ? static_cast<uint32_t>(std::atoi(e)) : LOOP_GRAB_LIGHT;
    }();
    static const uint32_t grab_heavy = []() {
      const char* e = std::getenv("PSXGPU_GRAB_HEAVY");
      return (e && std::atoi(e) > 0) ? static_cast<uint32_t>(std::atoi(e)) : LOOP_GRAB_HEAVY;
    }();
    const uint32_t grab = (tiled_cmds <= static_cast<uint64_t>(n_items) * LOOP_LIGHT_CMDS_PER_ITEM) ? grab_light : grab_heavy;
    raster_loop_kernel<<<ctas, 32, 0, st>>>(items, n_items, setups, tile_cmds, vram_pool, clut_pool, payload, clut_slots, loop_counter, grab);
    return cudaGetLastError();
  }
  // gridDim.y is limited to 65535: chunk the wave
  for (uint32_t off = 0; off < n_items; off += 32768)
  {
    const uint32_t n = (n_items - off) < 32768u ? (n_items - off) : 32768u;
    raster_kernel<<<dim3(NUM_TILES / PSX_RASTER_WARPS, n), 32 * PSX_RASTER_WARPS, 0, st>>>(items + off, setups, tile_cmds, vram_pool, clut_pool, payload, clut_slots);
  }
  return cudaGetLastError();
}

cudaError_t launch_item_union(WaveItem* items, uint32_t n_items, const BinInfo* bins, cudaStream_t st)
{
  if (n_items == 0)
    return cudaSuccess;
  item_union_kernel<<<(n_items + 3) / 4, 128, 0, st>>>(items, n_items, bins);
  return cudaGetLastError();
}

cudaError_t launch_serial(const WaveItem* items, uint32_t n_items, const PrimSetup* setups, uint16_t* vram_pool,
                          const uint16_t* clut_pool, const uint16_t* payload, uint32_t clut_slots, cudaStream_t st)
{
  if (n_items == 0)
    return cudaSuccess;
  serial_kernel<<<n_items, SERIAL_THREADS, 0, st>>>(items, setups, vram_pool, clut_pool, payload, clut_slots);
  return cudaGetLastError();
}

int persistent_max_instances(int device)
{
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raster_persistent_kernel, PERSIST_THREADS, 0) != cudaSuccess)
    return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess)
    return 0;
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
  return coop ? (per_sm * sms) / static_cast<int>(NUM_TILES) : 0;
}

cudaError_t launch_persistent(const WaveItem* items, const void* waves, uint32_t n_waves, const uint32_t* instances, uint32_t n_instances,
                              const PrimSetup* setups, const uint32_t* tile_cmds, const ClutOp* clut_ops, uint16_t* vram_pool,
                              uint16_t* clut_pool, const uint16_t* payload, uint32_t clut_slots, unsigned int* barrier_counter,
                              cudaStream_t st)
{
  if (n_waves == 0 || n_instances == 0)
    return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(barrier_counter, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess)
    return e;
  const PersistWave* wv = static_cast<const PersistWave*>(waves);
  void* args[] = {&items, &wv, &n_waves, &instances, &setups, &tile_cmds, &clut_ops, &vram_pool, &clut_pool, &payload, &clut_slots,
                  &barrier_counter};
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(raster_persistent_kernel), dim3(NUM_TILES, n_instances),
                                     dim3(PERSIST_THREADS), args, 0, st);
}

// Instances up to which a flush in instance mode runs as one cluster per instance (PSXGPU_CLUSTER_MAX; 0 = never).
uint32_t cluster_max_instances()
{
  static const uint32_t v = []() {
    const char* e = std::getenv("PSXGPU_CLUSTER_MAX");
    return e ? static_cast<uint32_t>(std::atoi(e)) : 8u;
  }();
  return v;
}

cudaError_t launch_instances(const InstWork* work, uint32_t n_instances, const Epoch* epochs, uint32_t* tile_sets, const PrimSetup* setups,
                             const BinInfo* bins, const uint32_t* tile_cmds, const ClutOp* clut_ops, uint16_t* vram_pool, uint16_t* clut_pool,
                             const uint16_t* payload, uint32_t clut_slots, bool wide, cudaStream_t st, unsigned int* inst_counter)
{
  if (n_instances == 0)
    return cudaSuccess;
  cudaError_t e;
  if (wide && n_instances <= cluster_max_instances())
  {
    // cluster mode: 16 CTAs per instance where the GPU can place such a cluster (non-portable size), else 8
    static const int rh = []() {
      const char* e = std::getenv("PSXGPU_CLUSTER_RH");
      return (e && std::atoi(e) == 8) ? 8 : 4; // rows per strip (measured: 4 is 8 % faster than 8 on the hazard stress stream)
    }();
    static const int warps = []() {
      const char* e = std::getenv("PSXGPU_CLUSTER_WARPS");
      const int v = e ? std::atoi(e) : 16; // (measured on the hazard stress stream: 16 warps 16.8 ms, 32 warps 18.0, 8 warps 22.0)
      return (v == 8 || v == 32) ? v : 16;
    }();
    using ClusterKernel = void (*)(const InstWork*, const Epoch*, const uint32_t*, const PrimSetup*, const uint32_t*, const ClutOp*, uint16_t*,
                                   uint16_t*, const uint16_t*, uint32_t, uint32_t);
    ClusterKernel const kernel = rh == 4 ? (warps == 8 ? raster_cluster_kernel<4, 8> : warps == 16 ? raster_cluster_kernel<4, 16> : raster_cluster_kernel<4, 32>)
                                         : (warps == 8 ? raster_cluster_kernel<8, 8> : warps == 16 ? raster_cluster_kernel<8, 16> : raster_cluster_kernel<8, 32>);
    const uint32_t CLUSTER_WARPS = static_cast<uint32_t>(warps);
    const size_t smem = sizeof(RasterShared<8>) * CLUSTER_WARPS + sizeof(InstanceShared);
    static const int cluster_size = [&]() {
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return 0;
      const char* env = std::getenv("PSXGPU_CLUSTER_SIZE");
      const int want = env ? std::atoi(env) : 16;
      for (int cs = want; cs >= 2; cs >>= 1)
      {
        if (cs > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)
        {
          cudaGetLastError();
          continue;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(static_cast<unsigned>(cs));
        cfg.blockDim = dim3(32 * CLUSTER_WARPS);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = static_cast<unsigned>(cs);
        attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) == cudaSuccess && n >= 1)
          return cs;
        cudaGetLastError();
      }
      return 0;
    }();
    if (cluster_size >= 2)
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(n_instances * static_cast<unsigned>(cluster_size));
      cfg.blockDim = dim3(32 * CLUSTER_WARPS);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = st;
      cudaLaunchAttribute attr;
      attr.id = cudaLaunchAttributeClusterDimension;
      attr.val.clusterDim.x = static_cast<unsigned>(cluster_size);
      attr.val.clusterDim.y = attr.val.clusterDim.z = 1;
      cfg.attrs = &attr;
      cfg.numAttrs = 1;
      epoch_tiles_kernel<<<n_instances, 128, 0, st>>>(work, epochs, tile_cmds, bins, tile_sets, 0u);
      const uint32_t* ts = tile_sets;
      static const uint32_t debug_skip = []() {
        const char* e = std::getenv("PSXGPU_DEBUG_SKIP"); // timing experiments: 1 = no serial epochs, 2 = no tiled epochs, 4 = no CLUT ops (results are wrong)
        return e ? static_cast<uint32_t>(std::atoi(e)) : 0u;
      }();
      return cudaLaunchKernelEx(&cfg, kernel, work, epochs, ts, setups, tile_cmds, clut_ops, vram_pool, clut_pool, payload,
                                clut_slots, debug_skip);
    }
  }
  // warps per CTA x CTAs per SM: 32 x 1 — one instance has an SM to itself, all of its epoch's tiles in flight at once and
  // its 1 MiB of VRAM the only one that SM pulls through L2.  Measured on 2048 instances of the hazard variant of the
  // batched workload: 8 x 4 23.5 ms, 16 x 2 20.1 ms, 32 x 1 18.6 ms (4 x 8: worse still).
  using InstKernel = void (*)(const InstWork*, const Epoch*, const uint32_t*, const PrimSetup*, const uint32_t*, const ClutOp*, uint16_t*,
                              uint16_t*, const uint16_t*, uint32_t, uint32_t);
  static const uint32_t run_max = []() {
    const char* v = std::getenv("PSXGPU_INST_RUN"); // tiled epochs in flight together (tile-granular dependencies); 1 = a barrier per epoch
    const int n = v ? std::atoi(v) : 1;
    return static_cast<uint32_t>(n < 1 ? 1 : (n > static_cast<int>(INST_RUN_MAX) ? static_cast<int>(INST_RUN_MAX) : n));
  }();
  constexpr int variant = 0; // (the other shapes are no longer built)
  InstKernel kernel = raster_instance_kernel<32, 1>;
  const uint32_t W = 32;
  const bool dag = run_max > 1 && W == 32;
  if (dag)
    kernel = raster_instance_dag_kernel<32, 1>;
  // slot mode (raster_multi_kernel): one 32-warp CTA per SM serves n_slots instances at a time
  static const uint32_t n_slots = []() {
    const char* v = std::getenv("PSXGPU_SLOTS"); // 0 = one instance per CTA and a barrier per epoch (raster_instance_kernel)
    const int n = v ? std::atoi(v) : 3;
    return static_cast<uint32_t>(n < 0 ? 0 : (n > static_cast<int>(MULTI_SLOTS_MAX) ? static_cast<int>(MULTI_SLOTS_MAX) : n));
  }();
  static const int n_sms = []() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    return n;
  }();
  // few instances: fewer slots per CTA, so that every SM gets one; with one instance per SM the barrier per epoch of
  // raster_instance_kernel is cheaper than the claim protocol (18.5 against 21.7 ms at one slot)
  static const bool force_slots = std::getenv("PSXGPU_SLOTS") != nullptr; // (set: that many slots whatever the instance count — tests)
  uint32_t slots = n_slots;
  while (!force_slots && slots > 1 && static_cast<uint64_t>(n_sms) * (slots - 1) >= n_instances)
    slots--;
  if (n_slots != 0 && (slots > 1 || force_slots) && variant == 0 && !dag && inst_counter != nullptr)
  {
    const uint32_t ctas = std::min<uint32_t>(static_cast<uint32_t>(n_sms), (n_instances + slots - 1) / slots);
    // (28 / 24 warps per CTA at 72 / 80 registers and more L1 were measured slower: 16.39 / 17.16 against 16.03 ms)
    constexpr uint32_t multi_warps = 32;
    auto const mk = raster_multi_kernel<32>;
    const size_t smem_multi = sizeof(RasterShared<TILE_H>) * multi_warps + sizeof(MultiShared);
    if ((e = cudaFuncSetAttribute(mk, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_multi))) != cudaSuccess)
      return e;
    if ((e = cudaMemsetAsync(inst_counter, 0, sizeof(unsigned int), st)) != cudaSuccess)
      return e;
    epoch_tiles_kernel<<<n_instances, 128, 0, st>>>(work, epochs, tile_cmds, bins, tile_sets, 0u);
    mk<<<ctas, 32 * multi_warps, smem_multi, st>>>(work, n_instances, epochs, tile_sets, setups, tile_cmds, clut_ops, vram_pool, clut_pool, payload,
                                                   clut_slots, slots, inst_counter);
    return cudaGetLastError();
  }
  epoch_tiles_kernel<<<n_instances, 128, 0, st>>>(work, epochs, tile_cmds, bins, tile_sets, dag ? 1u : 0u);
  const size_t smem = sizeof(RasterShared<TILE_H>) * W + (dag ? sizeof(InstanceDagShared) : sizeof(InstanceShared));
  if ((e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))) != cudaSuccess)
    return e;
  kernel<<<n_instances, 32 * W, smem, st>>>(work, epochs, tile_sets, setups, tile_cmds, clut_ops, vram_pool, clut_pool, payload, clut_slots,
                                            run_max);
  return cudaGetLastError();
}

cudaError_t launch_hash(const uint16_t* vram_pool, uint32_t first, uint32_t count, unsigned long long* out, cudaStream_t st)
{
  if (count == 0)
    return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(out + first, 0, sizeof(unsigned long long) * count, st);
  if (e != cudaSuccess)
    return e;
  for (uint32_t off = 0; off < count; off += 32768)
  {
    const uint32_t n = (count - off) < 32768u ? (count - off) : 32768u;
    hash_kernel<<<dim3(HASH_BLOCKS_PER_INSTANCE, n), 256, 0, st>>>(vram_pool, first + off, out);
  }
  return cudaGetLastError();
}

cudaError_t launch_scanout(const uint16_t* vram, const ScanoutParams& p, uint8_t* out, cudaStream_t st)
{
  if (p.width == 0 || p.height == 0)
    return cudaSuccess;
  scanout_kernel<<<dim3((p.width + 127) / 128, p.height), 128, 0, st>>>(vram, p, out);
  return cudaGetLastError();
}

} // namespace psx
