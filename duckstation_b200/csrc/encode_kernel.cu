// encode_kernel.cu — the command encoder of the batched path, one warp per VRAM instance (see device_encoder.h).
//
// It restates encoder.cpp for a warp: every lane runs the same (uniform) control flow over the instance's records, so
// the sequential semantics of the stream are kept, and the warp's width is spent on the three things the host encoder
// loops over — the 32x64-cell hazard grids (lane = 32-pixel column, one u64 of 8-row cells per lane and set, all in
// registers), the CLUT snapshot table (8 slots per lane) and the record fetch (one coalesced 128-byte load per record,
// fields handed round with shuffles).  Reference semantics cited in encoder.cpp apply unchanged:
// GPUBackend::HandleCommand (src/core/gpu_backend.cpp:383-544), GPU_SW::DrawPolygon's quad split (gpu_sw.cpp:113-115),
// the rasteriser's size culls (gpu_sw_rasterizer.inl:1459, :834-835), UpdateCLUT as a cache fill
// (gpu_sw_rasterizer.cpp:43-66).
#include "device_encoder.h"

#include "../../include/psxgpu_wire.h"

namespace psx {
namespace {

#ifndef PSX_ENC_WARPS
#define PSX_ENC_WARPS 4
#endif
constexpr uint32_t ENC_WARPS = PSX_ENC_WARPS;
constexpr uint32_t WALK_WARPS = PSX_ENC_WARPS;
constexpr uint32_t FULL = 0xFFFFFFFFu;

// Following the record sizes is a chain of dependent loads, and on this part a global load that misses L1 costs
// 600+ cycles even when it hits L2.  So the stream is staged through shared memory 4 KB at a time — one coalesced
// round trip per window (128 bytes per lane) — and the chain runs on shared-memory latency instead.
constexpr uint32_t WIN_BYTES = 4096;

struct StreamWindow
{
  const uint8_t* wire;
  uint32_t* smem;    // [WIN_BYTES / 4], this warp's
  uint32_t base;     // byte offset of smem[0] in the wire buffer (128-byte aligned); base == end => empty
  uint32_t limit;    // bytes of the wire buffer this warp may touch (stream end + slack)
  uint32_t lane;

  __device__ __forceinline__ void fill(uint32_t pos)
  {
    __syncwarp();
    base = pos & ~127u;
    const uint32_t at = base + lane * 128u;
    const uint4* src = reinterpret_cast<const uint4*>(wire + at);
    uint4* dst = reinterpret_cast<uint4*>(smem + lane * 32u);
    if (at < limit)
    {
#pragma unroll
      for (int i = 0; i < 8; i++)
        dst[i] = __ldg(src + i);
    }
    if (at + WIN_BYTES < limit) // the stream arrives from HBM: have the next window on its way to L2
      asm volatile("prefetch.global.L2 [%0];" ::"l"(wire + at + WIN_BYTES));
    __syncwarp();
  }
  // the 32-bit word at byte offset `pos` (4-byte aligned), refilling the window when pos lies outside it
  __device__ __forceinline__ uint32_t word(uint32_t pos)
  {
    if (pos < base || pos + 4u > base + WIN_BYTES)
      fill(pos);
    return smem[(pos - base) >> 2];
  }
  __device__ __forceinline__ bool holds(uint32_t pos, uint32_t bytes) const { return pos >= base && pos + bytes <= base + WIN_BYTES; }
};

struct Rects // up to four wrapped pieces of one VRAM rectangle (RectSet in encoder.h); uniform across the warp
{
  uint32_t n;
  uint32_t x0[4], y0[4], x1[4], y1[4];
};

__device__ __forceinline__ void rs_push(Rects& s, uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1)
{
#pragma unroll
  for (uint32_t k = 0; k < 4; k++)
    if (s.n == k)
    {
      s.x0[k] = x0;
      s.y0[k] = y0;
      s.x1[k] = x1;
      s.y1[k] = y1;
    }
  if (s.n < 4)
    s.n++;
}

__device__ __forceinline__ void rs_add_wrapped(Rects& s, uint32_t x, uint32_t y, uint32_t w, uint32_t h)
{
  if (w == 0 || h == 0)
    return;
  x &= VRAM_W - 1;
  y &= VRAM_H - 1;
  w = min(w, VRAM_W);
  h = min(h, VRAM_H);
  const bool wx = x + w > VRAM_W, wy = y + h > VRAM_H;
  if (!wx && !wy && s.n == 0)
  {
    s.x0[0] = x;
    s.y0[0] = y;
    s.x1[0] = x + w - 1;
    s.y1[0] = y + h - 1;
    s.n = 1;
    return;
  }
  const uint32_t xa1 = min(x + w, VRAM_W) - 1, xb1 = wx ? (x + w - VRAM_W - 1) : 0;
  const uint32_t ya1 = min(y + h, VRAM_H) - 1, yb1 = wy ? (y + h - VRAM_H - 1) : 0;
  rs_push(s, x, y, xa1, ya1);
  if (wx)
    rs_push(s, 0, y, xb1, ya1);
  if (wy)
  {
    rs_push(s, x, 0, xa1, yb1);
    if (wx)
      rs_push(s, 0, 0, xb1, yb1);
  }
}

// Nearly every set is a single rectangle (no wrap): piece 0 is handled straight-line and the (uniform) branch for
// pieces 1..3 is almost never taken.
__device__ __forceinline__ bool rs_overlap(const Rects& a, uint32_t i, const Rects& b, uint32_t j)
{
  return a.x0[i] <= b.x1[j] && b.x0[j] <= a.x1[i] && a.y0[i] <= b.y1[j] && b.y0[j] <= a.y1[i];
}

__device__ __forceinline__ bool rs_intersects(const Rects& a, const Rects& b)
{
  if (a.n == 0 || b.n == 0)
    return false;
  bool hit = rs_overlap(a, 0, b, 0);
  if ((a.n | b.n) > 1)
  {
#pragma unroll
    for (uint32_t i = 0; i < 4; i++)
#pragma unroll
      for (uint32_t j = 0; j < 4; j++)
        if (i + j != 0)
          hit |= (i < a.n && j < b.n && rs_overlap(a, i, b, j));
  }
  return hit;
}

__device__ __forceinline__ uint64_t rowmask(uint32_t y0, uint32_t y1)
{
  const uint32_t a = y0 >> 3, b = y1 >> 3; // 0..63
  return ((~0ull) >> (63u - b)) & ((~0ull) << a);
}

// Cell rows of `s` inside this lane's 32-pixel column.
__device__ __forceinline__ uint64_t rs_column(const Rects& s, uint32_t lane)
{
  uint64_t m = 0;
  if (s.n != 0)
  {
    if (lane >= (s.x0[0] >> 5) && lane <= (s.x1[0] >> 5))
      m = rowmask(s.y0[0], s.y1[0]);
    if (s.n > 1)
    {
#pragma unroll
      for (uint32_t k = 1; k < 4; k++)
        if (k < s.n && lane >= (s.x0[k] >> 5) && lane <= (s.x1[k] >> 5))
          m |= rowmask(s.y0[k], s.y1[k]);
    }
  }
  return m;
}

// HazardSet (encoder.h): bounding box prefilter + cell grid, the grid spread over the warp.
struct Hazard
{
  uint64_t col; // this lane's column
  uint32_t any, x0, y0, x1, y1;
};

__device__ __forceinline__ void hz_clear(Hazard& h)
{
  h.col = 0;
  h.any = 0;
  h.x0 = h.y0 = 0xFFFFu; // empty box: min/max below need no special first case
  h.x1 = h.y1 = 0;
}

__device__ __forceinline__ void hz_add(Hazard& h, const Rects& s, uint32_t lane)
{
  if (s.n == 0)
    return;
  h.any = 1;
  h.x0 = min(h.x0, s.x0[0]);
  h.y0 = min(h.y0, s.y0[0]);
  h.x1 = max(h.x1, s.x1[0]);
  h.y1 = max(h.y1, s.y1[0]);
  if (s.n > 1)
  {
#pragma unroll
    for (uint32_t k = 1; k < 4; k++)
      if (k < s.n)
      {
        h.x0 = min(h.x0, s.x0[k]);
        h.y0 = min(h.y0, s.y0[k]);
        h.x1 = max(h.x1, s.x1[k]);
        h.y1 = max(h.y1, s.y1[k]);
      }
  }
  h.col |= rs_column(s, lane);
}

// bounding-box half of hz_add: the cell grid is brought up to date lazily (WarpEncoder::ensure_grids)
__device__ __forceinline__ void hz_add_box(Hazard& h, const Rects& s)
{
  if (s.n == 0)
    return;
  h.any = 1;
  h.x0 = min(h.x0, s.x0[0]);
  h.y0 = min(h.y0, s.y0[0]);
  h.x1 = max(h.x1, s.x1[0]);
  h.y1 = max(h.y1, s.y1[0]);
  if (s.n > 1)
  {
#pragma unroll
    for (uint32_t k = 1; k < 4; k++)
      if (k < s.n)
      {
        h.x0 = min(h.x0, s.x0[k]);
        h.y0 = min(h.y0, s.y0[k]);
        h.x1 = max(h.x1, s.x1[k]);
        h.y1 = max(h.y1, s.y1[k]);
      }
  }
}

__device__ __forceinline__ bool hz_maybe(const Hazard& h, const Rects& s)
{
  if (!h.any || s.n == 0)
    return false;
  bool maybe = s.x0[0] <= h.x1 && h.x0 <= s.x1[0] && s.y0[0] <= h.y1 && h.y0 <= s.y1[0];
  if (s.n > 1)
  {
#pragma unroll
    for (uint32_t k = 1; k < 4; k++)
      maybe |= (k < s.n && s.x0[k] <= h.x1 && h.x0 <= s.x1[k] && s.y0[k] <= h.y1 && h.y0 <= s.y1[k]);
  }
  return maybe;
}

__device__ __forceinline__ bool hz_grid(const Hazard& h, const Rects& s, uint32_t lane)
{
  return __any_sync(FULL, (h.col & rs_column(s, lane)) != 0);
}

__device__ __forceinline__ bool hz_hits(const Hazard& h, const Rects& s, uint32_t lane)
{
  if (!h.any || s.n == 0)
    return false;
  bool maybe = s.x0[0] <= h.x1 && h.x0 <= s.x1[0] && s.y0[0] <= h.y1 && h.y0 <= s.y1[0];
  if (s.n > 1)
  {
#pragma unroll
    for (uint32_t k = 1; k < 4; k++)
      maybe |= (k < s.n && s.x0[k] <= h.x1 && h.x0 <= s.x1[k] && s.y0[k] <= h.y1 && h.y0 <= s.y1[k]);
  }
  if (!maybe)
    return false;
  return __any_sync(FULL, (h.col & rs_column(s, lane)) != 0);
}

__device__ __forceinline__ void store_cmd(PackedCmd* dst, const uint32_t (&w)[16], uint32_t lane)
{
  if (lane == 0)
  {
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d[1] = make_uint4(w[4], w[5], w[6], w[7]);
    d[2] = make_uint4(w[8], w[9], w[10], w[11]);
    d[3] = make_uint4(w[12], w[13], w[14], w[15]);
  }
}

struct Vtx
{
  int32_t x, y;
  uint32_t rgb; // r | g << 8 | b << 16
  uint32_t u, v;
};

// One instance's encoder.  Everything is uniform across the warp except the Hazard columns.
struct WarpEncoder
{
  uint32_t lane;
  // outputs
  PackedCmd* cmds;
  ClutOp* clut_ops;
  Epoch* epochs;
  uint32_t cmd_cap, clut_cap, epoch_cap;
  uint32_t n_cmds, n_clut, n_epochs, overflow;
  uint32_t epoch_cmd_begin, epoch_clut_begin, epoch_serial;
  // GPU state
  uint32_t area[4];
  uint32_t cur_lo, cur_hi, hand, slots, mod_crop;
  Hazard writes, reads, clut_src;
  uint32_t* key;  // shared, [256]
  uint32_t* last; // shared, [256]: serial of the last epoch that references the slot's contents
  unsigned long long* colscratch; // shared, [32]
  unsigned long long* colscratch2; // shared, [32]
  uint8_t* quick; // shared, [64]: position hash -> slot it was last found in (validated against key[])
  // Footprints of the open epoch's commands, one uint4 per command slot: {wr0, wr1, rd0, rd1} packed x | y << 16
  // (0xFFFFFFFF = none).  `writes` / `reads` keep their bounding boxes exact at all times, but their cell grids cover
  // only commands below grid_upto; the rest is folded in from here the first time a box test is inconclusive.
  uint4* rects;
  uint32_t grid_upto;
  // statistics
  uint32_t st_prims, st_raw, st_war, st_clutcut, st_self, st_copy, st_upload, st_clutops, st_hits, st_rejected;

  __device__ __forceinline__ void close_epoch(uint32_t kind)
  {
    if (n_epochs < epoch_cap)
    {
      if (lane == 0)
      {
        Epoch e;
        e.cmd_begin = epoch_cmd_begin;
        e.cmd_end = n_cmds;
        e.kind = kind;
        e.clut_begin = epoch_clut_begin;
        e.clut_end = n_clut;
        epochs[n_epochs] = e;
      }
    }
    else
    {
      overflow = 1;
    }
    n_epochs++;
    epoch_cmd_begin = n_cmds;
    epoch_clut_begin = n_clut;
    hz_clear(writes);
    hz_clear(reads);
    grid_upto = n_cmds;
    epoch_serial++;
  }

  __device__ __forceinline__ void ensure_grids()
  {
    if (grid_upto >= n_cmds)
      return;
    colscratch[lane] = 0;
    colscratch2[lane] = 0;
    __syncwarp();
    const uint32_t end = min(n_cmds, cmd_cap);
    for (uint32_t i = grid_upto + lane; i < end; i += 32)
    {
      const uint4 e = rects[i];
      if (e.x != 0xFFFFFFFFu)
      {
        const unsigned long long rows = rowmask(e.x >> 16, e.y >> 16);
        for (uint32_t c = (e.x & 0xFFFFu) >> 5; c <= ((e.y & 0xFFFFu) >> 5); c++)
          atomicOr(&colscratch[c], rows);
      }
      if (e.z != 0xFFFFFFFFu)
      {
        const unsigned long long rows = rowmask(e.z >> 16, e.w >> 16);
        for (uint32_t c = (e.z & 0xFFFFu) >> 5; c <= ((e.w & 0xFFFFu) >> 5); c++)
          atomicOr(&colscratch2[c], rows);
      }
    }
    __syncwarp();
    writes.col |= colscratch[lane];
    reads.col |= colscratch2[lane];
    grid_upto = n_cmds;
    __syncwarp();
  }

  __device__ __forceinline__ void slot_rects(uint32_t k, Rects& src) const
  {
    src.n = 0;
    rs_add_wrapped(src, k & 0x3FFu, (k >> 10) & 0x1FFu, ((k >> 19) & 1u) ? 256u : 16u, 1);
  }

  // m_clut_sources: the valid slots' source pixels, as a hazard set
  __device__ __forceinline__ void rebuild_clut_sources()
  {
    colscratch[lane] = 0;
    __syncwarp();
    uint32_t x0 = 0xFFFFu, y0 = 0xFFFFu, x1 = 0, y1 = 0, any = 0;
    for (uint32_t s = lane; s < slots; s += 32)
    {
      const uint32_t k = key[s];
      if (!(k & ENC_SLOT_VALID))
        continue;
      Rects src;
      slot_rects(k, src);
#pragma unroll
      for (uint32_t i = 0; i < 2; i++)
        if (i < src.n)
        {
          for (uint32_t c = src.x0[i] >> 5; c <= (src.x1[i] >> 5); c++)
            atomicOr(&colscratch[c], 1ull << (src.y0[i] >> 3));
          x0 = min(x0, src.x0[i]);
          y0 = min(y0, src.y0[i]);
          x1 = max(x1, src.x1[i]);
          y1 = max(y1, src.y1[i]);
          any = 1;
        }
    }
    __syncwarp();
    clut_src.col = colscratch[lane];
    clut_src.any = __any_sync(FULL, any) ? 1u : 0u;
    clut_src.x0 = __reduce_min_sync(FULL, x0);
    clut_src.y0 = __reduce_min_sync(FULL, y0);
    clut_src.x1 = __reduce_max_sync(FULL, x1);
    clut_src.y1 = __reduce_max_sync(FULL, y1);
    __syncwarp();
  }

  __device__ __forceinline__ void invalidate_cluts(const Rects& w)
  {
    bool any = false;
    for (uint32_t s = lane; s < slots; s += 32)
    {
      const uint32_t k = key[s];
      if (!(k & ENC_SLOT_VALID))
        continue;
      Rects src;
      slot_rects(k, src);
      if (rs_intersects(src, w))
      {
        key[s] = k & ~ENC_SLOT_VALID; // contents stay in use by draws that already reference the slot
        any = true;
      }
    }
    __syncwarp();
    if (__any_sync(FULL, any))
      rebuild_clut_sources();
  }

  static __device__ __forceinline__ uint32_t quick_index(uint32_t want) { return ((want >> 4) ^ (want >> 10) ^ (want >> 15)) & 63u; }

  // The quick-table half of find_slot on its own: lane-private (no collectives), for probing several palettes at once.
  __device__ __forceinline__ uint32_t quick_probe(uint32_t x, uint32_t y, bool is8) const
  {
    const uint32_t want = x | (y << 10);
    const uint32_t s = quick[quick_index(want)], k = key[s];
    return ((k & ENC_SLOT_VALID) && (k & 0x7FFFFu) == want && (((k >> 19) & 1u) || !is8)) ? s : 0xFFFFFFFFu;
  }

  __device__ __forceinline__ uint32_t find_slot(uint32_t x, uint32_t y, bool is8) const
  {
    const uint32_t want = x | (y << 10);
    // Palettes repeat (946 of the 958 CLUT updates of a BASELINE config-4 stream hit the cache): a 64-entry table of
    // "slot this position was last found in" answers those with two shared-memory loads instead of a table search.
    const uint32_t qi = quick_index(want);
    {
      const uint32_t s = quick[qi], k = key[s];
      if ((k & ENC_SLOT_VALID) && (k & 0x7FFFFu) == want && (((k >> 19) & 1u) || !is8))
        return s;
    }
    uint32_t best = 0xFFFFFFFFu;
    for (uint32_t s = lane; s < slots; s += 32)
    {
      const uint32_t k = key[s];
      // a valid 8-bit snapshot at the same position also serves a 4-bit load
      if ((k & ENC_SLOT_VALID) && (k & 0x7FFFFu) == want && (((k >> 19) & 1u) || !is8))
        best = min(best, s);
    }
    best = __reduce_min_sync(FULL, best);
    if (best != 0xFFFFFFFFu)
    {
      __syncwarp();
      if (lane == 0)
        quick[qi] = static_cast<uint8_t>(best);
      __syncwarp();
    }
    return best;
  }

  __device__ __forceinline__ uint32_t alloc_slot()
  {
    for (int pass = 0; pass < 2; pass++)
    {
      uint32_t best = 0xFFFFFFFFu;
      for (uint32_t s = lane; s < slots; s += 32)
        if (s >= 1 && s != cur_lo && s != cur_hi && last[s] != epoch_serial)
          best = min(best, (s >= hand) ? (s - hand) : (s + (slots - 1) - hand));
      best = __reduce_min_sync(FULL, best);
      if (best != 0xFFFFFFFFu)
      {
        uint32_t s = hand + best;
        if (s >= slots)
          s -= (slots - 1);
        hand = (s + 1 >= slots) ? 1 : (s + 1);
        return s;
      }
      // every slot is in use by the open epoch: cut, which frees all but the current pair
      close_epoch(EPOCH_TILED);
      st_clutcut++;
    }
    return (cur_lo == 1 || cur_hi == 1) ? ((cur_lo == 2 || cur_hi == 2) ? 3u : 2u) : 1u;
  }

  // check_writes = false: the caller has already established that the source row is not written in the open epoch
  __device__ __forceinline__ void update_clut(uint32_t reg, bool is8, bool check_writes = true)
  {
    const uint32_t x = PSX_PAL_X(reg), y = PSX_PAL_Y(reg);
    uint32_t slot = find_slot(x, y, is8);
    if (slot != 0xFFFFFFFFu)
    {
      st_hits++;
    }
    else
    {
      Rects src;
      src.n = 0;
      rs_add_wrapped(src, x, y, is8 ? 256u : 16u, 1);
      // the snapshot is taken before the epoch rasterises, so only writes EARLIER in this epoch matter
      bool hit = false;
      if (check_writes && hz_maybe(writes, src))
      {
        ensure_grids();
        hit = hz_grid(writes, src, lane);
      }
      if (hit)
      {
        close_epoch(EPOCH_TILED);
        st_clutcut++;
      }
      slot = alloc_slot();
      __syncwarp();
      if (lane == 0)
      {
        key[slot] = x | (y << 10) | (is8 ? (1u << 19) : 0u) | ENC_SLOT_VALID;
        quick[quick_index(x | (y << 10))] = static_cast<uint8_t>(slot);
      }
      __syncwarp();
      hz_add(clut_src, src, lane);
      if (n_clut < clut_cap)
      {
        if (lane == 0)
        {
          ClutOp op;
          op.dst_slot = static_cast<uint16_t>(slot);
          op.x = static_cast<uint16_t>(x);
          op.y = static_cast<uint16_t>(y);
          op.is8bit = is8 ? 1 : 0;
          clut_ops[n_clut] = op;
        }
      }
      else
      {
        overflow = 1;
      }
      n_clut++;
      st_clutops++;
    }
    if (lane == 0)
      last[slot] = epoch_serial;
    __syncwarp();
    cur_lo = slot;
    if (is8)
      cur_hi = slot;
  }

  // Hazard logic of one primitive (Encoder::emit): returns the command slot it goes to (NO_SLOT on overflow); the
  // caller stores the record.
  static constexpr uint32_t NO_SLOT = 0xFFFFFFFFu;
  __device__ __forceinline__ uint32_t take_slot()
  {
    const uint32_t idx = (n_cmds < cmd_cap) ? n_cmds : NO_SLOT;
    if (idx == NO_SLOT)
      overflow = 1;
    n_cmds++;
    st_prims++;
    return idx;
  }

  __device__ __forceinline__ uint4 side_entry(const Rects& wr, const Rects& rd)
  {
    // single rectangles go to the side buffer (lazy grid); anything wrapped is put into the grids right away
    uint4 e = make_uint4(0xFFFFFFFFu, 0, 0xFFFFFFFFu, 0);
    if (wr.n == 1)
    {
      e.x = wr.x0[0] | (wr.y0[0] << 16);
      e.y = wr.x1[0] | (wr.y1[0] << 16);
    }
    else if (wr.n > 1)
    {
      writes.col |= rs_column(wr, lane);
    }
    if (rd.n == 1)
    {
      e.z = rd.x0[0] | (rd.y0[0] << 16);
      e.w = rd.x1[0] | (rd.y1[0] << 16);
    }
    else if (rd.n > 1)
    {
      reads.col |= rs_column(rd, lane);
    }
    return e;
  }

  __device__ __forceinline__ uint32_t emit_slot(const Rects& wr, const Rects& rd, uint32_t kind, bool textured, uint4& side)
  {
    side = make_uint4(0xFFFFFFFFu, 0, 0xFFFFFFFFu, 0);
    if (hz_hits(clut_src, wr, lane))
      invalidate_cluts(wr);
    if (kind != EPOCH_TILED)
    {
      if (n_cmds > epoch_cmd_begin)
        close_epoch(EPOCH_TILED);
      if (textured && lane == 0)
        last[cur_lo] = last[cur_hi] = epoch_serial;
      __syncwarp();
      const uint32_t idx = take_slot();
      close_epoch(kind);
      return idx;
    }
    const bool raw_maybe = hz_maybe(writes, rd);
    const bool war_maybe = hz_maybe(reads, wr);
    bool raw = false, war = false;
    if (raw_maybe || war_maybe)
    {
      ensure_grids();
      raw = raw_maybe && hz_grid(writes, rd, lane);
      war = !raw && war_maybe && hz_grid(reads, wr, lane);
    }
    const bool full = (n_cmds - epoch_cmd_begin) >= 65535u;
    if (raw || war || full)
    {
      close_epoch(EPOCH_TILED);
      if (raw)
        st_raw++;
      else if (war)
        st_war++;
    }
    if (textured && lane == 0)
      last[cur_lo] = last[cur_hi] = epoch_serial;
    __syncwarp();
    const uint32_t idx = take_slot();
    hz_add_box(writes, wr);
    hz_add_box(reads, rd);
    side = side_entry(wr, rd);
    return idx;
  }

  __device__ __forceinline__ void emit(const uint32_t (&w)[16], const Rects& wr, const Rects& rd, uint32_t kind, bool textured)
  {
    uint4 side;
    const uint32_t idx = emit_slot(wr, rd, kind, textured, side);
    if (idx != NO_SLOT)
    {
      store_cmd(cmds + idx, w, lane);
      if (lane == 0)
        rects[idx] = side;
    }
  }

  __device__ __forceinline__ void header_words(uint32_t (&w)[16], uint32_t op, uint32_t wire_flags0, uint32_t wire_flags1, uint32_t draw_mode,
                               uint32_t window) const
  {
#pragma unroll
    for (int i = 0; i < 16; i++)
      w[i] = 0;
    const uint32_t f1 = ((wire_flags1 & PSX_DF1_DITHER) ? F1_DITHER : 0u) |
                        ((mod_crop && (wire_flags0 & PSX_DF_TEXTURE) && !(wire_flags0 & PSX_DF_RAW_TEXTURE)) ? F1_MOD5 : 0u);
    w[0] = op | (wire_flags0 << 8) | (f1 << 16);
    w[1] = draw_mode | (cur_lo << 16) | (cur_hi << 24);
    w[2] = window;
    w[3] = area[0] | (area[1] << 16);
    w[4] = area[2] | (area[3] << 16);
  }

  // VRAM rectangles a textured primitive may sample, from its (pre-window) texcoord range
  __device__ __forceinline__ void texture_footprint(uint32_t draw_mode, uint32_t window, uint32_t umin, uint32_t umax, uint32_t vmin, uint32_t vmax,
                                    Rects& rd) const
  {
    const uint32_t and_x = window & 0xFFu, and_y = (window >> 8) & 0xFFu, or_x = (window >> 16) & 0xFFu, or_y = window >> 24;
    if (and_x != 0xFF || or_x != 0)
    {
      umin = or_x;
      umax = or_x | and_x;
    }
    if (and_y != 0xFF || or_y != 0)
    {
      vmin = or_y;
      vmax = or_y | and_y;
    }
    const uint32_t mode = (draw_mode >> 7) & 3u;
    const uint32_t sh = (mode == 0) ? 2u : (mode == 1 ? 1u : 0u);
    const uint32_t bx = (draw_mode & 0xFu) * 64u, by = ((draw_mode >> 4) & 1u) * 256u;
    rs_add_wrapped(rd, bx + (umin >> sh), by + vmin, (umax >> sh) - (umin >> sh) + 1, vmax - vmin + 1);
  }

};

__device__ __forceinline__ bool vertex_ok(int32_t x, int32_t y)
{
  return x >= -32768 && x <= 32767 && y >= -32768 && y <= 32767;
}

#ifdef PSX_ENC_MAXREG
__global__ void __maxnreg__(PSX_ENC_MAXREG) encode_kernel(EncodeArgs A)
#else
__global__ void __launch_bounds__(32 * ENC_WARPS) encode_kernel(EncodeArgs A)
#endif
{
  __shared__ uint32_t sh_key[ENC_WARPS][256];
  __shared__ uint32_t sh_last[ENC_WARPS][256];
  __shared__ unsigned long long sh_col[ENC_WARPS][32];
  __shared__ unsigned long long sh_col2[ENC_WARPS][32];
  __shared__ uint8_t sh_quick[ENC_WARPS][64];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t idx = blockIdx.x * ENC_WARPS + warp;
  if (idx >= A.n_inst)
    return;
  const EncInst ii = A.inst[idx];
  const EncState* st = A.state_in + idx;
  EncState* st_out = A.state_out + idx;

  WarpEncoder E;
  E.lane = lane;
  E.cmds = A.cmds + ii.cmd_base;
  E.clut_ops = A.clut_ops + ii.clut_base;
  E.epochs = A.epochs + ii.epoch_base;
  E.cmd_cap = ii.cmd_cap;
  E.clut_cap = ii.clut_cap;
  E.epoch_cap = ii.epoch_cap;
  E.n_cmds = E.n_clut = E.n_epochs = E.overflow = 0;
  E.epoch_cmd_begin = E.epoch_clut_begin = E.epoch_serial = 0;
  for (int i = 0; i < 4; i++)
    E.area[i] = st->area[i];
  E.cur_lo = st->cur_lo;
  E.cur_hi = st->cur_hi;
  E.hand = st->alloc_hand;
  E.slots = A.clut_slots;
  E.mod_crop = A.modulation_crop;
  E.key = sh_key[warp];
  E.last = sh_last[warp];
  E.colscratch = sh_col[warp];
  E.colscratch2 = sh_col2[warp];
  E.quick = sh_quick[warp];
  E.quick[lane] = 0;
  E.quick[lane + 32] = 0;
  E.rects = A.rects + ii.cmd_base;
  E.grid_upto = 0;
  E.st_prims = E.st_raw = E.st_war = E.st_clutcut = E.st_self = E.st_copy = E.st_upload = E.st_clutops = E.st_hits = E.st_rejected = 0;
  for (uint32_t s = lane; s < 256; s += 32)
  {
    E.key[s] = (s < E.slots) ? st->slot_key[s] : 0u;
    E.last[s] = 0xFFFFFFFFu;
  }
  hz_clear(E.writes);
  E.reads = E.writes;
  E.clut_src = E.writes;
  __syncwarp();
  E.rebuild_clut_sources();

  const uint32_t* wire32 = reinterpret_cast<const uint32_t*>(A.wire);
  // Records that the lane-parallel parser below does not take (precise records, poly-lines, transfers, rectangles
  // that wrap, unaligned records) go through this warp-uniform path: one coalesced 128-byte load of the record,
  // fields handed round with shuffles.
  auto generic_record = [&](const uint32_t cur_off) {
    const uint32_t* rec = wire32 + (cur_off >> 2);
    const uint32_t cur = __ldg(rec + lane); // the wire buffer has slack at its end
    auto W = [&](uint32_t k) -> uint32_t { return (k < 32u) ? __shfl_sync(FULL, cur, k) : __ldg(rec + k); };
    const uint32_t size = W(0), type = W(1) & 0xFFu;

    // state records and the number of primitives the record holds
    uint32_t nsub = 0, w2 = 0, draw_mode = 0, window = 0, nv = 0;
    bool precise = false;
    switch (type)
    {
      case PSX_CMD_SET_DRAWING_AREA:
        if (size >= 24)
        {
          E.area[0] = min(W(2), 2047u);
          E.area[1] = min(W(3), 2047u);
          E.area[2] = min(W(4), 1023u);
          E.area[3] = min(W(5), 1023u);
        }
        break;
      case PSX_CMD_UPDATE_CLUT:
        if (size >= 12)
        {
          const uint32_t v = W(2);
          E.update_clut(v & 0xFFFFu, ((v >> 16) & 0xFFu) != 0);
        }
        break;
      case PSX_CMD_DRAW_PRECISE_POLYGON: precise = true; // fall through
      case PSX_CMD_DRAW_POLYGON:
      {
        w2 = W(2);
        draw_mode = W(3) & 0xFFFFu;
        window = W(4);
        nv = min(w2 >> 16, 4u);
        const uint32_t need = precise ? (40u + 28u * nv) : (20u + 16u * nv);
        bool ok = nv >= 3 && need <= size;
        for (uint32_t i = 0; i < nv && ok; i++)
        {
          const uint32_t b = precise ? (13u + 7u * i) : (5u + 4u * i);
          ok = vertex_ok(static_cast<int32_t>(W(b)), static_cast<int32_t>(W(b + 1)));
        }
        if (ok)
          nsub = (nv > 3) ? 2u : 1u; // gpu_sw.cpp:113-115: (v0,v1,v2) then (v2,v1,v3)
        else
          E.st_rejected++;
      }
      break;
      case PSX_CMD_DRAW_PRECISE_LINE: precise = true; // fall through
      case PSX_CMD_DRAW_LINE:
        w2 = W(2);
        draw_mode = W(3) & 0xFFFFu;
        window = W(4);
        nv = w2 >> 16;
        if (static_cast<uint64_t>(20u) + static_cast<uint64_t>(precise ? 24u : 12u) * nv > size)
          E.st_rejected++;
        else
          nsub = nv / 2u;
        break;
      case PSX_CMD_DRAW_RECTANGLE:
      case PSX_CMD_FILL_VRAM:
      case PSX_CMD_COPY_VRAM:
      case PSX_CMD_UPDATE_VRAM: nsub = 1; break;
      default: break; // ClearCache / BufferSwapped / SubmitFrame / ClearDisplay: no VRAM effect
    }

    for (uint32_t j = 0; j < nsub; j++)
    {
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; i++)
        w[i] = 0;
      Rects wr, rd;
      wr.n = rd.n = 0;
      uint32_t kind = EPOCH_TILED;
      bool textured = false, do_emit = false;
      switch (type)
      {
        case PSX_CMD_DRAW_POLYGON:
        case PSX_CMD_DRAW_PRECISE_POLYGON:
        {
          const uint32_t flags0 = w2 & 0xFFu, flags1 = (w2 >> 8) & 0xFFu;
          const uint32_t vi[3] = {j ? 2u : 0u, 1u, j ? 3u : 2u};
          Vtx v[3];
#pragma unroll
          for (int i = 0; i < 3; i++)
          {
            const uint32_t b = precise ? (13u + 7u * vi[i]) : (5u + 4u * vi[i]);
            v[i].x = static_cast<int32_t>(W(b));
            v[i].y = static_cast<int32_t>(W(b + 1));
            v[i].rgb = W(b + 2) & 0xFFFFFFu;
            const uint32_t tc = W(b + 3);
            v[i].u = tc & 0xFFu;
            v[i].v = (tc >> 8) & 0xFFu;
          }
          const Vtx &a = v[0], &b = v[1], &c = v[2];
          const int32_t minx = min(a.x, min(b.x, c.x)), maxx = max(a.x, max(b.x, c.x));
          const int32_t miny = min(a.y, min(b.y, c.y)), maxy = max(a.y, max(b.y, c.y));
          if ((maxx - minx) >= static_cast<int32_t>(VRAM_W) || (maxy - miny) >= static_cast<int32_t>(VRAM_H))
          {
            E.st_rejected++; // the decoder culls these (gpu.cpp:3283-3350); a backend never sees them
            break;
          }
          if (miny == maxy)
            break; // v0.y == v2.y: rejected by the rasteriser (inl:1459)
          const int32_t cl = static_cast<int32_t>(E.area[0]), ct = static_cast<int32_t>(E.area[1]);
          const int32_t cr = static_cast<int32_t>(E.area[2]), cb = static_cast<int32_t>(E.area[3]);
          int32_t x0 = cl, x1 = cr, y0 = 0, y1 = static_cast<int32_t>(VRAM_H) - 1;
          if (minx > -1023 && maxx < 1023)
          {
            x0 = max(minx - 1, cl);
            x1 = min(maxx + 1, cr);
          }
          if (miny >= -1024 && maxy <= 1024)
          {
            y0 = max(miny, ct);
            y1 = min(maxy - 1, cb);
          }
          if (x1 < x0 || y1 < y0)
            break; // nothing can be drawn
          E.header_words(w, OP_TRIANGLE, flags0, flags1, draw_mode, window);
          w[5] = (static_cast<uint32_t>(a.x) & 0xFFFFu) | (static_cast<uint32_t>(a.y) << 16);
          w[6] = a.rgb | (a.u << 24);
          w[7] = a.v | ((static_cast<uint32_t>(b.x) & 0xFFFFu) << 16);
          w[8] = (static_cast<uint32_t>(b.y) & 0xFFFFu) | ((b.rgb & 0xFFFFu) << 16);
          w[9] = (b.rgb >> 16) | (b.u << 8) | (b.v << 16);
          w[10] = (static_cast<uint32_t>(c.x) & 0xFFFFu) | (static_cast<uint32_t>(c.y) << 16);
          w[11] = c.rgb | (c.u << 24);
          w[12] = c.v;
          rs_add_wrapped(wr, static_cast<uint32_t>(x0), static_cast<uint32_t>(y0), static_cast<uint32_t>(x1 - x0 + 1),
                         static_cast<uint32_t>(y1 - y0 + 1));
          textured = (flags0 & PSX_DF_TEXTURE) != 0;
          if (textured)
          {
            uint32_t umin = min(a.u, min(b.u, c.u)), umax = max(a.u, max(b.u, c.u));
            uint32_t vmin = min(a.v, min(b.v, c.v)), vmax = max(a.v, max(b.v, c.v));
            // interpolated coordinates are u8 and may wrap by one step at the extremes
            if (umin == 0 || umax == 255)
            {
              umin = 0;
              umax = 255;
            }
            else
            {
              umin--;
              umax++;
            }
            if (vmin == 0 || vmax == 255)
            {
              vmin = 0;
              vmax = 255;
            }
            else
            {
              vmin--;
              vmax++;
            }
            E.texture_footprint(draw_mode, window, umin, umax, vmin, vmax, rd);
            if (rs_intersects(rd, wr))
            {
              kind = EPOCH_SELF_SAMPLING;
              E.st_self++;
            }
          }
          do_emit = true;
        }
        break;

        case PSX_CMD_DRAW_RECTANGLE:
        {
          if (size < 40)
          {
            E.st_rejected++;
            break;
          }
          const uint32_t r2 = W(2), dm = W(3) & 0xFFFFu, win = W(4), wh = W(5), texcoord = W(6) & 0xFFFFu, color = W(9);
          const int32_t x = static_cast<int32_t>(W(7)), y = static_cast<int32_t>(W(8));
          const uint32_t flags0 = r2 & 0xFFu, flags1 = (r2 >> 8) & 0xFFu;
          const uint32_t rw = wh & 0xFFFFu, rh = wh >> 16;
          if (!vertex_ok(x, y) || rw > VRAM_W || rh > VRAM_H)
          {
            E.st_rejected++;
            break;
          }
          const int32_t x0 = max(x, static_cast<int32_t>(E.area[0])), x1 = min(x + static_cast<int32_t>(rw) - 1, static_cast<int32_t>(E.area[2]));
          const int32_t y0 = max(y, static_cast<int32_t>(E.area[1])), y1 = min(y + static_cast<int32_t>(rh) - 1, static_cast<int32_t>(E.area[3]));
          if (x1 < x0 || y1 < y0)
            break;
          const uint32_t u0 = texcoord & 0xFFu, v0 = texcoord >> 8;
          E.header_words(w, OP_RECT, flags0, flags1, dm, win);
          w[5] = (static_cast<uint32_t>(x) & 0xFFFFu) | (static_cast<uint32_t>(y) << 16);
          w[6] = rw | (rh << 16);
          w[7] = (color & 0xFFFFFFu) | (u0 << 24);
          w[8] = v0;
          rs_add_wrapped(wr, static_cast<uint32_t>(x0), static_cast<uint32_t>(y0), static_cast<uint32_t>(x1 - x0 + 1),
                         static_cast<uint32_t>(y1 - y0 + 1));
          textured = (flags0 & PSX_DF_TEXTURE) != 0;
          if (textured)
          {
            uint32_t umin = u0, umax = u0 + rw - 1, vmin = v0, vmax = v0 + rh - 1;
            if (umax > 255)
            {
              umin = 0;
              umax = 255;
            }
            if (vmax > 255)
            {
              vmin = 0;
              vmax = 255;
            }
            E.texture_footprint(dm, win, umin, umax, vmin, vmax, rd);
            if (rs_intersects(rd, wr))
            {
              kind = EPOCH_SELF_SAMPLING;
              E.st_self++;
            }
          }
          do_emit = true;
        }
        break;

        case PSX_CMD_DRAW_LINE:
        case PSX_CMD_DRAW_PRECISE_LINE:
        {
          // lines are never textured (gpu.cpp:2974)
          const uint32_t flags0 = (w2 & 0xFFu) & ~static_cast<uint32_t>(PSX_DF_TEXTURE | PSX_DF_RAW_TEXTURE), flags1 = (w2 >> 8) & 0xFFu;
          int32_t px[2], py[2];
          uint32_t pc[2];
#pragma unroll
          for (uint32_t q = 0; q < 2; q++)
          {
            const uint32_t b = precise ? (5u + 6u * (2u * j + q) + 3u) : (5u + 3u * (2u * j + q));
            px[q] = static_cast<int32_t>(W(b));
            py[q] = static_cast<int32_t>(W(b + 1));
            pc[q] = W(b + 2) & 0xFFFFFFu;
          }
          if (!vertex_ok(px[0], py[0]) || !vertex_ok(px[1], py[1]))
          {
            E.st_rejected++;
            break;
          }
          if (abs(px[1] - px[0]) >= static_cast<int32_t>(VRAM_W) || abs(py[1] - py[0]) >= static_cast<int32_t>(VRAM_H))
            break; // rasteriser rejects (inl:834-835)
          const int32_t minx = min(px[0], px[1]), maxx = max(px[0], px[1]), miny = min(py[0], py[1]), maxy = max(py[0], py[1]);
          int32_t x0 = static_cast<int32_t>(E.area[0]), x1 = static_cast<int32_t>(E.area[2]), y0 = 0, y1 = static_cast<int32_t>(VRAM_H) - 1;
          if (minx >= -1024 && maxx <= 1023 && miny >= -1024 && maxy <= 1023)
          {
            x0 = max(minx, static_cast<int32_t>(E.area[0]));
            x1 = min(maxx, static_cast<int32_t>(E.area[2]));
            y0 = max(miny, static_cast<int32_t>(E.area[1]));
            y1 = min(maxy, static_cast<int32_t>(E.area[3]));
          }
          if (x1 < x0 || y1 < y0)
            break;
          E.header_words(w, OP_LINE, flags0, flags1, draw_mode, window);
          w[5] = (static_cast<uint32_t>(px[0]) & 0xFFFFu) | (static_cast<uint32_t>(py[0]) << 16);
          w[6] = pc[0];
          w[7] = (static_cast<uint32_t>(px[1]) & 0xFFFFu) << 16;
          w[8] = (static_cast<uint32_t>(py[1]) & 0xFFFFu) | ((pc[1] & 0xFFFFu) << 16);
          w[9] = pc[1] >> 16;
          rs_add_wrapped(wr, static_cast<uint32_t>(x0), static_cast<uint32_t>(y0), static_cast<uint32_t>(x1 - x0 + 1),
                         static_cast<uint32_t>(y1 - y0 + 1));
          do_emit = true;
        }
        break;

        case PSX_CMD_FILL_VRAM:
        {
          if (size < 24)
          {
            E.st_rejected++;
            break;
          }
          const uint32_t xy = W(2), wh = W(3), color = W(4), il = W(5);
          const uint32_t x = xy & 0xFFFFu, y = xy >> 16, fw = wh & 0xFFFFu, fh = wh >> 16;
          if (fw > VRAM_W || fh > VRAM_H || x >= VRAM_W || y >= VRAM_H)
          {
            E.st_rejected++;
            break;
          }
          if (fw == 0 || fh == 0)
            break;
          w[0] = OP_FILL;
          w[5] = xy;
          w[6] = wh;
          // VRAMRGBA8888ToRGBA5551, gpu_helpers.h:32-39
          const uint32_t c16 = ((color & 0xFFu) >> 3) | ((((color >> 8) & 0xFFu) >> 3) << 5) | ((((color >> 16) & 0xFFu) >> 3) << 10) |
                               (((color >> 24) & 1u) << 15);
          w[7] = c16 | (((il & 0xFFu) ? 1u : 0u) << 16) | (((il >> 8) & 1u) << 24);
          rs_add_wrapped(wr, x, y, fw, fh);
          do_emit = true;
        }
        break;

        case PSX_CMD_COPY_VRAM:
        {
          if (size < 24)
          {
            E.st_rejected++;
            break;
          }
          const uint32_t s = W(2), d = W(3), wh = W(4), m = W(5);
          const uint32_t sx = s & 0xFFFFu, sy = s >> 16, dx = d & 0xFFFFu, dy = d >> 16, cw = wh & 0xFFFFu, ch = wh >> 16;
          if (cw > VRAM_W || ch > VRAM_H || sx >= VRAM_W || dx >= VRAM_W || sy >= VRAM_H || dy >= VRAM_H)
          {
            E.st_rejected++;
            break;
          }
          if (cw == 0 || ch == 0)
            break;
          const uint32_t f0 = ((m & 0xFFu) ? F_SET_MASK : 0u) | (((m >> 8) & 0xFFu) ? F_CHECK_MASK : 0u);
          w[0] = OP_COPY | (f0 << 8);
          w[5] = s;
          w[6] = d;
          w[7] = wh;
          rs_add_wrapped(wr, dx, dy, cw, ch);
          rs_add_wrapped(rd, sx, sy, cw, ch);
          if (rs_intersects(rd, wr))
          {
            kind = EPOCH_COPY_OVERLAP;
            E.st_copy++;
          }
          do_emit = true;
        }
        break;

        case PSX_CMD_UPDATE_VRAM:
        {
          const uint32_t xy = W(2), wh = W(3), m = W(4);
          const uint32_t x = xy & 0xFFFFu, y = xy >> 16, uw = wh & 0xFFFFu, uh = wh >> 16;
          const uint64_t need = PSX_UPDATE_VRAM_DATA_OFFSET + static_cast<uint64_t>(uw) * uh * 2u;
          if (size < 20 || uw > VRAM_W || uh > VRAM_H || x >= VRAM_W || y >= VRAM_H || need > size)
          {
            E.st_rejected++;
            break;
          }
          if (uw == 0 || uh == 0)
            break;
          const bool set_mask = (m & 0xFFu) != 0, check_mask = ((m >> 8) & 0xFFu) != 0;
          w[0] = OP_UPLOAD | (((set_mask ? F_SET_MASK : 0u) | (check_mask ? F_CHECK_MASK : 0u)) << 8);
          w[5] = xy;
          w[6] = wh;
          w[7] = (cur_off + PSX_UPDATE_VRAM_DATA_OFFSET) >> 1; // the pixels stay where they are: inside the uploaded wire buffer
          rs_add_wrapped(wr, x, y, uw, uh);
          // Q2: with mask checking, source pixels are consumed only by written pixels in the scalar tail of each row
          const uint32_t aw = min(uw, VRAM_W - x) & ~7u;
          if (check_mask && aw != uw)
          {
            kind = EPOCH_UPLOAD_MASKED;
            E.st_upload++;
          }
          do_emit = true;
        }
        break;

        default: break;
      }
      if (do_emit)
        E.emit(w, wr, rd, kind, textured);
    }
    };

  // Main loop: up to 32 records at a time.  Phase 1 (lane-parallel, one record per lane): fetch, classify and fully
  // parse the common stateless cases — triangles / quads, rectangles, two-point lines — into packed words and
  // footprints.  Phase 2 (warp-uniform, in stream order): the stateful part only — drawing area, CLUT cache, hazard
  // sets, epoch cuts, output slots.  A SetDrawingArea record ends the batch because later records clip against it.
  enum : uint32_t
  {
    CLS_IGNORE = 0,
    CLS_FAST = 1,
    CLS_SLOW = 2,
    CLS_AREA = 3,
    CLS_CLUT = 4,
    FP_EMIT = 1,
    FP_TEXTURED = 2,
    FP_SELF = 4,
    FP_REJECTED = 8
  };
  // The framing was checked by walk_kernel, which also left the byte offset of every record: lane i takes record
  // base + i.
  const uint32_t n_rec = A.walk[idx].n_rec;
  const uint32_t* rec_off = A.rec_off + ii.rec_base;
  uint32_t base = 0, n_seen = 0;
  while (base < n_rec)
  {
    uint32_t nb = min(32u, n_rec - base);
    const bool have = lane < nb;
    const uint32_t my_off = have ? rec_off[base + lane] : 0u;
    uint32_t rw[24];
    {
      const uint4* q = reinterpret_cast<const uint4*>(A.wire + (my_off & ~15u));
#pragma unroll
      for (int i = 0; i < 6; i++)
      {
        const uint4 v = have ? __ldg(q + i) : make_uint4(0, 0, 0, 0); // the wire buffer has slack behind its last stream
        rw[4 * i + 0] = v.x;
        rw[4 * i + 1] = v.y;
        rw[4 * i + 2] = v.z;
        rw[4 * i + 3] = v.w;
      }
    }
    // The streams come straight from the H2D copy, i.e. from HBM: ask for the next batch's records now (two lines per
    // record at most), so that its loads find them in L2 when this batch has been committed.
    if (base + 32u + lane < n_rec)
    {
      const uint8_t* nx = A.wire + (rec_off[base + 32u + lane] & ~15u);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 95));
    }
    uint32_t cls = CLS_IGNORE, nprim = 0, hw0 = 0, hdm = 0, hwin = 0;
    uint32_t pv[2][8], pwr0[2], pwr1[2], prd0[2], prd1[2], pf[2];
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
#pragma unroll
      for (int i = 0; i < 8; i++)
        pv[j][i] = 0;
      pwr0[j] = pwr1[j] = prd0[j] = prd1[j] = pf[j] = 0;
    }
    // records that are not 16-byte aligned (never produced by the reference, whose sizes are multiples of 16) send
    // the whole batch down the uniform path, where the drawing area is tracked record by record
    const bool aligned = __all_sync(FULL, !have || (my_off & 15u) == 0);
    if (have && !aligned)
      cls = CLS_SLOW;
    if (have && aligned)
    {
      const uint32_t size = rw[0], type = rw[1] & 0xFFu;
      const int32_t cl = static_cast<int32_t>(E.area[0]), ct = static_cast<int32_t>(E.area[1]);
      const int32_t cr = static_cast<int32_t>(E.area[2]), cb = static_cast<int32_t>(E.area[3]);
      const uint32_t flags0 = rw[2] & 0xFFu, flags1 = (rw[2] >> 8) & 0xFFu;
      const uint32_t f1 = ((flags1 & PSX_DF1_DITHER) ? F1_DITHER : 0u) |
                          ((E.mod_crop && (flags0 & PSX_DF_TEXTURE) && !(flags0 & PSX_DF_RAW_TEXTURE)) ? F1_MOD5 : 0u);
      hdm = rw[3] & 0xFFFFu;
      hwin = rw[4];
      // texture footprint of a u/v range as one rectangle; false if it wraps (slow path)
      auto footprint = [&](uint32_t umin, uint32_t umax, uint32_t vmin, uint32_t vmax, uint32_t& r0, uint32_t& r1) -> bool {
        const uint32_t and_x = hwin & 0xFFu, and_y = (hwin >> 8) & 0xFFu, or_x = (hwin >> 16) & 0xFFu, or_y = hwin >> 24;
        if (and_x != 0xFF || or_x != 0)
        {
          umin = or_x;
          umax = or_x | and_x;
        }
        if (and_y != 0xFF || or_y != 0)
        {
          vmin = or_y;
          vmax = or_y | and_y;
        }
        const uint32_t mode = (hdm >> 7) & 3u;
        const uint32_t sh = (mode == 0) ? 2u : (mode == 1 ? 1u : 0u);
        const uint32_t bx = (hdm & 0xFu) * 64u, by = ((hdm >> 4) & 1u) * 256u;
        const uint32_t x0 = bx + (umin >> sh), x1 = bx + (umax >> sh), y0 = by + vmin, y1 = by + vmax;
        r0 = x0 | (y0 << 16);
        r1 = x1 | (y1 << 16);
        return x1 < VRAM_W && y1 < VRAM_H;
      };
      auto overlap = [](uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) -> bool {
        return (a0 & 0xFFFFu) <= (b1 & 0xFFFFu) && (b0 & 0xFFFFu) <= (a1 & 0xFFFFu) && (a0 >> 16) <= (b1 >> 16) && (b0 >> 16) <= (a1 >> 16);
      };
      if (type == PSX_CMD_SET_DRAWING_AREA)
      {
        cls = (size >= 24) ? CLS_AREA : CLS_IGNORE;
        pv[0][0] = min(rw[2], 2047u);
        pv[0][1] = min(rw[3], 2047u);
        pv[0][2] = min(rw[4], 1023u);
        pv[0][3] = min(rw[5], 1023u);
      }
      else if (type == PSX_CMD_UPDATE_CLUT)
      {
        cls = (size >= 12) ? CLS_CLUT : CLS_IGNORE;
        pv[0][0] = rw[2];
      }
      else if (type == PSX_CMD_DRAW_POLYGON)
      {
        const uint32_t nv = min(rw[2] >> 16, 4u);
        bool ok = nv >= 3 && (20u + 16u * nv) <= size;
#pragma unroll
        for (uint32_t i = 0; i < 4; i++)
          if (i < nv)
            ok = ok && vertex_ok(static_cast<int32_t>(rw[5 + 4 * i]), static_cast<int32_t>(rw[6 + 4 * i]));
        cls = CLS_FAST;
        if (!ok)
        {
          nprim = 1;
          pf[0] = FP_REJECTED;
        }
        else
        {
          nprim = (nv > 3) ? 2u : 1u;
          hw0 = OP_TRIANGLE | (flags0 << 8) | (f1 << 16);
#pragma unroll
          for (uint32_t j = 0; j < 2; j++)
          {
            if (j < nprim)
            {
              // gpu_sw.cpp:113-115: (v0,v1,v2) then (v2,v1,v3)
              const uint32_t ia = j ? 2u : 0u, ib = 1u, ic = j ? 3u : 2u;
              const int32_t ax = static_cast<int32_t>(rw[5 + 4 * ia]), ay = static_cast<int32_t>(rw[6 + 4 * ia]);
              const int32_t bx = static_cast<int32_t>(rw[5 + 4 * ib]), by = static_cast<int32_t>(rw[6 + 4 * ib]);
              const int32_t cx = static_cast<int32_t>(rw[5 + 4 * ic]), cy = static_cast<int32_t>(rw[6 + 4 * ic]);
              const uint32_t argb = rw[7 + 4 * ia] & 0xFFFFFFu, brgb = rw[7 + 4 * ib] & 0xFFFFFFu, crgb = rw[7 + 4 * ic] & 0xFFFFFFu;
              const uint32_t au = rw[8 + 4 * ia] & 0xFFu, av = (rw[8 + 4 * ia] >> 8) & 0xFFu;
              const uint32_t bu = rw[8 + 4 * ib] & 0xFFu, bv = (rw[8 + 4 * ib] >> 8) & 0xFFu;
              const uint32_t cu = rw[8 + 4 * ic] & 0xFFu, cv = (rw[8 + 4 * ic] >> 8) & 0xFFu;
              const int32_t minx = min(ax, min(bx, cx)), maxx = max(ax, max(bx, cx));
              const int32_t miny = min(ay, min(by, cy)), maxy = max(ay, max(by, cy));
              if ((maxx - minx) >= static_cast<int32_t>(VRAM_W) || (maxy - miny) >= static_cast<int32_t>(VRAM_H))
              {
                pf[j] = FP_REJECTED; // the decoder culls these (gpu.cpp:3283-3350); a backend never sees them
                continue;
              }
              if (miny == maxy)
                continue; // v0.y == v2.y: rejected by the rasteriser (inl:1459)
              int32_t x0 = cl, x1 = cr, y0 = 0, y1 = static_cast<int32_t>(VRAM_H) - 1;
              if (minx > -1023 && maxx < 1023)
              {
                x0 = max(minx - 1, cl);
                x1 = min(maxx + 1, cr);
              }
              if (miny >= -1024 && maxy <= 1024)
              {
                y0 = max(miny, ct);
                y1 = min(maxy - 1, cb);
              }
              if (x1 < x0 || y1 < y0)
                continue; // nothing can be drawn
              if (x0 < 0 || y0 < 0 || x1 >= static_cast<int32_t>(VRAM_W) || y1 >= static_cast<int32_t>(VRAM_H))
              {
                cls = CLS_SLOW; // footprint wraps
                continue;
              }
              pv[j][0] = (static_cast<uint32_t>(ax) & 0xFFFFu) | (static_cast<uint32_t>(ay) << 16);
              pv[j][1] = argb | (au << 24);
              pv[j][2] = av | ((static_cast<uint32_t>(bx) & 0xFFFFu) << 16);
              pv[j][3] = (static_cast<uint32_t>(by) & 0xFFFFu) | ((brgb & 0xFFFFu) << 16);
              pv[j][4] = (brgb >> 16) | (bu << 8) | (bv << 16);
              pv[j][5] = (static_cast<uint32_t>(cx) & 0xFFFFu) | (static_cast<uint32_t>(cy) << 16);
              pv[j][6] = crgb | (cu << 24);
              pv[j][7] = cv;
              pwr0[j] = static_cast<uint32_t>(x0) | (static_cast<uint32_t>(y0) << 16);
              pwr1[j] = static_cast<uint32_t>(x1) | (static_cast<uint32_t>(y1) << 16);
              pf[j] = FP_EMIT;
              if (flags0 & PSX_DF_TEXTURE)
              {
                uint32_t umin = min(au, min(bu, cu)), umax = max(au, max(bu, cu));
                uint32_t vmin = min(av, min(bv, cv)), vmax = max(av, max(bv, cv));
                // interpolated coordinates are u8 and may wrap by one step at the extremes
                if (umin == 0 || umax == 255)
                {
                  umin = 0;
                  umax = 255;
                }
                else
                {
                  umin--;
                  umax++;
                }
                if (vmin == 0 || vmax == 255)
                {
                  vmin = 0;
                  vmax = 255;
                }
                else
                {
                  vmin--;
                  vmax++;
                }
                if (!footprint(umin, umax, vmin, vmax, prd0[j], prd1[j]))
                  cls = CLS_SLOW;
                pf[j] |= FP_TEXTURED;
                if (overlap(prd0[j], prd1[j], pwr0[j], pwr1[j]))
                  pf[j] |= FP_SELF;
              }
            }
          }
        }
      }
      else if (type == PSX_CMD_DRAW_RECTANGLE)
      {
        cls = CLS_FAST;
        nprim = 1;
        const uint32_t wh = rw[5], texcoord = rw[6] & 0xFFFFu, color = rw[9];
        const int32_t x = static_cast<int32_t>(rw[7]), y = static_cast<int32_t>(rw[8]);
        const uint32_t rwid = wh & 0xFFFFu, rhei = wh >> 16;
        if (size < 40 || !vertex_ok(x, y) || rwid > VRAM_W || rhei > VRAM_H)
        {
          pf[0] = FP_REJECTED;
        }
        else
        {
          const int32_t x0 = max(x, cl), x1 = min(x + static_cast<int32_t>(rwid) - 1, cr);
          const int32_t y0 = max(y, ct), y1 = min(y + static_cast<int32_t>(rhei) - 1, cb);
          if (!(x1 < x0 || y1 < y0))
          {
            if (x0 < 0 || y0 < 0 || x1 >= static_cast<int32_t>(VRAM_W) || y1 >= static_cast<int32_t>(VRAM_H))
              cls = CLS_SLOW;
            const uint32_t u0 = texcoord & 0xFFu, v0 = texcoord >> 8;
            hw0 = OP_RECT | (flags0 << 8) | (f1 << 16);
            pv[0][0] = (static_cast<uint32_t>(x) & 0xFFFFu) | (static_cast<uint32_t>(y) << 16);
            pv[0][1] = rwid | (rhei << 16);
            pv[0][2] = (color & 0xFFFFFFu) | (u0 << 24);
            pv[0][3] = v0;
            pwr0[0] = static_cast<uint32_t>(x0) | (static_cast<uint32_t>(y0) << 16);
            pwr1[0] = static_cast<uint32_t>(x1) | (static_cast<uint32_t>(y1) << 16);
            pf[0] = FP_EMIT;
            if (flags0 & PSX_DF_TEXTURE)
            {
              uint32_t umin = u0, umax = u0 + rwid - 1, vmin = v0, vmax = v0 + rhei - 1;
              if (umax > 255)
              {
                umin = 0;
                umax = 255;
              }
              if (vmax > 255)
              {
                vmin = 0;
                vmax = 255;
              }
              if (!footprint(umin, umax, vmin, vmax, prd0[0], prd1[0]))
                cls = CLS_SLOW;
              pf[0] |= FP_TEXTURED;
              if (overlap(prd0[0], prd1[0], pwr0[0], pwr1[0]))
                pf[0] |= FP_SELF;
            }
          }
        }
      }
      else if (type == PSX_CMD_DRAW_LINE && (rw[2] >> 16) == 2u && size >= 44)
      {
        cls = CLS_FAST;
        nprim = 1;
        // lines are never textured (gpu.cpp:2974)
        const uint32_t lf0 = flags0 & ~static_cast<uint32_t>(PSX_DF_TEXTURE | PSX_DF_RAW_TEXTURE);
        const uint32_t lf1 = (flags1 & PSX_DF1_DITHER) ? F1_DITHER : 0u;
        const int32_t x0p = static_cast<int32_t>(rw[5]), y0p = static_cast<int32_t>(rw[6]);
        const int32_t x1p = static_cast<int32_t>(rw[8]), y1p = static_cast<int32_t>(rw[9]);
        const uint32_t c0 = rw[7] & 0xFFFFFFu, c1 = rw[10] & 0xFFFFFFu;
        if (!vertex_ok(x0p, y0p) || !vertex_ok(x1p, y1p))
        {
          pf[0] = FP_REJECTED;
        }
        else if (!(abs(x1p - x0p) >= static_cast<int32_t>(VRAM_W) || abs(y1p - y0p) >= static_cast<int32_t>(VRAM_H)))
        {
          const int32_t minx = min(x0p, x1p), maxx = max(x0p, x1p), miny = min(y0p, y1p), maxy = max(y0p, y1p);
          int32_t x0 = cl, x1 = cr, y0 = 0, y1 = static_cast<int32_t>(VRAM_H) - 1;
          if (minx >= -1024 && maxx <= 1023 && miny >= -1024 && maxy <= 1023)
          {
            x0 = max(minx, cl);
            x1 = min(maxx, cr);
            y0 = max(miny, ct);
            y1 = min(maxy, cb);
          }
          if (!(x1 < x0 || y1 < y0))
          {
            if (x0 < 0 || y0 < 0 || x1 >= static_cast<int32_t>(VRAM_W) || y1 >= static_cast<int32_t>(VRAM_H))
              cls = CLS_SLOW;
            hw0 = OP_LINE | (lf0 << 8) | (lf1 << 16);
            pv[0][0] = (static_cast<uint32_t>(x0p) & 0xFFFFu) | (static_cast<uint32_t>(y0p) << 16);
            pv[0][1] = c0;
            pv[0][2] = (static_cast<uint32_t>(x1p) & 0xFFFFu) << 16;
            pv[0][3] = (static_cast<uint32_t>(y1p) & 0xFFFFu) | ((c1 & 0xFFFFu) << 16);
            pv[0][4] = c1 >> 16;
            pwr0[0] = static_cast<uint32_t>(x0) | (static_cast<uint32_t>(y0) << 16);
            pwr1[0] = static_cast<uint32_t>(x1) | (static_cast<uint32_t>(y1) << 16);
            pf[0] = FP_EMIT;
          }
        }
      }
      else
      {
        switch (type)
        {
          case PSX_CMD_DRAW_PRECISE_POLYGON:
          case PSX_CMD_DRAW_LINE:
          case PSX_CMD_DRAW_PRECISE_LINE:
          case PSX_CMD_FILL_VRAM:
          case PSX_CMD_COPY_VRAM:
          case PSX_CMD_UPDATE_VRAM: cls = CLS_SLOW; break;
          default: cls = CLS_IGNORE; break; // ClearCache / BufferSwapped / SubmitFrame / ClearDisplay: no VRAM effect
        }
      }
    }
    // a drawing-area record ends the batch: the records after it were clipped against the old area
    const uint32_t area_mask = __ballot_sync(FULL, cls == CLS_AREA);
    if (area_mask)
      nb = static_cast<uint32_t>(__ffs(static_cast<int>(area_mask)));
    base += nb;
    n_seen += nb;
    const uint32_t summary = cls | (nprim << 4) | (pf[0] << 8) | (pf[1] << 16);

    // ---- Batch-parallel commit.  If bounding boxes alone prove that nothing in the batch can conflict — not with the
    // open epoch's read / write sets, not with a CLUT source, not with another record of the batch — the hazard
    // logic has nothing to decide: output slots are a prefix sum, every lane stores its own records, and only the
    // CLUT-cache updates are walked in order.  Anything else takes the record-by-record loop below.
    {
      const bool in_batch = lane < nb;
      const bool fast = in_batch && cls == CLS_FAST;
      const bool e0 = fast && (pf[0] & FP_EMIT), e1 = fast && (pf[1] & FP_EMIT);
      const bool t0 = e0 && (pf[0] & FP_TEXTURED), t1 = e1 && (pf[1] & FP_TEXTURED);
      auto box_of = [&](uint32_t from, uint32_t (&wb)[4], uint32_t (&rb)[4]) {
        uint32_t a[4] = {0xFFFFu, 0xFFFFu, 0u, 0u}, b[4] = {0xFFFFu, 0xFFFFu, 0u, 0u};
        if (lane >= from)
        {
          if (e0)
          {
            a[0] = pwr0[0] & 0xFFFFu, a[1] = pwr0[0] >> 16, a[2] = pwr1[0] & 0xFFFFu, a[3] = pwr1[0] >> 16;
          }
          if (e1)
          {
            a[0] = min(a[0], pwr0[1] & 0xFFFFu), a[1] = min(a[1], pwr0[1] >> 16);
            a[2] = max(a[2], pwr1[1] & 0xFFFFu), a[3] = max(a[3], pwr1[1] >> 16);
          }
          if (t0)
          {
            b[0] = prd0[0] & 0xFFFFu, b[1] = prd0[0] >> 16, b[2] = prd1[0] & 0xFFFFu, b[3] = prd1[0] >> 16;
          }
          if (t1)
          {
            b[0] = min(b[0], prd0[1] & 0xFFFFu), b[1] = min(b[1], prd0[1] >> 16);
            b[2] = max(b[2], prd1[1] & 0xFFFFu), b[3] = max(b[3], prd1[1] >> 16);
          }
        }
        wb[0] = __reduce_min_sync(FULL, a[0]);
        wb[1] = __reduce_min_sync(FULL, a[1]);
        wb[2] = __reduce_max_sync(FULL, a[2]);
        wb[3] = __reduce_max_sync(FULL, a[3]);
        rb[0] = __reduce_min_sync(FULL, b[0]);
        rb[1] = __reduce_min_sync(FULL, b[1]);
        rb[2] = __reduce_max_sync(FULL, b[2]);
        rb[3] = __reduce_max_sync(FULL, b[3]);
      };
      auto boxes_meet = [](const uint32_t (&p)[4], uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) -> bool {
        return p[0] <= p[2] && x0 <= x1 && p[0] <= x1 && x0 <= p[2] && p[1] <= y1 && y0 <= p[3];
      };
      uint32_t wb[4], rb[4];
      box_of(0, wb, rb);
      const uint32_t mine_emit = (e0 ? 1u : 0u) + (e1 ? 1u : 0u);
      uint32_t incl = mine_emit;
#pragma unroll
      for (uint32_t d = 1; d < 32; d <<= 1)
      {
        const uint32_t t = __shfl_up_sync(FULL, incl, d);
        if (lane >= d)
          incl += t;
      }
      const uint32_t rank = incl - mine_emit, total = __shfl_sync(FULL, incl, 31);
      // per-lane obstacles
      bool bad = in_batch && (cls == CLS_SLOW || (fast && ((pf[0] | pf[1]) & FP_SELF)));
      if (in_batch && cls == CLS_CLUT)
      {
        // source row of the snapshot this record may take (whole row if it wraps)
        const uint32_t reg = pv[0][0] & 0xFFFFu, is8 = (pv[0][0] >> 16) & 0xFFu;
        uint32_t sx0 = PSX_PAL_X(reg), sx1 = sx0 + (is8 ? 255u : 15u);
        const uint32_t sy = PSX_PAL_Y(reg);
        if (sx1 >= VRAM_W)
        {
          sx0 = 0;
          sx1 = VRAM_W - 1;
        }
        bad = bad || boxes_meet(wb, sx0, sy, sx1, sy) ||
              (E.writes.any && sx0 <= E.writes.x1 && E.writes.x0 <= sx1 && sy <= E.writes.y1 && E.writes.y0 <= sy);
      }
      bool ok = !__any_sync(FULL, bad);
      ok = ok && !boxes_meet(wb, rb[0], rb[1], rb[2], rb[3]);
      ok = ok && !(E.reads.any && boxes_meet(wb, E.reads.x0, E.reads.y0, E.reads.x1, E.reads.y1));
      ok = ok && !(E.writes.any && boxes_meet(rb, E.writes.x0, E.writes.y0, E.writes.x1, E.writes.y1));
      ok = ok && !(E.clut_src.any && boxes_meet(wb, E.clut_src.x0, E.clut_src.y0, E.clut_src.x1, E.clut_src.y1));
      ok = ok && (E.n_cmds - E.epoch_cmd_begin) + total < 65535u && E.n_cmds + total <= E.cmd_cap;
      if (ok)
      {
        uint32_t mylo = E.cur_lo, myhi = E.cur_hi, myserial = E.epoch_serial;
        const uint32_t n0 = E.n_cmds;
        const bool is_clut = in_batch && cls == CLS_CLUT;
        uint32_t clut_mask = __ballot_sync(FULL, is_clut);
        uint32_t prev = 0, set_from = 0;
        // Palettes repeat: usually every CLUT update of the batch finds its snapshot in the quick table.  Then nothing
        // is allocated and no epoch is cut — all that is left to decide is which slots are current for which record
        // ("the nearest CLUT update in front of me"), which is a ballot and a shuffle instead of a walk in order.
        uint32_t my_slot = 0xFFFFFFFFu;
        bool my_is8 = false;
        if (is_clut)
        {
          const uint32_t reg = pv[0][0] & 0xFFFFu;
          my_is8 = ((pv[0][0] >> 16) & 0xFFu) != 0;
          my_slot = E.quick_probe(PSX_PAL_X(reg), PSX_PAL_Y(reg), my_is8);
        }
        if (clut_mask != 0 && !__any_sync(FULL, is_clut && my_slot == 0xFFFFFFFFu))
        {
          if (is_clut)
            E.last[my_slot] = myserial;
          const uint32_t lt = (1u << lane) - 1u;
          const uint32_t mask8 = __ballot_sync(FULL, is_clut && my_is8);
          const uint32_t below = clut_mask & lt, below8 = mask8 & lt;
          const uint32_t lo_v = __shfl_sync(FULL, my_slot, below ? (31u - static_cast<uint32_t>(__clz(static_cast<int>(below)))) : 0u);
          const uint32_t hi_v = __shfl_sync(FULL, my_slot, below8 ? (31u - static_cast<uint32_t>(__clz(static_cast<int>(below8)))) : 0u);
          if (below)
            mylo = lo_v;
          if (below8)
            myhi = hi_v;
          E.cur_lo = __shfl_sync(FULL, my_slot, 31u - static_cast<uint32_t>(__clz(static_cast<int>(clut_mask))));
          if (mask8)
            E.cur_hi = __shfl_sync(FULL, my_slot, 31u - static_cast<uint32_t>(__clz(static_cast<int>(mask8))));
          E.st_hits += __popc(clut_mask);
          clut_mask = 0;
          __syncwarp();
        }
        while (clut_mask)
        {
          const uint32_t c = static_cast<uint32_t>(__ffs(static_cast<int>(clut_mask))) - 1u;
          clut_mask &= clut_mask - 1u;
          // the primitives in front of this CLUT update use the slots that are current now
          if (lane >= prev && lane < c && (t0 || t1))
          {
            E.last[mylo] = myserial;
            E.last[myhi] = myserial;
          }
          __syncwarp();
          E.n_cmds = n0 + __shfl_sync(FULL, rank, c);
          const uint32_t epochs_before = E.n_epochs;
          const uint32_t v = __shfl_sync(FULL, pv[0][0], c);
          E.update_clut(v & 0xFFFFu, ((v >> 16) & 0xFFu) != 0, false);
          if (E.n_epochs != epochs_before)
            set_from = c; // slot pressure cut the epoch: everything in front of c belongs to the closed one
          if (lane > c)
          {
            mylo = E.cur_lo;
            myhi = E.cur_hi;
            myserial = E.epoch_serial;
          }
          prev = c + 1;
        }
        if (lane >= prev && (t0 || t1))
        {
          E.last[mylo] = myserial;
          E.last[myhi] = myserial;
        }
        __syncwarp();
        const uint32_t w1 = hdm | (mylo << 16) | (myhi << 24);
        const uint32_t w3 = E.area[0] | (E.area[1] << 16), w4 = E.area[2] | (E.area[3] << 16);
        if (e0)
        {
          const uint32_t at = n0 + rank;
          uint4* d = reinterpret_cast<uint4*>(E.cmds + at);
          d[0] = make_uint4(hw0, w1, hwin, w3);
          d[1] = make_uint4(w4, pv[0][0], pv[0][1], pv[0][2]);
          d[2] = make_uint4(pv[0][3], pv[0][4], pv[0][5], pv[0][6]);
          d[3] = make_uint4(pv[0][7], 0, 0, 0);
          E.rects[at] = make_uint4(pwr0[0], pwr1[0], t0 ? prd0[0] : 0xFFFFFFFFu, prd1[0]);
        }
        if (e1)
        {
          const uint32_t at = n0 + rank + (e0 ? 1u : 0u);
          uint4* d = reinterpret_cast<uint4*>(E.cmds + at);
          d[0] = make_uint4(hw0, w1, hwin, w3);
          d[1] = make_uint4(w4, pv[1][0], pv[1][1], pv[1][2]);
          d[2] = make_uint4(pv[1][3], pv[1][4], pv[1][5], pv[1][6]);
          d[3] = make_uint4(pv[1][7], 0, 0, 0);
          E.rects[at] = make_uint4(pwr0[1], pwr1[1], t1 ? prd0[1] : 0xFFFFFFFFu, prd1[1]);
        }
        E.n_cmds = n0 + total;
        E.st_prims += total;
        E.st_rejected += __popc(__ballot_sync(FULL, fast && (pf[0] & FP_REJECTED))) + __popc(__ballot_sync(FULL, fast && (pf[1] & FP_REJECTED)));
        if (set_from != 0)
          box_of(set_from, wb, rb);
        if (wb[0] <= wb[2])
        {
          E.writes.any = 1;
          E.writes.x0 = min(E.writes.x0, wb[0]);
          E.writes.y0 = min(E.writes.y0, wb[1]);
          E.writes.x1 = max(E.writes.x1, wb[2]);
          E.writes.y1 = max(E.writes.y1, wb[3]);
        }
        if (rb[0] <= rb[2])
        {
          E.reads.any = 1;
          E.reads.x0 = min(E.reads.x0, rb[0]);
          E.reads.y0 = min(E.reads.y0, rb[1]);
          E.reads.x1 = max(E.reads.x1, rb[2]);
          E.reads.y1 = max(E.reads.y1, rb[3]);
        }
        if (area_mask)
        {
          const uint32_t r = nb - 1;
          E.area[0] = __shfl_sync(FULL, pv[0][0], r);
          E.area[1] = __shfl_sync(FULL, pv[0][1], r);
          E.area[2] = __shfl_sync(FULL, pv[0][2], r);
          E.area[3] = __shfl_sync(FULL, pv[0][3], r);
        }
        continue;
      }
    }

    for (uint32_t r = 0; r < nb; r++)
    {
      const uint32_t sm = __shfl_sync(FULL, summary, r);
      const uint32_t rcls = sm & 0xFu;
      if (rcls == CLS_FAST)
      {
        const uint32_t np = (sm >> 4) & 0xFu;
        for (uint32_t j = 0; j < np; j++)
        {
          const uint32_t f = (j ? (sm >> 16) : (sm >> 8)) & 0xFFu;
          if (f & FP_REJECTED)
            E.st_rejected++;
          if (!(f & FP_EMIT))
            continue;
          const uint32_t wr0 = __shfl_sync(FULL, j ? pwr0[1] : pwr0[0], r), wr1 = __shfl_sync(FULL, j ? pwr1[1] : pwr1[0], r);
          Rects wr, rd;
          wr.n = 1;
          wr.x0[0] = wr0 & 0xFFFFu;
          wr.y0[0] = wr0 >> 16;
          wr.x1[0] = wr1 & 0xFFFFu;
          wr.y1[0] = wr1 >> 16;
          rd.n = 0;
          const bool textured = (f & FP_TEXTURED) != 0;
          if (textured)
          {
            const uint32_t rd0 = __shfl_sync(FULL, j ? prd0[1] : prd0[0], r), rd1 = __shfl_sync(FULL, j ? prd1[1] : prd1[0], r);
            rd.n = 1;
            rd.x0[0] = rd0 & 0xFFFFu;
            rd.y0[0] = rd0 >> 16;
            rd.x1[0] = rd1 & 0xFFFFu;
            rd.y1[0] = rd1 >> 16;
          }
          uint32_t kind = EPOCH_TILED;
          if (f & FP_SELF)
          {
            kind = EPOCH_SELF_SAMPLING;
            E.st_self++;
          }
          uint4 side;
          const uint32_t idx = E.emit_slot(wr, rd, kind, textured, side);
          if (lane == r && idx != WarpEncoder::NO_SLOT)
          {
            E.rects[idx] = side;
            uint4* d = reinterpret_cast<uint4*>(E.cmds + idx);
            const uint32_t w1 = hdm | (E.cur_lo << 16) | (E.cur_hi << 24);
            const uint32_t w3 = E.area[0] | (E.area[1] << 16), w4 = E.area[2] | (E.area[3] << 16);
            if (j == 0)
            {
              d[0] = make_uint4(hw0, w1, hwin, w3);
              d[1] = make_uint4(w4, pv[0][0], pv[0][1], pv[0][2]);
              d[2] = make_uint4(pv[0][3], pv[0][4], pv[0][5], pv[0][6]);
              d[3] = make_uint4(pv[0][7], 0, 0, 0);
            }
            else
            {
              d[0] = make_uint4(hw0, w1, hwin, w3);
              d[1] = make_uint4(w4, pv[1][0], pv[1][1], pv[1][2]);
              d[2] = make_uint4(pv[1][3], pv[1][4], pv[1][5], pv[1][6]);
              d[3] = make_uint4(pv[1][7], 0, 0, 0);
            }
          }
        }
      }
      else if (rcls == CLS_CLUT)
      {
        const uint32_t v = __shfl_sync(FULL, pv[0][0], r);
        E.update_clut(v & 0xFFFFu, ((v >> 16) & 0xFFu) != 0);
      }
      else if (rcls == CLS_AREA)
      {
        E.area[0] = __shfl_sync(FULL, pv[0][0], r);
        E.area[1] = __shfl_sync(FULL, pv[0][1], r);
        E.area[2] = __shfl_sync(FULL, pv[0][2], r);
        E.area[3] = __shfl_sync(FULL, pv[0][3], r);
      }
      else if (rcls == CLS_SLOW)
      {
        generic_record(__shfl_sync(FULL, my_off, r));
      }
    }
  }
  if (E.n_cmds > E.epoch_cmd_begin || E.n_clut > E.epoch_clut_begin)
    E.close_epoch(EPOCH_TILED);

  // unused tail of the command slots: setup_kernel runs over all of them
  for (uint32_t i = E.n_cmds + lane; i < E.cmd_cap; i += 32)
  {
    uint4* d = reinterpret_cast<uint4*>(E.cmds + i);
    d[0] = make_uint4(OP_NOP, 0, 0, 0);
    d[1] = d[2] = d[3] = make_uint4(0, 0, 0, 0);
  }
  __syncwarp();
  // state that outlives the flush
  if (lane == 0)
  {
    for (int i = 0; i < 4; i++)
      st_out->area[i] = static_cast<uint16_t>(E.area[i]);
    st_out->cur_lo = static_cast<uint8_t>(E.cur_lo);
    st_out->cur_hi = static_cast<uint8_t>(E.cur_hi);
    st_out->alloc_hand = static_cast<uint16_t>(E.hand);
    st_out->pad = 0;
    EncResult res;
    res.n_cmds = min(E.n_cmds, E.cmd_cap);
    res.n_clut = min(E.n_clut, E.clut_cap);
    res.n_epochs = min(E.n_epochs, E.epoch_cap);
    res.pad = 0;
    A.result[idx] = res;
    atomicMax(&A.summary->max_epochs, res.n_epochs);
    atomicAdd(&A.summary->total_items, res.n_epochs);
    if (E.overflow)
      atomicOr(&A.summary->overflow, 1u);
    unsigned long long* S = A.summary->stats;
    atomicAdd(&S[ENC_STAT_WIRE], static_cast<unsigned long long>(n_seen));
    atomicAdd(&S[ENC_STAT_PRIMS], static_cast<unsigned long long>(E.st_prims));
    atomicAdd(&S[ENC_STAT_EPOCHS], static_cast<unsigned long long>(E.n_epochs));
    atomicAdd(&S[ENC_STAT_CUTS_RAW], static_cast<unsigned long long>(E.st_raw));
    atomicAdd(&S[ENC_STAT_CUTS_WAR], static_cast<unsigned long long>(E.st_war));
    atomicAdd(&S[ENC_STAT_CUTS_CLUT], static_cast<unsigned long long>(E.st_clutcut));
    atomicAdd(&S[ENC_STAT_SELF], static_cast<unsigned long long>(E.st_self));
    atomicAdd(&S[ENC_STAT_COPY], static_cast<unsigned long long>(E.st_copy));
    atomicAdd(&S[ENC_STAT_UPLOAD], static_cast<unsigned long long>(E.st_upload));
    atomicAdd(&S[ENC_STAT_CLUT_OPS], static_cast<unsigned long long>(E.st_clutops));
    atomicAdd(&S[ENC_STAT_CLUT_HITS], static_cast<unsigned long long>(E.st_hits));
    atomicAdd(&S[ENC_STAT_REJECTED], static_cast<unsigned long long>(E.st_rejected));
  }
  for (uint32_t s = lane; s < 256; s += 32)
    st_out->slot_key[s] = E.key[s];
  // wave k = epoch k of every instance: count the items and what they contain
  const uint32_t ne = min(E.n_epochs, E.epoch_cap);
  for (uint32_t k = lane; k < ne && k < A.wave_cap; k += 32)
  {
    const Epoch e = E.epochs[k];
    uint32_t f = (e.clut_end > e.clut_begin) ? 1u : 0u;
    if (e.cmd_end > e.cmd_begin)
      f |= (e.kind == EPOCH_TILED) ? 2u : 4u;
    atomicAdd(&A.waves[k].item_count, 1u);
    atomicOr(&A.waves[k].flags, f);
    if (e.kind == EPOCH_TILED)
      atomicAdd(&A.waves[k].tiled_cmds, e.cmd_end - e.cmd_begin);
  }
}

// Record framing of one stream per warp: sizes add up, no barrier records; counts what the encoder can produce and
// notes where every record starts.  Only following the sizes is sequential (a short loop on shared-memory latency, 32
// records at a time); classifying the 32 records, counting and storing their offsets is done one record per lane.
__global__ void __launch_bounds__(32 * WALK_WARPS) walk_kernel(const uint8_t* __restrict__ wire, const EncInst* __restrict__ in, WalkOut* __restrict__ out,
                                                   uint32_t* __restrict__ rec_off, uint32_t n_inst)
{
  __shared__ __align__(16) uint32_t sh_win[WALK_WARPS][WIN_BYTES / 4];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t idx = blockIdx.x * WALK_WARPS + warp;
  if (idx >= n_inst)
    return;
  const EncInst wi = in[idx];
  const uint32_t* w32 = reinterpret_cast<const uint32_t*>(wire);
  WalkOut r;
  r.n_rec = r.cmd_cap = r.clut_cap = 0;
  r.status = 0;
  uint32_t pos = wi.wire_off;
  const uint32_t end = wi.wire_off + wi.wire_bytes;
  StreamWindow win;
  win.wire = wire;
  win.smem = sh_win[warp];
  win.limit = end + 128u;
  win.lane = lane;
  win.base = 0xFFFFFF00u;
  while (pos < end)
  {
    // -- follow up to 32 sizes --
    uint32_t nb = 0, my_off = 0, my_size = 0, scan = pos;
    int32_t chase_status = 0;
    for (uint32_t i = 0; i < 32; i++)
    {
      if (scan >= end)
        break;
      if (end - scan < 8u)
      {
        chase_status = -3;
        break;
      }
      const uint32_t sz = win.word(scan);
      if (sz < 8u || (sz & 3u) != 0 || sz > end - scan)
      {
        chase_status = -3;
        break;
      }
      if (lane == i)
      {
        my_off = scan;
        my_size = sz;
      }
      nb = i + 1;
      scan += sz;
    }
    // -- one record per lane --
    const bool have = lane < nb;
    uint32_t type = 0, w2 = 0;
    if (have)
    {
      if (win.holds(my_off, 12u))
      {
        type = win.smem[(my_off + 4u - win.base) >> 2] & 0xFFu;
        w2 = win.smem[(my_off + 8u - win.base) >> 2];
      }
      else
      {
        type = __ldg(w32 + (my_off >> 2) + 1) & 0xFFu;
        w2 = (my_size >= 12u) ? __ldg(w32 + (my_off >> 2) + 2) : 0u;
      }
    }
    const bool barrier =
      have && (type == PSX_CMD_CLEAR_VRAM || type == PSX_CMD_LOAD_STATE || type == PSX_CMD_READ_VRAM || type == PSX_CMD_UPDATE_DISPLAY);
    const uint32_t barrier_mask = __ballot_sync(FULL, barrier);
    const uint32_t counted = barrier_mask ? (static_cast<uint32_t>(__ffs(static_cast<int>(barrier_mask))) - 1u) : nb; // stop in front of a barrier
    uint32_t cmd_inc = 0, clut_inc = 0;
    if (lane < counted)
    {
      switch (type)
      {
        case PSX_CMD_DRAW_POLYGON:
        case PSX_CMD_DRAW_PRECISE_POLYGON:
          if (my_size >= 20u)
            cmd_inc = ((w2 >> 16) > 3u) ? 2u : 1u;
          break;
        case PSX_CMD_DRAW_LINE:
        case PSX_CMD_DRAW_PRECISE_LINE:
          if (my_size >= 20u)
            cmd_inc = (w2 >> 16) / 2u;
          break;
        case PSX_CMD_DRAW_RECTANGLE:
        case PSX_CMD_FILL_VRAM:
        case PSX_CMD_COPY_VRAM:
        case PSX_CMD_UPDATE_VRAM: cmd_inc = 1; break;
        case PSX_CMD_UPDATE_CLUT: clut_inc = 1; break;
        default: break;
      }
      rec_off[wi.rec_base + r.n_rec + lane] = my_off; // the encoder reads these 32 at a time
    }
    r.cmd_cap += __reduce_add_sync(FULL, cmd_inc);
    r.clut_cap += __reduce_add_sync(FULL, clut_inc);
    r.n_rec += counted;
    if (barrier_mask)
    {
      r.status = -5;
      break;
    }
    if (chase_status)
    {
      r.status = chase_status;
      break;
    }
    pos = scan;
  }
  if (lane == 0)
    out[idx] = r;
}

// Block-wide exclusive scan of one value per thread (1024 threads); returns the block total through `total`.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total)
{
  uint32_t incl = v;
#pragma unroll
  for (uint32_t d = 1; d < 32; d <<= 1)
  {
    const uint32_t t = __shfl_up_sync(FULL, incl, d);
    if ((threadIdx.x & 31u) >= d)
      incl += t;
  }
  __syncthreads();
  if ((threadIdx.x & 31u) == 31u)
    warp_sums[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32)
  {
    uint32_t s = warp_sums[threadIdx.x];
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1)
    {
      const uint32_t t = __shfl_up_sync(FULL, s, d);
      if (threadIdx.x >= d)
        s += t;
    }
    warp_sums[threadIdx.x] = s;
  }
  __syncthreads();
  total = warp_sums[31];
  return ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0u) + (incl - v);
}

// Output layout of a chunk from the framing counts: instance k gets cmd / CLUT-op / epoch slots behind instance k-1's.
__global__ void __launch_bounds__(1024) layout_kernel(const WalkOut* __restrict__ walk, EncInst* __restrict__ inst, uint32_t n_inst)
{
  __shared__ uint32_t sums[32];
  uint32_t cmd_carry = 0, clut_carry = 0;
  for (uint32_t base = 0; base < n_inst; base += 1024)
  {
    const uint32_t k = base + threadIdx.x;
    const uint32_t c = (k < n_inst) ? walk[k].cmd_cap : 0u, u = (k < n_inst) ? walk[k].clut_cap : 0u;
    uint32_t ctot, utot;
    const uint32_t cpre = block_exclusive_scan(c, sums, ctot);
    const uint32_t upre = block_exclusive_scan(u, sums, utot);
    if (k < n_inst)
    {
      inst[k].cmd_base = cmd_carry + cpre;
      inst[k].cmd_cap = c;
      inst[k].clut_base = clut_carry + upre;
      inst[k].clut_cap = u;
      inst[k].epoch_base = cmd_carry + cpre + clut_carry + upre + k; // capacity c + u + 1 each
      inst[k].epoch_cap = c + u + 1;
    }
    cmd_carry += ctot;
    clut_carry += utot;
  }
}

// Exclusive scan of the per-wave item counts (one block; the number of waves is data dependent).
__global__ void __launch_bounds__(1024) wave_scan_kernel(EncWave* waves, const EncSummary* summary, uint32_t wave_cap)
{
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  const uint32_t M = min(summary->max_epochs, wave_cap);
  if (threadIdx.x == 0)
    carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < M; base += 1024)
  {
    const uint32_t k = base + threadIdx.x;
    const uint32_t v = (k < M) ? waves[k].item_count : 0u;
    uint32_t incl = v;
#pragma unroll
    for (uint32_t d = 1; d < 32; d <<= 1)
    {
      const uint32_t t = __shfl_up_sync(FULL, incl, d);
      if ((threadIdx.x & 31u) >= d)
        incl += t;
    }
    if ((threadIdx.x & 31u) == 31u)
      warp_sums[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32)
    {
      uint32_t s = warp_sums[threadIdx.x];
#pragma unroll
      for (uint32_t d = 1; d < 32; d <<= 1)
      {
        const uint32_t t = __shfl_up_sync(FULL, s, d);
        if (threadIdx.x >= d)
          s += t;
      }
      warp_sums[threadIdx.x] = s;
    }
    __syncthreads();
    const uint32_t before = carry + ((threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0u) + (incl - v);
    if (k < M)
      waves[k].item_off = before;
    __syncthreads();
    if (threadIdx.x == 1023)
      carry = before + v;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128) wave_fill_kernel(EncodeArgs A)
{
  const uint32_t idx = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
  if (idx >= A.n_inst)
    return;
  const EncInst ii = A.inst[idx];
  const uint32_t ne = A.result[idx].n_epochs;
  if (lane == 0)
  {
    InstWork w;
    w.instance = ii.instance;
    w.cmd_base = ii.cmd_base;
    w.clut_base = ii.clut_base;
    w.epoch_first = ii.epoch_base;
    w.n_epochs = min(ne, A.wave_cap);
    w.pad[0] = w.pad[1] = w.pad[2] = 0;
    A.inst_work[idx] = w;
  }
  for (uint32_t k = lane; k < ne && k < A.wave_cap; k += 32)
  {
    const Epoch e = A.epochs[ii.epoch_base + k];
    const uint32_t pos = A.waves[k].item_off + atomicAdd(&A.wave_cursor[k], 1u);
    WaveItem it;
    it.instance = ii.instance;
    it.cmd_begin = ii.cmd_base + e.cmd_begin;
    it.cmd_end = ii.cmd_base + e.cmd_end;
    it.kind = e.kind;
    it.clut_begin = ii.clut_base + e.clut_begin;
    it.clut_end = ii.clut_base + e.clut_end;
    it.pad[0] = it.pad[1] = 0;
    A.items[pos] = it;
  }
}

} // namespace

cudaError_t launch_walk(const uint8_t* wire, const EncInst* inst, WalkOut* out, uint32_t* rec_off, uint32_t n_inst, cudaStream_t st)
{
  if (n_inst == 0)
    return cudaSuccess;
  walk_kernel<<<(n_inst + WALK_WARPS - 1) / WALK_WARPS, 32 * WALK_WARPS, 0, st>>>(wire, inst, out, rec_off, n_inst);
  return cudaGetLastError();
}

cudaError_t launch_layout(const WalkOut* walk, EncInst* inst, uint32_t n_inst, cudaStream_t st)
{
  if (n_inst == 0)
    return cudaSuccess;
  layout_kernel<<<1, 1024, 0, st>>>(walk, inst, n_inst);
  return cudaGetLastError();
}

cudaError_t launch_encode(const EncodeArgs& a, cudaStream_t st)
{
  if (a.n_inst == 0)
    return cudaSuccess;
  cudaError_t e;
  if ((e = cudaMemsetAsync(a.waves, 0, static_cast<size_t>(a.wave_cap) * sizeof(EncWave), st)) != cudaSuccess)
    return e;
  if ((e = cudaMemsetAsync(a.wave_cursor, 0, static_cast<size_t>(a.wave_cap) * sizeof(uint32_t), st)) != cudaSuccess)
    return e;
  if ((e = cudaMemsetAsync(a.summary, 0, sizeof(EncSummary), st)) != cudaSuccess)
    return e;
  const uint32_t blocks = (a.n_inst + ENC_WARPS - 1) / ENC_WARPS;
  encode_kernel<<<blocks, 32 * ENC_WARPS, 0, st>>>(a);
  wave_scan_kernel<<<1, 1024, 0, st>>>(a.waves, a.summary, a.wave_cap);
  wave_fill_kernel<<<(a.n_inst + 3) / 4, 128, 0, st>>>(a);
  return cudaGetLastError();
}

} // namespace psx
