// context.cpp — the C ABI (include/psxgpu.h): device memory, the batched-instance scheduler and flush logic.
//
// One context = one CUDA device, one stream, N independent 1024x512 VRAM instances resident in HBM
// (1 MiB each) plus a ring of CLUT snapshot slots per instance.  Commands are encoded per instance on the
// host (encoder.cpp), packed into pinned staging buffers, copied once per flush and executed as "waves":
// wave k runs epoch k of every instance that has one, as a single grid of (instance, tile) CTAs.
#include "../../include/psxgpu.h"

#include "device_encoder.h"
#include "encoder.h"
#include "kernels.h"

#include <algorithm>
#include <new>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

using namespace psx;

namespace {

struct PinnedBuf
{
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap)
      return cudaSuccess;
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e == cudaSuccess)
      cap = want;
    return e;
  }
  void release()
  {
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

struct DevBuf
{
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap)
      return cudaSuccess;
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess)
      cap = want;
    return e;
  }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

template<typename F>
void parallel_for(uint32_t n, uint32_t threads, F&& fn)
{
  if (n == 0)
    return;
  threads = std::max(1u, std::min(threads, n));
  if (threads == 1)
  {
    for (uint32_t i = 0; i < n; i++)
      fn(i);
    return;
  }
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (uint32_t t = 0; t < threads; t++)
  {
    pool.emplace_back([=, &fn]() {
      const uint32_t lo = static_cast<uint32_t>((static_cast<uint64_t>(n) * t) / threads);
      const uint32_t hi = static_cast<uint32_t>((static_cast<uint64_t>(n) * (t + 1)) / threads);
      for (uint32_t i = lo; i < hi; i++)
        fn(i);
    });
  }
  for (auto& th : pool)
    th.join();
}

// Like parallel_for, but hands every worker its contiguous range once: fn(worker, lo, hi).
template<typename F>
uint32_t parallel_ranges(uint32_t n, uint32_t threads, F&& fn)
{
  threads = std::max(1u, std::min(threads, std::max(1u, n)));
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (uint32_t t = 0; t < threads; t++)
  {
    const uint32_t lo = static_cast<uint32_t>((static_cast<uint64_t>(n) * t) / threads);
    const uint32_t hi = static_cast<uint32_t>((static_cast<uint64_t>(n) * (t + 1)) / threads);
    if (threads == 1)
      fn(t, lo, hi);
    else
      pool.emplace_back([=, &fn]() { fn(t, lo, hi); });
  }
  for (auto& th : pool)
    th.join();
  return threads;
}

double now_ms()
{
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

struct psxgpu_ctx
{
  int device = 0;
  uint32_t n = 0;
  psxgpu_opts opts{};
  uint32_t clut_slots = 256;
  uint32_t threads = 1;
  cudaStream_t stream = nullptr;
  // batched path with device-side encoding: raw streams are uploaded on copy_stream, encoded on enc_stream, and the
  // main stream picks the chunk up once its wave table is known
  cudaStream_t copy_stream = nullptr, enc_streams[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t walk_stream = nullptr; // framing kernels: tiny, high priority (their result gates the host)
  // PSXGPU_TRACE=1: per-chunk device timeline of the batched path (copy, encode, waves), printed by psxgpu_sync
  bool trace = false;
  struct TraceChunk
  {
    cudaEvent_t copy0, copy1, enc1, run0, run1;
  };
  std::vector<TraceChunk> trace_chunks;
  cudaEvent_t trace_origin = nullptr;
  uint32_t encode_mode = 0; // opts.reserved[1]: 0 auto (device encoder for batches of >= 16 instances), 1 host, 2 device

  uint16_t* d_vram = nullptr;
  uint16_t* d_clut = nullptr;
  unsigned long long* d_hash = nullptr;
  unsigned int* d_barrier = nullptr;

  std::vector<Encoder> enc;
  std::vector<uint8_t> dirty;
  // Encoder state of instances last encoded on the device (blob_truth[i] != 0): imported into enc[i] on demand.
  std::vector<EncState> enc_blob;
  std::vector<uint8_t> blob_truth;

  PinnedBuf h_misc;
  DevBuf d_scan;

  // One chunk = one flush unit: device-resident command buffers + launch descriptors.  A "group" is the list of chunks
  // launched since the previous psxgpu_flush (a batched submit is cut into several so that host encoding of chunk k+1
  // overlaps the copy + kernels of chunk k); psxgpu_run_staged replays the last closed group.
  struct Chunk
  {
    DevBuf d_cmds, d_setups, d_bins, d_tilecmds, d_payload, d_clutops, d_items, d_misc;
    // device-encoded chunks: d_payload holds the raw wire streams (upload pixels are read in place)
    DevBuf d_recoff, d_rec, d_inst, d_state, d_state_out, d_epochs, d_result, d_waves, d_cursor, d_summary, d_rects;
    DevBuf d_instwork; // per instance: where its epochs / commands / CLUT ops start (instance mode; d_epochs holds the epochs)
    bool instance_mode = false; // many short epochs per instance: one CTA per instance for the whole flush (raster_instance_kernel)
    bool instance_wide = false; // ... with 32 warps per CTA (a handful of instances)
    EncodeArgs enc_args{};
    bool device_encoded = false; // psxgpu_run_staged replays the encode kernels too (from the resident wire streams)
    cudaEvent_t ev_copied = nullptr, ev_walked = nullptr, ev_encoded = nullptr;
    cudaEvent_t ev_done = nullptr; // main stream: last launch that reads this chunk's buffers
    bool done_valid = false;
    uint32_t n_cmds = 0;
    bool persistent = false; // few instances: one cooperative launch for all waves (raster_persistent_kernel)
    uint32_t n_instances = 0;
    std::vector<uint32_t> wave_off{0}; // wave k = items [wave_off[k], wave_off[k+1])
    std::vector<uint8_t> wave_has_clut, wave_has_tiled, wave_has_serial;
    std::vector<uint32_t> wave_tiled_cmds; // commands of each wave's tiled epochs (raster loop grab size)
    std::vector<cudaEvent_t> ev_raster; // pairs, profile mode
    uint32_t ev_raster_used = 0;
  };
  std::vector<Chunk*> chunks;
  uint32_t group_n = 0;     // chunks of the open group
  bool group_open = false;
  uint32_t staged_n = 0;    // chunks of the last closed group
  uint32_t batch_chunk = 0; // instances per chunk in psxgpu_submit_batch (0 = auto)
  uint32_t persist_max = 0; // largest instance count the persistent kernel can hold co-resident (0 = unavailable)

  // pinned staging, triple buffered: a set may be refilled once the copies that read it have completed
  struct PinnedSet
  {
    PinnedBuf cmds, payload, clutops, items, misc;
    PinnedBuf wire, recoff, inst, state, result; // device-encoded chunks (result = D2H: summary, wave table, states)
    cudaEvent_t free_ev = nullptr;
    bool in_flight = false;
  };
  // sets 0-2: host-encoded flushes (triple buffered); sets 3-8: device-encoded chunks, up to six in flight
  PinnedSet pinned[9];
  uint32_t pin_next = 0, pin_next_dev = 0;

  std::unordered_map<uint32_t, std::vector<uint16_t>> shadow;
  std::vector<uint8_t> display;
  uint32_t disp_w = 0, disp_h = 0, disp_stride = 0, disp_rgba8 = 1;

  cudaEvent_t ev_start = nullptr, ev_end = nullptr, ev_hash0 = nullptr, ev_hash1 = nullptr;
  bool timing_pending = false;
  uint32_t timing_chunks = 0;
  double stage_ms_acc = 0, scan_ms_acc = 0, wait_ms_acc = 0, launch_ms_acc = 0;

  psxgpu_stats counters{};
  psxgpu_timing timing{};
  std::vector<float> raster_ms;
  double encode_ms_acc = 0;
  std::string error;

  int fail(cudaError_t e, const char* what)
  {
    char buf[256];
    std::snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    error = buf;
    return PSXGPU_E_CUDA;
  }
  int invalid(const char* what)
  {
    error = what;
    return PSXGPU_E_INVALID;
  }
};

#define CK(expr, what)                                                                                                 \
  do                                                                                                                   \
  {                                                                                                                    \
    cudaError_t e__ = (expr);                                                                                          \
    if (e__ != cudaSuccess)                                                                                            \
      return ctx->fail(e__, what);                                                                                     \
  } while (0)

namespace {

uint16_t* inst_vram(psxgpu_ctx* ctx, uint32_t i)
{
  return ctx->d_vram + static_cast<size_t>(i) * VRAM_PIXELS;
}
uint16_t* inst_clut(psxgpu_ctx* ctx, uint32_t i, uint32_t slot)
{
  return ctx->d_clut + (static_cast<size_t>(i) * ctx->clut_slots + slot) * CLUT_ENTRIES;
}

int collect_timing(psxgpu_ctx* ctx)
{
  if (!ctx->timing_pending)
    return PSXGPU_OK;
  CK(cudaEventSynchronize(ctx->ev_end), "event sync");
  float a = 0;
  cudaEventElapsedTime(&a, ctx->ev_start, ctx->ev_end);
  ctx->timing.last_flush_device_ms = a;
  ctx->timing.last_kernels_ms = a;
  ctx->raster_ms.clear();
  float sum = 0, mx = 0;
  uint32_t waves = 0;
  for (uint32_t c = 0; c < ctx->timing_chunks && c < ctx->chunks.size(); c++)
  {
    psxgpu_ctx::Chunk& ch = *ctx->chunks[c];
    waves += static_cast<uint32_t>(ch.wave_off.size()) - 1;
    for (uint32_t i = 0; i + 1 < ch.ev_raster_used; i += 2)
    {
      float t = 0;
      cudaEventElapsedTime(&t, ch.ev_raster[i], ch.ev_raster[i + 1]);
      ctx->raster_ms.push_back(t);
      sum += t;
      mx = std::max(mx, t);
    }
  }
  ctx->timing.last_raster_ms = sum;
  ctx->timing.last_raster_max_ms = mx;
  ctx->timing.last_raster_launches = static_cast<uint32_t>(ctx->raster_ms.size());
  ctx->timing.last_waves = waves;
  ctx->timing_pending = false;
  return PSXGPU_OK;
}

// Instance mode pays off when a chunk's instances are cut into many short epochs: per epoch the wave modes pay a
// kernel launch (or a grid barrier) and one pass over every touched tile.  PSXGPU_INSTANCE_MODE = 0 / 1 forces it.
bool instance_mode_wanted(psxgpu_ctx* ctx, uint64_t total_epochs, uint32_t n_instances, uint64_t n_cmds)
{
  static const int forced = []() {
    const char* e = std::getenv("PSXGPU_INSTANCE_MODE");
    return e ? std::atoi(e) : -1;
  }();
  if (ctx->opts.reserved[0] & 2u) // instance mode disabled (tests: one launch per epoch wanted)
    return false;
  if (forced >= 0)
    return forced != 0 && total_epochs > 0;
  if (n_instances == 0 || total_epochs < 8ull * n_instances)
    return false;
  return n_cmds < 64ull * total_epochs; // fewer than 64 commands per epoch on average
}

// setup + all waves of one chunk, on ctx->stream
int run_chunk(psxgpu_ctx* ctx, psxgpu_ctx::Chunk& S)
{
#ifdef PSX_DEBUG_BOUNDS
  {
    // what the chunk's kernels may index (kernels.cu, PSX_CHECK): the allocations' extents, like a memory checker would
    DebugLimits lim;
    lim.n_cmds = S.n_cmds;
    lim.n_payload = static_cast<uint32_t>(std::min<size_t>(S.d_payload.cap / 2, 0xFFFFFFFFu));
    lim.n_clutops = static_cast<uint32_t>(S.d_clutops.cap / sizeof(ClutOp));
    lim.n_items = S.wave_off.empty() ? 0u : S.wave_off.back();
    lim.n_instances = ctx->n;
    lim.clut_slots = ctx->clut_slots;
    CK(launch_debug_limits(lim, ctx->stream), "debug limits");
  }
#endif
  CK(launch_setup(static_cast<const PackedCmd*>(S.d_cmds.p), static_cast<PrimSetup*>(S.d_setups.p),
                  static_cast<BinInfo*>(S.d_bins.p), static_cast<uint32_t*>(S.d_tilecmds.p), S.n_cmds, ctx->stream),
     "setup_kernel");
  if (S.n_cmds)
    ctx->counters.kernel_launches++;
  const uint32_t waves = static_cast<uint32_t>(S.wave_off.size()) - 1;
  if (!S.instance_mode && !S.wave_off.empty() && S.wave_off.back() != 0)
  {
    CK(launch_item_union(static_cast<WaveItem*>(S.d_items.p), S.wave_off.back(), static_cast<const BinInfo*>(S.d_bins.p), ctx->stream),
       "item_union_kernel");
    ctx->counters.kernel_launches++;
  }
  S.ev_raster_used = 0;
  if (S.instance_mode)
  {
    const bool prof_i = ctx->opts.profile != 0;
    while (prof_i && S.ev_raster.size() < 2)
    {
      cudaEvent_t e;
      CK(cudaEventCreate(&e), "event create");
      S.ev_raster.push_back(e);
    }
    if (prof_i)
      CK(cudaEventRecord(S.ev_raster[S.ev_raster_used++], ctx->stream), "event record");
    CK(launch_instances(static_cast<const InstWork*>(S.d_instwork.p), S.n_instances, static_cast<const Epoch*>(S.d_epochs.p),
                        static_cast<uint32_t*>(S.d_items.p), // (the wave items are not used in this mode: their buffer, 32 bytes per
                                                             // epoch, holds the epochs' tile sets)
                        static_cast<const PrimSetup*>(S.d_setups.p), static_cast<const BinInfo*>(S.d_bins.p),
                        static_cast<const uint32_t*>(S.d_tilecmds.p), static_cast<const ClutOp*>(S.d_clutops.p), ctx->d_vram, ctx->d_clut,
                        static_cast<const uint16_t*>(S.d_payload.p), ctx->clut_slots, S.instance_wide, ctx->stream, ctx->d_barrier + 8),
       "raster_instance_kernel");
    if (prof_i)
      CK(cudaEventRecord(S.ev_raster[S.ev_raster_used++], ctx->stream), "event record");
    ctx->counters.kernel_launches += 2;
    ctx->counters.waves += waves;
    if (S.ev_done)
    {
      CK(cudaEventRecord(S.ev_done, ctx->stream), "event record");
      S.done_valid = true;
    }
    return PSXGPU_OK;
  }
  if (S.persistent)
  {
    const uint8_t* misc = static_cast<const uint8_t*>(S.d_misc.p);
    CK(launch_persistent(static_cast<const WaveItem*>(S.d_items.p), misc, waves,
                         reinterpret_cast<const uint32_t*>(misc + static_cast<size_t>(waves) * 16), S.n_instances,
                         static_cast<const PrimSetup*>(S.d_setups.p), static_cast<const uint32_t*>(S.d_tilecmds.p),
                         static_cast<const ClutOp*>(S.d_clutops.p), ctx->d_vram, ctx->d_clut,
                         static_cast<const uint16_t*>(S.d_payload.p), ctx->clut_slots, ctx->d_barrier, ctx->stream),
       "raster_persistent_kernel");
    ctx->counters.kernel_launches++;
    ctx->counters.waves += waves;
    return PSXGPU_OK;
  }
  const bool prof = ctx->opts.profile != 0;
  if (prof)
  {
    while (S.ev_raster.size() < static_cast<size_t>(waves) * 2)
    {
      cudaEvent_t e;
      CK(cudaEventCreate(&e), "event create");
      S.ev_raster.push_back(e);
    }
  }
  const WaveItem* items = static_cast<const WaveItem*>(S.d_items.p);
  for (uint32_t k = 0; k < waves; k++)
  {
    const uint32_t off = S.wave_off[k], cnt = S.wave_off[k + 1] - off;
    if (cnt == 0)
      continue;
    if (S.wave_has_clut[k])
    {
      CK(launch_clut(items + off, cnt, static_cast<const ClutOp*>(S.d_clutops.p), ctx->d_vram, ctx->d_clut, ctx->clut_slots,
                     ctx->stream),
         "clut_kernel");
      ctx->counters.kernel_launches++;
    }
    if (prof)
      CK(cudaEventRecord(S.ev_raster[S.ev_raster_used++], ctx->stream), "event record");
    if (S.wave_has_tiled[k])
    {
      // A wave of a pipelined (device-encoded) batch goes out in PSXGPU_WAVE_SPLIT pieces: the persistent raster CTAs of
      // a piece all end together, and at that seam the next chunk's framing / encode kernels, which wait on a
      // higher-priority stream, get onto the SMs instead of waiting for the whole wave.
      static const uint32_t wave_split = []() {
        const char* e = std::getenv("PSXGPU_WAVE_SPLIT");
        const int v = e ? std::atoi(e) : 1;
        return static_cast<uint32_t>(v < 1 ? 1 : (v > 8 ? 8 : v));
      }();
      const uint32_t pieces = (S.device_encoded && cnt >= 128u * wave_split) ? wave_split : 1u;
      for (uint32_t pc = 0; pc < pieces; pc++)
      {
        const uint32_t lo = static_cast<uint32_t>(static_cast<uint64_t>(cnt) * pc / pieces);
        const uint32_t hi = static_cast<uint32_t>(static_cast<uint64_t>(cnt) * (pc + 1) / pieces);
        CK(launch_raster(items + off + lo, hi - lo, static_cast<const PrimSetup*>(S.d_setups.p), static_cast<const uint32_t*>(S.d_tilecmds.p),
                         ctx->d_vram, ctx->d_clut, static_cast<const uint16_t*>(S.d_payload.p), ctx->clut_slots, ctx->stream,
                         static_cast<uint32_t>(static_cast<uint64_t>(S.wave_tiled_cmds[k]) * (hi - lo) / cnt), ctx->d_barrier + 4),
           "raster_kernel");
        ctx->counters.kernel_launches += (hi - lo + 32767) / 32768;
      }
    }
    if (S.wave_has_serial[k])
    {
      CK(launch_serial(items + off, cnt, static_cast<const PrimSetup*>(S.d_setups.p), ctx->d_vram, ctx->d_clut,
                       static_cast<const uint16_t*>(S.d_payload.p), ctx->clut_slots, ctx->stream),
         "serial_kernel");
      ctx->counters.kernel_launches++;
    }
    if (prof)
      CK(cudaEventRecord(S.ev_raster[S.ev_raster_used++], ctx->stream), "event record");
  }
  ctx->counters.waves += waves;
  if (S.ev_done)
  {
    CK(cudaEventRecord(S.ev_done, ctx->stream), "event record");
    S.done_valid = true;
  }
  return PSXGPU_OK;
}

int open_group(psxgpu_ctx* ctx)
{
  if (ctx->group_open)
    return PSXGPU_OK;
  int rc = collect_timing(ctx); // the events are about to be reused
  if (rc != PSXGPU_OK)
    return rc;
  ctx->group_open = true;
  ctx->group_n = 0;
  ctx->stage_ms_acc = 0;
  CK(cudaEventRecord(ctx->ev_start, ctx->stream), "event record");
  return PSXGPU_OK;
}

int close_group(psxgpu_ctx* ctx)
{
  if (!ctx->group_open)
    return PSXGPU_OK;
  CK(cudaEventRecord(ctx->ev_end, ctx->stream), "event record");
  ctx->group_open = false;
  ctx->staged_n = ctx->group_n;
  ctx->timing_chunks = ctx->group_n;
  ctx->timing_pending = true;
  ctx->timing.last_stage_ms = static_cast<float>(ctx->stage_ms_acc);
  ctx->timing.last_encode_ms = static_cast<float>(ctx->encode_ms_acc);
  ctx->timing.last_scan_ms = static_cast<float>(ctx->scan_ms_acc);
  ctx->timing.last_pinned_wait_ms = static_cast<float>(ctx->wait_ms_acc);
  ctx->timing.last_launch_ms = static_cast<float>(ctx->launch_ms_acc);
  ctx->encode_ms_acc = ctx->scan_ms_acc = ctx->wait_ms_acc = ctx->launch_ms_acc = 0;
  ctx->counters.flushes++;
  return PSXGPU_OK;
}

int acquire_pinned(psxgpu_ctx* ctx, psxgpu_ctx::PinnedSet*& out, bool device_path = false)
{
  psxgpu_ctx::PinnedSet& P = device_path ? ctx->pinned[3 + ctx->pin_next_dev] : ctx->pinned[ctx->pin_next];
  if (device_path)
    ctx->pin_next_dev = (ctx->pin_next_dev + 1) % 6;
  else
    ctx->pin_next = (ctx->pin_next + 1) % 3;
  if (!P.free_ev)
    CK(cudaEventCreate(&P.free_ev), "event create");
  if (P.in_flight)
  {
    const double t0 = now_ms();
    CK(cudaEventSynchronize(P.free_ev), "pinned buffer sync");
    ctx->wait_ms_acc += now_ms() - t0;
    P.in_flight = false;
  }
  out = &P;
  return PSXGPU_OK;
}

// Common tail of a flush unit.  The packed records and upload pixels of work[w] already sit in the pinned set at
// cmd_base[w] / pay_base[w] (cmd_base[nw] / pay_base[nw] = extents); this adds CLUT ops and wave descriptors, copies
// everything to a chunk's device buffers and launches setup + waves.  Asynchronous.
// A run of records / upload pixels that is contiguous both in the pinned set and on the device.
struct CopySeg
{
  uint64_t host_off, dev_off, count;
};
// Where the pinned copy differs from the device layout (per-thread arenas): host position of each instance's first
// record and the copy segments.  Null = the pinned set is already laid out like the device buffers.
struct HostLayout
{
  std::vector<uint64_t> host_cmd_base;
  std::vector<CopySeg> cmd_segs, pay_segs;
};

int launch_chunk_impl(psxgpu_ctx* ctx, const std::vector<uint32_t>& work, const std::vector<uint64_t>& cmd_base,
                      const std::vector<uint64_t>& pay_base, psxgpu_ctx::PinnedSet& P, const HostLayout* layout);
int launch_chunk(psxgpu_ctx* ctx, const std::vector<uint32_t>& work, const std::vector<uint64_t>& cmd_base,
                 const std::vector<uint64_t>& pay_base, psxgpu_ctx::PinnedSet& P, const HostLayout* layout = nullptr)
{
  const double t0 = now_ms();
  const int rc = launch_chunk_impl(ctx, work, cmd_base, pay_base, P, layout);
  ctx->launch_ms_acc += now_ms() - t0;
  return rc;
}
int launch_chunk_impl(psxgpu_ctx* ctx, const std::vector<uint32_t>& work, const std::vector<uint64_t>& cmd_base,
                      const std::vector<uint64_t>& pay_base, psxgpu_ctx::PinnedSet& P, const HostLayout* layout)
{
  const uint32_t nw = static_cast<uint32_t>(work.size());
  const uint64_t n_cmds = cmd_base[nw], n_pay = pay_base[nw];
  if (n_cmds > 0xFFFFFFF0ull || n_pay > 0xFFFFFFF0ull)
    return ctx->invalid("flush too large: more than 2^32 commands or payload pixels");
  std::vector<uint64_t> clut_base(nw + 1, 0);
  size_t max_epochs = 0, total_epochs = 0;
  for (uint32_t w = 0; w < nw; w++)
  {
    const Encoder& e = ctx->enc[work[w]];
    clut_base[w + 1] = clut_base[w] + e.clut_ops.size();
    max_epochs = std::max(max_epochs, e.epochs.size());
    total_epochs += e.epochs.size();
  }
  const uint64_t n_clut = clut_base[nw];

  if (ctx->group_n >= ctx->chunks.size())
    ctx->chunks.push_back(new psxgpu_ctx::Chunk());
  psxgpu_ctx::Chunk& S = *ctx->chunks[ctx->group_n++];
  CK(P.clutops.ensure(n_clut * sizeof(ClutOp)), "pinned alloc (clut ops)");
  CK(P.items.ensure(total_epochs * sizeof(WaveItem)), "pinned alloc (wave items)");
  CK(S.d_cmds.ensure(n_cmds * sizeof(PackedCmd)), "device alloc (commands)");
  CK(S.d_setups.ensure(n_cmds * sizeof(PrimSetup)), "device alloc (setup records)");
  CK(S.d_bins.ensure(n_cmds * sizeof(BinInfo) + 64), "device alloc (bins)");
  CK(S.d_tilecmds.ensure(tile_cmds_bytes(n_cmds)), "device alloc (tile command bitmaps)");
  CK(S.d_payload.ensure(n_pay * 2 + 16), "device alloc (payload)");
  CK(S.d_clutops.ensure(n_clut * sizeof(ClutOp) + 16), "device alloc (clut ops)");
  CK(S.d_items.ensure(total_epochs * 64 + 16), "device alloc (wave items / epoch tile sets)"); // (instance mode keeps 64 bytes of tile sets per epoch here)

  PackedCmd* hc = static_cast<PackedCmd*>(P.cmds.p);
  uint16_t* hp = static_cast<uint16_t*>(P.payload.p);
  ClutOp* ho = static_cast<ClutOp*>(P.clutops.p);
  for (uint32_t w = 0; w < nw; w++)
  {
    const Encoder& e = ctx->enc[work[w]];
    const uint64_t host_first = layout ? layout->host_cmd_base[w] : cmd_base[w];
    for (uint32_t ui : e.upload_cmds)
      hc[host_first + ui].upload.payload += static_cast<uint32_t>(pay_base[w]);
    if (!e.clut_ops.empty())
      std::memcpy(ho + clut_base[w], e.clut_ops.data(), e.clut_ops.size() * sizeof(ClutOp));
  }

  // waves: wave k = epoch k of every instance that has one
  WaveItem* hi = static_cast<WaveItem*>(P.items.p);
  S.wave_off.assign(1, 0);
  S.wave_has_clut.clear();
  S.wave_has_tiled.clear();
  S.wave_has_serial.clear();
  S.wave_tiled_cmds.clear();
  uint32_t n_items = 0;
  for (size_t k = 0; k < max_epochs; k++)
  {
    uint8_t has_clut = 0, has_tiled = 0, has_serial = 0;
    uint64_t tiled_cmds = 0;
    for (uint32_t w = 0; w < nw; w++)
    {
      const Encoder& e = ctx->enc[work[w]];
      if (e.epochs.size() <= k)
        continue;
      const Epoch& ep = e.epochs[k];
      WaveItem it{};
      it.instance = work[w];
      it.cmd_begin = static_cast<uint32_t>(cmd_base[w] + ep.cmd_begin);
      it.cmd_end = static_cast<uint32_t>(cmd_base[w] + ep.cmd_end);
      it.kind = ep.kind;
      it.clut_begin = static_cast<uint32_t>(clut_base[w] + ep.clut_begin);
      it.clut_end = static_cast<uint32_t>(clut_base[w] + ep.clut_end);
      has_clut |= (it.clut_end > it.clut_begin) ? 1 : 0;
      if (it.cmd_end > it.cmd_begin)
      {
        if (ep.kind == EPOCH_TILED)
        {
          has_tiled = 1;
          tiled_cmds += it.cmd_end - it.cmd_begin;
        }
        else
          has_serial = 1;
      }
      hi[n_items++] = it;
    }
    S.wave_off.push_back(n_items);
    S.wave_has_clut.push_back(has_clut);
    S.wave_has_tiled.push_back(has_tiled);
    S.wave_has_serial.push_back(has_serial);
    S.wave_tiled_cmds.push_back(static_cast<uint32_t>(std::min<uint64_t>(tiled_cmds, 0xFFFFFFFFull)));
  }
  S.n_cmds = static_cast<uint32_t>(n_cmds);
  S.device_encoded = false;
  S.n_instances = nw;
  const uint32_t n_waves = static_cast<uint32_t>(S.wave_off.size()) - 1;
  S.instance_mode = instance_mode_wanted(ctx, total_epochs, nw, n_cmds);
  S.instance_wide = nw <= 64;
  size_t inst_bytes = 0;
  if (S.instance_mode)
  {
    // the same epochs, instance by instance
    inst_bytes = static_cast<size_t>(nw) * sizeof(InstWork) + total_epochs * sizeof(Epoch);
    CK(P.inst.ensure(inst_bytes), "pinned alloc (instance table)");
    CK(S.d_instwork.ensure(static_cast<size_t>(nw) * sizeof(InstWork)), "device alloc (instance table)");
    CK(S.d_epochs.ensure(total_epochs * sizeof(Epoch) + 16), "device alloc (epochs)");
    InstWork* hw = static_cast<InstWork*>(P.inst.p);
    Epoch* he = reinterpret_cast<Epoch*>(hw + nw);
    uint32_t first = 0;
    for (uint32_t w = 0; w < nw; w++)
    {
      const Encoder& e = ctx->enc[work[w]];
      hw[w] = InstWork{work[w], static_cast<uint32_t>(cmd_base[w]), static_cast<uint32_t>(clut_base[w]), first,
                       static_cast<uint32_t>(e.epochs.size()), {0, 0, 0}};
      if (!e.epochs.empty())
        std::memcpy(he + first, e.epochs.data(), e.epochs.size() * sizeof(Epoch));
      first += static_cast<uint32_t>(e.epochs.size());
    }
  }
  S.persistent = !S.instance_mode && ctx->persist_max != 0 && nw <= ctx->persist_max && n_waves > 0 && !ctx->opts.profile;
  size_t misc_bytes = 0;
  if (S.persistent)
  {
    misc_bytes = static_cast<size_t>(n_waves) * 16 + static_cast<size_t>(nw) * 4;
    CK(P.misc.ensure(misc_bytes), "pinned alloc (wave table)");
    CK(S.d_misc.ensure(misc_bytes + 16), "device alloc (wave table)");
    uint32_t* wv = static_cast<uint32_t*>(P.misc.p);
    for (uint32_t k = 0; k < n_waves; k++)
    {
      wv[k * 4 + 0] = S.wave_off[k];
      wv[k * 4 + 1] = S.wave_off[k + 1] - S.wave_off[k];
      wv[k * 4 + 2] = S.wave_has_clut[k];
      wv[k * 4 + 3] = 0;
    }
    for (uint32_t w = 0; w < nw; w++)
      wv[n_waves * 4 + w] = work[w];
  }
  // everything the encoders held is in the pinned set now
  for (uint32_t w = 0; w < nw; w++)
  {
    ctx->enc[work[w]].begin_flush();
    ctx->dirty[work[w]] = 0;
  }

  if (layout)
  {
    for (const CopySeg& g : layout->cmd_segs)
      if (g.count)
        CK(cudaMemcpyAsync(static_cast<PackedCmd*>(S.d_cmds.p) + g.dev_off, hc + g.host_off, g.count * sizeof(PackedCmd),
                           cudaMemcpyHostToDevice, ctx->stream),
           "H2D commands");
    for (const CopySeg& g : layout->pay_segs)
      if (g.count)
        CK(cudaMemcpyAsync(static_cast<uint16_t*>(S.d_payload.p) + g.dev_off, hp + g.host_off, g.count * 2, cudaMemcpyHostToDevice,
                           ctx->stream),
           "H2D payload");
  }
  else
  {
    if (n_cmds)
      CK(cudaMemcpyAsync(S.d_cmds.p, hc, n_cmds * sizeof(PackedCmd), cudaMemcpyHostToDevice, ctx->stream), "H2D commands");
    if (n_pay)
      CK(cudaMemcpyAsync(S.d_payload.p, hp, n_pay * 2, cudaMemcpyHostToDevice, ctx->stream), "H2D payload");
  }
  if (n_clut)
    CK(cudaMemcpyAsync(S.d_clutops.p, ho, n_clut * sizeof(ClutOp), cudaMemcpyHostToDevice, ctx->stream), "H2D clut ops");
  if (n_items)
    CK(cudaMemcpyAsync(S.d_items.p, hi, n_items * sizeof(WaveItem), cudaMemcpyHostToDevice, ctx->stream), "H2D wave items");
  if (misc_bytes)
    CK(cudaMemcpyAsync(S.d_misc.p, P.misc.p, misc_bytes, cudaMemcpyHostToDevice, ctx->stream), "H2D wave table");
  if (inst_bytes)
  {
    CK(cudaMemcpyAsync(S.d_instwork.p, P.inst.p, static_cast<size_t>(nw) * sizeof(InstWork), cudaMemcpyHostToDevice, ctx->stream),
       "H2D instance table");
    if (total_epochs)
      CK(cudaMemcpyAsync(S.d_epochs.p, static_cast<const uint8_t*>(P.inst.p) + static_cast<size_t>(nw) * sizeof(InstWork),
                         total_epochs * sizeof(Epoch), cudaMemcpyHostToDevice, ctx->stream),
         "H2D epochs");
  }
  CK(cudaEventRecord(P.free_ev, ctx->stream), "event record");
  P.in_flight = true;
  ctx->counters.h2d_bytes += n_cmds * sizeof(PackedCmd) + n_pay * 2 + n_clut * sizeof(ClutOp) + n_items * sizeof(WaveItem);
  return run_chunk(ctx, S);
}

// Vector mode (records arrived one at a time through psxgpu_submit_wire): packs the finished encoders of `work` into a
// pinned set and launches them as one chunk.
int flush_work(psxgpu_ctx* ctx, const std::vector<uint32_t>& work)
{
  if (work.empty())
    return PSXGPU_OK;
  int rc = open_group(ctx);
  if (rc != PSXGPU_OK)
    return rc;
  const double t0 = now_ms();
  const uint32_t nw = static_cast<uint32_t>(work.size());
  parallel_for(nw, ctx->threads, [&](uint32_t w) { ctx->enc[work[w]].finish_flush(); });
  std::vector<uint64_t> cmd_base(nw + 1, 0), pay_base(nw + 1, 0);
  for (uint32_t w = 0; w < nw; w++)
  {
    const Encoder& e = ctx->enc[work[w]];
    cmd_base[w + 1] = cmd_base[w] + e.n_cmds();
    pay_base[w + 1] = pay_base[w] + e.n_payload();
  }
  psxgpu_ctx::PinnedSet* P = nullptr;
  rc = acquire_pinned(ctx, P);
  if (rc != PSXGPU_OK)
    return rc;
  CK(P->cmds.ensure(cmd_base[nw] * sizeof(PackedCmd)), "pinned alloc (commands)");
  CK(P->payload.ensure(pay_base[nw] * 2), "pinned alloc (payload)");
  PackedCmd* hc = static_cast<PackedCmd*>(P->cmds.p);
  uint16_t* hp = static_cast<uint16_t*>(P->payload.p);
  parallel_for(nw, ctx->threads, [&](uint32_t w) {
    const Encoder& e = ctx->enc[work[w]];
    if (!e.cmds.empty())
      std::memcpy(hc + cmd_base[w], e.cmds.data(), e.cmds.size() * sizeof(PackedCmd));
    if (!e.payload.empty())
      std::memcpy(hp + pay_base[w], e.payload.data(), e.payload.size() * 2);
  });
  ctx->stage_ms_acc += now_ms() - t0;
  return launch_chunk(ctx, work, cmd_base, pay_base, *P);
}


// ---------------------------------------------------------------------------------------------------------------
// Batched path with device-side encoding (device_encoder.h).  Per chunk the host validates the record framing, copies
// the raw streams into pinned memory and uploads them on copy_stream; enc_stream encodes them (one warp per instance)
// and builds the wave table; the host reads the small table back and launches setup + waves on the main stream.  The
// read-back of chunk k is awaited after chunk k+1 has been staged, so the host never idles on the device.
void ensure_host_state(psxgpu_ctx* ctx, uint32_t i)
{
  if (ctx->blob_truth[i])
  {
    ctx->enc[i].import_state(ctx->enc_blob[i]);
    ctx->blob_truth[i] = 0;
  }
}

struct PendingChunk
{
  psxgpu_ctx::Chunk* S = nullptr;
  psxgpu_ctx::PinnedSet* P = nullptr;
  cudaStream_t es = nullptr;
  psxgpu_ctx::TraceChunk trace{};
  uint32_t first_instance = 0, cn = 0, wave_cap = 0;
  size_t waves_off = 0, state_off = 0;
  uint64_t cmd_total = 0;
};

// True if [p, p + bytes) starts and ends in page-locked host memory (cudaHostAlloc / cudaHostRegister).
bool is_pinned_host(const void* p, size_t bytes)
{
  if (!p || bytes == 0)
    return false;
  cudaPointerAttributes a0{}, a1{};
  if (cudaPointerGetAttributes(&a0, p) != cudaSuccess || cudaPointerGetAttributes(&a1, static_cast<const uint8_t*>(p) + bytes - 1) != cudaSuccess)
  {
    cudaGetLastError();
    return false;
  }
  return a0.type == cudaMemoryTypeHost && a1.type == cudaMemoryTypeHost;
}

// Stage A of a device-encoded chunk: upload the raw streams (copy_stream) and check their framing on the device
// (walk_kernel on the chunk's encode stream); the per-instance counts come back into the pinned set.
int stage_upload(psxgpu_ctx* ctx, PendingChunk& pc, uint32_t chunk_index, const void* const* streams, const size_t* bytes)
{
  const uint32_t cn = pc.cn;
  psxgpu_ctx::PinnedSet* P = nullptr;
  int rc = acquire_pinned(ctx, P, true);
  if (rc != PSXGPU_OK)
    return rc;
  pc.P = P;
  const double t0 = now_ms();
  // Where the streams go in the chunk's device wire buffer.  Streams that already sit in page-locked memory are
  // copied from where they are (neighbours in one copy); anything else is packed into our pinned staging buffer.
  bool direct = true;
  for (uint32_t k = 0; k < cn && direct; k++)
    direct = bytes[k] == 0 || ((reinterpret_cast<uintptr_t>(streams[k]) & 15u) == 0 && is_pinned_host(streams[k], bytes[k]));
  struct Run
  {
    const uint8_t* src;
    uint64_t dev_off, bytes;
  };
  std::vector<Run> runs;
  std::vector<uint64_t> wire_off(cn + 1, 0);
  if (direct)
  {
    uint64_t dev = 0;
    for (uint32_t k = 0; k < cn; k++)
    {
      const uint8_t* src = static_cast<const uint8_t*>(streams[k]);
      const uint64_t n = bytes[k];
      if (n == 0)
      {
        wire_off[k] = dev;
        continue;
      }
      if (!runs.empty() && src >= runs.back().src + runs.back().bytes && src - (runs.back().src + runs.back().bytes) < 64)
      {
        wire_off[k] = runs.back().dev_off + static_cast<uint64_t>(src - runs.back().src);
        runs.back().bytes = static_cast<uint64_t>(src - runs.back().src) + n;
      }
      else
      {
        dev = runs.empty() ? 0 : ((runs.back().dev_off + runs.back().bytes + 15u) & ~static_cast<uint64_t>(15));
        wire_off[k] = dev;
        runs.push_back(Run{src, dev, n});
      }
    }
    wire_off[cn] = runs.empty() ? 0 : ((runs.back().dev_off + runs.back().bytes + 15u) & ~static_cast<uint64_t>(15));
  }
  else
  {
    for (uint32_t k = 0; k < cn; k++)
      wire_off[k + 1] = wire_off[k] + ((bytes[k] + 15u) & ~static_cast<uint64_t>(15));
  }
  for (uint32_t k = 0; k < cn; k++)
    if (bytes[k] > 0xFFFFFF00ull)
      return ctx->invalid("submit_batch: stream larger than 4 GiB");
  if (wire_off[cn] > 0xFFFFFF00ull)
    return ctx->invalid("submit_batch: chunk larger than 4 GiB of records; use a smaller batch chunk");
  const size_t wire_bytes = static_cast<size_t>(wire_off[cn]) + 256;
  CK(P->inst.ensure(static_cast<size_t>(cn) * sizeof(EncInst)), "pinned alloc (instances)");
  CK(P->recoff.ensure(static_cast<size_t>(cn) * sizeof(WalkOut)), "pinned alloc (framing)");
  CK(P->state.ensure(static_cast<size_t>(cn) * sizeof(EncState)), "pinned alloc (encoder state)");
  EncInst* hi = static_cast<EncInst*>(P->inst.p);
  EncState* hs = static_cast<EncState*>(P->state.p);
  uint64_t rec_cap = 0;
  for (uint32_t k = 0; k < cn; k++)
  {
    std::memset(&hi[k], 0, sizeof(EncInst));
    hi[k].instance = pc.first_instance + k;
    hi[k].wire_off = static_cast<uint32_t>(wire_off[k]);
    hi[k].wire_bytes = static_cast<uint32_t>(bytes[k]);
    hi[k].rec_base = static_cast<uint32_t>(rec_cap);
    rec_cap += bytes[k] / 8 + 2;
    const uint32_t inst = pc.first_instance + k;
    if (ctx->blob_truth[inst])
      hs[k] = ctx->enc_blob[inst];
    else
      ctx->enc[inst].export_state(hs[k]);
  }
  uint8_t* hw = nullptr;
  if (!direct)
  {
    CK(P->wire.ensure(wire_bytes), "pinned alloc (wire)");
    hw = static_cast<uint8_t*>(P->wire.p);
    parallel_for(cn, ctx->threads, [&](uint32_t k) {
      if (bytes[k])
        std::memcpy(hw + wire_off[k], streams[k], bytes[k]);
    });
    std::memset(hw + wire_off[cn], 0, 256);
  }
  ctx->encode_ms_acc += now_ms() - t0;

  const double t2 = now_ms();
  if (ctx->group_n >= ctx->chunks.size())
    ctx->chunks.push_back(new psxgpu_ctx::Chunk());
  psxgpu_ctx::Chunk& S = *ctx->chunks[ctx->group_n++];
  pc.S = &S;
  pc.es = ctx->enc_streams[chunk_index % 4];
  for (cudaEvent_t* ev : {&S.ev_copied, &S.ev_walked, &S.ev_encoded, &S.ev_done})
    if (!*ev)
      CK(cudaEventCreateWithFlags(ev, cudaEventDisableTiming), "event create");
  if (!S.done_valid) // first device-encoded use of this slot: anything queued on the main stream may still read it
  {
    CK(cudaEventRecord(S.ev_done, ctx->stream), "event record");
    S.done_valid = true;
  }
  CK(cudaStreamWaitEvent(ctx->copy_stream, S.ev_done, 0), "stream wait");
  CK(S.d_payload.ensure(wire_bytes), "device alloc (wire)");
  CK(S.d_recoff.ensure(static_cast<size_t>(cn) * sizeof(WalkOut)), "device alloc (framing)");
  CK(S.d_rec.ensure(rec_cap * 4 + 16), "device alloc (record offsets)");
  CK(S.d_inst.ensure(static_cast<size_t>(cn) * sizeof(EncInst)), "device alloc (instances)");
  CK(S.d_state.ensure(static_cast<size_t>(cn) * sizeof(EncState)), "device alloc (encoder state)");
  if (ctx->trace)
  {
    for (cudaEvent_t* ev : {&pc.trace.copy0, &pc.trace.copy1, &pc.trace.enc1, &pc.trace.run0, &pc.trace.run1})
      cudaEventCreate(ev);
    if (!ctx->trace_origin)
    {
      cudaEventCreate(&ctx->trace_origin);
      cudaEventRecord(ctx->trace_origin, ctx->stream);
    }
    cudaEventRecord(pc.trace.copy0, ctx->copy_stream);
  }
  uint8_t* dw = static_cast<uint8_t*>(S.d_payload.p);
  uint64_t copied = 0;
  // the small tables go first: nothing small should ever queue behind a 150 MB copy
  CK(cudaMemcpyAsync(S.d_inst.p, hi, static_cast<size_t>(cn) * sizeof(EncInst), cudaMemcpyHostToDevice, ctx->copy_stream), "H2D instances");
  CK(cudaMemcpyAsync(S.d_state.p, hs, static_cast<size_t>(cn) * sizeof(EncState), cudaMemcpyHostToDevice, ctx->copy_stream), "H2D encoder state");
  if (direct)
  {
    for (const Run& r : runs)
    {
      CK(cudaMemcpyAsync(dw + r.dev_off, r.src, r.bytes, cudaMemcpyHostToDevice, ctx->copy_stream), "H2D wire");
      copied += r.bytes;
    }
    CK(cudaMemsetAsync(dw + wire_off[cn], 0, 256, ctx->copy_stream), "wire slack");
  }
  else
  {
    CK(cudaMemcpyAsync(dw, hw, wire_bytes, cudaMemcpyHostToDevice, ctx->copy_stream), "H2D wire");
    copied += wire_bytes;
  }
  if (ctx->trace)
    cudaEventRecord(pc.trace.copy1, ctx->copy_stream);
  CK(cudaEventRecord(S.ev_copied, ctx->copy_stream), "event record");
  CK(cudaEventRecord(P->free_ev, ctx->copy_stream), "event record");
  P->in_flight = true;
  ctx->counters.h2d_bytes += copied + static_cast<uint64_t>(cn) * (sizeof(EncInst) + sizeof(EncState));

  CK(cudaStreamWaitEvent(ctx->walk_stream, S.ev_copied, 0), "stream wait");
  WalkOut* d_wo = static_cast<WalkOut*>(S.d_recoff.p);
  CK(launch_walk(dw, static_cast<const EncInst*>(S.d_inst.p), d_wo, static_cast<uint32_t*>(S.d_rec.p), cn, ctx->walk_stream), "walk_kernel");
  ctx->counters.kernel_launches++;
  CK(cudaMemcpyAsync(P->recoff.p, d_wo, static_cast<size_t>(cn) * sizeof(WalkOut), cudaMemcpyDeviceToHost, ctx->walk_stream), "D2H framing result");
  CK(cudaEventRecord(S.ev_walked, ctx->walk_stream), "event record");
  ctx->counters.d2h_bytes += static_cast<uint64_t>(cn) * sizeof(WalkOut);
  ctx->launch_ms_acc += now_ms() - t2;
  return PSXGPU_OK;
}

// Stage B: the framing result is back — report errors, size the chunk's buffers exactly, launch encode + wave kernels.
int stage_encode(psxgpu_ctx* ctx, PendingChunk& pc)
{
  psxgpu_ctx::Chunk& S = *pc.S;
  psxgpu_ctx::PinnedSet* P = pc.P;
  const uint32_t cn = pc.cn;
  const double t0 = now_ms();
  CK(cudaEventSynchronize(S.ev_walked), "framing sync");
  ctx->wait_ms_acc += now_ms() - t0;
  const double t1 = now_ms();
  const WalkOut* wo = static_cast<const WalkOut*>(P->recoff.p);
  for (uint32_t k = 0; k < cn; k++)
    if (wo[k].status != 0)
    {
      ctx->error = (wo[k].status == PSXGPU_E_MALFORMED) ? "submit_batch: malformed record stream"
                                                       : "submit_batch: read-back / state records are not allowed in a batch";
      return wo[k].status == PSXGPU_E_MALFORMED ? PSXGPU_E_MALFORMED : PSXGPU_E_UNSUPPORTED;
    }
  uint64_t cmd = 0, clut = 0, epoch = 0;
  uint32_t wave_cap = 1;
  for (uint32_t k = 0; k < cn; k++)
  {
    const uint32_t ecap = wo[k].cmd_cap + wo[k].clut_cap + 1;
    cmd += wo[k].cmd_cap;
    clut += wo[k].clut_cap;
    epoch += ecap;
    wave_cap = std::max(wave_cap, ecap);
  }
  if (cmd > 0xFFFFFFF0ull || epoch > 0xFFFFFFF0ull)
    return ctx->invalid("submit_batch: chunk holds more than 2^32 records; use a smaller batch chunk");
  const size_t waves_off = 256, waves_inline = std::min<size_t>(wave_cap, ENC_WAVES_INLINE);
  const size_t state_off = waves_off + ENC_WAVES_INLINE * sizeof(EncWave);
  CK(P->result.ensure(state_off + static_cast<size_t>(cn) * sizeof(EncState)), "pinned alloc (encode result)");
  CK(S.d_state_out.ensure(static_cast<size_t>(cn) * sizeof(EncState)), "device alloc (encoder state)");
  CK(S.d_cmds.ensure(cmd * sizeof(PackedCmd) + 64), "device alloc (commands)");
  CK(S.d_setups.ensure(cmd * sizeof(PrimSetup) + 64), "device alloc (setup records)");
  CK(S.d_bins.ensure(cmd * sizeof(BinInfo) + 64), "device alloc (bins)");
  CK(S.d_tilecmds.ensure(tile_cmds_bytes(cmd)), "device alloc (tile command bitmaps)");
  CK(S.d_rects.ensure(cmd * 16 + 64), "device alloc (footprints)");
  CK(S.d_clutops.ensure(clut * sizeof(ClutOp) + 16), "device alloc (clut ops)");
  CK(S.d_epochs.ensure(epoch * sizeof(Epoch) + 16), "device alloc (epochs)");
  CK(S.d_items.ensure(static_cast<size_t>(epoch) * 64 + 16), "device alloc (wave items / epoch tile sets)");
  CK(S.d_result.ensure(static_cast<size_t>(cn) * sizeof(EncResult)), "device alloc (encode result)");
  CK(S.d_waves.ensure(static_cast<size_t>(wave_cap) * sizeof(EncWave)), "device alloc (wave table)");
  CK(S.d_cursor.ensure(static_cast<size_t>(wave_cap) * 4), "device alloc (wave cursors)");
  CK(S.d_summary.ensure(sizeof(EncSummary)), "device alloc (summary)");
  CK(S.d_instwork.ensure(static_cast<size_t>(cn) * sizeof(InstWork)), "device alloc (instance table)");
  // no host-to-device copy here: it would queue behind the next chunk's wire upload
  CK(cudaStreamWaitEvent(pc.es, S.ev_walked, 0), "stream wait");
  CK(launch_layout(static_cast<const WalkOut*>(S.d_recoff.p), static_cast<EncInst*>(S.d_inst.p), cn, pc.es), "layout_kernel");
  ctx->counters.kernel_launches++;
  EncodeArgs a{};
  a.wire = static_cast<const uint8_t*>(S.d_payload.p);
  a.rec_off = static_cast<const uint32_t*>(S.d_rec.p);
  a.walk = static_cast<const WalkOut*>(S.d_recoff.p);
  a.inst = static_cast<const EncInst*>(S.d_inst.p);
  a.n_inst = cn;
  a.clut_slots = ctx->clut_slots;
  a.modulation_crop = ctx->opts.modulation_crop ? 1u : 0u;
  a.state_in = static_cast<const EncState*>(S.d_state.p);
  a.state_out = static_cast<EncState*>(S.d_state_out.p);
  a.cmds = static_cast<PackedCmd*>(S.d_cmds.p);
  a.rects = static_cast<uint4*>(S.d_rects.p);
  a.clut_ops = static_cast<ClutOp*>(S.d_clutops.p);
  a.epochs = static_cast<Epoch*>(S.d_epochs.p);
  a.result = static_cast<EncResult*>(S.d_result.p);
  a.summary = static_cast<EncSummary*>(S.d_summary.p);
  a.waves = static_cast<EncWave*>(S.d_waves.p);
  a.wave_cursor = static_cast<uint32_t*>(S.d_cursor.p);
  a.wave_cap = wave_cap;
  a.items = static_cast<WaveItem*>(S.d_items.p);
  a.inst_work = static_cast<InstWork*>(S.d_instwork.p);
  CK(launch_encode(a, pc.es), "encode_kernel");
  ctx->counters.kernel_launches += 3;
  S.enc_args = a;
  S.device_encoded = true;
  uint8_t* res = static_cast<uint8_t*>(P->result.p);
  CK(cudaMemcpyAsync(res, S.d_summary.p, sizeof(EncSummary), cudaMemcpyDeviceToHost, pc.es), "D2H summary");
  CK(cudaMemcpyAsync(res + waves_off, S.d_waves.p, waves_inline * sizeof(EncWave), cudaMemcpyDeviceToHost, pc.es), "D2H wave table");
  CK(cudaMemcpyAsync(res + state_off, S.d_state_out.p, static_cast<size_t>(cn) * sizeof(EncState), cudaMemcpyDeviceToHost, pc.es),
     "D2H encoder state");
  if (ctx->trace)
    cudaEventRecord(pc.trace.enc1, pc.es);
  CK(cudaEventRecord(S.ev_encoded, pc.es), "event record");
  ctx->counters.d2h_bytes += sizeof(EncSummary) + waves_inline * sizeof(EncWave) + static_cast<uint64_t>(cn) * sizeof(EncState);
  pc.wave_cap = wave_cap;
  pc.waves_off = waves_off;
  pc.state_off = state_off;
  pc.cmd_total = cmd;
  ctx->launch_ms_acc += now_ms() - t1;
  return PSXGPU_OK;
}

// Stage C: the wave table is back — launch setup + waves on the main stream.
int finalize_device_chunk(psxgpu_ctx* ctx, PendingChunk& pc)
{
  psxgpu_ctx::Chunk& S = *pc.S;
  const double t0 = now_ms();
  CK(cudaEventSynchronize(S.ev_encoded), "encode sync");
  ctx->wait_ms_acc += now_ms() - t0;
  const double t1 = now_ms();
  const uint8_t* res = static_cast<const uint8_t*>(pc.P->result.p);
  const EncSummary* sm = reinterpret_cast<const EncSummary*>(res);
  if (sm->overflow)
    return ctx->invalid("device encoder: output capacity exceeded");
  const uint32_t M = std::min(sm->max_epochs, pc.wave_cap);
  const EncWave* wv = reinterpret_cast<const EncWave*>(res + pc.waves_off);
  std::vector<EncWave> big;
  if (M > ENC_WAVES_INLINE)
  {
    big.resize(M);
    CK(cudaMemcpyAsync(big.data(), S.d_waves.p, static_cast<size_t>(M) * sizeof(EncWave), cudaMemcpyDeviceToHost, pc.es),
       "D2H wave table");
    CK(cudaStreamSynchronize(pc.es), "encode sync");
    ctx->counters.d2h_bytes += static_cast<uint64_t>(M) * sizeof(EncWave);
    wv = big.data();
  }
  S.wave_off.assign(1, 0);
  S.wave_has_clut.clear();
  S.wave_has_tiled.clear();
  S.wave_has_serial.clear();
  S.wave_tiled_cmds.clear();
  for (uint32_t k = 0; k < M; k++)
  {
    S.wave_off.push_back(wv[k].item_off + wv[k].item_count);
    S.wave_has_clut.push_back((wv[k].flags & 1u) ? 1 : 0);
    S.wave_has_tiled.push_back((wv[k].flags & 2u) ? 1 : 0);
    S.wave_has_serial.push_back((wv[k].flags & 4u) ? 1 : 0);
    S.wave_tiled_cmds.push_back(wv[k].tiled_cmds);
  }
  S.n_cmds = static_cast<uint32_t>(pc.cmd_total);
  S.n_instances = pc.cn;
  S.persistent = false;
  // many short epochs per instance (render-to-texture feedback): a launch per epoch costs more than the epochs' work
  S.instance_mode = instance_mode_wanted(ctx, sm->total_items, pc.cn, pc.cmd_total) && M <= pc.wave_cap;
  S.instance_wide = pc.cn <= 64;
  const EncState* states = reinterpret_cast<const EncState*>(res + pc.state_off);
  for (uint32_t k = 0; k < pc.cn; k++)
  {
    ctx->enc_blob[pc.first_instance + k] = states[k];
    ctx->blob_truth[pc.first_instance + k] = 1;
  }
  psxgpu_stats& c = ctx->counters;
  c.wire_commands += sm->stats[ENC_STAT_WIRE];
  c.packed_prims += sm->stats[ENC_STAT_PRIMS];
  c.epochs += sm->stats[ENC_STAT_EPOCHS];
  c.cuts_raw += sm->stats[ENC_STAT_CUTS_RAW];
  c.cuts_war += sm->stats[ENC_STAT_CUTS_WAR];
  c.cuts_clut += sm->stats[ENC_STAT_CUTS_CLUT];
  c.serial_self_sampling += sm->stats[ENC_STAT_SELF];
  c.serial_copy += sm->stats[ENC_STAT_COPY];
  c.serial_upload += sm->stats[ENC_STAT_UPLOAD];
  c.clut_ops += sm->stats[ENC_STAT_CLUT_OPS];
  c.clut_cache_hits += sm->stats[ENC_STAT_CLUT_HITS];
  c.rejected += sm->stats[ENC_STAT_REJECTED];
  CK(cudaStreamWaitEvent(ctx->stream, S.ev_encoded, 0), "stream wait");
  if (ctx->trace)
    cudaEventRecord(pc.trace.run0, ctx->stream);
  const int rc = run_chunk(ctx, S);
  if (ctx->trace)
  {
    cudaEventRecord(pc.trace.run1, ctx->stream);
    ctx->trace_chunks.push_back(pc.trace);
  }
  ctx->launch_ms_acc += now_ms() - t1;
  return rc;
}

// Up to six chunks are in flight: while chunk k is being uploaded and framed, older chunks are sized and encoded and
// still older ones have their waves launched — the host never walks a record, and it polls for whichever result
// arrives first.
int submit_batch_device(psxgpu_ctx* ctx, uint32_t first, uint32_t count, const void* const* streams, const size_t* bytes)
{
  uint32_t per = ctx->batch_chunk;
  if (per == 0)
  {
    static const uint32_t div = []() {
      const char* e = std::getenv("PSXGPU_CHUNK_DIV"); // experiments: chunks per batch in automatic mode
      const int v = e ? std::atoi(e) : 0;
      return v > 0 ? static_cast<uint32_t>(v) : 12u; // (e2e step of 4096 instances: 8 chunks 26.95 ms, 12 chunks 26.45, 16 chunks 26.65; 3-6 chunks 27.4-28.8)
    }();
    per = (count <= 256) ? count : std::max<uint32_t>(128, (count + div - 1) / div);
  }
  std::vector<PendingChunk> uploaded, encoded; // waiting for stage B / stage C, oldest first
  int status = PSXGPU_OK;
  auto finalize_front = [&]() {
    PendingChunk pc = encoded.front();
    encoded.erase(encoded.begin());
    const int rc = finalize_device_chunk(ctx, pc);
    if (rc != PSXGPU_OK && status == PSXGPU_OK)
      status = rc;
  };
  auto advance = [&](size_t keep_uploaded, size_t keep_encoded) -> int {
    while (uploaded.size() > keep_uploaded)
    {
      PendingChunk pc = uploaded.front();
      uploaded.erase(uploaded.begin());
      if (status != PSXGPU_OK)
        continue;
      // wait for the framing result — and launch the waves of an older chunk the moment its encode completes, instead
      // of leaving the main stream idle until this wait is over
      const double t0 = now_ms();
      for (;;)
      {
        if (!encoded.empty() && cudaEventQuery(encoded.front().S->ev_encoded) == cudaSuccess)
        {
          finalize_front();
          continue;
        }
        const cudaError_t q = cudaEventQuery(pc.S->ev_walked);
        if (q == cudaSuccess)
          break;
        if (q != cudaErrorNotReady)
        {
          status = ctx->fail(q, "framing sync");
          break;
        }
        std::this_thread::yield();
      }
      ctx->wait_ms_acc += now_ms() - t0;
      if (status != PSXGPU_OK)
        continue;
      status = stage_encode(ctx, pc);
      if (status == PSXGPU_OK)
        encoded.push_back(pc);
    }
    while (encoded.size() > keep_encoded)
      finalize_front();
    return status;
  };
  uint32_t chunk_index = 0;
  static const bool taper = []() {
    const char* e = std::getenv("PSXGPU_CHUNK_TAPER");
    return !e || std::atoi(e) != 0;
  }();
  const bool ramp = ctx->batch_chunk == 0 && count > 256; // automatic chunking: small first chunks fill the pipeline sooner
  for (uint32_t c0 = 0, cn = 0; c0 < count && status == PSXGPU_OK; c0 += cn, chunk_index++)
  {
    uint32_t want = per;
    if (ramp && chunk_index == 0)
      want = std::max<uint32_t>(64, per / 4);
    else if (ramp && chunk_index == 1)
      want = std::max<uint32_t>(64, per / 2);
    else if (ramp && count - c0 <= per + per / 2 && taper)
      want = std::max<uint32_t>(128, (count - c0) / 2); // ... and small last chunks shorten what is left to do after the last copy
    cn = std::min(want, count - c0);
    if (ramp && count - c0 - cn < 64)
      cn = count - c0; // (no crumbs)
    int rc = open_group(ctx);
    if (rc != PSXGPU_OK)
      return rc;
    // the pinned sets rotate by six: chunk k reuses the set of chunk k-6, whose waves must have been launched
    advance(2, 3);
    if (status != PSXGPU_OK)
      break;
    PendingChunk pc;
    pc.first_instance = first + c0;
    pc.cn = cn;
    rc = stage_upload(ctx, pc, chunk_index, streams + c0, bytes + c0);
    if (rc != PSXGPU_OK)
    {
      status = rc;
      break;
    }
    uploaded.push_back(pc);
  }
  advance(0, 0);
  return status;
}

// Flushes every instance that has queued work and closes the group.
int do_flush(psxgpu_ctx* ctx)
{
  std::vector<uint32_t> work;
  for (uint32_t i = 0; i < ctx->n; i++)
    if (ctx->dirty[i])
      work.push_back(i);
  int rc = flush_work(ctx, work);
  if (rc != PSXGPU_OK)
    return rc;
  return close_group(ctx);
}

void scanout_params(const psxgpu_ctx* ctx, const psx_update_display* cmd, ScanoutParams& p, bool& disabled)
{
  // GPU_SW::UpdateDisplay, gpu_sw.cpp:391-448
  std::memset(&p, 0, sizeof(p));
  disabled = false;
  if (ctx->opts.show_vram)
  {
    p.width = VRAM_W;
    p.height = VRAM_H;
    p.is24 = 0;
  }
  else
  {
    if (cmd->flags & PSX_UD_DISPLAY_DISABLED)
    {
      disabled = true;
      return;
    }
    p.is24 = (cmd->flags & PSX_UD_DISPLAY_24BIT) ? 1 : 0;
    const uint32_t inter = (cmd->flags & PSX_UD_INTERLACED_INTERLEAVED) ? 1u : 0u;
    const uint32_t field = (cmd->flags & PSX_UD_INTERLACED_FIELD) ? 1u : 0u;
    p.src_x = p.is24 ? cmd->X : cmd->display_vram_left;
    p.skip_x = p.is24 ? static_cast<uint32_t>(cmd->display_vram_left - cmd->X) : 0;
    p.src_y = cmd->display_vram_top + (field & inter);
    p.width = cmd->display_vram_width;
    p.height = cmd->display_vram_height;
    p.line_skip = (cmd->flags & PSX_UD_INTERLACED_ENABLED) ? inter : 0;
  }
  if (p.is24)
    p.fast = ((p.src_x + p.width) <= VRAM_W && (p.src_y + (p.height << p.line_skip)) <= VRAM_H) ? 1 : 0; // gpu_sw.cpp:305
  else
    p.fast = ((p.src_x + p.width) <= VRAM_W && (p.src_y + p.height) <= VRAM_H) ? 1 : 0; // gpu_sw.cpp:241
  p.format = p.is24 ? 0 : ctx->opts.display_format;
  const uint32_t bpp = (p.is24 || p.format == PSXGPU_FMT_RGBA8) ? 4 : 2;
  p.out_stride = (p.width * bpp + 3u) & ~3u; // gpu_sw.cpp:236 / :301
}

int do_scanout(psxgpu_ctx* ctx, uint32_t instance, const psx_update_display* cmd, void* dst, uint32_t dst_stride, uint32_t* ow,
               uint32_t* oh, uint32_t* obpp)
{
  ScanoutParams p;
  bool disabled;
  scanout_params(ctx, cmd, p, disabled);
  if (disabled)
    return 1;
  if (p.width > 4096 || p.height > 4096)
    return ctx->invalid("scan-out size out of range");
  const uint32_t bpp = (p.is24 || p.format == PSXGPU_FMT_RGBA8) ? 4 : 2;
  if (ow)
    *ow = p.width;
  if (oh)
    *oh = p.height;
  if (obpp)
    *obpp = bpp;
  if (p.width == 0 || p.height == 0)
    return PSXGPU_OK;
  if (dst_stride < p.width * bpp)
    return ctx->invalid("scan-out destination stride too small");
  int rc = do_flush(ctx);
  if (rc != PSXGPU_OK)
    return rc;
  CK(ctx->d_scan.ensure(static_cast<size_t>(p.out_stride) * p.height), "device alloc (scan-out)");
  CK(launch_scanout(inst_vram(ctx, instance), p, static_cast<uint8_t*>(ctx->d_scan.p), ctx->stream), "scanout_kernel");
  ctx->counters.kernel_launches++;
  CK(cudaMemcpy2DAsync(dst, dst_stride, ctx->d_scan.p, p.out_stride, static_cast<size_t>(p.width) * bpp, p.height,
                       cudaMemcpyDeviceToHost, ctx->stream),
     "D2H scan-out");
  CK(cudaStreamSynchronize(ctx->stream), "stream sync");
  ctx->counters.d2h_bytes += static_cast<uint64_t>(p.width) * bpp * p.height;
  return PSXGPU_OK;
}

int do_read_vram(psxgpu_ctx* ctx, uint32_t instance, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint16_t* dst, uint32_t pitch)
{
  if (w == 0 || h == 0)
    return PSXGPU_OK;
  if (w > VRAM_W || h > VRAM_H || pitch < w)
    return ctx->invalid("read_vram: rectangle larger than VRAM or pitch too small");
  int rc = do_flush(ctx);
  if (rc != PSXGPU_OK)
    return rc;
  x &= VRAM_W - 1;
  y &= VRAM_H - 1;
  // up to 2x2 pieces when the rectangle wraps (GPUREAD wraps, gpu.cpp:1896-1929)
  const uint32_t w0 = std::min(w, VRAM_W - x), h0 = std::min(h, VRAM_H - y);
  const uint32_t xs[2] = {x, 0}, ws[2] = {w0, w - w0}, ys[2] = {y, 0}, hs[2] = {h0, h - h0};
  for (int j = 0; j < 2; j++)
    for (int i = 0; i < 2; i++)
    {
      if (ws[i] == 0 || hs[j] == 0)
        continue;
      uint16_t* d = dst + static_cast<size_t>(j ? h0 : 0) * pitch + (i ? w0 : 0);
      CK(cudaMemcpy2DAsync(d, static_cast<size_t>(pitch) * 2, inst_vram(ctx, instance) + ys[j] * VRAM_W + xs[i], VRAM_W * 2,
                           static_cast<size_t>(ws[i]) * 2, hs[j], cudaMemcpyDeviceToHost, ctx->stream),
         "D2H read_vram");
    }
  CK(cudaStreamSynchronize(ctx->stream), "stream sync");
  ctx->counters.d2h_bytes += static_cast<uint64_t>(w) * h * 2;
  return PSXGPU_OK;
}

} // namespace

extern "C" {

int psxgpu_abi_version(void)
{
  return PSXGPU_ABI_VERSION;
}

int psxgpu_create(int device, uint32_t n_instances, const psxgpu_opts* opts, psxgpu_ctx** out)
{
  if (!out || n_instances == 0)
    return PSXGPU_E_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device < 0 || device >= count)
    return PSXGPU_E_CUDA; // no CPU fallback: the product path needs a CUDA device
  psxgpu_ctx* ctx = new (std::nothrow) psxgpu_ctx();
  if (!ctx)
    return PSXGPU_E_NOMEM;
  ctx->device = device;
  ctx->n = n_instances;
  if (opts)
    std::memcpy(&ctx->opts, opts, std::min<size_t>(opts->struct_size ? opts->struct_size : sizeof(psxgpu_opts), sizeof(psxgpu_opts)));
  ctx->clut_slots = ctx->opts.clut_slots ? std::min(256u, std::max(4u, ctx->opts.clut_slots)) : 256u;
  ctx->threads = ctx->opts.encode_threads ? ctx->opts.encode_threads : std::max(1u, std::thread::hardware_concurrency());
  if (ctx->opts.display_format > PSXGPU_FMT_RGB565)
    ctx->opts.display_format = PSXGPU_FMT_RGBA8;

  auto bail = [&](cudaError_t e, const char* what) {
    std::fprintf(stderr, "psxgpu_create: %s: %s\n", what, cudaGetErrorString(e));
    psxgpu_destroy(ctx);
    return PSXGPU_E_CUDA;
  };
  cudaError_t e;
  if ((e = cudaSetDevice(device)) != cudaSuccess)
    return bail(e, "cudaSetDevice");
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
    return bail(e, "stream");
  // Stream priorities: measured (profiles/README.md), high-priority encode streams make the co-running raster waves
  // 60 % slower for the whole time an encode kernel is resident (its 168-register CTAs displace ten one-warp raster
  // CTAs each); at equal priority the encodes take longer but six chunks in flight hide that.  PSXGPU_ENC_PRIO=1
  // restores the high priority for experiments.
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  {
    const char* pe = std::getenv("PSXGPU_ENC_PRIO");
    if (!(pe && pe[0] == '1'))
      prio_hi = prio_lo;
  }
  if ((e = cudaStreamCreateWithPriority(&ctx->copy_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess)
    return bail(e, "stream");
  {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if ((e = cudaStreamCreateWithPriority(&ctx->walk_stream, cudaStreamNonBlocking, hi)) != cudaSuccess)
      return bail(e, "stream");
  }
  for (cudaStream_t& es : ctx->enc_streams)
    if ((e = cudaStreamCreateWithPriority(&es, cudaStreamNonBlocking, prio_hi)) != cudaSuccess)
      return bail(e, "stream");
  ctx->encode_mode = ctx->opts.reserved[1];
  ctx->trace = std::getenv("PSXGPU_TRACE") != nullptr;
  if ((e = cudaMalloc(&ctx->d_vram, static_cast<size_t>(n_instances) * VRAM_PIXELS * 2)) != cudaSuccess)
    return bail(e, "VRAM pool");
  if ((e = cudaMalloc(&ctx->d_clut, static_cast<size_t>(n_instances) * ctx->clut_slots * CLUT_ENTRIES * 2)) != cudaSuccess)
    return bail(e, "CLUT pool");
  if ((e = cudaMalloc(&ctx->d_hash, static_cast<size_t>(n_instances) * 8)) != cudaSuccess)
    return bail(e, "hash buffer");
  if ((e = cudaMalloc(&ctx->d_barrier, 64)) != cudaSuccess)
    return bail(e, "barrier counter");
  if ((e = cudaMemsetAsync(ctx->d_vram, 0, static_cast<size_t>(n_instances) * VRAM_PIXELS * 2, ctx->stream)) != cudaSuccess)
    return bail(e, "VRAM clear");
  if ((e = cudaMemsetAsync(ctx->d_clut, 0, static_cast<size_t>(n_instances) * ctx->clut_slots * CLUT_ENTRIES * 2, ctx->stream)) !=
      cudaSuccess)
    return bail(e, "CLUT clear");
  for (cudaEvent_t* ev : {&ctx->ev_start, &ctx->ev_end, &ctx->ev_hash0, &ctx->ev_hash1})
    if ((e = cudaEventCreate(ev)) != cudaSuccess)
      return bail(e, "event");
  if ((e = ctx->h_misc.ensure(std::max<size_t>(static_cast<size_t>(n_instances) * 8, 4096))) != cudaSuccess)
    return bail(e, "pinned hash buffer");
  if ((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess)
    return bail(e, "sync");

  ctx->persist_max = static_cast<uint32_t>(std::max(0, persistent_max_instances(device)));
  if (ctx->opts.reserved[0] & 1u) // latency mode disabled (opts.reserved[0] bit 0 = no_persistent)
    ctx->persist_max = 0;
  ctx->enc.reserve(n_instances);
  for (uint32_t i = 0; i < n_instances; i++)
  {
    ctx->enc.emplace_back(ctx->clut_slots, ctx->opts.modulation_crop != 0);
    ctx->enc.back().begin_flush();
  }
  ctx->dirty.assign(n_instances, 0);
  ctx->enc_blob.resize(n_instances);
  ctx->blob_truth.assign(n_instances, 0);
  *out = ctx;
  return PSXGPU_OK;
}

void psxgpu_destroy(psxgpu_ctx* ctx)
{
  if (!ctx)
    return;
  cudaSetDevice(ctx->device);
  for (cudaStream_t st : {ctx->copy_stream, ctx->walk_stream, ctx->enc_streams[0], ctx->enc_streams[1], ctx->enc_streams[2], ctx->enc_streams[3], ctx->stream})
    if (st)
      cudaStreamSynchronize(st);
  ctx->h_misc.release();
  ctx->d_scan.release();
  for (psxgpu_ctx::PinnedSet& ps : ctx->pinned)
  {
    for (PinnedBuf* b : {&ps.cmds, &ps.payload, &ps.clutops, &ps.items, &ps.misc, &ps.wire, &ps.recoff, &ps.inst, &ps.state, &ps.result})
      b->release();
    if (ps.free_ev)
      cudaEventDestroy(ps.free_ev);
  }
  for (psxgpu_ctx::Chunk* ch : ctx->chunks)
  {
    for (DevBuf* b : {&ch->d_cmds, &ch->d_setups, &ch->d_bins, &ch->d_tilecmds, &ch->d_payload, &ch->d_clutops, &ch->d_items, &ch->d_misc, &ch->d_recoff, &ch->d_rec,
                      &ch->d_inst, &ch->d_state, &ch->d_state_out, &ch->d_epochs, &ch->d_result, &ch->d_waves, &ch->d_cursor, &ch->d_summary, &ch->d_rects,
                      &ch->d_instwork})
      b->release();
    for (cudaEvent_t ev : {ch->ev_copied, ch->ev_walked, ch->ev_encoded, ch->ev_done})
      if (ev)
        cudaEventDestroy(ev);
    for (cudaEvent_t ev : ch->ev_raster)
      cudaEventDestroy(ev);
    delete ch;
  }
  if (ctx->d_vram)
    cudaFree(ctx->d_vram);
  if (ctx->d_clut)
    cudaFree(ctx->d_clut);
  if (ctx->d_hash)
    cudaFree(ctx->d_hash);
  if (ctx->d_barrier)
    cudaFree(ctx->d_barrier);
  for (cudaEvent_t ev : {ctx->ev_start, ctx->ev_end, ctx->ev_hash0, ctx->ev_hash1})
    if (ev)
      cudaEventDestroy(ev);
  for (cudaStream_t st : {ctx->copy_stream, ctx->walk_stream, ctx->enc_streams[0], ctx->enc_streams[1], ctx->enc_streams[2], ctx->enc_streams[3], ctx->stream})
    if (st)
      cudaStreamDestroy(st);
  delete ctx;
}

const char* psxgpu_last_error(const psxgpu_ctx* ctx)
{
  return ctx ? ctx->error.c_str() : "null context";
}

static int psxgpu_flush_impl(psxgpu_ctx* ctx)
{
  if (!ctx)
    return PSXGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  return do_flush(ctx);
}

// The C ABI never throws: an allocation failure inside becomes PSXGPU_E_NOMEM.
int psxgpu_flush(psxgpu_ctx* ctx)
{
  try
  {
    return psxgpu_flush_impl(ctx);
  }
  catch (const std::bad_alloc&)
  {
    if (ctx)
      ctx->error = "psxgpu_flush: out of host memory";
    return PSXGPU_E_NOMEM;
  }
  catch (...)
  {
    if (ctx)
      ctx->error = "psxgpu_flush: unexpected exception";
    return PSXGPU_E_INVALID;
  }
}

int psxgpu_sync(psxgpu_ctx* ctx)
{
  if (!ctx)
    return PSXGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  CK(cudaStreamSynchronize(ctx->stream), "stream sync");
  if (ctx->trace && !ctx->trace_chunks.empty())
  {
    std::fprintf(stderr, "psxgpu trace (ms since first chunk): chunk  copy[begin end]  encode_end  waves[begin end]\n");
    uint32_t i = 0;
    for (psxgpu_ctx::TraceChunk& tc : ctx->trace_chunks)
    {
      float t[5] = {0, 0, 0, 0, 0};
      cudaEvent_t evs[5] = {tc.copy0, tc.copy1, tc.enc1, tc.run0, tc.run1};
      for (int k = 0; k < 5; k++)
      {
        cudaEventSynchronize(evs[k]);
        cudaEventElapsedTime(&t[k], ctx->trace_origin, evs[k]);
        cudaEventDestroy(evs[k]);
      }
      std::fprintf(stderr, "  %2u  copy %7.2f %7.2f  enc %7.2f  waves %7.2f %7.2f\n", i++, t[0], t[1], t[2], t[3], t[4]);
    }
    ctx->trace_chunks.clear();
    cudaEventDestroy(ctx->trace_origin);
    ctx->trace_origin = nullptr;
  }
  return collect_timing(ctx);
}

int psxgpu_run_staged(psxgpu_ctx* ctx)
{
  if (!ctx)
    return PSXGPU_E_INVALID;
  if (ctx->group_open || ctx->staged_n == 0)
    return ctx->invalid("run_staged: no flushed command buffers are resident");
  cudaSetDevice(ctx->device);
  int rc = collect_timing(ctx);
  if (rc != PSXGPU_OK)
    return rc;
  CK(cudaEventRecord(ctx->ev_start, ctx->stream), "event record");
  for (uint32_t c = 0; c < ctx->staged_n; c++)
  {
    if (ctx->chunks[c]->device_encoded)
    {
      // the resident inputs of a device-encoded flush are the raw wire streams: encode them again
      CK(launch_encode(ctx->chunks[c]->enc_args, ctx->stream), "encode_kernel");
      ctx->counters.kernel_launches += 3;
    }
    rc = run_chunk(ctx, *ctx->chunks[c]);
    if (rc != PSXGPU_OK)
      return rc;
  }
  CK(cudaEventRecord(ctx->ev_end, ctx->stream), "event record");
  ctx->timing_chunks = ctx->staged_n;
  ctx->timing_pending = true;
  return PSXGPU_OK;
}

int psxgpu_set_batch_chunk(psxgpu_ctx* ctx, uint32_t instances_per_chunk)
{
  if (!ctx)
    return PSXGPU_E_INVALID;
  ctx->batch_chunk = instances_per_chunk;
  return PSXGPU_OK;
}

static int psxgpu_submit_wire_impl(psxgpu_ctx* ctx, uint32_t instance, const void* wire, size_t bytes)
{
  if (!ctx || instance >= ctx->n || (!wire && bytes))
    return ctx ? ctx->invalid("submit_wire: bad instance or buffer") : PSXGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  ensure_host_state(ctx, instance);
  const uint8_t* p = static_cast<const uint8_t*>(wire);
  size_t off = 0;
  while (off < bytes)
  {
    size_t used = 0;
    const double t0 = now_ms();
    // A long stream is flushed every AUTO_FLUSH primitives (asynchronously: nothing waits for the device), so that the
    // device works on the first part while the host encodes the rest; PSXGPU_AUTO_FLUSH=0 turns that off.
    static const uint32_t auto_flush = []() {
      const char* e = std::getenv("PSXGPU_AUTO_FLUSH");
      return e ? static_cast<uint32_t>(std::atoi(e)) : 2048u;
    }();
    ctx->enc[instance].set_cmd_budget(auto_flush);
    const EncodeResult res = ctx->enc[instance].add_wire(p + off, bytes - off, &used);
    ctx->enc[instance].set_cmd_budget(0);
    ctx->encode_ms_acc += now_ms() - t0;
    if (used)
      ctx->dirty[instance] = 1;
    off += used;
    if (res == EncodeResult::Ok && auto_flush != 0 && ctx->enc[instance].n_cmds() >= auto_flush)
    {
      const int frc = flush_work(ctx, std::vector<uint32_t>{instance});
      if (frc != PSXGPU_OK)
        return frc;
    }
    if (res == EncodeResult::Malformed)
    {
      ctx->error = "submit_wire: malformed record stream";
      return PSXGPU_E_MALFORMED;
    }
    if (res != EncodeResult::Barrier)
      continue;
    const psx_cmd_header* h = reinterpret_cast<const psx_cmd_header*>(p + off);
    int rc = PSXGPU_OK;
    switch (h->type)
    {
      case PSX_CMD_CLEAR_VRAM: rc = psxgpu_clear_vram(ctx, instance); break;
      case PSX_CMD_LOAD_STATE:
      {
        if (h->size < sizeof(psx_load_state))
          return ctx->invalid("LoadState record too small");
        const psx_load_state* ls = reinterpret_cast<const psx_load_state*>(h);
        rc = psxgpu_load_state(ctx, instance, ls->vram_data, ls->clut_data);
      }
      break;
      case PSX_CMD_READ_VRAM:
      {
        if (h->size < sizeof(psx_read_vram))
          return ctx->invalid("ReadVRAM record too small");
        const psx_read_vram* rv = reinterpret_cast<const psx_read_vram*>(h);
        if (rv->width > VRAM_W || rv->height > VRAM_H)
          return ctx->invalid("ReadVRAM rectangle larger than VRAM");
        uint16_t* sh = psxgpu_shadow_vram(ctx, instance);
        // refresh the shadow rectangle in place (wrapping): row by row pieces handled by do_read_vram into a temp
        std::vector<uint16_t> tmp(static_cast<size_t>(rv->width) * rv->height);
        rc = do_read_vram(ctx, instance, rv->x, rv->y, rv->width, rv->height, tmp.data(), rv->width);
        if (rc == PSXGPU_OK)
          for (uint32_t r = 0; r < rv->height; r++)
            for (uint32_t c = 0; c < rv->width; c++)
              sh[((rv->y + r) & (VRAM_H - 1)) * VRAM_W + ((rv->x + c) & (VRAM_W - 1))] = tmp[static_cast<size_t>(r) * rv->width + c];
      }
      break;
      case PSX_CMD_UPDATE_DISPLAY:
      {
        if (h->size < sizeof(psx_update_display))
          return ctx->invalid("UpdateDisplay record too small");
        const psx_update_display* ud = reinterpret_cast<const psx_update_display*>(h);
        ScanoutParams sp;
        bool disabled;
        scanout_params(ctx, ud, sp, disabled);
        if (!disabled && (sp.width > 4096 || sp.height > 4096))
          return ctx->invalid("UpdateDisplay larger than 4096 x 4096");
        if (disabled)
        {
          ctx->disp_w = ctx->disp_h = 0;
          rc = do_flush(ctx);
          break;
        }
        const uint32_t bpp = (sp.is24 || sp.format == PSXGPU_FMT_RGBA8) ? 4 : 2;
        ctx->display.resize(static_cast<size_t>(sp.out_stride) * sp.height);
        uint32_t w = 0, hh = 0;
        rc = do_scanout(ctx, instance, ud, ctx->display.data(), sp.out_stride, &w, &hh, nullptr);
        ctx->disp_w = w;
        ctx->disp_h = hh;
        ctx->disp_stride = sp.out_stride;
        ctx->disp_rgba8 = (bpp == 4);
      }
      break;
      default: break;
    }
    if (rc < 0)
      return rc;
    off += h->size;
  }
  return PSXGPU_OK;
}

// The C ABI never throws: an allocation failure inside becomes PSXGPU_E_NOMEM.
int psxgpu_submit_wire(psxgpu_ctx* ctx, uint32_t instance, const void* wire, size_t bytes)
{
  try
  {
    return psxgpu_submit_wire_impl(ctx, instance, wire, bytes);
  }
  catch (const std::bad_alloc&)
  {
    if (ctx)
      ctx->error = "psxgpu_submit_wire: out of host memory";
    return PSXGPU_E_NOMEM;
  }
  catch (...)
  {
    if (ctx)
      ctx->error = "psxgpu_submit_wire: unexpected exception";
    return PSXGPU_E_INVALID;
  }
}

static int psxgpu_submit_batch_impl(psxgpu_ctx* ctx, uint32_t first, uint32_t count, const void* const* streams, const size_t* bytes)
{
  if (!ctx || !streams || !bytes || first + count > ctx->n || first + count < first)
    return ctx ? ctx->invalid("submit_batch: bad arguments") : PSXGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  // records queued one at a time for these instances go first (they are older)
  {
    std::vector<uint32_t> pending;
    for (uint32_t i = 0; i < count; i++)
      if (ctx->dirty[first + i])
        pending.push_back(first + i);
    const int rc = flush_work(ctx, pending);
    if (rc != PSXGPU_OK)
      return rc;
  }
  if (ctx->encode_mode == 2 || (ctx->encode_mode == 0 && count >= 16))
    return submit_batch_device(ctx, first, count, streams, bytes);
  for (uint32_t i = 0; i < count; i++)
    ensure_host_state(ctx, first + i);
  // Host-encoded path (small batches: the latency mode).  Cut the batch into chunks: while the copy engine and the SMs
  // work on chunk k, the host threads encode chunk k+1 — straight into pinned memory, so nothing is copied twice.
  uint32_t per = ctx->batch_chunk;
  if (per == 0)
  {
    static const uint32_t div = []() {
      const char* e = std::getenv("PSXGPU_CHUNK_DIV"); // experiments: chunks per batch in automatic mode
      const int v = e ? std::atoi(e) : 0;
      return v > 0 ? static_cast<uint32_t>(v) : 8u;
    }();
    per = (count <= 256) ? count : std::max<uint32_t>(128, (count + div - 1) / div);
  }
  std::vector<int> status(count, 0);
  std::vector<uint32_t> work;
  std::vector<uint64_t> cmd_base, pay_base;
  for (uint32_t c0 = 0; c0 < count; c0 += per)
  {
    const uint32_t cn = std::min(per, count - c0);
    int rc = open_group(ctx);
    if (rc != PSXGPU_OK)
      return rc;
    // Every worker thread appends its instances back to back into its own arena of the pinned set, so there are no
    // gaps to fill and no sizing pass over the streams: an arena is sized from the stream bytes alone (a record that
    // yields a primitive is at least 24 bytes long; an upload pixel is 2 bytes).
    psxgpu_ctx::PinnedSet* P = nullptr;
    rc = acquire_pinned(ctx, P);
    if (rc != PSXGPU_OK)
      return rc;
    const double t0 = now_ms();
    const uint32_t T = std::max(1u, std::min(ctx->threads, cn));
    std::vector<uint64_t> arena_c(T + 1, 0), arena_p(T + 1, 0);
    for (uint32_t t = 0; t < T; t++)
    {
      const uint32_t lo = static_cast<uint32_t>((static_cast<uint64_t>(cn) * t) / T);
      const uint32_t hi = static_cast<uint32_t>((static_cast<uint64_t>(cn) * (t + 1)) / T);
      uint64_t c = 0, p = 0;
      for (uint32_t k = lo; k < hi; k++)
      {
        c += bytes[c0 + k] / 24 + 2;
        p += ((bytes[c0 + k] / 2 + 8) + 7) & ~static_cast<uint64_t>(7);
      }
      arena_c[t + 1] = arena_c[t] + c;
      arena_p[t + 1] = arena_p[t] + p;
    }
    CK(P->cmds.ensure(arena_c[T] * sizeof(PackedCmd)), "pinned alloc (commands)");
    CK(P->payload.ensure(arena_p[T] * 2), "pinned alloc (payload)");
    PackedCmd* hc = static_cast<PackedCmd*>(P->cmds.p);
    uint16_t* hp = static_cast<uint16_t*>(P->payload.p);
    HostLayout layout;
    layout.host_cmd_base.assign(cn, 0);
    std::vector<uint64_t> inst_p(cn, 0), used_c(T, 0), used_p(T, 0);
    // the encoders' persistent state (drawing area, current CLUT slots, snapshot table) as it is before this chunk: a
    // chunk that is dropped because one of its streams is malformed must not leave slots marked valid whose snapshot
    // ops were never run
    std::vector<EncoderStateBlob> saved(cn);
    for (uint32_t k = 0; k < cn; k++)
      ctx->enc[first + c0 + k].export_state(saved[k]);
    parallel_ranges(cn, T, [&](uint32_t t, uint32_t lo, uint32_t hi) {
      uint64_t cc = arena_c[t], cp = arena_p[t];
      for (uint32_t k = lo; k < hi; k++)
      {
        Encoder& e = ctx->enc[first + c0 + k];
        e.set_output(hc + cc, static_cast<uint32_t>(std::min<uint64_t>(arena_c[t + 1] - cc, 0xFFFFFFFFu)), hp + cp, arena_p[t + 1] - cp);
        size_t used = 0;
        const EncodeResult res = e.add_wire(streams[c0 + k], bytes[c0 + k], &used);
        e.finish_flush();
        status[c0 + k] = (res == EncodeResult::Ok && !e.overflowed()) ? 0 : (res == EncodeResult::Barrier ? PSXGPU_E_UNSUPPORTED : PSXGPU_E_MALFORMED);
        layout.host_cmd_base[k] = cc;
        inst_p[k] = cp;
        cc += e.n_cmds();
        cp += e.n_payload();
      }
      used_c[t] = cc - arena_c[t];
      used_p[t] = cp - arena_p[t];
    });
    ctx->encode_ms_acc += now_ms() - t0;
    work.clear();
    for (uint32_t k = 0; k < cn; k++)
    {
      if (status[c0 + k] != 0)
      {
        ctx->error = (status[c0 + k] == PSXGPU_E_MALFORMED) ? "submit_batch: malformed record stream"
                                                            : "submit_batch: read-back / state records are not allowed in a batch";
        // Drop the half-encoded chunk and put its encoders back where they were.  Earlier chunks of the same batch have
        // already been launched: a failing batch is applied up to the chunk that holds the bad stream (documented in
        // include/psxgpu.h).
        for (uint32_t j = 0; j < cn; j++)
        {
          ctx->enc[first + c0 + j].begin_flush();
          ctx->enc[first + c0 + j].import_state(saved[j]);
        }
        return status[c0 + k];
      }
      work.push_back(first + c0 + k);
    }
    // device layout: the arenas' used prefixes, concatenated
    cmd_base.assign(cn + 1, 0);
    pay_base.assign(cn + 1, 0);
    uint64_t dev_c = 0, dev_p = 0;
    for (uint32_t t = 0; t < T; t++)
    {
      const uint32_t lo = static_cast<uint32_t>((static_cast<uint64_t>(cn) * t) / T);
      const uint32_t hi = static_cast<uint32_t>((static_cast<uint64_t>(cn) * (t + 1)) / T);
      for (uint32_t k = lo; k < hi; k++)
      {
        cmd_base[k] = dev_c + (layout.host_cmd_base[k] - arena_c[t]);
        pay_base[k] = dev_p + (inst_p[k] - arena_p[t]);
      }
      layout.cmd_segs.push_back(CopySeg{arena_c[t], dev_c, used_c[t]});
      layout.pay_segs.push_back(CopySeg{arena_p[t], dev_p, used_p[t]});
      dev_c += used_c[t];
      dev_p += used_p[t];
    }
    cmd_base[cn] = dev_c;
    pay_base[cn] = dev_p;
    rc = launch_chunk(ctx, work, cmd_base, pay_base, *P, &layout);
    if (rc != PSXGPU_OK)
      return rc;
  }
  return PSXGPU_OK;
}

// The C ABI never throws: an allocation failure inside becomes PSXGPU_E_NOMEM.
int psxgpu_submit_batch(psxgpu_ctx* ctx, uint32_t first, uint32_t count, const void* const* streams, const size_t* bytes)
{
  try
  {
    return psxgpu_submit_batch_impl(ctx, first, count, streams, bytes);
  }
  catch (const std::bad_alloc&)
  {
    if (ctx)
      ctx->error = "psxgpu_submit_batch: out of host memory";
    return PSXGPU_E_NOMEM;
  }
  catch (...)
  {
    if (ctx)
      ctx->error = "psxgpu_submit_batch: unexpected exception";
    return PSXGPU_E_INVALID;
  }
}

int psxgpu_read_vram(psxgpu_ctx* ctx, uint32_t instance, uint32_t x, uint32_t y, uint32_t w, uint32_t h, uint16_t* dst,
                     uint32_t dst_pitch_px)
{
  if (!ctx || instance >= ctx->n || !dst)
    return ctx ? ctx->invalid("read_vram: bad arguments") : PSXGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  return do_read_vram(ctx, instance, x, y, w, h, dst, dst_pitch_px);
}

uint16_t* psxgpu_shadow_vram(psxgpu_ctx* ctx, uint32_t instance)
{
  if (!ctx || instance >= ctx->n)
    return nullptr;
  auto& v = ctx->shadow[instance];
  if (v.empty())
    v.assign(VRAM_PIXELS, 0);
  return v.data();
}

int psxgpu_clear_vram(psxgpu_ctx* ctx, uint32_t instance)
{
  if (!ctx || instance >= ctx->n)
    return ctx ? ctx->invalid("clear_vram: bad instance") : PSXGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  int rc = do_flush(ctx);
  if (rc != PSXGPU_OK)
    return rc;
  CK(cudaMemsetAsync(inst_vram(ctx, instance), 0, VRAM_PIXELS * 2, ctx->stream), "VRAM clear");
  CK(cudaMemsetAsync(inst_clut(ctx, instance, 0), 0, CLUT_ENTRIES * 2, ctx->stream), "CLUT clear");
  ensure_host_state(ctx, instance);
  ctx->enc[instance].reset_clut_state();
  return PSXGPU_OK;
}

int psxgpu_load_state(psxgpu_ctx* ctx, uint32_t instance, const uint16_t* vram, const uint16_t* clut)
{
  if (!ctx || instance >= ctx->n || !vram)
    return ctx ? ctx->invalid("load_state: bad arguments") : PSXGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  int rc = do_flush(ctx);
  if (rc != PSXGPU_OK)
    return rc;
  CK(cudaMemcpyAsync(inst_vram(ctx, instance), vram, VRAM_PIXELS * 2, cudaMemcpyHostToDevice, ctx->stream), "H2D VRAM");
  if (clut)
    CK(cudaMemcpyAsync(inst_clut(ctx, instance, 0), clut, CLUT_ENTRIES * 2, cudaMemcpyHostToDevice, ctx->stream), "H2D CLUT");
  else
    CK(cudaMemsetAsync(inst_clut(ctx, instance, 0), 0, CLUT_ENTRIES * 2, ctx->stream), "CLUT clear");
  CK(cudaStreamSynchronize(ctx->stream), "stream sync"); // caller's buffers may be reused on return
  ctx->counters.h2d_bytes += VRAM_PIXELS * 2 + CLUT_ENTRIES * 2;
  ensure_host_state(ctx, instance);
  ctx->enc[instance].reset_clut_state();
  return PSXGPU_OK;
}

int psxgpu_save_state(psxgpu_ctx* ctx, uint32_t instance, uint16_t* vram, uint16_t* clut)
{
  if (!ctx || instance >= ctx->n || !vram)
    return ctx ? ctx->invalid("save_state: bad arguments") : PSXGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  int rc = do_flush(ctx);
  if (rc != PSXGPU_OK)
    return rc;
  CK(cudaMemcpyAsync(vram, inst_vram(ctx, instance), VRAM_PIXELS * 2, cudaMemcpyDeviceToHost, ctx->stream), "D2H VRAM");
  if (clut)
  {
    ensure_host_state(ctx, instance);
    const Encoder& e = ctx->enc[instance];
    CK(cudaMemcpyAsync(clut, inst_clut(ctx, instance, e.clut_lo_slot()), 16 * 2, cudaMemcpyDeviceToHost, ctx->stream), "D2H CLUT");
    CK(cudaMemcpyAsync(clut + 16, inst_clut(ctx, instance, e.clut_hi_slot()) + 16, 240 * 2, cudaMemcpyDeviceToHost, ctx->stream),
       "D2H CLUT");
  }
  CK(cudaStreamSynchronize(ctx->stream), "stream sync");
  ctx->counters.d2h_bytes += VRAM_PIXELS * 2 + CLUT_ENTRIES * 2;
  return PSXGPU_OK;
}

int psxgpu_scanout(psxgpu_ctx* ctx, uint32_t instance, const psx_update_display* cmd, void* dst, uint32_t dst_stride,
                   uint32_t* out_width, uint32_t* out_height)
{
  if (!ctx || instance >= ctx->n || !cmd || !dst)
    return ctx ? ctx->invalid("scanout: bad arguments") : PSXGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  return do_scanout(ctx, instance, cmd, dst, dst_stride, out_width, out_height, nullptr);
}

// ---- presenter: GPU_SW::UpdateDisplay with the VideoPresenter steps (gpu_sw.cpp:386-441) ----
struct psxgpu_presenter
{
  psxgpu_ctx* ctx = nullptr;
  uint32_t mode = PSXGPU_DEINT_DISABLED;
  bool chroma = false;
  // s_locals.display_texture / display_texture_rect: what is shown (points at one of the buffers below, or nothing)
  const uint32_t* shown = nullptr;
  uint32_t shown_w = 0, shown_h = 0;
  DevBuf scan, smooth, target;           // m_upload_texture, chroma_smoothing_texture, deinterlace_texture
  uint32_t target_w = 0, target_h = 0;
  DevBuf field_buf[4];                   // deinterlace_buffers
  uint32_t field_w[4] = {0, 0, 0, 0}, field_h[4] = {0, 0, 0, 0};
  bool field_live[4] = {false, false, false, false};
  uint32_t current = 0;                  // current_deinterlace_buffer
};

namespace {

PresentField field_of(const psxgpu_presenter* p, uint32_t i)
{
  if (!p->field_live[i])
    return PresentField{nullptr, 0, 0};
  return PresentField{static_cast<const uint32_t*>(p->field_buf[i].p), p->field_w[i], p->field_h[i]};
}

// DeinterlaceSetTargetSize (video_presenter.cpp:1373-1383): a resize with preserve keeps the overlapping part, a new
// texture starts out cleared.
int present_target_size(psxgpu_presenter* p, uint32_t w, uint32_t h, bool preserve)
{
  psxgpu_ctx* ctx = p->ctx;
  if (p->target.p && p->target_w == w && p->target_h == h)
    return PSXGPU_OK;
  DevBuf fresh;
  CK(fresh.ensure(static_cast<size_t>(w) * h * 4), "device alloc (deinterlace target)");
  CK(cudaMemsetAsync(fresh.p, 0, static_cast<size_t>(w) * h * 4, ctx->stream), "clear deinterlace target");
  if (preserve && p->target.p && p->target_w && p->target_h)
    CK(cudaMemcpy2DAsync(fresh.p, static_cast<size_t>(w) * 4, p->target.p, static_cast<size_t>(p->target_w) * 4,
                         static_cast<size_t>(std::min(w, p->target_w)) * 4, std::min(h, p->target_h), cudaMemcpyDeviceToDevice,
                         ctx->stream),
       "preserve deinterlace target");
  CK(cudaStreamSynchronize(ctx->stream), "stream sync");
  p->target.release();
  p->target = fresh;
  p->target_w = w;
  p->target_h = h;
  return PSXGPU_OK;
}

// VideoPresenter::Deinterlace (video_presenter.cpp:1208-1371).  p->shown is the field just scanned out, or nullptr
// after ClearDisplayTexture.
int present_deinterlace(psxgpu_presenter* p, uint32_t field)
{
  psxgpu_ctx* ctx = p->ctx;
  const uint32_t* src = p->shown;
  const uint32_t w = p->shown_w, h = p->shown_h;
  auto copy_to_field_buffer = [&](uint32_t b) -> int {
    CK(p->field_buf[b].ensure(static_cast<size_t>(w) * h * 4), "device alloc (field buffer)");
    CK(cudaMemcpyAsync(p->field_buf[b].p, src, static_cast<size_t>(w) * h * 4, cudaMemcpyDeviceToDevice, ctx->stream), "field copy");
    p->field_w[b] = w;
    p->field_h[b] = h;
    p->field_live[b] = true;
    return PSXGPU_OK;
  };
  switch (p->mode)
  {
    case PSXGPU_DEINT_WEAVE:
    {
      if (src)
      {
        int rc = present_target_size(p, w, h * 2, true);
        if (rc != PSXGPU_OK)
          return rc;
      }
      else if (!p->target.p)
        return PSXGPU_OK;
      CK(launch_weave(PresentField{src, w, h}, field, static_cast<uint32_t*>(p->target.p), p->target_w, p->target_h, ctx->stream), "weave_kernel");
      ctx->counters.kernel_launches++;
    }
    break;
    case PSXGPU_DEINT_BLEND:
    {
      const uint32_t cur = p->current;
      p->current = (p->current + 1u) % 2u;
      if (src)
      {
        int rc = present_target_size(p, w, h, false);
        if (rc == PSXGPU_OK)
          rc = copy_to_field_buffer(cur);
        if (rc != PSXGPU_OK)
          return rc;
      }
      else
      {
        if (!p->target.p)
          return PSXGPU_OK;
        p->field_live[cur] = false; // recycled: samples black
      }
      CK(launch_blend(field_of(p, cur), field_of(p, (cur - 1u) % 2u), static_cast<uint32_t*>(p->target.p), p->target_w, p->target_h, ctx->stream),
         "blend_kernel");
      ctx->counters.kernel_launches++;
    }
    break;
    case PSXGPU_DEINT_ADAPTIVE:
    {
      const uint32_t cur = p->current;
      p->current = (p->current + 1u) % 4u;
      if (src)
      {
        int rc = present_target_size(p, w, h * 2, false);
        if (rc == PSXGPU_OK)
          rc = copy_to_field_buffer(cur);
        if (rc != PSXGPU_OK)
          return rc;
      }
      else
      {
        if (!p->target.p)
          return PSXGPU_OK;
        p->field_live[cur] = false;
      }
      const PresentField f[4] = {field_of(p, cur), field_of(p, (cur - 1u) % 4u), field_of(p, (cur - 2u) % 4u), field_of(p, (cur - 3u) % 4u)};
      CK(launch_adaptive(f, field, static_cast<uint32_t*>(p->target.p), p->target_w, p->target_h, ctx->stream), "adaptive_kernel");
      ctx->counters.kernel_launches++;
    }
    break;
    default: return PSXGPU_OK; // Disabled (and Progressive, which never produces interlaced frames): show the field
  }
  p->shown = static_cast<const uint32_t*>(p->target.p);
  p->shown_w = p->target_w;
  p->shown_h = p->target_h;
  return PSXGPU_OK;
}

// VideoPresenter::ApplyChromaSmoothing (video_presenter.cpp:1385-1413)
int present_chroma(psxgpu_presenter* p)
{
  psxgpu_ctx* ctx = p->ctx;
  if (!p->shown || p->shown_w == 0 || p->shown_h == 0)
    return PSXGPU_OK;
  CK(p->smooth.ensure(static_cast<size_t>(p->shown_w) * p->shown_h * 4), "device alloc (chroma smoothing)");
  CK(launch_chroma_smoothing(PresentField{p->shown, p->shown_w, p->shown_h}, static_cast<uint32_t*>(p->smooth.p), ctx->stream), "chroma_kernel");
  ctx->counters.kernel_launches++;
  p->shown = static_cast<const uint32_t*>(p->smooth.p);
  return PSXGPU_OK;
}

} // namespace

int psxgpu_presenter_create(psxgpu_ctx* ctx, uint32_t deinterlace_mode, uint32_t chroma_smoothing, psxgpu_presenter** out)
{
  if (!ctx || !out || deinterlace_mode > PSXGPU_DEINT_PROGRESSIVE)
    return ctx ? ctx->invalid("presenter_create: bad arguments") : PSXGPU_E_INVALID;
  psxgpu_presenter* p = new (std::nothrow) psxgpu_presenter();
  if (!p)
    return PSXGPU_E_NOMEM;
  p->ctx = ctx;
  p->mode = deinterlace_mode;
  p->chroma = chroma_smoothing != 0;
  *out = p;
  return PSXGPU_OK;
}

void psxgpu_presenter_destroy(psxgpu_presenter* p)
{
  if (!p)
    return;
  cudaSetDevice(p->ctx->device);
  cudaStreamSynchronize(p->ctx->stream);
  p->scan.release();
  p->smooth.release();
  p->target.release();
  for (DevBuf& b : p->field_buf)
    b.release();
  delete p;
}

int psxgpu_present(psxgpu_presenter* p, uint32_t instance, const psx_update_display* cmd, void* dst, size_t dst_bytes,
                   uint32_t* out_width, uint32_t* out_height)
{
  if (!p || !cmd || instance >= p->ctx->n)
    return p ? p->ctx->invalid("present: bad arguments") : PSXGPU_E_INVALID;
  psxgpu_ctx* ctx = p->ctx;
  cudaSetDevice(ctx->device);
  ScanoutParams sp;
  bool disabled;
  scanout_params(ctx, cmd, sp, disabled);
  const uint32_t field = (cmd->flags & PSX_UD_INTERLACED_FIELD) ? 1u : 0u;
  const bool interlaced = (cmd->flags & PSX_UD_INTERLACED_ENABLED) != 0 && !ctx->opts.show_vram;
  int rc = do_flush(ctx);
  if (rc != PSXGPU_OK)
    return rc;
  if (disabled)
  {
    p->shown = nullptr; // ClearDisplayTexture
    p->shown_w = p->shown_h = 0;
    if (interlaced)
      rc = present_deinterlace(p, field);
  }
  else
  {
    if (sp.width > 4096 || sp.height > 4096)
      return ctx->invalid("scan-out size out of range");
    sp.format = PSXGPU_FMT_RGBA8; // the presenter's shaders are defined on the RGBA8 display texture
    sp.out_stride = sp.width * 4;
    p->shown = nullptr;
    p->shown_w = p->shown_h = 0;
    if (sp.width && sp.height)
    {
      CK(p->scan.ensure(static_cast<size_t>(sp.out_stride) * sp.height), "device alloc (scan-out)");
      CK(launch_scanout(inst_vram(ctx, instance), sp, static_cast<uint8_t*>(p->scan.p), ctx->stream), "scanout_kernel");
      ctx->counters.kernel_launches++;
      p->shown = static_cast<const uint32_t*>(p->scan.p);
      p->shown_w = sp.width;
      p->shown_h = sp.height;
      if (sp.is24 && p->chroma && !ctx->opts.show_vram)
        rc = present_chroma(p);
      if (rc == PSXGPU_OK && interlaced)
        rc = present_deinterlace(p, field);
    }
  }
  if (rc != PSXGPU_OK)
    return rc;
  if (out_width)
    *out_width = p->shown_w;
  if (out_height)
    *out_height = p->shown_h;
  const size_t bytes = static_cast<size_t>(p->shown_w) * p->shown_h * 4;
  if (bytes && dst)
  {
    if (dst_bytes < bytes)
      return ctx->invalid("present: destination too small");
    CK(cudaMemcpyAsync(dst, p->shown, bytes, cudaMemcpyDeviceToHost, ctx->stream), "D2H presented frame");
    ctx->counters.d2h_bytes += bytes;
  }
  CK(cudaStreamSynchronize(ctx->stream), "stream sync");
  return p->shown ? PSXGPU_OK : 1;
}

const void* psxgpu_display(psxgpu_ctx* ctx, uint32_t* width, uint32_t* height, uint32_t* stride, uint32_t* is_rgba8)
{
  if (!ctx)
    return nullptr;
  if (width)
    *width = ctx->disp_w;
  if (height)
    *height = ctx->disp_h;
  if (stride)
    *stride = ctx->disp_stride;
  if (is_rgba8)
    *is_rgba8 = ctx->disp_rgba8;
  return ctx->display.data();
}

static int hash_common(psxgpu_ctx* ctx, uint32_t first, uint32_t count)
{
  if (first + count > ctx->n || first + count < first)
    return ctx->invalid("hash: bad instance range");
  cudaSetDevice(ctx->device);
  int rc = do_flush(ctx);
  if (rc != PSXGPU_OK)
    return rc;
  CK(cudaEventRecord(ctx->ev_hash0, ctx->stream), "event record");
  CK(launch_hash(ctx->d_vram, first, count, ctx->d_hash, ctx->stream), "hash_kernel");
  CK(cudaEventRecord(ctx->ev_hash1, ctx->stream), "event record");
  ctx->counters.kernel_launches += (count + 32767) / 32768;
  return PSXGPU_OK;
}

int psxgpu_hash(psxgpu_ctx* ctx, uint32_t first, uint32_t count, uint64_t* out_host)
{
  if (!ctx || !out_host)
    return ctx ? ctx->invalid("hash: bad arguments") : PSXGPU_E_INVALID;
  int rc = hash_common(ctx, first, count);
  if (rc != PSXGPU_OK)
    return rc;
  CK(cudaMemcpyAsync(ctx->h_misc.p, ctx->d_hash + first, static_cast<size_t>(count) * 8, cudaMemcpyDeviceToHost, ctx->stream),
     "D2H hashes");
  CK(cudaStreamSynchronize(ctx->stream), "stream sync");
  std::memcpy(out_host, ctx->h_misc.p, static_cast<size_t>(count) * 8);
  ctx->counters.d2h_bytes += static_cast<uint64_t>(count) * 8;
  cudaEventElapsedTime(&ctx->timing.last_hash_ms, ctx->ev_hash0, ctx->ev_hash1);
  return collect_timing(ctx);
}

int psxgpu_hash_device(psxgpu_ctx* ctx, uint32_t first, uint32_t count, void* out_device)
{
  if (!ctx || !out_device)
    return ctx ? ctx->invalid("hash_device: bad arguments") : PSXGPU_E_INVALID;
  int rc = hash_common(ctx, first, count);
  if (rc != PSXGPU_OK)
    return rc;
  CK(cudaMemcpyAsync(out_device, ctx->d_hash + first, static_cast<size_t>(count) * 8, cudaMemcpyDeviceToDevice, ctx->stream),
     "D2D hashes");
  CK(cudaStreamSynchronize(ctx->stream), "stream sync");
  cudaEventElapsedTime(&ctx->timing.last_hash_ms, ctx->ev_hash0, ctx->ev_hash1);
  return collect_timing(ctx);
}

int psxgpu_get_stats(psxgpu_ctx* ctx, psxgpu_stats* out)
{
  if (!ctx || !out)
    return PSXGPU_E_INVALID;
  psxgpu_stats s = ctx->counters;
  for (const Encoder& e : ctx->enc)
  {
    s.wire_commands += e.stats.wire_commands;
    s.packed_prims += e.stats.packed_prims;
    s.epochs += e.stats.epochs;
    s.cuts_raw += e.stats.cuts_raw;
    s.cuts_war += e.stats.cuts_war;
    s.cuts_clut += e.stats.cuts_clut;
    s.serial_self_sampling += e.stats.serial_self_sampling;
    s.serial_copy += e.stats.serial_copy;
    s.serial_upload += e.stats.serial_upload;
    s.clut_ops += e.stats.clut_ops;
    s.clut_cache_hits += e.stats.clut_cache_hits;
    s.rejected += e.stats.rejected;
  }
  *out = s;
  return PSXGPU_OK;
}

int psxgpu_reset_stats(psxgpu_ctx* ctx)
{
  if (!ctx)
    return PSXGPU_E_INVALID;
  ctx->counters = psxgpu_stats{};
  for (Encoder& e : ctx->enc)
    e.stats = EncoderStats{};
  return PSXGPU_OK;
}

int psxgpu_get_timing(psxgpu_ctx* ctx, psxgpu_timing* out)
{
  if (!ctx || !out)
    return PSXGPU_E_INVALID;
  cudaSetDevice(ctx->device);
  int rc = collect_timing(ctx);
  *out = ctx->timing;
  return rc;
}

int psxgpu_get_raster_launch_ms(psxgpu_ctx* ctx, float* out, uint32_t max_count)
{
  if (!ctx || !out)
    return PSXGPU_E_INVALID;
  collect_timing(ctx);
  const uint32_t n = std::min<uint32_t>(max_count, static_cast<uint32_t>(ctx->raster_ms.size()));
  for (uint32_t i = 0; i < n; i++)
    out[i] = ctx->raster_ms[i];
  return static_cast<int>(ctx->raster_ms.size());
}

} // extern "C"
