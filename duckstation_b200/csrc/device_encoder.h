// device_encoder.h — the batched path's command encoder, run on the GPU (encode_kernel.cu).
//
// psxgpu_submit_batch hands the host N independent wire streams.  Encoding them on the host (encoder.cpp) costs more
// than rasterising them on the device, so for batches the host only validates record sizes, copies the raw streams
// into pinned memory and uploads them; one warp per instance then does everything Encoder::handle_record does
// (src/core/gpu_backend.cpp:383-544 dispatch, drawing area, CLUT cache, quad split, culls) plus the hazard / epoch
// splitter, and small follow-up kernels turn the per-instance epoch lists into wave descriptors.
// Upload pixels are not copied again: OP_UPLOAD records point into the uploaded wire buffer.
//
// The encoder state that outlives a flush (drawing area, current CLUT slots, the CLUT snapshot table) travels as an
// EncState blob per instance, so the host encoder (psxgpu_submit_wire path) and the device encoder can be mixed.
#pragma once

#include "encoder.h"
#include "psx_types.h"

#include <cuda_runtime.h>

namespace psx {

constexpr uint32_t ENC_SLOT_VALID = 0x80000000u; // slot_key: x | y << 10 | is8 << 19 | VALID

// == EncoderStateBlob (encoder.h): raw drawing area l, t, r, b (clamped like Encoder::m_area), current CLUT slots,
// allocation hand, snapshot table.
using EncState = EncoderStateBlob;
static_assert(sizeof(EncState) == 1040, "EncState");

// One instance of a chunk: where its stream sits in the chunk's wire buffer and where its outputs go.
struct alignas(16) EncInst
{
  uint32_t instance;  // VRAM instance
  uint32_t wire_off;  // byte offset of the stream in the chunk's wire buffer (16-byte aligned)
  uint32_t wire_bytes;
  uint32_t cmd_base, cmd_cap;
  uint32_t clut_base, clut_cap;
  uint32_t epoch_base, epoch_cap;
  uint32_t rec_base;  // first entry of this stream in rec_off[] (capacity wire_bytes / 8 + 2: a record is >= 8 bytes)
  uint32_t pad[2];
};
static_assert(sizeof(EncInst) == 48, "EncInst");

// Record framing, checked on the device right after the upload (walk_kernel): what Encoder::add_wire checks, plus the
// exact output capacities the stream needs.  The host sizes the chunk's buffers from these.
struct WalkOut
{
  uint32_t n_rec, cmd_cap, clut_cap;
  int32_t status; // 0, PSXGPU_E_MALFORMED (-3) or PSXGPU_E_UNSUPPORTED (-5: a barrier record)
};

struct EncResult
{
  uint32_t n_cmds, n_clut, n_epochs, pad;
};

enum
{
  ENC_STAT_WIRE = 0,
  ENC_STAT_PRIMS,
  ENC_STAT_EPOCHS,
  ENC_STAT_CUTS_RAW,
  ENC_STAT_CUTS_WAR,
  ENC_STAT_CUTS_CLUT,
  ENC_STAT_SELF,
  ENC_STAT_COPY,
  ENC_STAT_UPLOAD,
  ENC_STAT_CLUT_OPS,
  ENC_STAT_CLUT_HITS,
  ENC_STAT_REJECTED,
  ENC_STAT_COUNT = 16
};

// D2H after a chunk has been encoded: everything the host needs to launch the chunk's waves.
struct EncSummary
{
  uint32_t max_epochs;  // number of waves
  uint32_t total_items; // sum of epochs
  uint32_t overflow;    // an output capacity was exceeded (cannot happen with capacities from the host walk)
  uint32_t pad;
  unsigned long long stats[ENC_STAT_COUNT];
};
// Per wave k (D2H, first ENC_WAVES_INLINE entries with the summary): item offset, count and what it contains.
struct EncWave
{
  uint32_t item_off, item_count;
  uint32_t flags; // 1 has CLUT ops, 2 has tiled epochs, 4 has serial epochs
  uint32_t tiled_cmds; // commands of the wave's tiled epochs (how heavy is a tile of this wave: the raster loop's grab size)
};
constexpr uint32_t ENC_WAVES_INLINE = 2048;

struct EncodeArgs
{
  const uint8_t* wire; // chunk wire buffer (instances back to back, 16-byte aligned, 256 bytes of slack at the end)
  const uint32_t* rec_off; // byte offset of every record in `wire`, written by walk_kernel (see EncInst::rec_base)
  const WalkOut* walk;     // per instance: number of records
  const EncInst* inst;
  uint32_t n_inst;
  uint32_t clut_slots;
  uint32_t modulation_crop;
  const EncState* state_in; // per chunk-local instance: encoder state at the start of the flush ...
  EncState* state_out;      // ... and after it (kept apart so that a flush can be replayed from its resident inputs)
  PackedCmd* cmds;
  uint4* rects; // per command slot: footprints of the open epoch (scratch of the encoder, see encode_kernel.cu)
  ClutOp* clut_ops;
  Epoch* epochs;
  EncResult* result;
  EncSummary* summary;
  EncWave* waves;       // [wave_cap]: item_count / flags accumulate here, item_off filled by the scan
  uint32_t* wave_cursor; // [wave_cap]
  uint32_t wave_cap;
  WaveItem* items;
  InstWork* inst_work; // [n_inst]: the same epochs, instance by instance (raster_instance_kernel)
};

// walk: reads inst[].wire_off / wire_bytes.  layout: exclusive prefix sums of the capacities into inst[].*_base / *_cap
// (the same sums the host makes from the read-back WalkOut to size the buffers).
cudaError_t launch_walk(const uint8_t* wire, const EncInst* inst, WalkOut* out, uint32_t* rec_off, uint32_t n_inst, cudaStream_t st);
cudaError_t launch_layout(const WalkOut* walk, EncInst* inst, uint32_t n_inst, cudaStream_t st);
cudaError_t launch_encode(const EncodeArgs& a, cudaStream_t st);

} // namespace psx
