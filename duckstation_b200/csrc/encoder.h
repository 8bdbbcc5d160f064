// encoder.h — GP0 command encoder + hazard/epoch splitter (host side, one per VRAM instance).
//
// Consumes the reference's backend command records (include/psxgpu_wire.h; reference
// src/core/video_thread_commands.h) exactly as GPUBackend::HandleCommand would see them
// (src/core/gpu_backend.cpp:383-544) and produces, per flush:
//   * PackedCmd[]  — one 64-byte device record per primitive, state resolved per record;
//   * payload[]    — CPU->VRAM upload pixels;
//   * Epoch[]      — cuts of the list such that, inside one epoch, no region of VRAM is both
//                    sampled (texture page, CLUT source, copy source) and written, so tile CTAs
//                    may run unordered relative to each other;
//   * ClutOp[]     — ordered CLUT snapshot operations (the reference's UpdateCLUT is a cache
//                    fill, src/core/gpu_sw_rasterizer.cpp:43-66), executed before their epoch.
#pragma once

#include "psx_types.h"

#include <cstddef>
#include <cstdint>
#include <vector>

namespace psx {

// (struct Epoch: psx_types.h)

struct EncoderStats
{
  uint64_t wire_commands = 0;
  uint64_t packed_prims = 0;
  uint64_t epochs = 0;
  uint64_t cuts_raw = 0;        // a read hit the epoch's write-set
  uint64_t cuts_war = 0;        // a write hit the epoch's read-set
  uint64_t cuts_clut = 0;       // a CLUT load hit the epoch's write-set / slot pressure
  uint64_t serial_self_sampling = 0;
  uint64_t serial_copy = 0;
  uint64_t serial_upload = 0;
  uint64_t clut_ops = 0;
  uint64_t clut_cache_hits = 0;
  uint64_t clut_invalidations = 0;
  uint64_t rejected = 0;        // records outside the decoder's invariants
  uint64_t upload_pixels = 0;
};

// 32 px x 8 row cells: 32 columns (one bit each) x 64 cell rows.
class HazardGrid
{
public:
  void clear();
  void mark(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1);       // inclusive, no wrap
  bool test(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) const; // inclusive, no wrap
  bool empty() const { return !m_any; }

private:
  uint32_t m_rows[64] = {};
  bool m_any = false;
};

struct RectSet
{
  struct R
  {
    uint16_t x0, y0, x1, y1; // inclusive
  };
  R r[4];
  int n = 0;
  void add_wrapped(uint32_t x, uint32_t y, uint32_t w, uint32_t h); // x,y may exceed the VRAM size (wrapped)
  bool intersects(const RectSet& o) const;
};

// A set of VRAM rectangles with a cheap negative test: a bounding box answers "cannot overlap" in four compares, and
// rectangles are only rasterised into the cell grid the first time the bounding box test is inconclusive.  In the common
// frame (draws into the framebuffer, textures elsewhere) the grid is never touched.
class HazardSet
{
public:
  void clear();
  void add(const RectSet& s);
  bool hits(const RectSet& s);
  bool empty() const { return !m_any; }

private:
  HazardGrid m_grid;
  std::vector<RectSet::R> m_pending;
  uint16_t m_x0 = 0, m_y0 = 0, m_x1 = 0, m_y1 = 0;
  bool m_any = false;
};

// Encoder state that outlives a flush, in the form the device encoder (device_encoder.h, EncState) keeps it.
struct EncoderStateBlob
{
  uint16_t area[4];
  uint8_t cur_lo, cur_hi;
  uint16_t alloc_hand;
  uint32_t pad;
  uint32_t slot_key[256]; // x | y << 10 | is8 << 19 | valid << 31
};

enum class EncodeResult
{
  Ok,        // whole buffer consumed
  Barrier,   // stopped in front of a record the encoder does not own (ReadVRAM, LoadState, UpdateDisplay, ClearVRAM, …)
  Malformed, // record sizes do not add up
};

class Encoder
{
public:
  explicit Encoder(uint32_t clut_slots = 256, bool modulation_crop = false);

  // Encodes records until the end or a barrier record.  *consumed = bytes eaten.
  EncodeResult add_wire(const void* wire, size_t bytes, size_t* consumed);

  // Closes the open epoch; after this cmds/payload/epochs/clut_ops describe the flush.
  void finish_flush();
  // Drops the flushed data, keeps GPU state (drawing area, live CLUT slots).
  void begin_flush();
  // VRAM + CLUT were replaced wholesale (ClearVRAM / LoadState): slot 0 now holds the CLUT.
  void reset_clut_state();

  bool has_work() const { return m_n_cmds != 0 || !clut_ops.empty(); }

  // Direct mode (batched path): the next flush's records and upload pixels are written straight into caller-provided
  // (pinned) memory instead of the vectors below.  Capacities come from scan_wire(); cleared by begin_flush().
  void set_output(PackedCmd* out_cmds, uint32_t cmd_capacity, uint16_t* out_payload, uint64_t payload_capacity);
  bool overflowed() const { return m_overflow; }
  uint32_t n_cmds() const { return m_n_cmds; }
  uint64_t n_payload() const { return m_n_payload; }
  // Upper bound of packed records and exact upper bound of (8-pixel padded) upload pixels a wire stream can produce.
  static bool scan_wire(const void* wire, size_t bytes, uint32_t* max_cmds, uint64_t* max_payload);

  std::vector<PackedCmd> cmds;
  std::vector<uint32_t> upload_cmds; // indices of OP_UPLOAD records (payload offsets get rebased at flush)
  std::vector<uint16_t> payload;
  std::vector<Epoch> epochs;
  std::vector<ClutOp> clut_ops;
  EncoderStats stats;

  void set_modulation_crop(bool on) { m_modulation_crop = on; }
  // add_wire returns (Ok, *consumed < bytes) once the open flush holds this many primitives; 0 = no limit.
  void set_cmd_budget(uint32_t n) { m_cmd_budget = n; }
  // Hand-over between this encoder and the device encoder of the batched path; only between flushes.
  void export_state(EncoderStateBlob& out) const;
  void import_state(const EncoderStateBlob& in);
  uint32_t clut_lo_slot() const { return m_cur_lo; }
  uint32_t clut_hi_slot() const { return m_cur_hi; }

private:
  void handle_record(const uint8_t* rec);
  void emit(const PackedCmd& c, const RectSet& writes, const RectSet& reads, uint32_t serial_kind);
  void close_epoch(uint32_t kind);
  void update_clut(uint16_t reg, bool is8);
  uint32_t alloc_slot();
  int find_slot(uint32_t x, uint32_t y, bool is8) const;
  void fill_draw_header(PackedCmd& c, const void* draw_header) const;
  void draw_triangle(const void* draw_header, const PackedVertex& a, const PackedVertex& b, const PackedVertex& c3);
  void texture_footprint(const PackedCmd& c, uint32_t umin, uint32_t umax, uint32_t vmin, uint32_t vmax, RectSet& reads) const;

  void push_cmd(const PackedCmd& c);

  uint32_t m_clut_slots;
  bool m_modulation_crop;
  PackedCmd* m_out_cmds = nullptr;
  uint16_t* m_out_payload = nullptr;
  uint32_t m_out_cmd_cap = 0;
  uint64_t m_out_pay_cap = 0;
  uint32_t m_n_cmds = 0;
  uint64_t m_n_payload = 0;
  bool m_overflow = false;
  uint32_t m_cmd_budget = 0;

  // GPU state mirrored from the command stream
  uint16_t m_area[4] = {0, 0, 0, 0}; // raw drawing area l, t, r, b

  // CLUT snapshot cache.  Slot 0 is the power-on / loaded-state CLUT and is never reallocated; slots 1.. hold
  // snapshots keyed by (x, y, depth) that stay valid until a command writes their source pixels.
  struct ClutSlot
  {
    uint16_t x = 0, y = 0;
    uint8_t is8 = 0;
    bool valid = false;
    uint32_t last_epoch = 0xFFFFFFFFu; // serial of the last epoch that references the slot's current contents
  };
  std::vector<ClutSlot> m_slots;
  // (x, y) -> chain of valid slots: bucket heads and per-slot links (0 = end; slot 0 is never cached)
  uint8_t m_bucket[1024] = {};
  std::vector<uint8_t> m_chain;
  static uint32_t bucket_of(uint32_t x, uint32_t y) { return ((x >> 4) ^ (y * 67u)) & 1023u; }
  void chain_insert(uint32_t slot);
  void chain_remove(uint32_t slot);
  uint32_t m_cur_lo = 0, m_cur_hi = 0; // slots supplying palette entries 0..15 / 16..255 right now
  uint32_t m_alloc_hand = 1;
  uint32_t m_epoch_serial = 0;
  HazardSet m_clut_sources; // superset of the valid slots' source pixels (prefilter for invalidation)
  void invalidate_cluts(const RectSet& writes);

  // open epoch
  uint32_t m_epoch_cmd_begin = 0;
  uint32_t m_epoch_clut_begin = 0;
  HazardSet m_writes, m_reads;
};

} // namespace psx
