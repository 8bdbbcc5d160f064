// psx_types.h — device command format shared by the host encoder and the CUDA kernels.
//
// PackedCmd is what the GP0 command encoder (encoder.cpp) writes into the pinned host
// command buffer and what travels over PCIe: one fixed 64-byte record per primitive
// (a quad becomes two triangle records, a polyline one record per segment), with every
// piece of state the reference keeps in globals (drawing area, CLUT cache, mask bits)
// resolved into the record so that records can be set up independently and in parallel.
//
// PrimSetup is the device-only result of the setup pre-pass (setup_kernel in kernels.cu):
// everything the raster kernel needs, with all divisions done.  The closed forms it encodes
// are derived in DESIGN.md §4 from the reference's DrawTriangle / DrawLine
// (src/core/gpu_sw_rasterizer.inl:1417-1566, :814-892).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define PSX_HD __host__ __device__ __forceinline__
#else
#define PSX_HD inline
#endif

namespace psx {

constexpr uint32_t VRAM_W = 1024, VRAM_H = 512, VRAM_PIXELS = VRAM_W * VRAM_H;
constexpr uint32_t TILE_W = 64, TILE_H = 32;         // one CTA owns one tile (4 KiB of u16)
constexpr uint32_t TILES_X = VRAM_W / TILE_W;        // 16
constexpr uint32_t TILES_Y = VRAM_H / TILE_H;        // 16
constexpr uint32_t NUM_TILES = TILES_X * TILES_Y;    // 256
// Raster kernel geometry: one warp (= one CTA) owns one 64x32 tile for the duration of an epoch.
constexpr uint32_t CLUT_ENTRIES = 256;

enum PackedOp : uint8_t
{
  OP_TRIANGLE = 0,
  OP_RECT = 1,
  OP_LINE = 2,
  OP_FILL = 3,
  OP_COPY = 4,
  OP_UPLOAD = 5,
  OP_NOP = 6, // rejected by setup (degenerate triangle, culled sprite, over-long line)
};

// flags0 = the wire flag byte (include/psxgpu_wire.h PSX_DF_*), verbatim.
enum : uint8_t
{
  F_INTERLACED = 0x01,
  F_ACTIVE_LSB = 0x02,
  F_SET_MASK = 0x04,
  F_CHECK_MASK = 0x08,
  F_TEXTURE = 0x10,
  F_RAW = 0x20,
  F_TRANSP = 0x40,
  F_SHADED = 0x80,
};
enum : uint8_t
{
  F1_DITHER = 0x01,
  F1_MOD5 = 0x02, // gpu_modulation_crop: Modulate5Bit instead of Modulate8Bit
};

// PrimSetup::tri.tflags
enum : uint8_t
{
  TRI_SHORT_IS_RIGHT = 0x01, // the short edges are the right-hand boundary (reference: right_facing)
  TRI_UPPER_UP = 0x02,       // rows [y0, y1) are walked upwards from v1 (reference: vo)
  TRI_LOWER_UP = 0x04,       // rows [y1, y2) are walked upwards from v2 (reference: vp != 0)
};

struct PackedVertex
{
  int16_t x, y;
  uint8_t r, g, b, u;
  uint8_t v, pad;
};
static_assert(sizeof(PackedVertex) == 10, "PackedVertex");

struct alignas(16) PackedCmd
{
  uint8_t op;
  uint8_t flags0;
  uint8_t flags1;
  uint8_t reserved;
  uint16_t draw_mode; // GPUDrawModeReg bits
  uint8_t clut_lo;    // CLUT snapshot slot supplying palette entries 0..15
  uint8_t clut_hi;    // slot supplying entries 16..255 (differs from clut_lo after a 4-bit load)
  uint8_t win_and_x, win_and_y, win_or_x, win_or_y;
  uint16_t clip_l, clip_t, clip_r, clip_b; // drawing area, inclusive, 0..1023
  union
  {
    PackedVertex v[3]; // triangle; line uses v[0], v[1]
    struct
    {
      int16_t x, y;
      uint16_t w, h;
      uint8_t r, g, b, u0;
      uint8_t v0;
    } rect;
    struct
    {
      uint16_t x, y, w, h;
      uint16_t color16;
      uint8_t interlaced, active_lsb;
    } fill;
    struct
    {
      uint16_t sx, sy, dx, dy, w, h;
    } copy;
    struct
    {
      uint16_t x, y, w, h;
      uint32_t payload; // offset into the payload buffer, in u16 units (global over the flush)
    } upload;
    uint8_t raw[44];
  };
};
static_assert(sizeof(PackedCmd) == 64, "PackedCmd must be 64 bytes");

// Epoch kinds (hazard/epoch splitter, encoder.cpp).
enum EpochKind : uint32_t
{
  EPOCH_TILED = 0,        // binned per tile, tiles run in parallel
  EPOCH_SELF_SAMPLING = 1, // one textured primitive whose texture footprint overlaps its own pixels
  EPOCH_COPY_OVERLAP = 2,  // one CopyVRAM whose source and destination alias
  EPOCH_UPLOAD_MASKED = 3, // one mask-checked CPU->VRAM upload
};

// One epoch of one instance, indices relative to the instance's own commands / CLUT ops.
struct Epoch
{
  uint32_t cmd_begin, cmd_end; // indices into cmds
  uint32_t kind;               // EpochKind
  uint32_t clut_begin, clut_end;
};

// One instance of a chunk for raster_instance_kernel: where its epochs, commands and CLUT ops start in the chunk's arrays.
struct InstWork
{
  uint32_t instance;
  uint32_t cmd_base, clut_base;
  uint32_t epoch_first, n_epochs;
  uint32_t pad[3];
};
static_assert(sizeof(InstWork) == 32, "InstWork");

struct WaveItem
{
  uint32_t instance;
  uint32_t cmd_begin, cmd_end; // global indices into the flush's command arrays
  uint32_t kind;
  uint32_t clut_begin, clut_end; // global indices into the flush's CLUT-op array
  uint32_t pad[2];
};
static_assert(sizeof(WaveItem) == 32, "WaveItem");

// CLUT snapshot: copy 16 (4-bit) or 256 (8-bit, x wraps at 1024) pixels of VRAM row y into a slot.
struct ClutOp
{
  uint16_t dst_slot;
  uint16_t x, y; // VRAM source
  uint16_t is8bit;
};
static_assert(sizeof(ClutOp) == 8, "ClutOp");

// What the raster kernel's word loop needs from a triangle / rectangle, every derived constant ready to use (the loop
// reads them from the staged record with LDS instead of holding them in registers; psx_setup.h:make_loop_desc).
enum : uint32_t
{
  LP_TEXTURED = 0x001,
  LP_PALETTED = 0x002,  // 4- or 8-bit texture: the texel is a palette index
  LP_MODULATED = 0x004, // the colours matter: untextured, or textured and not raw
  LP_GOURAUD = 0x008,   // per-pixel colours (shaded triangle whose colours matter)
  LP_DITHER = 0x010,
  LP_TRANSP = 0x020,
  LP_TMODE_SHIFT = 6,   // bits 6-7: semi-transparency mode
  LP_TRI = 0x100,
  LP_WINDOW = 0x200,    // a texture window other than and = 0xFF, or = 0
  LP_INTERLACED = 0x400, // interlaced rendering: rows of the displayed field are skipped
  LP_COMPLEX = 0x800,   // triangle with a coordinate outside 11 bits: rows / columns wrap, the general tri_row decides
};
struct alignas(16) LoopDesc
{
  uint32_t and_x2, or_x2, and_y2, or_y2; // texture window, each byte value in both halves
  uint32_t tex_bx2;   // texture page base column in both halves
  uint32_t sub_mask2; // (texels per VRAM word - 1) in both halves
  uint32_t tex_rows;  // texture page base row * 1024
  uint32_t tex_shift; // log2(texels per VRAM word)
  uint32_t sub_shift; // log2(bits per texel)
  uint32_t idx_mask;  // palette index mask (15 / 255)
  uint32_t chk15;     // 0x80008000 if the mask bit is tested
  uint32_t set15;     // 0x80008000 if the mask bit is set
  uint32_t fr, fg, fb; // flat colour: textured = the 8-bit values (Modulate5Bit crop applied), untextured = 16 c in both halves
  uint32_t cmask;     // 0xFF, or 0xF8 with Modulate5Bit: applied to interpolated colours
  uint32_t path;      // LP_*
  uint32_t ubase;     // rectangle: (u0 - x) & 0xFF
  uint32_t clut_key;  // paletted: lo slot | hi slot << 8 | 8-bit << 16 | 1 << 31 (which palette the loop wants staged); else 0
  uint32_t pad;
};
static_assert(sizeof(LoopDesc) == 80, "LoopDesc");

// One y-range of a triangle: both edges as x(cy) = start + (cy - lo) * step (mod 2^64, then >> 32), lo = first row.
struct TriPart
{
  uint64_t l_start, l_step, r_start, r_step;
};

// ---- device-side setup record: 256 bytes, read by the raster kernel ----
struct alignas(16) PrimSetup
{
  // 32-byte common header
  uint8_t op, flags0, flags1, texmode; // texmode: 0 4-bit, 1 8-bit, 2 direct
  uint16_t draw_mode;
  uint8_t clut_lo, clut_hi;
  uint8_t win_and_x, win_and_y, win_or_x, win_or_y;
  int16_t clip_l, clip_t, clip_r, clip_b;
  int16_t bx0, bx1;      // VRAM column range touched (inclusive), already clipped
  uint16_t by0, bycount; // VRAM rows touched: (by0 + i) & 511, i < bycount
  uint16_t tex_bx, tex_by; // texture page base

  union
  {
    struct
    {
      TriPart part[2];  // [0] rows cy < y1, [1] rows cy >= y1; left / right edge, anchored at the part's first row
      int16_t y0, y1;
      int16_t lo[2], hi[2]; // effective row range [lo, hi) of each part after break emulation (hi >= lo); unless
                            // LP_COMPLEX also cut to the drawing area's rows
      uint8_t tflags;       // TRI_SHORT_IS_RIGHT | TRI_UPPER_UP | TRI_LOWER_UP
      uint8_t flat_r, flat_g, flat_b;
      uint32_t base[5], ddx[5], ddy[5]; // r g b u v : value(x,cy) = base + ddx*x + ddy*cy, top byte
      uint32_t pad;
    } tri; // 64 + 16 + 60 + 4 = 144
    struct
    {
      int16_t x, y;
      uint16_t w, h;
      uint8_t r, g, b, u0, v0;
    } rect;
    struct
    {
      uint64_t curx0, cury0, dxdk, dydk;
      uint32_t r0, g0, b0;
      int32_t dr, dg, db;
      int32_t k;
      int16_t x0, y0; // p0 after the leftmost swap
      uint8_t x_major, flat_r, flat_g, flat_b;
    } line;
    struct
    {
      uint16_t x, y, w, h;
      uint16_t color16;
      uint8_t interlaced, active_lsb;
    } fill;
    struct
    {
      uint16_t sx, sy, dx, dy, w, h;
    } copy;
    struct
    {
      uint16_t x, y, w, h;
      uint32_t payload;
    } upload;
    uint32_t words[36];
  };
  LoopDesc desc; // triangles and rectangles
};
static_assert(sizeof(PrimSetup) == 256, "PrimSetup must be 256 bytes");

// Binning words produced by setup, read by every tile CTA of the epoch.
struct BinInfo
{
  uint32_t tilemask; // bits 0-15: 64-px tile columns touched, bits 16-31: 32-row tile rows touched
  uint32_t readmask; // the same form: tiles the command may read while it runs (texture footprint, copy source); 0 = none
};

PSX_HD int32_t sext11(int32_t v)
{
  return static_cast<int32_t>(static_cast<uint32_t>(v) << 21) >> 21;
}

PSX_HD int32_t imin(int32_t a, int32_t b)
{
  return a < b ? a : b;
}
PSX_HD int32_t imax(int32_t a, int32_t b)
{
  return a > b ? a : b;
}

PSX_HD uint64_t mix64(uint64_t z)
{
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}

} // namespace psx
