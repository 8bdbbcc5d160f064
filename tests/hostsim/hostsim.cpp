// hostsim.cpp — TEST-ONLY host build of the device headers (psx_setup.h / psx_pixel.h) and the encoder.
//
// Executes an encoded flush on the CPU with the same execution MODEL as the CUDA kernels, so the closed forms,
// the encoder and the hazard/epoch splitter can be checked against the oracle without a GPU:
//   * per epoch, texture / CLUT-source / copy-source reads come from a snapshot of VRAM taken at epoch start
//     (a tile CTA reads those from global memory while other tiles' writes are still in shared memory);
//   * CLUT snapshot ops of an epoch run before its primitives;
//   * inside an epoch primitives are applied in order, pixel by pixel, in arbitrary (here: raster) pixel order.
// It is not part of the product library and nothing in duckstation_b200/ links it.
#include "../../duckstation_b200/csrc/encoder.h"
#include "../../duckstation_b200/csrc/psx_pair.h"
#include "../../duckstation_b200/csrc/psx_pixel.h"
#include "../../duckstation_b200/csrc/psx_setup.h"
#include "../../include/psxgpu_wire.h"

#include <climits>
#include <cstring>
#include <vector>

using namespace psx;

namespace {
constexpr uint32_t SLOTS = 256;
std::vector<uint16_t> g_vram(VRAM_PIXELS), g_snap(VRAM_PIXELS);
std::vector<uint16_t> g_clut(SLOTS * 256);
Encoder* g_enc = nullptr;
EncoderStats g_stats;

void shade_and_store(const PrimSetup& s, const uint16_t*, uint32_t px, uint32_t Y, uint32_t r, uint32_t g, uint32_t b,
                     uint32_t u, uint32_t v)
{
  const ShadeState st = make_shade_state(s);
  uint32_t texel = 0;
  if (s.flags0 & F_TEXTURE)
    texel = fetch_texel<false>(st, g_snap.data(), &g_clut[s.clut_lo * 256], &g_clut[s.clut_hi * 256], u, v);
  uint32_t colour;
  if (!shade_colour(st, px, Y, r, g, b, texel, colour))
    return;
  uint16_t& dst = g_vram[Y * VRAM_W + px];
  dst = static_cast<uint16_t>(blend_mask(st, dst, colour));
}

// Work model of the raster kernel's word loop (profiling aid): per (primitive, tile) visit the 32-bit words its spans
// touch, the batches of 32 words the warp walks, and the pixels covered.
uint64_t g_work[8]; // visits, words, batches, pixels, visits by words bucket: <=32, <=64, <=128, >128
void account_span(std::vector<uint32_t>& words, std::vector<uint32_t>& touched, int32_t lo, int32_t hi, uint32_t Y)
{
  for (int32_t tx = lo >> 6; tx <= (hi - 1) >> 6; tx++)
  {
    const int32_t a = lo > tx * 64 ? lo : tx * 64, b = hi < tx * 64 + 64 ? hi : tx * 64 + 64;
    const uint32_t t = (Y >> 5) * 16u + (static_cast<uint32_t>(tx) & 15u);
    if (words[t] == 0)
      touched.push_back(t);
    words[t] += static_cast<uint32_t>(((b - 1) >> 1) - (a >> 1) + 1);
    g_work[3] += static_cast<uint32_t>(b - a);
  }
}
void account_flush(std::vector<uint32_t>& words, std::vector<uint32_t>& touched)
{
  for (uint32_t t : touched)
  {
    const uint32_t w = words[t];
    g_work[0]++;
    g_work[1] += w;
    g_work[2] += (w + 31) / 32;
    g_work[w <= 32 ? 4 : w <= 64 ? 5 : w <= 128 ? 6 : 7]++;
    words[t] = 0;
  }
  touched.clear();
}

void run_tiled_prim(const PrimSetup& s, const uint16_t* payload)
{
  const uint16_t* clut = nullptr;
  if (s.op == OP_NOP)
    return;
  static std::vector<uint32_t> acc_words(NUM_TILES, 0), acc_touched;
  for (uint32_t i = 0; i < s.bycount; i++)
  {
    const uint32_t Y = (s.by0 + i) & (VRAM_H - 1);
    switch (s.op)
    {
      case OP_TRIANGLE:
      {
        TriRow row;
        if (!tri_row(s, Y, row))
          break;
        if (row.x_hi > row.x_lo)
          account_span(acc_words, acc_touched, row.x_lo, row.x_hi, Y);
        for (int32_t px = row.x_lo; px < row.x_hi; px++)
        {
          uint32_t r, g, b, u, v;
          tri_attrs(s, row, px, r, g, b, u, v);
          shade_and_store(s, clut, static_cast<uint32_t>(px), Y, r, g, b, u, v);
        }
      }
      break;
      case OP_RECT:
      {
        RectRow row;
        if (!rect_row(s, Y, row))
          break;
        if (row.x_hi > row.x_lo)
          account_span(acc_words, acc_touched, row.x_lo, row.x_hi, Y);
        for (int32_t px = row.x_lo; px < row.x_hi; px++)
          shade_and_store(s, clut, static_cast<uint32_t>(px), Y, s.rect.r, s.rect.g, s.rect.b,
                          (s.rect.u0 + static_cast<uint32_t>(px - s.rect.x)) & 0xFFu, row.v);
      }
      break;
      case OP_LINE:
      {
        if (s.line.x_major)
        {
          for (int32_t px = s.bx0; px <= s.bx1; px++)
          {
            LinePixel p;
            if (line_step(s, line_step_for_column(s, px), p) && p.x == px && (static_cast<uint32_t>(p.y) & (VRAM_H - 1)) == Y)
              shade_and_store(s, clut, static_cast<uint32_t>(px), Y, p.r, p.g, p.b, 0, 0);
          }
        }
        else
        {
          LinePixel p;
          if (line_step(s, line_step_for_row(s, Y), p) && (static_cast<uint32_t>(p.y) & (VRAM_H - 1)) == Y)
            shade_and_store(s, clut, static_cast<uint32_t>(p.x), Y, p.r, p.g, p.b, 0, 0);
        }
      }
      break;
      case OP_FILL:
        for (uint32_t px = 0; px < VRAM_W; px++)
          if (fill_covers(s, px, Y))
            g_vram[Y * VRAM_W + px] = s.fill.color16;
        break;
      case OP_COPY:
        for (uint32_t px = 0; px < VRAM_W; px++)
        {
          uint32_t si;
          if (copy_covers(s, px, Y, si))
            g_vram[Y * VRAM_W + px] = transfer_write(s, g_vram[Y * VRAM_W + px], g_snap[si]);
        }
        break;
      case OP_UPLOAD:
        for (uint32_t px = 0; px < VRAM_W; px++)
        {
          uint32_t pi;
          if (upload_covers(s, px, Y, pi))
            g_vram[Y * VRAM_W + px] = transfer_write(s, g_vram[Y * VRAM_W + px], payload[pi]);
        }
        break;
      default: break;
    }
  }
  account_flush(acc_words, acc_touched);
}

// ---- serial kinds: reference-order emulation (what the single-CTA kernels do) ----
void shade_group(const PrimSetup& s, const uint16_t*, int32_t x0, uint32_t Y, uint32_t active, const uint32_t* r,
                 const uint32_t* g, const uint32_t* b, const uint32_t* u, const uint32_t* v)
{
  const ShadeState st = make_shade_state(s);
  uint32_t tex[4] = {0, 0, 0, 0};
  for (int l = 0; l < 4; l++)
    if ((active >> l) & 1u)
      tex[l] = fetch_texel<true>(st, g_vram.data(), &g_clut[s.clut_lo * 256], &g_clut[s.clut_hi * 256], u[l], v[l]); // LIVE vram
  for (int l = 0; l < 4; l++)
  {
    if (!((active >> l) & 1u))
      continue;
    uint32_t colour;
    if (!shade_colour(st, static_cast<uint32_t>(x0 + l), Y, r[l], g[l], b[l], tex[l], colour))
      continue;
    uint16_t& dst = g_vram[Y * VRAM_W + static_cast<uint32_t>(x0 + l)];
    dst = static_cast<uint16_t>(blend_mask(st, dst, colour));
  }
}

void run_self_sampling(const PrimSetup& s)
{
  const uint16_t* clut = nullptr;
  if (s.op == OP_RECT)
  {
    for (uint32_t oy = 0; oy < s.rect.h; oy++)
    {
      const uint32_t Y = static_cast<uint32_t>(s.rect.y + static_cast<int32_t>(oy)) & (VRAM_H - 1);
      RectRow row;
      if (!rect_row(s, Y, row) || row.y != s.rect.y + static_cast<int32_t>(oy))
        continue;
      for (uint32_t ox = 0; ox < s.rect.w; ox += 4)
      {
        uint32_t active = 0, r[4], g[4], b[4], u[4], v[4];
        for (uint32_t l = 0; l < 4; l++)
        {
          const int32_t x = s.rect.x + static_cast<int32_t>(ox + l);
          if (x >= row.x_lo && x < row.x_hi)
            active |= 1u << l;
          r[l] = s.rect.r;
          g[l] = s.rect.g;
          b[l] = s.rect.b;
          u[l] = (s.rect.u0 + ox + l) & 0xFFu;
          v[l] = row.v;
        }
        if (active)
          shade_group(s, clut, s.rect.x + static_cast<int32_t>(ox), Y, active, r, g, b, u, v);
      }
    }
    return;
  }
  if (s.op != OP_TRIANGLE)
    return;
  const auto& T = s.tri;
  // part order: triparts[0] then [1] (inl:1557-1560).  The upper part is triparts[vo]; vo == walks-upwards.
  struct Part
  {
    int32_t lo, hi;
    bool up;
  };
  const bool upper_is_up = (T.tflags & TRI_UPPER_UP) != 0, lower_is_up = (T.tflags & TRI_LOWER_UP) != 0;
  Part upper{T.lo[0], T.hi[0], upper_is_up}, lower{T.lo[1], T.hi[1], lower_is_up};
  Part order[2];
  if (upper_is_up)
  {
    order[0] = lower;
    order[1] = upper;
  }
  else
  {
    order[0] = upper;
    order[1] = lower;
  }
  for (const Part& p : order)
  {
    for (int32_t k = 0; k < p.hi - p.lo; k++)
    {
      const int32_t cy = p.up ? (p.hi - 1 - k) : (p.lo + k);
      const uint32_t Y = static_cast<uint32_t>(sext11(cy)) & (VRAM_H - 1);
      TriRow row;
      if (!tri_row(s, Y, row) || row.cy != cy)
        continue;
      for (int32_t gx = row.x_lo; gx < row.x_hi; gx += 4)
      {
        uint32_t active = 0, r[4], g[4], b[4], u[4], v[4];
        for (int32_t l = 0; l < 4; l++)
        {
          if (gx + l < row.x_hi)
            active |= 1u << l;
          tri_attrs(s, row, gx + l, r[l], g[l], b[l], u[l], v[l]);
        }
        shade_group(s, clut, gx, Y, active, r, g, b, u, v);
      }
    }
  }
}

void copy_serial(uint32_t sx, uint32_t sy, uint32_t dx, uint32_t dy, uint32_t w, uint32_t h, uint16_t mand, uint16_t mor)
{
  if ((sx + w) > VRAM_W || (dx + w) > VRAM_W)
  {
    uint32_t rem_rows = h, csy = sy, cdy = dy;
    while (rem_rows > 0)
    {
      const uint32_t rows = std::min(rem_rows, std::min(VRAM_H - csy, VRAM_H - cdy));
      uint32_t rem_cols = w, csx = sx, cdx = dx;
      while (rem_cols > 0)
      {
        const uint32_t cols = std::min(rem_cols, std::min(VRAM_W - csx, VRAM_W - cdx));
        copy_serial(csx, csy, cdx, cdy, cols, rows, mand, mor);
        csx = (csx + cols) % VRAM_W;
        cdx = (cdx + cols) % VRAM_W;
        rem_cols -= cols;
      }
      csy = (csy + rows) % VRAM_H;
      cdy = (cdy + rows) % VRAM_H;
      rem_rows -= rows;
    }
    return;
  }
  static uint16_t rowbuf[VRAM_W];
  for (uint32_t r = 0; r < h; r++)
  {
    const uint16_t* srow = &g_vram[((sy + r) % VRAM_H) * VRAM_W];
    uint16_t* drow = &g_vram[((dy + r) % VRAM_H) * VRAM_W];
    for (uint32_t c = 0; c < w; c++)
      rowbuf[c] = srow[(sx + c) % VRAM_W];
    for (uint32_t c = 0; c < w; c++)
    {
      uint16_t& d = drow[(dx + c) % VRAM_W];
      if ((d & mand) == 0)
        d = static_cast<uint16_t>(rowbuf[c] | mor);
    }
  }
}

void upload_masked(const PrimSetup& s, const uint16_t* payload)
{
  const auto& U = s.upload;
  const uint32_t aw = std::min<uint32_t>(U.w, VRAM_W - U.x) & ~7u;
  const uint16_t mor = (s.flags0 & F_SET_MASK) ? 0x8000u : 0;
  uint32_t deficit = 0;
  for (uint32_t r = 0; r < U.h; r++)
  {
    uint16_t* drow = &g_vram[((U.y + r) % VRAM_H) * VRAM_W];
    for (uint32_t c = 0; c < U.w; c++)
    {
      uint16_t& d = drow[(U.x + c) % VRAM_W];
      const bool masked = (d & 0x8000u) != 0;
      if (!masked)
        d = static_cast<uint16_t>(payload[U.payload + r * U.w + c - deficit] | mor);
      else if (c >= aw)
        deficit++;
    }
  }
}
} // namespace

extern "C" {

void hostsim_reset(int modulation_crop)
{
  delete g_enc;
  g_enc = new Encoder(SLOTS, modulation_crop != 0);
  std::fill(g_vram.begin(), g_vram.end(), 0);
  std::fill(g_clut.begin(), g_clut.end(), 0);
}

uint16_t* hostsim_vram()
{
  return g_vram.data();
}

// Encodes and executes one wire stream (barrier records are skipped).  Returns the number of epochs, or -1.
long hostsim_exec(const void* wire, size_t bytes)
{
  if (!g_enc)
    hostsim_reset(0);
  const uint8_t* p = static_cast<const uint8_t*>(wire);
  size_t off = 0;
  g_enc->begin_flush();
  while (off < bytes)
  {
    size_t used = 0;
    const EncodeResult res = g_enc->add_wire(p + off, bytes - off, &used);
    off += used;
    if (res == EncodeResult::Malformed)
      return -1;
    if (res == EncodeResult::Barrier)
      off += reinterpret_cast<const psx_cmd_header*>(p + off)->size;
  }
  g_enc->finish_flush();
  std::vector<PrimSetup> setups(g_enc->cmds.size());
  std::vector<BinInfo> bins(g_enc->cmds.size());
  for (size_t i = 0; i < g_enc->cmds.size(); i++)
  {
    std::memset(&setups[i], 0, sizeof(PrimSetup));
    setup_prim(g_enc->cmds[i], setups[i], bins[i]);
  }
  const uint16_t* payload = g_enc->payload.data();
  for (const Epoch& e : g_enc->epochs)
  {
    for (uint32_t o = e.clut_begin; o < e.clut_end; o++)
    {
      const ClutOp& op = g_enc->clut_ops[o];
      for (uint32_t i = 0; i < (op.is8bit ? 256u : 16u); i++)
        g_clut[op.dst_slot * 256 + i] = g_vram[op.y * VRAM_W + ((op.x + i) & (VRAM_W - 1))];
    }
    if (e.kind == EPOCH_TILED)
    {
      g_snap = g_vram;
      for (uint32_t i = e.cmd_begin; i < e.cmd_end; i++)
        run_tiled_prim(setups[i], payload);
    }
    else
    {
      const PrimSetup& s = setups[e.cmd_begin];
      if (s.op == OP_NOP)
        continue;
      if (e.kind == EPOCH_SELF_SAMPLING)
        run_self_sampling(s);
      else if (e.kind == EPOCH_COPY_OVERLAP)
        copy_serial(s.copy.sx, s.copy.sy, s.copy.dx, s.copy.dy, s.copy.w, s.copy.h, (s.flags0 & F_CHECK_MASK) ? 0x8000u : 0,
                    (s.flags0 & F_SET_MASK) ? 0x8000u : 0);
      else if (e.kind == EPOCH_UPLOAD_MASKED)
        upload_masked(s, payload);
    }
  }
  g_stats = g_enc->stats;
  return static_cast<long>(g_enc->epochs.size());
}

// Replaces the encoder by a fresh one that only knows the exported state blob — what happens when an instance moves
// between the host encoder and the device encoder of the batched path (Encoder::export_state / import_state).
void hostsim_handover(int modulation_crop)
{
  if (!g_enc)
    return;
  EncoderStateBlob blob;
  g_enc->export_state(blob);
  delete g_enc;
  g_enc = new Encoder(SLOTS, modulation_crop != 0);
  g_enc->import_state(blob);
}

void hostsim_work(uint64_t* out, int reset)
{
  for (int i = 0; i < 8; i++)
  {
    out[i] = g_work[i];
    if (reset)
      g_work[i] = 0;
  }
}

void hostsim_stats(uint64_t* out, int n)
{
  const uint64_t v[] = {g_stats.wire_commands, g_stats.packed_prims, g_stats.epochs, g_stats.cuts_raw, g_stats.cuts_war,
                        g_stats.cuts_clut, g_stats.serial_self_sampling, g_stats.serial_copy, g_stats.serial_upload,
                        g_stats.clut_ops, g_stats.clut_cache_hits, g_stats.rejected};
  for (int i = 0; i < n && i < 12; i++)
    out[i] = v[i];
}
}

// ---- the set-up's divisions through double precision (psx_setup.h: trunc_div_rcp, makestep_xy) against integer division ----
extern "C" long hostsim_division_check(long n)
{
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return s;
  };
  long bad = 0;
  for (long it = 0; it < n; it++)
  {
    int32_t a = static_cast<int32_t>(rnd()), d = static_cast<int32_t>(rnd());
    const int mode = static_cast<int>(it & 7);
    if (mode < 4) // determinants of 11-bit coordinate differences
      d = static_cast<int32_t>(rnd() % (1u << (1 + rnd() % 22))) - static_cast<int32_t>(1u << (rnd() % 21));
    if (mode == 1) // exact multiples: the case the remainder test exists for
      a = static_cast<int32_t>(static_cast<uint32_t>(static_cast<int32_t>(rnd() % 4096) - 2048) * static_cast<uint32_t>(d));
    if (mode == 2) // ... and the values right below the next multiple
      a = static_cast<int32_t>(static_cast<uint32_t>(static_cast<int32_t>(rnd() % 4096) - 2048) * static_cast<uint32_t>(d) +
                               static_cast<uint32_t>((d > 0) ? (d - 1) : (d + 1)));
    if (d == 0 || d == INT32_MIN || (a == INT32_MIN && d == -1))
      continue;
    if (trunc_div_rcp(a, d, 1.0 / static_cast<double>(d)) != a / d)
      bad++;
    int32_t dx = static_cast<int32_t>(rnd() % 4096) - 2048, dy = 1 + static_cast<int32_t>(rnd() % 2047);
    if (it & 1)
    {
      dx = static_cast<int32_t>(rnd() % 2097152) - 1048576;
      dy = 1 + static_cast<int32_t>(rnd() % 1048575);
    }
    const int64_t adj = (dx < 0) ? -static_cast<int64_t>(dy - 1) : ((dx > 0) ? static_cast<int64_t>(dy - 1) : 0);
    const int64_t num = static_cast<int64_t>((static_cast<uint64_t>(static_cast<int64_t>(dx)) << 32) + static_cast<uint64_t>(adj));
    if (makestep_xy(dx, dy) != num / dy)
      bad++;
  }
  return bad;
}

// ---- psx_pair.h (two pixels per register) against the scalar pipeline of psx_pixel.h ----
namespace {
uint64_t pair_rng(uint64_t& s)
{
  s += 0x9E3779B97F4A7C15ull;
  return mix64(s);
}
} // namespace

extern "C" {

// n random pairs: random flags / transparency mode / background / texels / colours / position / coverage.
// Returns the number of pairs whose result differs from shade_colour + blend_mask applied to each half.
uint64_t hostsim_pair_selftest(uint64_t seed, uint64_t n)
{
  uint64_t bad = 0, s = seed;
  for (uint64_t i = 0; i < n; i++)
  {
    const uint64_t a = pair_rng(s), b = pair_rng(s), c = pair_rng(s);
    ShadeState st;
    st.flags = static_cast<uint32_t>(a) & 0xFFu;
    st.flags |= (static_cast<uint32_t>(a >> 8) & 1u) ? (F1_DITHER << 8) : 0u;
    st.flags |= ((static_cast<uint32_t>(a >> 9) & 7u) == 0) ? (F1_MOD5 << 8) : 0u;
    st.modes = (static_cast<uint32_t>(a >> 12) & 3u) << 8;
    st.window = 0x0000FFFFu;
    st.texbase = 0;
    uint32_t bg = static_cast<uint32_t>(b), T = static_cast<uint32_t>(b >> 32);
    if (((a >> 16) & 7u) == 0)
      T &= 0xFFFF0000u; // transparent texel in the low half
    if (((a >> 19) & 7u) == 0)
      T &= 0x0000FFFFu;
    if (((a >> 22) & 3u) == 0)
      T &= 0x80008000u | static_cast<uint32_t>(c >> 40); // mask bit alone stays: texel != 0 but colour 0
    const uint32_t cov = static_cast<uint32_t>(a >> 24) & 3u;
    const uint32_t cov15 = ((cov & 1u) ? 0x8000u : 0u) | ((cov & 2u) ? 0x80000000u : 0u);
    const uint32_t cr = static_cast<uint32_t>(c) & 0x00FF00FFu, cg = static_cast<uint32_t>(c >> 8) & 0x00FF00FFu,
                   cb = static_cast<uint32_t>(c >> 32) & 0x00FF00FFu;
    const uint32_t px = (static_cast<uint32_t>(a >> 32) & 1023u) & ~1u, Y = static_cast<uint32_t>(a >> 42) & 511u;
    const bool dither = (st.flags & (F1_DITHER << 8)) != 0;
    const uint32_t d16 = pair_dither_entry(dither ? ((Y & 3u) * 2u + ((px >> 1) & 1u)) : 8u);
    const PairState ps = make_pair_state(st);
    const uint32_t got = pair_shade(ps, bg, cov15, T, cr, cg, cb, d16);
    uint32_t want = bg;
    for (uint32_t h = 0; h < 2; h++)
    {
      if (!((cov >> h) & 1u))
        continue;
      const uint32_t sh = 16u * h;
      uint32_t colour;
      if (!shade_colour(st, px + h, Y, (cr >> sh) & 0xFFu, (cg >> sh) & 0xFFu, (cb >> sh) & 0xFFu, (T >> sh) & 0xFFFFu, colour))
        continue;
      const uint32_t out = blend_mask(st, (bg >> sh) & 0xFFFFu, colour) & 0xFFFFu;
      want = (want & ~(0xFFFFu << sh)) | (out << sh);
    }
    bad += (got != want);
  }
  return bad;
}

// Exhaustive: every (background, foreground) pair of 15-bit colours through the four blend equations, the other half
// of the register filled with noise; compared with blend_mask (textured, texel bit 15 set).
uint64_t hostsim_pair_blend_exhaustive(uint32_t mode, uint32_t bg_step)
{
  uint64_t bad = 0, s = 1234 + mode;
  ShadeState st;
  st.flags = F_TEXTURE | F_TRANSP;
  st.modes = mode << 8;
  st.window = 0x0000FFFFu;
  st.texbase = 0;
  for (uint32_t bg = 0; bg < 0x10000u; bg += bg_step)
    for (uint32_t fg = 0; fg < 0x8000u; fg++)
    {
      const uint32_t noise = static_cast<uint32_t>(pair_rng(s));
      const uint32_t b2 = (bg & 0x7FFFu) | ((noise & 0x7FFFu) << 16), f2 = fg | (noise & 0x7FFF0000u);
      const uint32_t got = pair_blend15(mode, b2, f2);
      const uint32_t w0 = blend_mask(st, bg, fg | 0x8000u), w1 = blend_mask(st, noise & 0x7FFFu, (noise >> 16) | 0x8000u);
      bad += ((got & 0xFFFFu) != (w0 & 0x7FFFu)) + ((got >> 16) != (w1 & 0x7FFFu)) + ((w0 & 0x8000u) == 0);
    }
  return bad;
}

} // extern "C"
