// psx_pixel.h — the PSX pixel pipeline and the closed-form coverage queries.
//
// "Given VRAM pixel (px, Y), is it covered by this primitive, and with which interpolants?"
// answered directly from a PrimSetup, so that any thread can evaluate any pixel in any order.
// Restates, per pixel, what the reference computes by stepping:
//   span walk / clip        src/core/gpu_sw_rasterizer.inl:1203-1312, 1314-1413
//   ShadePixel              src/core/gpu_sw_rasterizer.inl:524-736 (vector form == scalar form per pixel)
//   DrawRectangle           src/core/gpu_sw_rasterizer.inl:738-809
//   DrawLine                src/core/gpu_sw_rasterizer.inl:814-892
//   FillVRAM / CopyVRAM / WriteVRAM  src/core/gpu_sw_rasterizer.inl:1639-1939
// Host-compilable for tests/hostsim; the product only calls it from device code.
#pragma once

#include "psx_types.h"

#if defined(__CUDA_ARCH__)
#define PSX_LD_TEX(ptr) __ldg(ptr) // texture pages / CLUT slots / upload payload: read-only within an epoch
#else
#define PSX_LD_TEX(ptr) (*(ptr))
#endif

namespace psx {

// ---- colour stage ----
PSX_HD int32_t dither_offset(uint32_t x, uint32_t y)
{
  // gpu_types.h:269-272 with +4 bias, 4 bits per entry, 16 bits per row:
  // {-4,0,-3,1} {2,-2,3,-1} {-3,1,-4,0} {3,-1,2,-2}
  const uint64_t m = 0x2637405137265140ull;
  return static_cast<int32_t>((m >> (((y & 3u) * 16u) + ((x & 3u) * 4u))) & 0xFu) - 4;
}
// clamp((v + d) >> 3, 0, 31) == clamp(v + d, 0, 255) >> 3 (the shift is monotone), which is one fused
// add-min-relu instruction on sm_90+.
PSX_HD uint32_t dither5(int32_t v, int32_t d)
{
#if defined(__CUDA_ARCH__)
  return static_cast<uint32_t>(__viaddmin_s32_relu(v, d, 255)) >> 3;
#else
  const int32_t t = v + d;
  return static_cast<uint32_t>(t < 0 ? 0 : (t > 255 ? 255 : t)) >> 3;
#endif
}

// Everything the pixel pipeline needs from a primitive, packed into five registers so that the raster kernel can keep
// it live across its pixel loop (reading it back from the staged record after every shared-memory store is what the
// compiler would otherwise have to do).
struct ShadeState
{
  uint32_t flags;   // flags0 | flags1 << 8
  uint32_t modes;   // texmode (0 4-bit, 1 8-bit, 2 direct) | transparency mode << 8
  uint32_t window;  // and_x | and_y << 8 | or_x << 16 | or_y << 24
  uint32_t texbase; // page base x | page base y << 16
};

PSX_HD ShadeState make_shade_state(const PrimSetup& s)
{
  ShadeState st;
  st.flags = static_cast<uint32_t>(s.flags0) | (static_cast<uint32_t>(s.flags1) << 8);
  st.modes = static_cast<uint32_t>(s.texmode) | (((static_cast<uint32_t>(s.draw_mode) >> 5) & 3u) << 8);
  st.window = static_cast<uint32_t>(s.win_and_x) | (static_cast<uint32_t>(s.win_and_y) << 8) |
              (static_cast<uint32_t>(s.win_or_x) << 16) | (static_cast<uint32_t>(s.win_or_y) << 24);
  st.texbase = static_cast<uint32_t>(s.tex_bx) | (static_cast<uint32_t>(s.tex_by) << 16);
  return st;
}

// Texture window + page addressing; returns the VRAM pixel index of the texel word, the sub-word shift and the palette
// index mask (0 = direct colour).  Branch-free in the texture mode: 4-bit packs 4 indices per word, 8-bit 2, direct 1.
PSX_HD uint32_t texel_addr(const ShadeState& st, uint32_t u, uint32_t v, uint32_t& shift, uint32_t& idxmask)
{
  u = (u & (st.window & 0xFFu)) | ((st.window >> 16) & 0xFFu);
  v = (v & ((st.window >> 8) & 0xFFu)) | (st.window >> 24);
  const uint32_t mode = st.modes & 0xFFu;
  const uint32_t sh = 2u - mode; // log2(texels per VRAM word)
  const uint32_t ty = ((st.texbase >> 16) + v) & (VRAM_H - 1);
  const uint32_t tx = ((st.texbase & 0xFFFFu) + (u >> sh)) & (VRAM_W - 1);
  shift = (u & ((1u << sh) - 1u)) << (2u + mode);
  idxmask = (mode == 2u) ? 0u : ((1u << (4u << mode)) - 1u);
  return ty * VRAM_W + tx;
}

// Full texel fetch through the CLUT snapshots.  `clut_lo` holds palette entries 0..15 and `clut_hi` entries 16..255:
// a 4-bit UpdateCLUT only replaces the first 16 entries of the reference's cache (gpu_sw_rasterizer.cpp:47-51).
// kLive = the texel word is read from VRAM that this very primitive is writing (self-sampling): the
// read must then be a coherent (volatile) load instead of going through the read-only path.
template<bool kLive, bool kClutInShared = false>
PSX_HD uint32_t fetch_texel(const ShadeState& st, const uint16_t* vram, const uint16_t* __restrict__ clut_lo,
                            const uint16_t* __restrict__ clut_hi, uint32_t u, uint32_t v)
{
  uint32_t shift, idxmask;
  const uint32_t a = texel_addr(st, u, v, shift, idxmask);
  uint32_t w;
  if (kLive)
    w = *reinterpret_cast<const volatile uint16_t*>(vram + a);
  else
    w = PSX_LD_TEX(vram + a);
  if (idxmask == 0)
    return w;
  const uint32_t idx = (w >> shift) & idxmask;
  if (kClutInShared)
    return clut_lo[idx]; // the raster kernel stages the merged 256-entry palette in shared memory
  return PSX_LD_TEX((idx < 16u ? clut_lo : clut_hi) + idx);
}

// Colour before blending.  Returns false if the pixel is dropped (texel == 0).
PSX_HD bool shade_colour(const ShadeState& st, uint32_t px, uint32_t Y, uint32_t r, uint32_t g, uint32_t b, uint32_t texel,
                         uint32_t& out)
{
  const int32_t dv = (st.flags & (F1_DITHER << 8)) ? dither_offset(px, Y) : 0;
  if (!(st.flags & F_TEXTURE))
  {
    out = dither5(static_cast<int32_t>(r), dv) | (dither5(static_cast<int32_t>(g), dv) << 5) |
          (dither5(static_cast<int32_t>(b), dv) << 10) | ((st.flags & F_TRANSP) ? 0x8000u : 0u);
    return true;
  }
  if (texel == 0)
    return false;
  if (st.flags & F_RAW)
  {
    out = texel;
    return true;
  }
  const uint32_t tr = texel & 31u, tg = (texel >> 5) & 31u, tb = (texel >> 10) & 31u;
  uint32_t cr, cg, cb;
  if (!(st.flags & (F1_MOD5 << 8)))
  {
    cr = dither5(static_cast<int32_t>((tr * r) >> 4), dv);
    cg = dither5(static_cast<int32_t>((tg * g) >> 4), dv);
    cb = dither5(static_cast<int32_t>((tb * b) >> 4), dv);
  }
  else
  {
    cr = dither5(static_cast<int32_t>((tr * (r >> 3)) >> 1), dv);
    cg = dither5(static_cast<int32_t>((tg * (g >> 3)) >> 1), dv);
    cb = dither5(static_cast<int32_t>((tb * (b >> 3)) >> 1), dv);
  }
  out = cr | (cg << 5) | (cb << 10) | (texel & 0x8000u);
  return true;
}

// Blend (if enabled and applicable), mask test, mask set.  Returns the new pixel value (bg if masked).
PSX_HD uint32_t blend_mask(const ShadeState& st, uint32_t bg, uint32_t color)
{
  if ((st.flags & F_CHECK_MASK) && (bg & 0x8000u))
    return bg;
  const bool textured = (st.flags & F_TEXTURE) != 0;
  if ((st.flags & F_TRANSP) && ((color & 0x8000u) || !textured))
  {
    uint32_t bgb = bg, fgb = color;
    switch ((st.modes >> 8) & 3u)
    {
      case 0:
        bgb |= 0x8000u;
        color = (((fgb + bgb) - ((fgb ^ bgb) & 0x0421u)) >> 1) & 0xFFFFu;
        break;
      case 1:
      {
        bgb &= ~0x8000u;
        const uint32_t sum = fgb + bgb;
        const uint32_t carry = (sum - ((fgb ^ bgb) & 0x8421u)) & 0x8420u;
        color = ((sum - carry) | (carry - (carry >> 5))) & 0xFFFFu;
      }
      break;
      case 2:
      {
        bgb |= 0x8000u;
        fgb &= ~0x8000u;
        const uint32_t diff = bgb - fgb + 0x108420u;
        const uint32_t borrow = (diff - ((bgb ^ fgb) & 0x108420u)) & 0x108420u;
        color = ((diff - borrow) & (borrow - (borrow >> 5))) & 0xFFFFu;
      }
      break;
      default:
      {
        bgb &= ~0x8000u;
        fgb = ((fgb >> 2) & 0x1CE7u) | 0x8000u;
        const uint32_t sum = fgb + bgb;
        const uint32_t carry = (sum - ((fgb ^ bgb) & 0x8421u)) & 0x8420u;
        color = ((sum - carry) | (carry - (carry >> 5))) & 0xFFFFu;
      }
      break;
    }
    if (!textured)
      color &= 0x7FFFu;
  }
  return color | ((st.flags & F_SET_MASK) ? 0x8000u : 0u);
}

PSX_HD bool interlace_skip(const PrimSetup& s, uint32_t line)
{
  return (s.flags0 & F_INTERLACED) && ((((s.flags0 & F_ACTIVE_LSB) != 0) ? 1u : 0u) == (line & 1u));
}

// ---- triangles ----
struct TriRow
{
  int32_t cy;     // un-truncated row
  int32_t x_lo;   // first covered VRAM column (clipped)
  int32_t x_hi;   // one past the last covered column
  int32_t xdelta; // x_raw = px + xdelta feeds the interpolants
};

// Is VRAM row Y (0..511) a drawn row of this triangle?  If so, its clipped span.
PSX_HD bool tri_row(const PrimSetup& s, uint32_t Y, TriRow& row)
{
  const auto& T = s.tri;
  const int32_t cy = T.y0 + static_cast<int32_t>((Y - static_cast<uint32_t>(static_cast<int32_t>(T.y0))) & (VRAM_H - 1));
  const uint32_t p = (cy >= T.y1) ? 1u : 0u;
  const uint32_t k = static_cast<uint32_t>(cy - T.lo[p]); // row within the part
  if (k >= static_cast<uint32_t>(T.hi[p] - T.lo[p]))
    return false;
  const int32_t y = sext11(cy);
  if (y < s.clip_t || y > s.clip_b || interlace_skip(s, static_cast<uint32_t>(cy)))
    return false;
  const TriPart& P = T.part[p];
  const int32_t x_start = static_cast<int32_t>((P.l_start + static_cast<uint64_t>(k) * P.l_step) >> 32);
  const int32_t x_bound = static_cast<int32_t>((P.r_start + static_cast<uint64_t>(k) * P.r_step) >> 32);
  int32_t width = x_bound - x_start;
  int32_t cx = sext11(x_start);
  row.xdelta = x_start - cx;
  if (cx < s.clip_l)
  {
    width -= (s.clip_l - cx);
    cx = s.clip_l;
  }
  if (cx + width > s.clip_r + 1)
    width = s.clip_r + 1 - cx;
  if (width <= 0)
    return false;
  row.cy = cy;
  row.x_lo = cx;
  row.x_hi = cx + width;
  return true;
}

PSX_HD void tri_attrs(const PrimSetup& s, const TriRow& row, int32_t px, uint32_t& r, uint32_t& g, uint32_t& b, uint32_t& u,
                      uint32_t& v)
{
  const auto& T = s.tri;
  const uint32_t xr = static_cast<uint32_t>(px + row.xdelta), cy = static_cast<uint32_t>(row.cy);
  if (s.flags0 & F_SHADED)
  {
    r = (T.base[0] + T.ddx[0] * xr + T.ddy[0] * cy) >> 24;
    g = (T.base[1] + T.ddx[1] * xr + T.ddy[1] * cy) >> 24;
    b = (T.base[2] + T.ddx[2] * xr + T.ddy[2] * cy) >> 24;
  }
  else
  {
    r = T.flat_r;
    g = T.flat_g;
    b = T.flat_b;
  }
  if (s.flags0 & F_TEXTURE)
  {
    u = (T.base[3] + T.ddx[3] * xr + T.ddy[3] * cy) >> 24;
    v = (T.base[4] + T.ddx[4] * xr + T.ddy[4] * cy) >> 24;
  }
  else
  {
    u = v = 0;
  }
}

// ---- rectangles ----
struct RectRow
{
  int32_t x_lo, x_hi; // covered VRAM columns [x_lo, x_hi)
  uint32_t v;         // texture row before the window
  int32_t y;          // un-wrapped y (interlace / dither use y & 511 == Y)
};
PSX_HD bool rect_row(const PrimSetup& s, uint32_t Y, RectRow& row)
{
  const auto& R = s.rect;
  const uint32_t oy = (Y - static_cast<uint32_t>(static_cast<int32_t>(R.y))) & (VRAM_H - 1);
  if (oy >= R.h)
    return false;
  const int32_t y = R.y + static_cast<int32_t>(oy);
  if (y < s.clip_t || y > s.clip_b || interlace_skip(s, static_cast<uint32_t>(y)))
    return false;
  const int32_t x_lo = imax(R.x, s.clip_l), x_hi = imin(R.x + static_cast<int32_t>(R.w), s.clip_r + 1);
  if (x_hi <= x_lo)
    return false;
  row.x_lo = x_lo;
  row.x_hi = x_hi;
  row.v = (R.v0 + oy) & 0xFFu;
  row.y = y;
  return true;
}

// ---- lines: the i-th DDA step that lands on column px (x-major) or on row Y (y-major), if any ----
struct LinePixel
{
  int32_t x, y; // 11-bit truncated position
  uint32_t r, g, b;
};
PSX_HD bool line_step(const PrimSetup& s, int32_t i, LinePixel& p)
{
  const auto& L = s.line;
  if (i < 0 || i > L.k)
    return false;
  const uint64_t cx = L.curx0 + static_cast<uint64_t>(static_cast<int64_t>(i)) * L.dxdk;
  const uint64_t cy = L.cury0 + static_cast<uint64_t>(static_cast<int64_t>(i)) * L.dydk;
  p.x = sext11(static_cast<int32_t>(static_cast<int64_t>(cx) >> 32));
  p.y = sext11(static_cast<int32_t>(static_cast<int64_t>(cy) >> 32));
  if (p.x < s.clip_l || p.x > s.clip_r || p.y < s.clip_t || p.y > s.clip_b || interlace_skip(s, static_cast<uint32_t>(p.y)))
    return false;
  if (s.flags0 & F_SHADED)
  {
    p.r = static_cast<uint32_t>(static_cast<int32_t>(L.r0 + static_cast<uint32_t>(L.dr) * static_cast<uint32_t>(i)) >> 12) & 0xFFu;
    p.g = static_cast<uint32_t>(static_cast<int32_t>(L.g0 + static_cast<uint32_t>(L.dg) * static_cast<uint32_t>(i)) >> 12) & 0xFFu;
    p.b = static_cast<uint32_t>(static_cast<int32_t>(L.b0 + static_cast<uint32_t>(L.db) * static_cast<uint32_t>(i)) >> 12) & 0xFFu;
  }
  else
  {
    p.r = L.flat_r;
    p.g = L.flat_g;
    p.b = L.flat_b;
  }
  return true;
}
// Candidate step for a column (x-major lines advance exactly one column per step).
PSX_HD int32_t line_step_for_column(const PrimSetup& s, int32_t px)
{
  return static_cast<int32_t>(static_cast<uint32_t>(px - s.line.x0) & 2047u);
}
// Candidate step for a VRAM row (y-major lines advance exactly one row per step, |dy| < 512).
PSX_HD int32_t line_step_for_row(const PrimSetup& s, uint32_t Y)
{
  const uint32_t d = Y - static_cast<uint32_t>(static_cast<int32_t>(s.line.y0));
  const bool up = static_cast<int64_t>(s.line.dydk) < 0;
  return static_cast<int32_t>((up ? (0u - d) : d) & (VRAM_H - 1));
}

// ---- transfers (tile-parallel forms; aliasing / mask-tail cases take the serial kernels) ----
PSX_HD bool fill_covers(const PrimSetup& s, uint32_t px, uint32_t Y)
{
  const auto& F = s.fill;
  if (((px - F.x) & (VRAM_W - 1)) >= F.w || ((Y - F.y) & (VRAM_H - 1)) >= F.h)
    return false;
  return !(F.interlaced && (Y & 1u) == F.active_lsb);
}
PSX_HD bool copy_covers(const PrimSetup& s, uint32_t px, uint32_t Y, uint32_t& src_index)
{
  const auto& C = s.copy;
  const uint32_t c = (px - C.dx) & (VRAM_W - 1), r = (Y - C.dy) & (VRAM_H - 1);
  if (c >= C.w || r >= C.h)
    return false;
  src_index = ((C.sy + r) & (VRAM_H - 1)) * VRAM_W + ((C.sx + c) & (VRAM_W - 1));
  return true;
}
PSX_HD bool upload_covers(const PrimSetup& s, uint32_t px, uint32_t Y, uint32_t& payload_index)
{
  const auto& U = s.upload;
  const uint32_t c = (px - U.x) & (VRAM_W - 1), r = (Y - U.y) & (VRAM_H - 1);
  if (c >= U.w || r >= U.h)
    return false;
  payload_index = U.payload + r * U.w + c;
  return true;
}
PSX_HD uint16_t transfer_write(const PrimSetup& s, uint16_t bg, uint16_t src)
{
  if ((s.flags0 & F_CHECK_MASK) && (bg & 0x8000u))
    return bg;
  return static_cast<uint16_t>(src | ((s.flags0 & F_SET_MASK) ? 0x8000u : 0u));
}

} // namespace psx
