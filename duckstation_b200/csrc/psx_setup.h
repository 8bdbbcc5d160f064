// psx_setup.h — per-primitive setup: PackedCmd -> PrimSetup + BinInfo.
//
// Runs once per primitive in the setup pre-pass (one CUDA thread each), so the raster
// kernel never divides.  Host-compilable (PSX_HD) so tests/hostsim can check the closed
// forms against the oracle without a GPU; the product only ever calls it from device code.
//
// Reference behaviour restated here (paths relative to the reference checkout):
//   triangle sort / edge / attribute setup  src/core/gpu_sw_rasterizer.inl:1417-1551
//   row walk, break and skip rules          src/core/gpu_sw_rasterizer.inl:1314-1413
//   line DDA setup                          src/core/gpu_sw_rasterizer.inl:814-862
//   sprite cull against the clamped area    src/core/gpu_sw.cpp:141-153, src/core/gpu.cpp:2190-2200
#pragma once

#include "psx_types.h"

namespace psx {

// Tile-column mask for VRAM columns (x + j) & 1023, j < w: a run of tiles starting at x's tile, rotated in 16 bits.
PSX_HD uint32_t tile_colmask(uint32_t x, uint32_t w)
{
  if (w == 0)
    return 0;
  x &= (VRAM_W - 1);
  const uint32_t n = ((x % TILE_W) + w + TILE_W - 1) / TILE_W;
  if (n >= TILES_X)
    return 0xFFFFu;
  const uint32_t m = ((1u << n) - 1u) << (x / TILE_W);
  return (m | (m >> TILES_X)) & 0xFFFFu;
}
PSX_HD uint32_t tile_rowmask(uint32_t y, uint32_t h)
{
  if (h == 0)
    return 0;
  y &= (VRAM_H - 1);
  const uint32_t n = ((y % TILE_H) + h + TILE_H - 1) / TILE_H;
  if (n >= TILES_Y)
    return 0xFFFFu;
  const uint32_t m = ((1u << n) - 1u) << (y / TILE_H);
  return (m | (m >> TILES_Y)) & 0xFFFFu;
}

// Tiles a command may READ from VRAM while it is rasterised — the texture footprint (the rule of
// Encoder::texture_footprint, encoder.cpp, at tile granularity: texcoord range of the vertices widened by one step, or
// the window's range, on the texture page) or a copy's source rectangle — as column bits | row bits << 16 of the
// bounding rectangle, wrapping.  0 = reads nothing.  The kernels that order work per tile instead of per epoch
// (raster_instance_kernel) build their dependencies from it; a superset is always safe.
PSX_HD uint32_t read_tilemask(const PackedCmd& c)
{
  if (c.op == OP_COPY)
    return tile_colmask(c.copy.sx, c.copy.w) | (tile_rowmask(c.copy.sy, c.copy.h) << 16);
  if ((c.op != OP_TRIANGLE && c.op != OP_RECT) || !(c.flags0 & F_TEXTURE))
    return 0;
  uint32_t umin, umax, vmin, vmax;
  if (c.op == OP_TRIANGLE)
  {
    umin = c.v[0].u < c.v[1].u ? c.v[0].u : c.v[1].u;
    umin = c.v[2].u < umin ? c.v[2].u : umin;
    umax = c.v[0].u > c.v[1].u ? c.v[0].u : c.v[1].u;
    umax = c.v[2].u > umax ? c.v[2].u : umax;
    vmin = c.v[0].v < c.v[1].v ? c.v[0].v : c.v[1].v;
    vmin = c.v[2].v < vmin ? c.v[2].v : vmin;
    vmax = c.v[0].v > c.v[1].v ? c.v[0].v : c.v[1].v;
    vmax = c.v[2].v > vmax ? c.v[2].v : vmax;
    // interpolated coordinates may overshoot by one step, and wrap at the ends of the range
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
  }
  else
  {
    umin = c.rect.u0;
    umax = c.rect.u0 + (c.rect.w ? c.rect.w - 1u : 0u);
    vmin = c.rect.v0;
    vmax = c.rect.v0 + (c.rect.h ? c.rect.h - 1u : 0u);
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
  }
  if (c.win_and_x != 0xFF || c.win_or_x != 0)
  {
    umin = c.win_or_x;
    umax = c.win_or_x | c.win_and_x;
  }
  if (c.win_and_y != 0xFF || c.win_or_y != 0)
  {
    vmin = c.win_or_y;
    vmax = c.win_or_y | c.win_and_y;
  }
  const uint32_t mode = (c.draw_mode >> 7) & 3u;
  const uint32_t sh = (mode == 0) ? 2u : (mode == 1 ? 1u : 0u);
  const uint32_t bx = (c.draw_mode & 0xFu) * 64u, by = ((c.draw_mode >> 4) & 1u) * 256u;
  return tile_colmask(bx + (umin >> sh), (umax >> sh) - (umin >> sh) + 1u) | (tile_rowmask(by + vmin, vmax - vmin + 1u) << 16);
}

PSX_HD void set_bins(PrimSetup& s, BinInfo& b, int32_t x0, int32_t x1, uint32_t row0, uint32_t rowcount, bool cols_wrap,
                     uint32_t wrap_x, uint32_t wrap_w)
{
  s.by0 = static_cast<uint16_t>(row0 & (VRAM_H - 1));
  s.bycount = static_cast<uint16_t>(rowcount > VRAM_H ? VRAM_H : rowcount);
  uint32_t colmask;
  if (cols_wrap)
  {
    colmask = tile_colmask(wrap_x, wrap_w);
    const bool wraps = ((wrap_x & (VRAM_W - 1)) + wrap_w) > VRAM_W;
    s.bx0 = static_cast<int16_t>(wraps ? 0 : (wrap_x & (VRAM_W - 1)));
    s.bx1 = static_cast<int16_t>(wraps ? (VRAM_W - 1) : ((wrap_x & (VRAM_W - 1)) + wrap_w - 1));
  }
  else
  {
    s.bx0 = static_cast<int16_t>(x0);
    s.bx1 = static_cast<int16_t>(x1);
    colmask = (x1 < x0) ? 0u : tile_colmask(static_cast<uint32_t>(x0), static_cast<uint32_t>(x1 - x0 + 1));
  }
  const uint32_t rowmask = tile_rowmask(s.by0, s.bycount);
  b.tilemask = (colmask && rowmask) ? (colmask | (rowmask << 16)) : 0u;
}

PSX_HD void set_nop(PrimSetup& s, BinInfo& b)
{
  s.op = OP_NOP;
  s.bx0 = 0;
  s.bx1 = -1;
  s.by0 = 0;
  s.bycount = 0;
  b.tilemask = 0;
}

PSX_HD int64_t makefp_xy(int32_t x) // inl:1463
{
  return static_cast<int64_t>((static_cast<uint64_t>(static_cast<int64_t>(x)) << 32) + ((1ull << 32) - (1u << 11)));
}
// The integer divisions of the set-up, through double precision (a 64-bit integer division is a ~140-instruction
// routine on the GPU, a 32-bit one ~25, and a shaded textured triangle has ten of the latter by the same determinant).
// Both are exact:
//  * makestep_xy: |n| < 2^53 is exact as a double and 0 < dy < 2^31; the rounded quotient is off by less than 1, so
//    one look at the remainder settles it.
//  * trunc_div_rcp: trunc(n / d) from rd = 1.0 / d.  n * rd is off by at most |q| * 2^-51 and a true quotient that is
//    not an integer lies at least 1 / |d| from the nearest one; |q| * |d| <= 2^31 makes the first smaller than the second, so
//    truncation can only go wrong when the true quotient IS an integer and the product lands just inside it — the
//    remainder is then +-d, which is what the fix tests.
PSX_HD int64_t makestep_xy(int32_t dx, int32_t dy) // inl:1464-1466, truncating s64 division
{
  const int64_t adj = (dx < 0) ? -static_cast<int64_t>(dy - 1) : ((dx > 0) ? static_cast<int64_t>(dy - 1) : 0);
  const int64_t n = static_cast<int64_t>((static_cast<uint64_t>(static_cast<int64_t>(dx)) << 32) + static_cast<uint64_t>(adj));
  const uint64_t a = static_cast<uint64_t>(n < 0 ? -n : n);
  if (a >= (1ull << 52) || dy <= 0)
    return n / dy; // (never for coordinates the decoder lets through; keeps the function total)
  uint64_t q = static_cast<uint64_t>(static_cast<double>(a) / static_cast<double>(dy));
  const int64_t r = static_cast<int64_t>(a - q * static_cast<uint64_t>(dy));
  if (r < 0)
    q--;
  else if (r >= dy)
    q++;
  return n < 0 ? -static_cast<int64_t>(q) : static_cast<int64_t>(q);
}
PSX_HD int32_t trunc_div_rcp(int32_t n, int32_t d, double rd) // trunc(n / d), rd = 1.0 / d, d != 0, (n, d) != (INT32_MIN, -1)
{
  int32_t q = static_cast<int32_t>(static_cast<double>(n) * rd);
  const int32_t r = static_cast<int32_t>(static_cast<uint32_t>(n) - static_cast<uint32_t>(q) * static_cast<uint32_t>(d));
  const int32_t ad = d < 0 ? -d : d;
  if (r >= ad || r <= -ad)
    q += ((r ^ d) < 0) ? -1 : 1;
  return q;
}
PSX_HD uint32_t attrib_step(int32_t a0, int32_t a1, int32_t a2, int32_t b0, int32_t b1, int32_t b2, int32_t det, double rdet) // inl:1495-1496
{
  const uint32_t num = (static_cast<uint32_t>(a1 - a0) * static_cast<uint32_t>(b2 - b1)) -
                       (static_cast<uint32_t>(a2 - a1) * static_cast<uint32_t>(b1 - b0));
  const int32_t scaled = static_cast<int32_t>(num * 4096u);
  const int32_t q = ((det == -1 && scaled == INT32_MIN) || det == INT32_MIN) ? ((det == -1) ? INT32_MIN : scaled / det) : trunc_div_rcp(scaled, det, rdet);
  return static_cast<uint32_t>(q) << 12;
}

// Downward parts stop at the first row whose 11-bit y is below the drawing area (inl:1391-1394);
// returns that row, or b if none in [a, b).
PSX_HD int32_t first_row_below_area(int32_t a, int32_t b, int32_t bottom)
{
  if (a >= -1024 && b <= 1024)
    return imin(b, imax(a, bottom + 1));
  for (int32_t cy = a; cy < b; cy++)
    if (sext11(cy) > bottom)
      return cy;
  return b;
}
// Upward parts walk cy = b-1 .. a and stop at the first row whose 11-bit y is above the area (inl:1351-1353);
// returns the first row that is still drawn-or-skipped, i.e. (stopping row + 1), or a if none stops it.
PSX_HD int32_t last_row_above_area(int32_t a, int32_t b, int32_t top)
{
  if (a >= -1024 && b <= 1024)
    return imax(a, imin(b, top));
  for (int32_t cy = b - 1; cy >= a; cy--)
    if (sext11(cy) < top)
      return cy + 1;
  return a;
}

// Everything the raster kernel's word loop reads per primitive, derived once here (psx_types.h:LoopDesc).
PSX_HD void make_loop_desc(PrimSetup& s, bool complex_tri)
{
  LoopDesc& D = s.desc;
  const bool tri = s.op == OP_TRIANGLE;
  const bool textured = (s.flags0 & F_TEXTURE) != 0, raw = textured && (s.flags0 & F_RAW);
  const bool mod5 = (s.flags1 & F1_MOD5) != 0;
  const uint32_t mode = s.texmode, tsh = 2u - mode;
  D.and_x2 = s.win_and_x * 0x00010001u;
  D.or_x2 = s.win_or_x * 0x00010001u;
  D.and_y2 = s.win_and_y * 0x00010001u;
  D.or_y2 = s.win_or_y * 0x00010001u;
  D.tex_bx2 = s.tex_bx * 0x00010001u;
  D.sub_mask2 = ((1u << tsh) - 1u) * 0x00010001u;
  D.tex_rows = static_cast<uint32_t>(s.tex_by) * VRAM_W;
  D.tex_shift = tsh;
  D.sub_shift = 2u + mode;
  D.idx_mask = mode < 2u ? ((1u << (4u << mode)) - 1u) : 0u;
  uint32_t r = tri ? s.tri.flat_r : s.rect.r, g = tri ? s.tri.flat_g : s.rect.g, b = tri ? s.tri.flat_b : s.rect.b;
  if (textured)
  {
    if (mod5)
    {
      r &= 0xF8u; // Modulate5Bit: (t * (c >> 3)) >> 1 == (t * (c & 0xF8)) >> 4
      g &= 0xF8u;
      b &= 0xF8u;
    }
  }
  else
  {
    r *= 0x00100010u;
    g *= 0x00100010u;
    b *= 0x00100010u;
  }
  D.fr = r;
  D.fg = g;
  D.fb = b;
  D.chk15 = (s.flags0 & F_CHECK_MASK) ? 0x80008000u : 0u;
  D.set15 = (s.flags0 & F_SET_MASK) ? 0x80008000u : 0u;
  uint32_t path = tri ? LP_TRI : 0u;
  if (textured)
    path |= LP_TEXTURED | (mode < 2u ? LP_PALETTED : 0u);
  if (!raw)
    path |= LP_MODULATED | ((tri && (s.flags0 & F_SHADED)) ? LP_GOURAUD : 0u);
  if (s.flags1 & F1_DITHER)
    path |= LP_DITHER;
  if (s.flags0 & F_TRANSP)
    path |= LP_TRANSP;
  path |= ((static_cast<uint32_t>(s.draw_mode) >> 5) & 3u) << LP_TMODE_SHIFT;
  if (s.win_and_x != 0xFF || s.win_and_y != 0xFF || s.win_or_x != 0 || s.win_or_y != 0)
    path |= LP_WINDOW;
  if (s.flags0 & F_INTERLACED)
    path |= LP_INTERLACED;
  if (complex_tri)
    path |= LP_COMPLEX;
  D.path = path;
  const uint32_t ubase = tri ? 0u : ((static_cast<uint32_t>(s.rect.u0) - static_cast<uint32_t>(static_cast<int32_t>(s.rect.x))) & 0xFFu);
  D.cmask = (textured && mod5) ? 0xF8u : 0xFFu;
  D.ubase = ubase;
  D.pad = 0;
  const uint32_t is8 = mode & 1u;
  D.clut_key = (textured && mode < 2u) ? (s.clut_lo | (static_cast<uint32_t>(is8 ? s.clut_hi : s.clut_lo) << 8) | (is8 << 16) | 0x80000000u) : 0u;
}

PSX_HD void setup_triangle(const PackedCmd& c, PrimSetup& s, BinInfo& bin)
{
  // The reference picks a "core" vertex from the x order of the vertices as submitted (inl:1432-1441: the leftmost,
  // with its tie rules), then sorts the three by y with three neighbour swaps and keeps track of where the core
  // vertex goes (inl:1443-1456).  Here: the core vertex by its submitted index, and the same three swaps on an index
  // permutation.
  const PackedVertex in[3] = {c.v[0], c.v[1], c.v[2]};
  uint32_t core;
  if (in[1].x <= in[0].x)
    core = (in[2].x <= in[1].x) ? 2u : 1u;
  else
    core = (in[2].x < in[0].x) ? 2u : 0u;
  uint32_t order[3] = {0, 1, 2};
  auto swap_if_above = [&](uint32_t a, uint32_t b) {
    if (in[order[b]].y < in[order[a]].y)
    {
      const uint32_t t = order[a];
      order[a] = order[b];
      order[b] = t;
    }
  };
  swap_if_above(1, 2);
  swap_if_above(0, 1);
  swap_if_above(1, 2);
  const PackedVertex v0 = in[order[0]], v1 = in[order[1]], v2 = in[order[2]];
  const uint32_t tl = (order[0] == core) ? 0u : ((order[1] == core) ? 1u : 2u); // position of the core vertex after the sort
  if (v0.y == v2.y)
    return set_nop(s, bin);
  const int32_t det = static_cast<int32_t>((static_cast<uint32_t>(v1.x - v0.x) * static_cast<uint32_t>(v2.y - v1.y)) -
                                           (static_cast<uint32_t>(v2.x - v1.x) * static_cast<uint32_t>(v1.y - v0.y)));
  if (det == 0)
    return set_nop(s, bin);

  const int64_t base_step = makestep_xy(v2.x - v0.x, v2.y - v0.y);
  const int64_t bound_us = (v1.y == v0.y) ? 0 : makestep_xy(v1.x - v0.x, v1.y - v0.y);
  const int64_t bound_ls = (v2.y == v1.y) ? 0 : makestep_xy(v2.x - v1.x, v2.y - v1.y);
  const bool upper_up = (tl != 0);  // vo: rows [y0, y1) are walked upwards from v1
  const bool lower_up = (tl == 2);  // vp: rows [y1, y2) are walked upwards from v2
  const bool right_facing = (v1.y == v0.y) ? (v1.x > v0.x) : (bound_us > base_step);

  auto& T = s.tri;
  T.y0 = v0.y;
  T.y1 = v1.y;
  T.tflags = static_cast<uint8_t>((right_facing ? TRI_SHORT_IS_RIGHT : 0) | (upper_up ? TRI_UPPER_UP : 0) | (lower_up ? TRI_LOWER_UP : 0));

  // effective row ranges (break emulation, inl:1351-1353 / 1391-1394)
  const int32_t top = c.clip_t, bottom = c.clip_b;
  int32_t lo[2], hi[2];
  if (upper_up)
  {
    lo[0] = last_row_above_area(v0.y, v1.y, top);
    hi[0] = v1.y;
  }
  else
  {
    lo[0] = v0.y;
    hi[0] = first_row_below_area(v0.y, v1.y, bottom);
  }
  if (lower_up)
  {
    lo[1] = last_row_above_area(v1.y, v2.y, top);
    hi[1] = v2.y;
  }
  else
  {
    lo[1] = v1.y;
    hi[1] = first_row_below_area(v1.y, v2.y, bottom);
  }
  // The long edge is one line over both parts (both TriangleParts start it at base + (start_y - v0.y) * step); a short
  // edge is anchored at the vertex its part is walked from (the steps are truncated quotients, so the anchor matters).
  // Both are re-anchored at the part's first effective row, in modular arithmetic, so that the kernel multiplies by a
  // row offset >= 0.
  const uint64_t long_start = static_cast<uint64_t>(makefp_xy(v0.x)), long_step = static_cast<uint64_t>(base_step);
  const uint64_t short_start[2] = {static_cast<uint64_t>(makefp_xy(upper_up ? v1.x : v0.x)), static_cast<uint64_t>(makefp_xy(lower_up ? v2.x : v1.x))};
  const int32_t short_anchor[2] = {upper_up ? v1.y : v0.y, lower_up ? v2.y : v1.y};
  const uint64_t short_step[2] = {static_cast<uint64_t>(bound_us), static_cast<uint64_t>(bound_ls)};
  // All coordinates inside 11 bits (with a pixel to spare for the edges' rounding): sext11 is the identity on every
  // row and column the walk can produce, so the drawing area's rows can be cut into the ranges here and the kernel
  // needs neither the row clip nor the column wrap.  Otherwise the kernel takes the general tri_row.
  const bool simple = v0.y >= -1024 && v2.y <= 1024 && imin(v0.x, imin(v1.x, v2.x)) >= -1023 && imax(v0.x, imax(v1.x, v2.x)) <= 1022;
  for (int p = 0; p < 2; p++)
  {
    if (simple)
    {
      lo[p] = imax(lo[p], top);
      hi[p] = imin(hi[p], bottom + 1);
    }
    if (hi[p] < lo[p])
      hi[p] = lo[p];
    T.lo[p] = static_cast<int16_t>(lo[p]);
    T.hi[p] = static_cast<int16_t>(hi[p]);
    const uint64_t ls = long_start + static_cast<uint64_t>(static_cast<int64_t>(lo[p] - v0.y)) * long_step;
    const uint64_t ss = short_start[p] + static_cast<uint64_t>(static_cast<int64_t>(lo[p] - short_anchor[p])) * short_step[p];
    T.part[p].l_start = right_facing ? ls : ss;
    T.part[p].l_step = right_facing ? long_step : short_step[p];
    T.part[p].r_start = right_facing ? ss : ls;
    T.part[p].r_step = right_facing ? short_step[p] : long_step;
  }

  const PackedVertex tv = (tl == 0) ? v0 : ((tl == 1) ? v1 : v2);
  const bool shaded = (c.flags0 & F_SHADED) != 0, textured = (c.flags0 & F_TEXTURE) != 0;
  const double rdet = 1.0 / static_cast<double>(det);
  const int32_t at[5][3] = {{v0.r, v1.r, v2.r}, {v0.g, v1.g, v2.g}, {v0.b, v1.b, v2.b}, {v0.u, v1.u, v2.u}, {v0.v, v1.v, v2.v}};
  const uint32_t start[5] = {tv.r, tv.g, tv.b, tv.u, tv.v};
  for (int i = 0; i < 5; i++)
  {
    const bool on = (i < 3) ? shaded : textured;
    uint32_t base = ((start[i] << 12) + (1u << 11)) << 12, ddx = 0, ddy = 0;
    if (on)
    {
      ddx = attrib_step(at[i][0], at[i][1], at[i][2], v0.y, v1.y, v2.y, det, rdet);
      ddy = attrib_step(v0.x, v1.x, v2.x, at[i][0], at[i][1], at[i][2], det, rdet);
      base += ddx * static_cast<uint32_t>(-static_cast<int32_t>(tv.x)) + ddy * static_cast<uint32_t>(-static_cast<int32_t>(tv.y));
    }
    T.base[i] = base;
    T.ddx[i] = ddx;
    T.ddy[i] = ddy;
  }
  T.flat_r = tv.r;
  T.flat_g = tv.g;
  T.flat_b = tv.b;

  // bounding box (conservative by one pixel; exact coverage is decided per pixel)
  const int32_t minx = imin(v0.x, imin(v1.x, v2.x)), maxx = imax(v0.x, imax(v1.x, v2.x));
  int32_t x0, x1;
  if (minx > -1023 && maxx < 1023)
  {
    x0 = imax(minx - 1, c.clip_l);
    x1 = imin(maxx + 1, c.clip_r);
  }
  else
  {
    x0 = c.clip_l;
    x1 = c.clip_r;
  }
  int32_t r0, rc;
  if (v0.y >= -1024 && v2.y <= 1024)
  {
    const int32_t ylo = imax(v0.y, top), yhi = imin(v2.y - 1, bottom);
    r0 = ylo;
    rc = yhi - ylo + 1;
  }
  else
  {
    r0 = 0;
    rc = static_cast<int32_t>(VRAM_H);
  }
  if (x1 < x0 || rc <= 0)
    return set_nop(s, bin);
  set_bins(s, bin, x0, x1, static_cast<uint32_t>(r0), static_cast<uint32_t>(rc), false, 0, 0);
  make_loop_desc(s, !simple);
}

PSX_HD void setup_rect(const PackedCmd& c, PrimSetup& s, BinInfo& bin)
{
  const int32_t x = c.rect.x, y = c.rect.y, w = c.rect.w, h = c.rect.h;
  // sprite cull against the CLAMPED drawing area (gpu_sw.cpp:141-153; clamp gpu.cpp:2190-2200)
  int32_t cl, ct, cr, cb;
  if (c.clip_l > c.clip_r || c.clip_t > c.clip_b)
  {
    cl = ct = cr = cb = 0;
  }
  else
  {
    cr = imin(c.clip_r + 1, static_cast<int32_t>(VRAM_W));
    cl = imin(c.clip_l, imin(c.clip_r, static_cast<int32_t>(VRAM_W) - 1));
    cb = imin(c.clip_b + 1, static_cast<int32_t>(VRAM_H));
    ct = imin(c.clip_t, imin(c.clip_b, static_cast<int32_t>(VRAM_H) - 1));
  }
  if (imin(cr, x + w) <= imax(cl, x) || imin(cb, y + h) <= imax(ct, y))
    return set_nop(s, bin);
  s.rect.x = c.rect.x;
  s.rect.y = c.rect.y;
  s.rect.w = c.rect.w;
  s.rect.h = c.rect.h;
  s.rect.r = c.rect.r;
  s.rect.g = c.rect.g;
  s.rect.b = c.rect.b;
  s.rect.u0 = c.rect.u0;
  s.rect.v0 = c.rect.v0;
  // the rasteriser itself clips to the RAW area (inl:767-786)
  const int32_t x0 = imax(x, c.clip_l), x1 = imin(x + w - 1, c.clip_r);
  const int32_t y0 = imax(y, c.clip_t), y1 = imin(y + h - 1, c.clip_b);
  if (x1 < x0 || y1 < y0)
    return set_nop(s, bin);
  set_bins(s, bin, x0, x1, static_cast<uint32_t>(y0), static_cast<uint32_t>(y1 - y0 + 1), false, 0, 0);
  make_loop_desc(s, false);
}

PSX_HD int32_t iabs(int32_t v)
{
  return v < 0 ? -v : v;
}

PSX_HD void setup_line(const PackedCmd& c, PrimSetup& s, BinInfo& bin)
{
  PackedVertex p0 = c.v[0], p1 = c.v[1];
  const int32_t adx = iabs(p1.x - p0.x), ady = iabs(p1.y - p0.y);
  const int32_t k = (adx > ady) ? adx : ady;
  if (adx >= 1024 || ady >= 512)
    return set_nop(s, bin);
  if (p0.x >= p1.x && k > 0)
  {
    const PackedVertex t = p0;
    p0 = p1;
    p1 = t;
  }
  auto& L = s.line;
  int64_t dxdk = 0, dydk = 0;
  int32_t dr = 0, dg = 0, db = 0;
  if (k != 0)
  {
    const int64_t ddx = p1.x - p0.x, ddy = p1.y - p0.y;
    dxdk = static_cast<int64_t>((static_cast<uint64_t>(ddx) << 32) - static_cast<uint64_t>((ddx < 0) ? (k - 1) : 0) +
                                static_cast<uint64_t>((ddx > 0) ? (k - 1) : 0)) /
           k;
    dydk = static_cast<int64_t>((static_cast<uint64_t>(ddy) << 32) - static_cast<uint64_t>((ddy < 0) ? (k - 1) : 0) +
                                static_cast<uint64_t>((ddy > 0) ? (k - 1) : 0)) /
           k;
    if (c.flags0 & F_SHADED)
    {
      dr = static_cast<int32_t>(static_cast<uint32_t>(static_cast<int32_t>(p1.r) - static_cast<int32_t>(p0.r)) << 12) / k;
      dg = static_cast<int32_t>(static_cast<uint32_t>(static_cast<int32_t>(p1.g) - static_cast<int32_t>(p0.g)) << 12) / k;
      db = static_cast<int32_t>(static_cast<uint32_t>(static_cast<int32_t>(p1.b) - static_cast<int32_t>(p0.b)) << 12) / k;
    }
  }
  L.curx0 = ((static_cast<uint64_t>(static_cast<int64_t>(p0.x)) << 32) | (1ull << 31)) - 1024u;
  L.cury0 = ((static_cast<uint64_t>(static_cast<int64_t>(p0.y)) << 32) | (1ull << 31)) - ((dydk < 0) ? 1024u : 0u);
  L.dxdk = static_cast<uint64_t>(dxdk);
  L.dydk = static_cast<uint64_t>(dydk);
  L.r0 = (static_cast<uint32_t>(p0.r) << 12) | (1u << 11);
  L.g0 = (static_cast<uint32_t>(p0.g) << 12) | (1u << 11);
  L.b0 = (static_cast<uint32_t>(p0.b) << 12) | (1u << 11);
  L.dr = dr;
  L.dg = dg;
  L.db = db;
  L.k = k;
  L.x0 = p0.x;
  L.y0 = p0.y;
  L.x_major = (adx > ady) ? 1 : 0; // k == adx; along x the DDA advances exactly one pixel per step
  L.flat_r = p0.r;
  L.flat_g = p0.g;
  L.flat_b = p0.b;

  const int32_t minx = imin(p0.x, p1.x), maxx = imax(p0.x, p1.x), miny = imin(p0.y, p1.y), maxy = imax(p0.y, p1.y);
  int32_t x0 = c.clip_l, x1 = c.clip_r, r0 = 0, rc = static_cast<int32_t>(VRAM_H);
  if (minx >= -1024 && maxx <= 1023 && miny >= -1024 && maxy <= 1023)
  {
    x0 = imax(minx, c.clip_l);
    x1 = imin(maxx, c.clip_r);
    const int32_t ylo = imax(miny, c.clip_t), yhi = imin(maxy, c.clip_b);
    r0 = ylo;
    rc = yhi - ylo + 1;
  }
  if (x1 < x0 || rc <= 0)
    return set_nop(s, bin);
  set_bins(s, bin, x0, x1, static_cast<uint32_t>(r0), static_cast<uint32_t>(rc), false, 0, 0);
}

PSX_HD void setup_prim(const PackedCmd& c, PrimSetup& s, BinInfo& bin)
{
  s.op = c.op;
  s.flags0 = c.flags0;
  s.flags1 = c.flags1;
  const uint32_t tm = (c.draw_mode >> 7) & 3u;
  s.texmode = static_cast<uint8_t>(tm > 2 ? 2 : tm); // mode 3 behaves as direct 15-bit (inl:117)
  s.draw_mode = c.draw_mode;
  s.clut_lo = c.clut_lo;
  s.clut_hi = c.clut_hi;
  s.win_and_x = c.win_and_x;
  s.win_and_y = c.win_and_y;
  s.win_or_x = c.win_or_x;
  s.win_or_y = c.win_or_y;
  s.clip_l = static_cast<int16_t>(c.clip_l);
  s.clip_t = static_cast<int16_t>(c.clip_t);
  s.clip_r = static_cast<int16_t>(c.clip_r);
  s.clip_b = static_cast<int16_t>(c.clip_b);
  s.tex_bx = static_cast<uint16_t>((c.draw_mode & 0xFu) * 64u);
  s.tex_by = static_cast<uint16_t>(((c.draw_mode >> 4) & 1u) * 256u);
  switch (c.op)
  {
    case OP_TRIANGLE: setup_triangle(c, s, bin); break;
    case OP_RECT: setup_rect(c, s, bin); break;
    case OP_LINE: setup_line(c, s, bin); break;
    case OP_FILL:
      s.fill.x = c.fill.x;
      s.fill.y = c.fill.y;
      s.fill.w = c.fill.w;
      s.fill.h = c.fill.h;
      s.fill.color16 = c.fill.color16;
      s.fill.interlaced = c.fill.interlaced;
      s.fill.active_lsb = c.fill.active_lsb;
      if (c.fill.w == 0 || c.fill.h == 0)
        return set_nop(s, bin);
      set_bins(s, bin, 0, 0, c.fill.y, c.fill.h, true, c.fill.x, c.fill.w);
      break;
    case OP_COPY:
      s.copy.sx = c.copy.sx;
      s.copy.sy = c.copy.sy;
      s.copy.dx = c.copy.dx;
      s.copy.dy = c.copy.dy;
      s.copy.w = c.copy.w;
      s.copy.h = c.copy.h;
      if (c.copy.w == 0 || c.copy.h == 0)
        return set_nop(s, bin);
      set_bins(s, bin, 0, 0, c.copy.dy, c.copy.h, true, c.copy.dx, c.copy.w);
      break;
    case OP_UPLOAD:
      s.upload.x = c.upload.x;
      s.upload.y = c.upload.y;
      s.upload.w = c.upload.w;
      s.upload.h = c.upload.h;
      s.upload.payload = c.upload.payload;
      if (c.upload.w == 0 || c.upload.h == 0)
        return set_nop(s, bin);
      set_bins(s, bin, 0, 0, c.upload.y, c.upload.h, true, c.upload.x, c.upload.w);
      break;
    default: set_nop(s, bin); break;
  }
}

} // namespace psx
