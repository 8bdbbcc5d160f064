/* psx_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of the PSX software rasteriser path of
 * stenzek/duckstation (GPU_SW + GPU_SW_Rasterizer), written from a reading of the
 * reference and pinned against the reference's own translation unit compiled in
 * place (oracle/_ref/libpsxref.so, see oracle/ref_build/) by tests/test_oracle_vs_ref.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the CUDA product never does.
 *
 * It reproduces the behaviour of the reference's SIMD instantiation (the x86 default,
 * 4 pixels per vector), including
 *   Q1  4-pixel gather-then-store groups when a primitive samples what it writes,
 *   Q2  the mask-checked CPU->VRAM upload tail (source pointer only advances on write).
 *
 * All arithmetic is integer.  Signed products that may wrap are done in uint32_t.
 * Every function cites the reference lines it follows (paths relative to the
 * reference checkout: src/core/...).
 */
#include "../include/psxgpu_wire.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t s32;
typedef int64_t s64;

#define VW 1024u
#define VH 512u

static u16 o_vram[VW * VH];
static u16 o_clut[256];
static psx_drawing_area o_area;      /* g_drawing_area, gpu_sw_rasterizer.cpp:40 */
static s32 o_clamped[4];             /* m_clamped_drawing_area, gpu.cpp:2190-2200 */
static int o_modulation_crop = 0;    /* settings.h:83 */

/* Pixel-class counters for the roofline byte model. */
typedef struct psxo_stats {
  u64 draw_px;      /* ShadePixel invocations (covered, clipped, not interlace-skipped) */
  u64 draw_px_rmw;  /* of which transparency_enable || check_mask */
  u64 draw_px_tex;  /* of which texture_enable */
  u64 fill_px, copy_px, copy_px_masked, upload_px, upload_px_masked;
  u64 prims;        /* triangles + rectangles + line segments handed to the rasteriser */
  u64 commands;
} psxo_stats;
static psxo_stats o_stats;

static inline s32 sext11(s32 v) /* TruncateGPUVertexPosition, gpu_helpers.h:137-140 */
{
  return (s32)((u32)v << 21) >> 21;
}

/* ---- dither: g_dither_lut[i][j][v] = clamp((v + M[i][j]) >> 3, 0, 31), gpu_sw_rasterizer.cpp:18-32;
 *      matrix gpu_types.h:269-272 ---- */
static const s32 k_dither[4][4] = {{-4, 0, -3, 1}, {2, -2, 3, -1}, {-3, 1, -4, 0}, {3, -1, 2, -2}};
static inline u32 dither5(s32 v, s32 d)
{
  s32 r = (v + d) >> 3;
  return (u32)(r < 0 ? 0 : (r > 31 ? 31 : r));
}

enum { MOD_DISABLED = 0, MOD_RAW = 1, MOD_8BIT = 2, MOD_5BIT = 3 }; /* gpu_sw_rasterizer.h:16-22 */

static inline int mod_mode(u8 flags0) /* GetModulationMode, gpu_sw_rasterizer.h:63-71 */
{
  if (!(flags0 & PSX_DF_TEXTURE))
    return MOD_DISABLED;
  if (flags0 & PSX_DF_RAW_TEXTURE)
    return MOD_RAW;
  return o_modulation_crop ? MOD_5BIT : MOD_8BIT;
}

/* Texel fetch incl. texture window and CLUT; ShadePixel (a)-(b), gpu_sw_rasterizer.inl:88-123 / 538-580. */
static inline u16 fetch_texel(const psx_draw_header* d, u8 u, u8 v)
{
  const u32 bx = PSX_DM_PAGE_X(d->draw_mode) * 64u, by = PSX_DM_PAGE_Y(d->draw_mode) * 256u;
  u = (u8)((u & d->window.and_x) | d->window.or_x);
  v = (u8)((v & d->window.and_y) | d->window.or_y);
  const u32 ty = (by + v) & (VH - 1);
  switch (PSX_DM_TEXMODE(d->draw_mode)) {
  case 0: {
    const u16 w = o_vram[ty * VW + ((bx + (u >> 2)) & (VW - 1))];
    return o_clut[(w >> ((u & 3u) * 4u)) & 0xFu];
  }
  case 1: {
    const u16 w = o_vram[ty * VW + ((bx + (u >> 1)) & (VW - 1))];
    return o_clut[(w >> ((u & 1u) * 8u)) & 0xFFu];
  }
  default: /* modes 2 and 3: direct 15-bit (inl:117) */
    return o_vram[ty * VW + ((bx + u) & (VW - 1))];
  }
}

/* Colour stage of ShadePixel (c)-(d): returns 0 when the pixel is skipped (texel 0), else 1 and *out.
 * inl:125-170 (scalar form), 589-644 (vector form: max(.,0)>>3 then min 31 — same function). */
static inline int shade_colour(const psx_draw_header* d, int mode, u32 x, u32 y, u8 r, u8 g, u8 b, u16 texel, u16* out)
{
  const int dith = (d->flags1 & PSX_DF1_DITHER) != 0;
  const s32 dv = dith ? k_dither[y & 3u][x & 3u] : 0;
  if (mode == MOD_DISABLED) {
    *out = (u16)(dither5(r, dv) | (dither5(g, dv) << 5) | (dither5(b, dv) << 10) |
                 ((d->flags0 & PSX_DF_TRANSPARENCY) ? 0x8000u : 0u));
    return 1;
  }
  if (texel == 0)
    return 0;
  if (mode == MOD_RAW) {
    *out = texel;
    return 1;
  }
  const u32 tr = texel & 31u, tg = (texel >> 5) & 31u, tb = (texel >> 10) & 31u;
  u32 cr, cg, cb;
  if (mode == MOD_8BIT) {
    cr = dither5((s32)((tr * r) >> 4), dv);
    cg = dither5((s32)((tg * g) >> 4), dv);
    cb = dither5((s32)((tb * b) >> 4), dv);
  } else {
    cr = dither5((s32)((tr * (u32)(r >> 3)) >> 1), dv);
    cg = dither5((s32)((tg * (u32)(g >> 3)) >> 1), dv);
    cb = dither5((s32)((tb * (u32)(b >> 3)) >> 1), dv);
  }
  *out = (u16)(cr | (cg << 5) | (cb << 10) | (texel & 0x8000u));
  return 1;
}

/* Blend + mask + store: ShadePixel (e)-(h), inl:172-239. */
static inline void blend_store(const psx_draw_header* d, int textured, u32 x, u32 y, u16 color)
{
  u16* p = &o_vram[y * VW + x];
  const u16 bg = *p;
  if ((d->flags0 & PSX_DF_TRANSPARENCY) && ((color & 0x8000u) || !textured)) {
    u32 bgb = bg, fgb = color;
    switch (PSX_DM_TRANSPARENCY(d->draw_mode)) {
    case 0:
      bgb |= 0x8000u;
      color = (u16)(((fgb + bgb) - ((fgb ^ bgb) & 0x0421u)) >> 1);
      break;
    case 1: {
      bgb &= ~0x8000u;
      const u32 sum = fgb + bgb;
      const u32 carry = (sum - ((fgb ^ bgb) & 0x8421u)) & 0x8420u;
      color = (u16)((sum - carry) | (carry - (carry >> 5)));
    } break;
    case 2: {
      bgb |= 0x8000u;
      fgb &= ~0x8000u;
      const u32 diff = bgb - fgb + 0x108420u;
      const u32 borrow = (diff - ((bgb ^ fgb) & 0x108420u)) & 0x108420u;
      color = (u16)((diff - borrow) & (borrow - (borrow >> 5)));
    } break;
    default: {
      bgb &= ~0x8000u;
      fgb = ((fgb >> 2) & 0x1CE7u) | 0x8000u;
      const u32 sum = fgb + bgb;
      const u32 carry = (sum - ((fgb ^ bgb) & 0x8421u)) & 0x8420u;
      color = (u16)((sum - carry) | (carry - (carry >> 5)));
    } break;
    }
    if (!textured)
      color &= 0x7FFFu;
  }
  if ((d->flags0 & PSX_DF_CHECK_MASK) && (bg & 0x8000u))
    return;
  *p = (u16)(color | ((d->flags0 & PSX_DF_SET_MASK) ? 0x8000u : 0u));
}

static inline void count_px(const psx_draw_header* d, u32 n)
{
  o_stats.draw_px += n;
  if (d->flags0 & (PSX_DF_TRANSPARENCY | PSX_DF_CHECK_MASK))
    o_stats.draw_px_rmw += n;
  if (d->flags0 & PSX_DF_TEXTURE)
    o_stats.draw_px_tex += n;
}

/* One vector group of up to 4 lanes at x0..x0+3 on VRAM row y: all texel/CLUT fetches first,
 * then the stores (SIMD ShadePixel, inl:524-736; Q1). `active` is a 4-bit lane mask. */
static void shade_group4(const psx_draw_header* d, int mode, s32 x0, u32 y, u32 active, const u8* r, const u8* g,
                         const u8* b, const u8* u, const u8* v)
{
  u16 tex[4] = {0, 0, 0, 0};
  if (mode != MOD_DISABLED)
    for (int l = 0; l < 4; l++)
      if (active & (1u << l))
        tex[l] = fetch_texel(d, u[l], v[l]);
  for (int l = 0; l < 4; l++) {
    if (!(active & (1u << l)))
      continue;
    u16 c;
    count_px(d, 1);
    if (!shade_colour(d, mode, (u32)(x0 + l), y, r[l], g[l], b[l], tex[l], &c))
      continue;
    blend_store(d, mode != MOD_DISABLED, (u32)(x0 + l), y, c);
  }
}

/* ---- rectangles: DrawRectangle, inl:738-809 (vector form; scalar form 244-277 differs only in Q1) ---- */
static void draw_rectangle(const psx_draw_rectangle* c)
{
  const psx_draw_header* d = &c->d;
  const int mode = mod_mode(d->flags0);
  const u8 r = (u8)c->color, g = (u8)(c->color >> 8), b = (u8)(c->color >> 16);
  const u8 u0 = (u8)c->texcoord, v0 = (u8)(c->texcoord >> 8);
  const s32 left = (s32)o_area.left, right = (s32)o_area.right, top = (s32)o_area.top, bottom = (s32)o_area.bottom;
  o_stats.prims++;
  for (u32 oy = 0; oy < c->height; oy++) {
    const s32 y = c->y + (s32)oy;
    if (y < top || y > bottom)
      continue;
    if ((d->flags0 & PSX_DF_INTERLACED) && (((d->flags0 & PSX_DF_ACTIVE_LINE_LSB) != 0) == (((u32)y & 1u) != 0)))
      continue;
    const u32 dy = (u32)y & (VH - 1);
    const u8 tv = (u8)(v0 + oy);
    for (u32 ox = 0; ox < c->width; ox += 4) {
      u32 active = 0;
      u8 rr[4], gg[4], bb[4], uu[4], vv[4];
      for (u32 l = 0; l < 4; l++) {
        const s32 x = c->x + (s32)(ox + l);
        if (ox + l < c->width && x >= left && x <= right)
          active |= 1u << l;
        rr[l] = r;
        gg[l] = g;
        bb[l] = b;
        uu[l] = (u8)(u0 + ox + l);
        vv[l] = tv;
      }
      if (active)
        shade_group4(d, mode, c->x + (s32)ox, dy, active, rr, gg, bb, uu, vv);
    }
  }
}

/* ---- lines: DrawLine, inl:814-892 (scalar only in the reference) ---- */
static void draw_line(const psx_draw_header* d, const psx_line_vertex* p0, const psx_line_vertex* p1)
{
  const s32 adx = abs(p1->x - p0->x), ady = abs(p1->y - p0->y);
  const s32 k = (adx > ady) ? adx : ady;
  if (adx >= PSX_MAX_PRIM_WIDTH || ady >= PSX_MAX_PRIM_HEIGHT)
    return;
  if (p0->x >= p1->x && k > 0) {
    const psx_line_vertex* t = p0;
    p0 = p1;
    p1 = t;
  }
  o_stats.prims++;
  const int shaded = (d->flags0 & PSX_DF_SHADING) != 0;
  s64 dxdk = 0, dydk = 0;
  s32 drdk = 0, dgdk = 0, dbdk = 0;
  if (k != 0) {
    const s64 ddx = p1->x - p0->x, ddy = p1->y - p0->y;
    /* div_xy: ((delta << 32) - (delta<0 ? k-1 : 0) + (delta>0 ? k-1 : 0)) / k, truncating */
    dxdk = (s64)(((u64)ddx << 32) - (u64)((ddx < 0) ? (k - 1) : 0) + (u64)((ddx > 0) ? (k - 1) : 0)) / k;
    dydk = (s64)(((u64)ddy << 32) - (u64)((ddy < 0) ? (k - 1) : 0) + (u64)((ddy > 0) ? (k - 1) : 0)) / k;
    if (shaded) {
      drdk = (s32)((u32)((s32)p1->r - (s32)p0->r) << 12) / k;
      dgdk = (s32)((u32)((s32)p1->g - (s32)p0->g) << 12) / k;
      dbdk = (s32)((u32)((s32)p1->b - (s32)p0->b) << 12) / k;
    }
  }
  u64 curx = (((u64)(s64)p0->x << 32) | (1ull << 31)) - 1024u;
  u64 cury = (((u64)(s64)p0->y << 32) | (1ull << 31)) - ((dydk < 0) ? 1024u : 0u);
  u32 curr = ((u32)p0->r << 12) | (1u << 11), curg = ((u32)p0->g << 12) | (1u << 11), curb = ((u32)p0->b << 12) | (1u << 11);
  const s32 left = (s32)o_area.left, right = (s32)o_area.right, top = (s32)o_area.top, bottom = (s32)o_area.bottom;
  for (s32 i = 0; i <= k; i++) {
    const s32 x = sext11((s32)((s64)curx >> 32));
    const s32 y = sext11((s32)((s64)cury >> 32));
    const int skip = (d->flags0 & PSX_DF_INTERLACED) && (((d->flags0 & PSX_DF_ACTIVE_LINE_LSB) != 0) == (((u32)y & 1u) != 0));
    if (!skip && x >= left && x <= right && y >= top && y <= bottom) {
      const u8 r = shaded ? (u8)((s32)curr >> 12) : p0->r;
      const u8 g = shaded ? (u8)((s32)curg >> 12) : p0->g;
      const u8 b = shaded ? (u8)((s32)curb >> 12) : p0->b;
      u16 c;
      count_px(d, 1);
      shade_colour(d, MOD_DISABLED, (u32)x, (u32)y & (VH - 1), r, g, b, 0, &c);
      blend_store(d, 0, (u32)x, (u32)y & (VH - 1), c);
    }
    curx += (u64)dxdk;
    cury += (u64)dydk;
    curr += (u32)drdk;
    curg += (u32)dgdk;
    curb += (u32)dbdk;
  }
}

/* ---- triangles ---- */
typedef struct {
  u32 base[5]; /* r g b u v accumulators at (x=0, y=0), 8.24 with the +0.5 bias folded in */
  u32 ddx[5], ddy[5];
} tri_attrs;

typedef struct {
  u64 start_x[2], step_x[2];
  s32 start_y, end_y;
  int upside_down;
} tri_part;

/* DrawSpan, inl:1203-1312 (vector form: groups of 4 from the CLIPPED start). */
static void draw_span(const psx_draw_header* d, int mode, int shaded, s32 y, s32 x_start, s32 x_bound,
                      const u32* rowacc /* r g b u v at x=0 for this row */, const tri_attrs* a, const u8* flat)
{
  s32 width = x_bound - x_start;
  s32 cx = sext11(x_start);
  const s32 left = (s32)o_area.left, right = (s32)o_area.right;
  if (cx < left) {
    const s32 delta = left - cx;
    x_start += delta;
    cx += delta;
    width -= delta;
  }
  if ((cx + width) > (right + 1))
    width = right + 1 - cx;
  if (width <= 0)
    return;
  for (s32 gx = 0; gx < width; gx += 4) {
    u32 active = 0;
    u8 rr[4], gg[4], bb[4], uu[4], vv[4];
    for (s32 l = 0; l < 4; l++) {
      const u32 xr = (u32)(x_start + gx + l); /* raw x feeds the interpolants (inl:1232-1248) */
      if (gx + l < width)
        active |= 1u << l;
      if (shaded) {
        rr[l] = (u8)((rowacc[0] + a->ddx[0] * xr) >> 24);
        gg[l] = (u8)((rowacc[1] + a->ddx[1] * xr) >> 24);
        bb[l] = (u8)((rowacc[2] + a->ddx[2] * xr) >> 24);
      } else {
        rr[l] = flat[0];
        gg[l] = flat[1];
        bb[l] = flat[2];
      }
      uu[l] = (u8)((rowacc[3] + a->ddx[3] * xr) >> 24);
      vv[l] = (u8)((rowacc[4] + a->ddx[4] * xr) >> 24);
    }
    shade_group4(d, mode, cx + gx, (u32)y, active, rr, gg, bb, uu, vv);
  }
}

/* DrawTrianglePart, inl:1314-1413. */
static void draw_tri_part(const psx_draw_header* d, int mode, int shaded, const tri_part* tp, const tri_attrs* a,
                          const u8* flat)
{
  u64 lx = tp->start_x[0], rx = tp->start_x[1];
  s32 cy = tp->start_y;
  const s32 top = (s32)o_area.top, bottom = (s32)o_area.bottom;
  const int ilace = (d->flags0 & PSX_DF_INTERLACED) != 0;
  const u32 lsb = (d->flags0 & PSX_DF_ACTIVE_LINE_LSB) ? 1u : 0u;
  u32 row[5];
  if (tp->upside_down) {
    if (cy <= tp->end_y)
      return;
    do {
      cy--;
      lx -= tp->step_x[0];
      rx -= tp->step_x[1];
      const s32 y = sext11(cy);
      if (y < top)
        break;
      if (y > bottom || (ilace && lsb == ((u32)cy & 1u)))
        continue;
      for (int i = 0; i < 5; i++)
        row[i] = a->base[i] + a->ddy[i] * (u32)cy;
      draw_span(d, mode, shaded, y & (s32)(VH - 1), (s32)(lx >> 32), (s32)(rx >> 32), row, a, flat);
    } while (cy > tp->end_y);
  } else {
    if (cy >= tp->end_y)
      return;
    do {
      const s32 y = sext11(cy);
      if (y > bottom)
        break;
      if (y >= top && !(ilace && lsb == ((u32)cy & 1u))) {
        for (int i = 0; i < 5; i++)
          row[i] = a->base[i] + a->ddy[i] * (u32)cy;
        draw_span(d, mode, shaded, y & (s32)(VH - 1), (s32)(lx >> 32), (s32)(rx >> 32), row, a, flat);
      }
      cy++;
      lx += tp->step_x[0];
      rx += tp->step_x[1];
    } while (cy < tp->end_y);
  }
}

static inline s64 makefp_xy(s32 x) /* inl:1463 */
{
  return (s64)(((u64)(s64)x << 32) + ((1ull << 32) - (1u << 11)));
}
static inline s64 makestep_xy(s32 dx, s32 dy) /* inl:1464-1466 */
{
  const s64 adj = (dx < 0) ? -(s64)(dy - 1) : ((dx > 0) ? (s64)(dy - 1) : 0);
  return (s64)(((u64)(s64)dx << 32) + (u64)adj) / dy;
}

/* ATTRIB_STEP, inl:1495-1496: s32 products (wrapping), s32 truncating division, then << 12 as u32. */
static inline u32 attrib_step(s32 a0, s32 a1, s32 a2, s32 b0, s32 b1, s32 b2, s32 det)
{
  const u32 num = ((u32)(a1 - a0) * (u32)(b2 - b1)) - ((u32)(a2 - a1) * (u32)(b1 - b0));
  const s32 scaled = (s32)(num * 4096u);
  /* INT_MIN / -1 cannot occur: |det| >= 1 and |scaled| == 2^31 only if det divides … guarded anyway */
  const s32 q = (det == -1 && scaled == INT32_MIN) ? INT32_MIN : scaled / det;
  return (u32)q << 12;
}

/* DrawTriangle, inl:1417-1566. */
static void draw_triangle(const psx_draw_header* d, const psx_poly_vertex* v0, const psx_poly_vertex* v1,
                          const psx_poly_vertex* v2)
{
  const int mode = mod_mode(d->flags0);
  const int shaded = (d->flags0 & PSX_DF_SHADING) != 0;
  const psx_poly_vertex* t;
  u32 tl;
  if (v1->x <= v0->x)
    tl = (v2->x <= v1->x) ? 4 : 2;
  else if (v2->x < v0->x)
    tl = 4;
  else
    tl = 1;
  if (v2->y < v1->y) {
    t = v2, v2 = v1, v1 = t;
    tl = ((tl >> 1) & 2u) | ((tl << 1) & 4u) | (tl & 1u);
  }
  if (v1->y < v0->y) {
    t = v1, v1 = v0, v0 = t;
    tl = ((tl >> 1) & 1u) | ((tl << 1) & 2u) | (tl & 4u);
  }
  if (v2->y < v1->y) {
    t = v2, v2 = v1, v1 = t;
    tl = ((tl >> 1) & 2u) | ((tl << 1) & 4u) | (tl & 1u);
  }
  const psx_poly_vertex* vs[3] = {v0, v1, v2};
  tl >>= 1;
  if (v0->y == v2->y)
    return;

  const s64 base_coord = makefp_xy(v0->x);
  const s64 base_step = makestep_xy(v2->x - v0->x, v2->y - v0->y);
  const s64 bound_us = (v1->y == v0->y) ? 0 : makestep_xy(v1->x - v0->x, v1->y - v0->y);
  const s64 bound_ls = (v2->y == v1->y) ? 0 : makestep_xy(v2->x - v1->x, v2->y - v1->y);
  const u32 vo = (tl != 0) ? 1u : 0u;
  const u32 vp = (tl == 2) ? 3u : 0u;
  const int right_facing = (v1->y == v0->y) ? (v1->x > v0->x) : (bound_us > base_step);
  const u32 rfi = right_facing ? 1u : 0u, ofi = right_facing ? 0u : 1u;

  tri_part parts[2];
  tri_part* tpo = &parts[vo];
  tri_part* tpp = &parts[vo ^ 1u];
  tpo->start_y = vs[0 ^ vo]->y;
  tpo->end_y = vs[1 ^ vo]->y;
  tpp->start_y = vs[1 ^ vp]->y;
  tpp->end_y = vs[2 ^ vp]->y;
  tpo->start_x[rfi] = (u64)makefp_xy(vs[0 ^ vo]->x);
  tpo->step_x[rfi] = (u64)bound_us;
  tpo->start_x[ofi] = (u64)base_coord + (u64)(s64)(vs[vo]->y - vs[0]->y) * (u64)base_step;
  tpo->step_x[ofi] = (u64)base_step;
  tpo->upside_down = (int)vo;
  tpp->start_x[rfi] = (u64)makefp_xy(vs[1 ^ vp]->x);
  tpp->step_x[rfi] = (u64)bound_ls;
  tpp->start_x[ofi] = (u64)base_coord + (u64)(s64)(vs[1 ^ vp]->y - vs[0]->y) * (u64)base_step;
  tpp->step_x[ofi] = (u64)base_step;
  tpp->upside_down = (vp != 0);

  const s32 det = (s32)(((u32)(v1->x - v0->x) * (u32)(v2->y - v1->y)) - ((u32)(v2->x - v1->x) * (u32)(v1->y - v0->y)));
  if (det == 0)
    return;
  o_stats.prims++;

  tri_attrs a;
  memset(&a, 0, sizeof(a));
  const psx_poly_vertex* tv = vs[tl];
  const s32 at[5][3] = {{v0->r, v1->r, v2->r}, {v0->g, v1->g, v2->g}, {v0->b, v1->b, v2->b},
                        {v0->u, v1->u, v2->u}, {v0->v, v1->v, v2->v}};
  const u32 start[5] = {tv->r, tv->g, tv->b, tv->u, tv->v};
  for (int i = 0; i < 5; i++) {
    const int on = (i < 3) ? shaded : (mode != MOD_DISABLED);
    a.base[i] = ((start[i] << 12) + (1u << 11)) << 12;
    if (!on)
      continue;
    a.ddx[i] = attrib_step(at[i][0], at[i][1], at[i][2], v0->y, v1->y, v2->y, det);
    a.ddy[i] = attrib_step(v0->x, v1->x, v2->x, at[i][0], at[i][1], at[i][2], det);
    a.base[i] += a.ddx[i] * (u32)(-tv->x) + a.ddy[i] * (u32)(-tv->y);
  }
  const u8 flat[3] = {tv->r, tv->g, tv->b};
  draw_tri_part(d, mode, shaded, &parts[0], &a, flat);
  draw_tri_part(d, mode, shaded, &parts[1], &a, flat);
}

/* ---- transfers ---- */
/* FillVRAMImpl, inl:1639-1755; colour conversion gpu_helpers.h:32-39. */
static void fill_vram(u32 x, u32 y, u32 w, u32 h, u32 color, int interlaced, u32 active_lsb)
{
  const u16 c16 = (u16)(((color & 0xFFu) >> 3) | ((((color >> 8) & 0xFFu) >> 3) << 5) |
                        ((((color >> 16) & 0xFFu) >> 3) << 10) | (((color >> 24) & 1u) << 15));
  for (u32 i = 0; i < h; i++) {
    const u32 row = (y + i) % VH;
    if (interlaced && (row & 1u) == active_lsb)
      continue;
    for (u32 j = 0; j < w; j++)
      o_vram[row * VW + ((x + j) % VW)] = c16;
    o_stats.fill_px += w;
  }
}

/* WriteVRAMImpl, inl:1757-1831, vector build (Q2). */
static void write_vram(u32 x, u32 y, u32 w, u32 h, const u16* src, int set_mask, int check_mask)
{
  const u16 mand = check_mask ? 0x8000u : 0, mor = set_mask ? 0x8000u : 0;
  o_stats.upload_px += (u64)w * h;
  if (check_mask)
    o_stats.upload_px_masked += (u64)w * h;
  if ((x + w) <= VW && (y + h) <= VH && !set_mask && !check_mask) {
    for (u32 r = 0; r < h; r++)
      memcpy(&o_vram[(y + r) * VW + x], src + (size_t)r * w, (size_t)w * 2);
    return;
  }
  const u32 room = VW - x;
  const u32 aw = ((w < room) ? w : room) & ~7u;
  for (u32 r = 0; r < h; r++) {
    u16* drow = &o_vram[((y + r) % VH) * VW];
    u32 col = 0;
    for (; col < aw; col++) { /* vector body: always consumes a source pixel */
      u16* p = &drow[x + col];
      const u16 s = *(src++);
      if ((*p & mand) == 0)
        *p = (u16)(s | mor);
    }
    for (; col < w; col++) { /* scalar tail: consumes only when written */
      u16* p = &drow[(x + col) % VW];
      if ((*p & mand) == 0)
        *p = (u16)(*(src++) | mor);
    }
  }
}

/* CopyVRAMImpl, inl:1833-1939: wrap split (columns inside row blocks), then per row memmove + mask. */
static void copy_vram(u32 sx, u32 sy, u32 dx, u32 dy, u32 w, u32 h, int set_mask, int check_mask)
{
  if ((sx + w) > VW || (dx + w) > VW) {
    u32 rem_rows = h, csy = sy, cdy = dy;
    while (rem_rows > 0) {
      u32 rows = VH - csy;
      if (VH - cdy < rows)
        rows = VH - cdy;
      if (rem_rows < rows)
        rows = rem_rows;
      u32 rem_cols = w, csx = sx, cdx = dx;
      while (rem_cols > 0) {
        u32 cols = VW - csx;
        if (VW - cdx < cols)
          cols = VW - cdx;
        if (rem_cols < cols)
          cols = rem_cols;
        copy_vram(csx, csy, cdx, cdy, cols, rows, set_mask, check_mask);
        csx = (csx + cols) % VW;
        cdx = (cdx + cols) % VW;
        rem_cols -= cols;
      }
      csy = (csy + rows) % VH;
      cdy = (cdy + rows) % VH;
      rem_rows -= rows;
    }
    return;
  }
  const u16 mand = check_mask ? 0x8000u : 0, mor = set_mask ? 0x8000u : 0;
  static u16 rowbuf[VW];
  o_stats.copy_px += (u64)w * h;
  if (check_mask)
    o_stats.copy_px_masked += (u64)w * h;
  for (u32 r = 0; r < h; r++) {
    const u16* srow = &o_vram[((sy + r) % VH) * VW];
    u16* drow = &o_vram[((dy + r) % VH) * VW];
    for (u32 c = 0; c < w; c++)
      rowbuf[c] = srow[(sx + c) % VW];
    for (u32 c = 0; c < w; c++) {
      u16* p = &drow[(dx + c) % VW];
      if ((*p & mand) == 0)
        *p = (u16)(rowbuf[c] | mor);
    }
  }
}

/* UpdateCLUT, gpu_sw_rasterizer.cpp:43-66. */
static void update_clut(u16 reg, int is8)
{
  const u16* row = &o_vram[PSX_PAL_Y(reg) * VW];
  const u32 sx = PSX_PAL_X(reg);
  const u32 n = is8 ? 256u : 16u;
  for (u32 i = 0; i < n; i++)
    o_clut[i] = row[(sx + i) % VW]; /* 4-bit: sx+16 <= 1024 always, so % is a no-op there */
}

/* GPU::GetClampedDrawingArea, gpu.cpp:2190-2200. */
static void set_area(const psx_drawing_area* a)
{
  o_area = *a;
  if (a->left > a->right || a->top > a->bottom) {
    o_clamped[0] = o_clamped[1] = o_clamped[2] = o_clamped[3] = 0;
    return;
  }
  const u32 r1 = a->right + 1, b1 = a->bottom + 1;
  const u32 lmin = a->right < VW - 1 ? a->right : VW - 1, tmin = a->bottom < VH - 1 ? a->bottom : VH - 1;
  o_clamped[2] = (s32)(r1 < VW ? r1 : VW);
  o_clamped[0] = (s32)(a->left < lmin ? a->left : lmin);
  o_clamped[3] = (s32)(b1 < VH ? b1 : VH);
  o_clamped[1] = (s32)(a->top < tmin ? a->top : tmin);
}

/* GPUBackend::HandleCommand (gpu_backend.cpp:383-544) + GPU_SW::Draw* (gpu_sw.cpp:105-190). */
static void handle_command(const psx_cmd_header* h)
{
  o_stats.commands++;
  switch (h->type) {
  case PSX_CMD_CLEAR_VRAM: /* gpu_sw.cpp:61-65 */
    memset(o_vram, 0, sizeof(o_vram));
    memset(o_clut, 0, sizeof(o_clut));
    break;
  case PSX_CMD_LOAD_STATE: { /* gpu_sw.cpp:67-71 */
    const psx_load_state* c = (const psx_load_state*)h;
    memcpy(o_vram, c->vram_data, sizeof(o_vram));
    memcpy(o_clut, c->clut_data, sizeof(o_clut));
  } break;
  case PSX_CMD_FILL_VRAM: {
    const psx_fill_vram* c = (const psx_fill_vram*)h;
    fill_vram(c->x, c->y, c->width, c->height, c->color, c->interlaced_rendering != 0, c->active_line_lsb);
  } break;
  case PSX_CMD_UPDATE_VRAM: {
    const psx_update_vram* c = (const psx_update_vram*)h;
    write_vram(c->x, c->y, c->width, c->height, (const u16*)((const u8*)h + PSX_UPDATE_VRAM_DATA_OFFSET),
               c->set_mask_while_drawing != 0, c->check_mask_before_draw != 0);
  } break;
  case PSX_CMD_COPY_VRAM: {
    const psx_copy_vram* c = (const psx_copy_vram*)h;
    copy_vram(c->src_x, c->src_y, c->dst_x, c->dst_y, c->width, c->height, c->set_mask_while_drawing != 0,
              c->check_mask_before_draw != 0);
  } break;
  case PSX_CMD_SET_DRAWING_AREA:
    set_area(&((const psx_set_drawing_area*)h)->new_area);
    break;
  case PSX_CMD_UPDATE_CLUT: {
    const psx_update_clut* c = (const psx_update_clut*)h;
    update_clut(c->reg, c->clut_is_8bit != 0);
  } break;
  case PSX_CMD_DRAW_POLYGON: {
    const psx_draw_polygon* c = (const psx_draw_polygon*)h;
    draw_triangle(&c->d, &c->vertices[0], &c->vertices[1], &c->vertices[2]);
    if (c->d.num_vertices > 3)
      draw_triangle(&c->d, &c->vertices[2], &c->vertices[1], &c->vertices[3]);
  } break;
  case PSX_CMD_DRAW_PRECISE_POLYGON: {
    const psx_draw_precise_polygon* c = (const psx_draw_precise_polygon*)h;
    psx_poly_vertex v[4];
    memset(v, 0, sizeof(v));
    for (u32 i = 0; i < c->d.num_vertices && i < 4; i++) {
      v[i].x = c->vertices[i].native_x;
      v[i].y = c->vertices[i].native_y;
      v[i].r = (u8)c->vertices[i].color;
      v[i].g = (u8)(c->vertices[i].color >> 8);
      v[i].b = (u8)(c->vertices[i].color >> 16);
      v[i].u = (u8)c->vertices[i].texcoord;
      v[i].v = (u8)(c->vertices[i].texcoord >> 8);
    }
    draw_triangle(&c->d, &v[0], &v[1], &v[2]);
    if (c->d.num_vertices > 3)
      draw_triangle(&c->d, &v[2], &v[1], &v[3]);
  } break;
  case PSX_CMD_DRAW_RECTANGLE: { /* cull: gpu_sw.cpp:141-153 */
    const psx_draw_rectangle* c = (const psx_draw_rectangle*)h;
    const s32 l = o_clamped[0] > c->x ? o_clamped[0] : c->x, t = o_clamped[1] > c->y ? o_clamped[1] : c->y;
    const s32 rx = c->x + (s32)c->width, by = c->y + (s32)c->height;
    const s32 r = o_clamped[2] < rx ? o_clamped[2] : rx, b = o_clamped[3] < by ? o_clamped[3] : by;
    if (r > l && b > t)
      draw_rectangle(c);
  } break;
  case PSX_CMD_DRAW_LINE: {
    const psx_draw_line* c = (const psx_draw_line*)h;
    for (u32 i = 0; i + 1 < c->d.num_vertices; i += 2)
      draw_line(&c->d, &c->vertices[i], &c->vertices[i + 1]);
  } break;
  case PSX_CMD_DRAW_PRECISE_LINE: {
    const psx_draw_header* d = (const psx_draw_header*)h;
    const psx_precise_line_vertex* pv = (const psx_precise_line_vertex*)((const u8*)h + sizeof(psx_draw_header));
    for (u32 i = 0; i + 1 < d->num_vertices; i += 2) {
      psx_line_vertex v[2];
      for (int j = 0; j < 2; j++) {
        v[j].x = pv[i + j].native_x;
        v[j].y = pv[i + j].native_y;
        v[j].r = (u8)pv[i + j].color;
        v[j].g = (u8)(pv[i + j].color >> 8);
        v[j].b = (u8)(pv[i + j].color >> 16);
        v[j].a = 0;
      }
      draw_line(d, &v[0], &v[1]);
    }
  } break;
  default:
    break; /* ReadVRAM, ClearCache, BufferSwapped, UpdateDisplay, SubmitFrame: no VRAM effect (gpu_sw.cpp:86-88,192-211) */
  }
}

/* ---- scan-out: GPU_SW::UpdateDisplay/CopyOut15Bit/CopyOut24Bit, gpu_sw.cpp:230-356, 391-448 ---- */
static inline u32 e5to8(u32 c)
{
  return (c * 527u + 23u) >> 6; /* gpu_helpers.h:18-31 */
}
static inline void convert_px(int fmt, u8** d, u16 c) /* ConvertVRAMPixel, gpu_helpers.h:107-134 */
{
  if (fmt == 0) {
    const u32 v = e5to8(c & 31u) | (e5to8((c >> 5) & 31u) << 8) | (e5to8((c >> 10) & 31u) << 16) | ((c >> 15) ? 0xFF000000u : 0u);
    memcpy(*d, &v, 4);
    *d += 4;
    return;
  }
  u16 o;
  if (fmt == 1)
    o = (u16)((c & 0x83E0u) | ((c >> 10) & 0x1Fu) | ((c & 0x1Fu) << 10));
  else if (fmt == 2)
    o = (u16)(((c & 0x3E0u) << 1) | ((c >> 9) & 0x3Eu) | ((c & 0x1Fu) << 11) | (c >> 15));
  else
    o = (u16)(((c & 0x3E0u) << 1) | ((c & 0x20u) << 1) | ((c >> 10) & 0x1Fu) | ((c & 0x1Fu) << 11));
  memcpy(*d, &o, 2);
  *d += 2;
}

/* ---- exported API (mirrors psxref_* in oracle/ref_build/ref_harness.cpp) ---- */
void psxo_init(void)
{
}
void psxo_set_modulation_crop(int on)
{
  o_modulation_crop = on != 0;
}
void psxo_reset(void)
{
  memset(o_vram, 0, sizeof(o_vram));
  memset(o_clut, 0, sizeof(o_clut));
  memset(&o_stats, 0, sizeof(o_stats));
  const psx_drawing_area z = {0, 0, 0, 0};
  set_area(&z);
}
u16* psxo_vram(void)
{
  return o_vram;
}
u16* psxo_clut(void)
{
  return o_clut;
}
void psxo_get_stats(psxo_stats* out)
{
  *out = o_stats;
}
long psxo_exec(const void* cmds, size_t bytes)
{
  const u8* p = (const u8*)cmds;
  const u8* end = p + bytes;
  long n = 0;
  while (p < end) {
    const psx_cmd_header* h = (const psx_cmd_header*)p;
    if (h->size < sizeof(psx_cmd_header) || (h->size & 3u) != 0 || p + h->size > end)
      return -1;
    handle_command(h);
    p += h->size;
    n++;
  }
  return n;
}

int psxo_scanout(const psx_update_display* ud, int fmt, int show_vram, void* dst, u32 dst_stride, u32* out_w, u32* out_h)
{
  u32 src_x, src_y, skip_x, width, height, line_skip;
  int is24;
  if (show_vram) { /* gpu_sw.cpp:443-447 */
    src_x = src_y = skip_x = line_skip = 0;
    width = VW;
    height = VH;
    is24 = 0;
  } else {
    if (ud->flags & PSX_UD_DISPLAY_DISABLED)
      return 1;
    is24 = (ud->flags & PSX_UD_DISPLAY_24BIT) != 0;
    const u32 inter = (ud->flags & PSX_UD_INTERLACED_INTERLEAVED) ? 1u : 0u;
    const u32 field = (ud->flags & PSX_UD_INTERLACED_FIELD) ? 1u : 0u;
    src_x = is24 ? ud->X : ud->display_vram_left;
    skip_x = is24 ? (u32)(ud->display_vram_left - ud->X) : 0;
    src_y = ud->display_vram_top + (field & inter);
    width = ud->display_vram_width;
    height = ud->display_vram_height;
    line_skip = (ud->flags & PSX_UD_INTERLACED_ENABLED) ? inter : 0; /* non-interlaced passes 0, gpu_sw.cpp:435 */
  }
  *out_w = width;
  *out_h = height;
  u8* drow = (u8*)dst;
  if (!is24) {
    if (fmt < 0 || fmt > 3)
      return -1;
    if ((src_x + width) <= VW && (src_y + height) <= VH) { /* fast path, gpu_sw.cpp:241-268 */
      const u16* sp = &o_vram[src_y * VW + src_x];
      for (u32 r = 0; r < height; r++) {
        u8* d = drow;
        for (u32 x = 0; x < width; x++)
          convert_px(fmt, &d, sp[x]);
        sp += (VW << line_skip);
        drow += dst_stride;
      }
    } else { /* wrap path, gpu_sw.cpp:270-285 */
      for (u32 r = 0; r < height; r++) {
        const u16* srow = &o_vram[(src_y % VH) * VW];
        u8* d = drow;
        for (u32 c = src_x; c < src_x + width; c++)
          convert_px(fmt, &d, srow[c % VW]);
        src_y += (1u << line_skip);
        drow += dst_stride;
      }
    }
    return 0;
  }
  if ((src_x + width) <= VW && (src_y + (height << line_skip)) <= VH) { /* gpu_sw.cpp:305-324 */
    const u8* sp = (const u8*)&o_vram[src_y * VW + src_x] + skip_x * 3u;
    for (u32 r = 0; r < height; r++) {
      const u8* s = sp;
      u8* d = drow;
      for (u32 c = 0; c < width; c++) {
        *d++ = *s++;
        *d++ = *s++;
        *d++ = *s++;
        *d++ = 0xFF;
      }
      sp += (VW << line_skip) * 2u;
      drow += dst_stride;
    }
  } else { /* gpu_sw.cpp:325-348 */
    for (u32 r = 0; r < height; r++) {
      const u16* srow = &o_vram[(src_y % VH) * VW];
      u32* d = (u32*)drow;
      for (u32 c = 0; c < width; c++) {
        const u32 off = src_x + (((skip_x + c) * 3u) / 2u);
        const u32 s0 = srow[off % VW], s1 = srow[(off + 1) % VW];
        *d++ = (((s1 << 16) | s0) >> ((c & 1u) * 8u)) | 0xFF000000u;
      }
      src_y += (1u << line_skip);
      drow += dst_stride;
    }
  }
  return 0;
}

/* vramhash64: the parallel-friendly per-instance hash the CUDA side also computes
 * (duckstation_b200/csrc; SURVEY appendix A "Hashing").  Sum over 32-bit words i of
 * mix64((i << 32) | word_i); commutative, so any reduction order gives the same value. */
static inline u64 mix64(u64 z)
{
  z ^= z >> 30;
  z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27;
  z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}
u64 psxo_vramhash64(const u16* vram)
{
  const u32* w = (const u32*)vram;
  u64 h = 0;
  for (u32 i = 0; i < VW * VH / 2; i++)
    h += mix64(((u64)i << 32) | w[i]);
  return h;
}
