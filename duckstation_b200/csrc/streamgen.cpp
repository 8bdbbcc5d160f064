// streamgen.cpp — seeded synthetic backend-command streams for the five BASELINE.json configs.
//
// Emits the wire format of include/psxgpu_wire.h (layout of the reference's
// src/core/video_thread_commands.h), in the order the reference's GP0 decoder would
// push it: [UpdateCLUT] (only when the CLUT register / depth changes, gpu.cpp:2157-2176),
// [SetDrawingArea] (only when changed, gpu.cpp:2956-2965), then the draw.  Flag rules
// follow GPU::FillDrawCommand (gpu.cpp:2968-2996): no texture on lines, raw only with
// texture, dither = reg bit for lines / off for rectangles / reg && (shaded || modulated
// texture) for polygons.  Size limits follow the decoder's culls (gpu.cpp:3283-3350,
// :3424-3425, :3527-3535).
//
// The same streams feed the CPU oracle, the compiled reference and the CUDA backend, so
// parity is always compared on identical bytes.  Workload definitions: SURVEY.md §8(d).
#include "../../include/psxgpu_wire.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct Rng
{
  uint32_t s;
  explicit Rng(uint32_t seed) : s(seed * 2654435761u + 0x9E3779B9u)
  {
    if (s == 0)
      s = 1;
    for (int i = 0; i < 4; i++)
      next();
  }
  uint32_t next()
  {
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
  }
  uint32_t below(uint32_t n) { return n ? static_cast<uint32_t>((static_cast<uint64_t>(next()) * n) >> 32) : 0; }
  int range(int lo, int hi) { return lo + static_cast<int>(below(static_cast<uint32_t>(hi - lo + 1))); }
  bool chance(uint32_t pct) { return below(100) < pct; }
};

struct Builder
{
  std::vector<uint8_t> out;
  psx_drawing_area area{0xFFFFFFFFu, 0, 0, 0};
  psx_drawing_area sent_area{0xFFFFFFFFu, 0, 0, 0};
  uint32_t clut_reg = 0xFFFFFFFFu; // GPU::InvalidateCLUT state
  bool clut_is_8bit = false;
  uint16_t draw_mode = 0;
  psx_texture_window window{0xFF, 0xFF, 0, 0};
  bool set_mask = false, check_mask = false;
  bool interlaced = false, active_lsb = false;

  template<typename T>
  T* emit(size_t bytes, uint8_t type)
  {
    const size_t sz = (bytes + 15u) & ~static_cast<size_t>(15);
    const size_t at = out.size();
    out.resize(at + sz, 0);
    T* p = reinterpret_cast<T*>(out.data() + at);
    reinterpret_cast<psx_cmd_header*>(p)->size = static_cast<uint32_t>(sz);
    reinterpret_cast<psx_cmd_header*>(p)->type = type;
    return p;
  }

  void set_area(uint32_t l, uint32_t t, uint32_t r, uint32_t b) { area = psx_drawing_area{l, t, r, b}; }

  void prepare_for_draw()
  {
    if (std::memcmp(&area, &sent_area, sizeof(area)) != 0)
    {
      auto* c = emit<psx_set_drawing_area>(sizeof(psx_set_drawing_area), PSX_CMD_SET_DRAWING_AREA);
      c->new_area = area;
      sent_area = area;
    }
  }

  void update_clut_if_needed(uint16_t reg)
  {
    const uint32_t mode = PSX_DM_TEXMODE(draw_mode);
    if (mode >= 2)
      return;
    const bool needs8 = (mode == 1);
    if (reg != clut_reg || (needs8 && !clut_is_8bit))
    {
      clut_reg = reg;
      clut_is_8bit = needs8;
      auto* c = emit<psx_update_clut>(sizeof(psx_update_clut), PSX_CMD_UPDATE_CLUT);
      c->reg = reg;
      c->clut_is_8bit = needs8;
    }
  }

  void fill_header(psx_draw_header* d, int prim /*0 poly 1 rect 2 line*/, bool textured, bool raw, bool transp, bool shaded,
                   bool quad, uint16_t palette)
  {
    const bool tex = (prim != 2) && textured;
    const bool rawt = tex && raw;
    uint8_t f0 = 0;
    if (interlaced)
      f0 |= PSX_DF_INTERLACED;
    if (active_lsb)
      f0 |= PSX_DF_ACTIVE_LINE_LSB;
    if (set_mask)
      f0 |= PSX_DF_SET_MASK;
    if (check_mask)
      f0 |= PSX_DF_CHECK_MASK;
    if (tex)
      f0 |= PSX_DF_TEXTURE;
    if (rawt)
      f0 |= PSX_DF_RAW_TEXTURE;
    if (transp)
      f0 |= PSX_DF_TRANSPARENCY;
    if (shaded)
      f0 |= PSX_DF_SHADING;
    const bool dreg = PSX_DM_DITHER(draw_mode) != 0;
    bool dither;
    if (prim == 2)
      dither = dreg;
    else if (prim == 1)
      dither = false;
    else
      dither = dreg && (shaded || (tex && !rawt));
    d->flags0 = f0;
    d->flags1 = static_cast<uint8_t>((quad ? PSX_DF1_QUAD : 0) | (dither ? PSX_DF1_DITHER : 0));
    d->draw_mode = draw_mode;
    d->palette = palette;
    d->window = window;
  }

  // returns false if culled like the decoder would
  bool polygon(int nv, const psx_poly_vertex* v, bool textured, bool raw, bool transp, bool shaded, uint16_t palette)
  {
    auto too_big = [](const psx_poly_vertex& a, const psx_poly_vertex& b, const psx_poly_vertex& c) {
      const int minx = std::min({a.x, b.x, c.x}), maxx = std::max({a.x, b.x, c.x});
      const int miny = std::min({a.y, b.y, c.y}), maxy = std::max({a.y, b.y, c.y});
      return (maxx - minx + 1) > PSX_MAX_PRIM_WIDTH || (maxy - miny + 1) > PSX_MAX_PRIM_HEIGHT;
    };
    psx_poly_vertex vv[4];
    std::memcpy(vv, v, sizeof(psx_poly_vertex) * nv);
    const bool first_culled = too_big(vv[0], vv[1], vv[2]);
    int n = nv;
    if (nv == 3)
    {
      if (first_culled)
        return false;
    }
    else
    {
      const bool second_culled = too_big(vv[2], vv[1], vv[3]);
      if (first_culled && second_culled)
        return false;
      if (second_culled)
        n = 3;
      else if (first_culled)
      {
        vv[0] = vv[2];
        vv[2] = vv[3];
        n = 3;
      }
    }
    if (textured)
      update_clut_if_needed(palette);
    prepare_for_draw();
    auto* c = emit<psx_draw_polygon>(sizeof(psx_draw_header) + sizeof(psx_poly_vertex) * n, PSX_CMD_DRAW_POLYGON);
    fill_header(&c->d, 0, textured, raw, transp, shaded, nv == 4, palette);
    c->d.num_vertices = static_cast<uint16_t>(n);
    for (int i = 0; i < n; i++)
    {
      c->vertices[i] = vv[i];
      if (!shaded)
      {
        c->vertices[i].r = vv[0].r;
        c->vertices[i].g = vv[0].g;
        c->vertices[i].b = vv[0].b;
      }
      c->vertices[i].a = 0;
      if (!textured)
        c->vertices[i].u = c->vertices[i].v = 0;
    }
    return true;
  }

  void rectangle(int x, int y, int w, int h, uint32_t color, uint8_t u, uint8_t v, bool textured, bool raw, bool transp,
                 uint16_t palette)
  {
    if (textured)
      update_clut_if_needed(palette);
    prepare_for_draw();
    auto* c = emit<psx_draw_rectangle>(sizeof(psx_draw_rectangle), PSX_CMD_DRAW_RECTANGLE);
    fill_header(&c->d, 1, textured, raw, transp, false, false, textured ? palette : 0);
    c->d.num_vertices = 1;
    c->x = (static_cast<int32_t>(static_cast<uint32_t>(x) << 21)) >> 21;
    c->y = (static_cast<int32_t>(static_cast<uint32_t>(y) << 21)) >> 21;
    c->width = static_cast<uint16_t>(w & 1023);
    c->height = static_cast<uint16_t>(h & 511);
    c->color = color & 0xFFFFFFu;
    c->texcoord = textured ? static_cast<uint16_t>(u | (v << 8)) : 0;
  }

  // vertex pairs; each pair culled like gpu.cpp:3527-3535 / 3663-3668
  void lines(int npairs, const psx_line_vertex* v, bool transp, bool shaded)
  {
    std::vector<psx_line_vertex> keep;
    for (int i = 0; i < npairs; i++)
    {
      const auto& a = v[2 * i];
      const auto& b = v[2 * i + 1];
      const int w = std::abs(a.x - b.x) + 1, h = std::abs(a.y - b.y) + 1;
      if (w > PSX_MAX_PRIM_WIDTH || h > PSX_MAX_PRIM_HEIGHT)
        continue;
      keep.push_back(a);
      keep.push_back(b);
    }
    if (keep.empty())
      return;
    prepare_for_draw();
    auto* c = emit<psx_draw_line>(sizeof(psx_draw_header) + sizeof(psx_line_vertex) * keep.size(), PSX_CMD_DRAW_LINE);
    fill_header(&c->d, 2, false, false, transp, shaded, false, 0);
    c->d.num_vertices = static_cast<uint16_t>(keep.size());
    for (size_t i = 0; i < keep.size(); i++)
    {
      c->vertices[i] = keep[i];
      c->vertices[i].a = 0;
      if (!shaded)
      {
        c->vertices[i].r = keep[i & ~static_cast<size_t>(1)].r;
        c->vertices[i].g = keep[i & ~static_cast<size_t>(1)].g;
        c->vertices[i].b = keep[i & ~static_cast<size_t>(1)].b;
      }
    }
  }

  void fill(int x, int y, int w, int h, uint32_t color)
  {
    // decoder masks: gpu.cpp:3702-3706
    x &= 0x3F0;
    y &= 0x1FF;
    w = ((w & 0x3FF) + 15) & ~15;
    h &= 0x1FF;
    if (w == 0 || h == 0)
      return;
    auto* c = emit<psx_fill_vram>(sizeof(psx_fill_vram), PSX_CMD_FILL_VRAM);
    c->x = static_cast<uint16_t>(x);
    c->y = static_cast<uint16_t>(y);
    c->width = static_cast<uint16_t>(w);
    c->height = static_cast<uint16_t>(h);
    c->color = color & 0xFFFFFFu;
    c->interlaced_rendering = interlaced;
    c->active_line_lsb = active_lsb;
  }

  void copy(int sx, int sy, int dx, int dy, int w, int h)
  {
    sx &= 1023;
    dx &= 1023;
    sy &= 511;
    dy &= 511;
    w = ((w - 1) & 1023) + 1;
    h = ((h - 1) & 511) + 1;
    if (sx == dx && sy == dy && !set_mask)
      return; // gpu.cpp:3871-3874
    auto* c = emit<psx_copy_vram>(sizeof(psx_copy_vram), PSX_CMD_COPY_VRAM);
    c->src_x = static_cast<uint16_t>(sx);
    c->src_y = static_cast<uint16_t>(sy);
    c->dst_x = static_cast<uint16_t>(dx);
    c->dst_y = static_cast<uint16_t>(dy);
    c->width = static_cast<uint16_t>(w);
    c->height = static_cast<uint16_t>(h);
    c->set_mask_while_drawing = set_mask;
    c->check_mask_before_draw = check_mask;
  }

  uint16_t* upload(int x, int y, int w, int h)
  {
    x &= 1023;
    y &= 511;
    auto* c = emit<psx_update_vram>(PSX_UPDATE_VRAM_DATA_OFFSET + static_cast<size_t>(w) * h * 2, PSX_CMD_UPDATE_VRAM);
    c->x = static_cast<uint16_t>(x);
    c->y = static_cast<uint16_t>(y);
    c->width = static_cast<uint16_t>(w);
    c->height = static_cast<uint16_t>(h);
    c->set_mask_while_drawing = set_mask;
    c->check_mask_before_draw = check_mask;
    return reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(c) + PSX_UPDATE_VRAM_DATA_OFFSET);
  }

  void update_display(int left, int top, int w, int h, bool is24, int X, bool interlaced_en = false, bool field = false,
                      bool interleaved = false)
  {
    auto* c = emit<psx_update_display>(sizeof(psx_update_display), PSX_CMD_UPDATE_DISPLAY);
    c->display_width = static_cast<uint16_t>(w);
    c->display_height = static_cast<uint16_t>(h);
    c->display_vram_left = static_cast<uint16_t>(left);
    c->display_vram_top = static_cast<uint16_t>(top);
    c->display_vram_width = static_cast<uint16_t>(w);
    c->display_vram_height = static_cast<uint16_t>(h);
    c->display_pixel_aspect_ratio = 1.0f;
    c->X = static_cast<uint16_t>(X);
    c->flags = static_cast<uint8_t>((interlaced_en ? PSX_UD_INTERLACED_ENABLED : 0) | (field ? PSX_UD_INTERLACED_FIELD : 0) |
                                    (interleaved ? PSX_UD_INTERLACED_INTERLEAVED : 0) | (is24 ? PSX_UD_DISPLAY_24BIT : 0) |
                                    PSX_UD_SUBMIT_FRAME);
  }
};

uint16_t noise_px(Rng& r)
{
  const uint32_t v = r.next();
  if ((v & 0xF0000u) == 0)
    return 0; // ~6% fully transparent texels
  return static_cast<uint16_t>(v);
}

void noise_upload(Builder& b, Rng& r, int x, int y, int w, int h)
{
  uint16_t* px = b.upload(x, y, w, h);
  for (int i = 0; i < w * h; i++)
    px[i] = noise_px(r);
}

uint16_t make_draw_mode(uint32_t page_x, uint32_t page_y, uint32_t transp_mode, uint32_t tex_mode, bool dither)
{
  return static_cast<uint16_t>((page_x & 15u) | ((page_y & 1u) << 4) | ((transp_mode & 3u) << 5) | ((tex_mode & 3u) << 7) |
                               (dither ? (1u << 9) : 0u));
}

psx_texture_window random_window(Rng& r)
{
  // GPU::SetTextureWindow, gpu.cpp:2221-2240
  const uint8_t mx = static_cast<uint8_t>(r.below(32)), my = static_cast<uint8_t>(r.below(32));
  const uint8_t ox = static_cast<uint8_t>(r.below(32)), oy = static_cast<uint8_t>(r.below(32));
  psx_texture_window w;
  w.and_x = static_cast<uint8_t>(~(mx * 8));
  w.and_y = static_cast<uint8_t>(~(my * 8));
  w.or_x = static_cast<uint8_t>((ox & mx) * 8u);
  w.or_y = static_cast<uint8_t>((oy & my) * 8u);
  return w;
}

struct TexLayout
{
  // texture atlas occupies [ax0, ax0+aw) x [ay0, ay0+ah); pages are chosen to stay inside it
  int ax0, ay0, aw, ah;
};

// Picks a texture page / CLUT inside the atlas for the given texture mode.
void pick_texture(Rng& r, const TexLayout& L, uint32_t tex_mode, uint32_t* page_x, uint32_t* page_y, uint16_t* palette)
{
  const int page_w = (tex_mode == 0) ? 64 : (tex_mode == 1 ? 128 : 256);
  const int first = L.ax0 / 64;
  const int npages = std::max(1, (L.aw - page_w) / 64 + 1);
  *page_x = static_cast<uint32_t>(first + static_cast<int>(r.below(static_cast<uint32_t>(npages))));
  *page_y = (L.ah > 256) ? r.below(2) : static_cast<uint32_t>(L.ay0 / 256);
  // CLUT rows live inside the atlas too (any noise row works as a palette); 8 distinct positions
  const int cx = L.ax0 + 16 * static_cast<int>(r.below(static_cast<uint32_t>(std::max(1, (L.aw - 256) / 16 + 1))));
  const int cy = L.ay0 + static_cast<int>(r.below(8)) * (L.ah / 8);
  *palette = static_cast<uint16_t>(((cx / 16) & 0x3F) | ((cy & 0x1FF) << 6));
}

struct MixParams
{
  int draw_x0, draw_y0, draw_w, draw_h; // where primitives land
  int tri_extent;                       // vertices = centre +- extent
  int sprite_min, sprite_max;
  uint32_t pct_tri;                     // remainder sprites
  uint32_t pct_quad;                    // of polygons
  bool textured_mix;                    // config-2 style state mix
};

// One primitive of the "config 2" mix: textured/untextured tris+sprites with the full state space.
void emit_mix_prim(Builder& b, Rng& r, const TexLayout& L, const MixParams& P, bool allow_self_sample)
{
  const bool textured = P.textured_mix ? r.chance(80) : false;
  const uint32_t tex_mode = r.below(3);
  const bool raw = textured && r.chance(25);
  const bool transp = P.textured_mix ? r.chance(50) : false;
  const uint32_t tmode = r.below(4);
  const bool dither_reg = r.chance(50);
  uint32_t px = 0, py = 0;
  uint16_t pal = 0;
  pick_texture(r, L, tex_mode, &px, &py, &pal);
  if (allow_self_sample)
  {
    // texture page overlapping the draw area (Q1)
    px = static_cast<uint32_t>(P.draw_x0 / 64) + r.below(static_cast<uint32_t>(std::max(1, P.draw_w / 64)));
    py = static_cast<uint32_t>(P.draw_y0 / 256);
  }
  b.draw_mode = make_draw_mode(px, py, tmode, tex_mode, dither_reg);
  if (P.textured_mix)
  {
    b.window = r.chance(25) ? random_window(r) : psx_texture_window{0xFF, 0xFF, 0, 0};
    b.set_mask = r.chance(25);
    b.check_mask = r.chance(25);
  }
  if (r.chance(P.pct_tri))
  {
    const bool shaded = r.chance(50);
    const bool quad = r.chance(P.pct_quad);
    const int cx = P.draw_x0 + static_cast<int>(r.below(static_cast<uint32_t>(P.draw_w)));
    const int cy = P.draw_y0 + static_cast<int>(r.below(static_cast<uint32_t>(P.draw_h)));
    psx_poly_vertex v[4];
    std::memset(v, 0, sizeof(v));
    for (int i = 0; i < 4; i++)
    {
      v[i].x = cx + r.range(-P.tri_extent, P.tri_extent);
      v[i].y = cy + r.range(-P.tri_extent, P.tri_extent);
      const uint32_t c = r.next();
      v[i].r = static_cast<uint8_t>(c);
      v[i].g = static_cast<uint8_t>(c >> 8);
      v[i].b = static_cast<uint8_t>(c >> 16);
      const uint32_t t = r.next();
      v[i].u = static_cast<uint8_t>(t);
      v[i].v = static_cast<uint8_t>(t >> 8);
    }
    b.polygon(quad ? 4 : 3, v, textured, raw, transp, shaded, pal);
  }
  else
  {
    const int w = r.range(P.sprite_min, P.sprite_max), h = r.range(P.sprite_min, P.sprite_max);
    const int x = P.draw_x0 + r.range(-w / 2, P.draw_w - w / 2), y = P.draw_y0 + r.range(-h / 2, P.draw_h - h / 2);
    const uint32_t t = r.next();
    b.rectangle(x, y, w, h, r.next(), static_cast<uint8_t>(t), static_cast<uint8_t>(t >> 8), textured, raw, transp, pal);
  }
}

// ---- config 1: 20k flat+Gouraud untextured opaque triangles over the whole VRAM ----
void gen_config1(Builder& b, Rng& r, uint32_t n)
{
  b.set_area(0, 0, 1023, 511);
  for (uint32_t i = 0; i < n; i++)
  {
    const bool shaded = r.chance(50);
    b.draw_mode = make_draw_mode(0, 0, 0, 0, shaded);
    const int cx = static_cast<int>(r.below(1024)), cy = static_cast<int>(r.below(512));
    psx_poly_vertex v[3];
    std::memset(v, 0, sizeof(v));
    for (auto& p : v)
    {
      p.x = cx + r.range(-64, 64);
      p.y = cy + r.range(-64, 64);
      const uint32_t c = r.next();
      p.r = static_cast<uint8_t>(c);
      p.g = static_cast<uint8_t>(c >> 8);
      p.b = static_cast<uint8_t>(c >> 16);
    }
    b.polygon(3, v, false, false, false, shaded, 0);
  }
}

// ---- config 2: textured tris + sprites, atlas at x>=512, draw at x<512, 1% self-sampling slice ----
void gen_config2(Builder& b, Rng& r, uint32_t n)
{
  const TexLayout L{512, 0, 512, 512};
  for (int ty = 0; ty < 512; ty += 128) // atlas in four uploads (keeps single records < 256 KiB)
    noise_upload(b, r, 512, ty, 512, 128);
  b.set_area(0, 0, 511, 511);
  const MixParams P{0, 0, 512, 512, 64, 8, 64, 70, 20, true};
  for (uint32_t i = 0; i < n; i++)
    emit_mix_prim(b, r, L, P, r.chance(1));
}

// ---- config 3: mixed stream with transfers and render-to-texture hazards ----
void gen_config3(Builder& b, Rng& r, uint32_t n)
{
  const TexLayout L{512, 0, 512, 512};
  for (int ty = 0; ty < 512; ty += 128)
    noise_upload(b, r, 512, ty, 512, 128);
  b.set_area(0, 0, 639, 479);
  const MixParams P{0, 0, 640, 480, 48, 4, 48, 100, 25, true};
  for (uint32_t i = 0; i < n; i++)
  {
    if (i % 50 == 49)
    {
      // render-to-texture pair: draw untextured into a texture page, then sample that page with a fresh CLUT
      const uint32_t page = 8 + r.below(8);
      const int x0 = static_cast<int>(page) * 64;
      b.set_mask = b.check_mask = false;
      b.window = psx_texture_window{0xFF, 0xFF, 0, 0};
      b.set_area(static_cast<uint32_t>(x0), 0, static_cast<uint32_t>(x0 + 63), 255);
      b.draw_mode = make_draw_mode(0, 0, 0, 0, true);
      psx_poly_vertex v[3];
      std::memset(v, 0, sizeof(v));
      for (auto& p : v)
      {
        p.x = x0 + r.range(-16, 80);
        p.y = r.range(-16, 272);
        const uint32_t c = r.next();
        p.r = static_cast<uint8_t>(c);
        p.g = static_cast<uint8_t>(c >> 8);
        p.b = static_cast<uint8_t>(c >> 16);
      }
      b.polygon(3, v, false, false, false, true, 0);
      // CLUT row inside the page just drawn (forces UpdateCLUT to read fresh pixels)
      b.set_area(0, 0, 639, 479);
      const uint32_t tex_mode = r.below(3);
      b.draw_mode = make_draw_mode(page - (tex_mode == 2 && page > 12 ? 4 : (tex_mode == 1 && page > 14 ? 1 : 0)), 0,
                                   r.below(4), tex_mode, false);
      const uint16_t pal = static_cast<uint16_t>(((x0 / 16) & 0x3F) | (static_cast<uint32_t>(r.below(256)) << 6));
      b.clut_reg = 0xFFFFFFFFu; // GP0(01h) clear-cache style invalidation (gpu.cpp:2842-2851)
      b.clut_is_8bit = false;
      b.rectangle(r.range(0, 600), r.range(0, 440), r.range(8, 64), r.range(8, 64), r.next(),
                  static_cast<uint8_t>(r.next()), static_cast<uint8_t>(r.next()), true, r.chance(50), r.chance(50), pal);
      continue;
    }
    const uint32_t k = r.below(100);
    if (k < 40)
    {
      emit_mix_prim(b, r, L, P, false);
    }
    else if (k < 55)
    {
      MixParams Q = P;
      Q.pct_tri = 0;
      emit_mix_prim(b, r, L, Q, false);
    }
    else if (k < 70)
    {
      // line / polyline of 1..6 segments
      const int segs = r.chance(50) ? 1 : r.range(2, 6);
      const bool shaded = r.chance(50), transp = r.chance(30);
      b.draw_mode = make_draw_mode(0, 0, r.below(4), 0, r.chance(50));
      b.set_mask = r.chance(15);
      b.check_mask = r.chance(15);
      std::vector<psx_line_vertex> v;
      psx_line_vertex cur{};
      cur.x = r.range(-50, 700);
      cur.y = r.range(-50, 520);
      uint32_t c = r.next();
      cur.r = static_cast<uint8_t>(c);
      cur.g = static_cast<uint8_t>(c >> 8);
      cur.b = static_cast<uint8_t>(c >> 16);
      for (int s = 0; s < segs; s++)
      {
        psx_line_vertex nx{};
        const int ext = r.chance(10) ? 600 : 120;
        nx.x = cur.x + r.range(-ext, ext);
        nx.y = cur.y + r.range(-ext / 2, ext / 2);
        c = r.next();
        nx.r = static_cast<uint8_t>(c);
        nx.g = static_cast<uint8_t>(c >> 8);
        nx.b = static_cast<uint8_t>(c >> 16);
        v.push_back(cur);
        v.push_back(nx);
        cur = nx;
      }
      b.lines(segs, v.data(), transp, shaded);
    }
    else if (k < 80)
    {
      b.interlaced = r.chance(10);
      b.active_lsb = r.chance(50);
      b.fill(static_cast<int>(r.below(1024)), static_cast<int>(r.below(512)), r.range(1, r.chance(10) ? 1023 : 200),
             r.range(1, r.chance(10) ? 511 : 120), r.next());
      b.interlaced = false;
    }
    else if (k < 90)
    {
      b.set_mask = r.chance(25);
      b.check_mask = r.chance(25);
      const int w = r.range(1, r.chance(10) ? 700 : 96), h = r.range(1, r.chance(10) ? 400 : 96);
      const int sx = static_cast<int>(r.below(1024)), sy = static_cast<int>(r.below(512));
      int dx, dy;
      if (r.chance(30))
      { // overlapping
        dx = sx + r.range(-20, 20);
        dy = sy + r.range(-20, 20);
      }
      else
      {
        dx = static_cast<int>(r.below(1024));
        dy = static_cast<int>(r.below(512));
      }
      b.copy(sx, sy, dx, dy, w, h);
    }
    else
    {
      b.set_mask = r.chance(30);
      b.check_mask = r.chance(40);
      const int w = r.range(4, 64), h = r.range(4, 64);
      int x = static_cast<int>(r.below(1024)), y = static_cast<int>(r.below(512));
      if (r.chance(10))
      {
        x = 1024 - r.range(1, w);
        y = 512 - r.range(1, h);
      }
      uint16_t* px = b.upload(x, y, w, h);
      for (int i2 = 0; i2 < w * h; i2++)
        px[i2] = static_cast<uint16_t>(r.next()); // ~50% have bit15 -> later mask checks hit (Q2)
    }
  }
  b.set_mask = b.check_mask = false;
}

// ---- config 4: one batched instance: clear + 256x256 atlas upload + n prims of the config-2 mix at ~256 px ----
void gen_config4(Builder& b, Rng& r, uint32_t n)
{
  const TexLayout L{512, 0, 256, 256};
  b.fill(0, 0, 512, 511, r.next());
  noise_upload(b, r, 512, 0, 256, 256);
  b.set_area(0, 0, 511, 510); // rows 0..510 == the cleared rows, so a re-run starts from the same pixels
  const MixParams P{0, 0, 512, 511, 29, 8, 24, 70, 20, true};
  for (uint32_t i = 0; i < n; i++)
    emit_mix_prim(b, r, L, P, false);
}

// ---- config 6: config 4 with hazards — the same clear + atlas + mix, plus config 2's 1 % self-sampling slice (Q1) and a
// render-to-texture pair every 100 primitives (an untextured triangle drawn INTO a texture page of the atlas, then
// a sprite that samples that page through a palette fetched from the pixels just drawn): every pair forces a
// write-after-read and a read-after-write cut, so an instance runs in tens of epochs instead of two ----
void gen_config6(Builder& b, Rng& r, uint32_t n)
{
  const TexLayout L{512, 0, 256, 256};
  b.fill(0, 0, 512, 511, r.next());
  noise_upload(b, r, 512, 0, 256, 256);
  b.set_area(0, 0, 511, 510);
  const MixParams P{0, 0, 512, 511, 29, 8, 24, 70, 20, true};
  for (uint32_t i = 0; i < n; i++)
  {
    if (i % 100 == 99)
    {
      const uint32_t page = 8 + r.below(4); // the atlas covers pages 8..11 (x 512..767), rows 0..255
      const int x0 = static_cast<int>(page) * 64;
      b.set_mask = b.check_mask = false;
      b.window = psx_texture_window{0xFF, 0xFF, 0, 0};
      b.set_area(static_cast<uint32_t>(x0), 0, static_cast<uint32_t>(x0 + 63), 255);
      b.draw_mode = make_draw_mode(0, 0, 0, 0, true);
      psx_poly_vertex v[3];
      std::memset(v, 0, sizeof(v));
      const int cx = x0 + r.range(8, 56), cy = r.range(16, 240);
      for (auto& p : v)
      {
        p.x = cx + r.range(-29, 29);
        p.y = cy + r.range(-29, 29);
        const uint32_t c = r.next();
        p.r = static_cast<uint8_t>(c);
        p.g = static_cast<uint8_t>(c >> 8);
        p.b = static_cast<uint8_t>(c >> 16);
      }
      b.polygon(3, v, false, false, false, true, 0);
      b.set_area(0, 0, 511, 510);
      const uint32_t tex_mode = r.below(2); // 4- or 8-bit: the palette row lies inside the page just drawn
      b.draw_mode = make_draw_mode(tex_mode == 1 && page > 10 ? 10 : page, 0, r.below(4), tex_mode, false);
      const uint16_t pal = static_cast<uint16_t>(((x0 / 16) & 0x3F) | (static_cast<uint32_t>(r.below(256)) << 6));
      b.clut_reg = 0xFFFFFFFFu; // GP0(01h)-style invalidation: UpdateCLUT re-reads the fresh pixels
      b.clut_is_8bit = false;
      b.rectangle(r.range(0, 480), r.range(0, 480), r.range(8, 24), r.range(8, 24), r.next(), static_cast<uint8_t>(r.next()),
                  static_cast<uint8_t>(r.next()), true, r.chance(50), r.chance(50), pal);
      continue;
    }
    emit_mix_prim(b, r, L, P, r.chance(1));
  }
}

// ---- config 5: game-like frames; n = number of frames ----
void gen_config5(Builder& b, Rng& r, uint32_t frames, uint32_t prims_per_frame)
{
  const TexLayout L{640, 0, 384, 512};
  for (int ty = 0; ty < 512; ty += 128)
    noise_upload(b, r, 640, ty, 384, 128);
  for (uint32_t f = 0; f < frames; f++)
  {
    const int by = (f & 1) ? 256 : 0; // back buffer
    const bool fmv = (f % 10) == 9;
    b.set_mask = b.check_mask = false;
    b.window = psx_texture_window{0xFF, 0xFF, 0, 0};
    if (fmv)
    {
      // 24-bit 320x240 frame = 480 halfwords per row
      uint16_t* px = b.upload(0, by, 480, 240);
      for (int i = 0; i < 480 * 240; i++)
        px[i] = static_cast<uint16_t>(r.next());
      b.update_display(0, by, 320, 240, true, 0);
      continue;
    }
    b.fill(0, by, 320, 240, r.next());
    b.set_area(0, static_cast<uint32_t>(by), 319, static_cast<uint32_t>(by + 239));
    // strips of textured Gouraud triangles
    uint32_t emitted = 0;
    while (emitted < prims_per_frame * 8 / 10)
    {
      const uint32_t tex_mode = r.below(3);
      uint32_t px = 0, py = 0;
      uint16_t pal = 0;
      pick_texture(r, L, tex_mode, &px, &py, &pal);
      b.draw_mode = make_draw_mode(px, py, r.below(4), tex_mode, true);
      const bool transp = r.chance(20);
      const int len = r.range(4, 24);
      int x = r.range(-20, 320), y = by + r.range(-20, 240);
      psx_poly_vertex a{}, c{};
      auto rndv = [&](psx_poly_vertex& p, int px_, int py_) {
        p.x = px_;
        p.y = py_;
        const uint32_t col = r.next();
        p.r = static_cast<uint8_t>(col);
        p.g = static_cast<uint8_t>(col >> 8);
        p.b = static_cast<uint8_t>(col >> 16);
        const uint32_t t = r.next();
        p.u = static_cast<uint8_t>(t);
        p.v = static_cast<uint8_t>(t >> 8);
      };
      rndv(a, x, y);
      rndv(c, x + r.range(-12, 12), y + r.range(8, 24));
      for (int s = 0; s < len; s++)
      {
        psx_poly_vertex d{};
        x += r.range(6, 20);
        rndv(d, x, y + ((s & 1) ? r.range(8, 24) : r.range(-6, 6)));
        psx_poly_vertex v[3] = {a, c, d};
        b.polygon(3, v, true, false, transp, true, pal);
        a = c;
        c = d;
        emitted++;
      }
    }
    // HUD / text sprites
    for (uint32_t s = 0; s < prims_per_frame * 18 / 100; s++)
    {
      const uint32_t tex_mode = r.below(2);
      uint32_t px = 0, py = 0;
      uint16_t pal = 0;
      pick_texture(r, L, tex_mode, &px, &py, &pal);
      b.draw_mode = make_draw_mode(px, py, 0, tex_mode, false);
      const int sz = r.chance(70) ? 8 : 16;
      b.rectangle(r.range(0, 312), by + r.range(0, 232), sz, sz, 0x808080u, static_cast<uint8_t>(r.next() & 0xF8),
                  static_cast<uint8_t>(r.next() & 0xF8), true, r.chance(50), false, pal);
    }
    // a few lines
    for (uint32_t s = 0; s < prims_per_frame * 2 / 100 + 1; s++)
    {
      psx_line_vertex v[2]{};
      for (auto& p : v)
      {
        p.x = r.range(0, 319);
        p.y = by + r.range(0, 239);
        const uint32_t col = r.next();
        p.r = static_cast<uint8_t>(col);
        p.g = static_cast<uint8_t>(col >> 8);
        p.b = static_cast<uint8_t>(col >> 16);
      }
      b.draw_mode = make_draw_mode(0, 0, 0, 0, true);
      b.lines(1, v, false, r.chance(50));
    }
    b.update_display(0, by, 320, 240, false, 0);
  }
}

// ---- config 0: fuzz — everything random incl. out-of-range coordinates, odd drawing areas, interlace ----
void gen_fuzz(Builder& b, Rng& r, uint32_t n, uint32_t flags)
{
  const bool overlap_textures = (flags & 1u) != 0; // sample anywhere, incl. the pixels being drawn (Q1)
  const bool with_transfers = (flags & 2u) != 0;
  // seed VRAM with noise
  for (int ty = 0; ty < 512; ty += 128)
    noise_upload(b, r, 0, ty, 1024, 128);
  for (uint32_t i = 0; i < n; i++)
  {
    if (i % 64 == 0)
    {
      uint32_t l = r.below(1024), rr = r.below(1024), t = r.below(1024), bb = r.below(1024);
      if (r.chance(90))
      {
        if (l > rr)
          std::swap(l, rr);
        if (t > bb)
          std::swap(t, bb);
      }
      if (r.chance(50))
      {
        l = r.below(256);
        t = r.below(128);
        rr = 1023 - r.below(256);
        bb = 511 - r.below(128);
      }
      if (!overlap_textures)
      {
        // keep draws at x<512 and textures at x>=512
        l = std::min(l, 511u);
        rr = std::min(rr, 511u);
      }
      b.set_area(l, t, rr, bb);
      b.interlaced = r.chance(15);
      b.active_lsb = r.chance(50);
    }
    const uint32_t tex_mode = r.below(4);
    const uint32_t pagex = overlap_textures ? r.below(16) : (8 + r.below(tex_mode == 0 ? 8 : (tex_mode == 1 ? 7 : 5)));
    b.draw_mode = make_draw_mode(pagex, r.below(2), r.below(4), tex_mode, r.chance(50));
    b.window = r.chance(30) ? random_window(r) : psx_texture_window{0xFF, 0xFF, 0, 0};
    b.set_mask = r.chance(25);
    b.check_mask = r.chance(25);
    const bool textured = r.chance(60), raw = r.chance(30), transp = r.chance(50);
    uint16_t pal;
    if (overlap_textures)
      pal = static_cast<uint16_t>(r.next() & 0x7FFF);
    else
      pal = static_cast<uint16_t>((32 + r.below(tex_mode == 1 ? 17 : 32)) | (r.below(512) << 6));
    const uint32_t k = r.below(100);
    if (with_transfers && k >= 85)
    {
      const uint32_t t = r.below(3);
      if (t == 0)
        b.fill(static_cast<int>(r.next()), static_cast<int>(r.next()), static_cast<int>(r.below(1024)),
               static_cast<int>(r.below(512)), r.next() | (r.chance(20) ? 0x01000000u : 0));
      else if (t == 1)
      {
        const int w = r.chance(5) ? 1024 : r.range(1, 300), h = r.chance(5) ? 512 : r.range(1, 200);
        const int sx = static_cast<int>(r.below(1024)), sy = static_cast<int>(r.below(512));
        const bool ov = r.chance(40);
        b.copy(sx, sy, ov ? sx + r.range(-20, 20) : static_cast<int>(r.below(1024)),
               ov ? sy + r.range(-20, 20) : static_cast<int>(r.below(512)), w, h);
      }
      else
      {
        const int w = r.chance(3) ? 1024 : r.range(1, 100), h = r.chance(3) ? 512 : r.range(1, 80);
        uint16_t* px = b.upload(static_cast<int>(r.below(1024)), static_cast<int>(r.below(512)), w, h);
        for (int j = 0; j < w * h; j++)
          px[j] = static_cast<uint16_t>(r.next());
      }
      if (!overlap_textures)
        i += 0; // transfers may dirty texture pages: that is what hazard tests want
      continue;
    }
    if (k < 55)
    {
      const bool big = r.chance(10);
      const int ext = big ? 500 : (r.chance(50) ? 24 : 90);
      const bool wild = r.chance(8); // coordinates beyond the 11-bit range (sext11 wrap paths)
      const int cx = wild ? r.range(-2000, 2000) : r.range(-100, 1100), cy = wild ? r.range(-2000, 2000) : r.range(-100, 600);
      psx_poly_vertex v[4];
      std::memset(v, 0, sizeof(v));
      for (auto& p : v)
      {
        p.x = std::clamp(cx + r.range(-ext, ext), -2048, 2046);
        p.y = std::clamp(cy + r.range(-ext / 2, ext / 2), -2048, 2046);
        const uint32_t c = r.next();
        p.r = static_cast<uint8_t>(c);
        p.g = static_cast<uint8_t>(c >> 8);
        p.b = static_cast<uint8_t>(c >> 16);
        const uint32_t t = r.next();
        p.u = static_cast<uint8_t>(t);
        p.v = static_cast<uint8_t>(t >> 8);
      }
      if (r.chance(10))
        v[1].y = v[0].y; // flat top/bottom
      if (r.chance(10))
        v[2].x = v[0].x; // vertical edge
      if (r.chance(3))
        v[2] = v[1]; // degenerate
      b.polygon(r.chance(25) ? 4 : 3, v, textured, raw, transp, r.chance(50), pal);
    }
    else if (k < 75)
    {
      const int w = r.chance(5) ? r.range(0, 1023) : r.range(0, 80), h = r.chance(5) ? r.range(0, 511) : r.range(0, 80);
      b.rectangle(r.range(-1100, 1100), r.range(-1100, 1100), w, h, r.next(), static_cast<uint8_t>(r.next()),
                  static_cast<uint8_t>(r.next()), textured, raw, transp, pal);
    }
    else
    {
      const int segs = r.range(1, 4);
      std::vector<psx_line_vertex> v;
      psx_line_vertex cur{};
      const bool wild = r.chance(8);
      cur.x = wild ? r.range(-2000, 2000) : r.range(-50, 1074);
      cur.y = wild ? r.range(-2000, 2000) : r.range(-50, 562);
      for (int s = 0; s < segs; s++)
      {
        psx_line_vertex nx{};
        const int ext = r.chance(15) ? 1023 : 100;
        nx.x = std::clamp(cur.x + r.range(-ext, ext), -2048, 2046);
        nx.y = std::clamp(cur.y + r.range(-ext / 2, ext / 2), -2048, 2046);
        if (r.chance(10))
          nx.x = cur.x;
        if (r.chance(10))
          nx.y = cur.y;
        uint32_t c = r.next();
        cur.r = static_cast<uint8_t>(c);
        cur.g = static_cast<uint8_t>(c >> 8);
        cur.b = static_cast<uint8_t>(c >> 16);
        c = r.next();
        nx.r = static_cast<uint8_t>(c);
        nx.g = static_cast<uint8_t>(c >> 8);
        nx.b = static_cast<uint8_t>(c >> 16);
        v.push_back(cur);
        v.push_back(nx);
        cur = nx;
      }
      b.lines(segs, v.data(), transp, r.chance(50));
    }
  }
}

} // namespace

extern "C" {

// config: 0 fuzz, 1..5 the BASELINE.json configs (1-based), 6 = config 4 with hazards.  n: primitives / commands / frames (config 5).
// flags: config 0 -> bit0 textures may overlap the draw area (Q1), bit1 include transfers;
//        config 5 -> primitives per frame (0 = 1500).
uint8_t* psxgen_config(int config, uint32_t seed, uint32_t n, uint32_t flags, size_t* out_bytes)
{
  Builder b;
  Rng r(seed);
  switch (config)
  {
    case 0: gen_fuzz(b, r, n, flags); break;
    case 1: gen_config1(b, r, n); break;
    case 2: gen_config2(b, r, n); break;
    case 3: gen_config3(b, r, n); break;
    case 4: gen_config4(b, r, n); break;
    case 5: gen_config5(b, r, n, flags ? flags : 1500u); break;
    case 6: gen_config6(b, r, n); break;
    default: *out_bytes = 0; return nullptr;
  }
  uint8_t* p = static_cast<uint8_t*>(std::malloc(b.out.size() ? b.out.size() : 1));
  if (!p)
  {
    *out_bytes = 0;
    return nullptr;
  }
  std::memcpy(p, b.out.data(), b.out.size());
  *out_bytes = b.out.size();
  return p;
}

void psxgen_free(uint8_t* p)
{
  std::free(p);
}

} // extern "C"
