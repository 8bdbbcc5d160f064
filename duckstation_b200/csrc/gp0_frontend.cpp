// gp0_frontend.cpp — see gp0_frontend.h.
#include "gp0_frontend.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

namespace psx {

namespace {
inline int32_t sext11(uint32_t v)
{
  return static_cast<int32_t>(v << 21) >> 21;
}
inline int32_t vertex_x(uint32_t word) // GPUVertexPosition, gpu_types.h:199-205
{
  return sext11(word & 0x7FFu);
}
inline int32_t vertex_y(uint32_t word)
{
  return sext11((word >> 16) & 0x7FFu);
}
inline uint32_t replace_zero(uint32_t v, uint32_t z)
{
  return v == 0 ? z : v;
}
// GPURenderCommand bits, gpu_types.h:122-135
inline bool rc_raw(uint32_t rc) { return (rc >> 24) & 1u; }
inline bool rc_transparent(uint32_t rc) { return (rc >> 25) & 1u; }
inline bool rc_textured(uint32_t rc) { return (rc >> 26) & 1u; }
inline bool rc_quad_or_polyline(uint32_t rc) { return (rc >> 27) & 1u; }
inline bool rc_shaded(uint32_t rc) { return (rc >> 28) & 1u; }
inline uint32_t rc_rect_size(uint32_t rc) { return (rc >> 27) & 3u; }
inline uint32_t rc_primitive(uint32_t rc) { return (rc >> 29) & 7u; }

constexpr uint16_t DRAW_MODE_MASK = 0x1FFF;            // GPUDrawModeReg::MASK
constexpr uint16_t POLYGON_TEXPAGE_MASK = 0x09FF;      // GPUDrawModeReg::POLYGON_TEXPAGE_MASK
constexpr size_t VRAM_WORDS_MAX = (1024 * 512) / 2;    // MAX_BLIT_BUFFER_SIZE, gpu.cpp:108
constexpr size_t MAX_POLYLINE_WORDS = 1024;            // sanity bound like the reference's MAX_POLYLINE_BUFFER_SIZE
} // namespace

Gp0Frontend::Gp0Frontend()
{
  reset(false);
}

void Gp0Frontend::reset(bool pal_region)
{
  m_console_pal = pal_region;
  m_gpustat_unmodelled = 0x14802000u & GPUSTAT_UNMODELLED;
  if (m_state == State::WritingVRAM)
    finish_vram_write();
  m_fifo.clear();
  m_head = 0;
  m_state = State::Idle;
  m_blit.clear();
  m_polyline.clear();
  m_blit_remaining = 0;
  m_area = psx_drawing_area{0, 0, 0, 0};
  m_area_changed = true;
  m_offset_x = m_offset_y = 0;
  m_set_mask = m_check_mask = false;
  m_pal = pal_region;
  m_display_disable = true;
  m_display_24 = m_vertical_interlace = m_vertical_res = false;
  m_hres1 = m_hres2 = 0;
  m_display_address_start = 0;
  m_h_range = 0xC60260;
  m_v_range = 0x3FC10;
  m_texture_window_reg = 0xFFFFFFFFu;
  set_draw_mode(0);
  m_palette = 0;
  set_texture_window(0);
  invalidate_clut();
  update_crtc_display();
}

template<typename T>
T* Gp0Frontend::emit(size_t bytes, uint8_t type)
{
  const size_t sz = (bytes + 15u) & ~static_cast<size_t>(15); // VideoThreadCommand::AlignCommandSize
  const size_t at = m_out.size();
  m_out.resize(at + sz, 0);
  T* p = reinterpret_cast<T*>(m_out.data() + at);
  reinterpret_cast<psx_cmd_header*>(p)->size = static_cast<uint32_t>(sz);
  reinterpret_cast<psx_cmd_header*>(p)->type = type;
  return p;
}

// ------------------------------------------------------------------------------------------------ state
void Gp0Frontend::set_draw_mode(uint16_t value) // gpu.cpp:2202-2214
{
  uint16_t v = value & DRAW_MODE_MASK;
  if (!m_set_texture_disable_mask)
    v &= static_cast<uint16_t>(~(1u << 11));
  m_draw_mode = v;
  m_draw_to_displayed_field = (v >> 10) & 1u;
}

void Gp0Frontend::set_texture_window(uint32_t value) // gpu.cpp:2221-2240
{
  value &= 0xFFFFFu;
  if (m_texture_window_reg == value)
    return;
  const uint8_t mask_x = value & 0x1Fu, mask_y = (value >> 5) & 0x1Fu;
  const uint8_t off_x = (value >> 10) & 0x1Fu, off_y = (value >> 15) & 0x1Fu;
  m_window.and_x = static_cast<uint8_t>(~(mask_x * 8));
  m_window.and_y = static_cast<uint8_t>(~(mask_y * 8));
  m_window.or_x = static_cast<uint8_t>((off_x & mask_x) * 8u);
  m_window.or_y = static_cast<uint8_t>((off_y & mask_y) * 8u);
  m_texture_window_reg = value;
}

void Gp0Frontend::invalidate_clut() // gpu.cpp:2178-2183
{
  m_clut_reg = 0xFFFFFFFFu;
  m_clut_is_8bit = false;
}

void Gp0Frontend::update_clut_if_needed() // gpu.cpp:2157-2176
{
  const uint32_t texmode = (m_draw_mode >> 7) & 3u;
  if (texmode >= 2)
    return;
  const bool needs8 = (texmode == 1);
  if (m_palette != m_clut_reg || (needs8 && !m_clut_is_8bit))
  {
    m_clut_reg = m_palette;
    m_clut_is_8bit = needs8;
    auto* c = emit<psx_update_clut>(sizeof(psx_update_clut), PSX_CMD_UPDATE_CLUT);
    c->reg = m_palette;
    c->clut_is_8bit = needs8;
  }
}

void Gp0Frontend::prepare_for_draw() // gpu.cpp:2956-2965
{
  if (m_area_changed)
  {
    m_area_changed = false;
    auto* c = emit<psx_set_drawing_area>(sizeof(psx_set_drawing_area), PSX_CMD_SET_DRAWING_AREA);
    c->new_area = m_area;
  }
}

bool Gp0Frontend::interlaced_rendering() const // GPUSTAT.SkipDrawingToActiveField, gpu_types.h:186-191
{
  return !m_force_progressive && m_vertical_res && m_vertical_interlace && !m_draw_to_displayed_field; // gpu.cpp:1475-1478
}

void Gp0Frontend::fill_draw_header(psx_draw_header* d, int primitive, uint32_t rc) // FillDrawCommand, gpu.cpp:2968-2996
{
  const bool tex = (primitive != 2) && rc_textured(rc);
  const bool raw = tex && rc_raw(rc);
  const bool shaded = rc_shaded(rc);
  uint8_t f0 = 0;
  if (interlaced_rendering())
    f0 |= PSX_DF_INTERLACED;
  if (m_active_line_lsb)
    f0 |= PSX_DF_ACTIVE_LINE_LSB;
  if (m_set_mask)
    f0 |= PSX_DF_SET_MASK;
  if (m_check_mask)
    f0 |= PSX_DF_CHECK_MASK;
  if (tex)
    f0 |= PSX_DF_TEXTURE;
  if (raw)
    f0 |= PSX_DF_RAW_TEXTURE;
  if (rc_transparent(rc))
    f0 |= PSX_DF_TRANSPARENCY;
  if (shaded)
    f0 |= PSX_DF_SHADING;
  const bool dreg = (m_draw_mode >> 9) & 1u;
  bool dither;
  if (primitive == 2)
    dither = dreg;
  else if (primitive == 1)
    dither = false;
  else
    dither = dreg && (shaded || (tex && !raw));
  d->flags0 = f0;
  d->flags1 = static_cast<uint8_t>((rc_quad_or_polyline(rc) ? PSX_DF1_QUAD : 0) | (dither ? PSX_DF1_DITHER : 0));
  d->draw_mode = m_draw_mode;
  d->palette = m_palette;
  d->window = m_window;
}

// ------------------------------------------------------------------------------------------------ FIFO
void Gp0Frontend::write_gp0(const uint32_t* words, size_t count)
{
  m_stats.gp0_words += count;
  if (m_head > 4096 && m_head * 2 > m_fifo.size())
  {
    m_fifo.erase(m_fifo.begin(), m_fifo.begin() + static_cast<ptrdiff_t>(m_head));
    m_head = 0;
  }
  m_fifo.insert(m_fifo.end(), words, words + count);
  execute();
}

void Gp0Frontend::execute() // GPU::TryExecuteCommands without timing, gpu.cpp:2728-2812
{
  while (fifo_size() > 0)
  {
    switch (m_state)
    {
      case State::Idle:
        if (!execute_command())
          return;
        break;

      case State::WritingVRAM:
      {
        const uint32_t n = std::min(m_blit_remaining, fifo_size());
        for (uint32_t i = 0; i < n; i++)
          m_blit.push_back(pop());
        m_blit_remaining -= n;
        if (m_blit_remaining == 0)
          finish_vram_write();
      }
      break;

      case State::ReadingVRAM: return; // the port is busy until the CPU has drained GPUREAD

      case State::DrawingPolyLine:
      {
        const bool shaded = rc_shaded(m_render_command);
        const uint32_t words_per_vertex = shaded ? 2 : 1;
        uint32_t term = shaded ? ((static_cast<uint32_t>(m_polyline.size()) & 1u) ^ 1u) : 0u;
        for (; term < fifo_size(); term += words_per_vertex)
          if ((peek(term) & 0xF000F000u) == 0x50005000u)
            break;
        const bool found = term < fifo_size();
        const uint32_t n = std::min(term, fifo_size());
        for (uint32_t i = 0; i < n; i++)
        {
          const uint32_t w = pop();
          if (m_polyline.size() < MAX_POLYLINE_WORDS)
            m_polyline.push_back(w);
        }
        if (!found)
          return;
        pop(); // the terminator
        finish_polyline();
        m_polyline.clear();
        end_command();
      }
      break;
    }
  }
}

bool Gp0Frontend::execute_command()
{
  const uint32_t word = peek(0);
  const uint32_t op = word >> 24;
  auto need = [&](uint32_t n) { return fifo_size() >= n; };
  m_stats.commands++;
  if (op >= 0x20 && op <= 0x7F)
  {
    bool done;
    switch (rc_primitive(word))
    {
      case 1: done = handle_polygon(word); break;
      case 2: done = (op & 0x08) ? handle_polyline_start(word) : handle_line(word); break;
      default: done = handle_rectangle(word); break;
    }
    if (!done)
      m_stats.commands--;
    return done;
  }
  switch (op)
  {
    case 0x01: // clear cache, gpu.cpp:2842-2851
      invalidate_clut();
      emit<psx_cmd_header>(sizeof(psx_cmd_header), PSX_CMD_CLEAR_CACHE);
      pop();
      return true;

    case 0x02: // fill, gpu.cpp:3695-3726
    {
      if (!need(3))
        break;
      const uint32_t color = pop() & 0x00FFFFFFu;
      const uint32_t xy = pop(), wh = pop();
      const uint32_t x = xy & 0x3F0u, y = (xy >> 16) & 0x1FFu;
      const uint32_t w = ((wh & 0x3FFu) + 0xFu) & ~0xFu, h = (wh >> 16) & 0x1FFu;
      if (w > 0 && h > 0)
      {
        auto* c = emit<psx_fill_vram>(sizeof(psx_fill_vram), PSX_CMD_FILL_VRAM);
        c->x = static_cast<uint16_t>(x);
        c->y = static_cast<uint16_t>(y);
        c->width = static_cast<uint16_t>(w);
        c->height = static_cast<uint16_t>(h);
        c->color = color;
        c->interlaced_rendering = interlaced_rendering();
        c->active_line_lsb = m_active_line_lsb;
      }
      return true;
    }

    case 0xE1: set_draw_mode(static_cast<uint16_t>(pop() & 0x00FFFFFFu)); return true;
    case 0xE2: set_texture_window(pop() & 0x00FFFFFFu); return true;
    case 0xE3: // gpu.cpp:2885-2903
    {
      const uint32_t p = pop() & 0x00FFFFFFu;
      const uint32_t l = p & 1023u, t = (p >> 10) & 1023u;
      if (m_area.left != l || m_area.top != t)
      {
        m_area.left = l;
        m_area.top = t;
        m_area_changed = true;
      }
      return true;
    }
    case 0xE4:
    {
      const uint32_t p = pop() & 0x00FFFFFFu;
      const uint32_t r = p & 1023u, b = (p >> 10) & 1023u;
      if (m_area.right != r || m_area.bottom != b)
      {
        m_area.right = r;
        m_area.bottom = b;
        m_area_changed = true;
      }
      return true;
    }
    case 0xE5: // gpu.cpp:2924-2939
    {
      const uint32_t p = pop() & 0x00FFFFFFu;
      m_offset_x = sext11(p & 0x7FFu);
      m_offset_y = sext11((p >> 11) & 0x7FFu);
      return true;
    }
    case 0xE6: // gpu.cpp:2941-2954
    {
      const uint32_t p = pop();
      m_set_mask = p & 1u;
      m_check_mask = (p >> 1) & 1u;
      return true;
    }
    default: break;
  }
  if (op >= 0x80 && op <= 0x9F) // VRAM -> VRAM, gpu.cpp:3856-3891
  {
    if (!need(4))
    {
      m_stats.commands--;
      return false;
    }
    pop();
    const uint32_t s = pop(), d = pop(), wh = pop();
    const uint32_t sx = s & 0x3FFu, sy = (s >> 16) & 0x1FFu, dx = d & 0x3FFu, dy = (d >> 16) & 0x1FFu;
    const uint32_t w = replace_zero(wh & 0x3FFu, 0x400), h = replace_zero((wh >> 16) & 0x1FFu, 0x200);
    if (!(sx == dx && sy == dy && !m_set_mask))
    {
      auto* c = emit<psx_copy_vram>(sizeof(psx_copy_vram), PSX_CMD_COPY_VRAM);
      c->src_x = static_cast<uint16_t>(sx);
      c->src_y = static_cast<uint16_t>(sy);
      c->dst_x = static_cast<uint16_t>(dx);
      c->dst_y = static_cast<uint16_t>(dy);
      c->width = static_cast<uint16_t>(w);
      c->height = static_cast<uint16_t>(h);
      c->set_mask_while_drawing = m_set_mask;
      c->check_mask_before_draw = m_check_mask;
    }
    return true;
  }
  if (op >= 0xA0 && op <= 0xBF) // CPU -> VRAM, gpu.cpp:3728-3765
  {
    if (!need(3))
    {
      m_stats.commands--;
      return false;
    }
    pop();
    const uint32_t coords = pop(), size = pop();
    if (size == 0xFFFFFFFFu)
      return true;
    m_xfer_x = static_cast<uint16_t>(coords & 0x3FFu);
    m_xfer_y = static_cast<uint16_t>((coords >> 16) & 0x1FFu);
    m_xfer_w = static_cast<uint16_t>(replace_zero(size & 0x3FFu, 0x400));
    m_xfer_h = static_cast<uint16_t>(replace_zero((size >> 16) & 0x1FFu, 0x200));
    m_blit_remaining = (static_cast<uint32_t>(m_xfer_w) * m_xfer_h + 1u) / 2u;
    m_blit.clear();
    m_blit.reserve(m_blit_remaining);
    m_state = State::WritingVRAM;
    return true;
  }
  if (op >= 0xC0 && op <= 0xDF) // VRAM -> CPU, gpu.cpp:3819-3854
  {
    if (!need(3))
    {
      m_stats.commands--;
      return false;
    }
    pop();
    const uint32_t coords = pop(), size = pop();
    auto* c = emit<psx_read_vram>(sizeof(psx_read_vram), PSX_CMD_READ_VRAM);
    c->x = static_cast<uint16_t>(coords & 0x3FFu);
    c->y = static_cast<uint16_t>((coords >> 16) & 0x1FFu);
    c->width = static_cast<uint16_t>((((size & 0xFFFFu) - 1u) & 0x3FFu) + 1u);
    c->height = static_cast<uint16_t>(((((size >> 16) & 0xFFFFu) - 1u) & 0x1FFu) + 1u);
    m_state = State::ReadingVRAM;
    return true;
  }
  if (op == 0x02) // incomplete fill
  {
    m_stats.commands--;
    return false;
  }
  // NOPs (0x00, 0x03-0x1E, 0xE0, 0xE7-0xEF, 0xFF), IRQ request (0x1F) and unknown opcodes all consume one word
  if (!(op == 0x00 || (op >= 0x03 && op <= 0x1F) || op == 0xE0 || (op >= 0xE7 && op <= 0xEF) || op == 0xFF))
    m_stats.unknown_commands++;
  pop();
  return true;
}

// ------------------------------------------------------------------------------------------------ primitives
bool Gp0Frontend::handle_polygon(uint32_t rc) // gpu.cpp:3099-3357
{
  const bool textured = rc_textured(rc), shaded = rc_shaded(rc), quad = rc_quad_or_polyline(rc);
  const uint32_t words_per_vertex = 1u + (textured ? 1u : 0u) + (shaded ? 1u : 0u);
  const uint32_t nv = quad ? 4u : 3u;
  const uint32_t total = words_per_vertex * nv + (shaded ? 0u : 1u);
  if (fifo_size() < total)
    return false;
  if (textured)
  {
    const uint16_t texpage = static_cast<uint16_t>((shaded ? peek(5) : peek(4)) >> 16);
    set_draw_mode(static_cast<uint16_t>((texpage & POLYGON_TEXPAGE_MASK) | (m_draw_mode & ~POLYGON_TEXPAGE_MASK)));
    m_palette = static_cast<uint16_t>((peek(2) >> 16) & 0x7FFFu);
    update_clut_if_needed();
  }
  m_render_command = rc;
  pop();
  prepare_for_draw();

  psx_poly_vertex v[4];
  std::memset(v, 0, sizeof(v));
  const uint32_t first_color = rc & 0x00FFFFFFu;
  for (uint32_t i = 0; i < nv; i++)
  {
    const uint32_t color = (shaded && i > 0) ? (pop() & 0x00FFFFFFu) : first_color;
    const uint32_t pos = pop();
    v[i].x = m_offset_x + vertex_x(pos);
    v[i].y = m_offset_y + vertex_y(pos);
    v[i].r = static_cast<uint8_t>(color);
    v[i].g = static_cast<uint8_t>(color >> 8);
    v[i].b = static_cast<uint8_t>(color >> 16);
    const uint16_t tc = textured ? static_cast<uint16_t>(pop()) : 0;
    v[i].u = static_cast<uint8_t>(tc);
    v[i].v = static_cast<uint8_t>(tc >> 8);
  }
  auto too_big = [](const psx_poly_vertex& a, const psx_poly_vertex& b, const psx_poly_vertex& c) {
    const int minx = std::min({a.x, b.x, c.x}), maxx = std::max({a.x, b.x, c.x});
    const int miny = std::min({a.y, b.y, c.y}), maxy = std::max({a.y, b.y, c.y});
    return (maxx - minx + 1) > PSX_MAX_PRIM_WIDTH || (maxy - miny + 1) > PSX_MAX_PRIM_HEIGHT;
  };
  uint32_t n = nv;
  const bool first_culled = too_big(v[0], v[1], v[2]);
  if (!quad)
  {
    if (first_culled)
    {
      m_stats.culled_primitives++;
      return true;
    }
  }
  else
  {
    const bool second_culled = too_big(v[2], v[1], v[3]);
    if (first_culled && second_culled)
    {
      m_stats.culled_primitives++;
      return true;
    }
    if (second_culled)
      n = 3;
    else if (first_culled)
    {
      v[0] = v[2];
      v[2] = v[3];
      n = 3;
    }
  }
  auto* c = emit<psx_draw_polygon>(sizeof(psx_draw_header) + sizeof(psx_poly_vertex) * n, PSX_CMD_DRAW_POLYGON);
  fill_draw_header(&c->d, 0, rc);
  c->d.num_vertices = static_cast<uint16_t>(n);
  std::memcpy(c->vertices, v, sizeof(psx_poly_vertex) * n);
  return true;
}

bool Gp0Frontend::handle_rectangle(uint32_t rc) // gpu.cpp:3359-3436
{
  const bool textured = rc_textured(rc);
  const uint32_t size_mode = rc_rect_size(rc);
  const uint32_t total = 2u + (textured ? 1u : 0u) + (size_mode == 0 ? 1u : 0u);
  if (fifo_size() < total)
    return false;
  if (textured)
  {
    m_palette = static_cast<uint16_t>((peek(2) >> 16) & 0x7FFFu);
    update_clut_if_needed();
  }
  m_render_command = rc;
  pop();
  prepare_for_draw();
  auto* c = emit<psx_draw_rectangle>(sizeof(psx_draw_rectangle), PSX_CMD_DRAW_RECTANGLE);
  fill_draw_header(&c->d, 1, rc);
  c->d.num_vertices = 0; // (the reference leaves the field of a rectangle record unset: gpu.cpp:3424-3470 never writes it)
  c->color = rc & 0x00FFFFFFu;
  const uint32_t pos = pop();
  c->x = sext11(static_cast<uint32_t>(m_offset_x + vertex_x(pos)));
  c->y = sext11(static_cast<uint32_t>(m_offset_y + vertex_y(pos)));
  if (textured)
  {
    const uint32_t tp = pop();
    c->d.palette = static_cast<uint16_t>(tp >> 16);
    c->texcoord = static_cast<uint16_t>(tp);
  }
  else
  {
    c->d.palette = 0;
    c->texcoord = 0;
  }
  switch (size_mode)
  {
    case 1: c->width = c->height = 1; break;
    case 2: c->width = c->height = 8; break;
    case 3: c->width = c->height = 16; break;
    default:
    {
      const uint32_t wh = pop();
      c->width = static_cast<uint16_t>(wh & 0x3FFu);
      c->height = static_cast<uint16_t>((wh >> 16) & 0x1FFu);
    }
    break;
  }
  return true;
}

static bool line_too_big(const psx_line_vertex& a, const psx_line_vertex& b) // gpu.cpp:3527-3535
{
  const int w = std::abs(a.x - b.x) + 1, h = std::abs(a.y - b.y) + 1;
  return w > PSX_MAX_PRIM_WIDTH || h > PSX_MAX_PRIM_HEIGHT;
}

bool Gp0Frontend::handle_line(uint32_t rc) // gpu.cpp:3438-3543
{
  const bool shaded = rc_shaded(rc);
  if (fifo_size() < (shaded ? 4u : 3u))
    return false;
  m_render_command = rc;
  pop();
  prepare_for_draw();
  psx_line_vertex v[2];
  std::memset(v, 0, sizeof(v));
  const uint32_t c0 = rc & 0x00FFFFFFu;
  const uint32_t p0 = pop();
  const uint32_t c1 = shaded ? (pop() & 0x00FFFFFFu) : c0;
  const uint32_t p1 = pop();
  const uint32_t cols[2] = {c0, c1}, poss[2] = {p0, p1};
  for (int i = 0; i < 2; i++)
  {
    v[i].x = m_offset_x + vertex_x(poss[i]);
    v[i].y = m_offset_y + vertex_y(poss[i]);
    v[i].r = static_cast<uint8_t>(cols[i]);
    v[i].g = static_cast<uint8_t>(cols[i] >> 8);
    v[i].b = static_cast<uint8_t>(cols[i] >> 16);
  }
  if (line_too_big(v[0], v[1]))
  {
    m_stats.culled_primitives++;
    return true;
  }
  auto* c = emit<psx_draw_line>(sizeof(psx_draw_header) + sizeof(psx_line_vertex) * 2, PSX_CMD_DRAW_LINE);
  fill_draw_header(&c->d, 2, rc);
  c->d.palette = 0;
  c->d.num_vertices = 2;
  c->vertices[0] = v[0];
  c->vertices[1] = v[1];
  return true;
}

bool Gp0Frontend::handle_polyline_start(uint32_t rc) // gpu.cpp:3545-3575 (the word counts are the reference's)
{
  const uint32_t min_words = rc_shaded(rc) ? 3u : 4u;
  if (fifo_size() < min_words)
    return false;
  m_render_command = rc;
  pop();
  m_polyline.clear();
  for (uint32_t i = 0; i + 1 < min_words; i++)
    m_polyline.push_back(pop());
  m_state = State::DrawingPolyLine;
  return true;
}

void Gp0Frontend::finish_polyline() // gpu.cpp:3577-3691
{
  prepare_for_draw();
  const bool shaded = rc_shaded(m_render_command);
  const uint32_t nv = (static_cast<uint32_t>(m_polyline.size()) + (shaded ? 1u : 0u)) >> (shaded ? 1 : 0);
  if (nv < 2)
    return;
  std::vector<psx_line_vertex> out;
  size_t pos = 0;
  psx_line_vertex start{};
  const uint32_t first_color = m_render_command & 0x00FFFFFFu;
  {
    const uint32_t p = m_polyline[pos++];
    start.x = m_offset_x + vertex_x(p);
    start.y = m_offset_y + vertex_y(p);
    start.r = static_cast<uint8_t>(first_color);
    start.g = static_cast<uint8_t>(first_color >> 8);
    start.b = static_cast<uint8_t>(first_color >> 16);
  }
  for (uint32_t i = 1; i < nv && pos < m_polyline.size(); i++)
  {
    const uint32_t color = shaded ? (m_polyline[pos++] & 0x00FFFFFFu) : first_color;
    if (pos >= m_polyline.size())
      break;
    const uint32_t p = m_polyline[pos++];
    psx_line_vertex end{};
    end.x = m_offset_x + vertex_x(p);
    end.y = m_offset_y + vertex_y(p);
    end.r = static_cast<uint8_t>(color);
    end.g = static_cast<uint8_t>(color >> 8);
    end.b = static_cast<uint8_t>(color >> 16);
    if (line_too_big(start, end))
      m_stats.culled_primitives++;
    else
    {
      out.push_back(start);
      out.push_back(end);
    }
    start = end;
  }
  if (out.empty())
    return;
  auto* c = emit<psx_draw_line>(sizeof(psx_draw_header) + sizeof(psx_line_vertex) * out.size(), PSX_CMD_DRAW_LINE);
  fill_draw_header(&c->d, 2, m_render_command);
  c->d.palette = 0;
  c->d.num_vertices = static_cast<uint16_t>(out.size());
  std::memcpy(c->vertices, out.data(), sizeof(psx_line_vertex) * out.size());
}

void Gp0Frontend::finish_vram_write() // gpu.cpp:3767-3817
{
  auto upload = [&](uint16_t x, uint16_t y, uint16_t w, uint16_t h, const uint8_t* px) {
    const size_t payload = static_cast<size_t>(w) * h * 2;
    auto* c = emit<psx_update_vram>(PSX_UPDATE_VRAM_DATA_OFFSET + payload, PSX_CMD_UPDATE_VRAM);
    c->x = x;
    c->y = y;
    c->width = w;
    c->height = h;
    c->set_mask_while_drawing = m_set_mask;
    c->check_mask_before_draw = m_check_mask;
    std::memcpy(reinterpret_cast<uint8_t*>(c) + PSX_UPDATE_VRAM_DATA_OFFSET, px, payload);
  };
  const uint8_t* data = reinterpret_cast<const uint8_t*>(m_blit.data());
  if (m_blit_remaining == 0)
  {
    if (m_xfer_w && m_xfer_h)
      upload(m_xfer_x, m_xfer_y, m_xfer_w, m_xfer_h, data);
  }
  else if (m_xfer_w)
  {
    // transfer cut short (GP1 reset / FIFO clear): full rows, then the partial last row
    const uint32_t pixels = static_cast<uint32_t>(m_blit.size()) * 2u;
    const uint32_t rows = pixels / m_xfer_w, last = pixels % m_xfer_w;
    if (rows > 0)
      upload(m_xfer_x, m_xfer_y, m_xfer_w, static_cast<uint16_t>(rows), data);
    if (last > 0)
      upload(m_xfer_x, static_cast<uint16_t>(m_xfer_y + rows), static_cast<uint16_t>(last), 1,
             data + static_cast<size_t>(m_xfer_w) * rows * 2);
  }
  m_blit.clear();
  m_blit_remaining = 0;
  m_xfer_x = m_xfer_y = m_xfer_w = m_xfer_h = 0;
  m_state = State::Idle;
}

void Gp0Frontend::finish_vram_read()
{
  if (m_state == State::ReadingVRAM)
  {
    m_state = State::Idle;
    execute();
  }
}

// ------------------------------------------------------------------------------------------------ GP1 / display
void Gp0Frontend::write_gp1(uint32_t value) // gpu.cpp:1931-2106
{
  m_stats.gp1_words++;
  const uint32_t command = (value >> 24) & 0x3Fu, param = value & 0x00FFFFFFu;
  switch (command)
  {
    case 0x00: reset(m_console_pal); break;
    case 0x01:
      if (m_state == State::WritingVRAM)
        finish_vram_write();
      m_state = State::Idle;
      m_fifo.clear();
      m_head = 0;
      m_blit.clear();
      m_polyline.clear();
      m_blit_remaining = 0;
      break;
    case 0x03: m_display_disable = value & 1u; break;
    case 0x05:
    {
      const uint32_t nv = param & 0x7FFFEu; // CRTCState::Regs::DISPLAY_ADDRESS_START_MASK (gpu.cpp:242): bit 0 is not kept
      if (m_display_address_start != nv)
      {
        m_display_address_start = nv;
        update_crtc_display();
        emit<psx_cmd_header>(sizeof(psx_cmd_header), PSX_CMD_BUFFER_SWAPPED);
      }
    }
    break;
    case 0x06:
      m_h_range = param & 0xFFFFFFu;
      update_crtc_display();
      break;
    case 0x07:
      m_v_range = param & 0xFFFFFu;
      update_crtc_display();
      break;
    case 0x08: // GP1SetDisplayMode, gpu_types.h:137-148
      if (!m_vertical_interlace && ((param >> 5) & 1u) && !m_force_progressive)
        emit<psx_cmd_header>(sizeof(psx_cmd_header), PSX_CMD_CLEAR_DISPLAY);
      m_hres1 = param & 3u;
      m_vertical_res = (param >> 2) & 1u;
      m_pal = (param >> 3) & 1u;
      m_display_24 = (param >> 4) & 1u;
      m_vertical_interlace = (param >> 5) & 1u;
      m_hres2 = (param >> 6) & 1u;
      update_crtc_display();
      break;
    case 0x09: m_set_texture_disable_mask = param & 1u; break;
    default: break; // ack IRQ, DMA direction, GPU info: no backend effect
  }
}

void Gp0Frontend::update_crtc_display() // UpdateCRTCConfig + UpdateCRTCDisplayParameters, gpu.cpp:1165-1431
{
  static const uint16_t dividers[8] = {10, 8, 5, 4, 7, 7, 7, 7};
  const uint16_t div = dividers[m_hres1 | (m_hres2 << 2)];
  const uint16_t h_total = m_pal ? 3406 : 3413, v_total = m_pal ? 314 : 263;
  const uint16_t X = m_display_address_start & 0x3FFu, Y = (m_display_address_start >> 10) & 0x1FFu;
  const uint16_t X1 = m_h_range & 0xFFFu, X2 = (m_h_range >> 12) & 0xFFFu;
  const uint16_t Y1 = m_v_range & 0x3FFu, Y2 = (m_v_range >> 10) & 0x3FFu;
  const uint16_t hds = static_cast<uint16_t>((std::min<uint16_t>(X1, h_total) / div) * div);
  const uint16_t hde = static_cast<uint16_t>((std::min<uint16_t>(X2, h_total) / div) * div);
  const uint16_t vds = std::min<uint16_t>(Y1, v_total), vde = std::min<uint16_t>(Y2, v_total);
  // active / overscan windows (gpu.cpp:82-105)
  const uint16_t act_hs = 488, act_he = m_pal ? 3300 : 3288, act_vs = m_pal ? 20 : 16, act_ve = m_pal ? 308 : 256;
  uint16_t hvs, hve, vvs, vve;
  switch (m_crop)
  {
    case CropMode::None:
      hvs = act_hs, hve = act_he, vvs = act_vs, vve = act_ve;
      break;
    case CropMode::Overscan:
      hvs = m_pal ? 628 : 608, hve = m_pal ? 3188 : 3168, vvs = m_pal ? 30 : 24, vve = m_pal ? 298 : 248;
      break;
    default: hvs = hds, hve = hde, vvs = vds, vve = vde; break;
  }
  hvs = std::clamp<uint16_t>(hvs, act_hs, act_he);
  hve = std::clamp<uint16_t>(hve, hvs, act_he);
  vvs = std::clamp<uint16_t>(vvs, act_vs, act_ve);
  vve = std::clamp<uint16_t>(vve, vvs, act_ve);
  const uint8_t y_shift = (m_vertical_interlace && m_vertical_res) ? 1 : 0;
  const uint8_t height_shift = m_force_progressive ? y_shift : (m_vertical_interlace ? 1 : 0);
  m_crtc.display_width = static_cast<uint16_t>((hve - hvs) / div);
  m_crtc.display_height = static_cast<uint16_t>((vve - vvs) << height_shift);
  const uint16_t ticks = (hde < hds) ? 0 : static_cast<uint16_t>(hde - hds);
  const uint16_t pixels = ticks / div;
  uint16_t vram_width = (pixels == 1u) ? 4u : static_cast<uint16_t>((pixels + 2u) & ~3u);
  uint16_t skip;
  if (hds >= hvs)
  {
    m_crtc.origin_left = static_cast<uint16_t>((hds - hvs) / div);
    m_crtc.vram_left = X;
    skip = 0;
  }
  else
  {
    skip = static_cast<uint16_t>((hvs - hds) / div);
    m_crtc.origin_left = 0;
    m_crtc.vram_left = static_cast<uint16_t>((X + skip) % PSX_VRAM_WIDTH);
  }
  vram_width = static_cast<uint16_t>(vram_width - std::min(vram_width, skip));
  vram_width = std::min<uint16_t>(vram_width, static_cast<uint16_t>(m_crtc.display_width - m_crtc.origin_left));
  m_crtc.vram_width = vram_width;
  if (vds >= vvs)
  {
    m_crtc.origin_top = static_cast<uint16_t>((vds - vvs) << y_shift);
    m_crtc.vram_top = Y;
  }
  else
  {
    m_crtc.origin_top = 0;
    m_crtc.vram_top = static_cast<uint16_t>((Y + ((vvs - vds) << y_shift)) % PSX_VRAM_HEIGHT);
  }
  const uint16_t v_end = (vde <= vve) ? vde : vve;
  m_crtc.vram_height = static_cast<uint16_t>((v_end - std::min(v_end, std::max(vds, vvs))) << height_shift);
}

void Gp0Frontend::vsync() // GPU::UpdateDisplay (gpu.cpp:2451-2481) + the field flip of the CRTC tick handler (:1679-1721)
{
  auto* c = emit<psx_update_display>(sizeof(psx_update_display), PSX_CMD_UPDATE_DISPLAY);
  const bool interlaced = !m_force_progressive && m_vertical_interlace; // IsInterlacedDisplayEnabled, gpu.cpp:1465-1468
  const bool in_480i = m_vertical_res && m_vertical_interlace;
  c->display_width = m_crtc.display_width;
  c->display_height = m_crtc.display_height;
  c->display_origin_left = m_crtc.origin_left;
  c->display_origin_top = m_crtc.origin_top;
  c->display_vram_left = m_crtc.vram_left;
  c->display_vram_top = m_crtc.vram_top;
  c->display_vram_width = m_crtc.vram_width;
  c->display_vram_height = static_cast<uint16_t>(m_crtc.vram_height >> (interlaced ? 1 : 0));
  c->display_pixel_aspect_ratio = 1.0f;
  c->X = m_display_address_start & 0x3FFu;
  const bool disabled = m_display_disable || m_crtc.vram_width == 0 || m_crtc.vram_height == 0;
  c->flags = static_cast<uint8_t>((interlaced ? PSX_UD_INTERLACED_ENABLED : 0) | (m_interlaced_field ? PSX_UD_INTERLACED_FIELD : 0) |
                                  ((interlaced && m_vertical_res) ? PSX_UD_INTERLACED_INTERLEAVED : 0) |
                                  (in_480i ? PSX_UD_480I_MODE : 0) | (m_display_24 ? PSX_UD_DISPLAY_24BIT : 0) |
                                  (disabled ? PSX_UD_DISPLAY_DISABLED : 0) | PSX_UD_SUBMIT_FRAME);
  m_stats.frames++;
  if (!m_field_per_vsync)
    return; // (the reference's trace replay: the beam does not run, the field bits stay)
  // next frame's field / active line
  const uint8_t display_field = in_480i ? (m_interlaced_field ^ 1u) : 0u;
  m_interlaced_field = m_vertical_interlace ? (m_interlaced_field ^ 1u) : 0u;
  m_active_line_lsb = in_480i ? static_cast<uint8_t>((((m_display_address_start >> 10) & 0x1FFu) + display_field) & 1u) : 0u;
}

// ------------------------------------------------------------------------------------------------ .psxgpu traces
long replay_psxgpu_dump(const void* data, size_t bytes, Gp0Frontend& fe, void (*on_frame)(void*, Gp0Frontend&), void* user)
{
  static const char magic[14] = {'P', 'S', 'X', 'G', 'P', 'U', 'D', 'U', 'M', 'P', 'v', '1', 0, 0}; // gpu_dump.cpp:33
  const uint8_t* p = static_cast<const uint8_t*>(data);
  if (bytes < sizeof(magic) || std::memcmp(p, magic, sizeof(magic)) != 0)
    return -1;
  size_t off = sizeof(magic);
  long frames = 0;
  std::vector<uint32_t> words;
  while (off + 4 <= bytes)
  {
    uint32_t header;
    std::memcpy(&header, p + off, 4);
    off += 4;
    const uint32_t length = header & 0xFFFFFFu, type = header >> 24; // gpu_dump.h:43-49
    if (static_cast<size_t>(length) * 4 > bytes - off)
      return -1;
    words.resize(length);
    if (length)
      std::memcpy(words.data(), p + off, static_cast<size_t>(length) * 4);
    off += static_cast<size_t>(length) * 4;
    switch (type)
    {
      case 0x00: fe.write_gp0(words.data(), words.size()); break; // GPUPort0Data
      case 0x01:                                                  // GPUPort1Data
        if (length == 1)
          fe.write_gp1(words[0]);
        break;
      case 0x02: // VSyncEvent
        fe.vsync();
        frames++;
        if (on_frame)
          on_frame(user, fe);
        break;
      case 0x03: fe.finish_vram_read(); break; // DiscardPort0Data: the CPU drained GPUREAD
      default: break;                          // read-back data, trace begin, version, game id, video format, comment
    }
  }
  return off == bytes ? frames : -1;
}

} // namespace psx

// ------------------------------------------------------------------------------------------------ C ABI (include/psxgpu.h)
#include "../../include/psxgpu.h"

// ------------------------------------------------------------------------------------------------ save states
namespace psx {
namespace {

// StateWrapper in read mode (src/util/state_wrapper.h:42-228): raw little-endian values, bool = one byte,
// strings / vectors / FIFOs = u32 count + elements; any short read poisons the rest.
struct StateReader
{
  const uint8_t* p;
  size_t size, pos = 0;
  bool error = false;
  template<typename T>
  T get()
  {
    T v{};
    if (error || size - pos < sizeof(T))
    {
      error = true;
      return v;
    }
    std::memcpy(&v, p + pos, sizeof(T));
    pos += sizeof(T);
    return v;
  }
  bool get_bool() { return get<uint8_t>() != 0; }
  void skip(size_t n)
  {
    if (error || size - pos < n)
      error = true;
    else
      pos += n;
  }
  bool marker(const char* text) // DoMarker, state_wrapper.cpp:104-116
  {
    const uint32_t len = get<uint32_t>();
    const size_t want = std::strlen(text);
    if (error || len != want || size - pos < len || std::memcmp(p + pos, text, len) != 0)
    {
      error = true;
      return false;
    }
    pos += len;
    return true;
  }
};

struct StateWriter
{
  std::vector<uint8_t>& out;
  template<typename T>
  void put(T v)
  {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&v);
    out.insert(out.end(), b, b + sizeof(T));
  }
  void put_bool(bool v) { out.push_back(v ? 1 : 0); }
  void marker(const char* text)
  {
    put<uint32_t>(static_cast<uint32_t>(std::strlen(text)));
    out.insert(out.end(), text, text + std::strlen(text));
  }
};

constexpr uint32_t TC_VRAM_WRITE_BYTES = 16 * 2 + 8 + 4;            // gpu_hw_texture_cache.cpp:62
constexpr uint32_t TC_PALETTE_RECORD_BYTES = 16 + 4 + 4 + 8 + 2 * 256; // gpu_hw_texture_cache.cpp:63-64

} // namespace

bool Gp0Frontend::load_state(const uint8_t* data, size_t bytes, uint32_t version, LoadedState* out)
{
  if (!data || !out || version < 42 || version > SAVE_STATE_VERSION)
    return false;
  StateReader r{data, bytes};
  const uint32_t gpustat = r.get<uint32_t>();
  const uint16_t draw_mode = r.get<uint16_t>();
  const uint16_t palette = r.get<uint16_t>();
  const uint32_t window_reg = r.get<uint32_t>();
  if (version < 62)
    r.skip(16); // texture_page_x/y, texture_palette_x/y
  psx_texture_window window;
  window.and_x = r.get<uint8_t>();
  window.and_y = r.get<uint8_t>();
  window.or_x = r.get<uint8_t>();
  window.or_y = r.get<uint8_t>();
  r.get_bool(); // texture_x_flip / texture_y_flip: bits 12 / 13 of the draw mode again
  r.get_bool();
  psx_drawing_area area;
  area.left = r.get<uint32_t>();
  area.top = r.get<uint32_t>();
  area.right = r.get<uint32_t>();
  area.bottom = r.get<uint32_t>();
  int32_t off_x = r.get<int32_t>();
  const int32_t off_y = r.get<int32_t>();
  off_x = r.get<int32_t>(); // the reference serialises drawing_offset.x a second time (gpu.cpp:609)
  const bool console_pal = r.get_bool();
  const bool set_texture_disable_mask = r.get_bool();
  const uint32_t display_address_start = r.get<uint32_t>();
  const uint32_t h_range = r.get<uint32_t>();
  const uint32_t v_range = r.get<uint32_t>();
  r.skip(2);      // dot_clock_divider
  r.skip(2 * 8);  // display_width .. display_vram_height: recomputed from the registers
  r.skip(2 * 10); // horizontal_total .. vertical_display_end
  r.skip(4 + 4 + 4); // fractional_ticks, current_tick_in_scanline, current_scanline (as u32)
  if (version >= 46)
    r.skip(4); // fractional_dot_ticks
  r.skip(2);   // in_hblank, in_vblank
  const uint8_t interlaced_field = r.get<uint8_t>();
  r.skip(1); // interlaced_display_field: follows interlaced_field at the next UpdateDisplay
  const uint8_t active_line_lsb = r.get<uint8_t>();
  uint8_t blitter = r.get<uint8_t>();
  if (blitter > 3)
    blitter = 3; // Do(T*, max_value), state_wrapper.h:78-96
  r.skip(4 + 4 + 4); // pending_command_ticks, command_total_words, GPUREAD_latch
  uint32_t clut_reg = 0xFFFFFFFFu;
  bool clut_is_8bit = false;
  out->clut_valid = false;
  std::memset(out->clut, 0, sizeof(out->clut));
  if (version >= 64)
  {
    clut_reg = r.get<uint32_t>();
    clut_is_8bit = r.get_bool();
    for (uint16_t& c : out->clut)
      c = r.get<uint16_t>();
    out->clut_valid = true;
  }
  const uint16_t xfer_x = r.get<uint16_t>(), xfer_y = r.get<uint16_t>(), xfer_w = r.get<uint16_t>(), xfer_h = r.get<uint16_t>();
  r.skip(4); // col, row of a VRAM -> CPU read in flight
  std::vector<uint32_t> fifo, blit, polyline;
  {
    const uint32_t n = r.get<uint32_t>();
    if (n > 256 || (bytes - r.pos) / 8 < n)
      r.error = true;
    for (uint32_t i = 0; i < n && !r.error; i++)
      fifo.push_back(static_cast<uint32_t>(r.get<uint64_t>())); // low half = the word, high half = its DMA address
  }
  {
    const uint32_t n = r.get<uint32_t>();
    if (n > VRAM_WORDS_MAX || (bytes - r.pos) / 4 < n)
      r.error = true;
    else
    {
      blit.resize(n);
      if (n && !r.error)
      {
        std::memcpy(blit.data(), data + r.pos, static_cast<size_t>(n) * 4);
        r.pos += static_cast<size_t>(n) * 4;
      }
    }
  }
  const uint32_t blit_remaining = r.get<uint32_t>();
  const uint32_t render_command = r.get<uint32_t>();
  if (version >= 85)
  {
    const uint32_t n = r.get<uint32_t>();
    if (n > 0x200000u || (bytes - r.pos) / 8 < n)
      r.error = true;
    for (uint32_t i = 0; i < n && !r.error; i++)
      polyline.push_back(static_cast<uint32_t>(r.get<uint64_t>()));
  }
  if (version < 83)
    r.skip(8);
  if (!r.marker("GPU-VRAM"))
    return false;
  out->vram = data + r.pos;
  r.skip(PSX_VRAM_PIXELS * 2);
  if (version >= 73) // GPUTextureCache::GetStateSize, gpu_hw_texture_cache.cpp:661-691
  {
    if (!r.marker("GPUTextureCache"))
      return false;
    const uint32_t writes = r.get<uint32_t>();
    if (version >= 86)
      r.skip(4);
    for (uint32_t i = 0; i < writes && !r.error; i++)
    {
      r.skip(version < 86 ? TC_VRAM_WRITE_BYTES - 4 : TC_VRAM_WRITE_BYTES);
      const uint32_t records = r.get<uint32_t>();
      if (records > (bytes - r.pos) / TC_PALETTE_RECORD_BYTES)
        r.error = true;
      else
        r.skip(static_cast<size_t>(records) * TC_PALETTE_RECORD_BYTES);
    }
  }
  if (r.error)
    return false;
  out->consumed = r.pos;

  // nothing failed: take the state over
  m_out.clear();
  m_fifo = std::move(fifo);
  m_head = 0;
  m_blit = std::move(blit);
  m_blit_remaining = blit_remaining;
  m_polyline = std::move(polyline);
  static const State states[4] = {State::Idle, State::ReadingVRAM, State::WritingVRAM, State::DrawingPolyLine};
  m_state = states[blitter];
  m_render_command = render_command;
  m_console_pal = console_pal;
  m_set_texture_disable_mask = set_texture_disable_mask;
  m_draw_mode = draw_mode;
  m_palette = palette;
  m_texture_window_reg = window_reg;
  m_window = window;
  m_area = area;
  m_area_changed = true; // gpu.cpp:704
  m_offset_x = off_x;
  m_offset_y = off_y;
  m_gpustat_unmodelled = gpustat & GPUSTAT_UNMODELLED;
  m_draw_to_displayed_field = (gpustat >> 10) & 1u;
  m_set_mask = (gpustat >> 11) & 1u;
  m_check_mask = (gpustat >> 12) & 1u;
  m_hres2 = (gpustat >> 16) & 1u;
  m_hres1 = (gpustat >> 17) & 3u;
  m_vertical_res = (gpustat >> 19) & 1u;
  m_pal = (gpustat >> 20) & 1u;
  m_display_24 = (gpustat >> 21) & 1u;
  m_vertical_interlace = (gpustat >> 22) & 1u;
  m_display_disable = (gpustat >> 23) & 1u;
  m_display_address_start = display_address_start;
  m_h_range = h_range;
  m_v_range = v_range;
  m_interlaced_field = interlaced_field;
  m_active_line_lsb = active_line_lsb;
  m_clut_reg = clut_reg;
  m_clut_is_8bit = clut_is_8bit;
  m_xfer_x = xfer_x;
  m_xfer_y = xfer_y;
  m_xfer_w = xfer_w;
  m_xfer_h = xfer_h;
  update_crtc_display();
  return true;
}

void Gp0Frontend::save_state(std::vector<uint8_t>& out, const uint16_t* vram, const uint16_t* clut) const
{
  StateWriter w{out};
  // GPUSTAT: draw-mode bits 0-10 and 15 mirror the register (gpu.cpp:2210-2213); the rest from the display state
  uint32_t gpustat = (m_draw_mode & 0x7FFu) | (((m_draw_mode >> 11) & 1u) << 15);
  gpustat |= (m_set_mask ? 1u : 0u) << 11 | (m_check_mask ? 1u : 0u) << 12;
  gpustat |= static_cast<uint32_t>(m_hres2) << 16 | static_cast<uint32_t>(m_hres1) << 17 | (m_vertical_res ? 1u : 0u) << 19 |
             (m_pal ? 1u : 0u) << 20 | (m_display_24 ? 1u : 0u) << 21 | (m_vertical_interlace ? 1u : 0u) << 22 |
             (m_display_disable ? 1u : 0u) << 23;
  // bits this front-end does not model (13 field status, 14 reverse flag, 24-31 the dynamic ready / DMA bits): as loaded,
  // or their reset values (gpu.cpp:499: 0x14802000)
  gpustat |= m_gpustat_unmodelled & GPUSTAT_UNMODELLED;
  w.put<uint32_t>(gpustat);
  w.put<uint16_t>(m_draw_mode);
  w.put<uint16_t>(m_palette);
  w.put<uint32_t>(m_texture_window_reg);
  w.put<uint8_t>(m_window.and_x);
  w.put<uint8_t>(m_window.and_y);
  w.put<uint8_t>(m_window.or_x);
  w.put<uint8_t>(m_window.or_y);
  w.put_bool((m_draw_mode >> 12) & 1u);
  w.put_bool((m_draw_mode >> 13) & 1u);
  w.put<uint32_t>(m_area.left);
  w.put<uint32_t>(m_area.top);
  w.put<uint32_t>(m_area.right);
  w.put<uint32_t>(m_area.bottom);
  w.put<int32_t>(m_offset_x);
  w.put<int32_t>(m_offset_y);
  w.put<int32_t>(m_offset_x);
  w.put_bool(m_console_pal);
  w.put_bool(m_set_texture_disable_mask);
  w.put<uint32_t>(m_display_address_start);
  w.put<uint32_t>(m_h_range);
  w.put<uint32_t>(m_v_range);
  static const uint16_t dividers[8] = {10, 8, 5, 4, 7, 7, 7, 7};
  w.put<uint16_t>(dividers[m_hres1 | (m_hres2 << 2)]);
  w.put<uint16_t>(m_crtc.display_width);
  w.put<uint16_t>(m_crtc.display_height);
  w.put<uint16_t>(m_crtc.origin_left);
  w.put<uint16_t>(m_crtc.origin_top);
  w.put<uint16_t>(m_crtc.vram_left);
  w.put<uint16_t>(m_crtc.vram_top);
  w.put<uint16_t>(m_crtc.vram_width);
  w.put<uint16_t>(m_crtc.vram_height);
  w.put<uint16_t>(m_pal ? 3406 : 3413); // horizontal_total
  for (int i = 0; i < 4; i++)
    w.put<uint16_t>(0); // horizontal_visible_start/end, horizontal_display_start/end (beam timing: not modelled)
  w.put<uint16_t>(m_pal ? 314 : 263); // vertical_total
  for (int i = 0; i < 4; i++)
    w.put<uint16_t>(0);
  w.put<int32_t>(0);  // fractional_ticks
  w.put<int32_t>(0);  // current_tick_in_scanline
  w.put<uint32_t>(0); // current_scanline
  w.put<int32_t>(0);  // fractional_dot_ticks
  w.put_bool(false);  // in_hblank
  w.put_bool(false);  // in_vblank
  w.put<uint8_t>(m_interlaced_field);
  w.put<uint8_t>(m_interlaced_field);
  w.put<uint8_t>(m_active_line_lsb);
  static const uint8_t blitter_of[4] = {0, 2, 1, 3}; // State -> BlitterState (gpu.cpp:63-69)
  w.put<uint8_t>(blitter_of[static_cast<uint8_t>(m_state)]);
  w.put<int32_t>(0);  // pending_command_ticks
  w.put<uint32_t>(0); // command_total_words
  w.put<uint32_t>(0); // GPUREAD_latch
  w.put<uint32_t>(m_clut_reg);
  w.put_bool(m_clut_is_8bit);
  for (int i = 0; i < 256; i++)
    w.put<uint16_t>(clut ? clut[i] : 0);
  w.put<uint16_t>(m_xfer_x);
  w.put<uint16_t>(m_xfer_y);
  w.put<uint16_t>(m_xfer_w);
  w.put<uint16_t>(m_xfer_h);
  w.put<uint16_t>(0);
  w.put<uint16_t>(0);
  w.put<uint32_t>(fifo_size());
  for (size_t i = m_head; i < m_fifo.size(); i++)
    w.put<uint64_t>(m_fifo[i]);
  w.put<uint32_t>(static_cast<uint32_t>(m_blit.size()));
  for (uint32_t v : m_blit)
    w.put<uint32_t>(v);
  w.put<uint32_t>(m_blit_remaining);
  w.put<uint32_t>(m_render_command);
  w.put<uint32_t>(static_cast<uint32_t>(m_polyline.size()));
  for (uint32_t v : m_polyline)
    w.put<uint64_t>(v);
  w.marker("GPU-VRAM");
  const uint8_t* vb = reinterpret_cast<const uint8_t*>(vram);
  out.insert(out.end(), vb, vb + PSX_VRAM_PIXELS * 2);
  w.marker("GPUTextureCache");
  w.put<uint32_t>(0);           // num_vram_writes
  w.put<uint32_t>(0xFFFFFFFFu); // last_vram_write_index = INVALID_VRAM_WRITE_INDEX
}

} // namespace psx

struct psxgpu_frontend
{
  psx::Gp0Frontend fe;
};

extern "C" {

int psxgpu_frontend_create(uint32_t flags, uint32_t crop_mode, psxgpu_frontend** out)
{
  if (!out || crop_mode > 2)
    return PSXGPU_E_INVALID;
  psxgpu_frontend* f = new (std::nothrow) psxgpu_frontend();
  if (!f)
    return PSXGPU_E_NOMEM;
  f->fe.set_crop_mode(static_cast<psx::CropMode>(crop_mode));
  f->fe.set_force_progressive(!(flags & PSXGPU_FE_INTERLACED));
  f->fe.set_field_per_vsync((flags & PSXGPU_FE_FIELD_PER_VSYNC) != 0);
  f->fe.reset((flags & PSXGPU_FE_PAL) != 0);
  f->fe.clear_output();
  *out = f;
  return PSXGPU_OK;
}

void psxgpu_frontend_destroy(psxgpu_frontend* f)
{
  delete f;
}

int psxgpu_frontend_gp0(psxgpu_frontend* f, const uint32_t* words, size_t count)
{
  if (!f || (!words && count))
    return PSXGPU_E_INVALID;
  f->fe.write_gp0(words, count);
  return PSXGPU_OK;
}

int psxgpu_frontend_gp1(psxgpu_frontend* f, uint32_t value)
{
  if (!f)
    return PSXGPU_E_INVALID;
  f->fe.write_gp1(value);
  return PSXGPU_OK;
}

int psxgpu_frontend_vsync(psxgpu_frontend* f)
{
  if (!f)
    return PSXGPU_E_INVALID;
  f->fe.vsync();
  return PSXGPU_OK;
}

int psxgpu_frontend_finish_read(psxgpu_frontend* f)
{
  if (!f)
    return PSXGPU_E_INVALID;
  f->fe.finish_vram_read();
  return PSXGPU_OK;
}

const void* psxgpu_frontend_records(psxgpu_frontend* f, size_t* bytes)
{
  if (!f)
    return nullptr;
  if (bytes)
    *bytes = f->fe.output().size();
  return f->fe.output().data();
}

int psxgpu_frontend_clear(psxgpu_frontend* f)
{
  if (!f)
    return PSXGPU_E_INVALID;
  f->fe.clear_output();
  return PSXGPU_OK;
}

int psxgpu_frontend_get_stats(psxgpu_frontend* f, psxgpu_frontend_stats* out)
{
  if (!f || !out)
    return PSXGPU_E_INVALID;
  const auto& s = f->fe.stats();
  std::memset(out, 0, sizeof(*out));
  out->gp0_words = s.gp0_words;
  out->gp1_words = s.gp1_words;
  out->commands = s.commands;
  out->unknown_commands = s.unknown_commands;
  out->culled_primitives = s.culled_primitives;
  out->frames = s.frames;
  out->idle = f->fe.idle();
  return PSXGPU_OK;
}

long psxgpu_frontend_feed_dump(psxgpu_frontend* f, const void* data, size_t bytes)
{
  if (!f || !data)
    return PSXGPU_E_INVALID;
  const long frames = psx::replay_psxgpu_dump(data, bytes, f->fe, nullptr, nullptr);
  return frames < 0 ? static_cast<long>(PSXGPU_E_MALFORMED) : frames;
}

namespace {
struct ReplayCtx
{
  psxgpu_ctx* ctx;
  uint32_t instance;
  psxgpu_frame_callback cb;
  void* user;
  long frame;
  int err;
};
void replay_frame(void* user, psx::Gp0Frontend& fe)
{
  ReplayCtx* r = static_cast<ReplayCtx*>(user);
  auto& out = fe.output();
  if (r->err == PSXGPU_OK && !out.empty())
    r->err = psxgpu_submit_wire(r->ctx, r->instance, out.data(), out.size());
  fe.clear_output();
  if (r->err == PSXGPU_OK && r->cb)
    r->cb(r->user, r->frame);
  r->frame++;
}
} // namespace

long psxgpu_replay_dump(psxgpu_ctx* ctx, uint32_t instance, const void* data, size_t bytes, uint32_t flags,
                        uint32_t crop_mode, psxgpu_frame_callback on_frame, void* user)
{
  if (!ctx || !data || crop_mode > 2)
    return PSXGPU_E_INVALID;
  psx::Gp0Frontend fe;
  fe.set_crop_mode(static_cast<psx::CropMode>(crop_mode));
  fe.set_force_progressive(!(flags & PSXGPU_FE_INTERLACED));
  fe.set_field_per_vsync((flags & PSXGPU_FE_FIELD_PER_VSYNC) != 0);
  fe.reset((flags & PSXGPU_FE_PAL) != 0);
  fe.clear_output();
  ReplayCtx r{ctx, instance, on_frame, user, 0, PSXGPU_OK};
  const long frames = psx::replay_psxgpu_dump(data, bytes, fe, &replay_frame, &r);
  if (frames < 0)
    return PSXGPU_E_MALFORMED;
  if (r.err == PSXGPU_OK && !fe.output().empty()) // records after the last VSync
    r.err = psxgpu_submit_wire(ctx, instance, fe.output().data(), fe.output().size());
  if (r.err != PSXGPU_OK)
    return r.err;
  return frames;
}


// ---- save states ----
int psxgpu_state_file_info(const void* file, size_t bytes, psxgpu_state_file* out)
{
  // SAVE_STATE_HEADER, src/core/save_state_version.h:14-56 (pack 4): magic, version, title[128], serial[32], then
  // twelve u32: media path (length, offset, subimage), screenshot (compression, width, height, size, offset),
  // data (compression, compressed size, uncompressed size, offset) — 216 bytes
  constexpr size_t HEADER_BYTES = 8 + 128 + 32 + 12 * 4;
  if (!file || !out || bytes < HEADER_BYTES)
    return PSXGPU_E_INVALID;
  const uint8_t* b = static_cast<const uint8_t*>(file);
  uint32_t head[2], f[12];
  std::memcpy(head, b, 8);
  std::memcpy(f, b + 8 + 128 + 32, sizeof(f));
  if (head[0] != 0x43435544u) // SAVE_STATE_MAGIC
    return PSXGPU_E_MALFORMED;
  std::memset(out, 0, sizeof(*out));
  out->version = head[1];
  std::memcpy(out->title, b + 8, 128);
  out->title[127] = 0;
  std::memcpy(out->serial, b + 8 + 128, 32);
  out->serial[31] = 0;
  out->data_compression = f[8];
  out->data_compressed_size = f[9];
  out->data_uncompressed_size = f[10];
  out->offset_to_data = f[11];
  if (static_cast<uint64_t>(out->offset_to_data) + out->data_compressed_size > bytes)
    return PSXGPU_E_MALFORMED;
  return PSXGPU_OK;
}

long psxgpu_state_find_gpu_section(const void* state, size_t bytes, uint32_t version)
{
  if (!state)
    return PSXGPU_E_INVALID;
  static const uint8_t gpu_marker[7] = {3, 0, 0, 0, 'G', 'P', 'U'};       // sw.DoMarker("GPU"), system.cpp:2497
  static const uint8_t next_marker[9] = {5, 0, 0, 0, 'C', 'D', 'R', 'O', 'M'}; // system.cpp:2500
  const uint8_t* b = static_cast<const uint8_t*>(state);
  for (size_t i = 0; i + sizeof(gpu_marker) <= bytes; i++)
  {
    if (std::memcmp(b + i, gpu_marker, sizeof(gpu_marker)) != 0)
      continue;
    const size_t at = i + sizeof(gpu_marker);
    psx::Gp0Frontend probe;
    psx::Gp0Frontend::LoadedState ls;
    if (!probe.load_state(b + at, bytes - at, version, &ls))
      continue;
    const size_t end = at + ls.consumed;
    if (end == bytes || (end + sizeof(next_marker) <= bytes && std::memcmp(b + end, next_marker, sizeof(next_marker)) == 0))
      return static_cast<long>(at);
  }
  return PSXGPU_E_MALFORMED;
}

int psxgpu_frontend_load_state(psxgpu_frontend* f, psxgpu_ctx* ctx, uint32_t instance, const void* section, size_t bytes,
                               uint32_t version, size_t* consumed)
{
  if (!f || !section)
    return PSXGPU_E_INVALID;
  psx::Gp0Frontend::LoadedState ls;
  if (!f->fe.load_state(static_cast<const uint8_t*>(section), bytes, version, &ls))
    return PSXGPU_E_MALFORMED;
  if (consumed)
    *consumed = ls.consumed;
  if (!ctx)
    return PSXGPU_OK;
  // the GPUBackendLoadStateCommand GPU::DoState pushes (gpu.cpp:692-701); `vram` may sit at any alignment in the file
  std::vector<uint16_t> vram(PSX_VRAM_PIXELS);
  std::memcpy(vram.data(), ls.vram, PSX_VRAM_PIXELS * 2);
  return psxgpu_load_state(ctx, instance, vram.data(), ls.clut);
}

int psxgpu_frontend_save_state(psxgpu_frontend* f, psxgpu_ctx* ctx, uint32_t instance, void* dst, size_t capacity, size_t* written)
{
  if (!f || !written)
    return PSXGPU_E_INVALID;
  std::vector<uint16_t> vram(PSX_VRAM_PIXELS), clut(PSX_CLUT_ENTRIES);
  if (ctx) // (without a context: registers only, VRAM and CLUT cache zero)
  {
    const int rc = psxgpu_save_state(ctx, instance, vram.data(), clut.data()); // GPU::DoState starts with a full ReadVRAM
    if (rc != PSXGPU_OK)
      return rc;
  }
  std::vector<uint8_t> out;
  out.reserve(PSX_VRAM_PIXELS * 2 + 4096);
  f->fe.save_state(out, vram.data(), clut.data());
  *written = out.size();
  if (!dst || capacity < out.size())
    return dst ? PSXGPU_E_NOMEM : PSXGPU_OK; // dst == nullptr: size query
  std::memcpy(dst, out.data(), out.size());
  return PSXGPU_OK;
}

} // extern "C"
