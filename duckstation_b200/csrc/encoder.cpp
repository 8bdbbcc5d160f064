// encoder.cpp — see encoder.h.
#include "encoder.h"

#include "../../include/psxgpu_wire.h"

#include <algorithm>
#include <cstring>

namespace psx {

// ---------------------------------------------------------------- HazardGrid
void HazardGrid::clear()
{
  if (m_any)
  {
    std::memset(m_rows, 0, sizeof(m_rows));
    m_any = false;
  }
}

static inline uint32_t col_bits(uint32_t x0, uint32_t x1)
{
  const uint32_t c0 = x0 >> 5, c1 = x1 >> 5;
  const uint32_t hi = (c1 == 31) ? 0xFFFFFFFFu : ((1u << (c1 + 1)) - 1u);
  return hi & ~((1u << c0) - 1u);
}

void HazardGrid::mark(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1)
{
  const uint32_t bits = col_bits(x0, x1);
  for (uint32_t r = y0 >> 3; r <= (y1 >> 3); r++)
    m_rows[r] |= bits;
  m_any = true;
}

bool HazardGrid::test(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) const
{
  if (!m_any)
    return false;
  const uint32_t bits = col_bits(x0, x1);
  for (uint32_t r = y0 >> 3; r <= (y1 >> 3); r++)
    if (m_rows[r] & bits)
      return true;
  return false;
}

// ---------------------------------------------------------------- RectSet
void RectSet::add_wrapped(uint32_t x, uint32_t y, uint32_t w, uint32_t h)
{
  if (w == 0 || h == 0)
    return;
  x &= VRAM_W - 1;
  y &= VRAM_H - 1;
  w = std::min(w, VRAM_W);
  h = std::min(h, VRAM_H);
  const uint32_t xs[2][2] = {{x, std::min(x + w, VRAM_W) - 1}, {0, (x + w > VRAM_W) ? (x + w - VRAM_W - 1) : 0}};
  const uint32_t ys[2][2] = {{y, std::min(y + h, VRAM_H) - 1}, {0, (y + h > VRAM_H) ? (y + h - VRAM_H - 1) : 0}};
  const int nx = (x + w > VRAM_W) ? 2 : 1, ny = (y + h > VRAM_H) ? 2 : 1;
  for (int j = 0; j < ny; j++)
    for (int i = 0; i < nx; i++)
      if (n < 4)
        r[n++] = R{static_cast<uint16_t>(xs[i][0]), static_cast<uint16_t>(ys[j][0]), static_cast<uint16_t>(xs[i][1]),
                   static_cast<uint16_t>(ys[j][1])};
}

bool RectSet::intersects(const RectSet& o) const
{
  for (int i = 0; i < n; i++)
    for (int j = 0; j < o.n; j++)
      if (r[i].x0 <= o.r[j].x1 && o.r[j].x0 <= r[i].x1 && r[i].y0 <= o.r[j].y1 && o.r[j].y0 <= r[i].y1)
        return true;
  return false;
}

static bool grid_hits(const HazardGrid& g, const RectSet& s)
{
  for (int i = 0; i < s.n; i++)
    if (g.test(s.r[i].x0, s.r[i].y0, s.r[i].x1, s.r[i].y1))
      return true;
  return false;
}
// ---------------------------------------------------------------- HazardSet
void HazardSet::clear()
{
  if (m_any)
  {
    m_grid.clear();
    m_pending.clear();
    m_any = false;
  }
}

void HazardSet::add(const RectSet& s)
{
  for (int i = 0; i < s.n; i++)
  {
    const RectSet::R& r = s.r[i];
    if (!m_any)
    {
      m_x0 = r.x0;
      m_y0 = r.y0;
      m_x1 = r.x1;
      m_y1 = r.y1;
      m_any = true;
    }
    else
    {
      m_x0 = std::min(m_x0, r.x0);
      m_y0 = std::min(m_y0, r.y0);
      m_x1 = std::max(m_x1, r.x1);
      m_y1 = std::max(m_y1, r.y1);
    }
    m_pending.push_back(r);
  }
}

bool HazardSet::hits(const RectSet& s)
{
  if (!m_any)
    return false;
  bool maybe = false;
  for (int i = 0; i < s.n && !maybe; i++)
    maybe = s.r[i].x0 <= m_x1 && m_x0 <= s.r[i].x1 && s.r[i].y0 <= m_y1 && m_y0 <= s.r[i].y1;
  if (!maybe)
    return false;
  for (const RectSet::R& r : m_pending)
    m_grid.mark(r.x0, r.y0, r.x1, r.y1);
  m_pending.clear();
  return grid_hits(m_grid, s);
}

// ---------------------------------------------------------------- Encoder
Encoder::Encoder(uint32_t clut_slots, bool modulation_crop)
  : m_clut_slots(std::min<uint32_t>(std::max<uint32_t>(clut_slots, 4), 256)), m_modulation_crop(modulation_crop)
{
  m_slots.resize(m_clut_slots);
  m_chain.assign(m_clut_slots, 0);
}

void Encoder::chain_insert(uint32_t slot)
{
  const uint32_t b = bucket_of(m_slots[slot].x, m_slots[slot].y);
  m_chain[slot] = m_bucket[b];
  m_bucket[b] = static_cast<uint8_t>(slot);
}

void Encoder::chain_remove(uint32_t slot)
{
  const uint32_t b = bucket_of(m_slots[slot].x, m_slots[slot].y);
  uint8_t* link = &m_bucket[b];
  while (*link != 0 && *link != slot)
    link = &m_chain[*link];
  if (*link == slot)
    *link = m_chain[slot];
  m_chain[slot] = 0;
}

void Encoder::begin_flush()
{
  cmds.clear();
  upload_cmds.clear();
  payload.clear();
  m_out_cmds = nullptr;
  m_out_payload = nullptr;
  m_out_cmd_cap = 0;
  m_out_pay_cap = 0;
  m_n_cmds = 0;
  m_n_payload = 0;
  m_overflow = false;
  epochs.clear();
  clut_ops.clear();
  m_epoch_cmd_begin = 0;
  m_epoch_clut_begin = 0;
  m_writes.clear();
  m_reads.clear();
}

void Encoder::finish_flush()
{
  if (m_n_cmds > m_epoch_cmd_begin || clut_ops.size() > m_epoch_clut_begin)
    close_epoch(EPOCH_TILED);
}

void Encoder::reset_clut_state()
{
  for (ClutSlot& s : m_slots)
    s = ClutSlot();
  std::memset(m_bucket, 0, sizeof(m_bucket));
  std::fill(m_chain.begin(), m_chain.end(), uint8_t(0));
  m_clut_sources.clear();
  m_cur_lo = m_cur_hi = 0;
  m_alloc_hand = 1;
}

void Encoder::export_state(EncoderStateBlob& out) const
{
  std::memset(&out, 0, sizeof(out));
  for (int i = 0; i < 4; i++)
    out.area[i] = m_area[i];
  out.cur_lo = static_cast<uint8_t>(m_cur_lo);
  out.cur_hi = static_cast<uint8_t>(m_cur_hi);
  out.alloc_hand = static_cast<uint16_t>(m_alloc_hand);
  for (uint32_t i = 1; i < m_clut_slots; i++)
  {
    const ClutSlot& s = m_slots[i];
    if (s.valid)
      out.slot_key[i] = s.x | (static_cast<uint32_t>(s.y) << 10) | (s.is8 ? (1u << 19) : 0u) | 0x80000000u;
  }
}

void Encoder::import_state(const EncoderStateBlob& in)
{
  reset_clut_state();
  for (int i = 0; i < 4; i++)
    m_area[i] = in.area[i];
  m_cur_lo = std::min<uint32_t>(in.cur_lo, m_clut_slots - 1);
  m_cur_hi = std::min<uint32_t>(in.cur_hi, m_clut_slots - 1);
  m_alloc_hand = (in.alloc_hand >= 1 && in.alloc_hand < m_clut_slots) ? in.alloc_hand : 1;
  for (uint32_t i = 1; i < m_clut_slots; i++)
  {
    const uint32_t k = in.slot_key[i];
    if (!(k & 0x80000000u))
      continue;
    ClutSlot& s = m_slots[i];
    s.x = static_cast<uint16_t>(k & 0x3FFu);
    s.y = static_cast<uint16_t>((k >> 10) & 0x1FFu);
    s.is8 = (k >> 19) & 1u;
    s.valid = true;
    chain_insert(i);
    RectSet src;
    src.add_wrapped(s.x, s.y, s.is8 ? 256u : 16u, 1);
    m_clut_sources.add(src);
  }
}

void Encoder::close_epoch(uint32_t kind)
{
  Epoch e;
  e.cmd_begin = m_epoch_cmd_begin;
  e.cmd_end = m_n_cmds;
  e.kind = kind;
  e.clut_begin = m_epoch_clut_begin;
  e.clut_end = static_cast<uint32_t>(clut_ops.size());
  epochs.push_back(e);
  stats.epochs++;
  m_epoch_cmd_begin = e.cmd_end;
  m_epoch_clut_begin = e.clut_end;
  m_writes.clear();
  m_reads.clear();
  m_epoch_serial++;
}

int Encoder::find_slot(uint32_t x, uint32_t y, bool is8) const
{
  // a valid 8-bit snapshot at the same position also serves a 4-bit load (its first 16 entries are the same pixels)
  for (uint32_t i = m_bucket[bucket_of(x, y)]; i != 0; i = m_chain[i])
  {
    const ClutSlot& s = m_slots[i];
    if (s.valid && s.x == x && s.y == y && (s.is8 || !is8))
      return static_cast<int>(i);
  }
  return -1;
}

uint32_t Encoder::alloc_slot()
{
  // A slot may be overwritten by this epoch's snapshot pass only if nothing in the open epoch samples its current
  // contents: snapshots of an epoch all run before its primitives.
  for (int pass = 0; pass < 2; pass++)
  {
    for (uint32_t n = 0; n + 1 < m_clut_slots; n++)
    {
      const uint32_t s = m_alloc_hand;
      m_alloc_hand = (m_alloc_hand + 1 >= m_clut_slots) ? 1 : (m_alloc_hand + 1);
      if (s == m_cur_lo || s == m_cur_hi)
        continue;
      if (m_slots[s].last_epoch == m_epoch_serial)
        continue;
      return s;
    }
    // every slot is in use by the open epoch: cut, which frees all but the current pair
    close_epoch(EPOCH_TILED);
    stats.cuts_clut++;
  }
  return (m_cur_lo == 1 || m_cur_hi == 1) ? ((m_cur_lo == 2 || m_cur_hi == 2) ? 3u : 2u) : 1u; // unreachable for >= 4 slots
}

void Encoder::invalidate_cluts(const RectSet& writes)
{
  bool any = false;
  for (uint32_t i = 1; i < m_clut_slots; i++)
  {
    ClutSlot& s = m_slots[i];
    if (!s.valid)
      continue;
    RectSet src;
    src.add_wrapped(s.x, s.y, s.is8 ? 256u : 16u, 1);
    if (src.intersects(writes))
    {
      s.valid = false; // contents stay in use by draws that already reference the slot
      chain_remove(i);
      stats.clut_invalidations++;
      any = true;
    }
  }
  if (any)
  {
    m_clut_sources.clear();
    for (uint32_t i = 1; i < m_clut_slots; i++)
    {
      const ClutSlot& s = m_slots[i];
      if (!s.valid)
        continue;
      RectSet src;
      src.add_wrapped(s.x, s.y, s.is8 ? 256u : 16u, 1);
      m_clut_sources.add(src);
    }
  }
}

void Encoder::update_clut(uint16_t reg, bool is8)
{
  const uint32_t x = PSX_PAL_X(reg), y = PSX_PAL_Y(reg);
  int slot = find_slot(x, y, is8);
  if (slot >= 0)
  {
    // still valid => nothing has written its source since the snapshot, so no ordering constraint either
    stats.clut_cache_hits++;
  }
  else
  {
    RectSet src;
    src.add_wrapped(x, y, is8 ? 256u : 16u, 1);
    // The snapshot is taken before the epoch rasterises, so only writes EARLIER in this epoch matter (RAW).
    if (m_writes.hits(src))
    {
      close_epoch(EPOCH_TILED);
      stats.cuts_clut++;
    }
    slot = static_cast<int>(alloc_slot());
    ClutSlot& s = m_slots[slot];
    if (s.valid)
      chain_remove(static_cast<uint32_t>(slot)); // evicted
    s.x = static_cast<uint16_t>(x);
    s.y = static_cast<uint16_t>(y);
    s.is8 = is8 ? 1 : 0;
    s.valid = true;
    chain_insert(static_cast<uint32_t>(slot));
    m_clut_sources.add(src);
    ClutOp op;
    op.dst_slot = static_cast<uint16_t>(slot);
    op.x = static_cast<uint16_t>(x);
    op.y = static_cast<uint16_t>(y);
    op.is8bit = is8 ? 1 : 0;
    clut_ops.push_back(op);
    stats.clut_ops++;
  }
  m_slots[slot].last_epoch = m_epoch_serial;
  m_cur_lo = static_cast<uint32_t>(slot);
  if (is8)
    m_cur_hi = static_cast<uint32_t>(slot);
}

void Encoder::set_output(PackedCmd* out_cmds, uint32_t cmd_capacity, uint16_t* out_payload, uint64_t payload_capacity)
{
  m_out_cmds = out_cmds;
  m_out_cmd_cap = cmd_capacity;
  m_out_payload = out_payload;
  m_out_pay_cap = payload_capacity;
}

void Encoder::push_cmd(const PackedCmd& c)
{
  if (m_out_cmds)
  {
    if (m_n_cmds < m_out_cmd_cap)
      m_out_cmds[m_n_cmds] = c;
    else
      m_overflow = true;
  }
  else
  {
    cmds.push_back(c);
  }
  m_n_cmds++;
}

static uint64_t record_bytes_needed(const uint8_t* rec, uint32_t size);

bool Encoder::scan_wire(const void* wire, size_t bytes, uint32_t* max_cmds, uint64_t* max_payload)
{
  const uint8_t* p = static_cast<const uint8_t*>(wire);
  const uint8_t* end = p + bytes;
  uint32_t nc = 0;
  uint64_t np = 0;
  while (p < end)
  {
    if (static_cast<size_t>(end - p) < sizeof(psx_cmd_header))
      return false;
    const psx_cmd_header* h = reinterpret_cast<const psx_cmd_header*>(p);
    if (h->size < sizeof(psx_cmd_header) || (h->size & 3u) != 0 || h->size > static_cast<size_t>(end - p))
      return false;
    if (record_bytes_needed(p, h->size) > h->size)
    {
      p += h->size; // handle_record drops it
      continue;
    }
    switch (h->type)
    {
      case PSX_CMD_DRAW_POLYGON:
      case PSX_CMD_DRAW_PRECISE_POLYGON: nc += (reinterpret_cast<const psx_draw_header*>(p)->num_vertices > 3) ? 2u : 1u; break;
      case PSX_CMD_DRAW_LINE:
      case PSX_CMD_DRAW_PRECISE_LINE: nc += reinterpret_cast<const psx_draw_header*>(p)->num_vertices / 2u; break;
      case PSX_CMD_DRAW_RECTANGLE:
      case PSX_CMD_FILL_VRAM:
      case PSX_CMD_COPY_VRAM: nc += 1; break;
      case PSX_CMD_UPDATE_VRAM:
      {
        const psx_update_vram* u = reinterpret_cast<const psx_update_vram*>(p);
        nc += 1;
        np += ((static_cast<uint64_t>(u->width) * u->height) + 7u) & ~static_cast<uint64_t>(7);
      }
      break;
      default: break;
    }
    p += h->size;
  }
  *max_cmds = nc;
  *max_payload = np;
  return true;
}

void Encoder::emit(const PackedCmd& c, const RectSet& writes, const RectSet& reads, uint32_t serial_kind)
{
  if (m_clut_sources.hits(writes))
    invalidate_cluts(writes);
  const bool textured = (c.op == OP_TRIANGLE || c.op == OP_RECT) && (c.flags0 & F_TEXTURE);
  if (serial_kind != EPOCH_TILED)
  {
    if (m_n_cmds > m_epoch_cmd_begin)
      close_epoch(EPOCH_TILED);
    if (textured)
      m_slots[c.clut_lo].last_epoch = m_slots[c.clut_hi].last_epoch = m_epoch_serial;
    push_cmd(c);
    stats.packed_prims++;
    close_epoch(serial_kind);
    return;
  }
  const bool raw = m_writes.hits(reads);
  const bool war = !raw && m_reads.hits(writes);
  const bool full = (m_n_cmds - m_epoch_cmd_begin) >= 65535u;
  if (raw || war || full)
  {
    close_epoch(EPOCH_TILED);
    if (raw)
      stats.cuts_raw++;
    else if (war)
      stats.cuts_war++;
  }
  if (textured)
    m_slots[c.clut_lo].last_epoch = m_slots[c.clut_hi].last_epoch = m_epoch_serial;
  push_cmd(c);
  stats.packed_prims++;
  m_writes.add(writes);
  m_reads.add(reads);
}

void Encoder::fill_draw_header(PackedCmd& c, const void* draw_header) const
{
  const psx_draw_header* d = static_cast<const psx_draw_header*>(draw_header);
  std::memset(&c, 0, sizeof(c));
  c.flags0 = d->flags0;
  c.flags1 = static_cast<uint8_t>(((d->flags1 & PSX_DF1_DITHER) ? F1_DITHER : 0) |
                                  ((m_modulation_crop && (d->flags0 & PSX_DF_TEXTURE) && !(d->flags0 & PSX_DF_RAW_TEXTURE)) ? F1_MOD5 : 0));
  c.draw_mode = d->draw_mode;
  c.clut_lo = static_cast<uint8_t>(m_cur_lo);
  c.clut_hi = static_cast<uint8_t>(m_cur_hi);
  c.win_and_x = d->window.and_x;
  c.win_and_y = d->window.and_y;
  c.win_or_x = d->window.or_x;
  c.win_or_y = d->window.or_y;
  c.clip_l = m_area[0];
  c.clip_t = m_area[1];
  c.clip_r = m_area[2];
  c.clip_b = m_area[3];
}

// VRAM rectangles a textured primitive may sample, from its (pre-window) texcoord range.
void Encoder::texture_footprint(const PackedCmd& c, uint32_t umin, uint32_t umax, uint32_t vmin, uint32_t vmax, RectSet& reads) const
{
  // texture window: every windowed coordinate lies in [or, or | and]
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
  reads.add_wrapped(bx + (umin >> sh), by + vmin, (umax >> sh) - (umin >> sh) + 1, vmax - vmin + 1);
}

void Encoder::draw_triangle(const void* draw_header, const PackedVertex& a, const PackedVertex& b, const PackedVertex& c3)
{
  PackedCmd c;
  fill_draw_header(c, draw_header);
  c.op = OP_TRIANGLE;
  c.v[0] = a;
  c.v[1] = b;
  c.v[2] = c3;
  const int32_t minx = std::min({a.x, b.x, c3.x}), maxx = std::max({a.x, b.x, c3.x});
  const int32_t miny = std::min({a.y, b.y, c3.y}), maxy = std::max({a.y, b.y, c3.y});
  if ((maxx - minx) >= static_cast<int32_t>(VRAM_W) || (maxy - miny) >= static_cast<int32_t>(VRAM_H))
  {
    stats.rejected++; // the decoder culls these (gpu.cpp:3283-3350); a backend never sees them
    return;
  }
  if (miny == maxy)
    return; // v0.y == v2.y: rejected by the rasteriser (inl:1459)
  // same bounding box as setup_triangle (psx_setup.h)
  int32_t x0, x1, y0, y1;
  if (minx > -1023 && maxx < 1023)
  {
    x0 = std::max<int32_t>(minx - 1, c.clip_l);
    x1 = std::min<int32_t>(maxx + 1, c.clip_r);
  }
  else
  {
    x0 = c.clip_l;
    x1 = c.clip_r;
  }
  if (miny >= -1024 && maxy <= 1024)
  {
    y0 = std::max<int32_t>(miny, c.clip_t);
    y1 = std::min<int32_t>(maxy - 1, c.clip_b);
  }
  else
  {
    y0 = 0;
    y1 = static_cast<int32_t>(VRAM_H) - 1;
  }
  if (x1 < x0 || y1 < y0)
    return; // nothing can be drawn
  RectSet writes, reads;
  writes.add_wrapped(static_cast<uint32_t>(x0), static_cast<uint32_t>(y0), static_cast<uint32_t>(x1 - x0 + 1),
                     static_cast<uint32_t>(y1 - y0 + 1));
  uint32_t kind = EPOCH_TILED;
  if (c.flags0 & F_TEXTURE)
  {
    uint32_t umin = std::min({a.u, b.u, c3.u}), umax = std::max({a.u, b.u, c3.u});
    uint32_t vmin = std::min({a.v, b.v, c3.v}), vmax = std::max({a.v, b.v, c3.v});
    // interpolated coordinates are u8 and may wrap by one step at the extremes
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
    texture_footprint(c, umin, umax, vmin, vmax, reads);
    if (reads.intersects(writes))
    {
      kind = EPOCH_SELF_SAMPLING;
      stats.serial_self_sampling++;
    }
  }
  emit(c, writes, reads, kind);
}

static inline bool vertex_ok(int32_t x, int32_t y)
{
  return x >= -32768 && x <= 32767 && y >= -32768 && y <= 32767;
}

// Bytes a record of this type must hold before any of its fields is read (its fixed part plus, for polygons, lines
// and uploads, what its own counts ask for).  `size` = the record's framed size.  0 = nothing to check.
// The same rules as the device encoder (encode_kernel.cu: every `size <` test): a record that is too short is dropped
// — counted as rejected if it is a primitive or a transfer, ignored if it is a state record — never read past.
static uint64_t record_bytes_needed(const uint8_t* rec, uint32_t size)
{
  const psx_cmd_header* h = reinterpret_cast<const psx_cmd_header*>(rec);
  switch (h->type)
  {
    case PSX_CMD_SET_DRAWING_AREA: return sizeof(psx_set_drawing_area);
    case PSX_CMD_UPDATE_CLUT: return sizeof(psx_update_clut);
    case PSX_CMD_DRAW_RECTANGLE: return sizeof(psx_draw_rectangle);
    case PSX_CMD_FILL_VRAM: return sizeof(psx_fill_vram);
    case PSX_CMD_COPY_VRAM: return sizeof(psx_copy_vram);
    case PSX_CMD_DRAW_POLYGON:
    case PSX_CMD_DRAW_PRECISE_POLYGON:
    case PSX_CMD_DRAW_LINE:
    case PSX_CMD_DRAW_PRECISE_LINE:
    {
      if (size < sizeof(psx_draw_header))
        return sizeof(psx_draw_header);
      const uint64_t nv = reinterpret_cast<const psx_draw_header*>(rec)->num_vertices;
      if (h->type == PSX_CMD_DRAW_POLYGON)
        return sizeof(psx_draw_header) + sizeof(psx_poly_vertex) * (nv < 4 ? nv : 4);
      if (h->type == PSX_CMD_DRAW_PRECISE_POLYGON)
        return offsetof(psx_draw_precise_polygon, vertices) + sizeof(psx_precise_poly_vertex) * (nv < 4 ? nv : 4);
      return sizeof(psx_draw_header) + (h->type == PSX_CMD_DRAW_LINE ? sizeof(psx_line_vertex) : sizeof(psx_precise_line_vertex)) * nv;
    }
    case PSX_CMD_UPDATE_VRAM:
    {
      if (size < 20)
        return 20;
      const psx_update_vram* u = reinterpret_cast<const psx_update_vram*>(rec);
      return PSX_UPDATE_VRAM_DATA_OFFSET + static_cast<uint64_t>(u->width) * u->height * 2u;
    }
    default: return 0;
  }
}

void Encoder::handle_record(const uint8_t* rec)
{
  const psx_cmd_header* h = reinterpret_cast<const psx_cmd_header*>(rec);
  stats.wire_commands++;
  if (record_bytes_needed(rec, h->size) > h->size)
  {
    if (h->type != PSX_CMD_SET_DRAWING_AREA && h->type != PSX_CMD_UPDATE_CLUT)
      stats.rejected++;
    return;
  }
  switch (h->type)
  {
    case PSX_CMD_SET_DRAWING_AREA:
    {
      const psx_set_drawing_area* c = reinterpret_cast<const psx_set_drawing_area*>(rec);
      m_area[0] = static_cast<uint16_t>(std::min<uint32_t>(c->new_area.left, 2047));
      m_area[1] = static_cast<uint16_t>(std::min<uint32_t>(c->new_area.top, 2047));
      m_area[2] = static_cast<uint16_t>(std::min<uint32_t>(c->new_area.right, 1023));
      m_area[3] = static_cast<uint16_t>(std::min<uint32_t>(c->new_area.bottom, 1023));
    }
    break;

    case PSX_CMD_UPDATE_CLUT:
    {
      const psx_update_clut* c = reinterpret_cast<const psx_update_clut*>(rec);
      update_clut(c->reg, c->clut_is_8bit != 0);
    }
    break;

    case PSX_CMD_DRAW_POLYGON:
    case PSX_CMD_DRAW_PRECISE_POLYGON:
    {
      const psx_draw_header* d = reinterpret_cast<const psx_draw_header*>(rec);
      PackedVertex v[4];
      const uint32_t nv = std::min<uint32_t>(d->num_vertices, 4);
      bool ok = nv >= 3;
      for (uint32_t i = 0; i < nv && ok; i++)
      {
        int32_t x, y;
        uint32_t color;
        uint16_t tc;
        if (h->type == PSX_CMD_DRAW_POLYGON)
        {
          const psx_poly_vertex& s = reinterpret_cast<const psx_draw_polygon*>(rec)->vertices[i];
          x = s.x;
          y = s.y;
          color = s.r | (s.g << 8) | (s.b << 16);
          tc = static_cast<uint16_t>(s.u | (s.v << 8));
        }
        else
        {
          const psx_precise_poly_vertex& s = reinterpret_cast<const psx_draw_precise_polygon*>(rec)->vertices[i];
          x = s.native_x;
          y = s.native_y;
          color = s.color;
          tc = s.texcoord;
        }
        ok = vertex_ok(x, y);
        v[i].x = static_cast<int16_t>(x);
        v[i].y = static_cast<int16_t>(y);
        v[i].r = static_cast<uint8_t>(color);
        v[i].g = static_cast<uint8_t>(color >> 8);
        v[i].b = static_cast<uint8_t>(color >> 16);
        v[i].u = static_cast<uint8_t>(tc);
        v[i].v = static_cast<uint8_t>(tc >> 8);
        v[i].pad = 0;
      }
      if (!ok)
      {
        stats.rejected++;
        break;
      }
      // gpu_sw.cpp:113-115: (v0,v1,v2) then (v2,v1,v3)
      draw_triangle(d, v[0], v[1], v[2]);
      if (nv > 3)
        draw_triangle(d, v[2], v[1], v[3]);
    }
    break;

    case PSX_CMD_DRAW_RECTANGLE:
    {
      const psx_draw_rectangle* r = reinterpret_cast<const psx_draw_rectangle*>(rec);
      if (!vertex_ok(r->x, r->y) || r->width > VRAM_W || r->height > VRAM_H)
      {
        stats.rejected++;
        break;
      }
      PackedCmd c;
      fill_draw_header(c, &r->d);
      c.op = OP_RECT;
      c.rect.x = static_cast<int16_t>(r->x);
      c.rect.y = static_cast<int16_t>(r->y);
      c.rect.w = r->width;
      c.rect.h = r->height;
      c.rect.r = static_cast<uint8_t>(r->color);
      c.rect.g = static_cast<uint8_t>(r->color >> 8);
      c.rect.b = static_cast<uint8_t>(r->color >> 16);
      c.rect.u0 = static_cast<uint8_t>(r->texcoord);
      c.rect.v0 = static_cast<uint8_t>(r->texcoord >> 8);
      const int32_t x0 = std::max<int32_t>(r->x, c.clip_l), x1 = std::min<int32_t>(r->x + r->width - 1, c.clip_r);
      const int32_t y0 = std::max<int32_t>(r->y, c.clip_t), y1 = std::min<int32_t>(r->y + r->height - 1, c.clip_b);
      if (x1 < x0 || y1 < y0)
        break;
      RectSet writes, reads;
      writes.add_wrapped(static_cast<uint32_t>(x0), static_cast<uint32_t>(y0), static_cast<uint32_t>(x1 - x0 + 1),
                         static_cast<uint32_t>(y1 - y0 + 1));
      uint32_t kind = EPOCH_TILED;
      if (c.flags0 & F_TEXTURE)
      {
        uint32_t umin = c.rect.u0, umax = c.rect.u0 + r->width - 1, vmin = c.rect.v0, vmax = c.rect.v0 + r->height - 1;
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
        texture_footprint(c, umin, umax, vmin, vmax, reads);
        if (reads.intersects(writes))
        {
          kind = EPOCH_SELF_SAMPLING;
          stats.serial_self_sampling++;
        }
      }
      emit(c, writes, reads, kind);
    }
    break;

    case PSX_CMD_DRAW_LINE:
    case PSX_CMD_DRAW_PRECISE_LINE:
    {
      const psx_draw_header* d = reinterpret_cast<const psx_draw_header*>(rec);
      for (uint32_t i = 0; i + 1 < d->num_vertices; i += 2)
      {
        PackedVertex p[2];
        bool ok = true;
        for (int j = 0; j < 2; j++)
        {
          int32_t x, y;
          uint32_t color;
          if (h->type == PSX_CMD_DRAW_LINE)
          {
            const psx_line_vertex& s = reinterpret_cast<const psx_draw_line*>(rec)->vertices[i + j];
            x = s.x;
            y = s.y;
            color = s.r | (s.g << 8) | (s.b << 16);
          }
          else
          {
            const psx_precise_line_vertex& s =
              reinterpret_cast<const psx_precise_line_vertex*>(rec + sizeof(psx_draw_header))[i + j];
            x = s.native_x;
            y = s.native_y;
            color = s.color;
          }
          ok = ok && vertex_ok(x, y);
          p[j].x = static_cast<int16_t>(x);
          p[j].y = static_cast<int16_t>(y);
          p[j].r = static_cast<uint8_t>(color);
          p[j].g = static_cast<uint8_t>(color >> 8);
          p[j].b = static_cast<uint8_t>(color >> 16);
          p[j].u = p[j].v = p[j].pad = 0;
        }
        if (!ok)
        {
          stats.rejected++;
          continue;
        }
        if (std::abs(p[1].x - p[0].x) >= static_cast<int32_t>(VRAM_W) || std::abs(p[1].y - p[0].y) >= static_cast<int32_t>(VRAM_H))
          continue; // rasteriser rejects (inl:834-835)
        PackedCmd c;
        fill_draw_header(c, d);
        c.flags0 &= static_cast<uint8_t>(~(F_TEXTURE | F_RAW)); // lines are never textured (gpu.cpp:2974)
        c.flags1 &= static_cast<uint8_t>(~F1_MOD5);
        c.op = OP_LINE;
        c.v[0] = p[0];
        c.v[1] = p[1];
        const int32_t minx = std::min(p[0].x, p[1].x), maxx = std::max(p[0].x, p[1].x);
        const int32_t miny = std::min(p[0].y, p[1].y), maxy = std::max(p[0].y, p[1].y);
        int32_t x0 = c.clip_l, x1 = c.clip_r, y0 = 0, y1 = static_cast<int32_t>(VRAM_H) - 1;
        if (minx >= -1024 && maxx <= 1023 && miny >= -1024 && maxy <= 1023)
        {
          x0 = std::max<int32_t>(minx, c.clip_l);
          x1 = std::min<int32_t>(maxx, c.clip_r);
          y0 = std::max<int32_t>(miny, c.clip_t);
          y1 = std::min<int32_t>(maxy, c.clip_b);
        }
        if (x1 < x0 || y1 < y0)
          continue;
        RectSet writes, reads;
        writes.add_wrapped(static_cast<uint32_t>(x0), static_cast<uint32_t>(y0), static_cast<uint32_t>(x1 - x0 + 1),
                           static_cast<uint32_t>(y1 - y0 + 1));
        emit(c, writes, reads, EPOCH_TILED);
      }
    }
    break;

    case PSX_CMD_FILL_VRAM:
    {
      const psx_fill_vram* f = reinterpret_cast<const psx_fill_vram*>(rec);
      if (f->width > VRAM_W || f->height > VRAM_H || f->x >= VRAM_W || f->y >= VRAM_H)
      {
        stats.rejected++;
        break;
      }
      if (f->width == 0 || f->height == 0)
        break;
      PackedCmd c;
      std::memset(&c, 0, sizeof(c));
      c.op = OP_FILL;
      c.fill.x = f->x;
      c.fill.y = f->y;
      c.fill.w = f->width;
      c.fill.h = f->height;
      // VRAMRGBA8888ToRGBA5551, gpu_helpers.h:32-39
      c.fill.color16 = static_cast<uint16_t>(((f->color & 0xFFu) >> 3) | ((((f->color >> 8) & 0xFFu) >> 3) << 5) |
                                             ((((f->color >> 16) & 0xFFu) >> 3) << 10) | (((f->color >> 24) & 1u) << 15));
      c.fill.interlaced = f->interlaced_rendering ? 1 : 0;
      c.fill.active_lsb = f->active_line_lsb & 1u;
      RectSet writes, reads;
      writes.add_wrapped(f->x, f->y, f->width, f->height);
      emit(c, writes, reads, EPOCH_TILED);
    }
    break;

    case PSX_CMD_COPY_VRAM:
    {
      const psx_copy_vram* p = reinterpret_cast<const psx_copy_vram*>(rec);
      if (p->width > VRAM_W || p->height > VRAM_H || p->src_x >= VRAM_W || p->dst_x >= VRAM_W || p->src_y >= VRAM_H ||
          p->dst_y >= VRAM_H)
      {
        stats.rejected++;
        break;
      }
      if (p->width == 0 || p->height == 0)
        break;
      PackedCmd c;
      std::memset(&c, 0, sizeof(c));
      c.op = OP_COPY;
      c.flags0 = static_cast<uint8_t>((p->set_mask_while_drawing ? F_SET_MASK : 0) | (p->check_mask_before_draw ? F_CHECK_MASK : 0));
      c.copy.sx = p->src_x;
      c.copy.sy = p->src_y;
      c.copy.dx = p->dst_x;
      c.copy.dy = p->dst_y;
      c.copy.w = p->width;
      c.copy.h = p->height;
      RectSet writes, reads;
      writes.add_wrapped(p->dst_x, p->dst_y, p->width, p->height);
      reads.add_wrapped(p->src_x, p->src_y, p->width, p->height);
      uint32_t kind = EPOCH_TILED;
      if (reads.intersects(writes))
      {
        kind = EPOCH_COPY_OVERLAP;
        stats.serial_copy++;
      }
      emit(c, writes, reads, kind);
    }
    break;

    case PSX_CMD_UPDATE_VRAM:
    {
      const psx_update_vram* u = reinterpret_cast<const psx_update_vram*>(rec);
      const size_t need = PSX_UPDATE_VRAM_DATA_OFFSET + static_cast<size_t>(u->width) * u->height * 2u;
      if (u->width > VRAM_W || u->height > VRAM_H || u->x >= VRAM_W || u->y >= VRAM_H || need > h->size)
      {
        stats.rejected++;
        break;
      }
      if (u->width == 0 || u->height == 0)
        break;
      PackedCmd c;
      std::memset(&c, 0, sizeof(c));
      c.op = OP_UPLOAD;
      c.flags0 = static_cast<uint8_t>((u->set_mask_while_drawing ? F_SET_MASK : 0) | (u->check_mask_before_draw ? F_CHECK_MASK : 0));
      c.upload.x = u->x;
      c.upload.y = u->y;
      c.upload.w = u->width;
      c.upload.h = u->height;
      c.upload.payload = static_cast<uint32_t>(m_n_payload);
      const uint16_t* px = reinterpret_cast<const uint16_t*>(rec + PSX_UPDATE_VRAM_DATA_OFFSET);
      const uint64_t npx = static_cast<uint64_t>(u->width) * u->height;
      const uint64_t padded = (npx + 7u) & ~static_cast<uint64_t>(7); // 8-pixel aligned so device reads can be vectorised
      if (m_out_payload)
      {
        if (m_n_payload + padded <= m_out_pay_cap)
        {
          std::memcpy(m_out_payload + m_n_payload, px, npx * 2);
          std::memset(m_out_payload + m_n_payload + npx, 0, (padded - npx) * 2);
        }
        else
        {
          m_overflow = true;
        }
      }
      else
      {
        payload.insert(payload.end(), px, px + npx);
        payload.resize(payload.size() + (padded - npx), 0);
      }
      m_n_payload += padded;
      stats.upload_pixels += static_cast<uint64_t>(u->width) * u->height;
      RectSet writes, reads;
      writes.add_wrapped(u->x, u->y, u->width, u->height);
      // Q2: with mask checking, source pixels are consumed only by written pixels in the scalar tail of each row
      // (inl:1822-1828); without a tail the mapping is the identity and the upload stays tile-parallel.
      const uint32_t aw = std::min<uint32_t>(u->width, VRAM_W - u->x) & ~7u;
      uint32_t kind = EPOCH_TILED;
      if (u->check_mask_before_draw && aw != u->width)
      {
        kind = EPOCH_UPLOAD_MASKED;
        stats.serial_upload++;
      }
      upload_cmds.push_back(m_n_cmds); // emit() appends exactly one record
      emit(c, writes, reads, kind);
    }
    break;

    default:
      break; // ClearCache / BufferSwapped / SubmitFrame / ClearDisplay: no VRAM effect (gpu_sw.cpp:192-211)
  }
}

static bool is_barrier(uint8_t type)
{
  switch (type)
  {
    case PSX_CMD_CLEAR_VRAM:
    case PSX_CMD_LOAD_STATE:
    case PSX_CMD_READ_VRAM:
    case PSX_CMD_UPDATE_DISPLAY:
      return true;
    default:
      return false;
  }
}

EncodeResult Encoder::add_wire(const void* wire, size_t bytes, size_t* consumed)
{
  const uint8_t* p = static_cast<const uint8_t*>(wire);
  const uint8_t* end = p + bytes;
  while (p < end)
  {
    if (static_cast<size_t>(end - p) < sizeof(psx_cmd_header))
    {
      *consumed = static_cast<size_t>(p - static_cast<const uint8_t*>(wire));
      return EncodeResult::Malformed;
    }
    const psx_cmd_header* h = reinterpret_cast<const psx_cmd_header*>(p);
    if (h->size < sizeof(psx_cmd_header) || (h->size & 3u) != 0 || h->size > static_cast<size_t>(end - p))
    {
      *consumed = static_cast<size_t>(p - static_cast<const uint8_t*>(wire));
      return EncodeResult::Malformed;
    }
    if (is_barrier(h->type))
    {
      *consumed = static_cast<size_t>(p - static_cast<const uint8_t*>(wire));
      return EncodeResult::Barrier;
    }
    handle_record(p);
    p += h->size;
    if (m_cmd_budget != 0 && m_n_cmds >= m_cmd_budget)
    {
      // the caller wants to flush every so many primitives (psxgpu_submit_wire: the device starts on the first part of a
      // long stream while the host encodes the rest)
      *consumed = static_cast<size_t>(p - static_cast<const uint8_t*>(wire));
      return EncodeResult::Ok;
    }
  }
  *consumed = bytes;
  return EncodeResult::Ok;
}

} // namespace psx
