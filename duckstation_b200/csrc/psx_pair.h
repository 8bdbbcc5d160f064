// psx_pair.h — the PSX pixel pipeline on TWO horizontally adjacent pixels held in one 32-bit register.
//
// A VRAM word is two 15+1-bit pixels (even column in bits 0-15, odd column in bits 16-31).  The raster kernel gives each
// lane one such word of its tile; everything after the texel fetch works on both halves at once:
//   * modulation       texel channel (5 bit) x colour (8 bit): one IMAD per channel when the colour is the same for both
//                      pixels (flat primitives, sprites), products < 2^13 so the halves never meet
//   * dither + clamp   one VIADDMNMX.S16x2.RELU per channel: clamp(p + 16 d, 0, 4095) >> 7 == clamp((p >> 4) + d, 0, 255) >> 3
//                      because 16 d is a multiple of 16 (the low four bits of p ride along and are shifted out)
//   * blending         the reference's packed 15-bit formulas (gpu_sw_rasterizer.inl:182-222) applied to both halves
//                      of the register; bit 15 is stripped first so that no carry crosses bit 16, and put back as what
//                      the reference produces (== the texel's bit 15 | set-mask, see pair_shade)
//   * mask test / transparency / coverage: masks at bit 15 of each half, widened to 0xFFFF with one PRMT
// Value-for-value the same function as shade_colour + blend_mask of psx_pixel.h on each half (tests/hostsim checks that
// on random and exhaustive inputs); host-compilable for that test, the product only calls it from device code.
#pragma once

#include "psx_pixel.h"

namespace psx {

// ---- 16x2 helpers (one instruction each on sm_100a) ----
PSX_HD uint32_t pk_addmin_relu(uint32_t a, uint32_t b, uint32_t c) // per half, signed: max(min(a + b, c), 0)
{
#if defined(__CUDA_ARCH__)
  return __viaddmin_s16x2_relu(a, b, c);
#else
  uint32_t out = 0;
  for (int h = 0; h < 2; h++)
  {
    int32_t t = static_cast<int16_t>(a >> (16 * h)) + static_cast<int16_t>(b >> (16 * h));
    const int32_t lim = static_cast<int16_t>(c >> (16 * h));
    t = t > lim ? lim : t;
    t = t < 0 ? 0 : t;
    out |= (static_cast<uint32_t>(t) & 0xFFFFu) << (16 * h);
  }
  return out;
#endif
}
// prmt.b32 with the full selector: bit 3 of a selector nibble replicates the sign of the chosen byte (__byte_perm masks
// the selector to three bits per nibble).
template<uint32_t kSel>
PSX_HD uint32_t pk_prmt(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "n"(kSel));
  return d;
#else
  const uint64_t src = (static_cast<uint64_t>(b) << 32) | a;
  uint32_t d = 0;
  for (uint32_t i = 0; i < 4; i++)
  {
    const uint32_t sel = (kSel >> (4 * i)) & 0xFu;
    uint32_t byte = static_cast<uint32_t>(src >> (8 * (sel & 7u))) & 0xFFu;
    if (sel & 8u)
      byte = (byte & 0x80u) ? 0xFFu : 0u;
    d |= byte << (8 * i);
  }
  return d;
#endif
}
PSX_HD uint32_t pk_expand15(uint32_t m) // bit 15 of each half -> 0xFFFF / 0x0000 in that half
{
  return pk_prmt<0xBB99>(m, 0);
}
PSX_HD uint32_t pk_dup_byte(uint32_t v, uint32_t byte) // byte `byte` of v in both halves, zero-extended
{
  const uint32_t b = (v >> (8u * byte)) & 0xFFu;
  return b | (b << 16);
}

// Dither offsets x 16 for a pixel pair: row y & 3, pair starting at column x & 3 in {0, 2}; entries 8..15 = no dither.
// gpu_types.h:269-272 {-4,0,-3,1} {2,-2,3,-1} {-3,1,-4,0} {3,-1,2,-2}
PSX_HD uint32_t pair_dither_entry(uint32_t i)
{
  if (i >= 8)
    return 0;
  const uint32_t y = i >> 1, x = (i & 1u) * 2u;
  const uint32_t d0 = static_cast<uint32_t>(dither_offset(x, y) * 16) & 0xFFFFu;
  const uint32_t d1 = static_cast<uint32_t>(dither_offset(x + 1, y) * 16) & 0xFFFFu;
  return d0 | (d1 << 16);
}

// What the pair pipeline needs from a primitive (warp-uniform).
struct PairState
{
  uint32_t flags;  // flags0 | flags1 << 8 (as ShadeState::flags)
  uint32_t tmode;  // semi-transparency mode 0..3
  uint32_t chk15;  // 0x80008000 if the mask bit is tested
  uint32_t set15;  // 0x80008000 if the mask bit is set
};
PSX_HD PairState make_pair_state(const ShadeState& st)
{
  PairState ps;
  ps.flags = st.flags;
  ps.tmode = (st.modes >> 8) & 3u;
  ps.chk15 = (st.flags & F_CHECK_MASK) ? 0x80008000u : 0u;
  ps.set15 = (st.flags & F_SET_MASK) ? 0x80008000u : 0u;
  return ps;
}

// The four blend equations on two 15-bit pixels (bit 15 of every half clear on entry and on return).
// Derived from the reference's packed forms; exhaustively equal to them on the low 15 bits (tests/hostsim).
PSX_HD uint32_t pair_blend15(uint32_t mode, uint32_t b, uint32_t f)
{
  switch (mode)
  {
    case 0: // B/2 + F/2: per channel (b + f) >> 1; the subtraction clears every channel's low bit, bit 16 included
      return ((f + b) - ((f ^ b) & 0x04210421u)) >> 1;
    case 2: // B - F, saturating: SWAR subtract with the channel top bits as borrow guards
    {
      const uint32_t H = 0x42104210u;
      const uint32_t t = (b | H) - (f & ~H);
      const uint32_t z = t ^ ((b ^ ~f) & H);
      const uint32_t bor = ((~b & f) | (~(b ^ f) & z)) & H;
      return z & ~((bor << 1) - (bor >> 4));
    }
    case 3: f = (f >> 2) & 0x1CE71CE7u; // B + F/4
      // fallthrough
    default: // 1: B + F, saturating
    {
      const uint32_t sum = f + b;
      const uint32_t carry = (sum - ((f ^ b) & 0x04210421u)) & 0x84208420u;
      return (sum - carry) | (carry - (carry >> 5));
    }
  }
}

// 15-bit colour of two pixels from channel products p = 16 x (8-bit value) + don't-care low bits, and the pair's
// dither word (pair_dither_entry).
PSX_HD uint32_t pair_pack15(uint32_t pr, uint32_t pg, uint32_t pb, uint32_t d16)
{
  const uint32_t vr = pk_addmin_relu(pr, d16, 0x0FFF0FFFu);
  const uint32_t vg = pk_addmin_relu(pg, d16, 0x0FFF0FFFu);
  const uint32_t vb = pk_addmin_relu(pb, d16, 0x0FFF0FFFu);
  return ((vr >> 7) & 0x001F001Fu) | ((vg >> 2) & 0x03E003E0u) | ((vb << 3) & 0x7C007C00u);
}

// Colour of a textured, modulated pair.  kFlat: cr/cg/cb are plain 8-bit values shared by both pixels; otherwise they
// are pairs (pixel 0 in bits 0-7, pixel 1 in bits 16-23).
template<bool kFlat>
PSX_HD uint32_t pair_modulate(uint32_t T, uint32_t cr, uint32_t cg, uint32_t cb, uint32_t d16)
{
  const uint32_t tr = T & 0x001F001Fu, tg = (T >> 5) & 0x001F001Fu, tb = (T >> 10) & 0x001F001Fu;
  uint32_t pr, pg, pb;
  if (kFlat)
  {
    pr = tr * cr;
    pg = tg * cg;
    pb = tb * cb;
  }
  else
  {
    // x * c0 + hi(x) * (c1 - c0): the low product stays below 2^13, so the high half ends up as hi(x) * c1
    pr = tr * (cr & 0xFFu) + (tr & 0xFFFF0000u) * ((cr >> 16) - (cr & 0xFFu));
    pg = tg * (cg & 0xFFu) + (tg & 0xFFFF0000u) * ((cg >> 16) - (cg & 0xFFu));
    pb = tb * (cb & 0xFFu) + (tb & 0xFFFF0000u) * ((cb >> 16) - (cb & 0xFFu));
  }
  return pair_pack15(pr, pg, pb, d16);
}

// Everything after the colour: transparency, blending, mask test / set, coverage.  `col15` = the pair's 15-bit colours,
// `t15` = bit 15 of the texels (0 for untextured primitives), `we15` = bit 15 of a half set if that pixel is covered,
// its texel is not the transparent one, and (with the mask test on) the background's mask bit is clear.
// Bit 15 of a written pixel: the reference ends with texel bit 15 | set-mask for textured pixels whether or not they
// were blended (a blended textured pixel has texel bit 15 set and every blend equation leaves bit 15 set), and with
// set-mask alone for untextured ones (blend_mask clears bit 15 after blending).
PSX_HD uint32_t pair_finish(const PairState& ps, uint32_t bg, uint32_t col15, uint32_t t15, uint32_t we15)
{
  if (ps.flags & F_TRANSP)
  {
    const uint32_t blended = pair_blend15(ps.tmode, bg & 0x7FFF7FFFu, col15);
    const uint32_t bm = (ps.flags & F_TEXTURE) ? pk_expand15(t15) : 0xFFFFFFFFu;
    col15 = (col15 & ~bm) | (blended & bm);
  }
  const uint32_t wm = pk_expand15(we15);
  return (bg & ~wm) | ((col15 | t15 | ps.set15) & wm);
}

// Write-enable bits of a textured pair: covered & texel != 0 & !(mask test & background mask bit).
PSX_HD uint32_t pair_we_textured(const PairState& ps, uint32_t bg, uint32_t T, uint32_t cov15)
{
  const uint32_t nz = ((T & 0x7FFF7FFFu) + 0x7FFF7FFFu) | T; // bit 15 of a half = that texel is not 0
  return cov15 & nz & ~(bg & ps.chk15);
}
PSX_HD uint32_t pair_we_untextured(const PairState& ps, uint32_t bg, uint32_t cov15)
{
  return cov15 & ~(bg & ps.chk15);
}

// One pair through the whole pipeline, general form (what the kernel's specialised paths must equal).
// cr/cg/cb: colour pairs (bits 0-7 / 16-23); T: texel pair; d16: pair_dither_entry of the position (0 if no dither).
PSX_HD uint32_t pair_shade(const PairState& ps, uint32_t bg, uint32_t cov15, uint32_t T, uint32_t cr, uint32_t cg, uint32_t cb,
                           uint32_t d16)
{
  uint32_t col15, t15 = 0, we15;
  if (ps.flags & F_TEXTURE)
  {
    we15 = pair_we_textured(ps, bg, T, cov15);
    t15 = T & 0x80008000u;
    if (ps.flags & F_RAW)
      col15 = T & 0x7FFF7FFFu;
    else
    {
      if (ps.flags & (F1_MOD5 << 8))
      {
        cr &= 0x00F800F8u; // Modulate5Bit: (t * (c >> 3)) >> 1 == (t * (c & 0xF8)) >> 4
        cg &= 0x00F800F8u;
        cb &= 0x00F800F8u;
      }
      col15 = pair_modulate<false>(T, cr, cg, cb, d16);
    }
  }
  else
  {
    we15 = pair_we_untextured(ps, bg, cov15);
    col15 = pair_pack15(cr << 4, cg << 4, cb << 4, d16);
  }
  return pair_finish(ps, bg, col15, t15, we15);
}

} // namespace psx
