"""Test infrastructure: an independent GP0/GP1 *assembler*.

It goes the opposite way of duckstation_b200/csrc/gp0_frontend.cpp — from backend command records (the wire format the
generator emits, whose rendering is pinned against the compiled reference) to the raw GP0 port words a PSX program
would have written to get them (command encodings per the public PSX GPU documentation, "GPU Render/Memory Transfer
Commands").  Round trip: records -> words -> front-end -> records' ; the oracle must render both streams identically.

Also writes `.psxgpu` trace containers (the layout the reference's src/core/gpu_dump.cpp documents: 14-byte magic,
then packets of `length | type << 24` followed by `length` words).
"""
from __future__ import annotations

import struct

import numpy as np

CMD_CLEAR_VRAM, CMD_UPDATE_DISPLAY, CMD_LOAD_STATE, CMD_READ_VRAM = 7, 9, 12, 15
CMD_FILL, CMD_UPLOAD, CMD_COPY, CMD_AREA, CMD_CLUT, CMD_CLEAR_CACHE = 16, 17, 18, 19, 20, 21
CMD_POLY, CMD_PRECISE_POLY, CMD_RECT, CMD_LINE, CMD_PRECISE_LINE = 22, 23, 24, 25, 26

DF_INTERLACED, DF_LSB, DF_SET_MASK, DF_CHECK_MASK, DF_TEXTURE, DF_RAW, DF_SEMI, DF_SHADED = (1 << i for i in range(8))
DF1_QUAD, DF1_DITHER = 1, 2


class Unsupported(Exception):
    """The record cannot be produced by GP0 writes (or is outside what this assembler models)."""


def _pos(x: int, y: int) -> int:
    return (x & 0x7FF) | ((y & 0x7FF) << 16)


class Gp0Assembler:
    """Mirrors the GPU state a program has to set up before each primitive and emits only the state writes needed."""

    def __init__(self):
        self.events: list[tuple] = []  # ("gp0", [words]) | ("gp1", word) | ("vsync",) | ("finish_read",)
        self.draw_mode = 0
        self.window = (0xFF, 0xFF, 0, 0)
        self.area = (0, 0, 0, 0)
        self.offset = (0, 0)
        self.mask = (0, 0)
        self.clut_reg = None
        self.clut8 = False
        # display registers (only used when UpdateDisplay records are translated, see add_record)
        self.display = False
        self.disp_mode = None
        self.disp_addr = None

    # ---------------------------------------------------------------- raw emitters
    def gp0(self, *words: int) -> None:
        ws = [w & 0xFFFFFFFF for w in words]
        if self.events and self.events[-1][0] == "gp0":
            self.events[-1][1].extend(ws)
        else:
            self.events.append(("gp0", ws))

    def gp1(self, word: int) -> None:
        self.events.append(("gp1", word & 0xFFFFFFFF))

    def vsync(self) -> None:
        self.events.append(("vsync",))

    def finish_read(self) -> None:
        self.events.append(("finish_read",))

    def words(self) -> list[int]:
        out = []
        for e in self.events:
            if e[0] == "gp0":
                out.extend(e[1])
        return out

    # ---------------------------------------------------------------- state
    def _set_mask(self, set_mask: int, check_mask: int) -> None:
        if self.mask != (set_mask, check_mask):
            self.gp0(0xE6000000 | set_mask | (check_mask << 1))
            self.mask = (set_mask, check_mask)

    def _set_offset_for(self, xs, ys) -> None:
        ox, oy = self.offset

        def fit(o, vs):
            lo, hi = min(vs), max(vs)
            if hi - lo > 2047:
                raise Unsupported("vertex span exceeds 11 bits")
            if lo - o >= -1024 and hi - o <= 1023:
                return o
            o2 = max(-1024, min(1023, lo + 1024))  # puts lo at -1024 when reachable
            if not (lo - o2 >= -1024 and hi - o2 <= 1023):
                o2 = max(-1024, min(1023, hi - 1023))
            if not (lo - o2 >= -1024 and hi - o2 <= 1023):
                raise Unsupported("position not reachable with an 11-bit offset")
            return o2

        nox, noy = fit(ox, xs), fit(oy, ys)
        if (nox, noy) != (ox, oy):
            self.gp0(0xE5000000 | (nox & 0x7FF) | ((noy & 0x7FF) << 11))
            self.offset = (nox, noy)

    def _set_window(self, win) -> None:
        if tuple(win) == self.window:
            return
        and_x, and_y, or_x, or_y = win
        mx, my = (~and_x & 0xFF), (~and_y & 0xFF)
        if (mx | my | or_x | or_y) & 7:
            raise Unsupported("texture window is not a multiple of 8")
        mx, my, ox, oy = mx >> 3, my >> 3, or_x >> 3, or_y >> 3
        if (ox & mx) != ox or (oy & my) != oy:
            raise Unsupported("texture window offset outside mask")
        self.gp0(0xE2000000 | mx | (my << 5) | (ox << 10) | (oy << 15))
        self.window = tuple(win)

    def _set_draw_mode(self, want: int, keep_mask: int) -> None:
        """Writes E1 unless the bits outside `keep_mask` (those the primitive itself supplies) already match."""
        want &= 0x17FF  # bit 11 (texture disable) needs GP1(09h); the generator never sets it
        if (self.draw_mode & ~keep_mask & 0x17FF) != (want & ~keep_mask):
            self.gp0(0xE1000000 | want)
            self.draw_mode = want

    def _clut_for(self, flags0: int, draw_mode: int, palette: int) -> None:
        """Mirror of the decoder's reload rule: a textured primitive in a palettised mode reloads the CLUT when the
        palette register differs or when it needs 256 entries and only 16 were loaded."""
        mode = (draw_mode >> 7) & 3
        if not (flags0 & DF_TEXTURE) or mode >= 2:
            return
        needs8 = mode == 1
        if self.pending_clut is not None:
            reg, is8 = self.pending_clut
            if reg != (palette & 0x7FFF) or is8 != needs8:
                raise Unsupported("UpdateCLUT does not belong to the next textured primitive")
            would_reload = self.clut_reg != reg or (needs8 and not self.clut8)
            if not would_reload:
                self.gp0(0x01000000)  # clear cache: forces the reload the record stream asks for
            self.clut_reg, self.clut8 = reg, needs8
            self.pending_clut = None
        else:
            would_reload = self.clut_reg != (palette & 0x7FFF) or (needs8 and not self.clut8)
            if would_reload:
                raise Unsupported("stream samples a stale CLUT (the decoder would reload here)")

    pending_clut = None

    # ---------------------------------------------------------------- records -> words
    def add_stream(self, stream: bytes) -> None:
        off, n = 0, len(stream)
        while off < n:
            size, typ = struct.unpack_from("<IB", stream, off)
            self.add_record(typ, stream[off:off + size])
            off += size
        if self.pending_clut is not None:
            raise Unsupported("trailing UpdateCLUT")

    def add_record(self, typ: int, rec: bytes) -> None:
        if typ == CMD_AREA:
            l, t, r, b = struct.unpack_from("<4I", rec, 8)
            if (l, t, r, b) != self.area:
                self.gp0(0xE3000000 | l | (t << 10), 0xE4000000 | r | (b << 10))
                self.area = (l, t, r, b)
        elif typ == CMD_CLUT:
            reg, is8 = struct.unpack_from("<HB", rec, 8)
            if self.pending_clut is not None:
                raise Unsupported("two UpdateCLUT records in a row")
            self.pending_clut = (reg, bool(is8))
        elif typ == CMD_CLEAR_CACHE:
            self.gp0(0x01000000)
            self.clut_reg, self.clut8 = None, False
        elif typ == CMD_FILL:
            x, y, w, h, color, il, lsb = struct.unpack_from("<4HIBB", rec, 8)
            if il or lsb:
                raise Unsupported("interlaced fill")
            if (x & 0xF) or (w & 0xF) or x > 0x3F0 or y > 0x1FF or w > 0x400 or h > 0x1FF or not w or not h:
                raise Unsupported("fill not expressible")
            self.gp0(0x02000000 | (color & 0xFFFFFF), x | (y << 16), (w & 0x3FF if w < 0x400 else 0x3F1) | (h << 16))
        elif typ == CMD_COPY:
            sx, sy, dx, dy, w, h, sm, cm = struct.unpack_from("<6HBB", rec, 8)
            self._set_mask(sm, cm)
            self.gp0(0x80000000, sx | (sy << 16), dx | (dy << 16), (w & 0x3FF) | ((h & 0x1FF) << 16))
        elif typ == CMD_UPLOAD:
            x, y, w, h, sm, cm = struct.unpack_from("<4HBB", rec, 8)
            self._set_mask(sm, cm)
            px = np.frombuffer(rec, dtype="<u2", count=w * h, offset=18)
            if px.size & 1:
                px = np.concatenate([px, np.zeros(1, dtype="<u2")])
            self.gp0(0xA0000000, x | (y << 16), (w & 0x3FF) | ((h & 0x1FF) << 16), *px.view("<u4").tolist())
        elif typ == CMD_READ_VRAM:
            x, y, w, h = struct.unpack_from("<4H", rec, 8)
            self.gp0(0xC0000000, x | (y << 16), (w & 0x3FF) | ((h & 0x1FF) << 16))
            self.finish_read()
        elif typ in (CMD_POLY, CMD_RECT, CMD_LINE):
            self._add_draw(typ, rec)
        elif typ == CMD_UPDATE_DISPLAY and self.display:
            self._add_display(rec)
        else:
            raise Unsupported(f"record type {typ}")

    def _add_display(self, rec: bytes) -> None:
        """A progressive 320x240 frame whose display rectangle starts at the display address: GP1 writes a game would
        make (ranges and enable once, mode and start address when they change) and the end of the frame.  With the
        front-end in `borders` crop mode this yields the same UpdateDisplay geometry as the record."""
        w, h, ol, ot, vl, vt, vw, vh = struct.unpack_from("<8H", rec, 8)
        x, _busy, flags = struct.unpack_from("<HBB", rec, 28)
        if (w, h, ol, ot, vw, vh) != (320, 240, 0, 0, 320, 240) or x != vl or (flags & 0x2F):
            raise Unsupported("display geometry")
        if self.disp_mode is None:
            self.gp1(0x06000000 | 0x260 | (0xC60 << 12))
            self.gp1(0x07000000 | 16 | (256 << 10))
            self.gp1(0x03000000)
        mode = 0x08000000 | 1 | (0x10 if flags & 0x10 else 0)
        if mode != self.disp_mode:
            self.gp1(mode)
            self.disp_mode = mode
        addr = vl | (vt << 10)
        if addr != self.disp_addr:
            self.gp1(0x05000000 | addr)
            self.disp_addr = addr
        self.vsync()

    def _add_draw(self, typ: int, rec: bytes) -> None:
        f0, f1, nv, draw_mode, palette = struct.unpack_from("<BBHHH", rec, 8)
        win = struct.unpack_from("<4B", rec, 16)
        if f0 & (DF_INTERLACED | DF_LSB):
            raise Unsupported("interlaced rendering flags")
        tex, raw, semi, shaded = bool(f0 & DF_TEXTURE), bool(f0 & DF_RAW), bool(f0 & DF_SEMI), bool(f0 & DF_SHADED)
        dither = bool(f1 & DF1_DITHER)
        self._set_mask(1 if f0 & DF_SET_MASK else 0, 1 if f0 & DF_CHECK_MASK else 0)
        if typ == CMD_LINE and tex:
            raise Unsupported("textured line")
        # the dither FLAG is derived by the decoder from E1 bit 9 and the primitive kind
        want_mode = draw_mode & 0x17FF
        if typ == CMD_POLY:
            relevant = shaded or (tex and not raw)
            if relevant:
                want_mode = (want_mode & ~0x200) | (0x200 if dither else 0)
            elif dither:
                raise Unsupported("dither flag on a primitive that cannot dither")
        elif typ == CMD_LINE:
            want_mode = (want_mode & ~0x200) | (0x200 if dither else 0)
        elif dither:
            raise Unsupported("dithered rectangle")
        if tex:
            self._set_window(win)

        if typ == CMD_POLY:
            if nv not in (3, 4) or bool(f1 & DF1_QUAD) != (nv == 4):
                raise Unsupported("vertex count / quad flag mismatch")
            vs = [struct.unpack_from("<iiBBBxBB", rec, 20 + 16 * i) for i in range(nv)]
            self._set_draw_mode(want_mode, 0x09FF if tex else 0)
            self._set_offset_for([v[0] for v in vs], [v[1] for v in vs])
            self._clut_for(f0, draw_mode, palette)
            ox, oy = self.offset
            cmd = 0x20 | (0x10 if shaded else 0) | (0x08 if nv == 4 else 0) | (0x04 if tex else 0) | (0x02 if semi else 0) | (
                0x01 if raw else 0)
            words = []
            for i, (x, y, r, g, b, u, v) in enumerate(vs):
                color = r | (g << 8) | (b << 16)
                if i == 0:
                    words.append((cmd << 24) | color)
                elif shaded:
                    words.append(color)
                elif (r, g, b) != vs[0][2:5]:
                    raise Unsupported("per-vertex colour on a flat polygon")
                words.append(_pos(x - ox, y - oy))
                if tex:
                    hi = palette if i == 0 else (draw_mode & 0x09FF) if i == 1 else 0
                    words.append(u | (v << 8) | (hi << 16))
            if tex:
                self.draw_mode = (self.draw_mode & ~0x09FF) | (draw_mode & 0x09FF)
            self.gp0(*words)
        elif typ == CMD_RECT:
            w, h, texcoord = struct.unpack_from("<HHH", rec, 20)
            x, y, color = struct.unpack_from("<iiI", rec, 28)
            self._set_draw_mode(want_mode, 0)
            self._clut_for(f0, draw_mode, palette)
            ox, oy = self.offset
            if w > 0x3FF or h > 0x1FF:
                raise Unsupported("rectangle size")
            size = {(1, 1): 1, (8, 8): 2, (16, 16): 3}.get((w, h), 0) if (x + y) & 1 else 0  # use both encodings
            cmd = 0x60 | (size << 3) | (0x04 if tex else 0) | (0x02 if semi else 0) | (0x01 if raw else 0)
            words = [(cmd << 24) | (color & 0xFFFFFF), _pos(x - ox, y - oy)]
            if tex:
                words.append(texcoord | (palette << 16))
            if size == 0:
                words.append(w | (h << 16))
            self.gp0(*words)
        else:
            if nv < 2 or nv & 1:
                raise Unsupported("line vertex count")
            vs = [struct.unpack_from("<iiBBB", rec, 20 + 12 * i) for i in range(nv)]
            self._set_draw_mode(want_mode, 0)
            self._set_offset_for([v[0] for v in vs], [v[1] for v in vs])
            ox, oy = self.offset
            # an unshaded line takes its colour from the first vertex of each pair; the second one's is ignored
            if shaded:
                chained = all(vs[i] == vs[i + 1] for i in range(1, nv - 1, 2))
            else:
                chained = (all(vs[i][:2] == vs[i + 1][:2] for i in range(1, nv - 1, 2)) and
                           all(vs[i][2:5] == vs[0][2:5] for i in range(0, nv, 2)))
            npts = nv // 2 + 1
            # the reference's decoder pops three words after an unshaded poly-line command, so an unshaded
            # poly-line needs three points to survive the round trip
            if chained and nv > 2 and (shaded or npts >= 3):
                pts = [vs[0]] + [vs[i] for i in range(1, nv, 2)]
                cmd = 0x48 | (0x10 if shaded else 0) | (0x02 if semi else 0)
                words = []
                for i, (x, y, r, g, b) in enumerate(pts):
                    color = r | (g << 8) | (b << 16)
                    if i == 0:
                        words.append((cmd << 24) | color)
                    elif shaded:
                        words.append(color)
                    words.append(_pos(x - ox, y - oy))
                words.append(0x55555555)
                self.gp0(*words)
            else:
                cmd = 0x40 | (0x10 if shaded else 0) | (0x02 if semi else 0)
                for i in range(0, nv, 2):
                    (x0, y0, r0, g0, b0), (x1, y1, r1, g1, b1) = vs[i], vs[i + 1]
                    words = [(cmd << 24) | r0 | (g0 << 8) | (b0 << 16), _pos(x0 - ox, y0 - oy)]
                    if shaded:
                        words.append(r1 | (g1 << 8) | (b1 << 16))
                    words.append(_pos(x1 - ox, y1 - oy))
                    self.gp0(*words)


# -------------------------------------------------------------------- .psxgpu container
DUMP_MAGIC = b"PSXGPUDUMPv1\x00\x00"
PKT_GP0, PKT_GP1, PKT_VSYNC, PKT_DISCARD_GP0, PKT_READBACK, PKT_TRACE_BEGIN, PKT_VERSION = 0, 1, 2, 3, 4, 5, 6
PKT_GAME_ID, PKT_VIDEO_FORMAT, PKT_COMMENT = 0x10, 0x11, 0x12


def _packet(ptype: int, words) -> bytes:
    words = [w & 0xFFFFFFFF for w in words]
    return struct.pack("<I", len(words) | (ptype << 24)) + struct.pack(f"<{len(words)}I", *words)


def write_dump(events, *, max_gp0_words: int = 0, extra_metadata: bool = True) -> bytes:
    """Serialises assembler events into a `.psxgpu` trace.  GP0 runs are cut into packets of at most `max_gp0_words`
    words (0 = one packet per run) so that commands straddle packet boundaries."""
    out = bytearray(DUMP_MAGIC)
    if extra_metadata:
        out += _packet(PKT_VERSION, [1])
        out += _packet(PKT_GAME_ID, list(struct.unpack("<3I", b"SLUS-00000\x00\x00")))
        out += _packet(PKT_VIDEO_FORMAT, [0])
        out += _packet(PKT_COMMENT, list(struct.unpack("<2I", b"roundtrp")))
        out += _packet(PKT_TRACE_BEGIN, [0] * 16)
    for e in events:
        if e[0] == "gp0":
            ws = e[1]
            step = max_gp0_words or len(ws) or 1
            for i in range(0, len(ws), step):
                out += _packet(PKT_GP0, ws[i:i + step])
        elif e[0] == "gp1":
            out += _packet(PKT_GP1, [e[1]])
        elif e[0] == "vsync":
            out += _packet(PKT_VSYNC, [0, 0])
        elif e[0] == "finish_read":
            out += _packet(PKT_DISCARD_GP0, [1])
        else:
            raise ValueError(e)
    return bytes(out)


def feed(frontend, events, rng=None) -> None:
    """Plays assembler events into a duckstation_b200.frontend.Gp0Frontend.  With `rng`, GP0 runs are written in
    random-sized pieces (1..40 words), the way DMA blocks and CPU stores interleave on hardware."""
    for e in events:
        if e[0] == "gp0":
            ws = e[1]
            if rng is None:
                frontend.gp0(ws)
            else:
                i = 0
                while i < len(ws):
                    k = int(rng.integers(1, 41))
                    frontend.gp0(ws[i:i + k])
                    i += k
        elif e[0] == "gp1":
            frontend.gp1(e[1])
        elif e[0] == "vsync":
            frontend.vsync()
        elif e[0] == "finish_read":
            frontend.finish_read()
