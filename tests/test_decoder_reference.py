"""The GP0/GP1 front-end (csrc/gp0_frontend.cpp) against the REFERENCE'S OWN decoder, executed: src/core/gpu.cpp compiled
in place with link stubs into oracle/_ref/libpsxdec.so (oracle/ref_build/Makefile.decoder, dec_harness.cpp).  Both are
fed the same port writes through the reference's trace-replay entry (GPU::ProcessGPUDumpPacket) and must emit the same
backend records: SURVEY §8 rows (a)16 decoder rules, (f)1 front-end, (f)2 CRTC display parameters, (f)4 save states.

What is compared: every record's meaningful bytes (header fields, the vertices / pixels its own counts announce).  Not
compared: allocation slack behind them (the reference sizes poly-line and upload records generously and a half-culled
quad leaves its fourth vertex there), the pixel aspect ratio, GPU-busy percentage and presentation parameters of
UpdateDisplay (front-end settings), and a rectangle's `num_vertices` (never written by the reference).
Host-only; skipped where neither the library nor the reference checkout is present."""
import ctypes as C
import difflib
import os
import struct

import numpy as np
import pytest

from duckstation_b200 import streams
from duckstation_b200.frontend import CROP_BORDERS, CROP_NONE, CROP_OVERSCAN, Gp0Frontend

from tests import gp0_assembler as asm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libpsxdec.so")
REF_CROP = {CROP_NONE: 0, CROP_OVERSCAN: 1, CROP_BORDERS: 3}  # DisplayCropMode (src/core/types.h:157-165)

pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libpsxdec.so not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    lib = C.CDLL(LIB)
    lib.psxdec_packet.argtypes = [C.c_uint, C.c_void_p, C.c_size_t]
    lib.psxdec_records_size.restype = C.c_size_t
    lib.psxdec_take_records.argtypes = [C.c_void_p]
    lib.psxdec_save_state.restype = C.c_size_t
    lib.psxdec_save_state.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
    lib.psxdec_load_state.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
    lib.psxdec_read_gpustat.restype = C.c_uint
    return lib


def ref_feed(lib, events):
    for e in events:
        if e[0] == "gp0":
            a = np.asarray(e[1], dtype=np.uint32)
            lib.psxdec_packet(0, a.ctypes.data, a.size)
        elif e[0] == "gp1":
            a = np.asarray([e[1]], dtype=np.uint32)
            lib.psxdec_packet(1, a.ctypes.data, 1)
        elif e[0] == "vsync":
            a = np.zeros(2, dtype=np.uint32)
            lib.psxdec_packet(2, a.ctypes.data, 2)


def ref_take(lib) -> bytes:
    n = lib.psxdec_records_size()
    buf = C.create_string_buffer(max(n, 1))
    lib.psxdec_take_records(buf)
    return buf.raw[:n]


FIXED = {7: 8, 8: 8, 10: 8, 11: 8, 21: 8, 15: 16, 16: 24, 18: 24, 19: 24, 20: 12, 24: 40}


def meaningful(stream: bytes, drop=()):
    out, off = [], 0
    while off < len(stream):
        size, typ = struct.unpack_from("<IB", stream, off)
        rec = stream[off:off + size]
        off += size
        if typ in drop:
            continue
        if typ in FIXED:
            n = FIXED[typ]
        elif typ in (22, 23, 25, 26):
            nv = struct.unpack_from("<H", rec, 10)[0]
            n = {22: 20 + 16 * nv, 23: 40 + 28 * nv, 25: 20 + 12 * nv, 26: 20 + 24 * nv}[typ]
        elif typ == 17:
            w, h = struct.unpack_from("<HH", rec, 12)
            n = 18 + 2 * w * h
        elif typ == 9:  # geometry, X, flags
            out.append(bytes([typ]) + rec[8:24] + rec[28:30] + rec[31:32])
            continue
        else:
            n = size
        assert n <= size, (typ, n, size)
        body = bytearray(rec[8:n])
        if typ == 24:
            body[2:4] = b"\0\0"
        out.append(bytes([typ]) + bytes(body))
    return out


def assert_same_records(ref_stream: bytes, our_stream: bytes, drop=()):
    a, b = meaningful(ref_stream, drop), meaningful(our_stream, drop)
    if a == b:
        return
    sm = difflib.SequenceMatcher(None, a, b, autojunk=False)
    msgs = []
    for tag, i1, i2, j1, j2 in sm.get_opcodes():
        if tag != "equal":
            msgs.append(f"{tag} reference[{i1}:{i2}] ours[{j1}:{j2}]: "
                        f"{[x[:48].hex() for x in a[i1:i2][:2]]} vs {[x[:48].hex() for x in b[j1:j2][:2]]}")
    raise AssertionError(f"{len(msgs)} divergences between the reference decoder and the front-end:\n" + "\n".join(msgs[:8]))


def strip_unsupported(stream: bytes) -> bytes:
    """Records the assembler cannot express as port writes (display records: the CRTC tests below write GP1 themselves;
    interlaced-rendering flags: they come from the beam, not from a command)."""
    out = bytearray()
    for typ, off, size in streams.iter_records(stream):
        rec = bytearray(stream[off:off + size])
        if typ in (7, 8, 9, 10, 11, 12, 15):
            continue
        if typ in (22, 23, 24, 25, 26):
            rec[8] &= ~3 & 0xFF
        if typ == 16:
            rec[20] = 0
            rec[21] = 0
        out += rec
    return bytes(out)


@pytest.mark.parametrize("case", [(1, 3, 1500, 0), (2, 3, 3000, 0), (3, 3, 3000, 0), (4, 9, 2000, 0), (6, 9, 2000, 0), (0, 33, 2500, 3),
                                  (5, 3, 4, 400)], ids=lambda c: "cfg{}-seed{}-n{}-f{}".format(*c))
def test_generated_streams_decode_to_the_same_records(ref, case):
    """Generator records -> GP0 words (tests/gp0_assembler.py) -> both decoders."""
    a = asm.Gp0Assembler()
    a.add_stream(strip_unsupported(streams.generate(*case)))
    ref.psxdec_init(0, REF_CROP[CROP_OVERSCAN], 1)
    ref_feed(ref, a.events)
    fe = Gp0Frontend()
    asm.feed(fe, a.events, rng=np.random.default_rng(case[1]))
    ours = fe.records()
    assert fe.stats()["unknown_commands"] == 0
    assert_same_records(ref_take(ref), ours)


def random_port_writes(seed: int, n: int):
    """Raw words: every GP0 command class with random parameters (coordinates mostly near the screen, sometimes anywhere),
    unknown command bytes, GP1 display writes and VSyncs in between.  VRAM->CPU commands are left out: draining
    GPUREAD is the CPU's side of that transfer."""
    rng = np.random.default_rng(seed)
    w = lambda: int(rng.integers(0, 1 << 32))

    def pos():
        if rng.random() < 0.8:
            return (int(rng.integers(-100, 700)) & 0x7FF) | ((int(rng.integers(-100, 600)) & 0x7FF) << 16)
        return w()

    ev = []
    for _ in range(n):
        k = rng.random()
        if k < 0.08:
            c = int(rng.choice([0x00, 0x01, 0x02, 0x03, 0x04, 0x05, 0x06, 0x07, 0x08, 0x09, 0x10]))
            ev.append(("gp1", (c << 24) | (w() & 0xFFFFFF)))
            continue
        if k < 0.10:
            ev.append(("vsync",))
            continue
        c = int(rng.integers(0, 256))
        if k < 0.5:
            c = int(rng.choice(list(range(0x20, 0x80)) + [0xE1, 0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0x02, 0x01, 0x80, 0xA0]))
        words = [(c << 24) | (w() & 0xFFFFFF)]
        if 0x20 <= c < 0x40:
            for v in range(4 if c & 8 else 3):
                if v > 0 and c & 0x10:
                    words.append(w() & 0xFFFFFF)
                words.append(pos())
                if c & 4:
                    words.append(w())
        elif 0x40 <= c < 0x60:
            if c & 8:
                for v in range(int(rng.integers(2, 6))):
                    if v > 0 and c & 0x10:
                        words.append(w() & 0xFFFFFF)
                    words.append(pos())
                words.append(0x55555555 if rng.random() < 0.7 else 0x50005000)
            else:
                words.append(pos())
                if c & 0x10:
                    words.append(w() & 0xFFFFFF)
                words.append(pos())
        elif 0x60 <= c < 0x80:
            words.append(pos())
            if c & 4:
                words.append(w())
            if ((c >> 3) & 3) == 0:
                words.append(int(rng.integers(0, 300)) | (int(rng.integers(0, 300)) << 16))
        elif c == 0x02:
            words += [pos(), int(rng.integers(0, 1024)) | (int(rng.integers(0, 512)) << 16)]
        elif 0x80 <= c < 0xA0:
            words += [w(), w(), int(rng.integers(0, 200)) | (int(rng.integers(0, 200)) << 16)]
        elif 0xA0 <= c < 0xC0:
            ww, hh = int(rng.integers(1, 24)), int(rng.integers(1, 8))
            words += [w(), ww | (hh << 16)] + [w() for _ in range((ww * hh + 1) // 2)]
        elif 0xC0 <= c < 0xE0:
            continue
        ev.append(("gp0", words))
    return ev


@pytest.mark.parametrize("seed,pal,crop,progressive", [(1, 0, CROP_OVERSCAN, 0), (2, 1, CROP_NONE, 1), (3, 0, CROP_BORDERS, 0),
                                                       (4, 1, CROP_BORDERS, 1), (5, 0, CROP_NONE, 0), (6, 1, CROP_OVERSCAN, 0),
                                                       (7, 0, CROP_OVERSCAN, 1), (8, 1, CROP_NONE, 0)])
def test_random_port_writes_decode_to_the_same_records(ref, seed, pal, crop, progressive):
    """Word-level fuzz, both video standards, the three crop modes, forced-progressive on and off: primitives, poly-lines,
    transfers, E1-E6 state, unknown commands, and the UpdateDisplay geometry after GP1 writes (UpdateCRTCDisplayParameters,
    gpu.cpp:1243-1431)."""
    ev = random_port_writes(seed, 2500)
    ref.psxdec_init(pal, REF_CROP[crop], progressive)
    ref_feed(ref, ev)
    fe = Gp0Frontend(pal=bool(pal), crop_mode=crop, interlaced=not progressive)
    asm.feed(fe, ev)
    assert_same_records(ref_take(ref), fe.records())


PRE_STATE = [("gp1", 0x08000027), ("gp1", 0x05000000 | 16 | (8 << 10)), ("gp1", 0x06000000 | 0x260 | (0xC60 << 12)),
             ("gp1", 0x07000000 | 0x10 | (0x100 << 10)),
             ("gp0", [0xE1000508, 0xE2012345, 0xE3000000 | 10 | (20 << 10), 0xE4000000 | 600 | (400 << 10), 0xE5000000 | 5 | (7 << 11),
                      0xE6000003]),
             ("gp0", [0x2C808080, 0x00100010, 0x7FC00000 | 0x0102, 0x00100090, 0x00150000 | 0x8070, 0x00900010, 0x0003, 0x00900090, 0x0004]),
             ("gp0", [0xA0000000, 0x00200010, 0x00040008, 0x11112222, 0x33334444])]  # an 8 x 4 upload: 16 words wanted, 2 sent
POST_STATE = [("gp0", [0x55556666] * 14 + [0x68FFFFFF, 0x00300030, 0x64FFFFFF, 0x00400040, 0x1234, 0x00080008]), ("vsync",)]


def test_save_state_sections_cross_load(ref):
    """GPU::DoState, both directions, current version.  The reference, stopped in the middle of a CPU->VRAM upload with
    every register away from its reset value, writes its "GPU" section; the front-end loads it and the rest of the port
    writes must decode to the same records as on the reference after loading its own section.  Then the front-end's own
    section (psxgpu_frontend_save_state) goes into the reference and both continue again."""
    version = 86
    ref.psxdec_init(0, REF_CROP[CROP_OVERSCAN], 0)
    ref_feed(ref, PRE_STATE)
    ref_take(ref)
    buf = C.create_string_buffer(4 << 20)
    n = ref.psxdec_save_state(buf, len(buf), version)
    assert n > (1 << 20)
    section = buf.raw[:n]
    stat_saved = ref.psxdec_read_gpustat()

    assert ref.psxdec_load_state(section, len(section), version) == 1
    ref_take(ref)
    ref_feed(ref, POST_STATE)
    want = ref_take(ref)

    fe = Gp0Frontend(interlaced=True)
    assert fe.load_state(section, version) == len(section)
    asm.feed(fe, POST_STATE)
    assert_same_records(want, fe.records())

    fe2 = Gp0Frontend(interlaced=True)
    fe2.load_state(section, version)
    mine = fe2.save_state()
    ref.psxdec_init(0, REF_CROP[CROP_OVERSCAN], 0)
    assert ref.psxdec_load_state(mine, len(mine), 86) == 1
    # static GPUSTAT bits survive the round trip (bits 24-28, 31 are recomputed from FIFO / DMA state on load)
    assert (ref.psxdec_read_gpustat() ^ stat_saved) & 0x60FFFFFF == 0, hex(ref.psxdec_read_gpustat() ^ stat_saved)
    ref_take(ref)
    ref_feed(ref, POST_STATE)
    asm.feed(fe2, POST_STATE)
    assert_same_records(ref_take(ref), fe2.records())


@pytest.mark.parametrize("version", [86, 85, 84, 82, 73, 72, 64, 63, 61, 46, 45, 42])
def test_older_save_state_versions_read_like_the_reference_reads_them(ref, version):
    """The reference only WRITES the current version; what it READS of the older ones (the fields dropped before 62,
    fractional_dot_ticks from 46, the CLUT cache from 64, the texture-cache tail from 73, the pair removed in 83, the
    poly-line buffer from 85, the tail's extra index from 86) is pinned by giving the same synthetic section
    (tests/test_save_state.py:build_section) to GPU::DoState and to the front-end: the reference must accept it, and
    both must decode what follows — the rest of an interrupted upload, a textured rectangle that depends on the loaded
    draw mode / window / offset, a VSync that depends on the loaded display registers — to the same records."""
    from tests.test_save_state import FIELDS, build_section

    rng = np.random.default_rng(version)
    vram = rng.integers(0, 65536, 512 * 1024, dtype=np.uint16).astype("<u2").tobytes()
    clut = rng.integers(0, 65536, 256, dtype=np.uint16)
    f = dict(FIELDS, field=0, line_lsb=0, blitter=2, xfer=(16, 32, 8, 4), blit=[0x11112222, 0x33334444], blit_remaining=14,
             render_command=0xA0000000)
    section = build_section(version, f, vram, clut, tc_writes=(2, 0, 1))
    ref.psxdec_init(1, REF_CROP[CROP_OVERSCAN], 0)
    assert ref.psxdec_load_state(section, len(section), version) == 1
    assert (ref.psxdec_read_gpustat() ^ f["gpustat"]) & 0x00FFDFFF == 0
    ref_take(ref)
    ref_feed(ref, POST_STATE)
    fe = Gp0Frontend(pal=True, interlaced=True)
    assert fe.load_state(section, version) == len(section)
    asm.feed(fe, POST_STATE)
    assert_same_records(ref_take(ref), fe.records())
