"""Save states: the GPU section GPU::DoState reads / writes (src/core/gpu.cpp:574-726).

The layout is written out here a second time, independently of csrc/gp0_frontend.cpp, straight from the reference's
DoState sequence and StateWrapper encodings — so a field that the C++ reader and writer both get wrong the same way
still fails.  No real save state exists in this environment: parity is pinned by this reading only."""
import struct

import numpy as np
import pytest

from duckstation_b200.frontend import Gp0Frontend, find_gpu_section, load_save_state_file, state_file_info

VRAM_BYTES = 1024 * 512 * 2


def marker(text: str) -> bytes:
    return struct.pack("<I", len(text)) + text.encode()


def build_section(version: int, f: dict, vram: bytes, clut: np.ndarray, tc_writes=()) -> bytes:
    """GPU::DoState in write order, with the per-version differences its reader handles."""
    b = bytearray()
    b += struct.pack("<IHHI", f["gpustat"], f["draw_mode"], f["palette"], f["window_reg"])
    if version < 62:
        b += bytes(16)
    b += struct.pack("<4B", *f["window"])
    b += struct.pack("<??", bool(f["draw_mode"] & 0x1000), bool(f["draw_mode"] & 0x2000))
    b += struct.pack("<4I", *f["area"])
    b += struct.pack("<iii", f["offset"][0], f["offset"][1], f["offset"][0])
    b += struct.pack("<??", f["console_pal"], f["tex_disable_mask"])
    b += struct.pack("<III", f["display_start"], f["h_range"], f["v_range"])
    b += struct.pack("<H", 8)              # dot_clock_divider
    b += struct.pack("<8H", *range(1, 9))  # display_width .. display_vram_height (stale on purpose: must be recomputed)
    b += struct.pack("<10H", *range(11, 21))
    b += struct.pack("<iiI", 5, 6, 7)      # fractional_ticks, current_tick_in_scanline, current_scanline
    if version >= 46:
        b += struct.pack("<i", 9)          # fractional_dot_ticks
    b += struct.pack("<??", True, False)   # in_hblank, in_vblank
    b += struct.pack("<BBB", f["field"], f["field"], f["line_lsb"])
    b += struct.pack("<B", f["blitter"])
    b += struct.pack("<iII", 11, 12, 13)   # pending_command_ticks, command_total_words, GPUREAD_latch
    if version >= 64:
        b += struct.pack("<I?", f["clut_reg"], f["clut_is_8bit"]) + clut.astype("<u2").tobytes()
    b += struct.pack("<6H", *f["xfer"], 0, 0)
    b += struct.pack("<I", len(f["fifo"])) + b"".join(struct.pack("<Q", w | (0x80001000 << 32)) for w in f["fifo"])
    b += struct.pack("<I", len(f["blit"])) + b"".join(struct.pack("<I", w) for w in f["blit"])
    b += struct.pack("<II", f["blit_remaining"], f["render_command"])
    if version >= 85:
        b += struct.pack("<I", len(f["polyline"])) + b"".join(struct.pack("<Q", w | (0x1234 << 32)) for w in f["polyline"])
    if version < 83:
        b += bytes(8)
    b += marker("GPU-VRAM") + vram
    if version >= 73:
        b += marker("GPUTextureCache") + struct.pack("<I", len(tc_writes))
        if version >= 86:
            b += struct.pack("<I", 0xFFFFFFFF)
        for n_records in tc_writes:
            b += bytes(44 if version >= 86 else 40) + struct.pack("<I", n_records) + bytes(544 * n_records)
    return bytes(b)


FIELDS = dict(gpustat=0x14802000 | (1 << 11) | (1 << 20) | (1 << 19) | (1 << 22) | (2 << 17) | 0x405, draw_mode=0x0405,
              palette=0x7ABC, window_reg=0x0C3A5, window=(0xD7, 0xEF, 0x28, 0x10), area=(16, 32, 300, 200), offset=(-5, 77),
              console_pal=True, tex_disable_mask=False, display_start=0x4010, h_range=0xC60260, v_range=0x4B410, field=1,
              line_lsb=1, blitter=0, clut_reg=0x7ABC, clut_is_8bit=True, xfer=(0, 0, 0, 0), fifo=[], blit=[], blit_remaining=0,
              render_command=0, polyline=[])


def v86_fields(section: bytes) -> dict:
    """Picks the interesting fields out of a version-86 section at their fixed offsets."""
    gpustat, draw_mode, palette, window_reg = struct.unpack_from("<IHHI", section, 0)
    out = dict(gpustat=gpustat, draw_mode=draw_mode, palette=palette, window_reg=window_reg,
               window=struct.unpack_from("<4B", section, 12), area=struct.unpack_from("<4I", section, 18),
               offset=struct.unpack_from("<ii", section, 34), offset_x_again=struct.unpack_from("<i", section, 42)[0],
               console_pal=section[46] != 0, tex_disable_mask=section[47] != 0)
    out["display_start"], out["h_range"], out["v_range"] = struct.unpack_from("<III", section, 48)
    out["crtc"] = struct.unpack_from("<8H", section, 62)
    at = 62 + 16 + 20 + 16 + 2  # crtc block, timing block, ticks (4 values incl. dot ticks), hblank / vblank
    out["field"], _, out["line_lsb"], out["blitter"] = struct.unpack_from("<4B", section, at)
    at += 4 + 12
    out["clut_reg"], out["clut_is_8bit"] = struct.unpack_from("<I?", section, at)
    out["clut"] = np.frombuffer(section, dtype="<u2", count=256, offset=at + 5)
    return out


@pytest.mark.parametrize("version", [86, 85, 84, 82, 72, 63, 61, 45])
def test_reader_follows_the_reference_layout_of_every_version(version):
    rng = np.random.default_rng(version)
    vram = rng.integers(0, 65536, 512 * 1024, dtype=np.uint16).astype("<u2").tobytes()
    clut = rng.integers(0, 65536, 256, dtype=np.uint16)
    f = dict(FIELDS, blitter=3 if version >= 85 else 0, polyline=[0x00100010, 0x00200040, 0x00FF00FF], fifo=[0x55555555],
             render_command=0x58112233)
    section = build_section(version, f, vram, clut, tc_writes=(2, 0, 1))
    fe = Gp0Frontend()
    assert fe.load_state(section + b"trailing bytes", version) == len(section)
    again = fe.save_state()
    got = v86_fields(again)
    for key in ("draw_mode", "palette", "window_reg", "window", "area", "offset", "console_pal", "tex_disable_mask",
                "display_start", "h_range", "v_range", "field", "line_lsb"):
        assert got[key] == f[key], key
    assert got["offset_x_again"] == f["offset"][0]
    assert got["gpustat"] & 0x00FF1FFF == f["gpustat"] & 0x00FF1FFF  # static bits; 13 (field) / dynamic bits aside
    assert got["crtc"] != tuple(range(1, 9))  # display geometry was recomputed from the registers, not taken over
    if version >= 64:
        assert got["clut_reg"] == f["clut_reg"] and got["clut_is_8bit"] == f["clut_is_8bit"]
    else:
        assert got["clut_reg"] == 0xFFFFFFFF  # CLUT cache invalid: reloaded by the next paletted draw (gpu.cpp:640-646)
    assert got["blitter"] == f["blitter"]
    for cut in (10, 60, 200, len(section) - VRAM_BYTES // 2, len(section) - 3):
        with pytest.raises(ValueError):
            Gp0Frontend().load_state(section[:cut], version)
    with pytest.raises(ValueError):
        Gp0Frontend().load_state(section.replace(b"GPU-VRAM", b"GPU-VRAN"), version)


def test_state_round_trip_in_the_middle_of_a_transfer_and_a_polyline():
    rng = np.random.default_rng(5)
    prefix = [0xE1000408, 0xE2000000 | 0x0C3A5, 0xE3000000 | 8 | (16 << 10), 0xE4000000 | 500 | (400 << 10),
              0xE5000000 | (30 & 0x7FF) | ((2040 & 0x7FF) << 11), 0xE6000001]
    upload = [0xA0000000, (40 << 16) | 24, (6 << 16) | 10] + [int(x) for x in rng.integers(0, 1 << 32, 30)]
    polyline = [0x58FF8040, 0x00100010, 0x0020FF40, 0x00400050, 0x00808080, 0x00600020, 0x55555555]
    rect = [0x64808080, (50 << 16) | 60, (0x7ABC << 16) | 0x0102, (16 << 16) | 16]
    words = prefix + upload + polyline + rect + [0x02123456, (8 << 16) | 16, (32 << 16) | 64]
    for cut in (3, len(prefix) + 2, len(prefix) + 11, len(prefix) + len(upload) + 3, len(prefix) + len(upload) + len(polyline) + 2):
        whole = Gp0Frontend()
        whole.gp1(0x08000023)
        whole.gp1(0x05000000 | 64 | (128 << 10))
        whole.gp0(words[:cut])
        head = whole.records()
        saved = whole.save_state()
        whole.gp0(words[cut:])
        whole.vsync()
        resumed = Gp0Frontend()
        assert resumed.load_state(saved) == len(saved)
        assert resumed.save_state() == saved
        resumed.gp0(words[cut:])
        resumed.vsync()
        tail_whole, tail_resumed = whole.records(), resumed.records()
        # the drawing area is re-sent after a load (drawing_area_changed = true, gpu.cpp:704): drop those records
        strip = lambda s: b"".join(s[o:o + n] for t, o, n in _records(s) if t != 19)
        assert strip(tail_resumed) == strip(tail_whole), cut
        assert head is not None


def _records(stream: bytes):
    off = 0
    while off < len(stream):
        size, typ = struct.unpack_from("<IB", stream, off)
        yield typ, off, size
        off += size


def test_gpu_section_is_found_between_its_markers():
    rng = np.random.default_rng(9)
    vram = rng.integers(0, 65536, 512 * 1024, dtype=np.uint16).astype("<u2").tobytes()
    section = build_section(86, FIELDS, vram, np.arange(256, dtype=np.uint16))
    decoy = marker("GPU") + b"not a gpu section at all" * 4
    data = marker("System") + bytes(100) + marker("CPU") + bytes(300) + decoy + marker("GPU") + section + marker("CDROM") + bytes(64)
    off = find_gpu_section(data)
    assert data[off:off + len(section)] == section
    with pytest.raises(ValueError):
        find_gpu_section(data[: off + 5000])
    with pytest.raises(ValueError):  # a section that parses but is followed by something else
        find_gpu_section(marker("GPU") + section + marker("Pad") + bytes(8))


@pytest.mark.parametrize("compression", [0, 2], ids=["uncompressed", "zstd"])
def test_save_state_file_container(compression):
    rng = np.random.default_rng(3)
    vram = rng.integers(0, 65536, 512 * 1024, dtype=np.uint16).astype("<u2").tobytes()
    clut = rng.integers(0, 65536, 256, dtype=np.uint16)
    section = build_section(84, FIELDS, vram, clut)
    data = marker("System") + bytes(40) + marker("GPU") + section + marker("CDROM") + bytes(200)
    blob = data
    if compression == 2:
        import pyarrow as pa

        blob = pa.compress(data, codec="zstd").to_pybytes()
    header_bytes = 8 + 128 + 32 + 12 * 4
    media = b"/games/Example (USA).chd"
    offset_to_data = header_bytes + len(media) + 16
    header = struct.pack("<II128s32s", 0x43435544, 84, b"Example Game", b"SLUS-00000")
    header += struct.pack("<12I", len(media), header_bytes, 0, 0, 0, 0, 0, 0, compression, len(blob), len(data), offset_to_data)
    file_bytes = header + media + bytes(16) + blob
    info = state_file_info(file_bytes)
    assert info["version"] == 84 and info["title"].startswith("Example Game") and info["serial"].startswith("SLUS-00000")
    assert info["data_compression"] == compression and info["offset_to_data"] == offset_to_data
    fe = Gp0Frontend()
    load_save_state_file(file_bytes, fe)
    got = v86_fields(fe.save_state())
    assert got["area"] == FIELDS["area"] and got["palette"] == FIELDS["palette"]
    with pytest.raises(ValueError):
        state_file_info(b"DUCK" + file_bytes[4:])
    with pytest.raises(ValueError):
        state_file_info(file_bytes[: offset_to_data + 10])


@pytest.mark.gpu
def test_state_seeds_an_instance_and_comes_back_out():
    """VRAM + CLUT cache of a section reach the device like the reference's LoadState command; a 4-bit sprite whose
    palette register equals the saved current_clut_reg must use the *saved* cache (no reload from VRAM), exactly what
    the oracle does after load_vram(vram, clut)."""
    from duckstation_b200 import Context, streams
    from tests.helpers import Oracle

    rng = np.random.default_rng(21)
    vram = rng.integers(0, 65536, (512, 1024), dtype=np.uint16)
    clut = rng.integers(0, 65536, 256, dtype=np.uint16)
    clut_reg = (640 // 16) | (300 << 6)
    f = dict(FIELDS, gpustat=0x14802000, draw_mode=8 | (0 << 7), palette=clut_reg, clut_reg=clut_reg, clut_is_8bit=False,
             area=(0, 0, 511, 511), offset=(0, 0), window_reg=0, window=(0xFF, 0xFF, 0, 0))
    section = build_section(86, f, vram.astype("<u2").tobytes(), clut)
    sprite = [0x65000000 | 0x808080, (40 << 16) | 30, (clut_reg << 16) | 0x0504, (24 << 16) | 48]
    oracle = Oracle()
    with Context(2) as ctx:
        fe = Gp0Frontend()
        assert fe.load_state(section, 86, ctx, instance=1) == len(section)
        assert np.array_equal(ctx.read_vram(1), vram) and not ctx.read_vram(0).any()
        fe.gp0(sprite)
        recs = fe.records()
        assert not any(t == 20 for t, _, _ in _records(recs)), "the saved CLUT cache is current: no UpdateCLUT"
        ctx.submit(recs, instance=1)
        oracle.reset()
        oracle.load_vram(vram, clut)
        oracle.exec(recs)
        got = ctx.read_vram(1)
        assert np.array_equal(got, oracle.vram())
        assert not np.array_equal(got, vram)
        out = fe.save_state(ctx, instance=1)
        back = v86_fields(out)
        assert np.array_equal(back["clut"], clut)
        at = out.index(b"GPU-VRAM") + 8
        assert np.array_equal(np.frombuffer(out, dtype="<u2", count=512 * 1024, offset=at).reshape(512, 1024), got)


def test_damaged_sections_are_rejected_or_survive():
    """Random byte damage and truncation: the reader either refuses the section or yields a front-end that keeps
    working — it never reads outside the buffer (sizes of the FIFO / blit / polyline / texture-cache parts are checked
    against what is left before anything is taken over)."""
    rng = np.random.default_rng(0)
    vram = bytes(VRAM_BYTES)
    for version in (86, 84, 72, 61):
        f = dict(FIELDS, fifo=[1, 2, 3], blit=[5] * 10, polyline=[7] * 6, blitter=2, xfer=(8, 8, 4, 5), blit_remaining=3)
        base = build_section(version, f, vram, np.zeros(256, np.uint16), tc_writes=(1, 2))
        for _ in range(400):
            b = bytearray(base)
            for _ in range(int(rng.integers(1, 6))):
                pos = int(rng.integers(0, 1400)) if rng.random() < 0.7 else len(b) - int(rng.integers(1, 2000))
                b[pos] = int(rng.integers(0, 256))
            if rng.random() < 0.3:
                b = b[: int(rng.integers(0, len(b)))]
            fe = Gp0Frontend()
            try:
                fe.load_state(bytes(b), version)
            except ValueError:
                continue
            fe.gp0([0x02000000, 0, 0x00100010, 0xE1000000])
            fe.vsync()
            fe.records()
