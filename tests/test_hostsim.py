"""Host logic without a GPU: the encoder, the hazard/epoch splitter, the CLUT snapshot cache and the closed-form
setup / coverage maths (the same headers the CUDA kernels compile), executed by tests/hostsim with the kernels'
execution model and compared with the oracle."""
import struct

import numpy as np
import pytest

from duckstation_b200 import streams
from tests.helpers import HostSim, Oracle, describe_diff


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture(scope="module")
def sim():
    return HostSim()


CASES = [(1, 3, 4000, 0), (2, 2, 5000, 0), (3, 2, 4000, 0), (4, 17, 2000, 0), (5, 4, 8, 300)] + \
        [(0, s, 2000, f) for s in (31, 32) for f in (0, 1, 2, 3)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "cfg{}-seed{}-n{}-f{}".format(*c))
def test_hostsim_matches_oracle(oracle, sim, case):
    s = streams.generate(*case)
    got, want = sim.render(s), oracle.render(s)
    assert np.array_equal(got, want), describe_diff(got, want)


def test_modulation_crop(oracle, sim):
    s = streams.generate(2, 3, 2000, 0)
    assert np.array_equal(sim.render(s, True), oracle.render(s, True))
    assert not np.array_equal(oracle.render(s, True), oracle.render(s, False))


def test_empty_stream(oracle, sim):
    assert not sim.render(b"").any()
    assert sim.stats()["epochs"] == 0


def test_malformed_stream_rejected(sim):
    bad = struct.pack("<IB3x", 12, 22) + b"\0" * 4  # size not a multiple of 4 past the header rule / truncated payload
    with pytest.raises(ValueError):
        sim.render(bad[:10])


def _truncated_records():
    """Records whose framed size is smaller than what their own fields ask for (a polygon that claims four vertices and
    carries one, a poly-line of 65535 vertices in 32 bytes, a rectangle cut in half, an upload without its pixels)."""
    poly = struct.pack("<IB3x", 36, 22) + struct.pack("<BBHHH4B", 0, 0, 4, 0, 0, 0xFF, 0xFF, 0, 0) + b"\x11" * 16
    line = struct.pack("<IB3x", 32, 25) + struct.pack("<BBHHH4B", 0, 0, 65535, 0, 0, 0xFF, 0xFF, 0, 0) + b"\x22" * 12
    rect = struct.pack("<IB3x", 24, 24) + b"\x33" * 16
    upload = struct.pack("<IB3x", 20, 17) + struct.pack("<4HBB2x", 0, 0, 64, 64, 0, 0)
    return [poly, line, rect, upload]


def test_truncated_records_are_dropped_not_read_past(oracle, sim):
    good = _area(0, 0, 1023, 511) + _rect(10, 10, 50, 50, 0x3050F0)
    s = good + b"".join(_truncated_records()) + _rect(100, 100, 20, 20, 0x10F010)
    got = sim.render(s)
    assert sim.stats()["rejected"] == 4
    assert np.array_equal(got, oracle.render(good + _rect(100, 100, 20, 20, 0x10F010)))


# ---- epoch splitter behaviour (stats come from the encoder) ----
def _rec(fmt, typ, *vals):
    body = struct.pack(fmt, *vals)
    size = (8 + len(body) + 15) & ~15
    return struct.pack("<IB3x", size, typ) + body + b"\0" * (size - 8 - len(body))


def _area(l, t, r, b):
    return _rec("<4I", 19, l, t, r, b)


def _rect(x, y, w, h, color=0x808080, textured=False, page_x=0, u=0, v=0, mode=2):
    flags0 = 0x10 if textured else 0
    draw_mode = (page_x & 15) | (mode << 7)
    return _rec("<BBHHH4BHHH2xiiI", 24, flags0, 0, 1, draw_mode, 0, 0xFF, 0xFF, 0, 0, w, h, u | (v << 8), x, y, color)


def _fill(x, y, w, h, color):
    return _rec("<4HIBB", 16, x, y, w, h, color, 0, 0)


def _copy(sx, sy, dx, dy, w, h):
    return _rec("<6HBB", 18, sx, sy, dx, dy, w, h, 0, 0)


def test_disjoint_draws_share_one_epoch(sim):
    s = _area(0, 0, 1023, 511) + b"".join(_rect(16 * i, 8 * i, 40, 40) for i in range(20))
    sim.render(s)
    assert sim.stats()["epochs"] == 1


def test_render_to_texture_forces_a_cut(oracle, sim):
    # draw into page 8 (x = 512..), then sample page 8: RAW; then draw over page 8 again while it was sampled: WAR
    s = (_area(0, 0, 1023, 511) + _fill(512, 0, 64, 64, 0x123456) + _rect(520, 8, 30, 30, 0xFFFFFF) +
         _rect(10, 10, 40, 40, textured=True, page_x=8) + _rect(530, 4, 20, 20, 0x0000FF) + _rect(100, 100, 16, 16))
    got = sim.render(s)
    st = sim.stats()
    assert st["cuts_raw"] == 1 and st["cuts_war"] == 1 and st["epochs"] == 3
    assert np.array_equal(got, oracle.render(s))


def test_self_sampling_goes_serial(oracle, sim):
    s = _area(0, 0, 1023, 511) + _fill(0, 0, 256, 256, 0x336699) + _rect(3, 0, 200, 100, textured=True, page_x=0)
    got = sim.render(s)
    assert sim.stats()["serial_self_sampling"] == 1
    assert np.array_equal(got, oracle.render(s))


def test_overlapping_copy_goes_serial(oracle, sim):
    noise = streams.generate(0, 5, 0, 0)  # just the noise upload of the fuzz generator
    s = noise + _copy(100, 100, 104, 110, 200, 150) + _copy(1000, 500, 300, 200, 60, 40) + _copy(0, 0, 600, 300, 64, 64)
    got = sim.render(s)
    assert sim.stats()["serial_copy"] == 1
    assert np.array_equal(got, oracle.render(s))


def test_clut_cache_reuses_snapshots(sim):
    s = streams.generate(4, 3, 2000, 0)
    sim.render(s)
    st = sim.stats()
    assert st["epochs"] == 2  # clear + atlas upload | all primitives
    assert st["clut_cache_hits"] > 10 * st["clut_ops"]


def test_precise_records_use_native_coordinates(oracle, sim):
    """DrawPrecisePolygon / DrawPreciseLine (PGXP record types) render exactly like their plain twins."""
    from tests.helpers import Reference, have_reference

    s = streams.generate(3, 6, 1500, 0)
    p = streams.to_precise(s)
    assert p != s and len(list(streams.iter_records(p))) == len(list(streams.iter_records(s)))
    want = oracle.render(s)
    assert np.array_equal(oracle.render(p), want)
    assert np.array_equal(sim.render(p), want)
    if have_reference():
        assert np.array_equal(Reference().render(p), want)


def test_degenerate_drawing_areas_and_extreme_sizes(oracle, sim):
    big = _rect(-1024, -1024, 1023, 511, 0x1F3F5F) + _rect(1, 1, 1023, 511, 0x204060) + _fill(0, 0, 1024, 512, 0xFFFFFF)
    cases = [
        _area(600, 300, 100, 200) + big,      # left > right: nothing is clipped in, sprite culled
        _area(0, 0, 1023, 1023) + big,        # bottom beyond VRAM: rows wrap at 512
        _area(1023, 511, 1023, 511) + big,    # one pixel
        _area(0, 0, 0, 0) + _copy(0, 0, 0, 0, 1024, 512) + _copy(5, 5, 1020, 508, 1024, 512),
    ]
    for s in cases:
        got, want = sim.render(s), oracle.render(s)
        assert np.array_equal(got, want), describe_diff(got, want)


@pytest.mark.parametrize("case", [(0, 31, 900, 3), (0, 32, 900, 1), (2, 4, 1200, 0), (3, 6, 700, 0), (4, 9, 800, 0)],
                         ids=lambda c: "cfg{}-seed{}".format(*c[:2]))
def test_encoder_state_survives_flushes_and_hand_over(case):
    """Drawing area, current CLUT slots and the snapshot table carry across flushes, and across
    Encoder::export_state -> import_state into a fresh encoder (how an instance moves between the host encoder and the
    device encoder of the batched path)."""
    s = streams.generate(*case)
    s = b"".join(s[off:off + size] for typ, off, size in streams.iter_records(s) if typ not in (7, 9, 12, 15))
    offs = [off for _, off, _ in streams.iter_records(s)]
    cuts = [0, offs[len(offs) // 5], offs[len(offs) // 2], offs[4 * len(offs) // 5], len(s)]
    parts = [s[a:b] for a, b in zip(cuts, cuts[1:])]
    want = Oracle().render(s)
    sim = HostSim()
    for handover in (False, True):
        got = sim.render_parts(parts, handover)
        assert np.array_equal(got, want), (handover, describe_diff(got, want))


def test_setup_divisions_through_double_precision_are_exact(sim):
    """psx_setup.h divides through double precision on the GPU (trunc_div_rcp: ten divisions by the determinant from one
    reciprocal; makestep_xy: the 64-bit edge steps): both must equal C integer division on every input class."""
    import ctypes as C
    sim.lib.hostsim_division_check.restype = C.c_long
    sim.lib.hostsim_division_check.argtypes = [C.c_long]
    assert sim.lib.hostsim_division_check(20_000_000) == 0
