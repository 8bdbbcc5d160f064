"""The CPU oracle (oracle/psx_oracle.c) against the committed golden vectors, which were produced by the reference's own
rasteriser TU (tests/golden/make_golden.py), and — where oracle/_ref was built — against the compiled reference live."""
import json
import os

import numpy as np
import pytest

from duckstation_b200 import streams
from tests.helpers import Oracle, Reference, describe_diff, have_reference, sha256

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")))


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.mark.parametrize("entry", GOLDEN["cases"], ids=lambda e: "cfg{}-seed{}-n{}-f{}".format(*e["case"]))
def test_oracle_matches_reference_golden(oracle, entry):
    s = streams.generate(*entry["case"])
    assert sha256(np.frombuffer(s, dtype=np.uint8)) == entry["stream_sha256"], "generator drifted: regenerate goldens"
    v = oracle.render(s)
    assert sha256(v) == entry["vram_sha256"]
    assert oracle.hash64(v) == entry["vramhash64"]


@pytest.mark.parametrize("entry", GOLDEN["modulation_crop_cases"], ids=lambda e: "crop-cfg{}-seed{}".format(*e["case"][:2]))
def test_oracle_modulation_crop_golden(oracle, entry):
    v = oracle.render(streams.generate(*entry["case"]), modulation_crop=True)
    assert sha256(v) == entry["vram_sha256"]


def test_oracle_scanout_golden(oracle):
    s = streams.generate(5, 3, 12, 200)
    oracle.reset()
    for entry, frame in zip(GOLDEN["scanout"], streams.split_frames(s)):
        oracle.exec(frame)
        assert sha256(oracle.vram()) == entry["vram_sha256"], f"frame {entry['frame']}"
        disp = [frame[o:o + z] for t, o, z in streams.iter_records(frame) if t == streams.CMD_UPDATE_DISPLAY]
        if "scanout_sha256" in entry:
            img = oracle.scanout(disp[-1], 0)
            w, h = oracle.last_size
            assert [w, h] == entry["size"]
            assert sha256(img[:, : w * 4]) == entry["scanout_sha256"], f"frame {entry['frame']}"
            for fmt in (1, 2, 3):
                key = f"scanout_fmt{fmt}_sha256"
                if key in entry:
                    assert sha256(oracle.scanout(disp[-1], fmt)[:, : w * 2]) == entry[key]
    for win in GOLDEN["scanout_windows"]:
        rec = streams.make_update_display(**win["params"])
        img = oracle.scanout(rec, win["fmt"])
        w, h = oracle.last_size
        bpp = 4 if (win["fmt"] == 0 or win["params"].get("is24")) else 2
        assert sha256(img[:, : w * bpp]) == win["sha256"], win


def test_scanout_disabled_returns_none(oracle):
    oracle.reset()
    assert oracle.scanout(streams.make_update_display(0, 0, 320, 240, disabled=True)) is None


def test_5551_to_8888_all_values(oracle):
    """(c5*527+23)>>6 expansion, alpha = bit 15 ? 255 : 0 — all 65536 inputs through the oracle's scan-out
    (gpu_helpers.h:18-31), checked against the closed form."""
    vram = np.zeros((512, 1024), dtype=np.uint16)
    vram[:64, :] = np.arange(65536, dtype=np.uint16).reshape(64, 1024)
    oracle.reset()
    oracle.load_vram(vram)
    img = oracle.scanout(streams.make_update_display(0, 0, 1024, 64), 0)[:, :4096].reshape(64, 1024, 4)
    c = np.arange(65536, dtype=np.uint32).reshape(64, 1024)
    e = lambda x: ((x * 527 + 23) >> 6).astype(np.uint8)
    assert np.array_equal(img[..., 0], e(c & 31))
    assert np.array_equal(img[..., 1], e((c >> 5) & 31))
    assert np.array_equal(img[..., 2], e((c >> 10) & 31))
    assert np.array_equal(img[..., 3], np.where(c >> 15, 255, 0).astype(np.uint8))


needs_ref = pytest.mark.skipif(not have_reference(), reason="oracle/_ref was not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("case", [(0, s, 2500, f) for s in (11, 12, 13) for f in (0, 1, 2, 3)] + [(2, 7, 3000, 0), (3, 4, 3000, 0)])
def test_oracle_matches_reference_live(oracle, case):
    """Fresh seeds that are NOT in the golden file, compared pixel for pixel with the compiled reference."""
    s = streams.generate(*case)
    ref = Reference()
    got, want = oracle.render(s), ref.render(s)
    assert np.array_equal(got, want), describe_diff(got, want)


@needs_ref
def test_reference_scalar_differs_only_on_quirks():
    """the oracle models the SIMD build.  The generator's non-overlapping streams must not depend on it."""
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r);"
            "from tests.helpers import Reference, sha256; from duckstation_b200 import streams;"
            "r = Reference(); print(sha256(r.render(streams.generate(0, 21, 2000, 0))), sha256(r.render(streams.generate(0, 21, 3000, 3))))"
            % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = {}
    for isa in ("SIMD", "Scalar"):
        env = dict(os.environ, SW_USE_ISA=isa)
        out[isa] = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.split()
    assert out["SIMD"][0] == out["Scalar"][0]  # textures disjoint from the draw area: ISA independent
    assert out["SIMD"][1] != out["Scalar"][1]  # mask-checked uploads (Q2) and self-sampling (Q1): build dependent
