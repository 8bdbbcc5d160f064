"""Generates tests/golden/golden.json from the REFERENCE ITSELF: the reference's rasteriser translation unit compiled in
place (oracle/_ref/libpsxref.so, built by oracle/ref_build/Makefile from /root/reference) is run on the seeded streams of
duckstation_b200/csrc/streamgen.cpp and the resulting VRAM / scan-out images are recorded as SHA-256 digests plus the
64-bit device hash.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no golden vectors of its own for this path, so these are the pins: the oracle
(oracle/psx_oracle.c) must reproduce every entry on CPU-only machines, and the CUDA path must on the GPU box.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from duckstation_b200 import streams  # noqa: E402
from tests.helpers import Oracle, Reference, sha256  # noqa: E402

# (config, seed, n, flags): small enough for the whole CPU suite to stay within minutes
CASES = [
    (1, 1, 20000, 0), (1, 2, 3000, 0),
    (2, 1, 20000, 0), (2, 5, 4000, 0),
    (3, 1, 10000, 0), (3, 9, 3000, 0),
    (4, 0, 2000, 0), (4, 1, 2000, 0), (4, 4095, 2000, 0),
    (5, 1, 60, 1500), (5, 2, 12, 300),
    (0, 1, 3000, 0), (0, 1, 3000, 1), (0, 1, 3000, 2), (0, 1, 3000, 3),
    (0, 7, 3000, 3), (0, 8, 3000, 3),
]
# Q3b: keep the reference's under-guarded fast paths inside VRAM
SCANOUT_WINDOWS = [
    dict(left=0, top=0, width=320, height=240),
    dict(left=16, top=256, width=640, height=240),
    dict(left=900, top=400, width=256, height=200),                      # wraps in x and y -> slow path
    dict(left=0, top=0, width=320, height=240, interlaced=True, field=True, interleaved=True),   # 480i field 1
    dict(left=0, top=16, width=512, height=240, interlaced=True, field=False, interleaved=True),
    dict(left=0, top=0, width=320, height=240, interlaced=True, field=True, interleaved=False),
    dict(left=0, top=0, width=320, height=240, is24=True),
    dict(left=8, top=256, width=320, height=200, is24=True, X=0),        # skip_x = 8
    dict(left=800, top=300, width=320, height=240, is24=True, X=800),    # 24-bit wrap path
    dict(left=0, top=0, width=256, height=120, is24=True, interlaced=True, field=True, interleaved=True),
    dict(left=0, top=0, width=0, height=0),
]
MODULATION_CROP_CASES = [(2, 3, 2000, 0), (0, 4, 2000, 1)]


def main() -> None:
    ref = Reference()
    orc = Oracle()
    out = {"generator": "tests/golden/make_golden.py", "reference_isa": os.environ.get("SW_USE_ISA", "SIMD (default)"),
           "cases": [], "modulation_crop_cases": [], "scanout": []}
    for case in CASES:
        s = streams.generate(*case)
        v = ref.render(s)
        out["cases"].append({"case": list(case), "stream_sha256": sha256(__import__("numpy").frombuffer(s, dtype="u1")),
                             "vram_sha256": sha256(v), "vramhash64": orc.hash64(v)})
        print(case, out["cases"][-1]["vram_sha256"][:16])
    for case in MODULATION_CROP_CASES:
        s = streams.generate(*case)
        v = ref.render(s, modulation_crop=True)
        out["modulation_crop_cases"].append({"case": list(case), "vram_sha256": sha256(v), "vramhash64": orc.hash64(v)})
    # scan-out: per-frame VRAM + display image of a short config-5 trace, all four 15-bit formats on frame 0
    s = streams.generate(5, 3, 12, 200)
    ref.reset()
    for i, frame in enumerate(streams.split_frames(s)):
        ref.exec(frame)
        disp = [frame[o:o + z] for t, o, z in streams.iter_records(frame) if t == streams.CMD_UPDATE_DISPLAY]
        entry = {"frame": i, "vram_sha256": sha256(ref.vram())}
        if disp:
            img = ref.scanout(disp[-1], 0)
            w, h = ref.last_size
            entry["scanout_sha256"] = sha256(img[:, : w * 4])
            entry["size"] = [w, h]
            if i == 0:
                for fmt in (1, 2, 3):
                    im = ref.scanout(disp[-1], fmt)
                    entry[f"scanout_fmt{fmt}_sha256"] = sha256(im[:, : w * 2])
        out["scanout"].append(entry)
    # hand-made display windows on the final VRAM: wrap paths, interlaced fields, 24-bit slow path
    out["scanout_windows"] = []
    for kw in SCANOUT_WINDOWS:
        rec = streams.make_update_display(**kw)
        for fmt in ((0,) if kw.get("is24") else (0, 1, 2, 3)):
            img = ref.scanout(rec, fmt)
            w, h = ref.last_size
            bpp = 4 if (fmt == 0 or kw.get("is24")) else 2
            out["scanout_windows"].append({"params": kw, "fmt": fmt, "sha256": sha256(img[:, : w * bpp])})
    with open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote golden.json:", len(out["cases"]), "cases,", len(out["scanout"]), "frames")


if __name__ == "__main__":
    main()
