"""Presenter kernels under ncu: a 640x240 interlaced 24-bit field through scan-out, chroma smoothing and each
deinterlacer (`ncu --metrics gpu__time_duration.sum -k regex:'scanout|chroma|weave|blend|adaptive'`)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duckstation_b200 import Context, streams  # noqa: E402
from duckstation_b200.presenter import Presenter  # noqa: E402

rng = np.random.default_rng(1)
with Context(1) as ctx:
    ctx.load_state(rng.integers(0, 65536, (512, 1024), dtype=np.uint16))
    for mode in (1, 2, 3):
        with Presenter(ctx, mode, True) as p:
            for field in (0, 1, 0, 1):
                rec = streams.make_update_display(0, 16, 640, 240, is24=True, X=0, interlaced=True, field=bool(field),
                                                  interleaved=True)
                img = p.present(rec)
    print("ok", img.shape)
