"""Where a frame of config 5 (60-frame trace, UpdateDisplay per frame) spends its wall time: per-frame submit() wall time
against the library's own timers."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duckstation_b200 import Context, streams  # noqa: E402

s = streams.generate(5, 1, 60, 1500)
frames = list(streams.split_frames(s))
with Context(1) as ctx:
    for rep in range(3):
        ctx.clear_vram(0)
        ctx.sync()
        acc = {}
        t_all = time.perf_counter()
        per = []
        for f in frames:
            t0 = time.perf_counter()
            ctx.submit(f, 0)
            per.append((time.perf_counter() - t0) * 1e3)
            t = ctx.timing()
            for k, v in t.items():
                if isinstance(v, float):
                    acc[k] = acc.get(k, 0.0) + v
        wall = (time.perf_counter() - t_all) * 1e3
    n = len(frames)
    print(f"{n} frames, wall {wall:.2f} ms ({wall / n * 1e3:.0f} us per frame); per-frame submit min/median/max "
          f"{min(per) * 1e3:.0f}/{sorted(per)[n // 2] * 1e3:.0f}/{max(per) * 1e3:.0f} us")
    print({k: round(v / n * 1e3, 1) for k, v in acc.items()}, "(us per frame)")
    print("records per frame:", sum(1 for _ in streams.iter_records(frames[1])), "bytes", len(frames[1]))
