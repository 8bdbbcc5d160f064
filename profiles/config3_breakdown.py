"""Config 3 (hazard stress, single VRAM) launched one kernel per epoch (profile=True) so that ncu can attribute the
time: raster epochs vs serial kinds (aliasing copies, mask-checked uploads, self-sampling primitives)."""
import sys
import time

sys.path.insert(0, ".")
from duckstation_b200 import Context, streams

s = streams.generate(3, 1, 10000, 0)
with Context(1, profile=True) as ctx:
    for rep in range(2):
        ctx.clear_vram()
        t0 = time.perf_counter()
        ctx.submit(s)
        ctx.flush()
        ctx.sync()
        print(f"rep {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms", ctx.stats())
