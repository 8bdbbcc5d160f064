"""Smallest program that touches every kernel and both raster modes: for compute-sanitizer runs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duckstation_b200 import Context, streams  # noqa: E402

mixed = streams.generate(3, 2, 400, 0) + streams.generate(0, 3, 300, 3)
with Context(1) as ctx:  # latency mode (persistent kernel), serial paths, scan-out, read-back
    ctx.submit(mixed + streams.generate(5, 1, 2, 100))
    print("single:", hex(int(ctx.hash()[0])))
    ctx.scanout(streams.make_update_display(900, 400, 256, 200))
    ctx.scanout(streams.make_update_display(0, 0, 320, 240, is24=True))
with Context(6, profile=True) as ctx:  # throughput kernels
    ctx.submit_batch([streams.generate(4, i, 200, 0) for i in range(5)] + [mixed])
    ctx.flush()
    print("batch:", [hex(int(h)) for h in ctx.hash()])
