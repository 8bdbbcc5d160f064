"""Small driver for ncu captures: N instances of the batched workload, staged once, replayed `reps` times.
Raster launches come in pairs per replay: wave 0 (FillVRAM + atlas upload), wave 1 (all primitives)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duckstation_b200 import Context, streams  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
config = int(sys.argv[3]) if len(sys.argv) > 3 else 4
prims = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
wire = [streams.generate(config, i, prims, 0) for i in range(n)]
with Context(n, profile=True) as ctx:
    ctx.set_batch_chunk(n)  # one chunk: one setup + one raster launch per wave
    ctx.submit_batch(wire)
    ctx.flush()
    ctx.sync()
    for _ in range(reps):
        ctx.run_staged()
        ctx.sync()
    h = ctx.hash()
    print("raster launch ms:", ctx.raster_launch_ms(), "hash[0]=%#x" % int(h[0]), ctx.timing())
