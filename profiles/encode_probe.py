"""Device-encoder probe: one batch of N config-4 instances through psxgpu_submit_batch(encode_mode=2).
Run under `ncu --metrics gpu__time_duration.sum` for per-kernel durations, or plain for wall-clock."""
import sys
import time

sys.path.insert(0, ".")
from duckstation_b200 import Context, streams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
batch = [streams.generate(4, i, 2000, 0) for i in range(n)]
with Context(n, encode_mode=2) as ctx:
    ctx.set_batch_chunk(n)
    for rep in range(3):
        t0 = time.perf_counter()
        ctx.submit_batch(batch)
        ctx.flush()
        ctx.sync()
        t1 = time.perf_counter()
        print(f"rep {rep}: {1e3 * (t1 - t0):.2f} ms wall", ctx.timing())
