"""Config 3 (hazard stress, single VRAM: 2470 epochs in 10 k commands) through the C ABI: wall / device time and parity,
for whatever mode the environment selects (PSXGPU_CLUSTER_MAX=0 -> one-CTA instance mode, PSXGPU_CLUSTER_SIZE=8|16,
PSXGPU_INSTANCE_MODE=0 -> cooperative persistent kernel).  usage: config3_cluster_probe.py [config] [n] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duckstation_b200 import Context, streams  # noqa: E402
from tests.helpers import Oracle  # noqa: E402

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
s = streams.generate(cfg, 1, n, 0 if cfg != 5 else 1500)
orc = Oracle()
orc.render(s)
want = orc.hash64()
with Context(1) as ctx:
    best = 1e9
    dev = 1e9
    for rep in range(reps):
        ctx.clear_vram(0)
        ctx.sync()
        t0 = time.perf_counter()
        ctx.submit(s, 0)
        ctx.flush()
        ctx.sync()
        best = min(best, (time.perf_counter() - t0) * 1e3)
        dev = min(dev, ctx.timing()["last_flush_device_ms"])
    ok = int(ctx.hash(0, 1)[0]) == want
    st = ctx.stats()
env = {k: v for k, v in os.environ.items() if k.startswith("PSXGPU_")}
print(f"config {cfg} n {n} env {env}: wall {best:.2f} ms, device {dev:.2f} ms, epochs {st['epochs'] // reps}, parity {ok}")
