"""Single-VRAM configs (BASELINE configs[0..2], [4]): wall time through the C ABI vs the CPU reference on this host."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from duckstation_b200 import Context, streams  # noqa: E402
from tests.helpers import Oracle, Reference, have_reference  # noqa: E402

ref = Reference() if have_reference() else Oracle()
orc = Oracle()
for name, case in (("config1", (1, 1, 20000, 0)), ("config2", (2, 1, 20000, 0)), ("config3", (3, 1, 10000, 0)), ("config5", (5, 1, 60, 1500))):
    s = streams.generate(*case)
    t0 = time.perf_counter()
    ref.reset()
    ref.exec(s)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    orc.render(s)
    px = orc.stats()
    with Context(1) as ctx:
        best = 1e9
        for rep in range(4):
            ctx.clear_vram(0)
            ctx.sync()
            t0 = time.perf_counter()
            ctx.submit(s, 0)
            ctx.flush()
            ctx.sync()
            best = min(best, (time.perf_counter() - t0) * 1e3)
        t = ctx.timing()
        st = ctx.stats()
        ok = int(ctx.hash(0, 1)[0]) == orc.hash64()
    print(f"{name}: cpu_ref {cpu_ms:.1f} ms | gpu e2e {best:.1f} ms (device {t['last_flush_device_ms']:.1f} ms, encode {t['last_encode_ms']:.1f} ms, "
          f"epochs {st['epochs'] // 4}, launches {st['kernel_launches'] // 4}) | draw px {px['draw_px'] / 1e6:.1f}M | parity {ok}")
