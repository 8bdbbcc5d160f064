"""Device timeline of the e2e path (PSXGPU_TRACE=1): per chunk copy / encode / wave times, plus raw H2D bandwidth."""
import ctypes as C
import os
import sys
import time

os.environ["PSXGPU_TRACE"] = "1"
sys.path.insert(0, ".")
import numpy as np
import torch

from duckstation_b200 import Context, streams

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wire = [streams.generate(4, i, 2000, 0) for i in range(n)]
offs, total = [], 0
for s in wire:
    offs.append(total)
    total += (len(s) + 15) & ~15
pinned = torch.empty(total + 64, dtype=torch.uint8).pin_memory()
base = pinned.data_ptr()
base += (-base) % 16
view = np.ctypeslib.as_array((C.c_uint8 * total).from_address(base))
ptrs = (C.c_void_p * n)()
sizes = (C.c_size_t * n)()
for i, s in enumerate(wire):
    view[offs[i]:offs[i] + len(s)] = np.frombuffer(s, dtype=np.uint8)
    ptrs[i] = base + offs[i]
    sizes[i] = len(s)
# raw H2D bandwidth of this box
dev = torch.empty(total, dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev.copy_(pinned[:total], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"raw pinned H2D: {total / dt / 1e9:.1f} GB/s ({dt * 1e3:.1f} ms for {total / 1e6:.0f} MB)", file=sys.stderr)
del dev
with Context(n) as ctx:
    if len(sys.argv) > 2:
        ctx.set_batch_chunk(int(sys.argv[2]))
    for rep in range(3):
        t0 = time.perf_counter()
        ctx.submit_batch_raw(ptrs, sizes, n)
        ctx.flush()
        h = ctx.hash()
        t1 = time.perf_counter()
        print(f"rep {rep}: {1e3 * (t1 - t0):.2f} ms wall", file=sys.stderr)
        ctx.sync()
