import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from duckstation_b200 import Context, streams
from tests.helpers import Oracle
o = Oracle()
bad = 0
n = 0
with Context(1) as ctx:
    for seed in range(100, 160):
        for flags in (0, 1, 2, 3):
            s = streams.generate(0, seed, 500, flags)
            want = o.render(s)
            ctx.clear_vram(0)
            ctx.submit(s)
            got = ctx.read_vram()
            n += 1
            if not np.array_equal(got, want):
                bad += 1
                print("MISMATCH seed", seed, "flags", flags, int((got != want).sum()))
# batched / device encoder path with interlace etc.
batch = []
for seed in range(200, 264):
    s = streams.generate(0, seed, 400, seed % 4)
    batch.append(b"".join(s[off:off + size] for typ, off, size in streams.iter_records(s) if typ not in (7, 9, 12, 15)))
want = [o.hash64(o.render(s)) for s in batch]
with Context(len(batch), encode_mode=2) as ctx:
    ctx.submit_batch(batch)
    ctx.flush()
    got = [int(h) for h in ctx.hash()]
bad2 = [i for i in range(len(batch)) if got[i] != want[i]]
print("single-stream cases", n, "mismatches", bad, "| batched mismatches", bad2)
