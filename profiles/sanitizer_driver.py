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
with Context(20, encode_mode=2) as ctx:  # device encoder: framing, layout, encode (fast + uniform paths), wave kernels
    batch = [streams.generate(4, i, 150, 0) for i in range(10)] + [streams.generate(0, 40 + i, 200, 3) for i in range(8)] + [
        b"".join(mixed[o:o + z] for t, o, z in streams.iter_records(mixed) if t not in (7, 9, 12, 15)), b""]
    ctx.set_batch_chunk(7)
    ctx.submit_batch(batch)
    ctx.flush()
    ctx.run_staged()
    print("device-encoded:", [hex(int(h)) for h in ctx.hash()][:4])
