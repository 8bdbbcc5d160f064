import sys
sys.path.insert(0, ".")
from duckstation_b200 import Context, streams
s = streams.generate(3, 1, 10000, 0)
with Context(1) as ctx:
    ctx.submit(s)
    ctx.flush()
    ctx.sync()
    print(ctx.timing()["last_flush_device_ms"])
