import os, sys, time, struct
sys.path.insert(0, '/root/repo')
from duckstation_b200 import Context
def _rec(fmt, typ, *vals):
    body = struct.pack(fmt, *vals); size = (8 + len(body) + 15) & ~15
    return struct.pack("<IB3x", size, typ) + body + b"\0" * (size - 8 - len(body))
def _area(l, t, r, b): return _rec("<4I", 19, l, t, r, b)
def _rect(x, y, w, h, color=0x808080, textured=False, page_x=0, u=0, v=0, mode=2):
    flags0 = 0x10 if textured else 0
    return _rec("<BBHHH4BHHH2xiiI", 24, flags0, 0, 1, (page_x & 15) | (mode << 7), 0, 0xFF, 0xFF, 0, 0, w, h, u | (v << 8), x, y, color)
N=2000
s=_area(0,0,1023,511)+b"".join(_rect(520,8,30,30,0xFFFFFF)+_rect(10+(i%50),10,40,40,textured=True,page_x=8) for i in range(N))
with Context(1) as ctx:
    for rep in range(3):
        ctx.clear_vram(0); ctx.sync()
        t0=time.perf_counter(); ctx.submit(s,0); ctx.flush(); ctx.sync(); dt=(time.perf_counter()-t0)*1e3
    st=ctx.stats(); t=ctx.timing()
    print("persistent: %.2f ms total, device %.2f ms, epochs %d -> %.2f us/epoch"%(dt,t['last_flush_device_ms'],st['epochs']//3,1e3*t['last_flush_device_ms']/(st['epochs']//3)))
with Context(1, profile=True) as ctx:
    for rep in range(3):
        ctx.clear_vram(0); ctx.sync()
        t0=time.perf_counter(); ctx.submit(s,0); ctx.flush(); ctx.sync(); dt=(time.perf_counter()-t0)*1e3
    st=ctx.stats(); t=ctx.timing()
    print("launch-per-epoch: %.2f ms total, device %.2f ms, epochs %d -> %.2f us/epoch"%(dt,t['last_flush_device_ms'],st['epochs']//3,1e3*t['last_flush_device_ms']/(st['epochs']//3)))
