"""cpu_farm.py — TEST / BASELINE INFRASTRUCTURE.  Runs the CPU oracle (oracle/libpsxoracle.so, "port") or the
reference's own rasteriser TU compiled in place (oracle/_ref/libpsxref.so, "reference") over many instances, one
instance at a time per worker PROCESS (the reference TU keeps VRAM / CLUT / drawing area in process globals; the
reference's own fan-out is process based too: scripts/run_regression_tests.py:50-60).

Used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
import struct
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "libpsxoracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libpsxref.so")
GEN_PATH = os.path.join(ROOT, "duckstation_b200", "libpsxgen.so")

STAT_NAMES = ("draw_px", "draw_px_rmw", "draw_px_tex", "fill_px", "copy_px", "copy_px_masked", "upload_px",
              "upload_px_masked", "prims", "commands")


def have_reference() -> bool:
    return os.path.exists(REF_PATH)


def _load_gen():
    g = C.CDLL(GEN_PATH)
    g.psxgen_config.restype = C.POINTER(C.c_uint8)
    g.psxgen_config.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_size_t)]
    g.psxgen_free.argtypes = [C.POINTER(C.c_uint8)]
    return g


def _gen(g, config, seed, n, flags) -> bytes:
    sz = C.c_size_t()
    p = g.psxgen_config(config, seed, n, flags, C.byref(sz))
    try:
        return C.string_at(p, sz.value)
    finally:
        g.psxgen_free(p)


class _Pipe:
    """Length-prefixed pickles over a pair of binary streams."""

    def __init__(self, rd, wr):
        self.rd, self.wr = rd, wr

    def send(self, obj) -> None:
        b = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
        self.wr.write(struct.pack("<Q", len(b)))
        self.wr.write(b)
        self.wr.flush()

    def recv(self):
        hdr = self.rd.read(8)
        if len(hdr) < 8:
            raise EOFError("cpu_farm worker pipe closed")
        (n,) = struct.unpack("<Q", hdr)
        return pickle.loads(self.rd.read(n))

    def close(self) -> None:
        try:
            self.wr.close()
        except Exception:
            pass


def _worker(conn):
    gen = _load_gen()
    orc = C.CDLL(ORACLE_PATH)
    orc.psxo_exec.restype = C.c_long
    orc.psxo_exec.argtypes = [C.c_char_p, C.c_size_t]
    orc.psxo_vram.restype = C.c_void_p
    orc.psxo_vramhash64.restype = C.c_uint64
    orc.psxo_vramhash64.argtypes = [C.c_void_p]
    orc.psxo_get_stats.argtypes = [C.c_void_p]
    ref = None
    if os.path.exists(REF_PATH):
        ref = C.CDLL(REF_PATH)
        ref.psxref_exec.restype = C.c_long
        ref.psxref_exec.argtypes = [C.c_char_p, C.c_size_t]
        ref.psxref_vram.restype = C.c_void_p
        ref.psxref_init()
    streams: list[bytes] = []
    while True:
        msg = conn.recv()
        op = msg[0]
        if op == "quit":
            break
        if op == "prepare":  # ("prepare", config, seeds, n, flags)
            _, config, seeds, n, flags = msg
            streams = [_gen(gen, config, s, n, flags) for s in seeds]
            conn.send(sum(len(s) for s in streams))
        elif op == "run":  # ("run", kind, want_stats)
            _, kind, want_stats = msg
            hashes, stats = [], [0] * len(STAT_NAMES)
            t0 = time.perf_counter()
            if kind == "reference" and ref is not None:
                for s in streams:
                    ref.psxref_reset()
                    ref.psxref_exec(s, len(s))
                    if want_stats:
                        hashes.append(int(orc.psxo_vramhash64(ref.psxref_vram())))
            else:
                buf = (C.c_uint64 * len(STAT_NAMES))()
                for s in streams:
                    orc.psxo_reset()
                    orc.psxo_exec(s, len(s))
                    if want_stats:
                        hashes.append(int(orc.psxo_vramhash64(orc.psxo_vram())))
                        orc.psxo_get_stats(buf)
                        for i in range(len(STAT_NAMES)):
                            stats[i] += int(buf[i])
            dt = time.perf_counter() - t0
            conn.send((dt, hashes, stats))
    conn.close()


class CpuFarm:
    """`nproc` worker processes, each owning a strided share of the instances."""

    def __init__(self, nproc: int | None = None, isa: str | None = None):
        self.nproc = max(1, nproc or (os.cpu_count() or 1))
        # plain subprocesses (not multiprocessing): independent of the parent's __main__, CUDA state and torch
        env = dict(os.environ)
        if isa:
            env["SW_USE_ISA"] = isa
        self.conns, self.procs = [], []
        for _ in range(self.nproc):
            p = subprocess.Popen([sys.executable, "-u", os.path.abspath(__file__), "--worker"], stdin=subprocess.PIPE,
                                 stdout=subprocess.PIPE, env=env)
            self.conns.append(_Pipe(p.stdout, p.stdin))
            self.procs.append(p)
        self.seeds: list[list[int]] = [[] for _ in range(self.nproc)]

    def prepare(self, config: int, seeds: list[int], n: int, flags: int = 0) -> int:
        self.seeds = [list(seeds[w::self.nproc]) for w in range(self.nproc)]
        for c, s in zip(self.conns, self.seeds):
            c.send(("prepare", config, s, n, flags))
        return sum(c.recv() for c in self.conns)

    def run(self, kind: str = "reference", want_stats: bool = False):
        """Executes every prepared stream once.  Returns (wall_seconds, {seed: hash64}, stats dict)."""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send(("run", kind, want_stats))
        res = [c.recv() for c in self.conns]
        wall = time.perf_counter() - t0
        hashes, stats = {}, dict.fromkeys(STAT_NAMES, 0)
        for seeds, (_, hs, st) in zip(self.seeds, res):
            hashes.update(zip(seeds, hs))
            for k, v in zip(STAT_NAMES, st):
                stats[k] += v
        return wall, hashes, stats

    def close(self):
        for c in self.conns:
            try:
                c.send(("quit",))
            except Exception:
                pass
        for c in self.conns:
            c.close()
        for p in self.procs:
            try:
                p.wait(timeout=5)
            except Exception:
                p.kill()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


if __name__ == "__main__" and "--worker" in sys.argv:
    _out = os.fdopen(os.dup(1), "wb")
    os.dup2(2, 1)  # anything the libraries print goes to stderr, the pipe stays clean
    _worker(_Pipe(sys.stdin.buffer, _out))
