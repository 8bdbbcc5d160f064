"""Test-side access to the oracle (oracle/libpsxoracle.so), the compiled reference (oracle/_ref/libpsxref.so,
only where it was built) and the host simulation.  TEST INFRASTRUCTURE: nothing under duckstation_b200/ imports this."""
from __future__ import annotations

import ctypes as C
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_PATH = os.path.join(ROOT, "oracle", "libpsxoracle.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libpsxref.so")
HOSTSIM_PATH = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")


class OracleStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("draw_px", "draw_px_rmw", "draw_px_tex", "fill_px", "copy_px", "copy_px_masked",
                                          "upload_px", "upload_px_masked", "prims", "commands")]


class _Renderer:
    def __init__(self, path: str, prefix: str):
        self.lib = C.CDLL(path)
        self.prefix = prefix
        f = lambda n: getattr(self.lib, f"{prefix}_{n}")
        f("vram").restype = C.POINTER(C.c_uint16)
        f("clut").restype = C.POINTER(C.c_uint16)
        f("exec").restype = C.c_long
        f("exec").argtypes = [C.c_char_p, C.c_size_t]
        f("scanout").restype = C.c_int
        f("scanout").argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        f("set_modulation_crop").argtypes = [C.c_int]
        f("init")()
        self._f = f

    def reset(self, modulation_crop: bool = False) -> None:
        self._f("set_modulation_crop")(int(modulation_crop))
        self._f("reset")()

    def exec(self, stream: bytes) -> int:
        n = self._f("exec")(stream, len(stream))
        if n < 0:
            raise ValueError("malformed stream")
        return n

    def vram(self) -> np.ndarray:
        return np.frombuffer(C.string_at(self._f("vram")(), 1 << 20), dtype=np.uint16).reshape(512, 1024).copy()

    def load_vram(self, vram: np.ndarray, clut: np.ndarray | None = None) -> None:
        C.memmove(self._f("vram")(), np.ascontiguousarray(vram, dtype=np.uint16).ctypes.data, 1 << 20)
        if clut is not None:
            C.memmove(self._f("clut")(), np.ascontiguousarray(clut, dtype=np.uint16).ctypes.data, 512)

    def clut(self) -> np.ndarray:
        return np.frombuffer(C.string_at(self._f("clut")(), 512), dtype=np.uint16).copy()

    def scanout(self, display_record: bytes, fmt: int = 0, show_vram: bool = False) -> np.ndarray | None:
        """Returns rows of 4096 bytes (like Context.scanout), or None if the display is disabled."""
        w, h = C.c_uint32(), C.c_uint32()
        buf = np.zeros(1024 * 512 * 4 + 64, dtype=np.uint8)
        rc = self._f("scanout")(display_record, fmt, int(show_vram), buf.ctypes.data, 4096, C.byref(w), C.byref(h))
        if rc == 1:
            return None
        if rc != 0:
            raise ValueError("scanout failed")
        self.last_size = (w.value, h.value)
        return buf[: 4096 * h.value].reshape(h.value, 4096)

    def render(self, stream: bytes, modulation_crop: bool = False) -> np.ndarray:
        self.reset(modulation_crop)
        self.exec(stream)
        return self.vram()


class Oracle(_Renderer):
    def __init__(self):
        super().__init__(ORACLE_PATH, "psxo")
        self.lib.psxo_get_stats.argtypes = [C.POINTER(OracleStats)]
        self.lib.psxo_vramhash64.restype = C.c_uint64
        self.lib.psxo_vramhash64.argtypes = [C.c_void_p]

    def stats(self) -> dict:
        s = OracleStats()
        self.lib.psxo_get_stats(C.byref(s))
        return {n: int(getattr(s, n)) for n, _ in OracleStats._fields_}

    def hash64(self, vram: np.ndarray | None = None) -> int:
        v = self.vram() if vram is None else np.ascontiguousarray(vram, dtype=np.uint16)
        return int(self.lib.psxo_vramhash64(v.ctypes.data))


def have_reference() -> bool:
    return os.path.exists(REF_PATH)


class Reference(_Renderer):
    """The reference's own rasteriser TU (SW_USE_ISA decides Scalar/SIMD; default SIMD)."""

    def __init__(self):
        super().__init__(REF_PATH, "psxref")


class HostSim:
    def __init__(self):
        self.lib = C.CDLL(HOSTSIM_PATH)
        self.lib.hostsim_vram.restype = C.POINTER(C.c_uint16)
        self.lib.hostsim_exec.restype = C.c_long
        self.lib.hostsim_exec.argtypes = [C.c_char_p, C.c_size_t]

    def render(self, stream: bytes, modulation_crop: bool = False) -> np.ndarray:
        self.lib.hostsim_reset(int(modulation_crop))
        if self.lib.hostsim_exec(stream, len(stream)) < 0:
            raise ValueError("malformed stream")
        return np.frombuffer(C.string_at(self.lib.hostsim_vram(), 1 << 20), dtype=np.uint16).reshape(512, 1024).copy()

    def render_parts(self, parts, handover: bool, modulation_crop: bool = False) -> np.ndarray:
        """Executes the parts as separate flushes; with `handover` the encoder is replaced between flushes by a fresh
        one that only knows the exported state blob (host encoder <-> device encoder hand-over)."""
        self.lib.hostsim_reset(int(modulation_crop))
        for i, part in enumerate(parts):
            if i and handover:
                self.lib.hostsim_handover(int(modulation_crop))
            if self.lib.hostsim_exec(part, len(part)) < 0:
                raise ValueError("malformed stream")
        return np.frombuffer(C.string_at(self.lib.hostsim_vram(), 1 << 20), dtype=np.uint16).reshape(512, 1024).copy()

    def stats(self) -> dict:
        v = (C.c_uint64 * 12)()
        self.lib.hostsim_stats(v, 12)
        names = ("wire_commands", "packed_prims", "epochs", "cuts_raw", "cuts_war", "cuts_clut", "serial_self_sampling",
                 "serial_copy", "serial_upload", "clut_ops", "clut_cache_hits", "rejected")
        return dict(zip(names, [int(x) for x in v]))


def sha256(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def describe_diff(a: np.ndarray, b: np.ndarray) -> str:
    d = np.argwhere(a != b)
    if len(d) == 0:
        return "identical"
    y, x = d[0]
    return f"{len(d)} pixels differ; first at (x={x}, y={y}): {int(a[y, x]):#06x} vs {int(b[y, x]):#06x}; rows {d[:,0].min()}..{d[:,0].max()} cols {d[:,1].min()}..{d[:,1].max()}"
