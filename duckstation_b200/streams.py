"""Seeded synthetic backend-command streams (wire format of include/psxgpu_wire.h) from libpsxgen.so.

Config numbers are BASELINE.json's, 1-based; config 0 is the fuzz generator.  See SURVEY.md §8(d) for the
workload definitions and duckstation_b200/csrc/streamgen.cpp for the implementation.
"""
from __future__ import annotations

import ctypes as C
import os
import struct

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpsxgen.so")
_lib = None

FUZZ_OVERLAP_TEXTURES = 1
FUZZ_TRANSFERS = 2

# wire record types (include/psxgpu_wire.h)
CMD_UPDATE_DISPLAY = 9
CMD_READ_VRAM = 15


def _load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(f"{_LIB_PATH} is missing: run `python -m duckstation_b200.build`")
        lib = C.CDLL(_LIB_PATH)
        lib.psxgen_config.restype = C.POINTER(C.c_uint8)
        lib.psxgen_config.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_size_t)]
        lib.psxgen_free.argtypes = [C.POINTER(C.c_uint8)]
        lib.psxgen_free.restype = None
        _lib = lib
    return _lib


def generate(config: int, seed: int, n: int, flags: int = 0) -> bytes:
    """One command stream.  ``n``: primitives (configs 0-4) or frames (config 5; ``flags`` = primitives per frame)."""
    lib = _load()
    size = C.c_size_t()
    p = lib.psxgen_config(config, seed, n, flags, C.byref(size))
    if not p:
        raise ValueError(f"unknown config {config}")
    try:
        return C.string_at(p, size.value)
    finally:
        lib.psxgen_free(p)


def iter_records(stream: bytes):
    """Yields (type, offset, size) for every record of a wire stream."""
    off, n = 0, len(stream)
    while off < n:
        size, typ = struct.unpack_from("<IB", stream, off)
        if size < 8 or off + size > n:
            raise ValueError("malformed stream")
        yield typ, off, size
        off += size


def split_frames(stream: bytes) -> list[bytes]:
    """Splits a stream after every UpdateDisplay record (config 5: one chunk per frame)."""
    frames, start = [], 0
    for typ, off, size in iter_records(stream):
        if typ == CMD_UPDATE_DISPLAY:
            frames.append(stream[start:off + size])
            start = off + size
    if start < len(stream):
        frames.append(stream[start:])
    return frames


# The five BASELINE.json configs at their full sizes: (config, seed, n, flags)
FULL_SIZE = {
    1: (1, 1, 20000, 0),
    2: (2, 1, 20000, 0),
    3: (3, 1, 10000, 0),
    4: (4, None, 2000, 0),  # seed = instance index
    5: (5, 1, 60, 1500),
}
BATCH_INSTANCES = 4096


def make_update_display(left: int, top: int, width: int, height: int, *, is24: bool = False, X: int | None = None,
                        interlaced: bool = False, field: bool = False, interleaved: bool = False,
                        disabled: bool = False) -> bytes:
    """One GPUBackendUpdateDisplayCommand record (64 bytes; layout include/psxgpu_wire.h psx_update_display)."""
    flags = ((1 if interlaced else 0) | (2 if field else 0) | (4 if interleaved else 0) | (0x10 if is24 else 0) |
             (0x20 if disabled else 0) | 0x40)
    x = left if X is None else X
    rec = struct.pack("<IB3x8HfHBB32x", 64, CMD_UPDATE_DISPLAY, width, height, 0, 0, left, top, width, height, 1.0, x, 0,
                      flags)
    assert len(rec) == 64
    return rec


def to_precise(stream: bytes) -> bytes:
    """Rewrites every DrawPolygon / DrawLine record as its DrawPrecisePolygon / DrawPreciseLine twin (the records the
    reference's decoder emits with PGXP on, video_thread_commands.h:304-317, :352-362): float positions are filled with
    garbage on purpose — the software renderer must only look at native_x / native_y (gpu_sw.cpp:118-139, :172-190)."""
    out = bytearray()
    for typ, off, size in iter_records(stream):
        rec = stream[off:off + size]
        if typ == 22:  # DrawPolygon -> DrawPrecisePolygon
            nv = struct.unpack_from("<H", rec, 10)[0]
            head = bytearray(rec[:20])
            body = bytearray(head[8:20]) + bytearray(8)  # `params` copy of the draw header payload (unused)
            verts = bytearray()
            for i in range(nv):
                x, y, color, tc = struct.unpack_from("<iiIH", rec, 20 + 16 * i)
                verts += struct.pack("<fffiiIH2x", 1e9, -3.5, 0.0, x, y, color, tc)
            new_size = (40 + len(verts) + 15) & ~15
            struct.pack_into("<IB", head, 0, new_size, 23)
            out += head + bytes(head[:20]) + verts + bytes(new_size - 40 - len(verts))
            del body
        elif typ == 25:  # DrawLine -> DrawPreciseLine
            nv = struct.unpack_from("<H", rec, 10)[0]
            head = bytearray(rec[:20])
            verts = bytearray()
            for i in range(nv):
                x, y, color = struct.unpack_from("<iiI", rec, 20 + 12 * i)
                verts += struct.pack("<fffiiI", -7.25, 1e6, 2.0, x, y, color)
            new_size = (20 + len(verts) + 15) & ~15
            struct.pack_into("<IB", head, 0, new_size, 26)
            out += head + verts + bytes(new_size - 20 - len(verts))
        else:
            out += rec
    return bytes(out)
