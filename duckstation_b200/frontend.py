"""GP0/GP1 word-level front-end (include/psxgpu.h `psxgpu_frontend_*`): raw PSX GPU port writes and `.psxgpu` traces in,
backend command records out.  Host-only; mirrors the decoder half of the reference's src/core/gpu.cpp."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

CROP_NONE, CROP_OVERSCAN, CROP_BORDERS = 0, 1, 2
FE_PAL, FE_INTERLACED, FE_FIELD_PER_VSYNC = 1, 2, 4


class Gp0Frontend:
    def __init__(self, *, pal: bool = False, interlaced: bool = False, crop_mode: int = CROP_OVERSCAN,
                 field_per_vsync: bool = False):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        flags = (FE_PAL if pal else 0) | (FE_INTERLACED if interlaced else 0) | (FE_FIELD_PER_VSYNC if field_per_vsync else 0)
        rc = self._lib.psxgpu_frontend_create(flags, crop_mode, C.byref(self._h))
        if rc != 0:
            raise ValueError(f"psxgpu_frontend_create failed ({rc})")

    def close(self) -> None:
        if self._h:
            self._lib.psxgpu_frontend_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def gp0(self, words) -> None:
        """Writes words to the GP0 port (a single int or any sequence of them)."""
        a = np.ascontiguousarray(np.atleast_1d(np.asarray(words, dtype=np.uint64)).astype(np.uint32))
        self._lib.psxgpu_frontend_gp0(self._h, a.ctypes.data, a.size)

    def gp1(self, value: int) -> None:
        self._lib.psxgpu_frontend_gp1(self._h, value & 0xFFFFFFFF)

    def vsync(self) -> None:
        self._lib.psxgpu_frontend_vsync(self._h)

    def finish_read(self) -> None:
        self._lib.psxgpu_frontend_finish_read(self._h)

    def feed_dump(self, data: bytes) -> int:
        n = self._lib.psxgpu_frontend_feed_dump(self._h, data, len(data))
        if n < 0:
            raise ValueError(f"malformed .psxgpu trace ({n})")
        return n

    def load_state(self, section: bytes, version: int = 86, ctx=None, instance: int = 0) -> int:
        """The GPU section of a save state (GPU::DoState, gpu.cpp:574-726): registers / FIFO / transfer state into this
        front-end, VRAM + CLUT cache into `instance` of `ctx` (if given).  Returns the bytes consumed."""
        n = C.c_size_t()
        rc = self._lib.psxgpu_frontend_load_state(self._h, ctx._ctx if ctx is not None else None, instance, section,
                                                  len(section), version, C.byref(n))
        if rc != 0:
            raise ValueError(f"psxgpu_frontend_load_state failed ({rc})")
        return n.value

    def save_state(self, ctx=None, instance: int = 0) -> bytes:
        """A version-86 GPU section of the current state (VRAM / CLUT cache from `ctx`, zero without one)."""
        n = C.c_size_t()
        h = ctx._ctx if ctx is not None else None
        rc = self._lib.psxgpu_frontend_save_state(self._h, h, instance, None, 0, C.byref(n))
        if rc != 0:
            raise ValueError(f"psxgpu_frontend_save_state failed ({rc})")
        buf = C.create_string_buffer(n.value)
        rc = self._lib.psxgpu_frontend_save_state(self._h, h, instance, buf, n.value, C.byref(n))
        if rc != 0:
            raise ValueError(f"psxgpu_frontend_save_state failed ({rc})")
        return buf.raw[: n.value]

    def records(self, clear: bool = True) -> bytes:
        """Backend records decoded so far (the psxgpu_submit_wire format)."""
        n = C.c_size_t()
        p = self._lib.psxgpu_frontend_records(self._h, C.byref(n))
        out = C.string_at(p, n.value) if n.value else b""
        if clear:
            self._lib.psxgpu_frontend_clear(self._h)
        return out

    def stats(self) -> dict:
        s = _lib.FrontendStats()
        self._lib.psxgpu_frontend_get_stats(self._h, C.byref(s))
        return {n: int(getattr(s, n)) for n, _ in s._fields_ if n != "reserved"}


def replay_dump(ctx, data: bytes, *, instance: int = 0, pal: bool = False, interlaced: bool = False,
                crop_mode: int = CROP_OVERSCAN, on_frame=None, field_per_vsync: bool = False) -> int:
    """Replays a `.psxgpu` trace into `instance` of a Context (psxgpu_replay_dump).  `on_frame(frame_index)` runs after
    each frame has been submitted; ctx.display() is that frame's scan-out."""
    lib = _lib.load()
    cb = _lib.FRAME_CALLBACK((lambda _u, i: on_frame(i)) if on_frame else (lambda _u, i: None))
    flags = (FE_PAL if pal else 0) | (FE_INTERLACED if interlaced else 0) | (FE_FIELD_PER_VSYNC if field_per_vsync else 0)
    n = lib.psxgpu_replay_dump(ctx._ctx, instance, data, len(data), flags, crop_mode, cb, None)
    if n < 0:
        raise ValueError(f"psxgpu_replay_dump failed ({n}): {lib.psxgpu_last_error(ctx._ctx).decode()}")
    return n


def state_file_info(file_bytes: bytes) -> dict:
    """SAVE_STATE_HEADER of a DuckStation `.sav` file (src/core/save_state_version.h)."""
    info = _lib.StateFile()
    rc = _lib.load().psxgpu_state_file_info(file_bytes, len(file_bytes), C.byref(info))
    if rc != 0:
        raise ValueError(f"not a save state ({rc})")
    return {"version": info.version, "data_compression": info.data_compression,
            "data_compressed_size": info.data_compressed_size, "data_uncompressed_size": info.data_uncompressed_size,
            "offset_to_data": info.offset_to_data, "title": info.title.decode(errors="replace"),
            "serial": info.serial.decode(errors="replace")}


def find_gpu_section(state_data: bytes, version: int = 86) -> int:
    """Offset of the GPU section in decompressed save-state data (after the "GPU" marker, system.cpp:2497)."""
    off = _lib.load().psxgpu_state_find_gpu_section(state_data, len(state_data), version)
    if off < 0:
        raise ValueError(f"no GPU section found ({off})")
    return off


def load_save_state_file(file_bytes: bytes, fe: "Gp0Frontend", ctx=None, instance: int = 0) -> dict:
    """Seeds `instance` (and the front-end's registers) from a DuckStation save state file: header, decompression
    (none or zstd — through pyarrow's codec, there is no zstd binding of our own), GPU section search, load."""
    info = state_file_info(file_bytes)
    blob = file_bytes[info["offset_to_data"]: info["offset_to_data"] + info["data_compressed_size"]]
    if info["data_compression"] == 2:
        import pyarrow as pa

        blob = pa.decompress(blob, decompressed_size=info["data_uncompressed_size"], codec="zstd").to_pybytes()
    elif info["data_compression"] != 0:
        raise ValueError("save state compression %d is not supported (none and zstd are)" % info["data_compression"])
    off = find_gpu_section(blob, info["version"])
    fe.load_state(blob[off:], info["version"], ctx, instance)
    return info
