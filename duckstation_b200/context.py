"""``Context`` — Python face of the C ABI (include/psxgpu.h).  Every call goes to libpsxgpu.so; nothing here
computes pixels, and there is no CPU fallback."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

VRAM_W, VRAM_H = 1024, 512


class PsxGpuError(RuntimeError):
    pass


class Context:
    """N independent 1024x512 VRAM instances on one CUDA device."""

    def __init__(self, n_instances: int = 1, device: int = 0, *, modulation_crop: bool = False, show_vram: bool = False,
                 display_format: int = 0, clut_slots: int = 0, encode_threads: int = 0, profile: bool = False,
                 no_persistent: bool = False, encode_mode: int = 0, launch_per_epoch: bool = False):
        """encode_mode: where submit_batch encodes — 0 auto (device for >= 16 instances), 1 host, 2 device.
        no_persistent: never use the cooperative latency-mode kernel; launch_per_epoch: neither that nor the instance-mode
        kernel (every epoch wave is its own launch)."""
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        opts = _lib.Opts()
        opts.struct_size = C.sizeof(_lib.Opts)
        opts.modulation_crop = int(modulation_crop)
        opts.show_vram = int(show_vram)
        opts.display_format = display_format
        opts.clut_slots = clut_slots
        opts.encode_threads = encode_threads
        opts.profile = int(profile)
        opts.reserved[0] = 3 if launch_per_epoch else int(no_persistent)
        opts.reserved[1] = encode_mode
        rc = self._lib.psxgpu_create(device, n_instances, C.byref(opts), C.byref(self._ctx))
        if rc != 0 or not self._ctx:
            self._ctx = C.c_void_p()
            raise PsxGpuError(f"psxgpu_create(device={device}, n={n_instances}) failed with {rc}: "
                              "a CUDA device is required (no CPU fallback)")
        self.n_instances = n_instances
        self.device = device
        self._keepalive = None

    # -- lifetime --
    def close(self) -> None:
        if getattr(self, "_ctx", None):
            self._lib.psxgpu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str) -> int:
        if rc < 0:
            raise PsxGpuError(f"{what} failed ({rc}): {self._lib.psxgpu_last_error(self._ctx).decode()}")
        return rc

    # -- commands --
    def submit(self, stream: bytes, instance: int = 0) -> None:
        self._check(self._lib.psxgpu_submit_wire(self._ctx, instance, stream, len(stream)), "psxgpu_submit_wire")

    def submit_batch(self, streams, first_instance: int = 0) -> None:
        n = len(streams)
        ptrs = (C.c_void_p * n)()
        sizes = (C.c_size_t * n)()
        keep = []
        for i, s in enumerate(streams):
            if isinstance(s, (bytes, bytearray)):
                buf = C.c_char_p(bytes(s)) if isinstance(s, bytearray) else C.c_char_p(s)
                keep.append(buf)
                ptrs[i] = C.cast(buf, C.c_void_p)
                sizes[i] = len(s)
            else:  # numpy / torch uint8 buffer
                arr = np.asarray(s)
                keep.append(arr)
                ptrs[i] = arr.ctypes.data
                sizes[i] = arr.nbytes
        self._check(self._lib.psxgpu_submit_batch(self._ctx, first_instance, n, ptrs, sizes), "psxgpu_submit_batch")

    def submit_batch_raw(self, ptrs, sizes, count: int, first_instance: int = 0) -> None:
        """``ptrs`` / ``sizes``: prebuilt ctypes arrays (avoids per-call marshalling in timed loops)."""
        self._check(self._lib.psxgpu_submit_batch(self._ctx, first_instance, count, ptrs, sizes), "psxgpu_submit_batch")

    def flush(self) -> None:
        self._check(self._lib.psxgpu_flush(self._ctx), "psxgpu_flush")

    def sync(self) -> None:
        self._check(self._lib.psxgpu_sync(self._ctx), "psxgpu_sync")

    def set_batch_chunk(self, instances_per_chunk: int) -> None:
        self._check(self._lib.psxgpu_set_batch_chunk(self._ctx, instances_per_chunk), "psxgpu_set_batch_chunk")

    def run_staged(self) -> None:
        self._check(self._lib.psxgpu_run_staged(self._ctx), "psxgpu_run_staged")

    # -- state --
    def clear_vram(self, instance: int = 0) -> None:
        self._check(self._lib.psxgpu_clear_vram(self._ctx, instance), "psxgpu_clear_vram")

    def load_state(self, vram: np.ndarray, clut: np.ndarray | None = None, instance: int = 0) -> None:
        v = np.ascontiguousarray(vram, dtype=np.uint16)
        assert v.size == VRAM_W * VRAM_H
        c = None if clut is None else np.ascontiguousarray(clut, dtype=np.uint16)
        self._check(self._lib.psxgpu_load_state(self._ctx, instance, v.ctypes.data, None if c is None else c.ctypes.data),
                    "psxgpu_load_state")

    def save_state(self, instance: int = 0) -> tuple[np.ndarray, np.ndarray]:
        v = np.empty((VRAM_H, VRAM_W), dtype=np.uint16)
        c = np.empty(256, dtype=np.uint16)
        self._check(self._lib.psxgpu_save_state(self._ctx, instance, v.ctypes.data, c.ctypes.data), "psxgpu_save_state")
        return v, c

    def read_vram(self, instance: int = 0, x: int = 0, y: int = 0, w: int = VRAM_W, h: int = VRAM_H) -> np.ndarray:
        out = np.empty((h, w), dtype=np.uint16)
        self._check(self._lib.psxgpu_read_vram(self._ctx, instance, x, y, w, h, out.ctypes.data, w), "psxgpu_read_vram")
        return out

    def scanout(self, display_record: bytes, instance: int = 0) -> np.ndarray | None:
        """GPU_SW::UpdateDisplay for one UpdateDisplay record; returns (h, w, 4) u8 or (h, w) u16, None if disabled."""
        w, h = C.c_uint32(), C.c_uint32()
        buf = np.zeros(1024 * 512 * 4 + 64, dtype=np.uint8)
        rec = C.create_string_buffer(bytes(display_record), len(display_record))
        rc = self._check(self._lib.psxgpu_scanout(self._ctx, instance, rec, buf.ctypes.data, 4096, C.byref(w), C.byref(h)),
                         "psxgpu_scanout")
        if rc == 1:
            return None
        rows = buf[: 4096 * h.value].reshape(h.value, 4096)
        return rows

    def display(self):
        w, h, s, r = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        p = self._lib.psxgpu_display(self._ctx, C.byref(w), C.byref(h), C.byref(s), C.byref(r))
        if not p or w.value == 0:
            return None
        raw = np.frombuffer(C.string_at(p, s.value * h.value), dtype=np.uint8).reshape(h.value, s.value)
        bpp = 4 if r.value else 2
        return raw[:, : w.value * bpp].copy()

    # -- hashes --
    def hash(self, first_instance: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n_instances - first_instance if count is None else count
        out = np.empty(count, dtype=np.uint64)
        self._check(self._lib.psxgpu_hash(self._ctx, first_instance, count, out.ctypes.data_as(C.POINTER(C.c_uint64))),
                    "psxgpu_hash")
        return out

    def hash_into_device(self, device_ptr: int, first_instance: int = 0, count: int | None = None) -> None:
        count = self.n_instances - first_instance if count is None else count
        self._check(self._lib.psxgpu_hash_device(self._ctx, first_instance, count, C.c_void_p(device_ptr)),
                    "psxgpu_hash_device")

    # -- introspection --
    def stats(self) -> dict:
        s = _lib.Stats()
        self._check(self._lib.psxgpu_get_stats(self._ctx, C.byref(s)), "psxgpu_get_stats")
        return {n: int(getattr(s, n)) for n, _ in _lib.Stats._fields_ if n != "reserved"}

    def reset_stats(self) -> None:
        self._check(self._lib.psxgpu_reset_stats(self._ctx), "psxgpu_reset_stats")

    def timing(self) -> dict:
        t = _lib.Timing()
        self._check(self._lib.psxgpu_get_timing(self._ctx, C.byref(t)), "psxgpu_get_timing")
        return {n: (float(getattr(t, n)) if "ms" in n else int(getattr(t, n))) for n, _ in _lib.Timing._fields_
                if n != "reserved"}

    def raster_launch_ms(self) -> list[float]:
        buf = (C.c_float * 4096)()
        n = self._check(self._lib.psxgpu_get_raster_launch_ms(self._ctx, buf, 4096), "psxgpu_get_raster_launch_ms")
        return [float(buf[i]) for i in range(min(n, 4096))]


def vramhash64(vram: np.ndarray) -> int:
    """Host implementation of the device hash (sum over 32-bit words of mix64((i<<32)|word)); for cross-checks."""
    w = np.ascontiguousarray(vram, dtype=np.uint16).reshape(-1).view(np.uint32).astype(np.uint64)
    z = (np.arange(w.size, dtype=np.uint64) << np.uint64(32)) | w
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
        return int(z.sum(dtype=np.uint64))
