"""Presenter: GPU_SW::UpdateDisplay with the VideoPresenter steps (scan-out, 24-bit chroma smoothing, deinterlacing).

Thin ctypes wrapper over psxgpu_presenter_* / psxgpu_present (include/psxgpu.h); mirrors the reference's
`display_deinterlacing_mode` / `display_24bit_chroma_smoothing` settings (src/core/settings.h).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .context import Context, PsxGpuError

DEINTERLACE_DISABLED, DEINTERLACE_WEAVE, DEINTERLACE_BLEND, DEINTERLACE_ADAPTIVE, DEINTERLACE_PROGRESSIVE = range(5)


class Presenter:
    def __init__(self, ctx: Context, deinterlace_mode: int = DEINTERLACE_ADAPTIVE, chroma_smoothing: bool = False):
        self._ctx = ctx
        self._lib = ctx._lib
        handle = C.c_void_p()
        rc = self._lib.psxgpu_presenter_create(ctx._ctx, deinterlace_mode, int(chroma_smoothing), C.byref(handle))
        if rc != 0:
            raise PsxGpuError(f"psxgpu_presenter_create failed ({rc})")
        self._p = handle
        self._buf = np.zeros(4096 * 4096 // 4 * 4, dtype=np.uint8)  # 720 x 1152 RGBA8 is the largest real frame

    def close(self) -> None:
        if self._p:
            self._lib.psxgpu_presenter_destroy(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def present(self, display_record: bytes, instance: int = 0) -> np.ndarray | None:
        """One UpdateDisplay record -> the frame shown, (h, w, 4) u8; None if nothing is shown."""
        w, h = C.c_uint32(), C.c_uint32()
        rec = C.create_string_buffer(bytes(display_record), len(display_record))
        rc = self._lib.psxgpu_present(self._p, instance, rec, self._buf.ctypes.data, self._buf.nbytes, C.byref(w), C.byref(h))
        if rc < 0:
            raise PsxGpuError(f"psxgpu_present failed ({rc}): {self._lib.psxgpu_last_error(self._ctx._ctx).decode()}")
        if rc == 1 or w.value == 0 or h.value == 0:
            return None
        return self._buf[: w.value * h.value * 4].reshape(h.value, w.value, 4).copy()
