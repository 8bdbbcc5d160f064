"""CPU restatement of the reference's presenter shaders — TEST INFRASTRUCTURE ONLY (see oracle/psx_oracle.c header).

Follows, in numpy float32 and in the order the shader text evaluates things:
  GenerateDeinterlaceWeaveFragmentShader     src/core/video_shadergen.cpp:166-185
  GenerateDeinterlaceBlendFragmentShader     src/core/video_shadergen.cpp:187-205
  GenerateFastMADReconstructFragmentShader   src/core/video_shadergen.cpp:207-264
  GenerateChromaSmoothingFragmentShader      src/core/video_shadergen.cpp:266-330
and the state machine that drives them:
  VideoPresenter::Deinterlace                src/core/video_presenter.cpp:1208-1371
  VideoPresenter::ApplyChromaSmoothing       src/core/video_presenter.cpp:1385-1413
  GPU_SW::UpdateDisplay                      src/core/gpu_sw.cpp:386-441

PARITY UNPINNED by execution: the reference runs these as fragment shaders on the host graphics API, which cannot
run here.  What the API leaves open is fixed as: UNORM8 -> float = v / 255; float -> UNORM8 = round-half-even of
clamp(x, 0, 1) * 255; lerp(a, b, t) = a + (b - a) * t; dot() summed left to right without fused multiply-adds; a texel
fetch outside a texture (or from a field buffer that does not exist yet) returns 0.  Frames are (h, w, 4) uint8 RGBA.
"""
from __future__ import annotations

import numpy as np

F = np.float32
DISABLED, WEAVE, BLEND, ADAPTIVE, PROGRESSIVE = range(5)


def _unorm(a: np.ndarray) -> np.ndarray:
    return a.astype(F) / F(255.0)


def _to_u8(x: np.ndarray) -> np.ndarray:
    x = np.clip(x.astype(F), F(0.0), F(1.0)) * F(255.0)
    return np.rint(x).astype(np.uint8)  # rint = round half to even, like cvt.rni


def _fetch(tex: np.ndarray | None, xs: np.ndarray, ys: np.ndarray) -> np.ndarray:
    """texelFetch with robust access: (…, 4) uint8, zero outside / for a missing texture."""
    out = np.zeros(xs.shape + (4,), dtype=np.uint8)
    if tex is None or tex.size == 0:
        return out
    h, w = tex.shape[:2]
    ok = (xs >= 0) & (ys >= 0) & (xs < w) & (ys < h)
    out[ok] = tex[ys[ok], xs[ok]]
    return out


def weave(src: np.ndarray | None, field: int, target: np.ndarray) -> np.ndarray:
    full_h, w = target.shape[:2]
    ys, xs = np.mgrid[0:full_h, 0:w]
    sel = (ys & 1) == field
    out = target.copy()
    out[sel] = _fetch(src, xs, ys >> 1)[sel]
    return out


def blend(cur: np.ndarray | None, prev: np.ndarray | None, w: int, h: int) -> np.ndarray:
    ys, xs = np.mgrid[0:h, 0:w]
    c0, c1 = _unorm(_fetch(cur, xs, ys)), _unorm(_fetch(prev, xs, ys))
    return _to_u8((c0 + c1) * F(0.5))


def adaptive(f: list[np.ndarray | None], field: int, w: int, full_h: int) -> np.ndarray:
    rows, xs = np.mgrid[0:full_h, 0:w]
    uy = rows >> 1
    cur = _fetch(f[0], xs, uy)
    hn, cn, ln = (_unorm(_fetch(f[0], xs, uy - 1)[..., :3]), _unorm(_fetch(f[1], xs, uy)[..., :3]),
                  _unorm(_fetch(f[0], xs, uy + 1)[..., :3]))
    ho, co, lo = (_unorm(_fetch(f[2], xs, uy - 1)[..., :3]), _unorm(_fetch(f[3], xs, uy)[..., :3]),
                  _unorm(_fetch(f[2], xs, uy + 1)[..., :3]))
    s = F(0.08)
    mh, mc, ml = np.abs(hn - ho) - s, np.abs(cn - co) - s, np.abs(ln - lo) - s
    mmax = np.maximum(mh, np.maximum(mc, ml)).max(axis=-1)
    out = np.zeros((full_h, w, 4), dtype=np.uint8)
    out[..., 3] = 255
    present = (rows & 1) == field
    recon = ~present & (rows > 0) & (rows < full_h) & (mmax > F(0.0))
    prev = ~present & ~recon
    out[present, :3] = cur[present][:, :3]
    out[recon, :3] = _to_u8((hn + ln) / F(2.0))[recon]
    out[prev, :3] = _fetch(f[1], xs, uy)[prev][:, :3]
    return out


def _dot(v: np.ndarray, k: tuple[float, float, float]) -> np.ndarray:
    return (v[..., 0] * F(k[0]) + v[..., 1] * F(k[1])) + v[..., 2] * F(k[2])


def chroma_smoothing(src: np.ndarray) -> np.ndarray:
    h, w = src.shape[:2]
    ys, xs = np.mgrid[0:h, 0:w]
    bx, by = xs - 1, ys - 1
    lx, ly = np.maximum(bx & ~1, 0), np.maximum(by & ~1, 0)
    hx, hy = np.minimum(lx + 2, w - 1), np.minimum(ly + 2, h - 1)
    cx = (bx & 1).astype(F) * F(0.5) + F(0.25)
    cy = (by & 1).astype(F) * F(0.5) + F(0.25)

    def avg(x, y):
        v = _unorm(_fetch(src, x, y)[..., :3])
        v = v + _unorm(_fetch(src, x, y + 1)[..., :3])
        v = v + _unorm(_fetch(src, x + 1, y)[..., :3])
        v = v + _unorm(_fetch(src, x + 1, y + 1)[..., :3])
        return v * F(0.25)

    def lerp(a, b, t):
        return a + (b - a) * t[..., None]

    p = _unorm(src[..., :3])
    p00, p01, p10, p11 = avg(lx, ly), avg(lx, hy), avg(hx, ly), avg(hx, hy)
    s = lerp(lerp(p00, p10, cx), lerp(p01, p11, cx), cy)
    yuv = np.stack([_dot(p, (0.299, 0.587, 0.114)), _dot(s, (-0.14713, -0.28886, 0.436)),
                    _dot(s, (0.615, -0.51499, -0.10001))], axis=-1)
    rgb = np.stack([_dot(yuv, (1.0, 0.0, 1.13983)), _dot(yuv, (1.0, -0.39465, -0.58060)),
                    _dot(yuv, (1.0, 2.03211, 0.0))], axis=-1)
    out = np.empty((h, w, 4), dtype=np.uint8)
    out[..., :3] = _to_u8(rgb)
    out[..., 3] = 255
    return out


class PresenterModel:
    """VideoPresenter's deinterlacing state + GPU_SW::UpdateDisplay's sequencing, fed with scanned-out RGBA8 frames."""

    def __init__(self, mode: int, chroma: bool):
        self.mode, self.chroma = mode, chroma
        self.target: np.ndarray | None = None
        self.fields: list[np.ndarray | None] = [None] * 4
        self.current = 0

    def _set_target(self, w: int, h: int, preserve: bool) -> None:
        if self.target is not None and self.target.shape[:2] == (h, w):
            return
        fresh = np.zeros((h, w, 4), dtype=np.uint8)
        if preserve and self.target is not None:
            oh, ow = self.target.shape[:2]
            fresh[: min(h, oh), : min(w, ow)] = self.target[: min(h, oh), : min(w, ow)]
        self.target = fresh

    def _deinterlace(self, src: np.ndarray | None, field: int) -> np.ndarray | None:
        if self.mode == WEAVE:
            if src is not None:
                self._set_target(src.shape[1], src.shape[0] * 2, True)
            elif self.target is None:
                return None
            self.target = weave(src, field, self.target)
            return self.target
        if self.mode in (BLEND, ADAPTIVE):
            n = 2 if self.mode == BLEND else 4
            cur = self.current
            self.current = (self.current + 1) % n
            if src is not None:
                self._set_target(src.shape[1], src.shape[0] * (1 if self.mode == BLEND else 2), False)
                self.fields[cur] = src.copy()
            else:
                if self.target is None:
                    return None
                self.fields[cur] = None
            h, w = self.target.shape[:2]
            if self.mode == BLEND:
                self.target = blend(self.fields[cur], self.fields[(cur - 1) % 2], w, h)
            else:
                self.target = adaptive([self.fields[(cur - k) % 4] for k in range(4)], field, w, h)
            return self.target
        return src

    def update_display(self, frame: np.ndarray | None, *, is24: bool, interlaced: bool, field: int) -> np.ndarray | None:
        """frame: the RGBA8 scan-out of this UpdateDisplay, None if the display is disabled."""
        if frame is None:
            return self._deinterlace(None, field) if interlaced else None
        shown = frame
        if is24 and self.chroma:
            shown = chroma_smoothing(shown)
        if interlaced:
            shown = self._deinterlace(shown, field)
        return shown
