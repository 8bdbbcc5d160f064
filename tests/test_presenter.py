"""Presenter steps after scan-out: deinterlacing (weave / blend / adaptive) and 24-bit chroma
smoothing.  CPU tests pin the numpy restatement of the shader text (oracle/presenter_oracle.py) with properties the
shaders must have; the GPU tests compare psxgpu_present with that restatement driven by the oracle's scan-out."""
import numpy as np
import pytest

from duckstation_b200 import streams
from oracle import presenter_oracle as po
from tests.helpers import Oracle


def _noise(h, w, seed):
    img = np.random.default_rng(seed).integers(0, 256, (h, w, 4), dtype=np.uint8)
    img[..., 3] = 255
    return img


# ---------------------------------------------------------------- the restatement itself (CPU)
def test_weave_keeps_the_other_field():
    a, b = _noise(5, 7, 1), _noise(5, 7, 2)
    t = po.weave(a, 0, np.zeros((10, 7, 4), np.uint8))
    assert np.array_equal(t[0::2], a) and not t[1::2].any()
    t = po.weave(b, 1, t)
    assert np.array_equal(t[0::2], a) and np.array_equal(t[1::2], b)
    t = po.weave(None, 0, t)  # cleared display: the field's rows go black, the other field stays
    assert not t[0::2].any() and np.array_equal(t[1::2], b)


def test_blend_is_the_rounded_average_for_every_pair_of_bytes():
    a = np.repeat(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    b = a.T
    fa, fb = np.stack([a] * 4, -1), np.stack([b] * 4, -1)
    got = po.blend(fa, fb, 256, 256)[..., 0].astype(int)
    s = a.astype(int) + b.astype(int)
    assert np.array_equal(got[s % 2 == 0], (s // 2)[s % 2 == 0])
    assert (np.abs(got * 2 - s) <= 1).all()  # ties may round either way (float32 evaluation of (a + b) / 510)
    assert np.array_equal(po.blend(fa, None, 256, 256)[..., 1], po.blend(None, fa, 256, 256)[..., 1])


def test_adaptive_motion_threshold_is_21_levels_for_every_pair_of_bytes():
    """abs(a/255 - b/255) - 0.08 > 0 in float32 <=> |a - b| >= 21: the decision is integer-exact."""
    a = np.repeat(np.arange(256, dtype=np.uint8), 256).reshape(256, 256)
    b = a.T.copy()
    h, w = 256, 256
    zeros = np.zeros((h, w, 4), np.uint8)

    def rgba(x):
        return np.stack([x, x, x, np.full_like(x, 255)], -1)

    # field 1 is current: even output rows are reconstructed.  cn (f1) vs co (f3) carries the difference; f0 = f2.
    up, low = 40, 200
    f0 = np.zeros((h, w, 4), np.uint8)
    f0[0::2, :, :3], f0[1::2, :, :3] = up, low
    out = po.adaptive([f0, rgba(a), f0.copy(), rgba(b)], 1, w, 2 * h)
    motion = np.abs(a.astype(int) - b.astype(int)) >= 21
    for y in range(1, h):  # output row 2y: hn = f0[y - 1], ln = f0[y + 1]
        row = out[2 * y, :, 0]
        hn, ln = int(f0[y - 1, 0, 0]), (int(f0[y + 1, 0, 0]) if y + 1 < h else 0)
        assert set(np.unique(row[motion[y]])) <= {(hn + ln) // 2, (hn + ln + 1) // 2}, y
        assert np.array_equal(row[~motion[y]], a[y][~motion[y]]), y
    assert np.array_equal(out[0, :, 0], a[0])  # row 0 never reconstructs
    assert np.array_equal(out[1::2, :, :3], f0[..., :3]) and (out[..., 3] == 255).all()
    assert zeros.shape == f0.shape


def test_chroma_smoothing_keeps_flat_colour_and_luma():
    flat = np.full((12, 16, 4), (10, 200, 77, 255), np.uint8)
    out = po.chroma_smoothing(flat)
    assert np.array_equal(out[:-1, :-1], flat[:-1, :-1])  # last row / column average with out-of-range zeros
    img = _noise(16, 24, 3)
    out = po.chroma_smoothing(img)
    k = np.array([0.299, 0.587, 0.114])
    luma_in, luma_out = img[..., :3] @ k, out[..., :3] @ k
    clipped = (out[..., :3] == 0).any(-1) | (out[..., :3] == 255).any(-1)
    assert np.abs(luma_in - luma_out)[~clipped].max() < 1.0
    assert (out[..., 3] == 255).all()


def test_presenter_model_sequencing():
    m = po.PresenterModel(po.ADAPTIVE, chroma=True)
    a = _noise(4, 6, 4)
    out = m.update_display(a, is24=False, interlaced=False, field=0)
    assert out is a  # progressive 15-bit: untouched
    out = m.update_display(a, is24=True, interlaced=False, field=0)
    assert np.array_equal(out, po.chroma_smoothing(a))
    out = m.update_display(a, is24=False, interlaced=True, field=1)
    assert out.shape == (8, 6, 4) and np.array_equal(out[1::2, :, :3], a[..., :3])
    assert m.update_display(None, is24=False, interlaced=False, field=0) is None
    out = m.update_display(None, is24=False, interlaced=True, field=0)  # cleared display keeps deinterlacing
    assert out.shape == (8, 6, 4) and not out[0::2, :, :3].any()
    w = po.PresenterModel(po.WEAVE, chroma=False)
    assert w.update_display(None, is24=False, interlaced=True, field=0) is None
    t = w.update_display(a, is24=False, interlaced=True, field=0)
    t2 = w.update_display(_noise(5, 8, 5), is24=False, interlaced=True, field=1)  # resize preserves the overlap
    assert t2.shape == (10, 8, 4) and np.array_equal(t2[0:8:2, :6], t[0::2])


# ---------------------------------------------------------------- psxgpu_present against it (GPU)
def _display_sequence():
    D = streams.make_update_display
    return [
        dict(rec=D(0, 0, 320, 240, interlaced=True, field=False), is24=False, interlaced=True, field=0),
        dict(rec=D(0, 0, 320, 240, interlaced=True, field=True), is24=False, interlaced=True, field=1),
        dict(rec=D(16, 8, 320, 240, interlaced=True, field=False, interleaved=True), is24=False, interlaced=True, field=0),
        dict(rec=D(16, 8, 320, 240, interlaced=True, field=True, interleaved=True), is24=False, interlaced=True, field=1),
        dict(rec=D(33, 16, 256, 224, is24=True, X=32, interlaced=True, field=False, interleaved=True), is24=True, interlaced=True, field=0),
        dict(rec=D(33, 16, 256, 224, is24=True, X=32, interlaced=True, field=True, interleaved=True), is24=True, interlaced=True, field=1),
        dict(rec=D(0, 0, 320, 240, disabled=True, interlaced=True, field=False), is24=False, interlaced=True, field=0),
        dict(rec=D(640, 256, 368, 240, interlaced=True, field=True), is24=False, interlaced=True, field=1),
        dict(rec=D(0, 0, 320, 240, is24=True), is24=True, interlaced=False, field=0),
        dict(rec=D(0, 0, 320, 240), is24=False, interlaced=False, field=0),
        dict(rec=D(0, 0, 320, 240, disabled=True), is24=False, interlaced=False, field=0),
        dict(rec=D(512, 100, 400, 200, interlaced=True, field=False), is24=False, interlaced=True, field=0),
    ]


@pytest.mark.gpu
@pytest.mark.parametrize("chroma", [False, True], ids=["plain", "chroma"])
@pytest.mark.parametrize("mode", [po.DISABLED, po.WEAVE, po.BLEND, po.ADAPTIVE], ids=["disabled", "weave", "blend", "adaptive"])
def test_present_matches_the_restated_presenter(mode, chroma):
    from duckstation_b200 import Context
    from duckstation_b200.presenter import Presenter

    oracle = Oracle()
    rng = np.random.default_rng(100 + mode)
    model = po.PresenterModel(mode, chroma)
    with Context(1) as ctx, Presenter(ctx, mode, chroma) as pres:
        for step, d in enumerate(_display_sequence()):
            if step % 3 == 0:  # new picture every few fields
                vram = rng.integers(0, 65536, (512, 1024), dtype=np.uint16)
                ctx.load_state(vram)
                oracle.reset()
                oracle.load_vram(vram)
            rows = oracle.scanout(d["rec"], 0)
            frame = None
            if rows is not None:
                w, h = oracle.last_size
                frame = rows[:, : w * 4].reshape(h, w, 4).copy()
            want = model.update_display(frame, is24=d["is24"], interlaced=d["interlaced"], field=d["field"])
            got = pres.present(d["rec"])
            assert (got is None) == (want is None), step
            if want is not None:
                assert got.shape == want.shape, step
                bad = np.argwhere(got != want)
                assert bad.size == 0, f"step {step}: {len(bad)} bytes differ, first at {bad[0]}: {got[tuple(bad[0])]} != {want[tuple(bad[0])]}"
