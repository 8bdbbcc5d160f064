"""GPU tests of everything around the raster kernel, all through the C ABI: golden vectors from the reference, scan-out,
read-back, state load/save, options, the batched scheduler, the C++ GPUBackend mirror and full-size property checks."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from duckstation_b200 import streams
from duckstation_b200.context import vramhash64
from tests.helpers import Oracle, describe_diff, sha256

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def _ctx(n=1, **kw):
    from duckstation_b200 import Context

    return Context(n, **kw)


@pytest.mark.parametrize("entry", GOLDEN["cases"], ids=lambda e: "cfg{}-seed{}-n{}-f{}".format(*e["case"]))
def test_reference_golden_vectors(entry):
    """SHA-256 of the full VRAM must equal what the reference's own rasteriser produced (tests/golden/make_golden.py)."""
    with _ctx() as ctx:
        ctx.submit(streams.generate(*entry["case"]))
        v = ctx.read_vram()
        assert sha256(v) == entry["vram_sha256"]
        assert int(ctx.hash()[0]) == entry["vramhash64"] == vramhash64(v)


@pytest.mark.parametrize("entry", GOLDEN["modulation_crop_cases"], ids=lambda e: "crop-cfg{}-seed{}".format(*e["case"][:2]))
def test_modulation_crop_option(entry):
    with _ctx(modulation_crop=True) as ctx:
        ctx.submit(streams.generate(*entry["case"]))
        assert sha256(ctx.read_vram()) == entry["vram_sha256"]


def test_trace_replay_with_scanout():
    """Config 5: per-frame VRAM and display image, 15-bit and 24-bit, through UpdateDisplay records."""
    s = streams.generate(5, 3, 12, 200)
    with _ctx() as ctx:
        for entry, frame in zip(GOLDEN["scanout"], streams.split_frames(s)):
            ctx.submit(frame)
            assert sha256(ctx.read_vram()) == entry["vram_sha256"], f"frame {entry['frame']}"
            if "scanout_sha256" in entry:
                img = ctx.display()
                assert img is not None and list(img.shape) == [entry["size"][1], entry["size"][0] * 4]
                assert sha256(img) == entry["scanout_sha256"], f"frame {entry['frame']}"


@pytest.mark.parametrize("fmt", [0, 1, 2, 3])
def test_scanout_windows_and_formats(fmt):
    s = streams.generate(5, 3, 12, 200)
    with _ctx(display_format=fmt) as ctx:
        ctx.submit(s)
        for win in GOLDEN["scanout_windows"]:
            if win["fmt"] != fmt and not win["params"].get("is24"):
                continue
            if win["params"].get("is24") and fmt != 0:
                continue
            p = win["params"]
            img = ctx.scanout(streams.make_update_display(**p))
            bpp = 4 if (fmt == 0 or p.get("is24")) else 2
            assert sha256(img[:, : p["width"] * bpp]) == win["sha256"], win
        assert ctx.scanout(streams.make_update_display(0, 0, 320, 240, disabled=True)) is None


def test_show_vram_option(oracle):
    s = streams.generate(2, 9, 500, 0)
    oracle.render(s)
    want = oracle.scanout(streams.make_update_display(0, 0, 1, 1), 0, show_vram=True)
    with _ctx(show_vram=True) as ctx:
        ctx.submit(s)
        got = ctx.scanout(streams.make_update_display(0, 0, 1, 1))
    assert got.shape == (512, 4096) and np.array_equal(got, want[:, :4096])


def test_read_vram_wraps_and_shadow(oracle):
    s = streams.generate(0, 41, 1500, 3)
    want = oracle.render(s)
    with _ctx() as ctx:
        ctx.submit(s)
        part = ctx.read_vram(0, 1000, 500, 60, 40)
        exp = want[np.ix_((500 + np.arange(40)) % 512, (1000 + np.arange(60)) % 1024)]
        assert np.array_equal(part, exp)
        # a ReadVRAM record refreshes the host shadow (the reference's g_vram contract)
        import struct

        ctx.submit(struct.pack("<IB3x4H", 16, streams.CMD_READ_VRAM, 0, 0, 1024, 512))
        from duckstation_b200 import _lib

        sh = _lib.load().psxgpu_shadow_vram(ctx._ctx, 0)
        got = np.frombuffer(C.string_at(sh, 1 << 20), dtype=np.uint16).reshape(512, 1024)
        assert np.array_equal(got, want)


def test_state_save_load_and_clear(oracle):
    a, b = streams.generate(2, 11, 1500, 0), streams.generate(0, 12, 800, 0)
    oracle.reset()
    oracle.exec(a)
    mid_vram, mid_clut = oracle.vram(), oracle.clut()
    oracle.exec(b)
    want = oracle.vram()
    with _ctx(2) as ctx:
        ctx.submit(a, 0)
        v, c = ctx.save_state(0)
        assert np.array_equal(v, mid_vram) and np.array_equal(c, mid_clut)
        ctx.load_state(v, c, instance=1)  # continue on another instance from the saved state
        ctx.submit(b, 1)
        assert np.array_equal(ctx.read_vram(1), want), describe_diff(ctx.read_vram(1), want)
        ctx.clear_vram(1)
        assert not ctx.read_vram(1).any()
        ctx.submit(a + b, 1)  # cleared == fresh
        assert np.array_equal(ctx.read_vram(1), want)


def test_incremental_submission_equals_one_shot(oracle):
    """Records fed one at a time with flushes in between (the drop-in call pattern) give the same VRAM."""
    s = streams.generate(3, 5, 1200, 0)
    want = oracle.render(s)
    with _ctx() as ctx:
        for i, (typ, off, size) in enumerate(streams.iter_records(s)):
            ctx.submit(s[off:off + size])
            if i % 97 == 0:
                ctx.flush()
        assert np.array_equal(ctx.read_vram(), want)


def test_batched_instances_and_hash(oracle):
    n = 48
    batch = [streams.generate(4, 1000 + i, 600 + 7 * i, 0) for i in range(n)]
    want = []
    for s in batch:
        oracle.render(s)
        want.append(oracle.hash64())
    with _ctx(n, profile=True) as ctx:
        ctx.submit_batch(batch)
        ctx.flush()
        got = [int(x) for x in ctx.hash()]
        assert got == want
        # replay from the HBM-resident command buffers is idempotent for this workload
        ctx.run_staged()
        assert [int(x) for x in ctx.hash()] == want
        # ragged: some instances idle, some with a second flush
        ctx.submit_batch(batch[:5], first_instance=10)
        ctx.flush()
        st = ctx.stats()
        assert st["kernel_launches"] > 0 and st["rejected"] == 0
        again = [int(x) for x in ctx.hash(10, 5)]
    for i in range(5):
        oracle.reset()
        oracle.exec(batch[10 + i])
        oracle.exec(batch[i])
        assert again[i] == oracle.hash64()


def test_batch_rejects_readback_records():
    from duckstation_b200 import PsxGpuError

    with _ctx(2) as ctx:
        with pytest.raises(PsxGpuError):
            ctx.submit_batch([streams.generate(5, 1, 1, 50), b""])
        with pytest.raises(PsxGpuError):
            ctx.submit(b"\x07\x00\x00\x00\x16\x00\x00\x00")  # size 7: malformed


def test_cpp_backend_mirror(oracle):
    """psx::CudaGpuBackend::HandleCommand record by record (the GPUBackend call pattern) + counters + display."""
    lib = C.CDLL(os.path.join(ROOT, "duckstation_b200", "libpsxgpu.so"))
    lib.psxbackend_create.restype = C.c_void_p
    lib.psxbackend_create.argtypes = [C.c_int, C.c_int]
    lib.psxbackend_handle_stream.restype = C.c_long
    lib.psxbackend_handle_stream.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.psxbackend_read_all.restype = C.POINTER(C.c_uint16)
    lib.psxbackend_read_all.argtypes = [C.c_void_p]
    lib.psxbackend_destroy.argtypes = [C.c_void_p]
    lib.psxbackend_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
    s = streams.generate(5, 6, 3, 300) + streams.generate(3, 6, 600, 0)
    want = oracle.render(s)
    b = lib.psxbackend_create(0, 0)
    assert b
    try:
        n = lib.psxbackend_handle_stream(b, s, len(s))
        assert n == sum(1 for _ in streams.iter_records(s))
        got = np.frombuffer(C.string_at(lib.psxbackend_read_all(b), 1 << 20), dtype=np.uint16).reshape(512, 1024)
        assert np.array_equal(got, want), describe_diff(got, want)
        cnt = (C.c_uint32 * 5)()
        lib.psxbackend_counters(b, cnt)
        # expected counters, counted the way GPUBackend::HandleCommand does (gpu_backend.cpp:449-537)
        import struct

        writes = copies = verts = prims = 0
        for typ, off, size in streams.iter_records(s):
            nv = struct.unpack_from("<H", s, off + 10)[0] if typ in (22, 24, 25) else 0
            if typ == 17:
                writes += 1
            elif typ == 18:
                copies += 1
            elif typ == 22:
                verts, prims = verts + nv, prims + 1
            elif typ == 24:
                verts, prims = verts + 1, prims + 1
            elif typ == 25:
                verts, prims = verts + nv, prims + nv // 2
        assert list(cnt) == [0, writes, copies, verts, prims]
    finally:
        lib.psxbackend_destroy(b)


def test_full_size_properties(oracle):
    """Size-independent properties at BASELINE sizes: determinism, order sensitivity, linearity of the hash under
    disjoint writes."""
    s = streams.generate(2, 1, 20000, 0)
    with _ctx(2) as ctx:
        ctx.submit(s, 0)
        ctx.submit(s, 1)
        h = ctx.hash()
        assert h[0] == h[1]  # deterministic across instances
        recs = [s[o:o + z] for _, o, z in streams.iter_records(s)]
        ctx.clear_vram(1)
        ctx.submit(b"".join(recs[:8] + recs[8:][::-1]), 1)  # same primitives, reversed order
        assert ctx.hash(1, 1)[0] != h[0]  # the path is order dependent: a shuffled bin order must change the result


@pytest.mark.parametrize("case", [(3, 8, 2500, 0), (2, 8, 3000, 0), (0, 51, 2500, 3)], ids=["mixed", "textured", "fuzz"])
def test_latency_mode_equals_launch_per_epoch(oracle, case):
    """The resident kernels (the cooperative latency-mode kernel, or the instance-mode kernel when the stream is cut into
    many short epochs) and the launch-per-epoch path must produce the same VRAM as the oracle."""
    s = streams.generate(*case)
    want = oracle.render(s)
    with _ctx(1) as ctx:  # latency mode
        ctx.submit(s)
        a = ctx.read_vram()
        launches_persistent = ctx.stats()["kernel_launches"]
    with _ctx(1, profile=True, launch_per_epoch=True) as ctx:  # one launch per epoch
        ctx.submit(s)
        b = ctx.read_vram()
        launches_per_epoch = ctx.stats()["kernel_launches"]
    assert np.array_equal(a, want), describe_diff(a, want)
    assert np.array_equal(b, want), describe_diff(b, want)
    if "PSXGPU_INSTANCE_MODE" not in os.environ:  # (forced instance mode: one launch either way)
        assert launches_persistent < launches_per_epoch
    with _ctx(6) as ctx:  # six instances: throughput kernels
        ctx.submit_batch([s] * 6)
        ctx.flush()
        hs = ctx.hash()
        assert all(int(h) == oracle.hash64(want) for h in hs)


def test_batch_chunking_is_transparent(oracle):
    n = 40
    batch = [streams.generate(4, 2000 + i, 300, 0) for i in range(n)]
    want = []
    for s in batch:
        oracle.render(s)
        want.append(oracle.hash64())
    for chunk in (0, 7, 16, n):
        with _ctx(n) as ctx:
            ctx.set_batch_chunk(chunk)
            ctx.submit_batch(batch)
            ctx.flush()
            assert [int(x) for x in ctx.hash()] == want, f"chunk={chunk}"
            ctx.run_staged()  # replays every chunk of the group
            assert [int(x) for x in ctx.hash()] == want, f"chunk={chunk} replay"


def test_precise_record_types(oracle):
    s = streams.to_precise(streams.generate(3, 6, 1500, 0))
    want = oracle.render(s)
    with _ctx() as ctx:
        ctx.submit(s)
        got = ctx.read_vram()
    assert np.array_equal(got, want), describe_diff(got, want)


@pytest.mark.parametrize("interlaced", [False, True], ids=["progressive", "interlaced"])
def test_psxgpu_trace_replay_through_the_gp0_frontend(oracle, interlaced):
    """A `.psxgpu` trace (GP0/GP1 port words + VSync packets) replayed by psxgpu_replay_dump: after every frame the VRAM
    and the scanned-out image equal what the oracle makes of the records the (host-only) front-end decodes."""
    from duckstation_b200.frontend import CROP_BORDERS, Gp0Frontend, replay_dump
    from tests import gp0_assembler as asm
    from tests.test_gp0_frontend import gp1_mode, strip_interlaced_and_display

    direct = strip_interlaced_and_display(streams.generate(0, 77, 900, 3))
    offs = [off for typ, off, _ in streams.iter_records(direct) if typ != asm.CMD_CLUT]
    cuts = [0] + [offs[len(offs) * k // 4] for k in (1, 2, 3)] + [len(direct)]
    a = asm.Gp0Assembler()
    a.gp1(0x00000000)
    a.gp1(gp1_mode(hres1=1, vres=1 if interlaced else 0, interlace=1 if interlaced else 0))
    a.gp1(0x06000000 | 0x260 | (0xC60 << 12))
    a.gp1(0x07000000 | 16 | (256 << 10))
    a.gp1(0x03000000)
    for k in range(4):
        # 480 interleaved lines from row 8k stay inside VRAM: past the end the reference's 15-bit fast path reads out
        # of bounds, which no implementation can reproduce
        a.gp1(0x05000000 | (k * 160) | ((k * 8) << 10))
        if k == 2:
            a.gp1(gp1_mode(hres1=1, depth24=1, vres=1 if interlaced else 0, interlace=1 if interlaced else 0))
        a.add_stream(direct[cuts[k]:cuts[k + 1]])
        if k == 1:  # interlaced rendering applies to these (line-skipping fill and rectangle)
            a.gp0(0x02102030, 0, 64 | (64 << 16), 0x60FFFFFF, 8 | (8 << 16), 40 | (40 << 16))
        a.vsync()
    dump = asm.write_dump(a.events, max_gp0_words=13)

    fe = Gp0Frontend(interlaced=interlaced, crop_mode=CROP_BORDERS)
    assert fe.feed_dump(dump) == 4
    frames = streams.split_frames(fe.records())
    assert len(frames) == 4
    want = []
    oracle.reset()
    for f in frames:
        oracle.exec(f)
        disp = [f[off:off + size] for typ, off, size in streams.iter_records(f) if typ == 9][-1]
        img = oracle.scanout(disp, 0)
        want.append((oracle.vram().copy(), None if img is None else img[:, : oracle.last_size[0] * 4].copy()))
    if interlaced:
        recs = [f[off:off + size] for f in frames for typ, off, size in streams.iter_records(f) if typ == 24]
        assert any(r[8] & 1 for r in recs)

    got = []
    with _ctx() as ctx:
        n = replay_dump(ctx, dump, interlaced=interlaced, crop_mode=CROP_BORDERS,
                        on_frame=lambda i: got.append((ctx.read_vram(), ctx.display())))
        assert n == 4 and len(got) == 4
    for i, ((gv, gi), (wv, wi)) in enumerate(zip(got, want)):
        assert np.array_equal(gv, wv), f"frame {i}: {describe_diff(gv, wv)}"
        assert (gi is None) == (wi is None)
        if wi is not None:
            assert gi.shape == wi.shape and np.array_equal(gi, wi), f"frame {i} scan-out"


DEVICE_ENCODER_CASES = ([(0, s, 600, f) for s in (3, 4, 5, 6, 7, 8) for f in (0, 1, 2, 3)] +
                        [(1, 2, 1500, 0), (2, 2, 1500, 0), (3, 2, 800, 0), (4, 2, 800, 0), (5, 2, 6, 200), (6, 2, 1200, 0), (6, 3, 2000, 0)])


def _strip_barriers(stream: bytes) -> bytes:
    return b"".join(stream[off:off + size] for typ, off, size in streams.iter_records(stream) if typ not in (7, 9, 12, 15))


def test_device_encoder_matches_oracle_and_host_encoder(oracle):
    """psxgpu_submit_batch with the encoder on the device (one warp per instance) against the oracle, on every kind of
    stream the generator makes; the host-encoded batch of the same streams must agree on the epoch statistics too."""
    batch = [_strip_barriers(streams.generate(*c)) for c in DEVICE_ENCODER_CASES]
    want = [oracle.hash64(oracle.render(s)) for s in batch]
    stats = {}
    for mode in (2, 1):
        with _ctx(len(batch), encode_mode=mode, no_persistent=True) as ctx:
            ctx.submit_batch(batch)
            ctx.flush()
            got = [int(h) for h in ctx.hash()]
            bad = [DEVICE_ENCODER_CASES[i] for i in range(len(batch)) if got[i] != want[i]]
            assert not bad, (mode, bad)
            stats[mode] = ctx.stats()
    for key in ("wire_commands", "packed_prims", "epochs", "cuts_raw", "cuts_war", "cuts_clut", "serial_self_sampling",
                "serial_copy", "serial_upload", "clut_ops", "clut_cache_hits", "rejected"):
        assert stats[2][key] == stats[1][key], (key, stats[2][key], stats[1][key])


def test_instance_mode_batch_with_hazards(oracle):
    """Generator config 6 (config 4 + self-sampling primitives + render-to-texture pairs: tens of epochs per instance):
    a device-encoded batch takes the instance-mode kernel (one CTA per instance, one launch for all epochs) and must
    match the oracle, and the launch-per-epoch path must give the same hashes."""
    batch = [streams.generate(6, 40 + i, 1500, 0) for i in range(24)]
    want = [oracle.hash64(oracle.render(s)) for s in batch]
    launches = {}
    for per_epoch in (False, True):
        with _ctx(len(batch), launch_per_epoch=per_epoch) as ctx:
            ctx.submit_batch(batch)
            ctx.flush()
            assert [int(h) for h in ctx.hash()] == want, per_epoch
            st = ctx.stats()
            assert st["epochs"] > 20 * len(batch)
            launches[per_epoch] = st["kernel_launches"]
    if "PSXGPU_INSTANCE_MODE" not in os.environ:
        assert launches[False] < launches[True] // 4


def test_truncated_records_host_and_device_encoders_agree(oracle):
    """A record that is shorter than its own counts ask for is dropped and counted by both encoders (never read past)."""
    from tests.test_hostsim import _area, _rect, _truncated_records
    good = _area(0, 0, 1023, 511) + _rect(10, 10, 50, 50, 0x3050F0)
    tail = _rect(100, 100, 20, 20, 0x10F010)
    s = good + b"".join(_truncated_records()) + tail
    want = oracle.hash64(oracle.render(good + tail))
    batch = [s] * 16
    for mode in (2, 1):
        with _ctx(len(batch), encode_mode=mode, no_persistent=True) as ctx:
            ctx.submit_batch(batch)
            ctx.flush()
            assert [int(h) for h in ctx.hash()] == [want] * len(batch), mode
            assert ctx.stats()["rejected"] == 4 * len(batch), (mode, ctx.stats()["rejected"])
    with _ctx(1) as ctx:
        ctx.submit(s)
        assert int(ctx.hash()[0]) == want and ctx.stats()["rejected"] == 4


def test_device_encoder_state_carries_across_flushes_and_paths(oracle):
    """Drawing area, current CLUT and the snapshot table survive a flush and the hand-over between the device encoder
    (batches) and the host encoder (psxgpu_submit_wire), in both directions."""
    n = 20
    full = [_strip_barriers(streams.generate(0, 100 + i, 900, 3)) for i in range(n)]
    cut = []
    for s in full:
        offs = [off for typ, off, _ in streams.iter_records(s)]
        cut.append([offs[len(offs) // 3], offs[2 * len(offs) // 3]])
    want = [oracle.hash64(oracle.render(s)) for s in full]
    with _ctx(n, encode_mode=2) as ctx:
        ctx.submit_batch([s[:c[0]] for s, c in zip(full, cut)])       # device
        ctx.flush()
        for i, (s, c) in enumerate(zip(full, cut)):                   # host, one instance at a time
            ctx.submit(s[c[0]:c[1]], instance=i)
        ctx.submit_batch([s[c[1]:] for s, c in zip(full, cut)])       # device again
        ctx.flush()
        got = [int(h) for h in ctx.hash()]
    assert got == want
    # several device-encoded flushes in a row, chunked
    with _ctx(n, encode_mode=2) as ctx:
        ctx.set_batch_chunk(7)
        for part in range(3):
            ctx.submit_batch([s[([0] + c)[part]:(c + [len(s)])[part]] for s, c in zip(full, cut)])
            ctx.flush()
        assert [int(h) for h in ctx.hash()] == want


def test_device_encoder_rejects_bad_streams_and_handles_empty_ones(oracle):
    from duckstation_b200 import PsxGpuError

    good = _strip_barriers(streams.generate(2, 5, 300, 0))
    with _ctx(4, encode_mode=2) as ctx:
        with pytest.raises(PsxGpuError):
            ctx.submit_batch([good, good[:-4], good, good])
        with pytest.raises(PsxGpuError):
            ctx.submit_batch([good, good + streams.make_update_display(0, 0, 320, 240), good, good])
    with _ctx(4, encode_mode=2) as ctx:
        ctx.submit_batch([good, b"", good, b""])
        ctx.flush()
        h = [int(x) for x in ctx.hash()]
        assert h[0] == h[2] == oracle.hash64(oracle.render(good))
        assert h[1] == h[3] == oracle.hash64(np.zeros((512, 1024), dtype=np.uint16))
        ctx.run_staged()  # replaying a device-encoded flush on top of its own result must be deterministic
        ctx.sync()
        again = [int(x) for x in ctx.hash()]
        oracle.render(good)
        oracle.exec(good)
        assert again[0] == oracle.hash64(oracle.vram())


def test_config5_psxgpu_trace_replay_matches_reference_goldens():
    """The config-5 frame trace as a `.psxgpu` file through psxgpu_replay_dump: per-frame VRAM and display hashes are
    the ones the reference's rasteriser produced (tests/golden)."""
    from duckstation_b200.frontend import CROP_BORDERS, replay_dump
    from tests.test_gp0_frontend import config5_as_psxgpu_trace

    dump = config5_as_psxgpu_trace()
    got = []
    with _ctx() as ctx:
        n = replay_dump(ctx, dump, crop_mode=CROP_BORDERS, on_frame=lambda i: got.append((sha256(ctx.read_vram()), ctx.display())))
    assert n == len(GOLDEN["scanout"]) == len(got)
    for entry, (vram_sha, img) in zip(GOLDEN["scanout"], got):
        assert vram_sha == entry["vram_sha256"], f"frame {entry['frame']}"
        if "scanout_sha256" in entry:
            assert img is not None and list(img.shape) == [entry["size"][1], entry["size"][0] * 4]
            assert sha256(img) == entry["scanout_sha256"], f"frame {entry['frame']}"


def _tight(stream: bytes) -> bytes:
    """Re-frames a stream with every record cut to the bytes it needs, rounded to 4 (the reference rounds to 16): records
    are then not 16-byte aligned any more, which the device encoder has to handle on its record-at-a-time path."""
    import struct

    out = bytearray()
    for typ, off, size in streams.iter_records(stream):
        rec = bytearray(stream[off:off + size])
        if typ == 22:
            need = 20 + 16 * struct.unpack_from("<H", rec, 10)[0]
        elif typ == 25:
            need = 20 + 12 * struct.unpack_from("<H", rec, 10)[0]
        elif typ == 24:
            need = 40
        elif typ in (16, 18, 19):
            need = 24
        elif typ == 20:
            need = 12
        elif typ == 17:
            w, h = struct.unpack_from("<HH", rec, 12)
            need = 18 + 2 * w * h
        else:
            need = size
        need = min(size, (need + 3) & ~3)
        struct.pack_into("<I", rec, 0, need)
        out += rec[:need]
    return bytes(out)


def test_device_encoder_precise_and_unaligned_records(oracle):
    base = [_strip_barriers(streams.generate(*c)) for c in [(0, 60 + i, 500, 3) for i in range(6)] + [(3, 9, 700, 0), (2, 3, 700, 0)]]
    batch = [streams.to_precise(s) for s in base] + [_tight(s) for s in base] + [_tight(streams.to_precise(s)) for s in base[:4]]
    want = [oracle.hash64(oracle.render(s)) for s in base]
    want = want + want + want[:4]
    assert any(any(off % 16 for _, off, _ in streams.iter_records(s)) for s in batch[8:16])
    with _ctx(len(batch), encode_mode=2) as ctx:
        ctx.submit_batch(batch)
        ctx.flush()
        got = [int(h) for h in ctx.hash()]
    assert got == want, [i for i in range(len(batch)) if got[i] != want[i]]


def test_device_encoder_page_locked_inputs(oracle):
    """Streams that already sit in page-locked memory are copied to the device from where they are: neighbours in one
    copy, any order, gaps; a misaligned or pageable stream in the chunk sends the whole chunk through the staging copy."""
    import ctypes as C

    import torch

    n = 24
    batch = [_strip_barriers(streams.generate(*c)) for c in [(0, 200 + i, 250, 3) for i in range(12)] + [(4, 300 + i, 200, 0) for i in range(12)]]
    want = [oracle.hash64(oracle.render(s)) for s in batch]

    def run(layout):
        """layout: list of byte offsets (into one pinned buffer) per stream"""
        total = max(o + len(s) for o, s in zip(layout, batch)) + 64
        pinned = torch.empty(total + 64, dtype=torch.uint8).pin_memory()
        base = pinned.data_ptr()
        base += (-base) % 16
        view = np.ctypeslib.as_array((C.c_uint8 * total).from_address(base))
        ptrs = (C.c_void_p * n)()
        sizes = (C.c_size_t * n)()
        for i, (o, s) in enumerate(zip(layout, batch)):
            view[o:o + len(s)] = np.frombuffer(s, dtype=np.uint8)
            ptrs[i] = base + o
            sizes[i] = len(s)
        with _ctx(n, encode_mode=2) as ctx:
            ctx.set_batch_chunk(10)
            ctx.submit_batch_raw(ptrs, sizes, n)
            ctx.flush()
            got = [int(h) for h in ctx.hash()]
            h2d = ctx.stats()["h2d_bytes"]
        del pinned
        return got, h2d

    a16 = lambda v: (v + 15) & ~15
    # back to back
    offs, at = [], 0
    for s in batch:
        offs.append(at)
        at = a16(at + len(s))
    got, h2d_packed = run(offs)
    assert got == want
    # gaps of 48 bytes (still one run) and of 4 KB (separate runs), streams in reverse address order
    offs, at = [], 0
    for i, s in enumerate(batch):
        offs.append(at)
        at = a16(at + len(s)) + (48 if i % 2 else 4096)
    got, _ = run(offs)
    assert got == want
    # streams in decreasing address order
    offs, at = [0] * n, 0
    for i in reversed(range(n)):
        offs[i] = at
        at = a16(at + len(batch[i])) + 32
    got, _ = run(offs)
    assert got == want
    # a misaligned stream: the chunk falls back to the staging copy, same result
    offs, at = [], 4
    for s in batch:
        offs.append(at)
        at = a16(at + len(s)) + 4
    got, _ = run(offs)
    assert got == want
    assert h2d_packed >= sum(len(s) for s in batch)


def _rec(typ: int, payload: bytes) -> bytes:
    import struct

    size = (8 + len(payload) + 15) & ~15
    return struct.pack("<IB3x", size, typ) + payload + bytes(size - 8 - len(payload))


def _clut_pressure_stream(seed: int, n: int) -> bytes:
    """More distinct palettes than CLUT snapshot slots inside one hazard-free run: noise in the right half of VRAM,
    then `n` small 4-bit textured sprites on the left, each with a palette register of its own."""
    import struct

    rng = np.random.default_rng(seed)
    out = bytearray()
    out += _rec(19, struct.pack("<4I", 0, 0, 511, 511))
    noise = rng.integers(0, 65536, size=(512, 512), dtype=np.uint16)
    out += _rec(17, struct.pack("<4HBB", 512, 0, 512, 512, 0, 0) + noise.tobytes())
    for i in range(n):
        reg = (32 + (i % 32)) | (((i // 32) % 512) << 6)  # x = 512 + 16*(i%32), y = i//32: all different
        out += _rec(20, struct.pack("<HB", reg, 0))
        flags0 = 0x10 | (0x20 if i % 3 == 0 else 0)
        draw_mode = 8 | (0 << 7)  # texture page x = 512, 4-bit
        hdr = struct.pack("<BBHHH4B", flags0, 0, 1, draw_mode, reg, 0xFF, 0xFF, 0, 0)
        x, y = int(rng.integers(0, 480)), int(rng.integers(0, 480))
        out += _rec(24, hdr + struct.pack("<HHH2xiiI", 12, 9, int(rng.integers(0, 65536)), x, y, 0x808080))
    return bytes(out)


def test_device_encoder_slot_pressure_and_long_epochs(oracle):
    """Two paths that ordinary streams rarely take: more live palettes than snapshot slots (the CLUT cache cuts the
    epoch) and more than 65535 primitives without a hazard (the epoch length cap)."""
    long_run = _strip_barriers(streams.generate(1, 5, 70000, 0))
    batch = [_clut_pressure_stream(1, 700), _clut_pressure_stream(2, 300), long_run] + [
        _strip_barriers(streams.generate(4, 900 + i, 100, 0)) for i in range(13)]
    want = [oracle.hash64(oracle.render(s)) for s in batch]
    stats = {}
    for mode, slots in ((2, 0), (1, 0), (2, 64), (1, 64)):
        with _ctx(len(batch), encode_mode=mode, clut_slots=slots) as ctx:
            ctx.submit_batch(batch)
            ctx.flush()
            got = [int(h) for h in ctx.hash()]
            assert got == want, (mode, slots, [i for i in range(len(batch)) if got[i] != want[i]])
            stats[(mode, slots)] = ctx.stats()
    for slots in (0, 64):
        for key in ("packed_prims", "epochs", "cuts_raw", "cuts_war", "cuts_clut", "clut_ops", "clut_cache_hits"):
            assert stats[(2, slots)][key] == stats[(1, slots)][key], (slots, key, stats[(2, slots)][key], stats[(1, slots)][key])
    assert stats[(2, 0)]["cuts_clut"] > 0 and stats[(2, 64)]["cuts_clut"] > stats[(2, 0)]["cuts_clut"]
    assert stats[(2, 0)]["epochs"] > len(batch) + 2


def test_every_command_on_every_tile_overflows_the_bin_lists(oracle):
    """Binning reads one word per (32 commands, tile) and expands it into a 160-entry list: 1100 blended rectangles that
    each cover all of VRAM fill every tile's list seven times over, with the epoch starting and ending inside a 32-command
    group (an odd number of state records in front) — and blending makes any reordering or loss visible."""
    import struct

    rng = np.random.default_rng(11)
    out = bytearray()
    out += _rec(19, struct.pack("<4I", 0, 0, 1023, 511))
    for _ in range(3):  # FillVRAM records in front: the rectangles do not start at a multiple of 32
        out += _rec(16, struct.pack("<4HIBB", int(rng.integers(0, 512)), int(rng.integers(0, 256)), 64, 64, int(rng.integers(0, 1 << 24)), 0, 0))
    for i in range(1100):
        mode = 0 if i % 10 < 7 else int(rng.integers(1, 4))  # mostly averaging: keeps the picture from saturating
        if i % 3:  # all of VRAM, skipping pixels whose mask bit is set
            flags0 = 0x40 | 0x08
            x, y, w, h = int(rng.integers(-3, 3)), int(rng.integers(-3, 3)), 1023, 511
        else:  # a small one that sets the mask bit: these pixels keep the value they have at this point of the order
            flags0 = 0x40 | 0x04 | (0x08 if i % 2 else 0)
            x, y, w, h = int(rng.integers(0, 1000)), int(rng.integers(0, 500)), int(rng.integers(8, 120)), int(rng.integers(8, 120))
        hdr = struct.pack("<BBHHH4B", flags0, 0, 1, mode << 5, 0, 0xFF, 0xFF, 0, 0)
        out += _rec(24, hdr + struct.pack("<HHH2xiiI", w, h, 0, x, y, int(rng.integers(0, 1 << 24))))
    stream = bytes(out)
    want = oracle.render(stream)
    with _ctx() as ctx:
        ctx.submit(stream)
        got = ctx.read_vram()
        assert np.array_equal(got, want), describe_diff(got, want)
        assert ctx.stats()["epochs"] <= 3
    with _ctx(16, encode_mode=2) as ctx:  # the same through the device encoder + raster_loop_kernel
        ctx.submit_batch([stream] * 16)
        ctx.flush()
        assert [int(h) for h in ctx.hash()] == [oracle.hash64(want)] * 16


def test_full_tile_uploads_with_odd_strides_and_mask_set(oracle):
    """Uploads large enough to cover whole 64x32 tiles take a straight-copy path (32-bit payload loads when the row
    starts on an even halfword): odd widths make the parity alternate row by row, odd origins shift the tile phase,
    set_mask ORs bit 15 in, check_mask must fall back to the general path — through both encoders."""
    import struct

    rng = np.random.default_rng(31)
    out = bytearray()
    out += _rec(16, struct.pack("<4HIBB", 0, 0, 1023, 511, 0x123456, 0, 0))
    for (x, y, w, h, set_mask, check_mask) in [(0, 0, 256, 128, 0, 0), (3, 5, 131, 97, 1, 0), (700, 300, 199, 64, 0, 0),
                                               (64, 32, 64, 32, 1, 0), (500, 100, 301, 130, 0, 1), (129, 257, 385, 200, 1, 1),
                                               (900, 480, 200, 60, 0, 0)]:
        pix = rng.integers(0, 65536, size=w * h, dtype=np.uint16)
        out += _rec(17, struct.pack("<4HBB", x, y, w, h, set_mask, check_mask) + pix.tobytes())
    stream = bytes(out)
    want = oracle.render(stream)
    with _ctx() as ctx:
        ctx.submit(stream)
        got = ctx.read_vram()
        assert np.array_equal(got, want), describe_diff(got, want)
    with _ctx(16, encode_mode=2) as ctx:
        ctx.submit_batch([stream] * 16)
        ctx.flush()
        assert [int(h) for h in ctx.hash()] == [oracle.hash64(want)] * 16
    with _ctx(16, encode_mode=1) as ctx:
        ctx.submit_batch([stream] * 16)
        ctx.flush()
        assert [int(h) for h in ctx.hash()] == [oracle.hash64(want)] * 16


def test_full_tile_copies(oracle):
    """Non-overlapping VRAM-to-VRAM copies large enough to cover whole tiles: even / odd source phase, source rows and
    columns that wrap (rows are fine for the straight-copy path, wrapping columns are not), mask set and mask test."""
    import struct

    rng = np.random.default_rng(41)
    out = bytearray()
    noise = rng.integers(0, 65536, size=(512, 512), dtype=np.uint16)
    out += _rec(17, struct.pack("<4HBB", 0, 0, 512, 512, 0, 0) + noise.tobytes())
    for (sx, sy, dx, dy, w, h, set_mask, check_mask) in [(0, 0, 512, 0, 256, 128, 0, 0), (3, 7, 515, 130, 199, 97, 1, 0),
                                                         (100, 400, 640, 256, 300, 200, 0, 0),   # source rows wrap past 511
                                                         (10, 10, 520, 300, 129, 65, 0, 1), (64, 64, 768, 384, 128, 64, 1, 1),
                                                         (900, 20, 512, 40, 250, 100, 0, 0)]:    # source columns wrap past 1023
        out += _rec(18, struct.pack("<6HBB", sx, sy, dx, dy, w, h, set_mask, check_mask))
    stream = bytes(out)
    want = oracle.render(stream)
    with _ctx() as ctx:
        ctx.submit(stream)
        got = ctx.read_vram()
        assert np.array_equal(got, want), describe_diff(got, want)
    for mode in (2, 1):
        with _ctx(16, encode_mode=mode) as ctx:
            ctx.submit_batch([stream] * 16)
            ctx.flush()
            assert [int(h) for h in ctx.hash()] == [oracle.hash64(want)] * 16, mode


@pytest.mark.gpu
def test_dropped_batch_chunk_restores_encoder_state(oracle):
    """A host-encoded batch that fails because one stream is malformed is dropped as a whole chunk, and the encoders go
    back to their state before it (the CLUT snapshot table must not keep slots whose snapshot ops were
    never run).  What follows must render as if the dropped chunk had never been submitted."""
    from duckstation_b200 import PsxGpuError

    n = 4
    parts = []
    for seed in range(n):
        s = _strip_barriers(streams.generate(2, 40 + seed, 1500, 0))
        cuts = [off for _, off, _ in streams.iter_records(s)]
        a, b = cuts[len(cuts) // 3], cuts[2 * len(cuts) // 3]
        parts.append((s[:a], s[a:b], s[b:]))
    want = [oracle.hash64(oracle.render(p[0] + p[2])) for p in parts]
    with _ctx(n, encode_mode=1) as ctx:
        ctx.submit_batch([p[0] for p in parts])
        bad = [p[1] for p in parts]
        bad[2] = bad[2] + b"\x07\x00\x00\x00\x16\x00\x00\x00"  # a record of size 7 at the end: malformed
        with pytest.raises(PsxGpuError):
            ctx.submit_batch(bad)
        ctx.submit_batch([p[2] for p in parts])
        ctx.flush()
        assert [int(h) for h in ctx.hash()] == want


@pytest.mark.gpu
def test_cluster_mode_several_instances(oracle):
    """Up to eight instances whose streams are cut into many short epochs run as one thread-block cluster per instance
    (raster_cluster_kernel): five different hazard-heavy streams in one flush, every VRAM against the oracle — aliasing
    copies (closed form over the cluster), mask-checked uploads and self-sampling primitives included."""
    cases = [(3, 11, 3000, 0), (3, 12, 2500, 0), (0, 77, 2500, 3), (6, 5, 1500, 0), (3, 13, 600, 0)]
    ss = [_strip_barriers(streams.generate(*c)) for c in cases]
    want = [oracle.hash64(oracle.render(s)) for s in ss]
    with _ctx(len(ss)) as ctx:
        for i, s in enumerate(ss):
            ctx.submit(s, i)
        ctx.flush()
        got = [int(h) for h in ctx.hash()]
        assert got == want, [cases[i] for i in range(len(ss)) if got[i] != want[i]]
        st = ctx.stats()
        assert st["epochs"] > 8 * len(ss) and st["serial_copy"] > 0 and st["serial_upload"] > 0


def test_slot_mode_mixed_batch(oracle):
    """More instances than SMs, cut into many short epochs: the batch runs in slot mode (raster_multi_kernel — several
    instances per CTA, each slot behind its own epoch counter).  The mix makes every path of the kernel happen: hazard
    streams (config 6), streams full of serial epoch kinds (config 3 and the fuzz generator: aliasing copies, mask-checked
    uploads, self-sampling primitives — the CTA-wide rendezvous), streams that start with a serial epoch, empty streams and
    single-command streams (slots that finish at once and pull the next instance), ragged lengths (slots that run out at
    different times).  Every instance against the oracle."""
    from tests.test_hostsim import _area, _copy, _rect
    cases = ([(6, 200 + i, 300 + 37 * (i % 9), 0) for i in range(120)] + [(3, 300 + i, 200 + 53 * (i % 7), 0) for i in range(120)] +
             [(0, 400 + i, 300, 3) for i in range(40)])
    ss = [_strip_barriers(streams.generate(*c)) for c in cases]
    noise = _strip_barriers(streams.generate(0, 5, 0, 0))
    ss += [b""] * 12                                                              # nothing to do
    ss += [_area(0, 0, 1023, 511) + _rect(3 * i, 5 * i, 40, 30, 0x40A0F0) for i in range(12)]  # one epoch, one primitive
    ss += [noise + _copy(100, 100, 104, 110, 200, 150) + _rect(7, 9, 50, 50, 0x3050F0)] * 6    # a serial epoch first
    want = [oracle.hash64(oracle.render(s)) for s in ss]
    assert len(ss) > 300
    with _ctx(len(ss)) as ctx:
        ctx.set_batch_chunk(len(ss))  # one chunk: more than two instances per SM, so the launcher picks three slots per CTA
        ctx.submit_batch(ss)
        ctx.flush()
        got = [int(h) for h in ctx.hash()]
        bad = [i for i in range(len(ss)) if got[i] != want[i]]
        assert not bad, bad[:10]
        st = ctx.stats()
        assert st["epochs"] > 8 * len(ss) and st["serial_copy"] > 0 and st["serial_upload"] > 0
        assert st["kernel_launches"] < 40  # one launch for all epochs, not one per epoch


@pytest.mark.parametrize("slots", [1, 2, 4])
def test_slot_mode_forced_slot_counts(slots):
    """raster_multi_kernel with 1, 2 and 4 instance slots per CTA (PSXGPU_SLOTS, read once per process: a worker process per
    value) on a batch of hazard-heavy streams with serial epochs; with the variable set the slot kernel runs whatever the
    instance count, so 40 instances put 1-4 of them on each of a few CTAs.  Worker: tests/slot_worker.py."""
    import subprocess
    import sys
    env = dict(os.environ, PSXGPU_SLOTS=str(slots), PSXGPU_CLUSTER_MAX="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "slot_worker.py"), "40"], env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
