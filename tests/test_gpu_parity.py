"""GPU parity: the CUDA rasteriser, driven through the C ABI, against the CPU oracle on the same seeded streams."""
import numpy as np
import pytest

from duckstation_b200 import streams
from tests.helpers import Oracle, describe_diff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


@pytest.fixture()
def ctx():
    from duckstation_b200 import Context

    c = Context(1)
    yield c
    c.close()


def _check(ctx, oracle, stream, **kw):
    want = oracle.render(stream)
    ctx.clear_vram(0)
    ctx.submit(stream, 0)
    got = ctx.read_vram(0)
    assert np.array_equal(got, want), describe_diff(got, want)
    assert ctx.hash(0, 1)[0] == oracle.hash64(want)


@pytest.mark.parametrize("config,seed,n,flags", [
    (1, 1, 20000, 0),
    (2, 1, 20000, 0),
    (3, 1, 10000, 0),
    (4, 0, 2000, 0),
    (4, 4095, 2000, 0),
    (5, 1, 12, 400),
])
def test_baseline_configs(ctx, oracle, config, seed, n, flags):
    _check(ctx, oracle, streams.generate(config, seed, n, flags))


@pytest.mark.parametrize("seed", range(1, 6))
@pytest.mark.parametrize("flags", [0, 1, 2, 3])
def test_fuzz(ctx, oracle, seed, flags):
    _check(ctx, oracle, streams.generate(0, seed, 3000, flags))


def test_fuzz_many_short_streams(oracle):
    """A wider net than test_fuzz: 160 short fuzz streams (interlaced rendering, mask bits, wrapping transfers, all
    primitive kinds) through the latency path, and 64 more as one device-encoded batch through the throughput kernels."""
    from duckstation_b200 import Context

    bad = []
    with Context(1) as ctx:
        for seed in range(100, 140):
            for flags in (0, 1, 2, 3):
                s = streams.generate(0, seed, 500, flags)
                want = oracle.render(s)
                ctx.clear_vram(0)
                ctx.submit(s)
                if not np.array_equal(ctx.read_vram(), want):
                    bad.append((seed, flags))
    assert not bad, bad
    batch = []
    for seed in range(200, 264):
        s = streams.generate(0, seed, 400, seed % 4)
        batch.append(b"".join(s[off:off + size] for typ, off, size in streams.iter_records(s) if typ not in (7, 9, 12, 15)))
    want = [oracle.hash64(oracle.render(s)) for s in batch]
    with Context(len(batch), encode_mode=2) as ctx:
        ctx.submit_batch(batch)
        ctx.flush()
        got = [int(h) for h in ctx.hash()]
    assert got == want, [i for i in range(len(batch)) if got[i] != want[i]]
