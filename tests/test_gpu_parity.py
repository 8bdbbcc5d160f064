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
