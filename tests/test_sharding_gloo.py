"""The N>1 path on CPU: block partition of instances + all-gather of per-instance hashes (gloo, world_size 2 and 3 with a
ragged block), compared with the single-process oracle."""
import json
import os
import subprocess
import sys

import pytest

from duckstation_b200 import streams
from duckstation_b200.sharding import instance_range, owner_of
from tests.helpers import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_is_a_partition():
    for n in (1, 7, 64, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                a, b = instance_range(r, world, n)
                assert a <= b
                seen += list(range(a, b))
            assert seen == list(range(n))
            for i in (0, n // 2, n - 1):
                a, b = instance_range(owner_of(i, world, n), world, n)
                assert a <= i < b
    with pytest.raises(ValueError):
        instance_range(2, 2, 8)


@pytest.mark.parametrize("world,n_total", [(2, 8), (3, 7)])
def test_gloo_gather_matches_single_process(world, n_total):
    prims = 150
    port = 29500 + (os.getpid() % 2000) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "gloo_worker.py"), str(n_total), str(prims)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("HASHES ")][-1]
    got = json.loads(line[len("HASHES "):])
    orc = Oracle()
    want = []
    for i in range(n_total):
        orc.render(streams.generate(4, i, prims, 0))
        want.append(orc.hash64())
    assert got == want
