"""World-size-N worker for tests/test_sharding_gloo.py: each rank renders its block of instances with the CPU host
simulation (no GPU here), all-gathers the per-instance hashes over gloo and rank 0 prints the merged list."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from duckstation_b200 import streams  # noqa: E402
from duckstation_b200.context import vramhash64  # noqa: E402
from duckstation_b200.sharding import gather_hashes, instance_range  # noqa: E402
from tests.helpers import HostSim  # noqa: E402


def main():
    n_total, prims = int(sys.argv[1]), int(sys.argv[2])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    first, last = instance_range(rank, world, n_total)
    sim = HostSim()
    local = [vramhash64(sim.render(streams.generate(4, i, prims, 0))) for i in range(first, last)]
    t = torch.tensor([h - (1 << 64) if h >= (1 << 63) else h for h in local], dtype=torch.int64)
    merged = gather_hashes(t, world, n_total, dist)
    if rank == 0:
        print("HASHES " + json.dumps([int(x) & 0xFFFFFFFFFFFFFFFF for x in merged.tolist()]), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
