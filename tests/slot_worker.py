"""Worker of test_slot_mode_forced_slot_counts: the environment variable PSXGPU_SLOTS is read once per process, so each
slot count runs in a process of its own.  Renders a hazard-heavy batch through the C ABI and checks every hash against
the CPU oracle; exits non-zero on a mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from duckstation_b200 import Context, streams  # noqa: E402
from tests.helpers import Oracle  # noqa: E402


def main() -> int:
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    oracle = Oracle()
    cases = [(6, 500 + i, 250 + 41 * (i % 5), 0) for i in range(n // 2)] + [(3, 600 + i, 150 + 29 * (i % 6), 0) for i in range(n - n // 2)]
    ss = [b"".join(s[o:o + z] for t, o, z in streams.iter_records(s) if t not in (7, 9, 12, 15)) for s in (streams.generate(*c) for c in cases)]
    want = [oracle.hash64(oracle.render(s)) for s in ss]
    with Context(len(ss)) as ctx:
        ctx.set_batch_chunk(len(ss))
        ctx.submit_batch(ss)
        ctx.flush()
        got = [int(h) for h in ctx.hash()]
        st = ctx.stats()
    bad = [cases[i] for i in range(len(ss)) if got[i] != want[i]]
    print(f"slots={os.environ.get('PSXGPU_SLOTS')} instances={len(ss)} epochs={st['epochs']} launches={st['kernel_launches']} bad={bad[:5]}")
    return 1 if bad or st["epochs"] < 8 * len(ss) else 0


if __name__ == "__main__":
    sys.exit(main())
