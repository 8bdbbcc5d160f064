"""duckstation_b200 — a B200-native (sm_100a CUDA) PSX GPU rasteriser behind DuckStation's GPUBackend interface.

The product is ``libpsxgpu.so`` (CUDA kernels + host encoder + C ABI, ``include/psxgpu.h``).  This package is
the thin Python host layer used by the tests and the benchmark:

* :mod:`duckstation_b200.context`  — ``Context``: ctypes wrapper of the C ABI (no CPU fallback)
* :mod:`duckstation_b200.frontend` — ``Gp0Frontend`` / ``replay_dump``: GP0/GP1 port words and `.psxgpu` traces -> records
* :mod:`duckstation_b200.streams`  — seeded synthetic command streams for the BASELINE.json configs
* :mod:`duckstation_b200.sharding` — instance partition across ranks + NCCL/gloo gather of per-instance hashes
* :mod:`duckstation_b200.build`    — in-tree build of every native library
"""
from .context import Context, PsxGpuError  # noqa: F401

__all__ = ["Context", "PsxGpuError"]
