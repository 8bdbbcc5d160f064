"""Builds every native library of the repository, in-tree.

* ``duckstation_b200/libpsxgpu.so``  — the product: CUDA kernels + host encoder + C ABI (nvcc, sm_100a only)
* ``duckstation_b200/libpsxgen.so``  — synthetic workload generator (g++)
* ``oracle/libpsxoracle.so``         — CPU restatement used as the parity oracle (gcc; test infrastructure)
* ``oracle/_ref/libpsxref.so``       — the reference's own rasteriser TU compiled in place (only where
                                        ``/root/reference`` exists; test infrastructure)
* ``oracle/_ref/libpsxdec.so``       — the reference's own GP0/GP1 decoder, CRTC and GPU::DoState (gpu.cpp) compiled in
                                        place with link stubs (same condition; test infrastructure)
* ``tests/hostsim/libhostsim.so``    — host build of the device headers for CPU-only tests

nvcc cross-compiles without a GPU, so this runs on the CPU-only build box too.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "duckstation_b200", "csrc")
REFERENCE = "/root/reference"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall", "-shared",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd: list[str], cwd: str | None = None) -> None:
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError(f"build step failed ({r.returncode}): {' '.join(cmd)}")


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the product library cannot be built (there is no CPU fallback)")


def build_product(force: bool = False) -> str:
    out = os.path.join(ROOT, "duckstation_b200", "libpsxgpu.so")
    srcs = [os.path.join(CSRC, f) for f in ("kernels.cu", "encode_kernel.cu", "presenter.cu", "context.cpp", "encoder.cpp", "backend.cpp", "gp0_frontend.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(ROOT, "include", "psxgpu.h"), os.path.join(ROOT, "include", "psxgpu_wire.h")]
    if force or _newer(out, deps):
        _run([_nvcc(), *NVCC_FLAGS, *os.environ.get("PSX_NVCC_EXTRA", "").split(), "-o", out, *srcs])
    return out


def build_generator(force: bool = False) -> str:
    out = os.path.join(ROOT, "duckstation_b200", "libpsxgen.so")
    src = os.path.join(CSRC, "streamgen.cpp")
    if force or _newer(out, [src, os.path.join(ROOT, "include", "psxgpu_wire.h")]):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", out, src])
    return out


def build_oracle(force: bool = False) -> str:
    out = os.path.join(ROOT, "oracle", "libpsxoracle.so")
    src = os.path.join(ROOT, "oracle", "psx_oracle.c")
    if force or _newer(out, [src, os.path.join(ROOT, "include", "psxgpu_wire.h")]):
        _run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fwrapv", "-Wall", "-o", out, src])
    return out


def build_reference(force: bool = False) -> str | None:
    """oracle/_ref: only possible where the reference checkout is mounted (never on the GPU box)."""
    out = os.path.join(ROOT, "oracle", "_ref", "libpsxref.so")
    if not os.path.isdir(os.path.join(REFERENCE, "src", "core")):
        return out if os.path.exists(out) else None
    if force:
        _run(["make", "clean"], cwd=os.path.join(ROOT, "oracle", "ref_build"))
    _run(["make"], cwd=os.path.join(ROOT, "oracle", "ref_build"))
    return out


def build_decoder_reference(force: bool = False) -> str | None:
    """oracle/_ref/libpsxdec.so: the reference's own GP0/GP1 decoder + CRTC + save-state code (src/core/gpu.cpp) compiled
    in place with link stubs (oracle/ref_build/dec_harness.cpp).  Only where the reference checkout is mounted."""
    out = os.path.join(ROOT, "oracle", "_ref", "libpsxdec.so")
    if not os.path.isfile(os.path.join(REFERENCE, "src", "core", "gpu.cpp")):
        return out if os.path.exists(out) else None
    if force:
        _run(["make", "-f", "Makefile.decoder", "clean"], cwd=os.path.join(ROOT, "oracle", "ref_build"))
    _run(["make", "-f", "Makefile.decoder"], cwd=os.path.join(ROOT, "oracle", "ref_build"))
    return out


def build_hostsim(force: bool = False) -> str:
    out = os.path.join(ROOT, "tests", "hostsim", "libhostsim.so")
    srcs = [os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp"), os.path.join(CSRC, "encoder.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    if force or _newer(out, deps):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", out, *srcs])
    return out


def build_all(force: bool = False) -> dict[str, str | None]:
    """Builds whatever is out of date.  Serialised across processes (torchrun ranks) with a file lock."""
    import fcntl

    with open(os.path.join(ROOT, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return {
                "product": build_product(force),
                "generator": build_generator(force),
                "oracle": build_oracle(force),
                "reference": build_reference(force),
                "decoder_reference": build_decoder_reference(force),
                "hostsim": build_hostsim(force),
            }
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    for k, v in build_all("--force" in sys.argv).items():
        print(f"{k}: {v}")
