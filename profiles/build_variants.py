"""Experiment helper: builds variants of libpsxgpu.so (extra -D flags) into profiles/variants/ (git-ignored), in parallel.
usage: build_variants.py name1="-DPSX_X=1 -DPSX_Y=2" name2="..."   (an empty flag string = the shipped build)"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "duckstation_b200", "csrc")
OUT = os.path.join(ROOT, "profiles", "variants")
os.makedirs(OUT, exist_ok=True)
OBJ = os.path.join(OUT, "obj")
os.makedirs(OBJ, exist_ok=True)
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-Wall"]
REST = ["presenter.cu", "context.cpp", "encoder.cpp", "backend.cpp", "gp0_frontend.cpp"]


def obj(src, extra, tag):
    o = os.path.join(OBJ, f"{os.path.splitext(src)[0]}_{tag}.o")
    r = subprocess.run(["nvcc", *FLAGS, *extra, "-Xptxas", "-v", "-c", "-o", o, os.path.join(CSRC, src)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise SystemExit(f"nvcc failed on {src} ({tag})")
    return o, r.stdout


def main():
    variants = dict(a.split("=", 1) for a in sys.argv[1:])
    with ThreadPoolExecutor(max_workers=8) as ex:
        rest = [ex.submit(obj, s, [], "base") for s in REST]
        kern = {n: ex.submit(obj, "kernels.cu", f.split(), n) for n, f in variants.items()}
        enc = {n: ex.submit(obj, "encode_kernel.cu", f.split(), n) for n, f in variants.items()}
        rest = [f.result()[0] for f in rest]
        for n, f in kern.items():
            o, log = f.result()
            eo, elog = enc[n].result()
            for i, ln in enumerate(elog.splitlines()):
                if "13encode_kernelE" in ln and "Compiling" in ln:
                    print(n, "encode_kernel:", " ".join(x.strip() for x in elog.splitlines()[i + 1:i + 4] if "registers" in x or "spill" in x))
            lines = log.splitlines()
            for i, ln in enumerate(lines):
                if "raster_loop_kernel" in ln and "Compiling" in ln:
                    print(n, "raster_loop_kernel:", " ".join(x.strip() for x in lines[i + 1:i + 4] if "registers" in x or "spill" in x))
            so = os.path.join(OUT, f"libpsxgpu_{n}.so")
            subprocess.check_call(["nvcc", "-shared", "-o", so, o, eo, *rest])
            print("built", so)


main()
