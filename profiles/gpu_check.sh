#!/bin/bash
# One gpurun call: GPU parity suite, draw-wave timing (profile_driver), bench line, optional ncu capture of the draw wave.
# usage: profiles/gpu_check.sh <tag> [tests|notests] [ncu|noncu]
tag=${1:-x}; tests=${2:-tests}; ncu=${3:-noncu}
mkdir -p gpurun_out
if [ "$tests" = tests ]; then timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log; fi
timeout 60 python profiles/profile_driver.py 512 3 > gpurun_out/${tag}_drv.log 2>&1; cat gpurun_out/${tag}_drv.log | tail -3
timeout 420 python bench.py --steps 3 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
if [ "$ncu" = ncu ]; then
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:raster_loop -s 3 -c 1 -f -o gpurun_out/${tag}_raster python profiles/profile_driver.py 512 2 > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
fi
