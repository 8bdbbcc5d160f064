#!/bin/bash
tag=${1:-mw}
mkdir -p gpurun_out; L=gpurun_out/${tag}_multiwarps.log
for w in 28 32 24; do for sl in 3 4; do
  echo -n "warps=$w slots=$sl n=2048: " >> $L
  PSXGPU_MULTI_WARPS=$w PSXGPU_SLOTS=$sl timeout 150 python profiles/profile_driver.py 2048 3 6 2>&1 | tail -1 | cut -c1-100 >> $L
done; done
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $L
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> $L
cat $L
