#!/bin/bash
tag=${1:-c2}
mkdir -p gpurun_out; L=gpurun_out/${tag}_cfg2modes.log
for e in "X=0" "PSXGPU_INSTANCE_MODE=1" "PSXGPU_INSTANCE_MODE=1 PSXGPU_CLUSTER_RH=8" "PSXGPU_INSTANCE_MODE=0"; do
  echo "== $e" >> $L
  env $e timeout 300 python profiles/single_vram_timing.py 2>&1 | cut -c1-200 >> $L
done
cat $L
