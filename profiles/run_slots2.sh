#!/bin/bash
# slot mode: prefetch variants x slots per CTA on the hazard variant (2048 instances of generator config 6)
tag=${1:-sl2}; shift
mkdir -p gpurun_out; L=gpurun_out/${tag}_slots.log
cp duckstation_b200/libpsxgpu.so /tmp/libpsxgpu_shipped.so
for v in "$@"; do
  cp profiles/variants/libpsxgpu_$v.so duckstation_b200/libpsxgpu.so
  for sl in 2 3 4; do
    echo -n "$v n=2048 slots=$sl: " >> $L
    PSXGPU_SLOTS=$sl timeout 150 python profiles/profile_driver.py 2048 3 6 2>&1 | tail -1 | cut -c1-100 >> $L
  done
done
cp /tmp/libpsxgpu_shipped.so duckstation_b200/libpsxgpu.so
cat $L
