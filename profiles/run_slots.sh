#!/bin/bash
# Slot mode (raster_multi_kernel): parity test first, then the hazard variant (generator config 6) with 0..6 slots per CTA.
tag=${1:-sl}
mkdir -p gpurun_out
L=gpurun_out/${tag}_slots.log
timeout 200 python -m pytest tests/test_gpu_features.py -x -q -m gpu -k "instance_mode or cluster_mode or dropped or resident" > gpurun_out/${tag}_t1.log 2>&1; echo "instance tests rc=$?" >> $L; tail -3 gpurun_out/${tag}_t1.log >> $L
for n in 512 2048; do
for sl in 0 1 2 3 4 6; do
  echo -n "n=$n slots=$sl: " >> $L
  PSXGPU_SLOTS=$sl timeout 150 python profiles/profile_driver.py $n 3 6 2>&1 | tail -1 | cut -c1-140 >> $L
done; done
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1; echo "all tests rc=$?" >> $L; tail -3 gpurun_out/${tag}_tests.log >> $L
cat $L
