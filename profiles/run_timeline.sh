#!/bin/bash
tag=${1:-tl}
mkdir -p gpurun_out
for d in 8 12 16; do
  echo "=== PSXGPU_CHUNK_DIV=$d" >> gpurun_out/${tag}_timeline.log
  PSXGPU_CHUNK_DIV=$d timeout 300 python profiles/e2e_timeline.py 4096 2>&1 | tail -30 >> gpurun_out/${tag}_timeline.log
done
cat gpurun_out/${tag}_timeline.log
