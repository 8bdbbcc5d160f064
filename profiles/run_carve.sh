#!/bin/bash
tag=${1:-cv}
mkdir -p gpurun_out; L=gpurun_out/${tag}_carve.log
for cfg in "32 0" "28 86" "28 0" "29 86" "24 75" "20 60"; do
  set -- $cfg
  for rep in 1 2; do
    echo -n "ctas/SM=$1 carveout=$2: " >> $L
    PSXGPU_LOOP_CTAS=$1 PSXGPU_LOOP_CARVEOUT=$2 timeout 120 python profiles/profile_driver.py 512 3 2>&1 | tail -1 | cut -c1-100 >> $L
  done
done
cat $L
