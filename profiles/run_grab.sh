#!/bin/bash
tag=${1:-gr}
mkdir -p gpurun_out; L=gpurun_out/${tag}_grab.log
cp duckstation_b200/libpsxgpu.so /tmp/libpsxgpu_shipped.so
run() { echo -n "$1 light=$2 heavy=$3: " >> $L; PSXGPU_GRAB_LIGHT=$2 PSXGPU_GRAB_HEAVY=$3 timeout 120 python profiles/profile_driver.py 4096 3 2>&1 | tail -1 | sed -e 's/last_flush.*last_hash_ms/last_hash_ms/' | cut -c1-125 >> $L; }
for v in base7 nopeek; do
  cp profiles/variants/libpsxgpu_$v.so duckstation_b200/libpsxgpu.so
  run $v 8 2; run $v 16 2; run $v 32 2; run $v 8 1; run $v 8 3; run $v 64 2
done
cp profiles/variants/libpsxgpu_hash2.so duckstation_b200/libpsxgpu.so 2>/dev/null && run hash2 8 2
cp /tmp/libpsxgpu_shipped.so duckstation_b200/libpsxgpu.so
cat $L
