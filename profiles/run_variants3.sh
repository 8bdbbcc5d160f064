#!/bin/bash
tag=${1:-v3}; shift
mkdir -p gpurun_out; L=gpurun_out/${tag}_variants.log
cp duckstation_b200/libpsxgpu.so /tmp/libpsxgpu_shipped.so
for v in "$@"; do
  cp profiles/variants/libpsxgpu_$v.so duckstation_b200/libpsxgpu.so
  for rep in 1 2; do
    echo -n "$v clean: " >> $L
    timeout 120 python profiles/profile_driver.py 512 3 2>&1 | tail -1 | sed -e 's/last_flush.*last_hash_ms/last_hash_ms/' | cut -c1-150 >> $L
  done
  for rep in 1 2; do
  echo -n "$v hazard n=2048: " >> $L
  timeout 150 python profiles/profile_driver.py 2048 3 6 2>&1 | tail -1 | sed -e 's/last_flush.*last_hash_ms/last_hash_ms/' | cut -c1-150 >> $L
  done
done
cp /tmp/libpsxgpu_shipped.so duckstation_b200/libpsxgpu.so
cat $L
