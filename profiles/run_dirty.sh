#!/bin/bash
tag=${1:-dr}
mkdir -p gpurun_out; L=gpurun_out/${tag}_dirty.log
cp duckstation_b200/libpsxgpu.so /tmp/libpsxgpu_shipped.so
for v in base9 dirty; do
  cp profiles/variants/libpsxgpu_$v.so duckstation_b200/libpsxgpu.so
  for rep in 1 2; do
    echo -n "$v clean: " >> $L; timeout 120 python profiles/profile_driver.py 512 3 2>&1 | tail -1 | cut -c1-90 >> $L
    echo -n "$v hazard 2048: " >> $L; timeout 150 python profiles/profile_driver.py 2048 3 6 2>&1 | tail -1 | cut -c1-90 >> $L
  done
  echo -n "$v config3: " >> $L; timeout 200 python profiles/config3_cluster_probe.py 2>&1 | tail -1 | cut -c1-120 >> $L
done
cp profiles/variants/libpsxgpu_dirty.so duckstation_b200/libpsxgpu.so
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> $L
cp /tmp/libpsxgpu_shipped.so duckstation_b200/libpsxgpu.so
cat $L
