#!/bin/bash
tag=${1:-sv}; shift
mkdir -p gpurun_out; L=gpurun_out/${tag}_setupvar.log
cp duckstation_b200/libpsxgpu.so /tmp/libpsxgpu_shipped.so
for v in "$@"; do
  cp profiles/variants/libpsxgpu_$v.so duckstation_b200/libpsxgpu.so
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"setup_kernel" --csv --log-file gpurun_out/${tag}_${v}.csv python profiles/profile_driver.py 4096 2 > gpurun_out/${tag}_ncu.log 2>&1
  echo "== $v: $(tail -1 gpurun_out/${tag}_ncu.log | cut -c1-90)" >> $L
  grep -E "setup_kernel" gpurun_out/${tag}_${v}.csv | awk -F'","' '{print $NF}' | tr '\n' ' ' >> $L; echo >> $L
  timeout 120 python profiles/profile_driver.py 512 2 6 2>&1 | tail -1 | cut -c1-90 >> $L
done
cp /tmp/libpsxgpu_shipped.so duckstation_b200/libpsxgpu.so
cat $L
