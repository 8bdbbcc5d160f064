#!/bin/bash
tag=${1:-en1}; shift
mkdir -p gpurun_out; L=gpurun_out/${tag}_enc.log
cp duckstation_b200/libpsxgpu.so /tmp/libpsxgpu_shipped.so
for v in "$@"; do
  cp profiles/variants/libpsxgpu_$v.so duckstation_b200/libpsxgpu.so
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"setup_kernel|encode_kernel|walk_kernel" --csv --log-file gpurun_out/${tag}_${v}.csv python profiles/profile_driver.py 4096 2 > gpurun_out/${tag}_ncu.log 2>&1
  echo "== $v: $(tail -1 gpurun_out/${tag}_ncu.log | cut -c1-90)" >> $L
  grep -E "setup_kernel|encode_kernel|walk_kernel" gpurun_out/${tag}_${v}.csv | awk -F'","' '{print $5, $NF}' | cut -c1-100 >> $L
done
cp /tmp/libpsxgpu_shipped.so duckstation_b200/libpsxgpu.so
timeout 300 ncu --set full --clock-control none --import-source on -k regex:setup_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_setup python profiles/profile_driver.py 4096 2 > gpurun_out/${tag}_ncu2.log 2>&1
ncu -i gpurun_out/${tag}_setup.ncu-rep --page source --csv > gpurun_out/${tag}_setup_source.csv 2>/dev/null
rm -f gpurun_out/${tag}_setup.ncu-rep
cat $L
