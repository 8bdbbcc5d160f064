#!/bin/bash
tag=${1:-m1}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:raster_multi -s 1 -c 1 -f -o gpurun_out/${tag}_multi python profiles/profile_driver.py 512 2 6 > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
ncu -i gpurun_out/${tag}_multi.ncu-rep --page source --csv > gpurun_out/${tag}_multi_source.csv 2>/dev/null
ncu -i gpurun_out/${tag}_multi.ncu-rep --page raw --csv > gpurun_out/${tag}_multi_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_multi.ncu-rep
ls -la gpurun_out | tail -5
