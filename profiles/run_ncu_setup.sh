#!/bin/bash
tag=${1:-su1}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:setup_kernel -s 1 -c 1 -f -o gpurun_out/${tag}_setup python profiles/profile_driver.py 4096 2 > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
ncu -i gpurun_out/${tag}_setup.ncu-rep --page source --csv > gpurun_out/${tag}_setup_source.csv 2>/dev/null
ncu -i gpurun_out/${tag}_setup.ncu-rep --page raw --csv > gpurun_out/${tag}_setup_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_setup.ncu-rep
