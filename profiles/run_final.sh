#!/bin/bash
# Final evidence run of a session: GPU suite, bench line, ncu launch list of the bench command, ncu --set full of the
# draw wave (raster_loop_kernel, 512 instances) and of the slot-mode kernel (512 instances of the hazard variant).
tag=${1:-f}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -2 gpurun_out/${tag}_tests.log
timeout 60 python profiles/profile_driver.py 512 3 > gpurun_out/${tag}_drv.log 2>&1; tail -1 gpurun_out/${tag}_drv.log | cut -c1-200
timeout 500 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${tag}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --headline-only > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:raster_loop -s 3 -c 1 -f -o gpurun_out/${tag}_raster python profiles/profile_driver.py 512 2 > gpurun_out/${tag}_ncu1.log 2>&1; tail -1 gpurun_out/${tag}_ncu1.log
ncu -i gpurun_out/${tag}_raster.ncu-rep --page raw --csv > gpurun_out/${tag}_raster_raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_raster.ncu-rep --page source --csv > gpurun_out/${tag}_raster_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:raster_multi -s 1 -c 1 -f -o gpurun_out/${tag}_multi python profiles/profile_driver.py 512 2 6 > gpurun_out/${tag}_ncu2.log 2>&1; tail -1 gpurun_out/${tag}_ncu2.log
ncu -i gpurun_out/${tag}_multi.ncu-rep --page raw --csv > gpurun_out/${tag}_multi_raw.csv 2>/dev/null
rm -f gpurun_out/${tag}_multi.ncu-rep
ls -la gpurun_out | grep ${tag}_
