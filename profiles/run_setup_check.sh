#!/bin/bash
tag=${1:-su2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -2 gpurun_out/${tag}_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"setup_kernel|encode_kernel|hash_kernel" --csv --log-file gpurun_out/${tag}_launches.csv python profiles/profile_driver.py 4096 2 > gpurun_out/${tag}_ncu.log 2>&1
grep -E "setup_kernel|encode_kernel|hash_kernel" gpurun_out/${tag}_launches.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | cut -c1-120
timeout 500 python bench.py --steps 5 --warmup 3 --headline-only > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/${tag}_bench.json
