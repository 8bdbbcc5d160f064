set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -c 2500 gpurun_out/bench_final2.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref2.json 2>/dev/null; tail -c 800 gpurun_out/bench_ref2.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:walk_kernel --launch-count 1 -f -o gpurun_out/r01_walk_v2 python profiles/encode_probe.py 512 > gpurun_out/ncu_walk.log 2>&1
python profiles/single_vram_timing.py 2>&1 | tail -12
