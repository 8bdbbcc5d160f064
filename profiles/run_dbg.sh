#!/bin/bash
# The GPU suite and the slot-mode / cluster probes under the PSX_DEBUG_BOUNDS build (index / extent checks that trap);
# also asks whether compute-sanitizer can be used on this box.
tag=${1:-dbg}
mkdir -p gpurun_out; L=gpurun_out/${tag}.log
cp duckstation_b200/libpsxgpu.so /tmp/libpsxgpu_shipped.so
cp profiles/variants/libpsxgpu_dbg.so duckstation_b200/libpsxgpu.so
echo "# PSX_DEBUG_BOUNDS build on a B200: GPU suite, slot mode on 512 instances of config 6, config 3 (cluster mode)" > $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 >> $L
timeout 200 python profiles/profile_driver.py 512 2 6 2>&1 | tail -1 | cut -c1-200 >> $L
timeout 200 python profiles/config3_cluster_probe.py 2>&1 | tail -2 | cut -c1-200 >> $L
cp /tmp/libpsxgpu_shipped.so duckstation_b200/libpsxgpu.so
echo "# compute-sanitizer memcheck on tests/slot_worker.py (shipped build):" >> $L
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tests/slot_worker.py 24 2>&1 | tail -6 | cut -c1-200 >> $L
cat $L
