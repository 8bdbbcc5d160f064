#!/bin/bash
tag=${1:-el}
mkdir -p gpurun_out; L=gpurun_out/${tag}_e2e_loop.log
for e in "X=0" "PSXGPU_RASTER_LOOP=0" "PSXGPU_RASTER_LOOP=0 PSXGPU_ENC_PRIO=1" "PSXGPU_ENC_PRIO=1" "PSXGPU_RASTER_LOOP=0 PSXGPU_ENC_PRIO=1 PSXGPU_CHUNK_DIV=8" "PSXGPU_RASTER_LOOP=0 PSXGPU_ENC_PRIO=1 PSXGPU_CHUNK_DIV=16"; do
env $e timeout 300 python bench.py --steps 5 --warmup 3 --headline-only --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.strip().startswith('{'):
        d=json.loads(l); print('$e', 'value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'e2e ms',round(d['e2e']['ms_per_step'],2), 'dev', round(d['e2e']['breakdown_ms_per_step']['device_ms'],2))
" >> $L
done
cat $L
