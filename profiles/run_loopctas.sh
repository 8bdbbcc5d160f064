#!/bin/bash
tag=${1:-lc}; shift
mkdir -p gpurun_out; L=gpurun_out/${tag}_loopctas.log
for d in "$@"; do
  PSXGPU_LOOP_CTAS=$d timeout 300 python bench.py --steps 5 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/${tag}_b.json 2> gpurun_out/${tag}_b.err
  python - <<PY >> $L
import json
d=json.load(open('gpurun_out/${tag}_b.json'))
print('ctas/SM=$d value %.1f ms %.2f | e2e %.1f ms %.2f parts %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], {k: round(v,2) for k,v in d['e2e']['breakdown_ms_per_step'].items()}))
PY
done
cat $L
