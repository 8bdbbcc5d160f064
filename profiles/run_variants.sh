#!/bin/bash
# Times the draw wave (profile_driver 512 3) with every library in profiles/variants/ (or the ones named), hash printed
# for comparison.  usage: run_variants.sh [tag] [names...]
tag=${1:-v}; shift
names="$@"; [ -z "$names" ] && names=$(ls profiles/variants/libpsxgpu_*.so | sed -e 's#.*libpsxgpu_##' -e 's#\.so##')
mkdir -p gpurun_out; cp duckstation_b200/libpsxgpu.so /tmp/libpsxgpu_shipped.so
for n in $names; do
  cp profiles/variants/libpsxgpu_$n.so duckstation_b200/libpsxgpu.so
  for rep in 1 2; do
    echo -n "$n: " >> gpurun_out/${tag}_variants.log
    timeout 120 python profiles/profile_driver.py 512 3 2>&1 | tail -1 | cut -c1-120 >> gpurun_out/${tag}_variants.log
  done
  if [ -n "$VARIANT_CFG6" ]; then
    echo -n "$n cfg6: " >> gpurun_out/${tag}_variants.log
    timeout 120 python profiles/profile_driver.py 512 3 6 2>&1 | tail -1 | cut -c1-120 >> gpurun_out/${tag}_variants.log
  fi
done
cp /tmp/libpsxgpu_shipped.so duckstation_b200/libpsxgpu.so
cat gpurun_out/${tag}_variants.log
