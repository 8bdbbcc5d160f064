mkdir -p gpurun_out
for e in "PSXGPU_CLUSTER_WARPS=32" "PSXGPU_CLUSTER_WARPS=16" "PSXGPU_CLUSTER_WARPS=16 PSXGPU_AUTO_FLUSH=0" "PSXGPU_CLUSTER_WARPS=16 PSXGPU_AUTO_FLUSH=1024" "PSXGPU_CLUSTER_WARPS=16 PSXGPU_AUTO_FLUSH=4096"; do
  env $e timeout 120 python profiles/config3_cluster_probe.py 3 10000 5 2>&1 | tail -1
done
for c in "1 20000" "2 20000" "5 60"; do PSXGPU_CLUSTER_WARPS=16 timeout 120 python profiles/config3_cluster_probe.py $c 5 2>&1 | tail -1;  PSXGPU_AUTO_FLUSH=0 timeout 120 python profiles/config3_cluster_probe.py $c 5 2>&1 | tail -1; done
PSXGPU_CLUSTER_WARPS=16 timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
