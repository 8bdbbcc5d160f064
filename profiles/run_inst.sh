for v in 1 16; do
  echo "run $v: $(PSXGPU_INST_RUN=$v timeout 100 python profiles/profile_driver.py 512 3 6 | cut -c1-60)"
done
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PSXGPU_INST_RUN=16 timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
