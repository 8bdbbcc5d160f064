for v in 162 321; do
  echo "variant $v: $(PSXGPU_INST_VARIANT=$v timeout 100 python profiles/profile_driver.py 512 3 6 | cut -c1-60)"
  echo "variant $v 2048: $(PSXGPU_INST_VARIANT=$v timeout 100 python profiles/profile_driver.py 2048 2 6 | cut -c1-60)"
done
echo "variant 0 2048: $(PSXGPU_INST_VARIANT=0 timeout 100 python profiles/profile_driver.py 2048 2 6 | cut -c1-60)"
