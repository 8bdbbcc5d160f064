bash profiles/gpu_check.sh s1 tests ncu
ncu -i gpurun_out/s1_raster.ncu-rep --page source --csv > gpurun_out/s1_raster_source.csv 2>/dev/null
ncu -i gpurun_out/s1_raster.ncu-rep --page raw --csv > gpurun_out/s1_raster_raw.csv 2>/dev/null
ls -la gpurun_out
