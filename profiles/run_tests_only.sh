#!/bin/bash
tag=${1:-t}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/${tag}_tests.log
