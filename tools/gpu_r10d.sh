#!/bin/bash
# 2-GPU session on the last commit: every GPU test
TAG=${1:-r10d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -q -m gpu -rs > $OUT/pytest_gpu_2gpu.log 2>&1; echo "rc=$?"; tail -12 $OUT/pytest_gpu_2gpu.log | cut -c1-200
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
