#!/bin/bash
# 2-GPU session: EVERY GPU test on a 2-GPU box (-rs: what is still skipped and why), then the A/B build
# (-DFSB_ALL_VARIANTS: all nine kernel variants, the dynamically scheduled kernel) through the variant tests.
TAG=${1:-r8c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu, product library, $(nvidia-smi -L | wc -l) GPUs"; timeout 1500 python -m pytest tests -q -m gpu -rs > $OUT/pytest_gpu_2gpu.log 2>&1; echo "rc=$?"; tail -25 $OUT/pytest_gpu_2gpu.log | cut -c1-220
echo "== A/B build"; LIBFLAGSTATS_CUDA_SO=$PWD/tools/bin/libflagstats_cuda_variants.so timeout 900 python -m pytest tests/test_dynamic_kernel.py tests/test_cuda_parity.py tests/test_samtools_mode.py -q -m gpu -rs > $OUT/pytest_ab_build.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest_ab_build.log | cut -c1-220
