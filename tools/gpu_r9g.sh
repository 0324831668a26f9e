#!/bin/bash
# 1-GPU session: compute-sanitizer with the Zstd decoders in the driver; A/B of 2 vs 3 hops per pointer-jumping round
TAG=${1:-r8a}; OUT=gpurun_out/$TAG; mkdir -p $OUT

python - <<PY
import os, subprocess, sys
sys.path.insert(0, os.getcwd())
from libflagstats_b200 import build as B
so = os.path.join("tools", "bin", "libflagstats_cuda_hopbytes.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
subprocess.check_call([B.nvcc()] + B.NVCC_FLAGS + ["-DFSB_L4_HOP_BYTES", "-o", so, os.path.join(B.CSRC, "flagstat_capi.cu")])
PY
for v in default hopbytes; do
  [ $v = hopbytes ] && export LIBFLAGSTATS_CUDA_SO=$PWD/tools/bin/libflagstats_cuda_hopbytes.so
  echo "== decode kernel times, $v"; FLAGSTAT_CUDA_DEBUG=1 timeout 600 python tools/lz4_bench.py --quick > $OUT/lz4_bench_$v.jsonl 2> $OUT/decode_times_$v.txt
  grep "block decode: 296 blocks" $OUT/decode_times_$v.txt | sort | uniq -c | sort -k7 | head -8
done
