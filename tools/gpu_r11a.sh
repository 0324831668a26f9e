#!/bin/bash
# 1-GPU session: LZ4 CTA decoder (and the Zstd copy stage that shares its copy phase), pointer-jumping rounds:
# (A) a pair leaves after a round that did not move its parents (product until r10) vs
# (B) a pair leaves as soon as it knows that it sits on roots (-DFSB_L4_EARLY_ROOT=1).
# The (B) library is built beforehand (tools/bin/libflagstats_cuda_earlyroot.so travels with the snapshot):
#   python - <<PY
#   from libflagstats_b200 import build as B; import subprocess, os
#   subprocess.check_call([B.nvcc()] + B.NVCC_FLAGS + ["-DFSB_L4_EARLY_ROOT=1", "-o", "tools/bin/libflagstats_cuda_earlyroot.so", os.path.join(B.CSRC, "flagstat_capi.cu")])
#   PY
TAG=${1:-r11a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
VAR=$PWD/tools/bin/libflagstats_cuda_earlyroot.so
VAR3=$PWD/tools/bin/libflagstats_cuda_earlyroot3.so   # (C) = (B) with three hops per round (-DFSB_L4_HOPS=3)
[ -f $VAR ] && [ -f $VAR3 ] || { echo "no $VAR / $VAR3"; exit 1; }
echo "== container parity tests on the early-root builds"
LIBFLAGSTATS_CUDA_SO=$VAR timeout 300 python -m pytest tests/test_blockfile.py -q -m gpu -x > $OUT/pytest_blockfile_earlyroot.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_blockfile_earlyroot.log | cut -c1-200
LIBFLAGSTATS_CUDA_SO=$VAR3 timeout 300 python -m pytest tests/test_blockfile.py -q -m gpu -x > $OUT/pytest_blockfile_earlyroot3.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_blockfile_earlyroot3.log | cut -c1-200
for v in default earlyroot earlyroot3; do
  unset LIBFLAGSTATS_CUDA_SO
  case $v in earlyroot3*) export LIBFLAGSTATS_CUDA_SO=$VAR3;; earlyroot*) export LIBFLAGSTATS_CUDA_SO=$VAR;; esac
  echo "== LZ4 decode kernel times, $v"
  FLAGSTAT_CUDA_DEBUG=1 FLAGSTAT_CUDA_LZ4_BATCH=296 timeout 120 python tools/lz4_bench.py --quick --only-default > $OUT/lz4_bench_$v.jsonl 2> $OUT/decode_times_$v.txt
  grep "block decode: 296 blocks" $OUT/decode_times_$v.txt | sort -k10 -n | awk '{print $6, $8, $10, $11, $12, $13, $14}' | sort | uniq -c | head -6
  cut -c1-200 $OUT/lz4_bench_$v.jsonl
done
for v in default earlyroot; do
  unset LIBFLAGSTATS_CUDA_SO
  case $v in earlyroot3*) export LIBFLAGSTATS_CUDA_SO=$VAR3;; earlyroot*) export LIBFLAGSTATS_CUDA_SO=$VAR;; esac
  echo "== Zstd containers, $v"
  FLAGSTAT_CUDA_DEBUG=1 timeout 120 python tools/zstd_bench.py --quick --only-default > $OUT/zstd_bench_$v.jsonl 2> $OUT/zstd_times_$v.txt
  cut -c1-200 $OUT/zstd_bench_$v.jsonl
  grep "block decode" $OUT/zstd_times_$v.txt | awk '{print $4, $10, $11}' | sort | uniq -c | head -6
done
