#!/bin/bash
# 1-GPU session: batch size of the container pipeline (FLAGSTAT_CUDA_LZ4_BATCH; default one wave = 2 x SMs blocks for the
# LZ4 CTA decoder, 2048 frames for Zstd) -- LZ4 and Zstd containers of 401 and 1601 blocks
TAG=${1:-r8d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for b in default 74 148 222 296 592; do
  [ $b = default ] && unset FLAGSTAT_CUDA_LZ4_BATCH || export FLAGSTAT_CUDA_LZ4_BATCH=$b
  echo "== lz4 batch=$b"; timeout 600 python tools/lz4_bench.py --only-default 2>/dev/null | grep '"cta"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', d['column'], d['blocks'], d['container_call_ms'], 'ms', d['gbs_records'], 'GB/s', d['same_counters'])" | tee -a $OUT/lz4_batch_$b.txt
done
for b in default 300 450 600 900 1200; do
  [ $b = default ] && unset FLAGSTAT_CUDA_LZ4_BATCH || export FLAGSTAT_CUDA_LZ4_BATCH=$b
  echo "== zstd batch=$b"; timeout 600 python tools/zstd_bench.py --only-default 2>/dev/null | grep '"two-stage"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  ', d['column'], d['frames'], d['container_call_ms'], 'ms', d['gbs_records'], 'GB/s', d['same_counters'])" | tee -a $OUT/zstd_batch_$b.txt
done
