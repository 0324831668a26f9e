#!/bin/bash
# 1-GPU session (the last GPU seconds of the round): the tile-index pass of the LZ4 / Zstd copy phase with its descriptor
# loads batched (eight descriptors per thread in flight) -- container parity on the new library, then decode-kernel times
# new vs previous (tools/bin/libflagstats_cuda_prev.so = the library of sessions r11b - r11d).
TAG=${1:-r11e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 100 python -m pytest tests/test_blockfile.py -q -m gpu -x > $OUT/pytest_blockfile.log 2>&1; echo "rc=$?"; tail -2 $OUT/pytest_blockfile.log | cut -c1-200
for v in new prev; do
  unset LIBFLAGSTATS_CUDA_SO
  [ $v = prev ] && export LIBFLAGSTATS_CUDA_SO=$PWD/tools/bin/libflagstats_cuda_prev.so
  echo "== LZ4 decode kernel times, $v"
  FLAGSTAT_CUDA_DEBUG=1 FLAGSTAT_CUDA_LZ4_BATCH=296 timeout 60 python tools/lz4_bench.py --quick --only-default > $OUT/lz4_bench_$v.jsonl 2> $OUT/decode_times_$v.txt
  grep "block decode: 296 blocks" $OUT/decode_times_$v.txt | awk '{print $6, $8, $10, $11, $12, $13, $14}' | sort | head -8
done
