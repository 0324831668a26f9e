#!/bin/bash
# 1-GPU session: LZ4 / Zstd containers after the host-side changes (3 lanes, gather with all CPUs + non-temporal stores)
TAG=${1:-r5g}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest blockfile"; timeout 900 python -m pytest tests/test_blockfile.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest.log
echo "== lz4 bench"; timeout 900 python tools/lz4_bench.py > $OUT/lz4_bench.jsonl 2> $OUT/lz4_bench.err; echo "rc=$?"; grep -v "warp sequence\|warp group" $OUT/lz4_bench.jsonl | cut -c1-260; tail -3 $OUT/lz4_bench.err
echo "== memcpy gather instead of NT"; FLAGSTAT_CUDA_STAGING_NT=0 timeout 900 python tools/lz4_bench.py --quick 2>/dev/null | grep '"cta"' | cut -c1-200
