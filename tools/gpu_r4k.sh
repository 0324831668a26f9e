#!/bin/bash
# 1-GPU session r4k: the CTA-per-block LZ4 decoder -- parity (incl. malformed blocks), then timings.
TAG=${1:-r4k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest blockfile"; timeout 900 python -m pytest tests/test_blockfile.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest.log
echo "== lz4 bench"; FLAGSTAT_CUDA_DEBUG=1 timeout 900 python tools/lz4_bench.py > $OUT/lz4_bench.jsonl 2> $OUT/lz4_bench.err; echo "rc=$?"; cut -c1-260 $OUT/lz4_bench.jsonl; grep "block decode" $OUT/lz4_bench.err | sort | uniq -c | sort -rn | head -30; grep -v "block decode" $OUT/lz4_bench.err | tail -5
