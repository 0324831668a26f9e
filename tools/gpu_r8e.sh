#!/bin/bash
TAG=${1:-r8e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest blockfile"; timeout 900 python -m pytest tests/test_blockfile.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest.log
echo "== lz4"; timeout 600 python tools/lz4_bench.py --only-default 2>/dev/null | cut -c1-200 | tee $OUT/lz4.jsonl
echo "== zstd"; timeout 600 python tools/zstd_bench.py --only-default 2>/dev/null | cut -c1-200 | tee $OUT/zstd.jsonl
