#!/bin/bash
TAG=${1:-r4v}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest blockfile"; timeout 900 python -m pytest tests/test_blockfile.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest.log
TOOLS="racecheck" TOOL_TIMEOUT=1200 bash tools/sanitize.sh $TAG
grep -A4 "Race reported" $OUT/racecheck.log | grep -oE "in [a-z0-9_]+\.cuh:[0-9]+" | sort | uniq -c
echo "== lz4 bench"; timeout 900 python tools/lz4_bench.py --quick > $OUT/lz4_bench.jsonl 2> $OUT/lz4_bench.err; grep '"cta"' $OUT/lz4_bench.jsonl | cut -c1-200
