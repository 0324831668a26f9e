#!/bin/bash
# 1-GPU session: Zstd two-stage decoder -- parity (both variants), throughput, kernel times
TAG=${1:-r6b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest blockfile"; timeout 900 python -m pytest tests/test_blockfile.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -12 $OUT/pytest.log
echo "== zstd bench"; timeout 900 python tools/zstd_bench.py > $OUT/zstd_bench.jsonl 2> $OUT/zstd_bench.err; echo "rc=$?"; cut -c1-260 $OUT/zstd_bench.jsonl; tail -3 $OUT/zstd_bench.err
echo "== kernel times"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:zstd -c 12 --csv --log-file $OUT/zstd_launches.csv python tools/zstd_bench.py --quick > /dev/null 2>&1; grep -o 'zstd_[a-z_]*kernel[^,]*\|"[0-9.]*"$' $OUT/zstd_launches.csv | paste - - | head -14
