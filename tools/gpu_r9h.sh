#!/bin/bash
# 1-GPU session: LZ4 after the hop change -- parity, end to end, decode kernel of a full wave (batch forced to 296)
TAG=${1:-r9h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest blockfile"; timeout 900 python -m pytest tests/test_blockfile.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest.log
echo "== lz4 end to end"; timeout 600 python tools/lz4_bench.py 2>/dev/null | grep -v "warp" | cut -c1-200 | tee $OUT/lz4_bench.jsonl
echo "== decode kernel, one wave of 296 blocks"; FLAGSTAT_CUDA_LZ4_BATCH=296 FLAGSTAT_CUDA_DEBUG=1 timeout 600 python tools/lz4_bench.py --quick --only-default > /dev/null 2> $OUT/decode_times_296.txt
grep "block decode: 296 blocks" $OUT/decode_times_296.txt | sort -k12 -n | awk '{print $6, $(NF-4), $(NF-3), $(NF-2), $(NF-1), $NF}' | sort | uniq | head -12
echo "== zstd end to end (its copy stage is the same code)"; timeout 600 python tools/zstd_bench.py --only-default 2>/dev/null | cut -c1-200 | tee $OUT/zstd_bench.jsonl
