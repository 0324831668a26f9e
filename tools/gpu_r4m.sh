#!/bin/bash
# 1-GPU session: CTA LZ4 decoder iteration -- parity, bench (quick), phase profile.
TAG=${1:-r4m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest blockfile"; timeout 900 python -m pytest tests/test_blockfile.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest.log
echo "== lz4 bench"; FLAGSTAT_CUDA_DEBUG=1 timeout 900 python tools/lz4_bench.py > $OUT/lz4_bench.jsonl 2> $OUT/lz4_bench.err; echo "rc=$?"; grep -v "warp sequence" $OUT/lz4_bench.jsonl | cut -c1-220; grep -v "block decode" $OUT/lz4_bench.err | tail -3
echo "== lz4 phases"; timeout 600 python tools/lz4_phase_probe.py 400 > $OUT/lz4_phases.jsonl 2> $OUT/lz4_phases.err; echo "rc=$?"; grep '"cta"' $OUT/lz4_phases.jsonl | cut -c1-1500; tail -3 $OUT/lz4_phases.err
