#!/bin/bash
# 1-GPU session: CTA LZ4 decoder with sequence-per-thread B1 -- parity (all decoders), throughput, phase shares
TAG=${1:-r5a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest blockfile"; timeout 900 python -m pytest tests/test_blockfile.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest.log
echo "== lz4 bench"; timeout 900 python tools/lz4_bench.py > $OUT/lz4_bench.jsonl 2> $OUT/lz4_bench.err; echo "rc=$?"; grep -v "warp sequence\|warp group" $OUT/lz4_bench.jsonl | cut -c1-260; tail -3 $OUT/lz4_bench.err
echo "== phases"; timeout 600 python tools/lz4_phase_probe.py 400 > $OUT/lz4_phases.jsonl 2> $OUT/lz4_phases.err; echo "rc=$?"; grep '"cta"' $OUT/lz4_phases.jsonl | cut -c1-900; tail -3 $OUT/lz4_phases.err
