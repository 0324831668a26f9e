#!/bin/bash
# 1-GPU session r4j: LZ4 group decoder phase profile, raw .bin reader sweep, dropin test fix.
TAG=${1:-r4j}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest dropin + blockfile"; timeout 900 python -m pytest tests/test_dropin.py tests/test_blockfile.py tests/test_ingest.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest.log
echo "== lz4 phases"; timeout 600 python tools/lz4_phase_probe.py 400 > $OUT/lz4_phases.jsonl 2> $OUT/lz4_phases.err; echo "rc=$?"; cut -c1-900 $OUT/lz4_phases.jsonl; tail -2 $OUT/lz4_phases.err
echo "== raw reader sweep"; timeout 600 python tools/raw_reader_sweep.py > $OUT/raw_reader_sweep.jsonl 2> $OUT/raw_reader.err; echo "rc=$?"; cut -c1-200 $OUT/raw_reader_sweep.jsonl; tail -2 $OUT/raw_reader.err
