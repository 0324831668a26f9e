#!/bin/bash
# GPU-box session for the LZ4 group decoder + pageable path: targeted tests first, then the
# whole GPU suite, sanitizer, file bench (A/B of both decoders), bench.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round2.sh [tag]
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/gpu.txt 2>&1
lscpu | head -25 > $OUT/cpu.txt 2>&1; nproc >> $OUT/cpu.txt
echo "== pytest new"; timeout 600 python -m pytest tests/test_blockfile.py "tests/test_cuda_parity.py::test_pageable_host_arrays_take_the_threaded_staging_path" -q -m gpu > $OUT/pytest_new.log 2>&1; echo "rc=$?"; tail -25 $OUT/pytest_new.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest_gpu.log
if [ "${SKIP_SAN:-0}" != "1" ]; then
for tool in memcheck racecheck; do
  echo "== sanitizer $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 python tools/sanitize_driver.py > $OUT/$tool.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize driver ok|Error|error" $OUT/$tool.log | head -8
done
fi
echo "== file bench"; FLAGSTAT_CUDA_DEBUG=1 timeout 900 python tools/file_bench.py ${FILE_BLOCKS:-1600} > $OUT/file_bench.jsonl 2> $OUT/file_bench.err; echo "rc=$?"; cat $OUT/file_bench.jsonl; grep "lz4 decode" $OUT/file_bench.err | sort | uniq -c | sort -rn | head -12
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
ls -la $OUT
