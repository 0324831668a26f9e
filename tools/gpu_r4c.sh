#!/bin/bash
# 1-GPU session r4c: what round 2 needs measured on the kernel that ships BEFORE changing it --
# GPU parity suite on the current tree, length sweep (configs[2] / configs[0]), drop-in block loop
# and GPU/CPU crossover, pageable staging sweep, bench line.
TAG=${1:-r4c}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_gpu.log
echo "== dropin --time"; timeout 300 oracle/_ref/dropin_check --time > $OUT/dropin_time.jsonl 2> $OUT/dropin_time.err; echo "rc=$?"; head -4 $OUT/dropin_time.jsonl; tail -3 $OUT/dropin_time.jsonl
echo "== length sweep"; timeout 600 python tools/length_sweep.py > $OUT/length_sweep.jsonl 2> $OUT/length_sweep.err; echo "rc=$?"; grep -E '"inmemory|ctas_per' $OUT/length_sweep.jsonl | cut -c1-230
echo "== pageable"; timeout 300 python tools/pageable_bench.py > $OUT/pageable.jsonl 2> $OUT/pageable.err; echo "rc=$?"; cut -c1-200 $OUT/pageable.jsonl
echo "== bench"; timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cut -c1-600 $OUT/bench.json; tail -3 $OUT/bench.err
