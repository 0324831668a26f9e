#!/bin/bash
# Final GPU-box session of a round on the CURRENT default kernel: smoke, all GPU parity tests,
# bench (both arms), ncu launch list, ncu --set full of the counting kernel on the bench workload
# (refreshes profiles/ncu_traffic.json through tools/summarize_ncu.py).
# Usage (from the repo root, under gpurun):  bash tools/gpu_round6.sh [tag]
TAG=${1:-r6}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/gpu.txt 2>&1
lscpu | head -25 > $OUT/cpu.txt 2>&1; nproc >> $OUT/cpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -3 $OUT/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cut -c1-600 $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cut -c1-300 $OUT/bench_ref.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-strong > $OUT/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:flagstat_kernel -s 3 -c 2 -f -o $OUT/prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-strong > $OUT/ncu_full.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_full.log
ls -la $OUT
