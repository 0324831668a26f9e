#!/bin/bash
# GPU-box session after the samtools mode went in: smoke, all GPU parity tests, variant A/B
# of the default kernel against the samtools instantiation, bench (both arms), ncu launch
# list and one ncu --set full of the samtools kernel.
# Usage (from the repo root, under gpurun):  bash tools/gpu_round4.sh [tag]
TAG=${1:-r4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/gpu.txt 2>&1
lscpu | head -25 > $OUT/cpu.txt 2>&1; nproc >> $OUT/cpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "rc=$?"; tail -3 $OUT/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -8 $OUT/pytest_gpu.log
echo "== variant ab"; timeout 300 python tools/variant_ab.py ${VARIANTS:-0 s} > $OUT/variant_ab.jsonl 2> $OUT/variant_ab.err; echo "rc=$?"; cut -c1-200 $OUT/variant_ab.jsonl | tail -12
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"; cat $OUT/bench_ref.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full samtools"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:flagstat_kernel_group -s 2 -c 1 -f -o $OUT/prof_samtools python tools/variant_ab.py s > $OUT/ncu_samtools.log 2>&1; echo "rc=$?"; tail -3 $OUT/ncu_samtools.log
fi
ls -la $OUT
