#!/bin/bash
# 1-GPU session on the last commit of the round: every GPU test, smoke(), both bench arms as the driver launches them
TAG=${1:-r11b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -q -m gpu -rs > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log | cut -c1-200
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1 | tee $OUT/smoke.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
r=json.loads(open("$OUT/bench_ref.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","value_serialised","ms_per_step","verified","gpu_launches")}, d["roofline"]["frac"], d["e2e"]["value"], d["clocks"])
print("stream note:", d["stream_e2e"].get("note","")[:60], "| ref", r["value"], r["cpu_baseline"]["cores"])
print("e2e / ref =", d["e2e"]["value"]/r["value"], " value / ref =", d["value"]/r["value"])
PY
