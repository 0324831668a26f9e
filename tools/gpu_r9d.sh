#!/bin/bash
# 2-GPU check of the bench line exactly as the driver launches it (default flags), both arms
TAG=${1:-r9d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29931 bench.py --gpus $NG --steps 20 --warmup 5 > $OUT/bench_${NG}gpu.json 2> $OUT/bench_${NG}gpu.err; echo "bench rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29932 bench.py --impl reference --gpus $NG --steps 5 --warmup 3 > $OUT/bench_ref_${NG}gpu.json 2> $OUT/bench_ref_${NG}gpu.err; echo "bench reference rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT/bench_${NG}gpu.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "ms_per_step_overlapped_wait_in_launch", "ms_per_step_overlapped_deferred_collection", "ms_per_step_serialised_launches", "ms_per_step_with_nccl_allreduce", "verified", "gpu_launches")})
print(d["exchange"][:200])
s = d.get("strong_2p34") or {}
print("strong", {k: s.get(k) for k in ("ms_per_step", "value", "verified", "deferred_collection", "error")})
print("e2e", d["e2e"]["value"], d["e2e"]["frac_of_pcie_probe"])
r = open("$OUT/bench_ref_${NG}gpu.json").read().strip().splitlines()
print("ref lines", len(r), r[-1][:200] if r else None)
PY
