#!/bin/bash
# N-GPU session r4f: multi-rank tests after the exchange protocol change (deferred collection, sticky
# failure), then the new bench line (strong_2p34 leg, extras at N = 1) at N = 1 and at the box's GPU count.
TAG=${1:-r4f}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
echo "== pytest multi-rank ($NG GPUs)"; timeout 900 python -m pytest tests/test_fused_exchange.py tests/test_sharded_nccl.py -x -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_multi.log
echo "== bench N=1"; timeout 900 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "rc=$?"; tail -3 $OUT/bench_1gpu.err
echo "== bench reference arm"; timeout 300 python bench.py --impl reference > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?"
if [ "$NG" -gt 1 ]; then
  echo "== bench N=$NG"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 20 --warmup 5 > $OUT/bench_${NG}gpu.json 2> $OUT/bench_${NG}gpu.err; echo "rc=$?"; tail -3 $OUT/bench_${NG}gpu.err
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*gpu.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, {k: d.get(k) for k in ("n_gpus", "value", "value_serialised", "ms_per_step", "ms_per_step_serialised_launches", "ms_per_step_overlapped_wait_in_launch", "ms_per_step_with_nccl_allreduce", "verified")})
    print("  strong", d.get("strong_2p34"))
    print("  e2e", d.get("e2e"))
    print("  pageable", d.get("e2e_pageable"))
    print("  inmemory", json.dumps(d.get("inmemory_100m"))[:900])
    print("  files", json.dumps(d.get("file_e2e"))[:2500])
    print("  cpu", json.dumps(d.get("cpu_baseline"))[:3000])
PY
