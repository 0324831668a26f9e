#!/bin/bash
# 8-GPU session: fused-exchange tests at world 8, bench at 8 and 4 GPUs.   bash tools/gpu_scale8.sh <tag>
TAG=${1:-scale8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== pytest fused exchange"; timeout 600 python -m pytest tests/test_fused_exchange.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest.log
for g in 8 4; do
  echo "== bench $g"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus $g --steps 20 --warmup 5 > $OUT/bench_$g.json 2> $OUT/bench_$g.err
  echo "rc=$?"; python - <<PY
import json
d=json.loads(open("$OUT/bench_$g.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","ms_per_step_with_nccl_allreduce","verified","n_gpus")}, d["roofline"]["kernel_ms"], d["roofline"]["kernel_ms_slowest_rank"], d["e2e"]["value"], d["stream_e2e"])
PY
  tail -2 $OUT/bench_$g.err
done
