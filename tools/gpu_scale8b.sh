#!/bin/bash
# 8-GPU session: multi-rank parity tests at world 8, bench at 8 GPUs on configs[1] per GPU (weak
# scaling, the driver's SCALE shape) and on configs[3] (2^34 records = 2^31 per GPU), then 4 GPUs.
#   bash tools/gpu_scale8b.sh <tag>
TAG=${1:-scale8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== pytest multi-rank"; timeout 600 python -m pytest tests/test_fused_exchange.py tests/test_sharded_nccl.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest.log
run() {  # name gpus extra-args
  echo "== bench $1"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2951$2 bench.py --gpus $2 --steps 20 --warmup 5 $3 > $OUT/bench_$1.json 2> $OUT/bench_$1.err
  echo "rc=$?"; python - <<PY
import json
d=json.loads(open("$OUT/bench_$1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","ms_per_step_with_nccl_allreduce","verified","n_gpus")}, d["config"]["global_records"], d["roofline"]["kernel_ms"], d["roofline"]["kernel_ms_slowest_rank"], d["e2e"]["value"], d["stream_e2e"].get("value"))
PY
  tail -2 $OUT/bench_$1.err
}
run 8gpu 8 ""
run 8gpu_16Grecords 8 "--records 2147483648"
run 4gpu 4 ""
