#!/bin/bash
# 2-GPU session (round 2): packed-word exchange -- multi-rank parity, the second-device LZ4 test, then the weak step at
# N = 1 and N = 2 on the SAME box (several repeats; deferred vs wait-in-launch vs serialised vs NCCL in every line).
TAG=${1:-r4x}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
echo "== pytest multi-rank ($NG GPUs)"; timeout 900 python -m pytest tests/test_fused_exchange.py tests/test_sharded_nccl.py "tests/test_blockfile.py::test_lz4_container_on_a_second_device_after_the_first" -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_multi.log
port=29640
for rep in 1 2 3; do
  for N in 1 $NG; do
    port=$((port+1))
    extra="--no-extras --no-strong"; [ $rep = 3 ] && extra="--no-extras"
    if [ $N = 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 $extra > $OUT/bench_${N}gpu_$rep.json 2> $OUT/bench_${N}gpu_$rep.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 $extra > $OUT/bench_${N}gpu_$rep.json 2> $OUT/bench_${N}gpu_$rep.err
    fi
    echo "bench N=$N rep=$rep rc=$?"
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*gpu_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, {k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "ms_per_step_serialised_launches", "ms_per_step_overlapped_wait_in_launch", "ms_per_step_with_nccl_allreduce", "verified")}, "kernel_ms", d["roofline"]["kernel_ms"])
    s = d.get("strong_2p34") or {}
    if s: print("  strong", {k: s.get(k) for k in ("ms_per_step", "value", "ms_per_step_serialised_launches", "gbs_per_gpu", "verified", "error")})
PY
