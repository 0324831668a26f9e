#!/bin/bash
# 8-GPU session (round 2): multi-rank parity at world 8, then the bench line at N = 8, 4, 2 (weak configs[1]
# per GPU + the strong 2^34-record leg, deferred collection vs wait-in-launch vs NCCL), as the driver's SCALE runs it.
TAG=${1:-r4_8gpu}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
echo "== pytest multi-rank ($NG GPUs)"; timeout 900 python -m pytest tests/test_fused_exchange.py tests/test_sharded_nccl.py "tests/test_blockfile.py::test_lz4_container_on_a_second_device_after_the_first" -x -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_multi.log
port=29540
for N in $NG; do
  [ "$N" -gt "$NG" ] && continue
  port=$((port+1))
  echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*" $OUT/bench_${N}gpu.err | tail -3
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*gpu.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, {k: d.get(k) for k in ("n_gpus", "value", "value_serialised", "ms_per_step", "ms_per_step_serialised_launches", "ms_per_step_overlapped_wait_in_launch", "ms_per_step_with_nccl_allreduce", "verified")})
    print("  kernel_ms", d["roofline"]["kernel_ms"], d["roofline"]["kernel_ms_slowest_rank"])
    s = d.get("strong_2p34") or {}
    print("  strong", {k: s.get(k) for k in ("ms_per_step", "value", "ms_per_step_serialised_launches", "value_serialised", "gbs_per_gpu", "verified", "efficiency_vs_n1_hint", "error")})
    print("  e2e", {k: d["e2e"].get(k) for k in ("value", "achieved_gbs_per_gpu", "pcie_h2d_probe_gbs", "frac_of_pcie_probe")})
    print("  stream", {k: (d.get("stream_e2e") or {}).get(k) for k in ("value", "gbs_per_gpu", "verified")})
PY
