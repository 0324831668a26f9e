#!/bin/bash
# 8-GPU session at the end of round 2: multi-rank parity at world 8 (packed-word exchange, deferred collection, the LZ4
# container on a second device), then the bench lines the driver's SCALE run produces -- N = 1, 2, 4, 8 on the SAME box
# (weak, configs[1] per GPU; the strong 2^34-record leg at N = 8 and N = 1).
TAG=${1:-r7_8gpu}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
echo "== pytest multi-rank ($NG GPUs)"; timeout 900 python -m pytest tests/test_fused_exchange.py tests/test_sharded_nccl.py "tests/test_blockfile.py::test_lz4_container_on_a_second_device_after_the_first" -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_multi.log
port=29840
for N in 1 2 4 8; do
  [ "$N" -gt "$NG" ] && continue
  port=$((port+1))
  extra="--no-extras --no-strong"; [ "$N" = "$NG" ] && extra="--no-extras"; [ "$N" = 1 ] && extra="--no-extras"
  if [ $N = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 $extra > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 $extra > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err
  fi
  echo "bench N=$N rc=$?"; grep -v "OMP_NUM_THREADS\|^\*\*\*" $OUT/bench_${N}gpu.err | tail -2
done
python - <<PY
import json, glob
v1 = None
for f in sorted(glob.glob("$OUT/bench_*gpu.json"), key=lambda s: int(s.split("bench_")[1].split("gpu")[0])):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    if d["n_gpus"] == 1: v1 = d["value"]
    eff = d["value"] / (d["n_gpus"] * v1) if v1 else None
    print(f, {k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "ms_per_step_overlapped_wait_in_launch", "ms_per_step_serialised_launches", "ms_per_step_with_nccl_allreduce", "verified")}, "efficiency vs N=1", eff)
    print("  kernel_ms", d["roofline"]["kernel_ms"], d["roofline"]["kernel_ms_slowest_rank"])
    s = d.get("strong_2p34") or {}
    if s: print("  strong", {k: s.get(k) for k in ("ms_per_step", "value", "ms_per_step_serialised_launches", "gbs_per_gpu", "verified", "efficiency_vs_n1_hint", "error")})
    print("  e2e", {k: d["e2e"].get(k) for k in ("value", "achieved_gbs_per_gpu", "pcie_h2d_probe_gbs", "frac_of_pcie_probe")})
PY
