#!/bin/bash
# First GPU session of the NEXT round (written at the end of round 1, when the GPU budget was
# spent): the measurements that are missing behind the round-1 code.
#   1. ncu --set full of zstd_decode_kernel (one thread per frame): where do the ~100 ms go --
#      long-scoreboard on table look-ups / bit reads / match sources?  
#   2. decode-kernel times of the three containers (FLAGSTAT_CUDA_DEBUG=1 prints them).
#   3. bench with overlapped steps at the box's GPU count (8-GPU overlapped was never measured).
# Usage (from the repo root, under gpurun [--gpus N]):  bash tools/gpu_round_next.sh [tag]
TAG=${1:-r4a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
echo "== zstd file bench (decode kernel times on stderr)"
FLAGSTAT_CUDA_DEBUG=1 timeout 120 python tools/zstd_file_bench.py 1600 1 > $OUT/zstd_file_bench.jsonl 2> $OUT/zstd_decode_kernel_times.txt; echo "rc=$?"
cut -c1-300 $OUT/zstd_file_bench.jsonl; grep -m4 "block decode" $OUT/zstd_decode_kernel_times.txt
echo "== ncu zstd_decode_kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zstd_decode_kernel -c 1 -f -o $OUT/prof_zstd \
    python tools/zstd_file_bench.py 400 1 > $OUT/ncu_zstd.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu_zstd.log
if [ "$NG" -gt 1 ]; then
  echo "== bench $NG GPUs, overlapped steps"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29520 \
      bench.py --gpus $NG --steps 20 --warmup 5 > $OUT/bench_${NG}gpu.json 2> $OUT/bench_${NG}gpu.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open("$OUT/bench_${NG}gpu.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","ms_per_step_serialised_launches","ms_per_step_with_nccl_allreduce","verified","n_gpus")}, d["roofline"]["kernel_ms_slowest_rank"])
PY
fi
ls -la $OUT
