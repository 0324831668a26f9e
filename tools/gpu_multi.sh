#!/bin/bash
# Multi-GPU session under `gpurun --gpus N`: fused-exchange + NCCL parity tests, then bench at 1..N GPUs.
# Usage: bash tools/gpu_multi.sh <tag> <ngpus>
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== pytest multi-gpu"; timeout 900 python -m pytest tests/test_fused_exchange.py tests/test_sharded_nccl.py -x -q -m gpu > $OUT/pytest_multi.log 2>&1; echo "rc=$?"; tail -15 $OUT/pytest_multi.log
for g in 1 2 4 8; do
  [ $g -gt $N ] && break
  echo "== bench $g gpu(s)"
  if [ $g -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_$g.json 2> $OUT/bench_$g.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 20 --warmup 5 > $OUT/bench_$g.json 2> $OUT/bench_$g.err
  fi
  echo "rc=$?"; cat $OUT/bench_$g.json | cut -c1-1500; tail -3 $OUT/bench_$g.err
done
