#!/bin/bash
# 2-GPU session: is the deferred-collection step slower because of the ORDER of the legs in bench.py (headline leg
# first, alternative later) or because of the collection order itself?  Default and --no-deferred, alternating.
TAG=${1:-r4z}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
port=29740
for rep in 1 2 3; do
  for mode in deferred immediate; do
    port=$((port+1))
    extra="--no-extras --no-strong"; [ $mode = immediate ] && extra="$extra --no-deferred"
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $port bench.py --gpus $NG --steps 20 --warmup 5 $extra > $OUT/bench_${mode}_$rep.json 2> $OUT/bench_${mode}_$rep.err
    echo "bench $mode rep=$rep rc=$?"
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, {k: d.get(k) for k in ("n_gpus", "ms_per_step", "ms_per_step_overlapped_wait_in_launch", "ms_per_step_overlapped_deferred_collection", "ms_per_step_serialised_launches", "verified")})
PY
