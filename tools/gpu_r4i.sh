#!/bin/bash
# 1-GPU session r4i: full GPU suite on the product library, the A/B variants (incl. the dynamic kernel)
# on the -DFSB_ALL_VARIANTS build, grid-size sweep for the mid-size rule, full bench line with extras.
TAG=${1:-r4i}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest gpu (product library)"; timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest_gpu.log
echo "== pytest A/B variants build"; LIBFLAGSTATS_CUDA_SO=$PWD/tools/bin/libflagstats_cuda_variants.so timeout 900 python -m pytest tests/test_dynamic_kernel.py tests/test_cuda_parity.py -x -q -m gpu > $OUT/pytest_variants.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_variants.log
echo "== length sweep"; timeout 900 python tools/length_sweep.py --quick > $OUT/length_sweep_quick.jsonl 2> $OUT/length_sweep.err; echo "rc=$?"; grep -E '"ctas_per' $OUT/length_sweep_quick.jsonl | cut -c1-175; tail -2 $OUT/length_sweep.err
echo "== bench"; timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "value_serialised", "ms_per_step", "verified")})
print("inmemory", json.dumps(d.get("inmemory_100m"))[:1200])
print("files", json.dumps(d.get("file_e2e"))[:4000])
print("pageable", d.get("e2e_pageable"))
print("cpu", json.dumps(d.get("cpu_baseline"))[:3500])
PY
