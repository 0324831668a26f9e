#!/bin/bash
# 1-GPU session r4e: timeline v2 (steady state + per-SM view), pageable in-place registration vs
# staging, quick length sweep after the launch-geometry / conditional-PDL change, drop-in loop.
TAG=${1:-r4e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest parity"; timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_fused_exchange.py tests/test_dropin.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest.log
echo "== timeline"; timeout 300 python tools/timeline_probe.py > $OUT/timeline.jsonl 2> $OUT/timeline.err; echo "rc=$?"; tail -3 $OUT/timeline.err
echo "== pageable"; timeout 400 python tools/pageable_bench.py > $OUT/pageable.jsonl 2> $OUT/pageable.err; echo "rc=$?"; cut -c1-130 $OUT/pageable.jsonl; tail -3 $OUT/pageable.err
echo "== length sweep quick"; timeout 600 python tools/length_sweep.py --quick > $OUT/length_sweep_quick.jsonl 2> $OUT/length_sweep.err; echo "rc=$?"; grep -E '"inmemory|"sweep hiseqx"' $OUT/length_sweep_quick.jsonl | cut -c1-200
echo "== dropin --time"; timeout 300 oracle/_ref/dropin_check --time > $OUT/dropin_time.jsonl 2> $OUT/dropin_time.err; echo "rc=$?"; head -4 $OUT/dropin_time.jsonl
