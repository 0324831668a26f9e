#!/bin/bash
# 1-GPU session r4d: per-CTA timeline of single launches (where the fixed ~6-13 us per launch go),
# pageable staging with non-temporal stores vs memcpy, fused-exchange tests after the protocol change.
TAG=${1:-r4d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest fused exchange + parity"; timeout 600 python -m pytest tests/test_fused_exchange.py tests/test_cuda_parity.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest.log
echo "== timeline"; timeout 300 python tools/timeline_probe.py > $OUT/timeline.jsonl 2> $OUT/timeline.err; echo "rc=$?"; cut -c1-700 $OUT/timeline.jsonl; tail -3 $OUT/timeline.err
echo "== pageable NT=1"; FLAGSTAT_CUDA_STAGING_NT=1 timeout 300 python tools/pageable_bench.py > $OUT/pageable_nt1.jsonl 2> $OUT/pageable.err; echo "rc=$?"; grep staged $OUT/pageable_nt1.jsonl | cut -c1-120
echo "== pageable NT=0"; FLAGSTAT_CUDA_STAGING_NT=0 timeout 300 python tools/pageable_bench.py > $OUT/pageable_nt0.jsonl 2>> $OUT/pageable.err; echo "rc=$?"; grep staged $OUT/pageable_nt0.jsonl | cut -c1-120
