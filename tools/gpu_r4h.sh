#!/bin/bash
# 1-GPU session r4h (was r4g): the dynamically scheduled kernel -- parity, timeline vs the static split, sweep.
TAG=${1:-r4h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest dynamic + parity + exchange"; timeout 900 python -m pytest tests/test_dynamic_kernel.py tests/test_cuda_parity.py tests/test_fused_exchange.py tests/test_samtools_mode.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -5 $OUT/pytest.log
echo "== timeline"; timeout 300 python tools/timeline_probe.py > $OUT/timeline.jsonl 2> $OUT/timeline.err; echo "rc=$?"; tail -2 $OUT/timeline.err
echo "== length sweep"; timeout 900 python tools/length_sweep.py --quick > $OUT/length_sweep_quick.jsonl 2> $OUT/length_sweep.err; echo "rc=$?"; grep -E '"split=|"inmemory' $OUT/length_sweep_quick.jsonl | cut -c1-175; tail -2 $OUT/length_sweep.err
echo "== bench"; timeout 600 python bench.py --no-extras > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?"; cut -c1-1500 $OUT/bench.json; tail -3 $OUT/bench.err
