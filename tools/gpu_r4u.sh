#!/bin/bash
# 1-GPU session r4u: LZ4 fuzz test on all three decoders, compute-sanitizer (memcheck, racecheck, initcheck) over
# the sanitize driver (now with the CTA decoder, malformed blocks, the deferred exchange at world 1).
TAG=${1:-r4u}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest blockfile"; timeout 900 python -m pytest tests/test_blockfile.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest.log
TOOLS="memcheck racecheck initcheck synccheck" TOOL_TIMEOUT=1200 bash tools/sanitize.sh $TAG
