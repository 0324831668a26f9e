#!/bin/bash
TAG=${1:-r4t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== lz4 bench (no debug sync)"; timeout 900 python tools/lz4_bench.py > $OUT/lz4_bench.jsonl 2> $OUT/lz4_bench.err; echo "rc=$?"; grep -v "warp sequence" $OUT/lz4_bench.jsonl | cut -c1-220; tail -3 $OUT/lz4_bench.err
