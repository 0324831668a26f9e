#!/bin/bash
# 1-GPU session r4l: phase profile of the CTA-per-block LZ4 decoder.
TAG=${1:-r4l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== lz4 phases"; timeout 600 python tools/lz4_phase_probe.py 400 > $OUT/lz4_phases.jsonl 2> $OUT/lz4_phases.err; echo "rc=$?"; grep '"cta"' $OUT/lz4_phases.jsonl | cut -c1-1400; tail -3 $OUT/lz4_phases.err
