#!/bin/bash
TAG=${1:-r4p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== lz4 phases"; timeout 600 python tools/lz4_phase_probe.py 400 > $OUT/lz4_phases.jsonl 2> $OUT/lz4_phases.err; echo "rc=$?"; grep '"cta"' $OUT/lz4_phases.jsonl | cut -c1-1800; tail -3 $OUT/lz4_phases.err
