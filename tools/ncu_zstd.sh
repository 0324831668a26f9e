#!/bin/bash
# ncu --set full of one launch each of the Zstd entropy stage (zstd_parse_kernel) and copy stage (zstd_copy_kernel)
# on the run-structured column.
TAG=${1:-ncu_zstd}; NB=${2:-400}; OUT=gpurun_out/$TAG; mkdir -p $OUT
cat > /tmp/ncu_zstd_drv.py <<PY
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import containers
import libflagstats_b200 as fs
from libflagstats_b200 import blockfile
col = containers.runs_column($NB * 512000)
blob = containers.container(col, "zstd", 1)
for _ in range(2):
    f, n = blockfile.flagstat_container(blob, blockfile.ZSTD)
print(n, int(f[9] + f[25]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zstd_ -s 2 -c 2 -f -o $OUT/prof_zstd python /tmp/ncu_zstd_drv.py > $OUT/ncu.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu.log
python tools/ncu_summary_short.py $OUT/prof_zstd.ncu-rep | tee $OUT/ncu_short.txt
