#!/bin/bash
# ncu --set full of one launch of the CTA-per-block LZ4 decoder on the i.i.d. column (ratio ~2.2).
#   bash tools/ncu_lz4.sh <tag> [blocks = 296]
TAG=${1:-ncu_lz4}; NB=${2:-296}; OUT=gpurun_out/$TAG; mkdir -p $OUT
cat > /tmp/ncu_lz4_drv.py <<PY
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import containers
import libflagstats_b200 as fs
from libflagstats_b200 import blockfile
col = containers.iid_column($NB * 512000)
blob = containers.container(col, "lz4")
for _ in range(3):
    f, n = blockfile.flagstat_container(blob, blockfile.LZ4)
print(n, int(f[9] + f[25]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_decode_cta -s 1 -c 1 -f -o $OUT/prof_lz4 python /tmp/ncu_lz4_drv.py > $OUT/ncu.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu.log
ncu -i $OUT/prof_lz4.ncu-rep --page raw --csv > $OUT/prof_lz4_raw.csv 2>/dev/null
python tools/ncu_summary_short.py $OUT/prof_lz4.ncu-rep | tee $OUT/ncu_short.txt
ncu -i $OUT/prof_lz4.ncu-rep --page source --csv > $OUT/prof_lz4_source.csv 2>/dev/null; wc -l $OUT/prof_lz4_source.csv
