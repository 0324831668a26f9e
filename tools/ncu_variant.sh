#!/bin/bash
# ncu --set full capture of one kernel variant:  bash tools/ncu_variant.sh <variant> <tag> [uniform]
V=${1:-0}; TAG=${2:-v$V}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export FLAGSTAT_CUDA_VARIANT=$V
cat > /tmp/ncu_drv.py <<PY
import sys, torch
sys.path.insert(0, "/root/repo")
import libflagstats_b200 as fs
from libflagstats_b200 import synth
mode = "${3:-hiseqx}"
d = synth.hiseqx_device(synth.HISEQX_N) if mode == "hiseqx" else synth.uniform_device(synth.HISEQX_N, 0, 0, 0x0FFF)
out = torch.zeros(32, dtype=torch.int64, device="cuda")
for _ in range(6):
    fs.flagstat_device(d, out=out)
torch.cuda.synchronize()
print(out.cpu().tolist()[:16])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flagstat_kernel -s 4 -c 1 -f -o $OUT/prof python /tmp/ncu_drv.py > $OUT/ncu.log 2>&1; echo "rc=$?"; tail -2 $OUT/ncu.log
