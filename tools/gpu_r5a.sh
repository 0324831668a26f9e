#!/bin/bash
# 1-GPU session: CTA LZ4 decoder with sequence-per-thread B1 -- parity (all decoders), throughput, phase shares
TAG=${1:-r5a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest blockfile"; timeout 900 python -m pytest tests/test_blockfile.py -x -q -m gpu > $OUT/pytest.log 2>&1; echo "rc=$?"; tail -4 $OUT/pytest.log
echo "== lz4 bench"; timeout 900 python tools/lz4_bench.py > $OUT/lz4_bench.jsonl 2> $OUT/lz4_bench.err; echo "rc=$?"; grep -v "warp sequence\|warp group" $OUT/lz4_bench.jsonl | cut -c1-260; tail -3 $OUT/lz4_bench.err
echo "== decode kernel times (FLAGSTAT_CUDA_DEBUG: batches serialised)"; FLAGSTAT_CUDA_DEBUG=1 timeout 600 python tools/lz4_bench.py --quick > /dev/null 2> $OUT/decode_times.txt; python - <<PY
import re,collections
rows=collections.defaultdict(list)
for l in open("$OUT/decode_times.txt"):
    m=re.search(r"block decode: (\d+) blocks, (\d+) -> (\d+) bytes, ([0-9.]+) ms", l)
    if m and int(m.group(1))>=200: rows[(int(m.group(1)), int(m.group(2)))].append(float(m.group(4)))
for (nb,cb),v in sorted(rows.items()):
    rb=nb*1024000
    print(f"decode {nb} blocks comp {cb/1e6:.0f} MB: best {min(v):.3f} ms = {rb/min(v)/1e6:.1f} GB/s out, median {sorted(v)[len(v)//2]:.3f} ms, n={len(v)}")
PY
echo "== phases"; timeout 600 python tools/lz4_phase_probe.py 400 > $OUT/lz4_phases.jsonl 2> $OUT/lz4_phases.err; echo "rc=$?"; grep '"cta"' $OUT/lz4_phases.jsonl | cut -c1-900; tail -3 $OUT/lz4_phases.err
echo "== ncu"; bash tools/ncu_lz4.sh ${TAG}_ncu 296 > $OUT/ncu_session.txt 2>&1; grep -v "^\[" $OUT/ncu_session.txt | head -30
