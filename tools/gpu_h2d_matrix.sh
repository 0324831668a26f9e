#!/bin/bash
# Host->device fan-out diagnosis on an N-GPU box.
#   gpurun --gpus 8 -- bash tools/gpu_h2d_matrix.sh r4a
TAG=${1:-r4_h2d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
(numactl -H || echo "numactl not installed") > $OUT/numactl.txt 2>&1
lscpu > $OUT/lscpu.txt 2>&1
(lspci -tv || echo "lspci not installed") > $OUT/lspci_tree.txt 2>&1
ls /sys/devices/system/node/ > $OUT/sys_nodes.txt 2>&1
cat /proc/meminfo | head -5 > $OUT/meminfo.txt
nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv > $OUT/pcie_links.csv 2>&1
timeout 300 tools/bin/h2d_matrix 512 6 > $OUT/h2d_matrix.jsonl 2> $OUT/h2d_matrix.err; echo "rc=$?"
wc -l $OUT/h2d_matrix.jsonl; head -12 $OUT/h2d_matrix.jsonl | cut -c1-400; grep -E '"(four|all)' $OUT/h2d_matrix.jsonl | cut -c1-300
cat $OUT/numactl.txt | head; cat $OUT/topo.txt | head -12
