#!/bin/bash
mkdir -p gpurun_out
for np in 16 148 444; do
echo "== kbench C4 10kx10k, $np pairs"; timeout 900 python tools/kbench.py --kind 0 --n 10000 --m 10000 --pairs $np --cap-per-pair 4096 wide_cta=0 wide_cta=1 2>&1 | tee -a gpurun_out/r01f_kbench_c4_small.txt
done
echo "== 1-warp kernel, 1700 pairs, 150 GB workspace"; timeout 900 python tools/kbench.py --kind 0 --n 10000 --m 10000 --pairs 1700 --workspace-gb 150 --cap-per-pair 4096 wide_cta=0 2>&1 | tee -a gpurun_out/r01f_kbench_c4_small.txt
