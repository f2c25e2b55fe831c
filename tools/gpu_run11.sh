#!/bin/bash
mkdir -p gpurun_out
echo "== parity (long pairs, golden, chunk)"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --timeout 600 -k "long_pairs or golden or multi or c1" 2>&1 | tail -3
for np in 16 444 1024; do
echo "== kbench C4 10kx10k, $np pairs"; timeout 900 python tools/kbench.py --check --check-pairs 2 --kind 0 --n 10000 --m 10000 --pairs $np --cap-per-pair 4096 wide_cta=0 wide_cta=1 2>&1 | tee -a gpurun_out/r01g_kbench_c4.txt
done
echo "== 1-warp kernel, 1700 pairs, 150 GB workspace"; timeout 900 python tools/kbench.py --kind 0 --n 10000 --m 10000 --pairs 1700 --workspace-gb 150 --cap-per-pair 4096 wide_cta=0 2>&1 | tee -a gpurun_out/r01g_kbench_c4.txt
