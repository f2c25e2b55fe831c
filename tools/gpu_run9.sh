#!/bin/bash
mkdir -p gpurun_out
echo "== W kernel parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x --timeout 600 -k "long_pairs or golden" 2>&1 | tail -5
echo "== kbench C4 10kx10k (kind 0), 512 pairs"; timeout 900 python tools/kbench.py --check --check-pairs 2 --kind 0 --n 10000 --m 10000 --pairs 512 --cap-per-pair 4096 wide_cta=0 wide_cta=1 2>&1 | tee gpurun_out/r01f_kbench_c4.txt
echo "== kbench C4 1024 pairs"; timeout 900 python tools/kbench.py --kind 0 --n 10000 --m 10000 --pairs 1024 --cap-per-pair 4096 wide_cta=0 wide_cta=1 2>&1 | tee -a gpurun_out/r01f_kbench_c4.txt
