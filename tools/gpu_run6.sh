#!/bin/bash
mkdir -p gpurun_out
echo "== long-pair tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 240 -k "long_pairs or cigar_to_bed or cpp_host" 2>&1 | tail -3
echo "== cpp host"; timeout 300 python -m pytest tests/test_cpp_host.py -m gpu -q -x --timeout 240 2>&1 | tail -2
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -2
echo "== C4 kbench"; timeout 600 python tools/kbench.py --check --check-pairs 4 --n 10000 --m 10000 --pairs 1000 --kind 0 --cap-per-pair 4000 long_warps=1 long_warps=4 2>&1 | tee gpurun_out/kbench_c4.txt
