#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
echo "== const kbench"; timeout 600 python tools/kbench.py --check --check-pairs 3000 --kind 2 --pairs 300000 --cap-per-pair 400 fill_impl=1 fill_impl=3 2>&1 | tee gpurun_out/kbench_const.txt
