#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== kbench"; timeout 900 python tools/kbench.py --check --pairs 500000 $KB_CONFIGS 2>&1 | tee gpurun_out/kbench.txt
