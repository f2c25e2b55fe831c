#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/kbench_sweep.txt
for cfg in "-DGNX_F3_GROUP4=0" "-DGNX_F3_GROUP4=1"; do
  echo "== build $cfg" | tee -a gpurun_out/kbench_sweep.txt
  GNX_NVCC_EXTRA="$cfg" python -c "from gonomics_b200 import build; build.build(force=True)" 2>&1 | grep -i error | head -3
  timeout 600 python tools/kbench.py --check --pairs 500000 tb_impl=2 2>&1 | grep trace | tee -a gpurun_out/kbench_sweep.txt
done
echo "== pytest gpu (GROUP4=1 build)"; timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -2
