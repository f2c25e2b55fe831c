#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/kbench_sweep.txt
for cfg in "-DGNX_F3_STAGE=1 -DGNX_F3_UNROLL_TRACE=1 -DGNX_F3_UNROLL_SCORE=1" "-DGNX_F3_STAGE=0 -DGNX_F3_UNROLL_TRACE=1 -DGNX_F3_UNROLL_SCORE=2" "-DGNX_F3_STAGE=0 -DGNX_F3_UNROLL_TRACE=2 -DGNX_F3_UNROLL_SCORE=4" "-DGNX_F3_STAGE=1 -DGNX_F3_UNROLL_TRACE=1 -DGNX_F3_UNROLL_SCORE=4 -DGNX_FILL3_MINB=12"; do
  echo "== build $cfg" | tee -a gpurun_out/kbench_sweep.txt
  GNX_NVCC_EXTRA="$cfg" python -c "from gonomics_b200 import build; build.build(force=True)" 2>&1 | grep -i error | head -3
  timeout 600 python tools/kbench.py --check --pairs 500000 fill_impl=3,skew=1 fill_impl=3,skew=2 2>&1 | tee -a gpurun_out/kbench_sweep.txt
done
