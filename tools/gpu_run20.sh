#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -3
echo "== ncu full: checkpoint path (final)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"affine_ckpt_trace|affine_fill16_kernel|ckpt_classify" --launch-skip 4 -c 4 -f -o gpurun_out/prof_ckpt_final python tools/kbench.py --pairs 262144 ckpt=1 > gpurun_out/ncu_ckpt_final.log 2>&1; echo rc=$?
