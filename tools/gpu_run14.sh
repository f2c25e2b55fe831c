#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:affine_ckpt_trace -c 1 -f -o gpurun_out/prof_ckpt python tools/kbench.py --pairs 262144 ckpt=1 > gpurun_out/ncu_ckpt.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_ckpt.log
