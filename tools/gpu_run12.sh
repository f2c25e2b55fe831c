#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:affine_fill3w -c 1 -f -o gpurun_out/prof_w python tools/kbench.py --kind 0 --n 10000 --m 10000 --pairs 64 --cap-per-pair 4096 wide_cta=1 > gpurun_out/ncu_w.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_w.log
