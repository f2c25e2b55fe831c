#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none -k regex:seed_kernel -c 2 -f -o gpurun_out/prof_seed python tools/seedprof.py > gpurun_out/ncu_seed.log 2>&1; echo rc=$?
timeout 300 ncu --set full --clock-control none -k regex:"affine_fill3w|traceback_affine_warp" -c 2 -f -o gpurun_out/prof_long python tools/kbench.py --kind 0 --n 10000 --m 10000 --pairs 444 --cap-per-pair 4096 "" > gpurun_out/ncu_long.log 2>&1; echo rc=$?
