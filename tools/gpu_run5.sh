#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:affine_fill3 -s 3 -c 2 -f -o gpurun_out/prof_fill3 python tools/kbench.py --pairs 200000 fill_impl=3 > gpurun_out/ncu_fill3.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_fill3.log
ls -la gpurun_out/*.ncu-rep
