#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traceback_kernel -c 1 -f -o gpurun_out/prof_tb python tools/kbench.py --pairs 262144 fill_impl=3 > gpurun_out/ncu_tb.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_tb.log
