#!/bin/bash
# final r01i evidence: bench, reference arm, ncu launch list, ncu full of the checkpoint path's two kernels
mkdir -p gpurun_out
echo "== bench"; timeout 1200 python bench.py > gpurun_out/r01i_bench.json 2> gpurun_out/r01i_bench.err; echo rc=$?
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01i_bench_reference.json 2>> gpurun_out/r01i_bench.err; echo rc=$?
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r01i_launches.csv python bench.py --steps 1 --warmup 1 --pairs 1000000 > gpurun_out/r01i_ncu_bench.log 2>&1; echo rc=$?; wc -l gpurun_out/r01i_launches.csv
echo "== ncu full: checkpoint path"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"affine_ckpt_trace|affine_fill16_kernel" -c 4 -f -o gpurun_out/prof_ckpt_path python tools/kbench.py --pairs 262144 ckpt=1 > gpurun_out/ncu_ckpt_path.log 2>&1; echo rc=$?
