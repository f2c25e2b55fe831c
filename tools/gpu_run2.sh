#!/bin/bash
# parity tests + full-size bench + ncu launch list + ncu full capture of the fill kernels
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench full"; timeout 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --quick --pairs 1000000 --steps 2 --warmup 1 > gpurun_out/ncu_launch_bench.log 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "== ncu full (score kernel, then traced kernel)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:affine_fill16 -c 1 -o gpurun_out/prof_fill16 -f python bench.py --quick --pairs 300000 --steps 1 --warmup 1 > gpurun_out/ncu_full16.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:affine_fill3 -c 1 -o gpurun_out/prof_fill3t -f python bench.py --quick --pairs 300000 --steps 1 --warmup 1 > gpurun_out/ncu_full3.log 2>&1; echo "rc=$?"
ls -la gpurun_out/
