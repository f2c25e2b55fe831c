#!/bin/bash
# first GPU contact: smoke, parity tests, microbench, a small bench
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== microbench"; timeout 300 ./tools/microbench > gpurun_out/microbench.txt 2>&1; echo "mb rc=$?"; cat gpurun_out/microbench.txt
echo "== bench small"; timeout 600 python bench.py --pairs 1000000 --steps 3 --warmup 3 > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "bench rc=$?"; cat gpurun_out/bench_1m.json; tail -5 gpurun_out/bench_1m.err
