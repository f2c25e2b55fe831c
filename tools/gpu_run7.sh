#!/bin/bash
# HEAD verification: gpu parity suite, smoke, default bench, reference arm, ncu launch list
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -4
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 900 python bench.py > gpurun_out/r01d_bench.json 2> gpurun_out/r01d_bench.err; echo rc=$?; tail -c 600 gpurun_out/r01d_bench.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01d_bench_reference.json 2>> gpurun_out/r01d_bench.err; echo rc=$?; tail -c 400 gpurun_out/r01d_bench_reference.json
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01d_launches.csv python bench.py --steps 1 --warmup 1 --pairs 1000000 > gpurun_out/r01d_ncu_bench.log 2>&1; echo rc=$?; wc -l gpurun_out/r01d_launches.csv
