#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -3
echo "== kbench C4 auto policy"; for np in 64 444 512 888 1024; do timeout 600 python tools/kbench.py --kind 0 --n 10000 --m 10000 --pairs $np --cap-per-pair 4096 "" 2>&1 | grep trace | sed "s/^/np=$np /"; done | tee gpurun_out/r01h_kbench_c4_auto.txt
echo "== bench"; timeout 1200 python bench.py > gpurun_out/r01h_bench.json 2> gpurun_out/r01h_bench.err; echo rc=$?; tail -c 300 gpurun_out/r01h_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r01h_bench.json'))
print(d['value'], d['e2e']['value'], d['traceback']['value'], d['traceback']['e2e']['value'], d['cpu_baseline'])
for k,v in d['other_workloads'].items(): print(k, v['value'], v['unit'])
P
