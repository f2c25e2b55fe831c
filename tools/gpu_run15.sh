#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -3
echo "== bench"; timeout 1200 python bench.py > gpurun_out/r01i_bench.json 2> gpurun_out/r01i_bench.err; echo rc=$?; tail -c 300 gpurun_out/r01i_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r01i_bench.json'))
print(d['value'], d['e2e']['value'], d['traceback']['value'], d['traceback']['e2e']['value'], d['traceback']['roofline'])
for k,v in d['other_workloads'].items(): print(k, v['value'], v['unit'])
P
