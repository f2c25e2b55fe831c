#!/bin/bash
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 1200 python bench.py > gpurun_out/r01o_bench.json 2> gpurun_out/r01o_bench.err; echo rc=$?; tail -c 300 gpurun_out/r01o_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r01o_bench.json'))
print(d['value'], d['e2e']['value'], d['traceback']['value'], d['traceback']['e2e']['value'], d['traceback']['roofline']['frac'], d['issue_roofline']['frac'])
for k,v in d['other_workloads'].items(): print(k, v['value'], v['unit'])
print(d['cpu_baseline'])
P
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01o_launches.csv python bench.py --steps 1 --warmup 1 --pairs 1000000 --no-cpu > gpurun_out/r01o_ncu_bench.log 2>&1; echo rc=$?; wc -l gpurun_out/r01o_launches.csv
