#!/bin/bash
# twobit/seed bench block + ncu capture of the pack kernel
mkdir -p gpurun_out
echo "== bench (1M pairs, with extras)"; timeout 900 python bench.py --pairs 1000000 --steps 2 --warmup 3 --no-traceback > gpurun_out/r01e_bench_small.json 2> gpurun_out/r01e_bench_small.err; echo rc=$?; tail -c 300 gpurun_out/r01e_bench_small.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r01e_bench_small.json'))
for k,v in d['other_workloads'].items(): print(k, json.dumps(v)[:600])
print('cpu', d.get('cpu_baseline'))
P
echo "== ncu pack"; cat > /tmp/packprof.py <<'P'
import sys; sys.path.insert(0,'.')
import torch
from gonomics_b200 import align, _lib
L=_lib.load(); ctx=align.Context(0)
n=1<<31
seq=torch.randint(0,4,(n,),dtype=torch.uint8,device='cuda'); words=torch.zeros(n//32,dtype=torch.int64,device='cuda')
for _ in range(2):
    ctx._check(L.gnx_twobit_pack_device(ctx._h, seq.data_ptr(), n, 0, words.data_ptr(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:twobit_pack -c 1 -f -o gpurun_out/prof_pack python /tmp/packprof.py > gpurun_out/ncu_pack.log 2>&1; echo rc=$?; tail -2 gpurun_out/ncu_pack.log
