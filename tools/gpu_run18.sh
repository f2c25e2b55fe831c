#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/racecheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_twobit.py -q -x --timeout 1200 -k "checkpoint_recompute_traceback[45-9] or checkpoint_recompute_traceback[64-30] or seed_index_and_seeds_match_oracle[20-8] or cta_per_pair_kernel[1]" 2>&1 | tail -4
echo rc=$?; grep -c "hazard" gpurun_out/racecheck.log; grep "hazard" gpurun_out/racecheck.log | sort | uniq -c | sort -rn | head -8; tail -3 gpurun_out/racecheck.log
