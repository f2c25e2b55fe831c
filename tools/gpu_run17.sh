#!/bin/bash
# compute-sanitizer memcheck over the new kernels (small cases)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log python -m pytest tests/test_gpu_parity.py tests/test_gpu_twobit.py -q -x --timeout 1200 -k "checkpoint_recompute_traceback[500-150] or checkpoint_recompute_traceback[45-9] or checkpoint_recompute_traceback[2-1] or cta_per_pair_kernel[1] or seed_index_and_seeds_match_oracle[20-8] or pack_ragged_batches_all_leads[13] or count_matches or smoke or golden_affine_local" 2>&1 | tail -4
echo rc=$?; grep -c "Invalid\|Error" gpurun_out/memcheck.log; tail -5 gpurun_out/memcheck.log
