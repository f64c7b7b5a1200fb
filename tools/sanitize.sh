#!/bin/bash
# compute-sanitizer over a small slice of the GPU parity tests (round-2 kernels): memcheck and racecheck.
mkdir -p gpurun_out
SEL="golden or ragged or multi_view or edge_cases or long_tile_lists"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/memcheck_r2.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_r2.log | tail -3
timeout 500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 0 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden or multi_view" > gpurun_out/racecheck_r2.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/racecheck_r2.log | tail -3
grep -E "Race reported|hazard" gpurun_out/racecheck_r2.log | sed -E 's/0x[0-9a-f]+//g; s/thread \([0-9,]+\)//g' | sort | uniq -c | sort -rn | head -20
