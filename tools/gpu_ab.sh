#!/bin/bash
# One GPU call: parity tests, then stage times for a list of env-variable variants (A/B of kernel options).
# usage: tools/gpu_ab.sh "VAR=val VAR2=val" "VAR=val" ...   (first variant "" = defaults)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
: > gpurun_out/ab.txt
for v in "" "$@"; do
  echo "== variant: [$v]" | tee -a gpurun_out/ab.txt
  env $v python tools/stage_times.py --views 1 8 2>&1 | tail -2 | tee -a gpurun_out/ab.txt
done
