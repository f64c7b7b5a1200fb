#!/bin/bash
# Quick GPU check: parity tests, stage times, short bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
[ $rc -ne 0 ] && exit 1
python tools/stage_times.py --views 1 8 > gpurun_out/stage_times.json 2>&1; cat gpurun_out/stage_times.json
python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 2500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
