#!/bin/bash
# One GPU call: parity tests, bench, stage times, ncu launch list and --set full capture of the hot kernels.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
python tools/stage_times.py --views 1 8 > gpurun_out/stage_times.json 2>&1; cat gpurun_out/stage_times.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py --steps 2 > gpurun_out/ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches.txt; cat gpurun_out/launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend|sort|merge|duplicate|scan|preprocess' -s 8 -c 8 -f -o gpurun_out/hot python tools/prof_step.py --steps 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
