#!/bin/bash
# Final evidence of a round on ONE GPU: bench lines, ncu launch list of the bench command, --set full of the step kernels.
mkdir -p gpurun_out
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json; tail -2 gpurun_out/final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err; cat gpurun_out/final_bench_ref.json
for c in c4 c5; do python bench.py --config $c > gpurun_out/final_bench_$c.json 2>> gpurun_out/final_bench.err; tail -c 400 gpurun_out/final_bench_$c.json; done
python tools/stage_times.py --views 1 8 > gpurun_out/final_stage_times.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 1 --quick --no-cpu > gpurun_out/final_ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/final_launches.csv --step-only > gpurun_out/final_launches.txt; python tools/launch_summary.py gpurun_out/final_launches.csv >> gpurun_out/final_launches.txt; head -14 gpurun_out/final_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend|sort|merge|duplicate|scan|preprocess' -s 9 -c 9 -f -o gpurun_out/final_hot python tools/prof_step.py --steps 2 > gpurun_out/final_ncu_full.log 2>&1
tail -2 gpurun_out/final_ncu_full.log
python tools/ncu_summary.py gpurun_out/final_hot.ncu-rep > gpurun_out/final_hot_summary.txt 2>&1; grep -c "====" gpurun_out/final_hot_summary.txt
