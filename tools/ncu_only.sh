mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py --steps 2 > gpurun_out/ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches.txt; cat gpurun_out/launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend|sort|merge|duplicate|scan|preprocess' -s 8 -c 8 -f -o gpurun_out/hot python tools/prof_step.py --steps 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
