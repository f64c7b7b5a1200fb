#!/bin/bash
# ncu --set full capture of the two blend kernels of one step (8 views of C2); extra env goes in "$@"
mkdir -p gpurun_out
env "$@" timeout 600 ncu --set full --clock-control none --import-source on -k regex:'blend' -s 2 -c 2 -f -o gpurun_out/blend python tools/prof_step.py --steps 2 > gpurun_out/ncu_blend.log 2>&1
tail -2 gpurun_out/ncu_blend.log
