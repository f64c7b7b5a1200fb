#!/bin/bash
# first-call probe of the GPU box
mkdir -p gpurun_out
{
echo "== nvidia-smi"; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv
echo "== cpu"; nproc; grep -m1 'model name' /proc/cpuinfo; grep -o -m1 -w 'avx2' /proc/cpuinfo; grep -o -m1 -w 'fma' /proc/cpuinfo; grep -o -m1 -w 'avx512f' /proc/cpuinfo; free -g | head -2
echo "== baseline/_ref"; ls -la baseline/_ref 2>&1 | head
echo "== upstream ext"; (cd /tmp && python -c "import diff_gaussian_rasterization as d; print(d.__file__)" 2>&1 | tail -1)
(cd /tmp && python -c "import gsplat; print('gsplat', gsplat.__file__)" 2>&1 | tail -1)
echo "== MEASURED_PEAKS"; cat MEASURED_PEAKS.json 2>/dev/null | head -5
} > gpurun_out/probe.txt 2>&1
cat gpurun_out/probe.txt
