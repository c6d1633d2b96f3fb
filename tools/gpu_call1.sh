#!/bin/bash
# round-2 first GPU call: prepared experiments + baseline sanity + independent record of the integer peak
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt
bash tools/gpu_probe_mem.sh > gpurun_out/c1_probe_mem.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:mb_imad_wide -s 1 -c 1 -o gpurun_out/prof_mb_imad_wide python tools/mb_imad.py > gpurun_out/c1_ncu_mb.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest_gpu.txt 2>&1
tail -3 gpurun_out/c1_pytest_gpu.txt
cat gpurun_out/c1_probe_mem.log | tail -45
