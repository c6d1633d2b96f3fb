#!/bin/bash
# round-2 call 2: kernel-structure experiments (kbench r2), full GPU test suite on the current tree, lincomb launch list
mkdir -p gpurun_out
timeout 400 tools/kbench 20 r2 > gpurun_out/c2_kbench_r2.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest_gpu.txt 2>&1
tail -5 gpurun_out/c2_pytest_gpu.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/c2_launches_lincomb.csv python bench.py --workload k256_lincomb --steps 2 --warmup 3 > gpurun_out/c2_lincomb_under_ncu.log 2>&1
cat gpurun_out/c2_kbench_r2.txt
