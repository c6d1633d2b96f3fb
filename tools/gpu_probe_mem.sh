#!/bin/bash
# first GPU call of the next session: the prepared "operands through memory" variants (OPT bit 8).
# Build first (here, no GPU needed):
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/kbench tools/kbench.cu
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DECG_FE_ALIGN=16 -o tools/kbench_a16 tools/kbench.cu
mkdir -p gpurun_out
timeout 120 tools/kbench 20 mem 2>&1 | tee gpurun_out/kbench_mem.txt
timeout 120 tools/kbench_a16 20 mem 2>&1 | tee gpurun_out/kbench_mem_a16.txt
