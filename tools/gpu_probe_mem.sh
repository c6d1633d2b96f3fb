#!/bin/bash
# first GPU call of the next session: the prepared "operands through memory" variants (OPT bit 8).
# Build first (here, no GPU needed):
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/kbench tools/kbench.cu
#   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DECG_FE_ALIGN=16 -o tools/kbench_a16 tools/kbench.cu
mkdir -p gpurun_out
timeout 120 tools/kbench 20 mem 2>&1 | tee gpurun_out/kbench_mem.txt
timeout 120 tools/kbench_a16 20 mem 2>&1 | tee gpurun_out/kbench_mem_a16.txt
# warp-balanced bucket kernel: lincomb bench with the default and the two sorted variants
for k in 1 4 8; do
  ECG_MSM_BUCKETS_PER_THREAD=$k timeout 150 python bench.py --workload k256_lincomb --steps 10 > gpurun_out/lincomb_bpt$k.json 2> gpurun_out/lincomb_bpt$k.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/lincomb_bpt$k.json").read().strip().splitlines()[-1]); print("lincomb buckets/thread=$k", d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"])
except Exception as e: print("bpt $k ERR", e)
PY
done
timeout 200 python -m pytest tests/test_zz_experimental_gpu.py -q 2>&1 | tail -2
