#!/bin/bash
# call 4: P-256 three-kernel path + size-ordered buckets: tests, sanitizer, bench lines, ncu of the new kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/c4_pytest_gpu.txt 2>&1; tail -4 gpurun_out/c4_pytest_gpu.txt
SAN_LIMIT=240 bash tools/gpu_sanitize.sh > gpurun_out/c4_sanitize.txt 2>&1; cat gpurun_out/san_summary.txt
for wl in p256_varbase k256_lincomb; do
  timeout 300 python bench.py --workload $wl --steps 10 --configs none > gpurun_out/c4_bench_$wl.json 2> gpurun_out/c4_bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c4_bench_$wl.json").read().strip().splitlines()[-1]); print("$wl", "%.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], d["bit_exact"], "ms %.3f"%d["ms_per_step"], "dom %.3f"%d["roofline"]["kernel_ms"], d["gpu_launches"])
except Exception as e: print("$wl ERR", e); print(open("gpurun_out/c4_bench_$wl.err").read()[-1500:])
PY
done
ECG_MSM_ORDER=0 timeout 300 python bench.py --workload k256_lincomb --steps 10 --configs none > gpurun_out/c4_bench_lincomb_noorder.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/c4_bench_lincomb_noorder.json').read().strip().splitlines()[-1]); print('lincomb natural order', '%.4g'%d['value'], 'ms %.3f'%d['ms_per_step'], 'dom %.3f'%d['roofline']['kernel_ms'])"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:generic_main_kernel -s 3 -c 1 -o gpurun_out/prof_p256_main_r02 python bench.py --workload p256_varbase --steps 1 --warmup 3 --configs none > gpurun_out/c4_ncu_p256.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:msm_bucket_kernel -s 2 -c 1 -o gpurun_out/prof_k256_bucket_r02 python bench.py --workload k256_lincomb --steps 1 --warmup 3 --configs none > gpurun_out/c4_ncu_bucket.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/c4_launches_p256.csv python bench.py --workload p256_varbase --steps 2 --warmup 3 --configs none > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
