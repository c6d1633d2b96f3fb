#!/bin/bash
# call 7: the phase-synchronised k256 kernel in the product: tests, bench, then the round's profiling evidence
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest_gpu.txt 2>&1; tail -4 gpurun_out/c7_pytest_gpu.txt
( time python bench.py --steps 20 --warmup 3 > gpurun_out/c7_bench_n1.json 2> gpurun_out/c7_bench_n1.err ) 2> gpurun_out/c7_time.txt
tail -3 gpurun_out/c7_bench_n1.err; cat gpurun_out/c7_time.txt
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c7_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], "%.3f"%d["roofline_int"]["frac"], d.get("configs_green"), d["clocks"])
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c["bit_exact"], "%.3f"%c["roofline_int"]["frac"], "%.3f ms"%c["ms_per_step"])
except Exception as e: print("ERR", e)
PY
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c7_bench_ref.json 2>/dev/null
# ncu --set full of the dominant kernel of every config (one launch each, after warm-up)
for spec in "k256_varbase:k256_varbase_kernel" "p256_varbase:generic_varbase_kernel" "k256_fixedbase:fixedbase_kernel" "k256_lincomb:msm_bucket_kernel"; do
  wl=${spec%%:*}; kn=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kn -s 3 -c 1 -o gpurun_out/prof_r02_$wl python bench.py --workload $wl --steps 1 --warmup 3 --configs none > gpurun_out/c7_ncu_$wl.log 2>&1
done
# launch lists (shares of a step) for the four configs
for wl in k256_varbase p256_varbase k256_fixedbase k256_lincomb; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/c7_launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --configs none > /dev/null 2>&1
done
ls -la gpurun_out/prof_r02_*.ncu-rep
