#!/bin/bash
# call 15: ECG_FLAG_CONSTTIME (tests + cost record), C++ mirror with hash to curve, default bench
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_consttime.py tests/test_host_cpp.py tests/test_h2c.py -m gpu -x -q ) > gpurun_out/c15_pytest_ct.txt 2>&1; tail -5 gpurun_out/c15_pytest_ct.txt
( time python bench.py --steps 20 --warmup 3 > gpurun_out/c15_bench_n1.json 2> gpurun_out/c15_bench_n1.err ) 2> gpurun_out/c15_bench_time.txt
tail -2 gpurun_out/c15_bench_n1.err; cat gpurun_out/c15_bench_time.txt
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c15_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], "%.3f"%d["roofline_int"]["frac"], d.get("configs_green"))
    print(json.dumps(d["configs"]["8_consttime_cost"]["curves"], indent=1))
except Exception as e: print("ERR", e)
PY
