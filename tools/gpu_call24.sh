#!/bin/bash
# call 24: ECDSA public-key recovery on the GPU (tests + the C++ mirror's full surface), then the default bench with the
# recovery record
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_ecdsa_recover.py tests/test_host_cpp.py tests/test_sm2dsa.py -m gpu -q -x ) > gpurun_out/c24_pytest_gpu.txt 2>&1; tail -5 gpurun_out/c24_pytest_gpu.txt
( time python bench.py --steps 20 --warmup 3 > gpurun_out/c24_bench_n1.json 2> gpurun_out/c24_bench_n1.err ) 2> gpurun_out/c24_bench_time.txt
tail -2 gpurun_out/c24_bench_n1.err; cat gpurun_out/c24_bench_time.txt
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c24_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], "%.3f"%d["roofline_int"]["frac"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c.get("bit_exact"), "%.3f ms"%c.get("ms_per_step",0))
except Exception as e: print("ERR", e)
PY
