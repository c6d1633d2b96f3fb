#!/bin/bash
# call 13: P-521 (17 limbs, 66-byte records) + everything again after the record-size decoupling; then the default bench
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/c13_pytest_gpu.txt 2>&1; tail -6 gpurun_out/c13_pytest_gpu.txt
( time python bench.py --steps 20 --warmup 3 > gpurun_out/c13_bench_n1.json 2> gpurun_out/c13_bench_n1.err ) 2> gpurun_out/c13_bench_time.txt
tail -2 gpurun_out/c13_bench_n1.err; cat gpurun_out/c13_bench_time.txt
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c13_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], "%.3f"%d["roofline_int"]["frac"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c["bit_exact"], "%.3f"%c["roofline_int"]["frac"], "%.3f ms"%c["ms_per_step"], "cpu %.4g"%c["cpu_baseline"]["value"])
        elif isinstance(c,dict) and "curves" in c:
            for n,v in c["curves"].items(): print("  ", n, "%.4g"%v["value"], "%.4g"%v["e2e"]["value"], v["bit_exact"], "%.3f"%v["roofline_int"]["frac"], "%.3f ms"%v["kernel_ms"], "cpu %.4g"%v["cpu_baseline"]["value"])
except Exception as e: print("ERR", e)
PY
