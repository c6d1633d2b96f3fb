#!/bin/bash
# call 20: every GPU test (with the generic SEC1 decompression / square root), then the default bench (hash-to-curve record,
# whole-wave batches for the other curves)
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/c20_pytest_gpu.txt 2>&1; tail -6 gpurun_out/c20_pytest_gpu.txt
( time python bench.py --steps 20 --warmup 3 > gpurun_out/c20_bench_n1.json 2> gpurun_out/c20_bench_n1.err ) 2> gpurun_out/c20_bench_time.txt
tail -2 gpurun_out/c20_bench_n1.err; cat gpurun_out/c20_bench_time.txt
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c20_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], "%.3f"%d["roofline_int"]["frac"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], c.get("bit_exact"), "%.3f ms"%c.get("ms_per_step",0))
        elif isinstance(c,dict) and "curves" in c:
            for n,v in c["curves"].items(): print("  ", n, "%.4g"%v["value"], "%.4g"%v["e2e"]["value"], v["bit_exact"], "%.3f"%v["roofline_int"]["frac"], "%.3f ms"%v["kernel_ms"], "cpu %.4g"%v["cpu_baseline"]["value"])
        else: print(k, json.dumps(c)[:400])
except Exception as e: print("ERR", e)
PY
