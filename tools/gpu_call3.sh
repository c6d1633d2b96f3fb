#!/bin/bash
mkdir -p gpurun_out
( time python bench.py --steps 20 --warmup 3 > gpurun_out/c3_bench_n1.json 2> gpurun_out/c3_bench_n1.err ) 2> gpurun_out/c3_bench_n1.time
tail -3 gpurun_out/c3_bench_n1.err; cat gpurun_out/c3_bench_n1.time
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c3_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", d["value"], d["e2e"]["value"], d["bit_exact"], d["roofline_int"]["frac"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c["bit_exact"], "%.3f"%c["roofline_int"]["frac"], "%.3f ms"%c["ms_per_step"], "cpu %.4g"%c["cpu_baseline"]["value"])
        else: print(k, json.dumps(c)[:600])
except Exception as e: print("ERR", e)
PY
( time python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c3_bench_ref.json 2> gpurun_out/c3_bench_ref.err ) 2>> gpurun_out/c3_bench_n1.time
tail -2 gpurun_out/c3_bench_ref.json | cut -c1-400
