#!/bin/bash
# call 26: the round's closing run on one B200 — smoke(), every GPU test, the default bench line
mkdir -p gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/c26_smoke.txt 2>&1; tail -3 gpurun_out/c26_smoke.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/c26_pytest_gpu.txt 2>&1; tail -5 gpurun_out/c26_pytest_gpu.txt
( time python bench.py --steps 20 --warmup 3 > gpurun_out/c26_bench_n1.json 2> gpurun_out/c26_bench_n1.err ) 2> gpurun_out/c26_bench_time.txt
tail -2 gpurun_out/c26_bench_n1.err; cat gpurun_out/c26_bench_time.txt
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c26_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], "%.3f"%d["roofline_int"]["frac"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c.get("bit_exact"), "%.3f ms"%c.get("ms_per_step",0))
    for n,v in d["configs"]["9_hash_to_curve"]["curves"].items(): print("h2c", n, "%.4g"%v["value"], "%.4g"%v["e2e"]["value"], v["bit_exact"])
except Exception as e: print("ERR", e)
PY
