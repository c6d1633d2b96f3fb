#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest_gpu.txt 2>&1; tail -6 gpurun_out/c6_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c6_smoke.txt 2>&1; tail -2 gpurun_out/c6_smoke.txt
( time python bench.py --steps 20 --warmup 3 > gpurun_out/c6_bench_n1.json 2> gpurun_out/c6_bench_n1.err ) 2> gpurun_out/c6_time.txt
tail -3 gpurun_out/c6_bench_n1.err; cat gpurun_out/c6_time.txt
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c6_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], "%.3f"%d["roofline_int"]["frac"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c["bit_exact"], "%.3f"%c["roofline_int"]["frac"], "%.3f ms"%c["ms_per_step"])
except Exception as e: print("ERR", e)
PY
