#!/bin/bash
# call 21: SEC1 decompression for all twelve curves, SM2DSA; memcheck over those kernels; the default bench with the
# hash-to-curve e2e leg on host buffers in the C ABI's own layout and the BIP340 verification record
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_sec1_ext.py tests/test_sm2dsa.py tests/test_h2c.py tests/test_gpu_curves_ext.py tests/test_gpu_p384.py -m gpu -q -x ) > gpurun_out/c21_pytest_gpu.txt 2>&1; tail -5 gpurun_out/c21_pytest_gpu.txt
( time timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_sec1.py ) > gpurun_out/c21_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/c21_memcheck.log
( time python bench.py --steps 20 --warmup 3 > gpurun_out/c21_bench_n1.json 2> gpurun_out/c21_bench_n1.err ) 2> gpurun_out/c21_bench_time.txt
tail -2 gpurun_out/c21_bench_n1.err; cat gpurun_out/c21_bench_time.txt
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c21_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], "%.3f"%d["roofline_int"]["frac"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c.get("bit_exact"), "%.3f ms"%c.get("ms_per_step",0))
    for n,v in d["configs"]["9_hash_to_curve"]["curves"].items(): print("h2c", n, "%.4g"%v["value"], "%.4g"%v["e2e"]["value"], v["bit_exact"])
except Exception as e: print("ERR", e)
PY
