#!/bin/bash
# quick confirmation after a host-side change: parity suite + the two e2e-sensitive bench workloads
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/check_pytest.txt
timeout 150 python bench.py > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err; python - <<'PY'
import json
for f in ("gpurun_out/check_bench.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"])
    except Exception as e: print(f, "ERR", e)
PY
timeout 150 python bench.py --workload k256_fixedbase > gpurun_out/check_bench_fb.json 2> gpurun_out/check_bench_fb.err; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/check_bench_fb.json").read().strip().splitlines()[-1]); print("fb", d["value"], d["e2e"]["value"])
except Exception as e: print("fb ERR", e)
PY
