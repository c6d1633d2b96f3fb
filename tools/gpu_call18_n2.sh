#!/bin/bash
# N = 2: NCCL scatter/gather + exchange paths, multi-device ctx, the full bench line with configs at world size 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c18_gpus.txt
timeout 900 python -m pytest tests/test_dist_nccl.py tests/test_gpu_parity.py -m gpu -q -k "nccl or multi_device" > gpurun_out/c18_pytest_n2.txt 2>&1; tail -4 gpurun_out/c18_pytest_n2.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c18_bench_n2.json 2> gpurun_out/c18_bench_n2.err ) 2> gpurun_out/c18_time.txt
tail -5 gpurun_out/c18_bench_n2.err; cat gpurun_out/c18_time.txt
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/c18_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c["bit_exact"], "%.3f ms"%c["ms_per_step"], c.get("exchange"))
        else: print(k, json.dumps(c)[:900])
except Exception as e: print("ERR", e)
PY
