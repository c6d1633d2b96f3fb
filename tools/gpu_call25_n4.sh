#!/bin/bash
# N = 4: the default bench line as the driver launches it (all configs, NCCL exchange, strong scaling, multi-device parity)
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/c25_ngpus.txt
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/c25_bench_n4.json 2> gpurun_out/c25_bench_n4.err ) 2> gpurun_out/c25_time.txt
tail -3 gpurun_out/c25_bench_n4.err; cat gpurun_out/c25_time.txt
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/c25_bench_n4.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c["bit_exact"], "%.3f ms"%c["ms_per_step"], c.get("exchange"))
        else: print(k, json.dumps(c)[:700])
except Exception as e: print("ERR", e)
PY
