#!/bin/bash
# N = 2 on the final tree: the default line under torchrun with the hash-to-curve and BIP340 records (all ranks), reference arm as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c22_gpus.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c22_bench_n2.json 2> gpurun_out/c22_bench_n2.err ) 2> gpurun_out/c22_time.txt
tail -5 gpurun_out/c22_bench_n2.err; cat gpurun_out/c22_time.txt
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/c22_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c["bit_exact"], "%.3f ms"%c["ms_per_step"], c.get("exchange"))
        else: print(k, json.dumps(c)[:900])
except Exception as e: print("ERR", e)
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/c22_ref_n2.json 2> gpurun_out/c22_ref_n2.err ) 2> gpurun_out/c22_ref_time.txt
tail -c 600 gpurun_out/c22_ref_n2.json; cat gpurun_out/c22_ref_time.txt
