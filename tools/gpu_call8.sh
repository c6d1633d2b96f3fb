#!/bin/bash
# call 8: final-tree tests + bench, then the round's profiling evidence (summaries made on the box: reports are too big to travel)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c8_pytest_gpu.txt 2>&1; tail -4 gpurun_out/c8_pytest_gpu.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/c8_bench_n1.json 2> gpurun_out/c8_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c8_bench_ref.json 2>/dev/null
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c8_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], "%.3f"%d["roofline_int"]["frac"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c["bit_exact"], "%.3f"%c["roofline_int"]["frac"], "%.3f ms"%c["ms_per_step"])
except Exception as e: print("ERR", e)
PY
for spec in "k256_varbase:k256_varbase_kernel" "p256_varbase:generic_varbase_kernel" "k256_fixedbase:fixedbase_kernel" "k256_lincomb:msm_bucket_kernel"; do
  wl=${spec%%:*}; kn=${spec##*:}
  timeout 400 ncu --set full --clock-control none -k regex:$kn -s 3 -c 1 -o /tmp/prof_r02_$wl python bench.py --workload $wl --steps 1 --warmup 3 --configs none > gpurun_out/c8_ncu_$wl.log 2>&1
  python tools/ncu_summary.py /tmp/prof_r02_$wl.ncu-rep gpurun_out/r02_ncu_$wl.json
  ncu -i /tmp/prof_r02_$wl.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; v=rows[2]
d={k:(x,uu) for k,uu,x in zip(h,u,v) if k.startswith('dram__bytes_') and k.endswith('.sum')}
print(json.dumps(d))" > gpurun_out/r02_dram_$wl.json
  ncu -i /tmp/prof_r02_$wl.ncu-rep --page details --csv 2>/dev/null | grep -i "inst_executed\|Executed Ipc\|Issue Slots\|Registers\|Local" | head -40 > gpurun_out/r02_details_$wl.csv
done
cp /tmp/prof_r02_k256_varbase.ncu-rep gpurun_out/ 2>/dev/null
for wl in k256_varbase p256_varbase k256_fixedbase k256_lincomb; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 120 --csv --log-file gpurun_out/r02_launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --configs none > /dev/null 2>&1
done
du -sh gpurun_out
