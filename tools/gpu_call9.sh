#!/bin/bash
# call 9: final tree: tests, default bench, ncu of the var-base kernels at the right launch (after the 9 table-building launches)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c9_pytest_gpu.txt 2>&1; tail -3 gpurun_out/c9_pytest_gpu.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/c9_bench_n1.json 2> gpurun_out/c9_bench_n1.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/c9_bench_n1.json").read().strip().splitlines()[-1])
    print("headline", "%.4g"%d["value"], "%.4g"%d["e2e"]["value"], d["bit_exact"], "%.3f"%d["roofline_int"]["frac"], d.get("configs_green"))
    for k,c in d.get("configs",{}).items():
        if isinstance(c,dict) and "value" in c: print(k, "%.4g"%c["value"], "%.4g"%c["e2e"]["value"], c["bit_exact"], "%.3f"%c["roofline_int"]["frac"], "%.3f ms"%c["ms_per_step"], "cpu %.4g"%c["cpu_baseline"]["value"])
except Exception as e: print("ERR", e)
PY
for spec in "k256_varbase:k256_varbase_kernel" "p256_varbase:generic_varbase_kernel"; do
  wl=${spec%%:*}; kn=${spec##*:}
  timeout 400 ncu --set full --clock-control none -k regex:$kn -s 12 -c 1 -o /tmp/prof_r02_$wl python bench.py --workload $wl --steps 1 --warmup 3 --configs none > gpurun_out/c9_ncu_$wl.log 2>&1
  python tools/ncu_summary.py /tmp/prof_r02_$wl.ncu-rep gpurun_out/r02_ncu_$wl.json
  ncu -i /tmp/prof_r02_$wl.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; v=rows[2]
d={k:(x,uu) for k,uu,x in zip(h,u,v) if k.startswith('dram__bytes_') and k.endswith('.sum')}
print(json.dumps(d))" > gpurun_out/r02_dram_$wl.json
  ncu -i /tmp/prof_r02_$wl.ncu-rep --page source --csv 2>/dev/null | python -c "
import csv,sys,collections
rows=list(csv.reader(sys.stdin))
h=rows[0]
try:
    si=h.index('Source'); ei=[i for i,x in enumerate(h) if x.startswith('# Instructions Executed') or x=='Instructions Executed'][0]
except Exception as e:
    print('no source page', h[:12]); sys.exit(0)
agg=collections.Counter()
for r in rows[1:]:
    try: n=float(r[ei].replace(',',''))
    except: continue
    op=r[si].split()[0] if r[si].split() else '?'
    if op.startswith('@'): op=r[si].split()[1]
    agg[op]+=n
tot=sum(agg.values())
for k,v in agg.most_common(25): print(f'{k:24s} {v:14.0f} {100*v/tot:6.2f}%')
" > gpurun_out/r02_instmix_$wl.txt
done
cp /tmp/prof_r02_k256_varbase.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
