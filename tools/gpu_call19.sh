#!/bin/bash
# call 19: compute-sanitizer memcheck over the new code paths; ncu of the generic var-base kernel (sm2, P-521) and of h2c_kernel
mkdir -p gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_ext.py sm2 bp256r1 bignp256 p224 p192 ) > gpurun_out/c19_memcheck_a.log 2>&1; echo "memcheck a rc=$?" | tee gpurun_out/c19_san_summary.txt; tail -3 gpurun_out/c19_memcheck_a.log
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_ext.py bp384r1 p384 p521 ) > gpurun_out/c19_memcheck_b.log 2>&1; echo "memcheck b rc=$?" | tee -a gpurun_out/c19_san_summary.txt; tail -3 gpurun_out/c19_memcheck_b.log
for spec in sm2:9 p521:18; do
  name=${spec%%:*}; skip=${spec##*:}   # skip the launches that build the curve's fixed-base table
  timeout 400 ncu --set full --clock-control none -k regex:generic_varbase_kernel -s $skip -c 1 -o /tmp/prof_r02_ext_$name python tools/gpu_ncu_ext.py $name > gpurun_out/c19_ncu_$name.log 2>&1
  python tools/ncu_summary.py /tmp/prof_r02_ext_$name.ncu-rep gpurun_out/r02_ncu_${name}_varbase.json; tail -2 gpurun_out/c19_ncu_$name.log
done
timeout 300 ncu --set full --clock-control none -k regex:h2c_kernel -s 1 -c 1 -o /tmp/prof_r02_h2c python tools/gpu_ncu_ext.py k256 > gpurun_out/c19_ncu_h2c.log 2>&1
python tools/ncu_summary.py /tmp/prof_r02_h2c.ncu-rep gpurun_out/r02_ncu_k256_h2c.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_sm2_varbase.csv python tools/gpu_ncu_ext.py sm2 > /dev/null 2>&1
ls -la gpurun_out | tail -12
