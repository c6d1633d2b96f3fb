#!/bin/bash
# round-end confirmation on the GPU box: parity suite, default bench, reference arm, then compute-sanitizer passes.
# every step has its own timeout so that the whole script stays inside the gpurun limit (sum of limits < 840 s)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/final_pytest.txt
timeout 150 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; cat gpurun_out/final_bench.json
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; cat gpurun_out/final_ref.json
for t in memcheck racecheck; do
  timeout 130 compute-sanitizer --tool $t --error-exitcode 9 python tools/sanitize.py > gpurun_out/san_$t.log 2>&1
  echo "$t rc=$?" | tee -a gpurun_out/san_summary.txt
  tail -3 gpurun_out/san_$t.log
done
