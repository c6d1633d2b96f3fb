#!/bin/bash
# compute-sanitizer passes over every entry point (bounded: each tool has its own timeout)
mkdir -p gpurun_out; rm -f gpurun_out/san_summary.txt
for t in memcheck racecheck; do
  timeout ${SAN_LIMIT:-200} compute-sanitizer --tool $t --error-exitcode 9 python tools/sanitize.py > gpurun_out/san_$t.log 2>&1
  echo "$t rc=$?" | tee -a gpurun_out/san_summary.txt
  tail -4 gpurun_out/san_$t.log
done
