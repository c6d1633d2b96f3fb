#!/bin/bash
# call 27 (the round's last GPU seconds): P-521 after the Mersenne reduction — every output of two whole waves vs the C
# restatement + kernel time, then the P-521 GPU tests
mkdir -p gpurun_out
timeout 80 python tools/gpu_p521_check.py > gpurun_out/c27_p521_check.json 2> gpurun_out/c27_p521_check.err; echo "check rc=$?"; cat gpurun_out/c27_p521_check.json; tail -2 gpurun_out/c27_p521_check.err
( time timeout 100 python -m pytest tests/test_gpu_curves_ext.py tests/test_ecdsa_ext.py tests/test_sec1_ext.py tests/test_h2c.py -m gpu -q -x -k "11 or p521" ) > gpurun_out/c27_pytest_p521.txt 2>&1; tail -4 gpurun_out/c27_pytest_p521.txt
