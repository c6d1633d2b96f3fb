#!/usr/bin/env python3
"""Runs the IMAD.WIDE issue-rate microbenchmark once (for an ncu capture: the independent record of the integer peak)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "elliptic-curves_b200"))
import ecgpu
eng = ecgpu.Engine(device_ptrs=True)
for which in (0, 1, 2):
    ops, ms = eng.microbench(which, 4000)
    print(which, f"{ops:.5g} ops/s", f"{ms:.3f} ms")
