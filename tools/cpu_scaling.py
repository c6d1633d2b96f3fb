#!/usr/bin/env python3
"""How does the CPU restatement scale with threads on this host? (development tool)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import ecref, pyref
from bench import synth_scalars
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
n = 1 << 15
k = synth_scalars("k256", 1, 0, n)
xy, _ = ecref.mul_gen_batch("k256", k, nthreads=32)
xy = np.ascontiguousarray(xy).reshape(-1)
for nt in (1, 4, 8, 16, 32, 64, 128, 256):
    t0 = time.perf_counter()
    ecref.mul_batch("k256", k, xy, None, nthreads=nt, variant=0)
    dt = time.perf_counter() - t0
    print(f"threads {nt:4d}: {n/dt:10.0f} mults/s  ({n/dt/nt:8.0f} per thread)")
