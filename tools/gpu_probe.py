#!/usr/bin/env python3
"""First-light GPU probe: integer-pipe microbenchmarks + a quick device-resident timing of ecg_mul_batch.
Writes gpurun_out/probe.json.  (Development tool; bench.py is the contract.)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "elliptic-curves_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch

import ecgpu

out = {}
eng = ecgpu.Engine(device_ptrs=True)
names = {0: "imad_wide_x", 1: "imad_lo", 2: "iadd3_x", 3: "fmul_k256", 4: "fmul_p256"}
for which in (0, 1, 2, 3, 4):
    try:
        ops, ms = eng.microbench(which, 4000)
        out[names[which]] = {"ops_per_s": ops, "ms": ms}
        print(names[which], f"{ops:.4g} ops/s  ({ms:.3f} ms)")
    except Exception as e:  # noqa
        print(names[which], "n/a", e)

dev = torch.device("cuda:0")
for curve in sys.argv[1:] or ["k256"]:
    for logn in (14, 17, 20):
        n = 1 << logn
        g = torch.Generator(device="cpu").manual_seed(1)
        k = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g)
        k[:, 0] = 0x7F  # < n for both curves
        import pyref

        c = pyref.CURVES[curve]
        Gxy, _ = pyref.enc_point(pyref.G(c))
        P = torch.from_numpy(np.frombuffer(Gxy * n, dtype=np.uint8).copy()).reshape(n, 64)
        kd, Pd = k.to(dev), P.to(dev)
        oxy = torch.empty((n, 64), dtype=torch.uint8, device=dev)
        oinf = torch.empty((n,), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        try:
            for rep in range(3):
                t0 = time.perf_counter()
                eng.mul_batch_ptr(curve, n, kd.data_ptr(), Pd.data_ptr(), 0, oxy.data_ptr(), oinf.data_ptr())
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            print(curve, f"n=2^{logn}: {dt*1e3:.2f} ms  -> {n/dt:.4g} mults/s")
            out[f"{curve}_mul_2^{logn}"] = {"ms": dt * 1e3, "per_s": n / dt}
        except Exception as e:  # noqa
            print(curve, "failed:", e)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"), indent=1)
