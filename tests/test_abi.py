"""CPU-only: the C-ABI library loads and exports every symbol include/ecgpu.h declares; host-side argument
checks that need no GPU; no silent CPU fallback."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "ecgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ecg_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import ecgpu

    lib = ecgpu.load_library()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ecgpu.h but not exported by libecgpu.so"
    assert sorted(ecgpu.EXPORTS) == syms
    assert b"sm_100a" in lib.ecg_version()


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device ctx creation must fail loudly (ECG_ECUDA), never compute on the CPU."""
    import torch

    import ecgpu

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ecgpu.EcgError) as ei:
        ecgpu.Engine()
    assert ei.value.code == ecgpu.ECG_ECUDA


def test_null_ctx_is_rejected():
    import ecgpu

    lib = ecgpu.load_library()
    assert lib.ecg_mul_batch(None, 0, 1, None, None, None, None, None) == ecgpu.ECG_EINVAL
    assert lib.ecg_mul_gen_batch(None, 0, 1, None, None, None) == ecgpu.ECG_EINVAL
    assert lib.ecg_lincomb(None, 0, 1, None, None, None, None, None) == ecgpu.ECG_EINVAL
    assert lib.ecg_kernel_launches(None) == 0
    out = ctypes.c_void_p(0)
    assert lib.ecg_ctx_create(None, 99, 0, ctypes.byref(out)) == ecgpu.ECG_EINVAL
    assert lib.ecg_ctx_create(None, 2, ecgpu.FLAG_DEVICE_PTRS, ctypes.byref(out)) == ecgpu.ECG_EINVAL


def test_product_does_not_import_oracle():
    """The product path must not touch oracle/ (parity claims are void otherwise)."""
    pkg = os.path.join(ROOT, "elliptic-curves_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "ecref" not in txt and "pyref" not in txt and "oracle/" not in txt.replace("see oracle/", ""), f


def test_sec1_compress_is_pure_byte_layout():
    """Engine.sec1_compress needs no device: tag by y parity, x, identity = 33 zero bytes; inverse of the decompressor's input"""
    import numpy as np
    import ecgpu
    import pyref

    c = pyref.K256
    pts = [pyref.mul(c, k, pyref.G(c)) for k in (1, 2, 3, 7)]
    xy = np.frombuffer(b"".join(P[0].to_bytes(32, "big") + P[1].to_bytes(32, "big") for P in pts) + bytes(64), np.uint8).reshape(5, 64)
    inf = np.array([0, 0, 0, 0, 1], np.uint8)
    rec = ecgpu.Engine.sec1_compress(xy, inf)
    assert rec.shape == (5, 33)
    for r, P in zip(rec, pts):
        assert r[0] == 2 + (P[1] & 1) and int.from_bytes(r[1:].tobytes(), "big") == P[0]
    assert not rec[4].any()
    # the generator's well-known compressed encoding
    assert rec[0].tobytes().hex() == "0279be667ef9dcbbac55a06295ce870b07029bfcdb2dce28d959f2815b16f81798"
