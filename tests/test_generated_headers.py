"""The two generated headers (curve / field / scalar-field constants, hash-to-curve constants) are exactly what their generators
write: tools/gen_curves_ext.py -> csrc/ecg_curves_ext.cuh, tools/gen_h2c_consts.py -> csrc/ecg_h2c_consts.cuh.  The generators
derive every constant (Montgomery forms, -p^-1 mod 2^32, (p+1)/4, Tonelli-Shanks constants, isogeny coefficients) from the public
curve parameters with Python integers, so a hand edit of a header — or a generator change without regenerating — fails here."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool, header", [("gen_curves_ext.py", "ecg_curves_ext.cuh"), ("gen_h2c_consts.py", "ecg_h2c_consts.cuh")])
def test_header_matches_its_generator(tmp_path, tool, header):
    spec = importlib.util.spec_from_file_location("gen_" + header.split(".")[0], os.path.join(ROOT, "tools", tool))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    committed = os.path.join(ROOT, "elliptic-curves_b200", "csrc", header)
    assert os.path.samefile(mod.OUT, committed)
    mod.OUT = str(tmp_path / header)
    mod.main()
    assert open(mod.OUT).read() == open(committed).read()
