"""CPU tier: the C-ABI library builds, loads, exports every symbol include/se3mpc.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls need a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "se3mpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(se3mpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from learningagileflight_se3_b200 import _lib
    _lib.build()
    L = C.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for nme in names:
        assert hasattr(L, nme), nme
    assert set(_lib.EXPORTS) <= set(names)


def test_default_params_are_the_reference_literals():
    from learningagileflight_se3_b200 import _lib
    p = _lib.default_params()
    assert (p.Jx, p.Jy, p.Jz, p.mass, p.arm_l, p.c_tau) == (0.0023, 0.0023, 0.004, 0.5, 0.35, 0.0245)   # quad_policy.py:37
    assert (p.wrt, p.wqt, p.wthrust, p.wrf, p.wvf, p.wwf) == (5, 80, 0.1, 5, 5, 3)                         # quad_policy.py:38
    assert p.N == 50 and p.dt == 0.1 and p.u_lb == 0 and abs(p.u_ub - 2.44) < 1e-15 and abs(p.w_bound - np.pi / 2) < 1e-15
    assert p.tol == 1e-8 and p.max_iter == 3000 and p.mu_init == 0.1


def test_params_layout_matches_oracle_and_emulation(oracle, emu):
    from learningagileflight_se3_b200 import _lib
    assert C.sizeof(_lib.Params) == C.sizeof(emu.Params) == C.sizeof(oracle.Params)
    po, pl = oracle.default_params(), _lib.default_params()
    for name, _ in _lib.Params._fields_:
        assert getattr(po, name) == getattr(pl, name), name


def test_compute_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from learningagileflight_se3_b200 import Se3MpcError, ops
    with pytest.raises(Se3MpcError):
        ops.solve_host(np.zeros((1, 13)), np.zeros((1, 3)), np.zeros((1, 3)), np.zeros((1, 3)), np.ones(1))
    with pytest.raises(Se3MpcError):
        ops.gate_frame_host(np.zeros((1, 12)), np.zeros((1, 3)), 0.0, 0.0, np.zeros((1, 13)), np.zeros((1, 3)))
    with pytest.raises(Se3MpcError):
        ops.plant_step_host(np.zeros((1, 13)), np.zeros((1, 4)))
    with pytest.raises(Se3MpcError):
        ops.solve(torch.zeros(1, 13, dtype=torch.float64), torch.zeros(1, 3, dtype=torch.float64), torch.zeros(1, 3, dtype=torch.float64),
                  torch.zeros(1, 3, dtype=torch.float64), torch.ones(1, dtype=torch.float64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "learningagileflight_se3_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "se3_oracle" not in txt and "se3o_" not in txt, f
                assert "SE3_EMULATE 1" not in txt, f
