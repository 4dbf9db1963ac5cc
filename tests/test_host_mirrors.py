"""CPU tier: host-side mirrors that need no GPU -- solid_geometry.plane / line against the imported reference classes, and the
batched-model helper of learning.rl_step against the reference's network (which takes its batch transposed, quad_nn.py:118)."""
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

REF = "/root/reference"


def _import_reference(name):
    if not os.path.isdir(REF):
        pytest.skip("/root/reference not present (GPU box)")
    cas = types.ModuleType("casadi")
    cas.numpy = np; cas.np = np; cas.atan = math.atan; cas.pi = math.pi
    saved = {k: sys.modules.get(k) for k in ("casadi", "matplotlib", "matplotlib.pyplot", "matplotlib.animation", name)}
    sys.modules["casadi"] = cas
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules.pop(name, None)
    sys.path.insert(0, REF)
    try:
        return __import__(name)
    finally:
        sys.path.remove(REF)
        mod = sys.modules.pop(name, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            elif k != name:
                sys.modules.pop(k, None)
        if mod is not None:
            sys.modules["_ref_" + name] = mod


def test_plane_and_line_match_the_reference_classes():
    ref = _import_reference("solid_geometry")
    from learningagileflight_se3_b200 import solid_geometry as mine
    rng = np.random.default_rng(3)
    for _ in range(50):
        a, b, c, x, y = (rng.normal(0, 2, 3) for _ in range(5))
        pr, pm = ref.plane(a, b, c), mine.plane(a, b, c)
        for f in ("nor_vec", "n1", "n2", "n3"):
            assert np.allclose(getattr(pr, f)(), getattr(pm, f)(), rtol=0, atol=1e-15)
        assert np.allclose(pr.interpoint(x, y), pm.interpoint(x, y), rtol=1e-12, atol=1e-12)
        lr, lm = ref.line(a, b), mine.line(a, b)
        assert abs(lr.vertical(x) - lm.vertical(x)) <= 1e-14
        for q in (x, a + 0.3 * (b - a) + 0.1 * c, a - 2 * (b - a), b + 3 * (b - a)):      # beside, within, beyond either end
            assert abs(lr.distance(q) - lm.distance(q)) <= 1e-14
    g = np.array([[-0.5, 0, 1], [0.5, 0, 1], [0.5, 0, -1], [-0.5, 0, -1.0]])
    o_r, o_m = ref.obstacle(*g), mine.obstacle(*g)
    for i in range(1, 5):
        assert np.allclose(getattr(o_r, f"plane{i}").nor_vec(), getattr(o_m, f"plane{i}").nor_vec())
        assert np.allclose(getattr(o_r, f"line{i}").dir, getattr(o_m, f"line{i}").dir)
    assert np.allclose(o_r.centroid, o_m.centroid)


def test_model_batch_helper_with_the_reference_network_and_row_models():
    from learningagileflight_se3_b200.learning import _model_batch
    x = np.random.default_rng(0).normal(size=(9, 9))             # B = 9 = D_in: a transposed batch must not be mistaken for a batch
    qn = _import_reference("quad_nn")
    net = qn.network(9, 64, 64, 7)
    with torch.no_grad():
        want = torch.stack([net(x[j]).reshape(7) for j in range(9)])
    got = _model_batch(net, x)
    assert got.shape == (9, 7) and got.requires_grad and float((got.detach() - want).abs().max()) <= 1e-6

    class Rows(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.l = torch.nn.Linear(9, 7)

        def forward(self, i):
            return self.l(torch.as_tensor(i, dtype=torch.float))

    class SingleOnly(Rows):
        def forward(self, i):
            assert np.asarray(i).ndim == 1
            return super().forward(i)

    for m in (Rows(), SingleOnly()):
        with torch.no_grad():
            want = torch.stack([m(x[j]) for j in range(9)])
        assert float((_model_batch(m, x).detach() - want).abs().max()) <= 1e-6


def test_shipped_dnn2_fixture_reproduces_the_reference_module():
    """tests/golden/nn3_1_weights.npz (exported from gym_pybullet_drone/nn3_1.pth by make_golden_nn3.py) re-assembled by
    learning.dnn_from_arrays gives the outputs the reference's own forward produced, batched."""
    from learningagileflight_se3_b200.learning import dnn_from_arrays
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "nn3_1_weights.npz"))
    model = dnn_from_arrays(d)
    with torch.no_grad():
        y = model(d["probe_in"]).numpy()
    assert y.shape == (16, 7) and np.abs(y - d["probe_out"]).max() <= 2e-5 * max(1.0, np.abs(d["probe_out"]).max())
