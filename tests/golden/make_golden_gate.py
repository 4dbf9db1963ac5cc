#!/usr/bin/env python
"""Generate tests/golden/gate_frame.npz by running the REFERENCE's own quad_model.gate and quad_moving.solver (read-only at
/root/reference; `casadi`/`matplotlib` are stubbed because those modules import them at the top, the code paths used here are
plain numpy + scipy).  Cases: the DNN-2 input packing of main.py:86-94 for random moving gates, and the traversal-time fixed
point of quad_moving.py:29-57 with a fixed small tanh network standing in for nn3_1.pth (weights stored in the fixture).

Usage:  python tests/golden/make_golden_gate.py [--n 64] [--seed 0]
"""
import argparse
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.common import make_gate  # noqa: E402


def install_reference():
    cas = types.ModuleType("casadi")
    cas.numpy = np; cas.np = np; cas.atan = math.atan; cas.pi = math.pi      # names leaked by `from casadi import *`
    sys.modules["casadi"] = cas
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.path.insert(0, "/root/reference")
    import quad_model, quad_moving  # noqa: E401
    return quad_model, quad_moving


class TinyNet:
    """18 -> 16 -> 7 tanh network; output 6 is squashed into (0.5, 3.5) like a traversal time."""
    def __init__(self, rng):
        self.W1 = rng.normal(0, 0.15, (16, 18)); self.b1 = rng.normal(0, 0.1, 16)
        self.W2 = rng.normal(0, 0.3, (7, 16)); self.b2 = rng.normal(0, 0.1, 7)

    def __call__(self, x):
        y = self.W2 @ np.tanh(self.W1 @ np.asarray(x, float) + self.b1) + self.b2
        y[6] = 2.0 + 1.5 * math.tanh(0.5 * y[6])
        return types.SimpleNamespace(data=types.SimpleNamespace(numpy=lambda y=y: y))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    quad_model, quad_moving = install_reference()
    rng = np.random.default_rng(args.seed)
    n = args.n
    width = np.clip(rng.normal(0.9, 0.3, n), 0.5, 1.25)
    pitch = rng.uniform(-1.2, 1.2, n)
    gates = np.stack([make_gate(w_, p_) for w_, p_ in zip(width, pitch)]) + np.tile(rng.uniform(-1, 1, (n, 3)), (1, 4))
    velo = rng.normal(0, 0.6, (n, 3))
    w = rng.uniform(-2, 2, n)
    t = rng.uniform(0.05, 3.0, n)
    state = np.zeros((n, 13))
    state[:, 0:3] = rng.uniform(-5, 5, (n, 3)) + np.array([0, -9, 0])
    state[:, 3:6] = rng.normal(0, 2, (n, 3))
    q = rng.normal(0, 1, (n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[n // 2:] *= rng.uniform(0.97, 1.03, (n - n // 2, 1))             # explicit Euler lets |q| drift: transform() re-normalises
    state[:, 6:10] = q
    state[:, 10:13] = rng.normal(0, 1, (n, 3))
    final = rng.uniform(-2, 2, (n, 3)) + np.array([0, 6, 0])
    inputs = np.zeros((n, 18))
    for i in range(n):
        g = quad_model.gate(gates[i].reshape(4, 3).copy())
        g.translate(t[i] * velo[i])                                     # main.py:86
        g.rotate_y(t[i] * w[i])                                         # main.py:87
        inputs[i, 16] = quad_moving.magni(g.gate_point[0, :] - g.gate_point[1, :])
        inputs[i, 17] = math.atan((g.gate_point[0, 2] - g.gate_point[1, 2]) / (g.gate_point[0, 0] - g.gate_point[1, 0]))
        inputs[i, 0:13] = g.transform(state[i])
        inputs[i, 13:16] = g.t_final(final[i])
    net = TinyNet(rng)
    m = min(n, 24)
    t_fix = np.zeros(m)
    for i in range(m):
        g1 = quad_model.gate(gates[i].reshape(4, 3).copy())
        t_fix[i] = quad_moving.solver(net, state[i], final[i], g1, velo[i], w[i])
    out = os.path.join(HERE, "gate_frame.npz")
    np.savez_compressed(out, gates=gates, velo=velo, w=w, t=t, state=state, final=final, inputs=inputs, t_fix=t_fix,
                        W1=net.W1, b1=net.b1, W2=net.W2, b2=net.b2)
    print("wrote", out, inputs.shape, t_fix[:4])


if __name__ == "__main__":
    main()
