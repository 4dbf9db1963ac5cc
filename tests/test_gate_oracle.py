"""CPU tier: the moving-gate oracle (oracle/gate_oracle.py) against tests/golden/gate_frame.npz, which the reference's own
quad_model.gate / quad_moving.solver produced (tests/golden/make_golden_gate.py)."""
import math
import os

import numpy as np

from oracle import gate_oracle as G

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gate_frame.npz")


def tiny_net(d):
    def f(x):
        y = d["W2"] @ np.tanh(d["W1"] @ np.asarray(x, float) + d["b1"]) + d["b2"]
        y[6] = 2.0 + 1.5 * math.tanh(0.5 * y[6])
        return y
    return f


def test_gate_frame_inputs_match_reference_gate_class():
    d = np.load(GOLD)
    for i in range(d["gates"].shape[0]):
        out = G.gate_frame_inputs(d["gates"][i], d["velo"][i], d["w"][i], d["t"][i], d["state"][i], d["final"][i])
        assert np.abs(out - d["inputs"][i]).max() <= 1e-12, i


def test_traversal_time_fixed_point_matches_reference_solver():
    d = np.load(GOLD)
    net = tiny_net(d)
    for i in range(d["t_fix"].shape[0]):
        t, it = G.traversal_time(net, d["state"][i], d["final"][i], d["gates"][i], d["velo"][i], d["w"][i])
        assert abs(t - d["t_fix"][i]) <= 1e-12, (i, t, d["t_fix"][i])


def test_plant_step_is_the_oracle_dynamics_with_dt_001(oracle):
    # the same Euler map the NLP uses (pinned to the reference's quad_model by tests/test_oracle_model.py), dt = 0.01 (main.py:35)
    O = oracle
    p = O.default_params(dt=0.01)
    rng = np.random.default_rng(3)
    for _ in range(10):
        x = rng.normal(0, 1, 13); u = rng.uniform(0, 2.4, 4)
        assert np.abs(G.plant_steps(x, u, 0.01, 1) - O.dyn(p, x, u)).max() <= 1e-13
