"""CPU tier: (1) the oracle's Riccati Newton step and inertia test against a dense KKT solve / eigen-inertia,
(2) the oracle's reward against the reference's own solid_geometry code (imported when /root/reference is present)
and against the golden rewards produced by the reference's run_quad.objective."""
import os
import sys

import numpy as np
import pytest

from tests.common import flat_w, sample_scenarios

N = 50
n, m = 17 * N + 13, 13 * N


def _bounds(p):
    lb, ub = np.full(n, -np.inf), np.full(n, np.inf)
    for k in range(N):
        lb[k * 17 + 13:k * 17 + 17] = p.u_lb - 1e-8 * max(1, abs(p.u_lb))
        ub[k * 17 + 13:k * 17 + 17] = p.u_ub + 1e-8 * max(1, abs(p.u_ub))
    for k in range(1, N + 1):
        lb[k * 17 + 10:k * 17 + 13] = -p.w_bound * (1 + 1e-8)
        ub[k * 17 + 10:k * 17 + 13] = p.w_bound * (1 + 1e-8)
    return lb, ub


def _random_iterate(O, p, rng, sc):
    """A mid-run-like primal-dual point: an optimal trajectory perturbed, random multipliers."""
    pb = O.make_problem(sc["x0"][0], sc["goal"][0], sc["tra_pos"][0], O.rd2quat(sc["tra_ang"][0]), sc["t"][0])
    sol = O.solve(p, pb)
    w = flat_w(sol["x_raw"], sol["u_raw"])
    w[13:] += rng.normal(0, 0.02, n - 13)
    lb, ub = _bounds(p)
    bounded = np.isfinite(lb)
    w[bounded] = np.clip(w[bounded], lb[bounded] + 1e-3, ub[bounded] - 1e-3)
    lam = sol["costate_traj_opt"].reshape(-1) * sol["obj_scale"] + rng.normal(0, 1, m)
    zL, zU = np.zeros(n), np.zeros(n)
    zL[bounded] = rng.uniform(0.01, 2, bounded.sum())
    zU[bounded] = rng.uniform(0.01, 2, bounded.sum())
    return pb, w, lam, zL, zU, sol["obj_scale"]


@pytest.mark.parametrize("delta_w", [0.0, 1.0, 300.0])
def test_riccati_step_equals_dense_kkt_solve(oracle, delta_w):
    O, p, rng = oracle, oracle.default_params(), np.random.default_rng(11)
    sc = sample_scenarios(1, 21)
    pb, w, lam, zL, zU, sf = _random_iterate(O, p, rng, sc)
    mu = 0.05
    lb, ub = _bounds(p)
    bounded, free = np.isfinite(lb), np.arange(13, n)
    f, grad, c, jac, hess = O.nlp_eval_dense(p, pb, w, lam, sf)
    Sig = np.zeros(n)
    Sig[bounded] = zL[bounded] / (w[bounded] - lb[bounded]) + zU[bounded] / (ub[bounded] - w[bounded])
    H = hess + np.diag(Sig) + delta_w * np.eye(n)
    K = np.block([[H[np.ix_(free, free)], jac[:, free].T], [jac[:, free], np.zeros((m, m))]])
    gphi = sf * grad
    gphi[bounded] += -mu / (w[bounded] - lb[bounded]) + mu / (ub[bounded] - w[bounded])
    ref = np.linalg.solve(K, -np.concatenate([gphi[free], c]))
    ev = np.linalg.eigvalsh(K)
    inertia_ok = (ev > 0).sum() == len(free) and (ev < 0).sum() == m
    ok, dw, lp = O.kkt_step(p, pb, w, lam, zL, zU, mu, delta_w, sf)
    assert ok == inertia_ok
    if ok:
        assert np.abs(dw[free] - ref[:len(free)]).max() <= 1e-8 * max(1.0, np.abs(ref[:len(free)]).max())
        assert np.abs(lp - ref[len(free):]).max() <= 1e-8 * max(1.0, np.abs(ref[len(free):]).max())


def test_riccati_inertia_flag_tracks_dense_inertia(oracle):
    """'every H_uu Cholesky succeeds' <=> dense KKT inertia (n, m, 0), on points where both answers occur."""
    O, p, rng = oracle, oracle.default_params(), np.random.default_rng(12)
    sc = sample_scenarios(1, 22)
    seen = set()
    for trial in range(6):
        pb, w, lam, zL, zU, sf = _random_iterate(O, p, rng, sc)
        lam = lam * (1 + 4 * trial)                 # larger multipliers -> more negative curvature
        lb, ub = _bounds(p)
        bounded, free = np.isfinite(lb), np.arange(13, n)
        _, _, _, jac, hess = O.nlp_eval_dense(p, pb, w, lam, sf)
        Sig = np.zeros(n)
        Sig[bounded] = zL[bounded] / (w[bounded] - lb[bounded]) + zU[bounded] / (ub[bounded] - w[bounded])
        for dw_ in (0.0, 30.0):
            H = hess + np.diag(Sig) + dw_ * np.eye(n)
            K = np.block([[H[np.ix_(free, free)], jac[:, free].T], [jac[:, free], np.zeros((m, m))]])
            ev = np.linalg.eigvalsh(K)
            dense_ok = (ev > 0).sum() == len(free) and (ev < 0).sum() == m
            ok, _, _ = O.kkt_step(p, pb, w, lam, zL, zU, 0.05, dw_, sf)
            assert ok == dense_ok
            seen.add(bool(ok))
    assert seen == {True, False}


# ---- reward -------------------------------------------------------------------------------------------------
def _reference_solid_geometry():
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "solid_geometry.py")):
        return None
    sys.path.insert(0, ref)
    try:
        import solid_geometry as sg   # numpy only: the reference's own code
        return sg
    finally:
        sys.path.remove(ref)


def _random_traj(rng, N=50):
    """A smooth fly-through with un-normalised quaternions (the reference never normalises)."""
    X = np.zeros((N + 1, 13))
    start = rng.uniform(-3, 3, 3) + np.array([0, -6, 0])
    end = rng.uniform(-2, 2, 3) + np.array([0, 6, 0])
    s = np.linspace(0, 1, N + 1)[:, None]
    X[:, 0:3] = start + (end - start) * s + 0.3 * np.sin(3 * s) * rng.normal(0, 1, 3)
    q = rng.normal(0, 0.5, 4); q[0] += 1.5
    X[:, 6:10] = q + 0.5 * s * rng.normal(0, 1, 4)
    return X


def test_reward_matches_reference_solid_geometry(oracle):
    sg = _reference_solid_geometry()
    if sg is None:
        pytest.skip("/root/reference not present (GPU box): covered by the golden rewards below")
    O, p, rng = oracle, oracle.default_params(), np.random.default_rng(5)
    sc = sample_scenarios(60, 7)
    n_cross = n_inside = 0
    for i in range(60):
        X = _random_traj(rng)
        if i % 7 == 0:
            X[:, 1] = -np.abs(X[:, 1]) - 0.5          # never crosses the plane
        if i % 11 == 0:
            X[:, 1] = np.abs(X[:, 1]) + 1.5           # starts behind the plane
        gate, goal = sc["gate"][i], sc["goal"][i]
        pos = O.rotor_positions(p, X)
        obs = sg.obstacle(gate[0:3], gate[3:6], gate[6:9], gate[9:12])
        col, co = 0.0, 0
        for c in range(4):
            v = obs.collis_det(pos[:, 3 * (c + 1):3 * (c + 2)], 50)
            vo, coo = O.collis_det(gate, pos[:, 3 * (c + 1):3 * (c + 2)], 50)
            assert abs(v - vo) <= 1e-12 and obs.co == coo
            col += v; co += obs.co
        path = sum(np.dot(pos[49 - q, 0:3] - goal, pos[49 - q, 0:3] - goal) for q in range(4))
        r, colo, patho, coo = O.reward(p, X, gate, goal)
        assert abs(r - (1000 * col - 0.5 * path + 100)) <= 1e-9 and abs(colo - col) <= 1e-12 and coo == co
        n_cross += col != 0
        n_inside += co > 0
    assert n_cross > 20 and n_inside > 5      # the sample exercises both the outside and the inside branches


@pytest.mark.parametrize("variant", ["root", "gym"])
def test_reward_matches_reference_objective_golden(oracle, golden_root, golden_gym, variant):
    """reward / collision / path / co / rotor-tip trajectory as computed by the reference's run_quad.objective."""
    O = oracle
    g = golden_root if variant == "root" else golden_gym
    p = O.default_params() if variant == "root" else O.default_params(att_power=1, u_ub=2.4)
    for i in range(int(g["n_cases"])):
        X = g[f"c{i}_oracle_X"]
        assert np.abs(O.rotor_positions(p, X) - g[f"c{i}_traj_ref"]).max() <= 1e-12
        r, col, path, co = O.reward(p, X, g[f"c{i}_gate"], g[f"c{i}_goal"])
        assert abs(r - float(g[f"c{i}_reward_ref"])) <= 1e-9
        assert abs(col - float(g[f"c{i}_collision_ref"])) <= 1e-12
        assert abs(path - float(g[f"c{i}_path_ref"])) <= 1e-10
        assert co == int(g[f"c{i}_co_ref"])
