"""GPU tier: the DEVICE code of the derivative tables and of the Riccati Newton step, tested directly (first row).

The CPU tier runs the kernel source as a host emulation whose `ddiv` / `dsqrt` / `rd2quat` / `uni` are libm alternates; the device
variants of those helpers and the table-driven A / B / Hessian records are only exercised on a GPU.  `se3mpc_kkt_step_device` runs
prepare_stages + riccati_factor (DMMA tiles) + riccati_forward on caller-supplied iterates and can dump the dense stage matrices
exactly as the factor sweep sees them; both are compared with the oracle's dense NLP derivatives (pinned to the reference's own NLP by
tests/golden) and its dense 1500 x 1500 KKT solve."""
import numpy as np
import pytest
import torch

from tests.common import flat_w, sample_scenarios

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


@pytest.fixture(scope="module")
def se3():
    import learningagileflight_se3_b200 as m
    assert m._lib.lib().se3mpc_device_count() >= 1
    return m


def _random_iterates(oracle, p, B, seed, ulast=None):
    """Interior iterates around oracle solutions: (problems, w, lam, zL, zU, obj_scale)."""
    rng = np.random.default_rng(seed)
    sc = sample_scenarios(B, seed)
    N = p.N
    n, m = 17 * N + 13, 13 * N
    out = {"pb": [], "w": np.zeros((B, n)), "lam": np.zeros((B, m)), "zL": np.zeros((B, n)), "zU": np.zeros((B, n)), "sf": np.zeros(B),
           "tq": np.zeros((B, 4)), "t": np.zeros(B), "sc": sc}
    for i in range(B):
        tq = oracle.rd2quat(sc["tra_ang"][i])
        t = oracle.round1(sc["t"][i]) + (0.03 if i % 2 else 0.0)
        pb = oracle.make_problem(sc["x0"][i], sc["goal"][i], sc["tra_pos"][i], tq, t, ulast)
        sol = oracle.solve(p, pb)
        w = flat_w(sol["x_raw"], sol["u_raw"])
        w[13:] += rng.normal(0, 0.02, n - 13)
        for k in range(N):
            w[k * 17 + 13:k * 17 + 17] = np.clip(w[k * 17 + 13:k * 17 + 17], 0.01, p.u_ub - 0.01)
            w[(k + 1) * 17 + 10:(k + 1) * 17 + 13] = np.clip(w[(k + 1) * 17 + 10:(k + 1) * 17 + 13], -1.5, 1.5)
        out["pb"].append(pb); out["w"][i] = w
        out["lam"][i] = sol["costate_traj_opt"].reshape(-1) * sol["obj_scale"] + rng.normal(0, 0.5, m)
        out["zL"][i], out["zU"][i] = rng.uniform(0.05, 1.5, n), rng.uniform(0.05, 1.5, n)
        out["sf"][i], out["tq"][i], out["t"][i] = sol["obj_scale"], tq, t
    return out


@pytest.mark.parametrize("variant", ["root", "gym_ulast"])
@pytest.mark.parametrize("delta_w", [0.0, 50.0])
def test_device_newton_step_matches_dense_kkt_solve(oracle, se3, variant, delta_w):
    ulast = None if variant == "root" else np.array([1.0, 1.3, 0.9, 1.1])
    kw = {} if variant == "root" else {"att_power": 1, "u_ub": 2.4}
    p, pg = oracle.default_params(**kw), se3.default_params(**kw)
    B, mu = 6, 0.02
    it = _random_iterates(oracle, p, B, 41, ulast)
    sc = it["sc"]
    for sf in sorted(set(np.round(it["sf"], 12))):            # obj_factor is one scalar per launch
        sel = np.where(np.round(it["sf"], 12) == sf)[0]
        ok, dw, lp = se3.ops.kkt_step(_t(sc["goal"][sel]), _t(sc["tra_pos"][sel]), _t(it["tq"][sel]), _t(it["t"][sel]), _t(it["w"][sel]),
                                      _t(it["lam"][sel]), _t(it["zL"][sel]), _t(it["zU"][sel]), mu, delta_w, float(it["sf"][sel[0]]),
                                      ulast=None if ulast is None else _t(np.tile(ulast, (len(sel), 1))), params=pg)
        torch.cuda.synchronize()
        ok, dw, lp = ok.cpu().numpy(), dw.cpu().numpy(), lp.cpu().numpy()
        for j, i in enumerate(sel):
            oko, dwo, lpo = oracle.kkt_step(p, it["pb"][i], it["w"][i], it["lam"][i], it["zL"][i], it["zU"][i], mu, delta_w, it["sf"][i])
            assert bool(ok[j]) == oko
            if oko:
                assert np.abs(dw[j] - dwo).max() <= 1e-9 * max(1.0, np.abs(dwo).max())
                assert np.abs(lp[j] - lpo).max() <= 1e-9 * max(1.0, np.abs(lpo).max())


@pytest.mark.parametrize("variant", ["root", "gym_ulast"])
def test_device_stage_matrices_match_oracle_derivatives(oracle, se3, variant):
    """A = d(x + dt f)/dx, B = d(x + dt f)/du, c, the Lagrangian Hessian blocks and the barrier-objective gradient of one stage, read back
    from the shared-memory operands of the factor sweep, against the oracle's dense Jacobian / Hessian of the reference-pinned NLP."""
    ulast = None if variant == "root" else np.array([1.0, 1.3, 0.9, 1.1])
    kw = {} if variant == "root" else {"att_power": 1, "u_ub": 2.4}
    p, pg = oracle.default_params(**kw), se3.default_params(**kw)
    B, mu, delta_w, N = 4, 0.03, 0.7, p.N
    it = _random_iterates(oracle, p, B, 43, ulast)
    sc = it["sc"]
    sf = 0.37                                                    # any objective scaling: applied to the whole batch
    wL, wU = -p.w_bound * (1 + 1e-8), p.w_bound * (1 + 1e-8)     # bound_relax_factor 1e-8 (w_bound > 1)
    uL, uU = p.u_lb - 1e-8, p.u_ub + 1e-8 * max(1.0, p.u_ub)
    for k in (0, 1, 17, N - 1):
        ok, dw, lp, d = se3.ops.kkt_step(_t(sc["goal"]), _t(sc["tra_pos"]), _t(it["tq"]), _t(it["t"]), _t(it["w"]), _t(it["lam"]),
                                         _t(it["zL"]), _t(it["zU"]), mu, delta_w, sf,
                                         ulast=None if ulast is None else _t(np.tile(ulast, (B, 1))), params=pg, dump_stage=k)
        torch.cuda.synchronize()
        d = {key: v.cpu().numpy() for key, v in d.items()}
        for i in range(B):
            w, lam, zL, zU = it["w"][i], it["lam"][i], it["zL"][i], it["zU"][i]
            f, grad, g, jac, hess = oracle.nlp_eval_dense(p, it["pb"][i], w, lam, sf, want_jac=True, want_hess=True)
            xs, us = slice(17 * k, 17 * k + 13), slice(17 * k + 13, 17 * k + 17)
            A, Bm = jac[13 * k:13 * k + 13, xs], jac[13 * k:13 * k + 13, us]
            assert np.abs(d["A"][i] - A).max() <= 1e-14 * max(1.0, np.abs(A).max())
            assert np.abs(d["B"][i] - Bm).max() <= 1e-14 * max(1.0, np.abs(Bm).max())
            assert np.abs(d["c"][i] - g[13 * k:13 * k + 13]).max() <= 1e-13 * max(1.0, np.abs(g).max())
            # Hessian of the Lagrangian over (x_k, u_k) + Sigma + delta_w; the u block in the augmented-state form of the
            # u_k - u_{k-1} coupling: R_k = sf (2 wth + 2 wdu) I + Sigma_u + delta_w I
            x, u = w[xs], w[us]
            Hxx = hess[xs, xs].copy()
            sig_w = np.zeros(13)
            bar_w = np.zeros(13)
            if k >= 1:
                for j in range(3):
                    v = x[10 + j]
                    sig_w[10 + j] = zL[17 * k + 10 + j] / (v - wL) + zU[17 * k + 10 + j] / (wU - v)
                    bar_w[10 + j] = mu * (1.0 / (wU - v) - 1.0 / (v - wL))
            Hxx_exp = Hxx + np.diag(sig_w) + delta_w * np.eye(13)
            if k == 0:
                Hxx_exp = None                                     # x_0 is fixed: its Hessian block is never used (dx_0 = 0)
            Hdev = d["H"][i]
            scale = max(1.0, np.abs(hess[xs, xs]).max())
            if Hxx_exp is not None:
                assert np.abs(Hdev[:13, :13] - Hxx_exp).max() <= 1e-13 * scale
            assert np.abs(Hdev[13:, :13] - hess[us, xs]).max() <= 1e-13 * scale
            sig_u = zL[us] / (u - uL) + zU[us] / (uU - u)
            Ruu = (sf * (2 * p.wthrust + 2 * p.wdu) + delta_w) * np.eye(4) + np.diag(sig_u)
            assert np.abs(Hdev[13:, 13:] - Ruu).max() <= 1e-13 * max(1.0, np.abs(Ruu).max())
            # gradient of the barrier objective: x part (stage cost gradient is separable in x_k), u part in the augmented form
            _, gxk, _, _, _ = oracle.stage_cost(p, it["pb"][i], k, x, u, np.zeros(4))
            gx_exp = sf * gxk + bar_w
            assert np.abs(d["g"][i][:13] - gx_exp).max() <= 1e-12 * max(1.0, np.abs(gx_exp).max())
            up = w[17 * (k - 1) + 13:17 * (k - 1) + 17] if k >= 1 else (np.zeros(4) if ulast is None else ulast)
            gu_exp = sf * (2 * p.wthrust * u + 2 * p.wdu * (u - up)) + mu * (1.0 / (uU - u) - 1.0 / (u - uL))
            assert np.abs(d["g"][i][13:] - gu_exp).max() <= 1e-12 * max(1.0, np.abs(gu_exp).max())


def test_device_math_helpers(oracle, se3):
    """ddiv / dsqrt (inline Newton sequences without the libm slow path), round1 and the libm-free rd2quat ON THE DEVICE."""
    rng = np.random.default_rng(11)
    n = 4096
    a = np.concatenate([rng.normal(0, 1, n // 2), rng.uniform(1.5, 4.5, n // 4), 10.0 ** rng.uniform(-12, 12, n // 4)])
    b = np.concatenate([rng.normal(0, 1, n // 2) + 3.0 * np.sign(rng.normal(size=n // 2)), 10.0 ** rng.uniform(-10, 10, n // 2)])
    c = rng.normal(0, 0.5, n)
    got = se3.ops.device_math(_t(np.stack([a, b, c], 1))).cpu().numpy()
    assert np.all(np.abs(got[:, 0] - a / b) <= 2.3e-16 * np.abs(a / b))             # within one ulp of the correctly rounded quotient
    assert np.all(np.abs(got[:, 1] - np.sqrt(np.abs(a))) <= 2.3e-16 * np.sqrt(np.abs(a)))
    r1 = np.array([oracle.round1(float(v)) for v in a[:n // 2 + n // 4]])
    assert np.array_equal(got[:n // 2 + n // 4, 2], r1)
    adversarial = np.array([0.05, 0.15, 0.25, 0.35, 2.45, 2.55, 3.05, 1.0000000000000002e-1 * 25, 2.675, 0.1 * 3])
    ga = se3.ops.device_math(_t(np.stack([adversarial, np.ones(10), np.zeros(10)], 1))).cpu().numpy()[:, 2]
    assert np.array_equal(ga, np.array([round(float(v), 1) for v in adversarial]))
    small = np.stack([a, b, c], 1)[:n // 2] * 0.3
    gq = se3.ops.device_math(_t(small)).cpu().numpy()[:, 3:7]
    eq = np.stack([oracle.rd2quat(v) for v in small])
    assert np.abs(gq - eq).max() <= 4e-16
