"""CPU tier: the oracle's model/NLP evaluation against (a) finite differences, (b) golden vectors produced by the
reference's own quad_model/quad_OC code through tests/golden/make_golden.py."""
import numpy as np
import pytest


def _rand_state(rng):
    x = rng.normal(0, 1, 13)
    x[6:10] = rng.normal(0, 1, 4)
    u = rng.uniform(0.1, 2.3, 4)
    return x, u


def test_dyn_jacobian_matches_finite_differences(oracle):
    O, p, rng = oracle, oracle.default_params(), np.random.default_rng(0)
    for _ in range(5):
        x, u = _rand_state(rng)
        A, B = O.dyn_jac(p, x, u)
        h = 1e-6
        for j in range(13):
            e = np.zeros(13); e[j] = h
            fd = (O.dyn(p, x + e, u) - O.dyn(p, x - e, u)) / (2 * h)
            assert np.allclose(A[:, j], fd, rtol=1e-7, atol=1e-8)
        for j in range(4):
            e = np.zeros(4); e[j] = h
            fd = (O.dyn(p, x, u + e) - O.dyn(p, x, u - e)) / (2 * h)
            assert np.allclose(B[:, j], fd, rtol=1e-7, atol=1e-8)


def test_dyn_hessian_matches_finite_differences(oracle):
    O, p, rng = oracle, oracle.default_params(), np.random.default_rng(1)
    x, u = _rand_state(rng)
    lam = rng.normal(0, 1, 13)
    H = O.dyn_hess(p, x, u, lam)
    assert np.allclose(H, H.T)
    z = np.concatenate([x, u])
    h = 1e-6

    def grad(zv):
        A, B = O.dyn_jac(p, zv[:13], zv[13:])
        return np.concatenate([A.T @ lam, B.T @ lam])
    for j in range(17):
        e = np.zeros(17); e[j] = h
        fd = (grad(z + e) - grad(z - e)) / (2 * h)
        assert np.allclose(H[:, j], fd, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("att_power", [2, 1])
def test_cost_gradient_and_hessian(oracle, att_power):
    O, rng = oracle, np.random.default_rng(2)
    p = O.default_params(att_power=att_power)
    pb = O.make_problem(np.zeros(13), rng.normal(0, 2, 3), rng.normal(0, 1, 3), O.rd2quat(rng.normal(0, 0.4, 3)), 2.3)
    x, u = _rand_state(rng)
    up = rng.uniform(0.1, 2.3, 4)
    k = 22
    v, gx, gu, gup, H = O.stage_cost(p, pb, k, x, u, up)
    h = 1e-6
    for j in range(13):
        e = np.zeros(13); e[j] = h
        vp, gp = O.stage_cost(p, pb, k, x + e, u, up)[:2]
        vm, gm = O.stage_cost(p, pb, k, x - e, u, up)[:2]
        assert abs((vp - vm) / (2 * h) - gx[j]) <= 1e-6 * max(1, abs(gx[j]))
        assert np.allclose((gp - gm) / (2 * h), H[:, j], rtol=1e-6, atol=1e-5)
    h = 1e-3   # the cost is quadratic in u: central differences are exact, a larger step only reduces round-off
    for j in range(4):
        e = np.zeros(4); e[j] = h
        assert abs((O.stage_cost(p, pb, k, x, u + e, up)[0] - O.stage_cost(p, pb, k, x, u - e, up)[0]) / (2 * h) - gu[j]) < 1e-6
        assert abs((O.stage_cost(p, pb, k, x, u, up + e)[0] - O.stage_cost(p, pb, k, x, u, up - e)[0]) / (2 * h) - gup[j]) < 1e-6


def test_round1_matches_python_round(oracle, emu):
    rng = np.random.default_rng(3)
    vals = list(rng.uniform(0, 6, 2000)) + [k / 20 for k in range(0, 121)] + [2.25, 2.35, 2.45, 2.55, 2.65, 0.25, 0.75, 3.05]
    vals += [np.nextafter(v, 10) for v in vals[-60:]] + [np.nextafter(v, -10) for v in vals[-60:]]
    for v in vals:
        v = float(v)
        assert oracle.round1(v) == round(v, 1), v
        assert emu.round1(v) == round(v, 1), v


def _problem_from_case(O, g, i):
    return O.make_problem(g[f"c{i}_x0"], g[f"c{i}_goal"], g[f"c{i}_tra_pos"], O.rd2quat(g[f"c{i}_tra_ang"]),
                          O.round1(float(g[f"c{i}_t"])))


@pytest.mark.parametrize("variant", ["root", "gym"])
def test_nlp_evaluation_matches_reference_golden(oracle, golden_root, golden_gym, variant):
    """f, g, grad f, jac g of the oracle == the values the reference's own NLP (quad_OC.py:114-167) evaluates to."""
    O = oracle
    g = golden_root if variant == "root" else golden_gym
    p = O.default_params() if variant == "root" else O.default_params(att_power=1, u_ub=2.4)
    for i in range(int(g["n_cases"])):
        pb = _problem_from_case(O, g, i)
        # the traversal quaternion the reference builds (Rd2Rp + toQuaternion)
        assert np.allclose(O.rd2quat(g[f"c{i}_tra_ang"]), g[f"c{i}_tra_q_ref"], rtol=0, atol=1e-15)
        w = g[f"c{i}_rand_w"]
        want_d = f"c{i}_rand_grad" in g.files
        f, grad, c, jac, _ = O.nlp_eval_dense(p, pb, w, None, 1.0, want_jac=want_d, want_hess=False)
        assert abs(f - float(g[f"c{i}_rand_f"])) <= 1e-12 * abs(f)
        assert np.allclose(c, g[f"c{i}_rand_g"], rtol=1e-13, atol=1e-12)
        if want_d:
            assert np.allclose(grad, g[f"c{i}_rand_grad"], rtol=1e-12, atol=1e-10)
            J = np.zeros_like(jac)
            J[g[f"c{i}_rand_jac_rows"], g[f"c{i}_rand_jac_cols"]] = g[f"c{i}_rand_jac_vals"]
            assert np.allclose(jac, J, rtol=1e-12, atol=1e-12)
        # initial guess and bounds exactly as the reference passes them to nlpsol (quad_OC.py:125-164)
        w0, lbx, ubx = g[f"c{i}_w0"], g[f"c{i}_lbx"], g[f"c{i}_ubx"]
        N = p.N
        assert np.array_equal(w0[:13], g[f"c{i}_x0"]) and np.array_equal(lbx[:13], ubx[:13])
        for k in range(N):
            assert np.allclose(w0[k * 17 + 13:k * 17 + 17], 0.5 * (p.u_lb + p.u_ub))
            assert np.all(lbx[k * 17 + 13:k * 17 + 17] == p.u_lb) and np.all(ubx[k * 17 + 13:k * 17 + 17] == p.u_ub)
            assert np.all(w0[(k + 1) * 17:(k + 1) * 17 + 13] == 0)
            assert np.allclose(ubx[(k + 1) * 17 + 10:(k + 1) * 17 + 13], p.w_bound)
            assert np.all(ubx[(k + 1) * 17:(k + 1) * 17 + 10] >= 1e19)


@pytest.mark.parametrize("variant", ["root", "gym"])
def test_oracle_solution_matches_certified_golden(oracle, golden_root, golden_gym, variant):
    """Re-solving reproduces the stored solution; the stored solution carries a KKT certificate computed in the
    reference-built NLP (stationarity, feasibility) -- see make_golden.py."""
    O = oracle
    g = golden_root if variant == "root" else golden_gym
    p = O.default_params() if variant == "root" else O.default_params(att_power=1, u_ub=2.4)
    for i in range(int(g["n_cases"])):
        assert float(g[f"c{i}_cert_stat_inf"]) < 1e-6 and float(g[f"c{i}_cert_g_inf"]) < 1e-9
        assert float(g[f"c{i}_cert_raw_bound_viol"]) <= 2.5e-8 and float(g[f"c{i}_cert_bound_viol"]) == 0.0
        sol = O.solve(p, _problem_from_case(O, g, i))
        assert sol["status"] == 0
        assert abs(sol["cost"] - float(g[f"c{i}_oracle_cost"])) <= 1e-9 * abs(sol["cost"])
        assert abs(sol["cost"] - float(g[f"c{i}_cert_f_raw"])) <= 1e-9 * abs(sol["cost"])   # reference objective at the same point
        assert np.abs(sol["state_traj_opt"] - g[f"c{i}_oracle_X"]).max() < 1e-7
        assert np.abs(sol["control_traj_opt"] - g[f"c{i}_oracle_U"]).max() < 1e-7
