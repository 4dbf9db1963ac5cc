"""CPU tier, SURVEY.md 8(c) items (4)-(5): solver-independent evidence for the restated interior-point method (the reference's
CasADi/IPOPT cannot run offline, so the solver iterations themselves are "parity unpinned").

 (4) an independent SQP (scipy SLSQP) on the same dense NLP -- objective/constraints/bounds exactly as the reference builds them
     (pinned by tests/test_oracle_model.py) -- started from the oracle's solution must not move it and must not find a better
     objective: the point is a local minimiser for a solver that shares no code with the oracle;
 (5) path independence: different barrier starts / scaling / multiplier initialisation reach the same optimum."""
import numpy as np
import pytest
from scipy.optimize import minimize

from tests.common import flat_w, sample_scenarios


def test_independent_sqp_started_at_the_oracle_solution_does_not_move(oracle):
    O = oracle
    p = O.default_params()
    sc = sample_scenarios(1, 77)
    pb = O.make_problem(sc["x0"][0], sc["goal"][0], sc["tra_pos"][0], O.rd2quat(sc["tra_ang"][0]), O.round1(sc["t"][0]))
    sol = O.solve(p, pb)
    assert sol["status"] == 0
    N = p.N
    n = 17 * N + 13
    w0 = flat_w(sol["state_traj_opt"], sol["control_traj_opt"])
    free = np.arange(13, n)                                     # x_0 is fixed by lbw = ubw (quad_OC.py:127-129)
    lb, ub = np.full(n, -np.inf), np.full(n, np.inf)
    for k in range(N):
        lb[k * 17 + 13:k * 17 + 17], ub[k * 17 + 13:k * 17 + 17] = p.u_lb, p.u_ub
        lb[(k + 1) * 17 + 10:(k + 1) * 17 + 13], ub[(k + 1) * 17 + 10:(k + 1) * 17 + 13] = -p.w_bound, p.w_bound

    def full(z):
        w = w0.copy(); w[free] = z
        return w

    fun = lambda z: O.nlp_eval_dense(p, pb, full(z), want_jac=False, want_hess=False)[0]            # noqa: E731
    jac = lambda z: O.nlp_eval_dense(p, pb, full(z), want_jac=False, want_hess=False)[1][free]      # noqa: E731
    con = {"type": "eq", "fun": lambda z: O.nlp_eval_dense(p, pb, full(z), want_hess=False)[2],
           "jac": lambda z: O.nlp_eval_dense(p, pb, full(z), want_hess=False)[3][:, free]}
    r = minimize(fun, w0[free], jac=jac, constraints=[con], bounds=list(zip(lb[free], ub[free])), method="SLSQP",
                 options={"maxiter": 15, "ftol": 1e-12})
    f0 = fun(w0[free])
    assert abs(f0 - sol["cost"]) <= 1e-9 * abs(f0)
    assert r.fun >= f0 - 1e-7 * abs(f0)                        # no better feasible point nearby
    assert np.abs(r.x - w0[free]).max() <= 1e-4                # and the SQP stays where the interior-point method stopped
    assert np.abs(con["fun"](r.x)).max() <= 1e-8


@pytest.mark.parametrize("change", [dict(mu_init=1.0), dict(mu_init=0.3), dict(flags=0)])
def test_path_dependence_of_the_optimum_is_rare_and_counted(oracle, change):
    """The NLP is nonconvex: a different barrier start / no scaling / no least-squares multipliers may end in a DIFFERENT local
    minimum.  Measured: 85-100% of the instances reach the same optimum (cost 1e-9 relative), the rest differ by 3e-5..6e-3
    relative in cost.  This is why north_star asks to COUNT different-minimum solves instead of failing on them; the test pins
    the rate (>= 75% identical) and that every run converges."""
    O = oracle
    sc = sample_scenarios(8, 78)
    same = 0
    for i in range(8):
        pb = O.make_problem(sc["x0"][i], sc["goal"][i], sc["tra_pos"][i], O.rd2quat(sc["tra_ang"][i]), O.round1(sc["t"][i]))
        a = O.solve(O.default_params(), pb)
        b = O.solve(O.default_params(**change), pb)
        assert a["status"] == 0 and b["status"] == 0
        if abs(a["cost"] - b["cost"]) <= 1e-6 * abs(a["cost"]):
            same += 1
            assert np.abs(a["state_traj_opt"] - b["state_traj_opt"]).max() <= 1e-5
            assert np.abs(a["control_traj_opt"] - b["control_traj_opt"]).max() <= 1e-5
    assert same >= 6, same


def test_independent_interior_point_from_the_reference_cold_start(oracle):
    """scipy's trust-constr (a barrier / SQP method that shares no code with the oracle) started from the REFERENCE's cold start
    (x_k = 0, u_k = 1.22: quad_OC.py:142,158) on the same dense NLP.  It lands on the oracle's optimum or, on the instances with
    several nearby local minima, on one that is not better: measured 2.8e-9 and 7e-5 relative cost difference on these two
    instances (the oracle's cost is the lower one in the second)."""
    from scipy.optimize import Bounds, NonlinearConstraint
    from scipy.sparse import csr_matrix
    O = oracle
    p = O.default_params()
    sc = sample_scenarios(2, 77)
    N = p.N
    n = 17 * N + 13
    free = np.arange(13, n)
    lb, ub = np.full(n, -np.inf), np.full(n, np.inf)
    for k in range(N):
        lb[k * 17 + 13:k * 17 + 17], ub[k * 17 + 13:k * 17 + 17] = p.u_lb, p.u_ub
        lb[(k + 1) * 17 + 10:(k + 1) * 17 + 13], ub[(k + 1) * 17 + 10:(k + 1) * 17 + 13] = -p.w_bound, p.w_bound
    close = 0
    for i in range(2):
        pb = O.make_problem(sc["x0"][i], sc["goal"][i], sc["tra_pos"][i], O.rd2quat(sc["tra_ang"][i]), O.round1(sc["t"][i]))
        sol = O.solve(p, pb)
        w0 = np.zeros(n)
        w0[:13] = sc["x0"][i]
        for k in range(N):
            w0[k * 17 + 13:k * 17 + 17] = 0.5 * (p.u_lb + p.u_ub)

        def full(z, w0=w0):
            w = w0.copy(); w[free] = z
            return w

        def ev(z, pb=pb, **kw):
            return O.nlp_eval_dense(p, pb, full(z), **kw)

        con = NonlinearConstraint(lambda z: ev(z, want_hess=False)[2], 0, 0, jac=lambda z: csr_matrix(ev(z, want_hess=False)[3][:, free]),
                                  hess=lambda z, v: ev(z, lam=v, obj_factor=0.0, want_jac=False)[4][np.ix_(free, free)])
        r = minimize(lambda z: ev(z, want_jac=False, want_hess=False)[0], w0[free], jac=lambda z: ev(z, want_jac=False, want_hess=False)[1][free],
                     hess=lambda z: ev(z, want_jac=False)[4][np.ix_(free, free)], constraints=[con], bounds=Bounds(lb[free], ub[free]),
                     method="trust-constr", options={"maxiter": 400, "gtol": 1e-8, "xtol": 1e-12})
        rel = (r.fun - sol["cost"]) / abs(sol["cost"])
        assert np.abs(con.fun(r.x)).max() <= 1e-8
        assert -1e-6 <= rel <= 1e-3, rel                      # never better than the oracle's point, never far from it
        close += abs(rel) <= 1e-6
    assert close >= 1
