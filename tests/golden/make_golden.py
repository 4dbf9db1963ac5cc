#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own Python code (read-only at /root/reference).

What runs unmodified from the reference:
  quad_model.Quadrotor (dynamics + costs), quad_model.gate, quad_model.toQuaternion, quad_nn.nn_sample,
  quad_OC.OCSys.ocSolver (NLP assembly: J, g, bounds, initial guess, solution unpacking),
  quad_policy.run_quad.objective / Rd2Rp (t rounding, reward assembly), solid_geometry.obstacle.collis_det.
What cannot run: casadi.nlpsol('ipopt') (casadi is not installable offline).  tests/golden/fake_casadi.py
stands in for casadi with sympy expressions; its nlpsol hands the reference-built NLP to a backend that
  (1) records w0 / lbx / ubx / lbg / ubg exactly as the reference passes them,
  (2) evaluates the reference's f, g, grad f, jac g at random points (pins the oracle's NLP evaluation),
  (3) obtains a candidate solution from the CPU oracle and CERTIFIES it against the reference NLP
      (objective value, constraint residual, KKT stationarity with the oracle's multipliers),
  (4) returns it so that the reference finishes objective() and produces the reward.
The solver iterations themselves therefore stay "parity unpinned" (see oracle/se3_oracle.h).

Usage:  python tests/golden/make_golden.py [--variant root|gym] [--n 6] [--seed 0]
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import fake_casadi  # noqa: E402
from oracle import oracle as O  # noqa: E402


def install_reference(variant: str):
    sys.modules["casadi"] = fake_casadi
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        sys.modules[n] = types.ModuleType(n)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    ref = "/root/reference"
    if variant == "gym":
        sys.path.insert(0, ref)                                   # quad_OC.py lives only in the root
        sys.path.insert(0, os.path.join(ref, "gym_pybullet_drone"))
    else:
        sys.path.insert(0, ref)
    import quad_model, quad_nn, quad_policy, solid_geometry  # noqa: E401
    return quad_model, quad_nn, quad_policy, solid_geometry


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="root", choices=["root", "gym"])
    ap.add_argument("--n", type=int, default=6)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    quad_model, quad_nn, quad_policy, solid_geometry = install_reference(args.variant)
    if args.variant == "gym":
        params = O.default_params(att_power=1, u_ub=2.4)          # gym_pybullet_drone/quad_model.py:200, quad_policy.py:48
    else:
        params = O.default_params()
    N = params.N
    n, m = 17 * N + 13, 13 * N

    rec = {}

    def backend(nlp, x0, lbx, ubx, lbg, ubg):
        ctx = rec["ctx"]
        w0, lbx, ubx = np.array(x0, float), np.array(lbx, float), np.array(ubx, float)
        out = rec["out"]
        out["w0"], out["lbx"], out["ubx"] = w0, lbx, ubx
        out["lbg"], out["ubg"] = np.array(lbg, float), np.array(ubg, float)
        assert len(nlp.x) == n and len(nlp.g) == m
        pb = O.make_problem(ctx["x0"], ctx["goal"], ctx["tra_pos"], O.rd2quat(ctx["tra_ang"]), O.round1(ctx["t"]),
                            ctx.get("ulast"))
        # (2) NLP evaluation at random points inside the bounds (pins oracle vs reference NLP)
        rng = np.random.default_rng(1234 + ctx["idx"])
        wr = rng.normal(0, 1.0, n)
        wr[:13] = w0[:13]
        for k in range(N):
            wr[k * 17 + 13:k * 17 + 17] = rng.uniform(0.1, 2.3, 4)
        f_ref, g_ref = nlp.eval_fg(wr)
        out["rand_w"], out["rand_f"], out["rand_g"] = wr, f_ref, g_ref
        if ctx["idx"] < 2:
            out["rand_grad"] = nlp.eval_grad(wr)
            J = nlp.eval_jac(wr)
            r_, c_ = np.nonzero(J)
            out["rand_jac_rows"], out["rand_jac_cols"], out["rand_jac_vals"] = r_.astype(np.int32), c_.astype(np.int32), J[r_, c_]
        # (3) candidate from the oracle, certified against the reference NLP
        sol = O.solve(params, pb)
        X, U, LAM = sol["state_traj_opt"], sol["control_traj_opt"], sol["costate_traj_opt"]
        w = np.zeros(n)
        for k in range(N):
            w[k * 17:k * 17 + 13] = X[k]
            w[k * 17 + 13:k * 17 + 17] = U[k]
        w[N * 17:] = X[N]
        wraw = np.zeros(n)                      # the iterate before IPOPT-style projection onto the original box
        for k in range(N):
            wraw[k * 17:k * 17 + 13] = sol["x_raw"][k]
            wraw[k * 17 + 13:k * 17 + 17] = sol["u_raw"][k]
        wraw[N * 17:] = sol["x_raw"][N]
        f_at, _ = nlp.eval_fg(w)
        f_raw, g_at = nlp.eval_fg(wraw)
        grad = nlp.eval_grad(wraw)
        J = nlp.eval_jac(wraw)
        sf = sol["obj_scale"]
        stat = grad + J.T @ LAM.reshape(-1) - sol["zL"] / sf + sol["zU"] / sf
        out["cert_f_ref"] = f_at
        out["cert_g_inf"] = float(np.abs(g_at).max())
        out["cert_stat_inf"] = float(np.abs(stat[13:]).max())
        out["cert_bound_viol"] = float(max((lbx - w).max(), (w - ubx).max(), 0.0))
        out["cert_raw_bound_viol"] = float(max((lbx - wraw).max(), (wraw - ubx).max(), 0.0))
        out["oracle_cost"] = sol["cost"]
        out["oracle_X"], out["oracle_U"], out["oracle_LAM"] = X, U, LAM
        out["cert_f_raw"] = f_raw
        out["oracle_status"], out["oracle_iters"], out["oracle_nfact"] = sol["status"], sol["iters"], sol["nfact"]
        out["oracle_kkt_error"], out["obj_scale"] = sol["kkt_error"], sf
        print(f"  [{ctx['idx']}] status={sol['status']} iters={sol['iters']} cost={sol['cost']:.9f} f_ref(w*)={f_at:.9f} "
              f"|g|={out['cert_g_inf']:.2e} |stat|={out['cert_stat_inf']:.2e}", flush=True)
        return {"x": w, "f": f_at, "lam_g": LAM.reshape(-1)}

    fake_casadi.NLP_BACKEND = backend
    np.random.seed(args.seed)
    rng = np.random.default_rng(args.seed + 100)
    cases = []
    for idx in range(args.n):
        inputs = quad_nn.nn_sample()                                               # quad_nn.py:18-48
        gate_point = np.array([[-inputs[7] / 2, 0, 1], [inputs[7] / 2, 0, 1], [inputs[7] / 2, 0, -1], [-inputs[7] / 2, 0, -1]])
        gate1 = quad_model.gate(gate_point)
        gate_point = gate1.rotate_y_out(inputs[8])                                  # deep_learning.py:25-27
        quad1 = quad_policy.run_quad(goal_pos=inputs[3:6], ini_r=inputs[0:3].tolist(),
                                     ini_q=quad_model.toQuaternion(inputs[6], [0, 0, 1]))   # deep_learning.py:29
        quad1.init_obstacle(gate_point.reshape(12))
        tra_pos = rng.normal(0, 0.2, 3)
        tra_ang = rng.normal(0, 0.2, 3)
        t = float(np.clip(round(solid_geometry.magni(inputs[0:3]) / 4, 1), 2, 4)) + (0.03 if idx % 2 else 0.0)
        rec["ctx"] = dict(idx=idx, x0=np.array(quad1.ini_state, float), goal=np.array(inputs[3:6], float),
                          tra_pos=tra_pos, tra_ang=tra_ang, t=t)
        rec["out"] = {}
        reward = quad1.objective(quad1.ini_state, tra_pos, tra_ang, t)              # quad_policy.py:67-91 (reference code)
        out = rec["out"]
        out.update(dict(nn_inputs=inputs, x0=rec["ctx"]["x0"], goal=rec["ctx"]["goal"], gate=gate_point.reshape(12),
                        tra_pos=tra_pos, tra_ang=tra_ang, t=t, reward_ref=float(reward),
                        collision_ref=float(quad1.collision), path_ref=float(quad1.path), co_ref=int(quad1.co),
                        traj_ref=np.array(quad1.traj, float),
                        tra_q_ref=np.array(quad_model.toQuaternion(*quad_policy.Rd2Rp(tra_ang)), float)))
        cases.append(out)
        print(f"case {idx}: reward_ref={reward:.9f} collision={quad1.collision:.6g} path={quad1.path:.6g} co={quad1.co}", flush=True)

    # also keep the reference's state/control/costate unpacking (quad_OC.py:175-210) for the last case
    flat = {}
    for i, c in enumerate(cases):
        for k, v in c.items():
            flat[f"c{i}_{k}"] = np.asarray(v)
    flat["n_cases"] = np.array(len(cases))
    flat["variant"] = np.array(args.variant)
    dst = os.path.join(HERE, f"nlp_{args.variant}.npz")
    np.savez_compressed(dst, **flat)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
