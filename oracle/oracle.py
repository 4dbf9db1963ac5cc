"""ctypes face of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  See oracle/se3_oracle.h for what it restates and for the "parity unpinned" statement.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libse3oracle.so")

NX, NU, NMAX = 13, 4, 64


class Params(C.Structure):
    _fields_ = [
        ("Jx", C.c_double), ("Jy", C.c_double), ("Jz", C.c_double), ("mass", C.c_double),
        ("arm_l", C.c_double), ("c_tau", C.c_double), ("grav", C.c_double),
        ("wrt", C.c_double), ("wqt", C.c_double), ("wthrust", C.c_double), ("wrf", C.c_double),
        ("wvf", C.c_double), ("wwf", C.c_double), ("wdu", C.c_double),
        ("gamma_amp", C.c_double), ("gamma_rate", C.c_double), ("dt", C.c_double),
        ("u_lb", C.c_double), ("u_ub", C.c_double), ("w_bound", C.c_double), ("wing_len", C.c_double),
        ("N", C.c_int), ("att_power", C.c_int),
        ("tol", C.c_double), ("acceptable_tol", C.c_double), ("mu_init", C.c_double),
        ("max_iter", C.c_int), ("acceptable_iter", C.c_int), ("max_soc", C.c_int), ("flags", C.c_int),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("x0", C.c_double * 13), ("goal", C.c_double * 3), ("tra_pos", C.c_double * 3),
        ("tra_q", C.c_double * 4), ("t_tra", C.c_double), ("ulast", C.c_double * 4),
    ]


class Solution(C.Structure):
    _fields_ = [
        ("X", C.c_double * ((NMAX + 1) * 13)), ("U", C.c_double * (NMAX * 4)), ("LAM", C.c_double * (NMAX * 13)),
        ("XRAW", C.c_double * ((NMAX + 1) * 13)), ("URAW", C.c_double * (NMAX * 4)),
        ("ZL", C.c_double * (NMAX * 17 + 13)), ("ZU", C.c_double * (NMAX * 17 + 13)),
        ("cost", C.c_double), ("kkt_error", C.c_double), ("mu", C.c_double), ("obj_scale", C.c_double),
        ("dual_inf", C.c_double), ("constr_viol", C.c_double), ("compl_inf", C.c_double),
        ("status", C.c_int), ("iters", C.c_int), ("nfact", C.c_int), ("nls", C.c_int),
        ("nsoc", C.c_int), ("nresto", C.c_int),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/libse3oracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "se3_oracle.c")
    hdr = os.path.join(_HERE, "se3_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libse3oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_native = None          # path of the -march=native build once use_native() switched to it


def _cpu_tag() -> str:
    """Short hash of this host's CPU model + ISA flags: a -march=native build is only valid on the CPU it was compiled on (the
    repo snapshot, built .so files included, travels between the build container and the GPU box)."""
    import hashlib
    model, flags = "", ""
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and not model:
                model = line.split(":", 1)[1].strip()
            if line.startswith("flags") and not flags:
                flags = line.split(":", 1)[1].strip()
            if model and flags:
                break
    except OSError:
        pass
    return hashlib.sha1((model + "|" + flags).encode()).hexdigest()[:10]


def use_native() -> str:
    """Switch to a build of the same source with `-O3 -march=native` for THIS host (the CPU baseline is compiled
    for the cores it is timed on); compiled on first use, tagged with the CPU it belongs to.  Returns the compiler flags in use
    (the portable x86-64-v3 build of the Makefile when the native build is not possible)."""
    global _lib, _native, _LIB_PATH
    if _native:
        return "-O3 -march=native"
    path = os.path.join(_HERE, f"libse3oracle_native_{_cpu_tag()}.so")
    src = os.path.join(_HERE, "se3_oracle.c")
    try:
        if not os.path.exists(path) or os.path.getmtime(src) > os.path.getmtime(path):
            subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-std=gnu11", "-shared", "-o", path, src, "-lm", "-lpthread"],
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except (OSError, subprocess.CalledProcessError):
        return "-O3 -march=x86-64-v3"
    _LIB_PATH, _native, _lib = path, path, None
    return "-O3 -march=native"


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        L.se3o_default_params.argtypes = [C.POINTER(Params)]
        L.se3o_rd2quat.argtypes = [dp, dp]
        L.se3o_round1.argtypes = [C.c_double]
        L.se3o_round1.restype = C.c_double
        L.se3o_dyn.argtypes = [C.POINTER(Params), dp, dp, dp]
        L.se3o_dyn_jac.argtypes = [C.POINTER(Params), dp, dp, dp, dp]
        L.se3o_dyn_hess.argtypes = [C.POINTER(Params), dp, dp, dp, dp]
        L.se3o_stage_cost.argtypes = [C.POINTER(Params), C.POINTER(Problem), C.c_int, dp, dp, dp, dp, dp, dp, dp]
        L.se3o_stage_cost.restype = C.c_double
        L.se3o_final_cost.argtypes = [C.POINTER(Params), C.POINTER(Problem), dp, dp, dp]
        L.se3o_final_cost.restype = C.c_double
        L.se3o_nlp_eval_dense.argtypes = [C.POINTER(Params), C.POINTER(Problem), dp, dp, C.c_double, dp, dp, dp, dp, dp]
        L.se3o_kkt_step.argtypes = [C.POINTER(Params), C.POINTER(Problem), dp, dp, dp, dp, C.c_double, C.c_double,
                                    C.c_double, dp, dp]
        L.se3o_kkt_step.restype = C.c_int
        L.se3o_solve.argtypes = [C.POINTER(Params), C.POINTER(Problem), C.POINTER(Solution)]
        L.se3o_solve.restype = C.c_int
        L.se3o_solve_from.argtypes = [C.POINTER(Params), C.POINTER(Problem), C.c_void_p, C.c_void_p, C.POINTER(Solution)]
        L.se3o_solve_from.restype = C.c_int
        L.se3o_rotor_positions.argtypes = [C.POINTER(Params), dp, C.c_int, dp]
        L.se3o_collis_det.argtypes = [dp, dp, C.c_int, C.c_int, ip]
        L.se3o_collis_det.restype = C.c_double
        L.se3o_reward.argtypes = [C.POINTER(Params), dp, dp, dp, dp, dp, ip]
        L.se3o_reward.restype = C.c_double
        L.se3o_objective.argtypes = [C.POINTER(Params), dp, dp, dp, dp, dp, C.c_double, dp, C.POINTER(Solution)]
        L.se3o_objective.restype = C.c_double
        L.se3o_sol_gradient.argtypes = [C.POINTER(Params), dp, dp, dp, dp, dp, C.c_double, dp, dp, ip]
        L.se3o_solve_batch.argtypes = [C.POINTER(Params), C.c_int, dp, dp, dp, dp, dp, dp, C.c_int, C.c_int,
                                       dp, dp, dp, dp, ip, ip, ip]
        L.se3o_objective_batch.argtypes = [C.POINTER(Params), C.c_int, dp, dp, dp, dp, dp, dp, dp, C.c_int,
                                           dp, dp, ip, ip]
        L.se3o_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def default_params(**overrides) -> Params:
    p = Params()
    lib().se3o_default_params(C.byref(p))
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def rd2quat(tra_ang):
    a = _f64(tra_ang, (3,))
    q = np.zeros(4)
    lib().se3o_rd2quat(_dp(a), _dp(q))
    return q


def round1(t: float) -> float:
    return float(lib().se3o_round1(float(t)))


def make_problem(x0, goal, tra_pos, tra_q, t_tra, ulast=None) -> Problem:
    pb = Problem()
    pb.x0[:] = list(np.asarray(x0, dtype=float).reshape(13))
    pb.goal[:] = list(np.asarray(goal, dtype=float).reshape(3))
    pb.tra_pos[:] = list(np.asarray(tra_pos, dtype=float).reshape(3))
    pb.tra_q[:] = list(np.asarray(tra_q, dtype=float).reshape(4))
    pb.t_tra = float(t_tra)
    pb.ulast[:] = [0.0] * 4 if ulast is None else list(np.asarray(ulast, dtype=float).reshape(4))
    return pb


def dyn(p, x, u):
    x, u = _f64(x, (13,)), _f64(u, (4,))
    out = np.zeros(13)
    lib().se3o_dyn(C.byref(p), _dp(x), _dp(u), _dp(out))
    return out


def dyn_jac(p, x, u):
    x, u = _f64(x, (13,)), _f64(u, (4,))
    A, B = np.zeros((13, 13)), np.zeros((13, 4))
    lib().se3o_dyn_jac(C.byref(p), _dp(x), _dp(u), _dp(A), _dp(B))
    return A, B


def dyn_hess(p, x, u, lam):
    x, u, lam = _f64(x, (13,)), _f64(u, (4,)), _f64(lam, (13,))
    H = np.zeros((17, 17))
    lib().se3o_dyn_hess(C.byref(p), _dp(x), _dp(u), _dp(lam), _dp(H))
    return H


def stage_cost(p, pb, k, x, u, uprev):
    x, u, uprev = _f64(x, (13,)), _f64(u, (4,)), _f64(uprev, (4,))
    gx, gu, gup, H = np.zeros(13), np.zeros(4), np.zeros(4), np.zeros((13, 13))
    v = lib().se3o_stage_cost(C.byref(p), C.byref(pb), int(k), _dp(x), _dp(u), _dp(uprev), _dp(gx), _dp(gu), _dp(gup), _dp(H))
    return v, gx, gu, gup, H


def final_cost(p, pb, x):
    x = _f64(x, (13,))
    gx, H = np.zeros(13), np.zeros((13, 13))
    v = lib().se3o_final_cost(C.byref(p), C.byref(pb), _dp(x), _dp(gx), _dp(H))
    return v, gx, H


def nlp_eval_dense(p, pb, w, lam=None, obj_factor=1.0, want_jac=True, want_hess=True):
    N = p.N
    n, m = 17 * N + 13, 13 * N
    w = _f64(w, (n,))
    lam_ = None if lam is None else _f64(lam, (m,))
    f = C.c_double(0.0)
    grad, g = np.zeros(n), np.zeros(m)
    jac = np.zeros((m, n)) if want_jac else None
    hess = np.zeros((n, n)) if want_hess else None
    lib().se3o_nlp_eval_dense(C.byref(p), C.byref(pb), _dp(w), _dp(lam_), float(obj_factor), C.byref(f),
                              _dp(grad), _dp(g), _dp(jac), _dp(hess))
    return f.value, grad, g, jac, hess


def kkt_step(p, pb, w, lam, zL, zU, mu, delta_w, obj_factor=1.0):
    N = p.N
    n, m = 17 * N + 13, 13 * N
    w, lam, zL, zU = _f64(w, (n,)), _f64(lam, (m,)), _f64(zL, (n,)), _f64(zU, (n,))
    dw, lp = np.zeros(n), np.zeros(m)
    ok = lib().se3o_kkt_step(C.byref(p), C.byref(pb), _dp(w), _dp(lam), _dp(zL), _dp(zU), float(mu), float(delta_w),
                             float(obj_factor), _dp(dw), _dp(lp))
    return bool(ok), dw, lp


def _sol_dict(p, sol: Solution):
    N = p.N
    return {
        "state_traj_opt": np.array(sol.X[: (N + 1) * 13]).reshape(N + 1, 13),
        "control_traj_opt": np.array(sol.U[: N * 4]).reshape(N, 4),
        "costate_traj_opt": np.array(sol.LAM[: N * 13]).reshape(N, 13),
        "x_raw": np.array(sol.XRAW[: (N + 1) * 13]).reshape(N + 1, 13), "u_raw": np.array(sol.URAW[: N * 4]).reshape(N, 4),
        "zL": np.array(sol.ZL[: N * 17 + 13]), "zU": np.array(sol.ZU[: N * 17 + 13]),
        "cost": float(sol.cost), "status": int(sol.status), "iters": int(sol.iters), "nfact": int(sol.nfact),
        "nls": int(sol.nls), "nsoc": int(sol.nsoc), "nresto": int(sol.nresto), "kkt_error": float(sol.kkt_error), "mu": float(sol.mu),
        "obj_scale": float(sol.obj_scale), "dual_inf": float(sol.dual_inf), "constr_viol": float(sol.constr_viol),
        "compl_inf": float(sol.compl_inf),
    }


def solve(p, pb, x_init=None, u_init=None):
    sol = Solution()
    if x_init is None and u_init is None:
        lib().se3o_solve(C.byref(p), C.byref(pb), C.byref(sol))
    else:
        xi = None if x_init is None else _f64(x_init, (p.N + 1, 13))
        ui = None if u_init is None else _f64(u_init, (p.N, 4))
        lib().se3o_solve_from(C.byref(p), C.byref(pb), None if xi is None else xi.ctypes.data, None if ui is None else ui.ctypes.data, C.byref(sol))
    return _sol_dict(p, sol)


def rotor_positions(p, X):
    X = _f64(X).reshape(-1, 13)
    pos = np.zeros((X.shape[0], 15))
    lib().se3o_rotor_positions(C.byref(p), _dp(X), X.shape[0], _dp(pos))
    return pos


def collis_det(gate12, V, horizon):
    g = _f64(gate12, (12,))
    V = _f64(V).reshape(-1, 3)
    co = C.c_int(0)
    v = lib().se3o_collis_det(_dp(g), _dp(V), 3, int(horizon), C.byref(co))
    return float(v), int(co.value)


def reward(p, X, gate12, goal):
    X, g, goal = _f64(X).reshape(-1, 13), _f64(gate12, (12,)), _f64(goal, (3,))
    col, path, co = C.c_double(0), C.c_double(0), C.c_int(0)
    r = lib().se3o_reward(C.byref(p), _dp(X), _dp(g), _dp(goal), C.byref(col), C.byref(path), C.byref(co))
    return float(r), float(col.value), float(path.value), int(co.value)


def objective(p, x0, goal, gate12, tra_pos, tra_ang, t, ulast=None):
    sol = Solution()
    ul = None if ulast is None else _f64(ulast, (4,))
    r = lib().se3o_objective(C.byref(p), _dp(_f64(x0, (13,))), _dp(_f64(goal, (3,))), _dp(_f64(gate12, (12,))),
                             _dp(_f64(tra_pos, (3,))), _dp(_f64(tra_ang, (3,))), float(t), _dp(ul), C.byref(sol))
    return float(r), _sol_dict(p, sol)


def sol_gradient(p, x0, goal, gate12, tra_pos, tra_ang, t, ulast=None):
    out = np.zeros(8)
    st = np.zeros(9, dtype=np.int32)
    ul = None if ulast is None else _f64(ulast, (4,))
    lib().se3o_sol_gradient(C.byref(p), _dp(_f64(x0, (13,))), _dp(_f64(goal, (3,))), _dp(_f64(gate12, (12,))),
                            _dp(_f64(tra_pos, (3,))), _dp(_f64(tra_ang, (3,))), float(t), _dp(ul), _dp(out), _ip(st))
    return out, st


def solve_batch(p, x0, goal, tra_pos, tra_ang, t, ulast=None, round_t=False, nthreads=0, want_traj=True):
    x0 = _f64(x0).reshape(-1, 13)
    B, N = x0.shape[0], p.N
    goal, tra_pos, tra_ang = _f64(goal, (B, 3)), _f64(tra_pos, (B, 3)), _f64(tra_ang, (B, 3))
    t = _f64(t, (B,))
    ul = None if ulast is None else _f64(ulast, (B, 4))
    X = np.zeros((B, N + 1, 13)) if want_traj else None
    U = np.zeros((B, N, 4)) if want_traj else None
    LAM = np.zeros((B, N, 13)) if want_traj else None
    cost = np.zeros(B)
    status, iters, nfact = (np.zeros(B, dtype=np.int32) for _ in range(3))
    lib().se3o_solve_batch(C.byref(p), B, _dp(x0), _dp(goal), _dp(tra_pos), _dp(tra_ang), _dp(t), _dp(ul),
                           int(round_t), int(nthreads), _dp(X), _dp(U), _dp(LAM), _dp(cost), _ip(status), _ip(iters),
                           _ip(nfact))
    return {"X": X, "U": U, "LAM": LAM, "cost": cost, "status": status, "iters": iters, "nfact": nfact}


def objective_batch(p, x0, goal, gate, tra_pos, tra_ang, t, ulast=None, nthreads=0):
    x0 = _f64(x0).reshape(-1, 13)
    B = x0.shape[0]
    goal, gate = _f64(goal, (B, 3)), _f64(gate, (B, 12))
    tra_pos, tra_ang, t = _f64(tra_pos, (B, 3)), _f64(tra_ang, (B, 3)), _f64(t, (B,))
    ul = None if ulast is None else _f64(ulast, (B, 4))
    rew, cost = np.zeros(B), np.zeros(B)
    status, iters = np.zeros(B, dtype=np.int32), np.zeros(B, dtype=np.int32)
    lib().se3o_objective_batch(C.byref(p), B, _dp(x0), _dp(goal), _dp(gate), _dp(tra_pos), _dp(tra_ang), _dp(t), _dp(ul),
                               int(nthreads), _dp(rew), _dp(cost), _ip(status), _ip(iters))
    return {"reward": rew, "cost": cost, "status": status, "iters": iters}


def num_threads() -> int:
    return int(lib().se3o_num_threads())
