"""ctypes face of the host warp-emulation of the CUDA solver (tests only; see se3mpc_emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libse3mpc_emu.so")


class Params(C.Structure):
    """Mirror of se3mpc_params (include/se3mpc.h)."""
    _fields_ = [
        ("Jx", C.c_double), ("Jy", C.c_double), ("Jz", C.c_double), ("mass", C.c_double),
        ("arm_l", C.c_double), ("c_tau", C.c_double), ("grav", C.c_double),
        ("wrt", C.c_double), ("wqt", C.c_double), ("wthrust", C.c_double), ("wrf", C.c_double),
        ("wvf", C.c_double), ("wwf", C.c_double), ("wdu", C.c_double),
        ("gamma_amp", C.c_double), ("gamma_rate", C.c_double), ("dt", C.c_double),
        ("u_lb", C.c_double), ("u_ub", C.c_double), ("w_bound", C.c_double), ("wing_len", C.c_double),
        ("N", C.c_int32), ("att_power", C.c_int32),
        ("tol", C.c_double), ("acceptable_tol", C.c_double), ("mu_init", C.c_double),
        ("max_iter", C.c_int32), ("acceptable_iter", C.c_int32), ("max_soc", C.c_int32), ("flags", C.c_int32),
    ]


def build(force=False):
    srcs = [os.path.join(_HERE, "se3mpc_emu.cpp"),
            os.path.join(_HERE, "../../learningagileflight_se3_b200/csrc/se3mpc_core.cuh"),
            os.path.join(_HERE, "../../include/se3mpc.h")]
    stale = (not os.path.exists(_LIB)) or any(os.path.getmtime(f) > os.path.getmtime(_LIB) for f in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.se3emu_solve.argtypes = [C.POINTER(Params), dp, dp, dp, dp, C.c_double, dp, C.c_int, dp, dp, dp, dp, ip, dp]
        L.se3emu_reward.argtypes = [C.POINTER(Params), dp, dp, dp, dp, dp, ip]
        L.se3emu_reward.restype = C.c_double
        L.se3emu_round1.argtypes = [C.c_double]
        L.se3emu_round1.restype = C.c_double
        L.se3emu_rd2quat.argtypes = [dp, dp]
        L.se3emu_smem_bytes.restype = C.c_size_t
        L.se3emu_slot_bytes.restype = C.c_size_t
        L.se3emu_kkt_step.argtypes = [C.POINTER(Params), dp, dp, dp, C.c_double, dp, dp, dp, dp, dp, C.c_double,
                                      C.c_double, C.c_double, dp, dp]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f(a, shape):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(shape)


def params_from(src) -> Params:
    """Copy any ctypes Params-like structure (e.g. the oracle's) field by field."""
    p = Params()
    for name, _ in Params._fields_:
        setattr(p, name, getattr(src, name))
    return p


def solve(p, x0, goal, tra_pos, tra_ang, t, ulast=None, round_t=False):
    N = p.N
    X, U, LAM = np.zeros((N + 1, 13)), np.zeros((N, 4)), np.zeros((N, 13))
    cost, kkt = C.c_double(0), C.c_double(0)
    st = (C.c_int * 4)()
    ul = None if ulast is None else _f(ulast, (4,))
    lib().se3emu_solve(C.byref(p), _dp(_f(x0, (13,))), _dp(_f(goal, (3,))), _dp(_f(tra_pos, (3,))), _dp(_f(tra_ang, (3,))),
                       float(t), _dp(ul), int(round_t), _dp(X), _dp(U), _dp(LAM), C.byref(cost), st, C.byref(kkt))
    return {"state_traj_opt": X, "control_traj_opt": U, "costate_traj_opt": LAM, "cost": cost.value,
            "status": st[0], "iters": st[1], "nfact": st[2], "nls": st[3], "kkt_error": kkt.value}


def reward(p, X, gate, goal):
    X = _f(X, (-1, 13))
    col, path, co = C.c_double(0), C.c_double(0), C.c_int(0)
    r = lib().se3emu_reward(C.byref(p), _dp(X), _dp(_f(gate, (12,))), _dp(_f(goal, (3,))), C.byref(col), C.byref(path), C.byref(co))
    return float(r), col.value, path.value, co.value


def round1(t):
    return float(lib().se3emu_round1(float(t)))


def rd2quat(a):
    q = np.zeros(4)
    lib().se3emu_rd2quat(_dp(_f(a, (3,))), _dp(q))
    return q


def kkt_step(p, goal, tra_pos, tra_q, t_tra, ulast, w, lam, zL, zU, mu, delta_w, obj_factor=1.0):
    N = p.N
    n, m = 17 * N + 13, 13 * N
    dw, lp = np.zeros(n), np.zeros(m)
    ul = None if ulast is None else _f(ulast, (4,))
    ok = lib().se3emu_kkt_step(C.byref(p), _dp(_f(goal, (3,))), _dp(_f(tra_pos, (3,))), _dp(_f(tra_q, (4,))), float(t_tra), _dp(ul),
                               _dp(_f(w, (n,))), _dp(_f(lam, (m,))), _dp(_f(zL, (n,))), _dp(_f(zU, (n,))), float(mu), float(delta_w),
                               float(obj_factor), _dp(dw), _dp(lp))
    return bool(ok), dw, lp
