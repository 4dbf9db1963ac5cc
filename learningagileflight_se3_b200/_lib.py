"""ctypes loader of libse3mpc.so (the C ABI in include/se3mpc.h).

There is deliberately no CPU implementation behind this module: if the CUDA library is missing, or no CUDA device
is present, compute entry points raise (`Se3MpcError`) instead of falling back to anything.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("SE3MPC_LIB") or os.path.join(CSRC, "libse3mpc.so")   # SE3MPC_LIB: tuning builds only


class Se3MpcError(RuntimeError):
    pass


class Params(C.Structure):
    """se3mpc_params (include/se3mpc.h): every literal of quad_policy.py:37-56 / quad_OC.py:145,150,170."""
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

    def copy(self) -> "Params":
        q = Params()
        C.memmove(C.byref(q), C.byref(self), C.sizeof(Params))
        return q


class Stats(C.Structure):
    _fields_ = [("status", C.c_void_p), ("iters", C.c_void_p), ("nfact", C.c_void_p), ("nls", C.c_void_p),
                ("kkt_error", C.c_void_p)]


EXPORTS = [
    "se3mpc_default_params", "se3mpc_last_error", "se3mpc_device_count", "se3mpc_version", "se3mpc_workspace_bytes",
    "se3mpc_solve_device", "se3mpc_reward_device", "se3mpc_collis_det_device", "se3mpc_objective_device",
    "se3mpc_sol_gradient_device", "se3mpc_solve_host", "se3mpc_reward_host", "se3mpc_collis_det_host",
    "se3mpc_objective_host", "se3mpc_sol_gradient_host", "se3mpc_host_release", "se3mpc_launch_count", "se3mpc_fp64_probe_device",
    "se3mpc_gate_frame_device", "se3mpc_gate_frame_host", "se3mpc_plant_step_device", "se3mpc_plant_step_host", "se3mpc_solve_warm_device", "se3mpc_rotor_positions_device", "se3mpc_rotor_positions_host",
    "se3mpc_kkt_step_device", "se3mpc_device_math_device", "se3mpc_episode_advance_device",
]


def build(force: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... (csrc/Makefile); cross-compiles without a GPU."""
    srcs = [os.path.join(CSRC, f) for f in ("se3mpc.cu", "se3mpc_core.cuh")] + [os.path.join(_HERE, "..", "include", "se3mpc.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(f) > os.path.getmtime(LIB_PATH) for f in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "-B", "libse3mpc.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """Load the CUDA library; raises Se3MpcError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Se3MpcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the se3mpc ops)")
    L = C.CDLL(LIB_PATH)
    vp, ci, pp = C.c_void_p, C.c_int, C.POINTER(Params)
    L.se3mpc_default_params.argtypes = [pp]
    L.se3mpc_last_error.restype = C.c_char_p
    L.se3mpc_device_count.restype = ci
    L.se3mpc_version.restype = C.c_char_p
    L.se3mpc_workspace_bytes.argtypes = [pp, ci]
    L.se3mpc_workspace_bytes.restype = C.c_size_t
    L.se3mpc_launch_count.restype = C.c_longlong
    sp = C.POINTER(Stats)
    L.se3mpc_solve_device.argtypes = [pp, ci, vp, vp, vp, vp, ci, vp, vp, ci, vp, vp, vp, vp, sp, vp, C.c_size_t, vp]
    L.se3mpc_solve_warm_device.argtypes = [pp, ci, vp, vp, vp, vp, ci, vp, vp, ci, vp, vp, vp, vp, vp, vp, sp, vp, C.c_size_t, vp]
    L.se3mpc_solve_warm_device.restype = ci
    L.se3mpc_rotor_positions_device.argtypes = [ci, C.c_double, vp, vp, vp]
    L.se3mpc_rotor_positions_host.argtypes = [ci, C.c_double, vp, vp]
    L.se3mpc_rotor_positions_device.restype = L.se3mpc_rotor_positions_host.restype = ci
    L.se3mpc_reward_device.argtypes = [pp, ci, vp, vp, vp, vp, vp, vp, vp, vp]
    L.se3mpc_fp64_probe_device.argtypes = [ci, vp, vp, C.POINTER(C.c_double)]
    L.se3mpc_fp64_probe_device.restype = ci
    L.se3mpc_collis_det_device.argtypes = [ci, ci, vp, vp, vp, vp, vp]
    L.se3mpc_collis_det_host.argtypes = [ci, ci, vp, vp, vp, vp]
    L.se3mpc_objective_device.argtypes = [pp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sp, vp, C.c_size_t, vp]
    L.se3mpc_sol_gradient_device.argtypes = [pp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, vp]
    L.se3mpc_solve_host.argtypes = [pp, ci, vp, vp, vp, vp, ci, vp, vp, ci, vp, vp, vp, vp, sp]
    L.se3mpc_reward_host.argtypes = [pp, ci, vp, vp, vp, vp, vp, vp, vp]
    L.se3mpc_objective_host.argtypes = [pp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sp]
    L.se3mpc_sol_gradient_host.argtypes = [pp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.se3mpc_gate_frame_device.argtypes = [ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.se3mpc_gate_frame_host.argtypes = [ci, vp, vp, vp, vp, vp, vp, vp, vp]
    L.se3mpc_plant_step_device.argtypes = [pp, ci, C.c_double, ci, vp, vp, vp]
    L.se3mpc_plant_step_host.argtypes = [pp, ci, C.c_double, ci, vp, vp]
    cd = C.c_double
    L.se3mpc_kkt_step_device.argtypes = [pp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, cd, cd, cd, vp, vp, vp, ci, vp, vp, C.c_size_t, vp]
    L.se3mpc_device_math_device.argtypes = [ci, vp, vp, vp]
    L.se3mpc_kkt_step_device.restype = L.se3mpc_device_math_device.restype = ci
    L.se3mpc_episode_advance_device.argtypes = [pp, ci, cd, ci, vp, vp, vp, vp, vp, vp, vp]
    L.se3mpc_episode_advance_device.restype = ci
    for f in ("se3mpc_gate_frame_device", "se3mpc_gate_frame_host", "se3mpc_plant_step_device", "se3mpc_plant_step_host"):
        getattr(L, f).restype = ci
    for f in ("se3mpc_collis_det_device", "se3mpc_collis_det_host", "se3mpc_solve_device", "se3mpc_reward_device", "se3mpc_objective_device", "se3mpc_sol_gradient_device",
              "se3mpc_solve_host", "se3mpc_reward_host", "se3mpc_objective_host", "se3mpc_sol_gradient_host"):
        getattr(L, f).restype = ci
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().se3mpc_last_error().decode(errors="replace")
        raise Se3MpcError(f"{what} failed (code {rc}): {msg}")


def default_params(**overrides) -> Params:
    p = Params()
    lib().se3mpc_default_params(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def launch_count() -> int:
    return int(lib().se3mpc_launch_count())
