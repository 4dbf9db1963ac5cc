"""Batched operators over the C ABI (include/se3mpc.h).

Two faces, same kernels:
  * device tensors (torch.float64 on CUDA): `solve`, `reward`, `objective`, `sol_gradient` enqueue on the current
    torch stream and return device tensors; also registered as `torch.ops.se3mpc.*` (CUDA only, no CPU kernel).
  * host buffers (numpy): `solve_host`, `reward_host`, `objective_host`, `sol_gradient_host` -- the call a user of
    the reference makes (numpy in, numpy out), with H2D / D2H inside the library.
PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import Params, Se3MpcError, Stats, default_params  # noqa: F401

_workspaces: dict = {}


def _require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise Se3MpcError("se3mpc device ops need CUDA tensors (there is no CPU fallback)")
        if t.dtype != torch.float64:
            raise Se3MpcError("se3mpc device ops compute in float64; got %s" % t.dtype)


def _c(t: Optional[torch.Tensor], shape):
    if t is None:
        return None
    return t.reshape(shape).contiguous()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _workspace(p: Params, nproblems: int, device) -> torch.Tensor:
    need = int(_lib.lib().se3mpc_workspace_bytes(C.byref(p), int(nproblems)))
    if need == 0:
        raise Se3MpcError("se3mpc_workspace_bytes returned 0 (bad params or no CUDA device)")
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _stats(B, device):
    st = {k: torch.empty(B, dtype=torch.int32, device=device) for k in ("status", "iters", "nfact", "nls")}
    st["kkt_error"] = torch.empty(B, dtype=torch.float64, device=device)
    cs = Stats(st["status"].data_ptr(), st["iters"].data_ptr(), st["nfact"].data_ptr(), st["nls"].data_ptr(), st["kkt_error"].data_ptr())
    return st, cs


def solve(x0, goal, tra_pos, tra_ang, t, ulast=None, round_t=False, params: Optional[Params] = None,
          want_traj=True, want_costate=True, want_state=True, x_init=None, u_init=None):
    """OCSys.ocSolver for a batch (quad_OC.py:104-212).  Returns dict of device tensors.

    tra_ang is [B,3] (the planner's Rodrigues-like vector, converted on the device) or [B,4] (traversal quaternion).
    x_init[B,N+1,13] / u_init[B,N,4]: optional primal starting point (se3mpc_solve_warm_device); default = the reference's cold start."""
    p = params or default_params()
    _require_cuda(x0, goal, tra_pos, tra_ang, t, ulast)
    B, N, dev = x0.reshape(-1, 13).shape[0], p.N, x0.device
    att_dim = 4 if tra_ang.numel() == 4 * B else 3
    x0, goal, tra_pos, tra_ang, t = _c(x0, (B, 13)), _c(goal, (B, 3)), _c(tra_pos, (B, 3)), _c(tra_ang, (B, att_dim)), _c(t, (B,))
    ulast = _c(ulast, (B, 4))
    X = torch.empty((B, N + 1, 13), dtype=torch.float64, device=dev) if (want_traj and want_state) else None
    U = torch.empty((B, N, 4), dtype=torch.float64, device=dev) if want_traj else None
    LAM = torch.empty((B, N, 13), dtype=torch.float64, device=dev) if want_costate else None
    cost = torch.empty(B, dtype=torch.float64, device=dev)
    st, cs = _stats(B, dev)
    _require_cuda(x_init, u_init)
    x_init, u_init = _c(x_init, (B, N + 1, 13)), _c(u_init, (B, N, 4))
    if B == 0:          # empty shard: nothing to launch (the C ABI accepts B = 0 as well), empty outputs
        return {"state_traj_opt": X, "control_traj_opt": U, "costate_traj_opt": LAM, "cost": cost, **st}
    with torch.cuda.device(dev):
        ws = _workspace(p, B, dev)
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if x_init is None and u_init is None:
            rc = _lib.lib().se3mpc_solve_device(C.byref(p), B, _ptr(x0), _ptr(goal), _ptr(tra_pos), _ptr(tra_ang), att_dim, _ptr(t), _ptr(ulast),
                                                int(bool(round_t)), _ptr(X), _ptr(U), _ptr(LAM), _ptr(cost), C.byref(cs), _ptr(ws),
                                                ws.numel(), stream)
        else:
            rc = _lib.lib().se3mpc_solve_warm_device(C.byref(p), B, _ptr(x0), _ptr(goal), _ptr(tra_pos), _ptr(tra_ang), att_dim, _ptr(t),
                                                     _ptr(ulast), int(bool(round_t)), _ptr(x_init), _ptr(u_init), _ptr(X), _ptr(U), _ptr(LAM),
                                                     _ptr(cost), C.byref(cs), _ptr(ws), ws.numel(), stream)
    _lib.check(rc, "se3mpc_solve_device")
    return {"state_traj_opt": X, "control_traj_opt": U, "costate_traj_opt": LAM, "cost": cost, **st}


def reward(X, gate, goal, params: Optional[Params] = None):
    """get_quadrotor_position + 4x collis_det + path + reward (quad_policy.py:78-91) for a batch of trajectories."""
    p = params or default_params()
    _require_cuda(X, gate, goal)
    N = p.N
    B, dev = X.reshape(-1, N + 1, 13).shape[0], X.device
    X, gate, goal = _c(X, (B, N + 1, 13)), _c(gate, (B, 12)), _c(goal, (B, 3))
    r = torch.empty(B, dtype=torch.float64, device=dev)
    col, path = torch.empty_like(r), torch.empty_like(r)
    co = torch.empty(B, dtype=torch.int32, device=dev)
    if B == 0:
        return {"reward": r, "collision": col, "path": path, "co": co}
    with torch.cuda.device(dev):
        rc = _lib.lib().se3mpc_reward_device(C.byref(p), B, _ptr(X), _ptr(gate), _ptr(goal), _ptr(r), _ptr(col), _ptr(path), _ptr(co),
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "se3mpc_reward_device")
    return {"reward": r, "collision": col, "path": path, "co": co}


def objective(x0, goal, gate, tra_pos, tra_ang, t, ulast=None, params: Optional[Params] = None, want_traj=False):
    """run_quad.objective for a batch (quad_policy.py:67-91): t rounded to 0.1, solve, reward -- one launch."""
    p = params or default_params()
    _require_cuda(x0, goal, gate, tra_pos, tra_ang, t, ulast)
    B, N, dev = x0.reshape(-1, 13).shape[0], p.N, x0.device
    x0, goal, gate = _c(x0, (B, 13)), _c(goal, (B, 3)), _c(gate, (B, 12))
    tra_pos, tra_ang, t, ulast = _c(tra_pos, (B, 3)), _c(tra_ang, (B, 3)), _c(t, (B,)), _c(ulast, (B, 4))
    r = torch.empty(B, dtype=torch.float64, device=dev)
    cost = torch.empty_like(r)
    X = torch.empty((B, N + 1, 13), dtype=torch.float64, device=dev) if want_traj else None
    U = torch.empty((B, N, 4), dtype=torch.float64, device=dev) if want_traj else None
    st, cs = _stats(B, dev)
    if B == 0:
        return {"reward": r, "cost": cost, "state_traj_opt": X, "control_traj_opt": U, **st}
    with torch.cuda.device(dev):
        ws = _workspace(p, B, dev)
        rc = _lib.lib().se3mpc_objective_device(C.byref(p), B, _ptr(x0), _ptr(goal), _ptr(gate), _ptr(tra_pos), _ptr(tra_ang), _ptr(t),
                                                _ptr(ulast), _ptr(r), _ptr(X), _ptr(U), _ptr(cost), C.byref(cs), _ptr(ws), ws.numel(),
                                                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "se3mpc_objective_device")
    return {"reward": r, "cost": cost, "state_traj_opt": X, "control_traj_opt": U, **st}


def sol_gradient(x0, goal, gate, tra_pos, tra_ang, t, ulast=None, params: Optional[Params] = None):
    """run_quad.sol_gradient for a batch (quad_policy.py:94-112): 9 solves per sample in one launch -> [B,8]."""
    p = params or default_params()
    _require_cuda(x0, goal, gate, tra_pos, tra_ang, t, ulast)
    B, dev = x0.reshape(-1, 13).shape[0], x0.device
    x0, goal, gate = _c(x0, (B, 13)), _c(goal, (B, 3)), _c(gate, (B, 12))
    tra_pos, tra_ang, t, ulast = _c(tra_pos, (B, 3)), _c(tra_ang, (B, 3)), _c(t, (B,)), _c(ulast, (B, 4))
    out = torch.empty((B, 8), dtype=torch.float64, device=dev)
    st9 = torch.empty((B, 9), dtype=torch.int32, device=dev)
    if B == 0:
        return out, st9
    with torch.cuda.device(dev):
        ws = _workspace(p, 9 * B, dev)
        rc = _lib.lib().se3mpc_sol_gradient_device(C.byref(p), B, _ptr(x0), _ptr(goal), _ptr(gate), _ptr(tra_pos), _ptr(tra_ang), _ptr(t),
                                                   _ptr(ulast), _ptr(out), _ptr(st9), _ptr(ws), ws.numel(),
                                                   C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "se3mpc_sol_gradient_device")
    return out, st9


# ---- host-buffer face (numpy in / numpy out; H2D + kernels + D2H inside libse3mpc.so) ----------------------
def _np(a, shape):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(shape)


def _np_ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _np_stats(B):
    st = {k: np.empty(B, dtype=np.int32) for k in ("status", "iters", "nfact", "nls")}
    st["kkt_error"] = np.empty(B, dtype=np.float64)
    cs = Stats(st["status"].ctypes.data, st["iters"].ctypes.data, st["nfact"].ctypes.data, st["nls"].ctypes.data,
               st["kkt_error"].ctypes.data)
    return st, cs


def solve_host(x0, goal, tra_pos, tra_ang, t, ulast=None, round_t=False, params: Optional[Params] = None,
               want_traj=True, want_costate=True):
    p = params or default_params()
    x0 = _np(x0, (-1, 13))
    B, N = x0.shape[0], p.N
    att_dim = 4 if np.asarray(tra_ang).size == 4 * B else 3
    goal, tra_pos, tra_ang, t, ulast = _np(goal, (B, 3)), _np(tra_pos, (B, 3)), _np(tra_ang, (B, att_dim)), _np(t, (B,)), _np(ulast, (B, 4))
    X = np.empty((B, N + 1, 13)) if want_traj else None
    U = np.empty((B, N, 4)) if want_traj else None
    LAM = np.empty((B, N, 13)) if want_costate else None
    cost = np.empty(B)
    st, cs = _np_stats(B)
    rc = _lib.lib().se3mpc_solve_host(C.byref(p), B, _np_ptr(x0), _np_ptr(goal), _np_ptr(tra_pos), _np_ptr(tra_ang), att_dim, _np_ptr(t),
                                      _np_ptr(ulast), int(bool(round_t)), _np_ptr(X), _np_ptr(U), _np_ptr(LAM), _np_ptr(cost), C.byref(cs))
    _lib.check(rc, "se3mpc_solve_host")
    return {"state_traj_opt": X, "control_traj_opt": U, "costate_traj_opt": LAM, "cost": cost, **st}


def reward_host(X, gate, goal, params: Optional[Params] = None):
    p = params or default_params()
    N = p.N
    X = _np(X, (-1, N + 1, 13))
    B = X.shape[0]
    gate, goal = _np(gate, (B, 12)), _np(goal, (B, 3))
    r, col, path, co = np.empty(B), np.empty(B), np.empty(B), np.empty(B, dtype=np.int32)
    rc = _lib.lib().se3mpc_reward_host(C.byref(p), B, _np_ptr(X), _np_ptr(gate), _np_ptr(goal), _np_ptr(r), _np_ptr(col), _np_ptr(path),
                                       C.c_void_p(co.ctypes.data))
    _lib.check(rc, "se3mpc_reward_host")
    return {"reward": r, "collision": col, "path": path, "co": co}


def collis_det_host(V, gate, horizon):
    """obstacle.collis_det for B polylines V[B,horizon+1,3] (solid_geometry.py:104-168) -> (score[B], co[B])."""
    V = _np(V, (-1, horizon + 1, 3))
    B = V.shape[0]
    gate = _np(gate, (B, 12))
    score, co = np.empty(B), np.empty(B, dtype=np.int32)
    rc = _lib.lib().se3mpc_collis_det_host(B, int(horizon), _np_ptr(V), _np_ptr(gate), _np_ptr(score), C.c_void_p(co.ctypes.data))
    _lib.check(rc, "se3mpc_collis_det_host")
    return score, co


def collis_det(V, gate, horizon):
    """Device-tensor variant of collis_det_host."""
    _require_cuda(V, gate)
    B, dev = V.reshape(-1, horizon + 1, 3).shape[0], V.device
    V, gate = _c(V, (B, horizon + 1, 3)), _c(gate, (B, 12))
    score = torch.empty(B, dtype=torch.float64, device=dev)
    co = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().se3mpc_collis_det_device(B, int(horizon), _ptr(V), _ptr(gate), _ptr(score), _ptr(co),
                                                 C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "se3mpc_collis_det_device")
    return score, co


def objective_host(x0, goal, gate, tra_pos, tra_ang, t, ulast=None, params: Optional[Params] = None, want_traj=False):
    p = params or default_params()
    x0 = _np(x0, (-1, 13))
    B, N = x0.shape[0], p.N
    goal, gate = _np(goal, (B, 3)), _np(gate, (B, 12))
    tra_pos, tra_ang, t, ulast = _np(tra_pos, (B, 3)), _np(tra_ang, (B, 3)), _np(t, (B,)), _np(ulast, (B, 4))
    r, cost = np.empty(B), np.empty(B)
    X = np.empty((B, N + 1, 13)) if want_traj else None
    U = np.empty((B, N, 4)) if want_traj else None
    st, cs = _np_stats(B)
    rc = _lib.lib().se3mpc_objective_host(C.byref(p), B, _np_ptr(x0), _np_ptr(goal), _np_ptr(gate), _np_ptr(tra_pos), _np_ptr(tra_ang),
                                          _np_ptr(t), _np_ptr(ulast), _np_ptr(r), _np_ptr(X), _np_ptr(U), _np_ptr(cost), C.byref(cs))
    _lib.check(rc, "se3mpc_objective_host")
    return {"reward": r, "cost": cost, "state_traj_opt": X, "control_traj_opt": U, **st}


def sol_gradient_host(x0, goal, gate, tra_pos, tra_ang, t, ulast=None, params: Optional[Params] = None):
    p = params or default_params()
    x0 = _np(x0, (-1, 13))
    B = x0.shape[0]
    goal, gate = _np(goal, (B, 3)), _np(gate, (B, 12))
    tra_pos, tra_ang, t, ulast = _np(tra_pos, (B, 3)), _np(tra_ang, (B, 3)), _np(t, (B,)), _np(ulast, (B, 4))
    out = np.empty((B, 8))
    st9 = np.empty((B, 9), dtype=np.int32)
    rc = _lib.lib().se3mpc_sol_gradient_host(C.byref(p), B, _np_ptr(x0), _np_ptr(goal), _np_ptr(gate), _np_ptr(tra_pos), _np_ptr(tra_ang),
                                             _np_ptr(t), _np_ptr(ulast), _np_ptr(out), C.c_void_p(st9.ctypes.data))
    _lib.check(rc, "se3mpc_sol_gradient_host")
    return out, st9


# ---- moving-gate closed loop (main.py:65-116) ------------------------------------------------------------------
def gate_frame(gate, velo, w, t, state, final_point, want_gate=False):
    """DNN-2 inputs for the gate predicted t seconds ahead (main.py:86-94, quad_moving.py:35-42), device tensors.
    gate[B,12], velo[B,3], w[B], t[B], state[B,13], final_point[B,3] -> inputs18[B,18] (and the predicted corners [B,12])."""
    _require_cuda(gate, velo, w, t, state, final_point)
    B, dev = state.reshape(-1, 13).shape[0], state.device
    gate, velo, w, t = _c(gate, (B, 12)), _c(velo, (B, 3)), _c(w, (B,)), _c(t, (B,))
    state, final_point = _c(state, (B, 13)), _c(final_point, (B, 3))
    out = torch.empty(B, 18, dtype=torch.float64, device=dev)
    gp = torch.empty(B, 12, dtype=torch.float64, device=dev) if want_gate else None
    with torch.cuda.device(dev):
        rc = _lib.lib().se3mpc_gate_frame_device(B, _ptr(gate), _ptr(velo), _ptr(w), _ptr(t), _ptr(state), _ptr(final_point), _ptr(out), _ptr(gp),
                                                 C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "se3mpc_gate_frame_device")
    return (out, gp) if want_gate else out


def plant_step(state, u, dt=0.01, nsteps=1, params: Optional[Params] = None):
    """nsteps explicit Euler steps of the plant with the control held (main.py:108); returns a NEW state tensor [B,13]."""
    _require_cuda(state, u)
    p = params or default_params()
    B, dev = state.reshape(-1, 13).shape[0], state.device
    x = state.reshape(B, 13).contiguous().clone()
    u = _c(u, (B, 4))
    with torch.cuda.device(dev):
        rc = _lib.lib().se3mpc_plant_step_device(C.byref(p), B, float(dt), int(nsteps), _ptr(u), _ptr(x),
                                                 C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "se3mpc_plant_step_device")
    return x


def episode_advance(state, gate, u, velo, w, dt=0.01, nsteps=10, noise=None, params: Optional[Params] = None):
    """One MPC tick between two re-solves (main.py:108 + gate.move): nsteps x {plant Euler step, gate rotation + translation} in ONE
    launch.  state[B,13] and gate[B,12] are updated IN PLACE (and returned); noise[nsteps,B,3] optional velocity noise."""
    _require_cuda(state, gate, u, velo, w, noise)
    p = params or default_params()
    B, dev = state.shape[0], state.device
    if not (state.is_contiguous() and gate.is_contiguous() and state.shape == (B, 13) and gate.shape == (B, 12)):
        raise Se3MpcError("episode_advance updates state[B,13] / gate[B,12] in place: pass contiguous tensors of that shape")
    u, velo, w = _c(u, (B, 4)), _c(velo, (B, 3)), _c(w, (B,))
    noise = _c(noise, (int(nsteps), B, 3))
    if B == 0:
        return state, gate
    with torch.cuda.device(dev):
        rc = _lib.lib().se3mpc_episode_advance_device(C.byref(p), B, float(dt), int(nsteps), _ptr(u), _ptr(state), _ptr(gate), _ptr(velo),
                                                      _ptr(w), _ptr(noise), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "se3mpc_episode_advance_device")
    return state, gate


def gate_frame_host(gate, velo, w, t, state, final_point, want_gate=False):
    """Host-buffer variant of gate_frame (numpy in, numpy out)."""
    state = _np(state, (-1, 13))
    B = state.shape[0]
    gate, velo, final_point = _np(gate, (B, 12)), _np(velo, (B, 3)), _np(final_point, (B, 3))
    w = _np(np.broadcast_to(np.asarray(w, dtype=np.float64), (B,)), (B,))
    t = _np(np.broadcast_to(np.asarray(t, dtype=np.float64), (B,)), (B,))
    out = np.empty((B, 18))
    gp = np.empty((B, 12)) if want_gate else None
    rc = _lib.lib().se3mpc_gate_frame_host(B, _np_ptr(gate), _np_ptr(velo), _np_ptr(w), _np_ptr(t), _np_ptr(state), _np_ptr(final_point),
                                           _np_ptr(out), _np_ptr(gp))
    _lib.check(rc, "se3mpc_gate_frame_host")
    return (out, gp) if want_gate else out


def plant_step_host(state, u, dt=0.01, nsteps=1, params: Optional[Params] = None):
    """Host-buffer variant of plant_step."""
    p = params or default_params()
    x = np.array(_np(state, (-1, 13)), copy=True)
    B = x.shape[0]
    u = _np(u, (B, 4))
    rc = _lib.lib().se3mpc_plant_step_host(C.byref(p), B, float(dt), int(nsteps), _np_ptr(u), _np_ptr(x))
    _lib.check(rc, "se3mpc_plant_step_host")
    return x


def rotor_positions_host(X, wing_len):
    """Quadrotor.get_quadrotor_position (quad_model.py:239-276) for any number of states: X[...,13] -> pos[...,15] (numpy)."""
    X = _np(X, (-1, 13))
    pos = np.empty((X.shape[0], 15))
    rc = _lib.lib().se3mpc_rotor_positions_host(X.shape[0], float(wing_len), _np_ptr(X), _np_ptr(pos))
    _lib.check(rc, "se3mpc_rotor_positions_host")
    return pos


# ---- diagnostics (tests): the device code of one Newton step / of the libm-free helpers ------------------------------
def kkt_step(goal, tra_pos, tra_q, t, w, lam, zL, zU, mu, delta_w, obj_factor=1.0, ulast=None, params: Optional[Params] = None,
             dump_stage=-1):
    """se3mpc_kkt_step_device: one interior-point Newton step from the iterates w[B,17N+13] (reference ordering), lam[B,13N],
    zL/zU[B,17N+13].  Returns (ok[B], dw[B,17N+13], lam_plus[B,13N]) and, with dump_stage >= 0, the dense stage matrices
    (A[B,13,13], Bm[B,13,4], c[B,13], H[B,17,17] over (x,u), g[B,17]) as the factor sweep sees them."""
    p = params or default_params()
    _require_cuda(goal, tra_pos, tra_q, t, w, lam, zL, zU, ulast)
    N = p.N
    nw = 17 * N + 13
    B, dev = w.reshape(-1, nw).shape[0], w.device
    goal, tra_pos, tra_q, t = _c(goal, (B, 3)), _c(tra_pos, (B, 3)), _c(tra_q, (B, 4)), _c(t, (B,))
    w, lam, zL, zU, ulast = _c(w, (B, nw)), _c(lam, (B, 13 * N)), _c(zL, (B, nw)), _c(zU, (B, nw)), _c(ulast, (B, 4))
    dw = torch.empty((B, nw), dtype=torch.float64, device=dev)
    lp = torch.empty((B, 13 * N), dtype=torch.float64, device=dev)
    ok = torch.empty(B, dtype=torch.int32, device=dev)
    dump = torch.zeros((B, 16 * 28 + 24 * 24), dtype=torch.float64, device=dev) if dump_stage >= 0 else None
    with torch.cuda.device(dev):
        slot = int(_lib.lib().se3mpc_workspace_bytes(C.byref(p), 1))      # header + one slot: an upper bound of the per-iterate slot
        ws = torch.empty(max(B, 1) * slot, dtype=torch.uint8, device=dev)
        rc = _lib.lib().se3mpc_kkt_step_device(C.byref(p), B, _ptr(goal), _ptr(tra_pos), _ptr(tra_q), _ptr(t), _ptr(ulast), _ptr(w), _ptr(lam),
                                               _ptr(zL), _ptr(zU), float(mu), float(delta_w), float(obj_factor), _ptr(dw), _ptr(lp), _ptr(ok),
                                               int(dump_stage), _ptr(dump), _ptr(ws), ws.numel(),
                                               C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "se3mpc_kkt_step_device")
    if dump is None:
        return ok, dw, lp
    Rb = dump[:, :16 * 28].reshape(B, 16, 28)
    Qa = dump[:, 16 * 28:].reshape(B, 24, 24)
    idx = list(range(13)) + [16, 17, 18, 19]
    H = Qa[:, idx][:, :, idx]
    g = Qa[:, 20][:, idx]
    return ok, dw, lp, {"A": Rb[:, :13, :13], "B": Rb[:, :13, 16:20], "c": Rb[:, :13, 20], "H": H, "g": g, "Qa": Qa, "Rb": Rb}


def device_math(in3):
    """se3mpc_device_math_device: [ddiv(a,b), dsqrt(|a|), round1(a), rd2quat(a,b,c)(4), 0] with the DEVICE variants of the helpers."""
    _require_cuda(in3)
    n, dev = in3.reshape(-1, 3).shape[0], in3.device
    in3 = _c(in3, (n, 3))
    out = torch.empty((n, 8), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().se3mpc_device_math_device(n, _ptr(in3), _ptr(out), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "se3mpc_device_math_device")
    return out


# ---- torch custom ops: torch.ops.se3mpc.{solve,reward,objective,sol_gradient} (CUDA dispatch key only) ------
_registered = False


def register_torch_ops():
    """Thin torch.library registration over the C ABI; default parameters (reference literals). CUDA only."""
    global _registered
    if _registered:
        return
    lib_def = torch.library.Library("se3mpc", "DEF")
    lib_def.define("solve(Tensor x0, Tensor goal, Tensor tra_pos, Tensor tra_ang, Tensor t, Tensor? ulast, bool round_t) -> (Tensor, Tensor, Tensor, Tensor, Tensor)")
    lib_def.define("reward(Tensor X, Tensor gate, Tensor goal) -> Tensor")
    lib_def.define("objective(Tensor x0, Tensor goal, Tensor gate, Tensor tra_pos, Tensor tra_ang, Tensor t, Tensor? ulast) -> (Tensor, Tensor)")
    lib_def.define("sol_gradient(Tensor x0, Tensor goal, Tensor gate, Tensor tra_pos, Tensor tra_ang, Tensor t, Tensor? ulast) -> Tensor")

    def _solve(x0, goal, tra_pos, tra_ang, t, ulast, round_t):
        r = solve(x0, goal, tra_pos, tra_ang, t, ulast, round_t)
        return r["state_traj_opt"], r["control_traj_opt"], r["costate_traj_opt"], r["cost"], r["status"]

    def _reward(X, gate, goal):
        return reward(X, gate, goal)["reward"]

    def _objective(x0, goal, gate, tra_pos, tra_ang, t, ulast):
        r = objective(x0, goal, gate, tra_pos, tra_ang, t, ulast)
        return r["reward"], r["status"]

    def _sol_gradient(x0, goal, gate, tra_pos, tra_ang, t, ulast):
        return sol_gradient(x0, goal, gate, tra_pos, tra_ang, t, ulast)[0]

    lib_def.impl("solve", _solve, "CUDA")
    lib_def.impl("reward", _reward, "CUDA")
    lib_def.impl("objective", _objective, "CUDA")
    lib_def.impl("sol_gradient", _sol_gradient, "CUDA")
    globals()["_torch_lib"] = lib_def
    _registered = True
