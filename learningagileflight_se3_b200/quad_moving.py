"""Moving-gate traversal: the reference's quad_moving.solver (traversal-time fixed point, quad_moving.py:29-57) and the batched
closed loop of main.py:65-116.

`solver(model, quad_state, final_point, gate1, velo, w)` keeps the reference's signature and result (one episode).
`traversal_time(model, state, final, gate, velo, w)` runs the same damped fixed point t1 <- t1 + (t2 - t1)/2 for B episodes at once,
masked per episode (each one stops when ITS |t2 - t1| <= 1e-3, exactly like its own while-loop would); the 18 DNN-2 inputs come
from the se3mpc_gate_frame kernel, the DNN stays in PyTorch.
`run_episodes(...)` is main.py's loop for B episodes: every 10th plant step one batched traversal-time solve, one gate-frame
packing, one DNN-2 forward, one batched run_quad.get_input launch (cold start, Ulast = previous control: main.py:105-106), then
10 plant steps on the device.  No host round trip inside a tick except what the caller's DNN does.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from . import ops
from ._lib import Params, default_params


def _centroid(gate):
    g = gate.reshape(-1, 4, 3)
    return torch.stack([g[:, :, 0].mean(1), g[:, :, 1].mean(1), g[:, :, 2].mean(1)], 1)


@torch.no_grad()
def traversal_time(model: Callable[[torch.Tensor], torch.Tensor], state, final_point, gate, velo, w, tol=1e-3, max_iter=200):
    """Batched quad_moving.solver.  model: [B,18] float64 CUDA tensor -> [B,7] (only column 6 is used).  Returns (t1[B], iterations[B])."""
    B = state.shape[0]
    w = torch.as_tensor(w, dtype=torch.float64, device=state.device).expand(B).contiguous()
    t1 = torch.sqrt(((_centroid(gate) - state[:, 0:3]) ** 2).sum(1)) / 3                     # quad_moving.py:32
    t2 = model(ops.gate_frame(gate, velo, w, t1, state, final_point))[:, 6].to(torch.float64)
    iters = torch.zeros(B, dtype=torch.int32, device=state.device)
    for _ in range(max_iter):
        active = (t2 - t1).abs() > tol
        if not bool(active.any()):
            break
        t1 = torch.where(active, t1 + (t2 - t1) / 2, t1)
        t2n = model(ops.gate_frame(gate, velo, w, t1, state, final_point))[:, 6].to(torch.float64)
        t2 = torch.where(active, t2n, t2)
        iters += active.to(torch.int32)
    return t1, iters


def solver(model, quad_state, final_point, gate1, velo, w):
    """Reference signature (quad_moving.py:29): `model` maps an 18-vector (numpy) to something with `.data.numpy()[6]`."""
    dev = torch.device("cuda", torch.cuda.current_device())

    def batched(x):
        y = model(x[0].cpu().numpy())
        y = y.data.numpy() if hasattr(y, "data") and hasattr(y.data, "numpy") else np.asarray(y)
        return torch.as_tensor(np.asarray(y, dtype=np.float64).reshape(1, -1), device=dev)

    f64 = dict(dtype=torch.float64, device=dev)
    t1, _ = traversal_time(batched, torch.as_tensor(np.asarray(quad_state, float).reshape(1, 13), **f64),
                           torch.as_tensor(np.asarray(final_point, float).reshape(1, 3), **f64),
                           torch.as_tensor(np.asarray(gate1.gate_point, float).reshape(1, 12), **f64),
                           torch.as_tensor(np.asarray(velo, float).reshape(1, 3), **f64), float(w), max_iter=100000)
    return float(t1[0])


@torch.no_grad()
def run_episodes(model, state0, final_point, gate0, velo, w, n_ticks=50, sub_steps=10, dt_sim=0.01, params: Optional[Params] = None,
                 gate_noise: Optional[torch.Tensor] = None, record=False, warm_start=False):
    """Batched main.py:65-116.  state0[B,13], final_point[B,3], gate0[B,12] (corners at time 0), velo[B,3], w (scalar or [B]).
    The gate moves like gate.move (rotation about its y axis by w dt, translation by (velo + noise) dt per plant step); `gate_noise`
    [n_ticks*sub_steps, B, 3] supplies the velocity noise (None: noise-free).  warm_start=False is the reference's behaviour (a fresh
    run_quad per tick, cold start); warm_start=True starts every re-solve from the previous tick's solution shifted by one stage (one
    MPC stage = one tick; BASELINE.json configs[3]).  Returns the final states, the controls of every tick,
    the per-tick solver status and traversal times (and the state trajectory at tick rate when record=True)."""
    p = params or default_params()
    dev = state0.device
    B = state0.shape[0]
    state = state0.clone()
    gate = gate0.reshape(B, 12).clone()
    w = torch.as_tensor(w, dtype=torch.float64, device=dev).expand(B).contiguous()
    u = torch.zeros(B, 4, dtype=torch.float64, device=dev)                                    # main.py:52
    zeros3 = torch.zeros(B, 3, dtype=torch.float64, device=dev)
    U, ST, TT, XS, IT = [], [], [], [], []
    prev = None
    for k in range(n_ticks):
        vel_k = velo if gate_noise is None else velo + gate_noise[k * sub_steps]
        t, _ = traversal_time(model, state, final_point, gate, vel_k, w)                      # main.py:67
        inp = ops.gate_frame(gate, vel_k, w, t, state, final_point)                           # main.py:86-94
        out = model(inp).to(torch.float64)                                                    # main.py:96
        xi = ui = None
        if warm_start and prev is not None:                                                   # previous plan, one stage later
            xi = torch.cat([prev["state_traj_opt"][:, 1:], prev["state_traj_opt"][:, -1:]], 1)
            ui = torch.cat([prev["control_traj_opt"][:, 1:], prev["control_traj_opt"][:, -1:]], 1)
        sol = ops.solve(inp[:, 0:13], inp[:, 13:16], out[:, 0:3], out[:, 3:6], out[:, 6], ulast=u, round_t=False, params=p,
                        want_costate=False, want_state=warm_start, x_init=xi, u_init=ui)      # main.py:105-106 (get_input)
        prev = sol if warm_start else None
        u = sol["control_traj_opt"][:, 0, :].contiguous()
        U.append(u); ST.append(sol["status"]); TT.append(t); IT.append(sol["iters"])
        if record:
            XS.append(state.clone())
        for j in range(sub_steps):                                                            # main.py:108 + gate.move
            state = ops.plant_step(state, u, dt_sim, 1, params=p)
            vel_j = velo if gate_noise is None else velo + gate_noise[k * sub_steps + j]
            # the gate as gate.move advances it: rotate about its centroid by w dt, then translate by vel dt = predicted gate at t = dt
            _, gate = ops.gate_frame(gate, vel_j, w, torch.full((B,), dt_sim, dtype=torch.float64, device=dev), state, final_point, want_gate=True)
    res = {"state": state, "controls": torch.stack(U, 1), "status": torch.stack(ST, 1), "t_tra": torch.stack(TT, 1), "gate": gate,
           "iters": torch.stack(IT, 1)}
    if record:
        res["states"] = torch.stack(XS, 1)
    return res
