"""Moving-gate traversal: the reference's quad_moving.solver (traversal-time fixed point, quad_moving.py:29-57) and the batched
closed loop of main.py:65-116.

`solver(model, quad_state, final_point, gate1, velo, w)` keeps the reference's signature and result (one episode).
`traversal_time(model, state, final, gate, velo, w)` runs the same damped fixed point t1 <- t1 + (t2 - t1)/2 for B episodes at once,
masked per episode (each one stops when ITS |t2 - t1| <= 1e-3, exactly like its own while-loop would); the 18 DNN-2 inputs come
from the se3mpc_gate_frame kernel, the DNN stays in PyTorch.
`run_episodes(...)` is main.py's loop for B episodes: every 10th plant step one batched traversal-time solve, one gate-frame
packing, one DNN-2 forward, one batched run_quad.get_input launch (cold start, Ulast = previous control: main.py:105-106), then
ONE fused launch for the 10 plant + gate sub-steps.  The traversal-time fixed point runs a fixed number of masked iterations inside
the loop, so a tick contains no device->host synchronisation at all; independent groups of episodes can run on separate streams.
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
def traversal_time(model: Callable[[torch.Tensor], torch.Tensor], state, final_point, gate, velo, w, tol=1e-3, max_iter=200, check_every=1,
                   return_active=False):
    """Batched quad_moving.solver.  model: [B,18] float64 CUDA tensor -> [B,7] (only column 6 is used).  Returns (t1[B], iterations[B]).

    Every episode is masked once ITS |t2 - t1| <= tol, so extra iterations leave it untouched: the result does not depend on how often
    the host looks.  check_every = k > 0: the host tests "any episode still active" every k iterations (k = 1: after every iteration,
    i.e. one device->host sync per iteration); check_every = 0: NO host synchronisation at all -- exactly max_iter masked iterations
    are enqueued (the closed loop uses this; return_active=True adds the mask of episodes that would still iterate)."""
    B = state.shape[0]
    w = torch.as_tensor(w, dtype=torch.float64, device=state.device).expand(B).contiguous()
    t1 = torch.sqrt(((_centroid(gate) - state[:, 0:3]) ** 2).sum(1)) / 3                     # quad_moving.py:32
    t2 = model(ops.gate_frame(gate, velo, w, t1, state, final_point))[:, 6].to(torch.float64)
    iters = torch.zeros(B, dtype=torch.int32, device=state.device)
    active = (t2 - t1).abs() > tol
    for it in range(max_iter):
        if check_every > 0 and it % check_every == 0 and not bool(active.any()):
            break
        t1 = torch.where(active, t1 + (t2 - t1) / 2, t1)
        t2n = model(ops.gate_frame(gate, velo, w, t1, state, final_point))[:, 6].to(torch.float64)
        t2 = torch.where(active, t2n, t2)
        iters += active.to(torch.int32)
        active = (t2 - t1).abs() > tol
    return (t1, iters, active) if return_active else (t1, iters)


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


def _episode_group(model, p, state, final_point, gate, velo, w, n_ticks, sub_steps, dt_sim, gate_noise, record, warm_start, tt_iters, lo, hi):
    """Generator: one MPC tick of one group of episodes per next() -- everything is enqueued on the current stream, nothing synchronises."""
    dev = state.device
    B = state.shape[0]
    u = torch.zeros(B, 4, dtype=torch.float64, device=dev)                                    # main.py:52
    out = {"controls": [], "status": [], "t_tra": [], "iters": [], "tt_active": [], "states": []}
    prev = None
    for k in range(n_ticks):
        vel_k = velo if gate_noise is None else velo + gate_noise[k * sub_steps, lo:hi]
        t, _, act = traversal_time(model, state, final_point, gate, vel_k, w, max_iter=tt_iters, check_every=0, return_active=True)   # main.py:67
        inp = ops.gate_frame(gate, vel_k, w, t, state, final_point)                           # main.py:86-94
        o = model(inp).to(torch.float64)                                                      # main.py:96
        xi = ui = None
        if warm_start and prev is not None:                                                   # previous plan, one stage later
            xi = torch.cat([prev["state_traj_opt"][:, 1:], prev["state_traj_opt"][:, -1:]], 1)
            ui = torch.cat([prev["control_traj_opt"][:, 1:], prev["control_traj_opt"][:, -1:]], 1)
        sol = ops.solve(inp[:, 0:13], inp[:, 13:16], o[:, 0:3], o[:, 3:6], o[:, 6], ulast=u, round_t=False, params=p,
                        want_costate=False, want_state=warm_start, x_init=xi, u_init=ui)      # main.py:105-106 (get_input)
        prev = sol if warm_start else None
        u = sol["control_traj_opt"][:, 0, :].contiguous()
        out["controls"].append(u); out["status"].append(sol["status"]); out["t_tra"].append(t); out["iters"].append(sol["iters"])
        out["tt_active"].append(act.sum())
        if record:
            out["states"].append(state.clone())
        noise = None if gate_noise is None else gate_noise[k * sub_steps:(k + 1) * sub_steps, lo:hi].contiguous()
        ops.episode_advance(state, gate, u, velo, w, dt_sim, sub_steps, noise=noise, params=p)       # main.py:108 + gate.move, one launch
        yield out
    yield out


@torch.no_grad()
def run_episodes(model, state0, final_point, gate0, velo, w, n_ticks=50, sub_steps=10, dt_sim=0.01, params: Optional[Params] = None,
                 gate_noise: Optional[torch.Tensor] = None, record=False, warm_start=False, groups=1, tt_iters=24):
    """Batched main.py:65-116.  state0[B,13], final_point[B,3], gate0[B,12] (corners at time 0), velo[B,3], w (scalar or [B]).
    The gate moves like gate.move (rotation about its y axis by w dt, translation by (velo + noise) dt per plant step); `gate_noise`
    [n_ticks*sub_steps, B, 3] supplies the velocity noise (None: noise-free).  warm_start=False is the reference's behaviour (a fresh
    run_quad per tick, cold start); warm_start=True starts every re-solve from the previous tick's solution shifted by one stage (one
    MPC stage = one tick; BASELINE.json configs[3]).

    Device-resident: a tick enqueues the traversal-time fixed point (tt_iters masked iterations, no host test), the gate-frame packing,
    DNN-2, ONE solver launch and ONE fused plant + gate advance launch; the host never waits inside the loop.  groups > 1 splits the
    episodes into that many independent groups on their own CUDA streams (the episodes do not interact), so the straggler tail of
    one group's solver launch overlaps the other groups' work instead of idling the GPU at every tick.  The results do not depend on
    `groups` (every episode is solved by one warp with a fixed arithmetic order).

    Returns the final states, the controls of every tick, the per-tick solver status, iteration counts and traversal times,
    `tt_unconverged` [n_ticks] (episodes whose fixed point still moved after tt_iters iterations: raise tt_iters if non-zero) and the
    state trajectory at tick rate when record=True."""
    p = params or default_params()
    dev = state0.device
    B = state0.shape[0]
    groups = max(1, min(int(groups), B)) if B > 0 else 1
    w = torch.as_tensor(w, dtype=torch.float64, device=dev).expand(B).contiguous()
    cuts = [(g * B // groups, (g + 1) * B // groups) for g in range(groups)]
    cur = torch.cuda.current_stream(dev)
    streams = [cur] if groups == 1 else [torch.cuda.Stream(dev) for _ in range(groups)]
    gens, states, gates = [], [], []
    for (lo, hi), st in zip(cuts, streams):
        if st is not cur:
            st.wait_stream(cur)
        with torch.cuda.stream(st):
            states.append(state0[lo:hi].clone())
            gates.append(gate0.reshape(B, 12)[lo:hi].clone())
            gens.append(_episode_group(model, p, states[-1], final_point[lo:hi].contiguous(), gates[-1], velo[lo:hi].contiguous(), w[lo:hi].contiguous(),
                                       n_ticks, sub_steps, dt_sim, gate_noise, record, warm_start, tt_iters, lo, hi))
    outs = [None] * groups
    for _ in range(n_ticks + 1):                              # tick-major issue order: the streams interleave on the device
        for g, st in enumerate(streams):
            with torch.cuda.stream(st):
                outs[g] = next(gens[g])
    for st in streams:
        if st is not cur:
            cur.wait_stream(st)

    def cat(key, dim=0):
        parts = [torch.stack(o[key], 1) for o in outs]
        for t_ in parts:
            t_.record_stream(cur)
        return torch.cat(parts, dim)

    res = {"state": torch.cat(states, 0), "gate": torch.cat(gates, 0)}
    if n_ticks > 0:
        res.update({"controls": cat("controls"), "status": cat("status"), "t_tra": cat("t_tra"), "iters": cat("iters"),
                    "tt_unconverged": torch.stack([torch.stack(o["tt_active"]) for o in outs], 0).sum(0)})
        if record:
            res["states"] = cat("states")
    for t_ in states + gates:
        t_.record_stream(cur)
    return res
