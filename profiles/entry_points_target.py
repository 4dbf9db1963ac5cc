"""One small call of every device entry point of libse3mpc.so with ragged batch sizes (a target for ncu / ad-hoc checks).

    python profiles/entry_points_target.py [big]

`big` adds one 1300-problem objective launch (more problems than resident warps: ticket queue + longest-first kernels).
It was written as a compute-sanitizer target; compute-sanitizer is closed on this GPU pool, so the out-of-bounds check
is done with guard bands instead (tests/test_gpu_guard_bands.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from learningagileflight_se3_b200 import ops  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

big = len(sys.argv) > 1 and sys.argv[1] == "big"
dev = torch.device("cuda:0")
T = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float64, device=dev)  # noqa: E731


def scen(B, seed):
    sc = sample_scenarios(B, seed)
    return {k: T(sc[k]) for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")}


s = scen(37, 3)
r = ops.objective(s["x0"], s["goal"], s["gate"], s["tra_pos"], s["tra_ang"], s["t"], want_traj=True)
print("objective   converged", int((r["status"] <= 1).sum()), "/ 37")
q = ops.solve(s["x0"], s["goal"], s["tra_pos"], s["tra_ang"], s["t"], round_t=True)
print("solve       converged", int((q["status"] <= 1).sum()), "/ 37")
X, U = q["state_traj_opt"], q["control_traj_opt"]
xi = torch.cat([X[:, 1:], X[:, -1:]], 1).contiguous()
ui = torch.cat([U[:, 1:], U[:, -1:]], 1).contiguous()
w = ops.solve(X[:, 1].contiguous(), s["goal"], s["tra_pos"], s["tra_ang"], (s["t"] - 0.1).clamp_min(0.0), ulast=U[:, 0].contiguous(),
              x_init=xi, u_init=ui)
print("solve_warm  converged", int((w["status"] <= 1).sum()), "/ 37 mean iters", float(w["iters"].double().mean()))
rw = ops.reward(X, s["gate"], s["goal"])
print("reward      finite", bool(torch.isfinite(rw["reward"]).all()))
sc5 = scen(5, 4)
g, st9 = ops.sol_gradient(sc5["x0"], sc5["goal"], sc5["gate"], sc5["tra_pos"], sc5["tra_ang"], sc5["t"])
print("sol_gradient finite", bool(torch.isfinite(g).all()), "converged", int((st9 <= 1).sum()), "/ 45")
score, co = ops.collis_det(X[:, :, 0:3].contiguous(), s["gate"], X.shape[1] - 1)
print("collis_det  finite", bool(torch.isfinite(score).all()))
velo, wz = T(np.random.default_rng(0).normal(size=(37, 3))), T(np.random.default_rng(1).normal(size=37))
gf, gp = ops.gate_frame(s["gate"], velo, wz, s["t"], s["x0"], s["goal"], want_gate=True)
print("gate_frame  finite", bool(torch.isfinite(gf).all() and torch.isfinite(gp).all()))
xn = ops.plant_step(s["x0"], U[:, 0].contiguous(), dt=0.01, nsteps=10)
print("plant_step  finite", bool(torch.isfinite(xn).all()))
if big:
    b = scen(1300, 5)
    r = ops.objective(b["x0"], b["goal"], b["gate"], b["tra_pos"], b["tra_ang"], b["t"])
    print("objective   converged", int((r["status"] <= 1).sum()), "/ 1300")
torch.cuda.synchronize()
print("launches", se3.launch_count())
