"""configs[3] with the shipped DNN-2: which re-solves do not converge?  Runs the closed loop (cold start) for B episodes and dumps the
solver inputs of the non-converged ticks (status >= 2) to gpurun_out/closed_loop_failures.npz for analysis with the CPU oracle."""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from learningagileflight_se3_b200 import learning, ops, quad_moving  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 50
dev = torch.device("cuda", 0)
f64 = dict(dtype=torch.float64, device=dev)
model = learning.dnn_from_arrays(np.load("tests/golden/nn3_1_weights.npz"), dev)
sc = sample_scenarios(B, 3)
x0, fin, g0 = (torch.as_tensor(sc[k], **f64) for k in ("x0", "goal", "gate"))
velo = torch.tensor([1.0, 0.3, 0.4], **f64).expand(B, 3).contiguous()
p = se3.default_params()
# the loop of quad_moving.run_episodes, one group, recording the solver inputs of every tick
state, gate = x0.clone(), g0.clone()
w = torch.full((B,), math.pi / 2, **f64)
u = torch.zeros(B, 4, **f64)
rows = []
hist = np.zeros((ticks, 5), dtype=np.int64)
torch.set_grad_enabled(False)
for k in range(ticks):
    t, _, act = quad_moving.traversal_time(model, state, fin, gate, velo, w, max_iter=24, check_every=0, return_active=True)
    inp = ops.gate_frame(gate, velo, w, t, state, fin)
    o = model(inp).to(torch.float64)
    sol = ops.solve(inp[:, 0:13], inp[:, 13:16], o[:, 0:3], o[:, 3:6], o[:, 6], ulast=u, round_t=False, params=p, want_costate=False)
    st = sol["status"].cpu().numpy()
    hist[k] = np.bincount(st, minlength=5)[:5]
    bad = np.where(st >= 2)[0]
    for b in bad[:8]:
        rows.append(np.concatenate([[k, b, st[b], int(sol["iters"][b])], inp[b, 0:16].cpu().numpy(), o[b].cpu().numpy(), u[b].cpu().numpy(), [float(t[b])]]))
    u = sol["control_traj_opt"][:, 0, :].contiguous()
    ops.episode_advance(state, gate, u, velo, w, 0.01, 10, params=p)
print("status histogram per tick (converged, acceptable, max_iter, failed, nan):")
torch.set_grad_enabled(False)
for k in range(ticks):
    print(k, hist[k].tolist(), "mean y of the state in the gate frame", float(inp[:, 1].mean()) if k == ticks - 1 else "")
os.makedirs("gpurun_out", exist_ok=True)
np.savez("gpurun_out/closed_loop_failures.npz", rows=np.array(rows), hist=hist)
print("dumped", len(rows), "failing problems")
