"""configs[2] (deep_learning.py RL step): 1024 samples x 9 perturbed problems each = 9216 solves + rewards + the clip/scale epilogue in
ONE se3mpc_sol_gradient launch.  Usage: python profiles/rl_bench.py [samples]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
sc = sample_scenarios(B, 2)
args = [torch.as_tensor(sc[k], dtype=torch.float64, device="cuda") for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")]
for _ in range(2):
    g, st = se3.ops.sol_gradient(*args)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g, st = se3.ops.sol_gradient(*args); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
conv = int((st <= 1).sum())
print(f"sol_gradient: {B} samples = {9 * B} problems in {min(ts):.1f} ms -> {9 * B / (min(ts) * 1e-3):.0f} solves/s, {B / (min(ts) * 1e-3):.0f} gradient samples/s, "
      f"{conv}/{9 * B} converged, mean reward {float(g[:, 7].mean()):.2f}")
