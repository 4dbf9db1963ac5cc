"""Stand-alone reward kernel (se3mpc_reward_kernel): achieved HBM GB/s on 2^18 trajectories (5456 algorithmic bytes each)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learningagileflight_se3_b200 import ops  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402
dev = torch.device("cuda", 0)
NB = 1 << 18
sc = sample_scenarios(1024, 7)
X = torch.randn(NB, 51, 13, dtype=torch.float64, device=dev)
X[:, :, 1] = torch.linspace(-6, 6, 51, dtype=torch.float64, device=dev)[None, :] + 0.3 * X[:, :, 1]      # fly through the gate plane
gate = torch.as_tensor(sc["gate"], dtype=torch.float64, device=dev).repeat(NB // 1024, 1)
goal = torch.as_tensor(sc["goal"], dtype=torch.float64, device=dev).repeat(NB // 1024, 1)
for _ in range(3): ops.reward(X, gate, goal)
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = ops.reward(X, gate, goal); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
rb = NB * (51 * 13 * 8 + 15 * 8 + 32)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists("MEASURED_PEAKS.json") else 6650.0
print(f"reward kernel: {min(ts):.3f} ms for {NB} trajectories -> {rb / (min(ts) * 1e-3) / 1e9:.0f} GB/s ({rb / (min(ts) * 1e-3) / 1e9 / peak * 100:.1f}% of {peak} GB/s); crossings scored: {int((r['collision'] != 0).sum())}")
