"""configs[3] (main.py moving narrow gate): B episodes in closed loop, batched -- throughput of the receding-horizon loop.
A random-init 18-64-64-7 ReLU network stands in for nn3_1.pth (quad_nn.network, fp32 like the reference), with its outputs squashed
into the planner's ranges (|tra_pos|, |tra_ang| <= 0.3, 1 <= t <= 4) so the MPC problems are of the usual kind.
Usage: python profiles/episodes_bench.py [B] [ticks]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from learningagileflight_se3_b200 import quad_moving  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(18, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 7)).to(dev)


def model(x):
    y = net(x.to(torch.float32)).to(torch.float64)
    return torch.cat([0.3 * torch.tanh(y[:, 0:6]), 2.5 + 1.5 * torch.tanh(y[:, 6:7])], 1)


sc = sample_scenarios(B, 3)
f64 = dict(dtype=torch.float64, device=dev)
x0 = torch.as_tensor(sc["x0"], **f64)
final = torch.as_tensor(sc["goal"], **f64)
gate0 = torch.as_tensor(sc["gate"], **f64)
velo = torch.tensor([1.0, 0.3, 0.4], **f64).expand(B, 3).contiguous()          # main.py:45
w = torch.pi / 2
quad_moving.run_episodes(model, x0[:256], final[:256], gate0[:256], velo[:256], w, n_ticks=1)      # warm-up
for warm in (False, True):
    torch.cuda.synchronize()
    l0 = se3.launch_count()
    t0 = time.perf_counter()
    res = quad_moving.run_episodes(model, x0, final, gate0, velo, w, n_ticks=ticks, warm_start=warm)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    conv = int((res["status"] <= 1).sum())
    print(f"B {B} ticks {ticks} {'warm' if warm else 'cold'} start: {dt:.3f} s -> {B * ticks / dt:.0f} closed-loop MPC solves/s ({conv}/{B * ticks} converged), "
          f"{dt / ticks * 1e3:.1f} ms per tick, {se3.launch_count() - l0} kernel launches, mean IPM iterations per re-solve "
          f"{res['iters'][:, 1:].double().mean().item():.1f}")
