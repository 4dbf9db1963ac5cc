"""Per-iteration latency of ONE warp alone vs under load: the makespan of a 4096-problem batch is set by its longest problem
(LPT starts it first), so what matters there is how fast a single warp walks through its iterations.

For each batch size B the first B problems of the same seeded batch are solved (so B=1 is the lone-warp case) and the launch
time is divided by the iteration count of the longest problem (a lower bound on that problem's latency share)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

sc = sample_scenarios(4096, 1)
full = [torch.as_tensor(sc[k], dtype=torch.float64, device="cuda") for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")]
r = se3.ops.objective(*full)
it = r["iters"].double()
order = torch.argsort(it, descending=True)
print("longest problems (iters):", it[order[:8]].tolist(), "mean", float(it.mean()))


def timed(args, reps=3):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); rr = se3.ops.objective(*args); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, rr


# the longest problem alone, then together with the next-longest ones (one warp per SM, 2, 4, 8 per SM), then the whole batch
for nb in (1, 2, 4, 37, 148, 296, 592, 1184, 2368, 4096):
    idx = order[:nb]
    args = [a[idx].contiguous() for a in full]
    ms, rr = timed(args)
    its = rr["iters"].double()
    print(f"B {nb:5d}  {ms:8.2f} ms  max iters {int(its.max()):4d}  sum iters {int(its.sum()):7d}  "
          f"ms per iteration of the longest {ms / float(its.max()):.4f}  warp-iterations/s {float(its.sum()) / ms * 1e3:10.0f}")
