"""How much could a better longest-first predictor give?  Each bench.py batch (seeds 1003..1007, 4096 problems) is solved
(a) with the built-in predictor, (b) in natural order (SE3MPC_NO_LPT=1 process) and (c) in natural order after sorting the batch
by the TRUE iteration counts of a previous solve (a perfect predictor).  Run twice: with and without SE3MPC_NO_LPT=1."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402


def timed(args):
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = se3.ops.objective(*args); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, r


nolpt = os.environ.get("SE3MPC_NO_LPT") is not None
tot = [0.0, 0.0]
for seed in range(1003, 1008):
    sc = sample_scenarios(4096, seed)
    args = [torch.as_tensor(sc[k], dtype=torch.float64, device="cuda") for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")]
    ms, r = timed(args)
    it = r["iters"].double()
    order = torch.argsort(it, descending=True)
    ms2, _ = timed([a[order].contiguous() for a in args])
    tot[0] += ms; tot[1] += ms2
    print(f"seed {seed} max iters {int(it.max())} ideal {float(it.sum()) / 3.13e3:.1f} ms | {'natural order' if nolpt else 'built-in LPT'} {ms:.1f} ms | "
          f"batch pre-sorted by true iterations ({'natural order = perfect LPT' if nolpt else 'then built-in LPT'}) {ms2:.1f} ms")
print("sums", [round(t, 1) for t in tot], "k solves/s", [round(5 * 4096 / t, 1) for t in tot])
