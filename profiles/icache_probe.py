"""Is instruction supply the limiter?  The same launch size with (a) B distinct problems and (b) B copies of ONE problem: in (b) all
resident warps run the same code at (nearly) the same time, so the SM / GPC instruction caches see one hot region instead of ~8; the
data side (slots, L2, DRAM, shared memory) is unchanged.  Reports warp-iterations per second for both."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

keys = ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")


def timed(args, reps=2):
    best, r = 1e9, None
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = se3.ops.objective(*args); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, r


for B in (1184, 16384):
    sc = sample_scenarios(B, 1)
    full = [torch.as_tensor(sc[k], dtype=torch.float64, device="cuda") for k in keys]
    ms, r = timed(full)
    its = r["iters"].double()
    print(f"B {B:6d} distinct problems : {ms:8.2f} ms  sum iters {int(its.sum()):8d} max {int(its.max()):4d} -> {float(its.sum()) / ms * 1e-3:6.3f} M warp-iterations/s")
    for pick in (0, 5, 9):
        one = [a[pick:pick + 1].expand(B, *a.shape[1:]).contiguous() for a in full]
        ms1, r1 = timed(one)
        it1 = r1["iters"].double()
        print(f"B {B:6d} copies of problem {pick}: {ms1:8.2f} ms  iters each {int(it1[0]):4d} -> {float(it1.sum()) / ms1 * 1e-3:6.3f} M warp-iterations/s "
              f"({ms1 / float(it1[0]) / ((B + 1183) // 1184):.4f} ms per iteration and wave)")
