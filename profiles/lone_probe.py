"""Single-warp latency of fixed problems (B = 1 launches): ms, iterations, factorisations, line-search trials of each, for the library named
by SE3MPC_LIB.  Apples-to-apples across builds as long as the iteration counts agree."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

sc = sample_scenarios(4096, 1)
full = [torch.as_tensor(sc[k], dtype=torch.float64, device="cuda") for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")]
se3.ops.objective(*[a[:4].contiguous() for a in full])
out = []
for i in (7, 100, 51, 2389, 300, 301):
    args = [a[i:i + 1].contiguous() for a in full]
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = se3.ops.objective(*args); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    out.append(f"p{i}: {best:6.2f} ms it {int(r['iters'][0])} nf {int(r['nfact'][0])} nls {int(r['nls'][0])} -> {best / int(r['nfact'][0]) * 1e3:5.1f} us/fact")
print(" | ".join(out))
