"""Small single-GPU target for ncu captures: W warm-up + K timed launches of the fused objective kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sc = sample_scenarios(B, 1)
args = [torch.as_tensor(sc[k], dtype=torch.float64, device="cuda") for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")]
for _ in range(K):
    r = se3.ops.objective(*args)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); r = se3.ops.objective(*args); b.record(); torch.cuda.synchronize()
print("B", B, "ms", a.elapsed_time(b), "converged", int((r["status"] <= 1).sum()), "iters mean", float(r["iters"].double().mean()),
      "nfact mean", float(r["nfact"].double().mean()))
it = r["iters"].cpu().numpy(); nf = r["nfact"].cpu().numpy()
import numpy as np
top = np.argsort(-nf)[:5]
print("top nfact problems", top.tolist(), "nfact", nf[top].tolist(), "iters", it[top].tolist())
