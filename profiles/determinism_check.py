"""Run-to-run determinism of the solver: the same batch solved several times (ticket order and slot reuse differ from run to
run) must give bit-identical trajectories and iteration counts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sc = sample_scenarios(B, 1)
args = [torch.as_tensor(sc[k], dtype=torch.float64, device="cuda") for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")]
ref = None
for rep in range(4):
    r = se3.ops.solve(*args[:2], *args[3:6]) if False else se3.ops.objective(*args)
    cur = {k: r[k].clone() for k in ("reward", "cost", "iters", "nfact", "nls", "status", "kkt_error")}
    if ref is None:
        ref = cur
        continue
    bad = torch.zeros(B, dtype=torch.bool, device="cuda")
    for k in cur:
        bad |= (cur[k] != ref[k]) & ~((cur[k] != cur[k]) & (ref[k] != ref[k]))
    idx = torch.nonzero(bad).flatten().tolist()
    print(f"rep {rep}: {len(idx)} problems differ from rep 0", idx[:10])
    for i in idx[:5]:
        print("   ", i, {k: (ref[k][i].item(), cur[k][i].item()) for k in cur})
