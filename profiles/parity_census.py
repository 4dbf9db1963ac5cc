"""Parity census at the BASELINE configs[1] size: every one of the 4096 GPU solves against the CPU oracle on the same inputs.
Counts agreement within the north_star tolerances (cost 1e-6 rel, X/U 1e-5, reward 1e-6) and solves that ended in a different
local minimum (cost differs by more than 1e-4 relative)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from oracle import oracle as O  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sc = sample_scenarios(B, seed)
t = lambda a: torch.as_tensor(a, dtype=torch.float64, device="cuda")  # noqa: E731
r = se3.ops.objective(t(sc["x0"]), t(sc["goal"]), t(sc["gate"]), t(sc["tra_pos"]), t(sc["tra_ang"]), t(sc["t"]), want_traj=True)
torch.cuda.synchronize()
p = O.default_params()
so = O.solve_batch(p, sc["x0"], sc["goal"], sc["tra_pos"], sc["tra_ang"], sc["t"], round_t=True)
ro = O.objective_batch(p, sc["x0"], sc["goal"], sc["gate"], sc["tra_pos"], sc["tra_ang"], sc["t"])
cg, co = r["cost"].cpu().numpy(), so["cost"]
rel = np.abs(cg - co) / np.abs(co)
dX = np.abs(r["state_traj_opt"].cpu().numpy() - so["X"].reshape(B, 51, 13)).max((1, 2))
dU = np.abs(r["control_traj_opt"].cpu().numpy() - so["U"].reshape(B, 50, 4)).max((1, 2))
dR = np.abs(r["reward"].cpu().numpy() - ro["reward"]) / np.maximum(1, np.abs(ro["reward"]))
stg, sto = r["status"].cpu().numpy(), so["status"]
itg, ito = r["iters"].cpu().numpy(), so["iters"]
ok = (rel <= 1e-6) & (dX <= 1e-5) & (dU <= 1e-5) & (dR <= 1e-6)
print(f"B {B} seed {seed}: within tolerance {int(ok.sum())}/{B}; different local minimum (cost rel diff > 1e-4): {int((rel > 1e-4).sum())}; "
      f"in between: {int(((rel > 1e-6) & (rel <= 1e-4)).sum())}; status GPU {np.bincount(stg, minlength=5).tolist()} oracle {np.bincount(sto, minlength=5).tolist()}; "
      f"identical iteration count {int((itg == ito).sum())}, |diff| <= 2: {int((np.abs(itg - ito) <= 2).sum())}; "
      f"max cost rel diff among agreeing {rel[ok].max():.1e}, max dX {dX[ok].max():.1e}, max dU {dU[ok].max():.1e}, max dReward {dR[ok].max():.1e}")
