"""A/B timing of one fused objective launch (B problems): LPT ordering on/off (SE3MPC_NO_LPT=1), L2 flush on/off."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3
from tests.common import sample_scenarios
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for seed in (1001, 1003, 1005):
    sc = sample_scenarios(B, seed)
    args = [torch.as_tensor(sc[k], dtype=torch.float64, device="cuda") for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")]
    for _ in range(2): r = se3.ops.objective(*args)
    torch.cuda.synchronize()
    out = {}
    for do_flush in (0, 1):
        ts = []
        for _ in range(3):
            if do_flush: flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = se3.ops.objective(*args); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        out[do_flush] = min(ts)
    nf = r["nfact"].cpu().numpy().astype(float)
    print("LPT off" if os.environ.get("SE3MPC_NO_LPT") else "LPT on", "seed", seed, "ms noflush %.1f flush %.1f" % (out[0], out[1]),
          "nfact mean %.1f max %d" % (nf.mean(), nf.max()), "conv", int((r["status"] <= 1).sum().item()))
