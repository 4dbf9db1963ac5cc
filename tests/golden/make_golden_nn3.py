#!/usr/bin/env python
"""Export the reference's shipped DNN-2 (gym_pybullet_drone/nn3_1.pth, the network main.py:41-42 loads) as a plain-array fixture:
tests/golden/nn3_1_weights.npz = the six Linear tensors (fp32, 18-128-128-7 ReLU: quad_nn.network) plus reference outputs on seeded
inputs, so that the GPU box (which has no /root/reference) can run BASELINE.json configs[3] with the shipped network and the CPU tier
can check that the re-assembled network reproduces the reference module.  Loading follows SURVEY.md 8(c): `quad_nn.network` must be
importable for torch.load, which needs `casadi` / `matplotlib` stubbed (imported at the top of the reference modules, unused here).

Usage:  python tests/golden/make_golden_nn3.py
"""
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    cas = types.ModuleType("casadi")
    cas.numpy = np; cas.np = np; cas.atan = math.atan; cas.pi = math.pi
    sys.modules["casadi"] = cas
    for n in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation"):
        sys.modules[n] = types.ModuleType(n)
    sys.path.insert(0, "/root/reference")
    import quad_nn  # noqa: F401  (torch.load needs quad_nn.network)
    model = torch.load("/root/reference/gym_pybullet_drone/nn3_1.pth", weights_only=False)
    sd = {k.replace(".", "_"): v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    rng = np.random.default_rng(0)
    x = rng.normal(0, 1, (16, 18))
    x[:, 1] -= 6.0; x[:, 14] += 6.0; x[:, 6] += 2.0; x[:, 16] = rng.uniform(0.5, 1.25, 16); x[:, 17] = rng.uniform(-1, 1, 16)
    with torch.no_grad():
        y = np.stack([model(x[i]).numpy() for i in range(16)])           # the reference forward, one 18-vector at a time (main.py:96)
    np.savez(os.path.join(HERE, "nn3_1_weights.npz"), probe_in=x, probe_out=y, **sd)
    print("wrote nn3_1_weights.npz:", {k: v.shape for k, v in sd.items()}, "probe_out[0] =", y[0])


if __name__ == "__main__":
    main()
