"""Iteration-count tail of the interior-point solves: the 1M problems of the configs[4] sweep (chunk seeds 400..415) and the 64 batches
the 8-rank bench draws (seeds 1000 (rank + 1) + s).  How many solves run past 200 / 300 / 500 / 1000 iterations, and what they end with."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learningagileflight_se3_b200 as se3  # noqa: E402
from tests.common import sample_scenarios  # noqa: E402

KEYS = ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")


def run(seeds, n):
    its, sts = [], []
    for sd in seeds:
        sc = sample_scenarios(n, sd)
        r = se3.ops.objective(*[torch.as_tensor(sc[k], dtype=torch.float64, device="cuda") for k in KEYS])
        its.append(r["iters"].cpu().numpy()); sts.append(r["status"].cpu().numpy())
    return np.concatenate(its), np.concatenate(sts)


for name, seeds, n in (("sweep chunks", range(400, 416), 62500), ("bench batches", [1000 * (r + 1) + s for r in range(8) for s in range(8)], 4096)):
    it, st = run(seeds, n)
    print(f"{name}: {it.size} problems, status {np.bincount(st, minlength=5).tolist()}, mean {it.mean():.2f}, p99 {np.quantile(it, 0.99):.0f}, p99.9 {np.quantile(it, 0.999):.0f}, "
          f"max {it.max()}; > 200: {(it > 200).sum()}, > 300: {(it > 300).sum()}, > 500: {(it > 500).sum()}, > 1000: {(it > 1000).sum()}; "
          f"converged with > 300: {((it > 300) & (st <= 1)).sum()}, largest converged {it[st <= 1].max()}")
