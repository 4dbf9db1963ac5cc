"""Batched learning loops around the solver: the RL step of deep_learning.py and the imitation-data generator of nn_train_2.py.

The DNN stays in PyTorch (north_star); what changes is that the process-per-sample loops of the reference
(`deep_learning.py:44-72`, `nn_train_2.py:44-69`) become ONE solver launch per mini-batch:

  * `scenario_from_nn_inputs`  the 9 sampled numbers of quad_nn.nn_sample -> x0, goal, gate corners (deep_learning.py:24-30)
  * `myloss_batch`             network.myloss (quad_nn.py:141-145) for a batch: loss_b = Dp_b . para_b
  * `rl_gradients`             DNN-1 outputs -> 9 solves + rewards per sample -> pseudo-gradients Dp[B,7] and rewards[B] (one launch)
  * `rl_step`                  one mini-batch of deep_learning.py:44-88.  mode="sequential" reproduces the reference's update order
                               exactly (gradients from the pre-update snapshot, then one Adam step per sample with the outputs
                               recomputed by the current weights); mode="batched" takes ONE optimizer step on the mean surrogate loss.
  * `imitation_dataset`        nn_train_2.py:29-40, 72-84: one get_input solve per sample -> 50 labelled (inputs18, target7) pairs
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import numpy as np
import torch

from . import ops
from ._lib import Params, default_params


def scenario_from_nn_inputs(inputs):
    """inputs[B,9] = nn_sample rows [r0(3), goal(3), yaw, width, pitch] -> dict of numpy arrays x0[B,13], goal[B,3], gate[B,12].
    deep_learning.py:24-30: gate corners (+-w/2, 0, +-1) rotated about y by the pitch (gate.rotate_y_out), x0 = [r0, 0, q(yaw about z), 0]."""
    inp = np.asarray(inputs, dtype=np.float64).reshape(-1, 9)
    B = inp.shape[0]
    x0 = np.zeros((B, 13))
    x0[:, 0:3] = inp[:, 0:3]
    x0[:, 6] = np.cos(inp[:, 6] / 2)
    x0[:, 9] = np.sin(inp[:, 6] / 2)
    w, pit = inp[:, 7], inp[:, 8]
    c, s = np.cos(pit), np.sin(pit)
    gx = np.stack([-w / 2, w / 2, w / 2, -w / 2], 1)
    gz = np.tile(np.array([1.0, 1.0, -1.0, -1.0]), (B, 1))
    gate = np.zeros((B, 4, 3))
    gate[:, :, 0] = c[:, None] * gx - s[:, None] * gz
    gate[:, :, 2] = s[:, None] * gx + c[:, None] * gz
    return {"x0": x0, "goal": inp[:, 3:6].copy(), "gate": gate.reshape(B, 12)}


def myloss_batch(para: torch.Tensor, dp) -> torch.Tensor:
    """quad_nn.py:141-145 for a batch: loss_b = sum_i Dp[b,i] * para[b,i] (Dp is data, not differentiated)."""
    Dp = torch.as_tensor(np.asarray(dp), dtype=para.dtype, device=para.device) if not isinstance(dp, torch.Tensor) else dp.to(para.dtype).to(para.device)
    return (Dp.detach() * para).sum(-1)


def rl_gradients(inputs, outs, params: Optional[Params] = None):
    """run_quad.sol_gradient for every sample (quad_policy.py:94-112): inputs[B,9] (nn_sample), outs[B,7] (DNN-1 outputs, numpy).
    One se3mpc_sol_gradient launch of 9B problems.  Returns (Dp[B,7], reward[B], status9[B,9])."""
    sc = scenario_from_nn_inputs(inputs)
    outs = np.asarray(outs, dtype=np.float64).reshape(-1, 7)
    g, st9 = ops.sol_gradient_host(sc["x0"], sc["goal"], sc["gate"], outs[:, 0:3], outs[:, 3:6], outs[:, 6], None, params=params or default_params())
    return g[:, 0:7], g[:, 7], st9


def rl_step(model: Callable, optimizer, inputs, mode: str = "sequential", params: Optional[Params] = None):
    """One mini-batch of the RL loop (deep_learning.py:44-88).  `model` maps a 9-vector (or [B,9] for mode="batched") to 7 outputs.
    Returns the rewards of the batch (n_gra[j][7])."""
    inputs = np.asarray(inputs, dtype=np.float64).reshape(-1, 9)
    B = inputs.shape[0]
    with torch.no_grad():                                                   # deep_learning.py:52-54 (snapshot outputs)
        outs = np.stack([np.asarray(model(inputs[j]).detach().cpu().numpy(), dtype=np.float64).reshape(7) for j in range(B)])
    Dp, reward, _ = rl_gradients(inputs, outs, params)                      # deep_learning.py:63-72, one launch instead of B processes
    if mode == "sequential":
        for j in range(B):                                                  # deep_learning.py:75-81
            outputs = model(inputs[j])
            loss = myloss_batch(outputs.reshape(1, 7), Dp[j:j + 1]).sum()
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
    elif mode == "batched":
        outputs = torch.stack([model(inputs[j]).reshape(7) for j in range(B)])
        loss = myloss_batch(outputs, Dp).mean()
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
    else:
        raise ValueError(mode)
    return reward


def imitation_dataset(inputs, outs, params: Optional[Params] = None, horizon_labels: int = 50):
    """nn_train_2.py:29-40 (traj) and :72-84 (labels): for every sample one get_input solve (Ulast = 0, t not rounded), then for
    i < 50: X = [state_traj[i] (world frame, 13) | final position (3) | width | pitch], Y = [out[0:6] | out[6] - 0.1 i].
    inputs[B,9], outs[B,7] -> (X[B*50,18], Y[B*50,7], status[B])."""
    inputs = np.asarray(inputs, dtype=np.float64).reshape(-1, 9)
    outs = np.asarray(outs, dtype=np.float64).reshape(-1, 7)
    B = inputs.shape[0]
    p = params or default_params()
    sc = scenario_from_nn_inputs(inputs)
    sol = ops.solve_host(sc["x0"], sc["goal"], outs[:, 0:3], outs[:, 3:6], outs[:, 6], np.zeros((B, 4)), round_t=False, params=p, want_costate=False)
    S = sol["state_traj_opt"][:, :horizon_labels, :]                          # [B, 50, 13]
    X = np.zeros((B, horizon_labels, 18))
    X[:, :, 0:13] = S
    X[:, :, 13:16] = inputs[:, None, 3:6]
    X[:, :, 16:18] = inputs[:, None, 7:9]
    Y = np.zeros((B, horizon_labels, 7))
    Y[:, :, 0:6] = outs[:, None, 0:6]
    Y[:, :, 6] = outs[:, None, 6] - 0.10 * np.arange(horizon_labels)[None, :]
    return X.reshape(B * horizon_labels, 18), Y.reshape(B * horizon_labels, 7), sol["status"]
