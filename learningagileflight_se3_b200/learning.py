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


def dnn_from_arrays(arrays, device=None) -> Callable:
    """Re-assemble a quad_nn.network (quad_nn.py:100-125: Linear-ReLU-Linear-ReLU-Linear, fp32) from plain arrays
    {l1_weight, l1_bias, l2_weight, l2_bias, l3_weight, l3_bias} (e.g. tests/golden/nn3_1_weights.npz, the reference's shipped DNN-2)
    as a batched callable: x[B,D_in] (any float dtype / device) -> [B,D_out] float64.  Like the reference's forward it computes in
    float32 (`torch.tensor(input, dtype=torch.float)`, quad_nn.py:118)."""
    layers = []
    for i in (1, 2, 3):
        W = torch.as_tensor(np.asarray(arrays[f"l{i}_weight"]), dtype=torch.float32)
        lin = torch.nn.Linear(W.shape[1], W.shape[0])
        with torch.no_grad():
            lin.weight.copy_(W)
            lin.bias.copy_(torch.as_tensor(np.asarray(arrays[f"l{i}_bias"]), dtype=torch.float32))
        layers += [lin] + ([torch.nn.ReLU()] if i < 3 else [])
    net = torch.nn.Sequential(*layers)
    if device is not None:
        net = net.to(device)
    net.eval()
    for q in net.parameters():          # an inference copy of shipped weights (train your own torch module instead)
        q.requires_grad_(False)

    def model(x):
        x = torch.as_tensor(x)
        return net(x.to(device=next(net.parameters()).device, dtype=torch.float32)).to(torch.float64)
    model.net = net
    return model


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


def _model_batch(model: Callable, inputs: np.ndarray) -> torch.Tensor:
    """DNN outputs [B,7] for inputs[B,9] in ONE forward pass when the model can take a batch, else one call per sample.  The
    reference's network transposes its input (quad_nn.py:118: `S.t()`, "column 2D tensor"), i.e. it takes [9,B]; other models take
    [B,9].  Which form (if any) reproduces the per-sample outputs is probed once on the first two samples."""
    B = inputs.shape[0]
    single = [model(inputs[j]).reshape(7) for j in range(min(B, 2))]
    if B >= 2:
        with torch.no_grad():
            ref2 = torch.stack([s.detach() for s in single])
            form = None
            for name, arg in (("rows", inputs[:2]), ("cols", np.ascontiguousarray(inputs[:2].T))):
                try:
                    o = model(arg)
                    if isinstance(o, torch.Tensor) and o.numel() == 14 and torch.allclose(o.reshape(2, 7), ref2, rtol=1e-5, atol=1e-6):
                        form = name
                        break
                except Exception:
                    continue
        if form is not None:
            return model(inputs if form == "rows" else np.ascontiguousarray(inputs.T)).reshape(B, 7)
    return torch.stack(single + [model(inputs[j]).reshape(7) for j in range(len(single), B)])


def rl_step(model: Callable, optimizer, inputs, mode: str = "sequential", params: Optional[Params] = None, return_info: bool = False):
    """One mini-batch of the RL loop (deep_learning.py:44-88).  `model` maps a 9-vector (and, for the batched passes, [B,9]) to 7
    outputs.  Returns the rewards of the batch (n_gra[j][7]); with return_info also {"skipped": indices, "status9": ...}.

    Samples whose 9 solves did not all converge (status > 1: the reference never reads IPOPT's status, quad_OC.py:174-175, and would
    feed such rewards straight into Adam) or whose pseudo-gradient is not finite are SKIPPED: they take no optimizer step
    ("sequential") / carry zero weight ("batched"), so one failed solve cannot poison the weights.  Their rewards are still returned."""
    inputs = np.asarray(inputs, dtype=np.float64).reshape(-1, 9)
    B = inputs.shape[0]
    with torch.no_grad():                                                   # deep_learning.py:52-54 (snapshot outputs)
        outs = np.asarray(_model_batch(model, inputs).detach().cpu().numpy(), dtype=np.float64).reshape(B, 7)
    Dp, reward, st9 = rl_gradients(inputs, outs, params)                    # deep_learning.py:63-72, one launch instead of B processes
    good = np.all(np.asarray(st9) <= 1, axis=1) & np.all(np.isfinite(Dp), axis=1) & np.isfinite(reward)
    if mode == "sequential":
        for j in range(B):                                                  # deep_learning.py:75-81
            if not good[j]:
                continue
            outputs = model(inputs[j])
            loss = myloss_batch(outputs.reshape(1, 7), Dp[j:j + 1]).sum()
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
    elif mode == "batched":
        if good.any():
            outputs = _model_batch(model, inputs)                           # one forward pass for the mini-batch
            wgt = torch.as_tensor(good.astype(np.float64), dtype=outputs.dtype, device=outputs.device)
            loss = (myloss_batch(outputs, np.where(good[:, None], Dp, 0.0)) * wgt).sum() / wgt.sum()
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
    else:
        raise ValueError(mode)
    if return_info:
        return reward, {"skipped": np.where(~good)[0], "status9": st9}
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
