"""Synthetic scenarios shared by tests and bench: the sampling rules of the reference's
quad_nn.nn_sample (quad_nn.py:18-48) re-stated with a seeded Generator, the gate construction of
deep_learning.py:25-27 / quad_model.py:735-741 and the pre-train traversal-time label of quad_nn.py:56."""
import math

import numpy as np


def make_gate(width, pitch):
    g = np.array([[-width / 2, 0, 1], [width / 2, 0, 1], [width / 2, 0, -1], [-width / 2, 0, -1]], dtype=float)
    c, s = math.cos(pitch), math.sin(pitch)
    out = g.copy()
    out[:, 0] = c * g[:, 0] - s * g[:, 2]        # rotate_y_out: [x,z] <- R(angle) [x,z] about the centroid (origin)
    out[:, 2] = s * g[:, 0] + c * g[:, 2]
    return out.reshape(12)


def make_gates(width, pitch):
    """make_gate for arrays width[B], pitch[B] -> [B,12] (same arithmetic, vectorised: the 1M-problem sweep of bench.py needs it)."""
    width, pitch = np.asarray(width, dtype=float), np.asarray(pitch, dtype=float)
    gx = np.stack([-width / 2, width / 2, width / 2, -width / 2], 1)
    gz = np.tile(np.array([1.0, 1.0, -1.0, -1.0]), (width.shape[0], 1))
    c, s = np.cos(pitch)[:, None], np.sin(pitch)[:, None]
    out = np.zeros((width.shape[0], 4, 3))
    out[:, :, 0] = c * gx - s * gz
    out[:, :, 2] = s * gx + c * gz
    return out.reshape(-1, 12)


def sample_scenarios(B, seed, noise=0.2):
    rng = np.random.default_rng(seed)
    r0 = rng.uniform(-5, 5, (B, 3)) + np.array([0, -9, 0])
    goal = rng.uniform(-2, 2, (B, 3)) + np.array([0, 6, 0])
    yaw = rng.uniform(-0.1, 0.1, B)
    width = np.clip(rng.normal(0.9, 0.3, B), 0.5, 1.25)
    angle = np.clip(1.3 * (1.2 - width), 0, math.pi / 3)
    angle1 = (math.pi / 2 - angle) / 3
    judge = rng.normal(0, 1, B)
    pos = np.clip(rng.normal(angle + angle1, 2 * angle1 / 3), angle, math.pi / 2)
    neg = np.clip(rng.normal(-angle - angle1, 2 * angle1 / 3), -math.pi / 2, -angle)
    pitch = np.where(judge > 0, pos, neg)
    x0 = np.zeros((B, 13))
    x0[:, 0:3] = r0
    x0[:, 6] = np.cos(yaw / 2)
    x0[:, 9] = np.sin(yaw / 2)
    gate = make_gates(width, pitch)
    tra_pos = rng.normal(0, noise, (B, 3))
    tra_ang = rng.normal(0, noise, (B, 3))
    t = np.clip(np.round(np.linalg.norm(r0, axis=1) / 4, 1), 2, 4)
    return {"x0": x0, "goal": goal, "gate": gate, "tra_pos": tra_pos, "tra_ang": tra_ang, "t": t,
            "width": width, "pitch": pitch}


def flat_w(X, U):
    """state/control trajectories -> the reference's decision vector w=[x0,u0,x1,u1,...,xN] (quad_OC.py:125-158)."""
    N = U.shape[0]
    w = np.zeros(17 * N + 13)
    for k in range(N):
        w[k * 17:k * 17 + 13] = X[k]
        w[k * 17 + 13:k * 17 + 17] = U[k]
    w[N * 17:] = X[N]
    return w
