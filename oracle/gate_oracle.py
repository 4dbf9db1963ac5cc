"""CPU oracle (TEST INFRASTRUCTURE) for the moving-gate closed loop of main.py: numpy restatement of

  * quad_model.gate: __init__ / translate / rotate_y / transform / t_final      (quad_model.py:671-700, 720-729, 793-815)
  * the DNN-2 input packing of main.py:86-94 (identical in quad_moving.py:35-42)
  * the plant step of main.py:108 (explicit Euler, dt = 0.01: quad_model.py:86-119 with uav1.setDyn(0.01), main.py:35)
  * the traversal-time fixed point of quad_moving.solver                         (quad_moving.py:29-57)

Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.  Parity is PINNED: tests/golden/gate_frame.npz
is produced by the reference's own quad_model.gate / quad_moving.solver (tests/golden/make_golden_gate.py) and this file is
checked against it (tests/test_gate_oracle.py).  gate.transform goes through scipy's Rotation.from_quat / from_matrix / as_quat;
`_mat_to_quat` restates scipy's branch rule (largest of m00, m11, m22, trace), including its sign convention.
"""
from __future__ import annotations

import math

import numpy as np


def _norm(v):
    return v / np.sqrt(np.sum(v * v))


def _frame(gp):
    """centroid and I_G = [ax; ay; az] rows as left by gate.translate / gate.rotate_y (quad_model.py:697-700, 726-729)."""
    cen = np.array([np.mean(gp[:, 0]), np.mean(gp[:, 1]), np.mean(gp[:, 2])])
    az = np.array([0.0, 0.0, 1.0])
    ay = _norm(np.cross(gp[1] - gp[0], gp[2] - gp[1]))
    ax = np.cross(ay, az)
    return cen, np.array([ax, ay, az])


def predict_gate(gate_pts, velo, w, t):
    """gate(gate_pts).translate(t*velo); .rotate_y(t*w)  ->  (points 4x3, centroid, I_G)."""
    gp = np.array(gate_pts, float).reshape(4, 3) + np.asarray(velo, float) * t           # translate (quad_model.py:720-722)
    cen, _ = _frame(gp)
    ang = w * t
    c, s = math.cos(ang), math.sin(ang)
    rel = gp - cen                                                                        # rotate_y (quad_model.py:686-695)
    out = rel.copy()
    out[:, 0] = c * rel[:, 0] - s * rel[:, 2]
    out[:, 2] = s * rel[:, 0] + c * rel[:, 2]
    gp = out + cen
    cen, IG = _frame(gp)
    return gp, cen, IG


def _quat_to_mat(q_wxyz):
    w, x, y, z = q_wxyz / np.sqrt(np.sum(q_wxyz * q_wxyz))                                # scipy from_quat normalises
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _mat_to_quat(m):
    """scipy Rotation.from_matrix(m).as_quat() -> (x, y, z, w)."""
    d = [m[0, 0], m[1, 1], m[2, 2], m[0, 0] + m[1, 1] + m[2, 2]]
    c = int(np.argmax(d))
    q = np.zeros(4)
    if c != 3:
        i = c; j = (i + 1) % 3; k = (j + 1) % 3
        q[i] = 1 - d[3] + 2 * m[i, i]; q[j] = m[j, i] + m[i, j]; q[k] = m[k, i] + m[i, k]; q[3] = m[k, j] - m[j, k]
    else:
        q[0] = m[2, 1] - m[1, 2]; q[1] = m[0, 2] - m[2, 0]; q[2] = m[1, 0] - m[0, 1]; q[3] = 1 + d[3]
    return q / np.sqrt(np.sum(q * q))


def gate_frame_inputs(gate_pts, velo, w, t, state, final_point):
    """main.py:86-94: the 18 DNN-2 inputs [state in the predicted gate frame (13) | final point (3) | width | pitch]."""
    gp, cen, IG = predict_gate(gate_pts, velo, w, t)
    state = np.asarray(state, float)
    out = np.zeros(18)
    out[0:3] = IG @ (state[0:3] - cen)                                                    # gate.transform (quad_model.py:793-810)
    out[3:6] = IG @ state[3:6]
    out[10:13] = state[10:13]
    q = _mat_to_quat(IG @ _quat_to_mat(state[6:10]))
    out[6] = q[3]; out[7:10] = q[0:3]
    out[13:16] = IG @ (np.asarray(final_point, float) - cen)                              # gate.t_final (quad_model.py:813-815)
    out[16] = np.sqrt(np.sum((gp[0] - gp[1]) ** 2))
    out[17] = math.atan((gp[0, 2] - gp[1, 2]) / (gp[0, 0] - gp[1, 0]))
    return out


def quad_f(x, u, Jx=0.0023, Jy=0.0023, Jz=0.004, mass=0.5, arm_l=0.35, c_tau=0.0245, grav=9.78):
    """continuous dynamics of quad_model.py:86-119 (parameters quad_policy.py:37, quad_model.py:37)."""
    r, v, q, w = x[0:3], x[3:6], x[6:10], x[10:13]
    q0, q1, q2, q3 = q
    T = u[0] + u[1] + u[2] + u[3]
    # thrust along body z rotated to the world frame: C_B_I^T [0,0,T] (quad_model.py:93-96, 637-643)
    dv = np.array([2 * (q1 * q3 + q0 * q2), 2 * (q2 * q3 - q0 * q1), 1 - 2 * (q1 * q1 + q2 * q2)]) * T / mass + np.array([0, 0, -grav])
    wx, wy, wz = w
    dq = 0.5 * np.array([-wx * q1 - wy * q2 - wz * q3, wx * q0 + wz * q2 - wy * q3, wy * q0 - wz * q1 + wx * q3, wz * q0 + wy * q1 - wx * q2])
    Mx, My, Mz = arm_l / 2 * (u[3] - u[1]), arm_l / 2 * (u[2] - u[0]), c_tau * (u[0] - u[1] + u[2] - u[3])
    dw = np.array([(Mx - (Jz - Jy) * wy * wz) / Jx, (My - (Jx - Jz) * wx * wz) / Jy, (Mz - (Jy - Jx) * wx * wy) / Jz])
    return np.concatenate([v, dv, dq, dw])


def plant_steps(x, u, dt=0.01, n=1):
    """main.py:108: state = uav1.dyn_fn(state, u), n times (the control is held between MPC ticks)."""
    x = np.array(x, float)
    for _ in range(n):
        x = x + dt * quad_f(x, np.asarray(u, float))
    return x


def traversal_time(model, state, final_point, gate_pts, velo, w, tol=1e-3, max_iter=1000):
    """quad_moving.solver: t1 <- t1 + (t2 - t1)/2 until |t2 - t1| <= tol, t2 = DNN2(inputs(t1))[6].  `model`: 18-vector -> 7-vector."""
    gp = np.array(gate_pts, float).reshape(4, 3)
    cen = np.array([np.mean(gp[:, 0]), np.mean(gp[:, 1]), np.mean(gp[:, 2])])
    t1 = float(np.sqrt(np.sum((cen - np.asarray(state, float)[0:3]) ** 2)) / 3)
    t2 = float(model(gate_frame_inputs(gp, velo, w, t1, state, final_point))[6])
    it = 0
    while abs(t2 - t1) > tol and it < max_iter:
        t1 += (t2 - t1) / 2
        t2 = float(model(gate_frame_inputs(gp, velo, w, t1, state, final_point))[6])
        it += 1
    return t1, it
