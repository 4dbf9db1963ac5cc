"""Host-side mirror of the parts of the reference's quad_model.py that feed the hot path.

`Quadrotor` keeps the reference's method names (initDyn / initCost / init_TraCost / setDyn /
get_quadrotor_position, quad_model.py:14-276) but holds NUMBERS, not casadi expressions: the dynamics, costs and
their derivatives live as device functions in csrc/se3mpc_core.cuh.  The attributes the reference hands to OCSys
(`X, U, f, thrust_cost, goal_cost, tra_cost, final_cost`) are opaque tokens that `quad_OC.OCSys` recognises.
"""
from __future__ import annotations

import math

import numpy as np


def toQuaternion(angle, dir):
    """Axis-angle to scalar-first quaternion, same contract as quad_model.py:818-825 (returns a list)."""
    d = np.asarray(dir, dtype=float)
    d = d / np.linalg.norm(d)
    h = 0.5 * float(angle)
    s = math.sin(h)
    return [math.cos(h), s * d[0], s * d[1], s * d[2]]


class _Token:
    """Opaque stand-in for a casadi expression; remembers which model quantity it denotes."""

    def __init__(self, owner, kind, n=1):
        self.owner, self.kind, self._n = owner, kind, n

    def numel(self):
        return self._n

    def __repr__(self):
        return f"<se3mpc {self.kind}>"


class Quadrotor:
    def __init__(self, project_name="my UAV"):
        self.project_name = project_name
        self.X = _Token(self, "state", 13)
        self.U = _Token(self, "control", 4)
        self.tra_r_I = [0.0, 0.0, 5.0]
        self.tra_q = toQuaternion(0.7, [0, 1, 0])
        self.att_power = 2          # quad_model.py:210 squares the attitude error; the gym variant does not
        self.dt_plant = None

    # quad_model.py:35-119 -- all six parameters are numeric in every caller (quad_policy.py:37)
    def initDyn(self, Jx=None, Jy=None, Jz=None, mass=None, l=None, c=None):
        for name, v in (("Jx", Jx), ("Jy", Jy), ("Jz", Jz), ("mass", mass), ("l", l), ("c", c)):
            if v is None:
                raise ValueError(f"symbolic {name} is not supported: pass a number (the reference always does)")
        self.Jx, self.Jy, self.Jz, self.mass, self.l, self.c = float(Jx), float(Jy), float(Jz), float(mass), float(l), float(c)
        self.g = 9.78               # quad_model.py:37
        self.u_m = np.array([[1, 1, 1, 1], [0, -self.l / 2, 0, self.l / 2], [-self.l / 2, 0, self.l / 2, 0],
                             [self.c, -self.c, self.c, -self.c]])
        self.f = _Token(self, "dynamics", 13)

    # quad_model.py:121-198
    def initCost(self, wrt=None, wqt=None, wrf=None, wvf=None, wqf=None, wwf=None, wthrust=0.5, goal_pos=[0, 9, 5],
                 goal_velo=[0, 0, 0], goal_atti=[0, [1, 0, 0]]):
        if any(w is None for w in (wrt, wqt, wrf, wvf, wqf, wwf)):
            raise ValueError("symbolic cost weights are not supported: pass numbers (the reference always does)")
        if float(wqf) != 0.0:
            raise NotImplementedError("wqf != 0 (goal attitude cost) is not part of the accelerated path; "
                                      "the reference fixes wqf=0 (quad_policy.py:38)")
        if any(float(v) != 0.0 for v in goal_velo):
            raise NotImplementedError("non-zero goal velocity is not part of the accelerated path (reference uses 0)")
        self.wrt, self.wqt, self.wrf, self.wvf, self.wqf, self.wwf = map(float, (wrt, wqt, wrf, wvf, wqf, wwf))
        self.wthrust = float(wthrust)
        self.goal_r_I = np.asarray(goal_pos, dtype=float).reshape(3)
        self.goal_q = toQuaternion(goal_atti[0], goal_atti[1])
        self.thrust_cost = _Token(self, "thrust_cost")
        self.goal_cost = _Token(self, "goal_cost")
        self.final_cost = _Token(self, "final_cost")

    # quad_model.py:200-213
    def init_TraCost(self, tra_pos=[0, 0, 5], tra_atti=[0.7, [0, 1, 0]]):
        self.tra_r_I = [float(v) for v in np.asarray(tra_pos, dtype=float).reshape(-1)[0:3]]
        self.tra_q = toQuaternion(tra_atti[0], tra_atti[1])
        self.tra_cost = _Token(self, "tra_cost")

    # quad_model.py:215-219: explicit Euler plant used by main.py:35,108
    def setDyn(self, dt):
        self.dt_plant = float(dt)

    def dyn_fn(self, state, control):
        """x + dt*f(x,u) (closed-loop plant of main.py:108) through se3mpc_plant_step (one Euler step on the GPU)."""
        if self.dt_plant is None:
            raise RuntimeError("call setDyn(dt) first")
        from . import ops
        p = ops.default_params(Jx=self.Jx, Jy=self.Jy, Jz=self.Jz, mass=self.mass, arm_l=self.l, c_tau=self.c, grav=self.g)
        return ops.plant_step_host(np.asarray(state, dtype=float).reshape(1, 13), np.asarray(control, dtype=float).reshape(1, 4),
                                   self.dt_plant, 1, params=p)[0]

    # quad_model.py:239-276
    def get_quadrotor_position(self, wing_len, state_traj):
        from . import ops
        return ops.rotor_positions_host(np.asarray(state_traj, dtype=float).reshape(-1, 13), wing_len)


class gate:
    """Gate kinematics with the reference's interface (quad_model.py:671-815): 4x3 corner array, centroid, I_G, translate /
    rotate_y / translate_out / rotate_y_out / move / transform / t_final.  The bookkeeping methods are tiny host-side numpy (they are
    host-side in the reference too); `transform` and `t_final` -- the per-tick part of main.py:86-94 -- run on the GPU through
    ops.gate_frame_host for the gate as it stands (velo = 0, t = 0).  Batched callers use ops.gate_frame directly."""

    def __init__(self, gate_point=None):
        self.gate_point = np.array(gate_point, dtype=float).reshape(4, 3)
        self._refresh(transposed=True)          # quad_model.py:684 stores [ax ay az]^T in __init__, the rows elsewhere

    def _refresh(self, transposed=False):
        gp = self.gate_point
        self.centroid = np.array([np.mean(gp[:, 0]), np.mean(gp[:, 1]), np.mean(gp[:, 2])])
        az = np.array([0.0, 0.0, 1.0])
        n = np.cross(gp[1] - gp[0], gp[2] - gp[1])
        ay = n / np.sqrt(np.sum(n * n))
        ax = np.cross(ay, az)
        self.ay = ay
        self.I_G = np.array([ax, ay, az]).T if transposed else np.array([ax, ay, az])

    def rotate_y_out(self, angle):
        c, s = math.cos(angle), math.sin(angle)
        rel = self.gate_point - self.centroid
        out = rel.copy()
        out[:, 0] = c * rel[:, 0] - s * rel[:, 2]
        out[:, 2] = s * rel[:, 0] + c * rel[:, 2]
        return out + self.centroid

    def rotate_y(self, angle):
        self.gate_point = self.rotate_y_out(angle)
        self._refresh()

    def translate_out(self, displace):
        return self.gate_point + np.asarray(displace, dtype=float).reshape(3)

    def translate(self, displace):
        self.gate_point = self.translate_out(displace)
        self._refresh()

    def move(self, T=5, dt=0.01, v=[0, 0, 0], w=0, rng=None):
        """quad_model.py:769-790: corner trajectory [T/dt+1, 4, 3] and per-step velocity (clipped Gaussian noise, sigma 0.1).
        `rng`: numpy Generator (the reference draws from the global numpy state; pass np.random for that behaviour)."""
        rng = np.random if rng is None else rng
        gp = self.gate_point.copy()
        velo = np.asarray(v, dtype=float)
        moves, V = [gp.copy()], [velo.copy()]
        c, s = math.cos(dt * w), math.sin(dt * w)
        for _ in range(int(T / dt)):
            noise = np.clip(rng.normal(0, 0.1, 3), -0.1, 0.1)
            cen = np.array([np.mean(gp[:, 0]), np.mean(gp[:, 1]), np.mean(gp[:, 2])])
            rel = gp - cen
            out = rel.copy()
            out[:, 0] = c * rel[:, 0] - s * rel[:, 2]
            out[:, 2] = s * rel[:, 0] + c * rel[:, 2]
            gp = out + cen + dt * (velo + noise)
            moves.append(gp.copy()); V.append(velo + noise)
        return np.array(moves), np.array(V)

    def _frame_inputs(self, state, final_point):
        from . import ops
        z3 = np.zeros((1, 3))
        return ops.gate_frame_host(self.gate_point.reshape(1, 12), z3, np.zeros(1), np.zeros(1), np.asarray(state, float).reshape(1, 13),
                                   np.asarray(final_point, float).reshape(1, 3))[0]

    def transform(self, inertial_state):
        """World-frame state -> gate-frame state (quad_model.py:793-810), computed by se3mpc_gate_frame (rows-form I_G, i.e. the
        frame after translate / rotate_y, which is the only one the reference calls transform() in)."""
        return self._frame_inputs(inertial_state, np.zeros(3))[0:13]

    def t_final(self, final_point):
        return self._frame_inputs(np.array([0, 0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 0, 0]), final_point)[13:16]
