"""`run_quad` with the reference's call signatures (quad_policy.py:15-211) on top of the batched CUDA engine.

Single-problem methods (`objective`, `sol_gradient`, `get_input`) take and return exactly what the reference's do
(host numpy / floats) and go through the host-buffer C ABI; the `*_batch` methods take a leading batch dimension
and are what replaces the reference's process-per-sample loops (deep_learning.py:66-72, nn_train_2.py:63-69).
"""
from __future__ import annotations

import math

import numpy as np

from . import ops
from .quad_model import Quadrotor, toQuaternion
from .quad_OC import OCSys
from .solid_geometry import magni, norm, obstacle


def Rd2Rp(tra_ang):
    """Rodrigues-like vector -> [angle, unit axis] (quad_policy.py:10-13)."""
    a = np.asarray(tra_ang, dtype=float)
    theta = 2 * math.atan(magni(a))
    vector = norm(a + np.array([1e-8, 0, 0]))
    return [theta, vector]


class run_quad:
    def __init__(self, goal_pos=[0, 8, 0], goal_atti=[0, [1, 0, 0]], ini_r=[0, -8, 0], ini_v_I=[0.0, 0.0, 0.0],
                 ini_q=toQuaternion(0.0, [3, 3, 5]), horizon=50):
        self.winglen = 1.5                                   # quad_policy.py:19
        self.goal_pos = goal_pos
        self.goal_atti = goal_atti
        if type(ini_r) is not list:
            ini_r = np.asarray(ini_r, dtype=float).tolist()
        self.ini_r, self.ini_v_I, self.ini_q = ini_r, list(ini_v_I), list(ini_q)
        self.ini_w = [0.0, 0.0, 0.0]
        self.ini_state = self.ini_r + self.ini_v_I + self.ini_q + self.ini_w
        self.horizon = horizon
        # model + OCP with the literals of quad_policy.py:35-56
        self.uav1 = Quadrotor()
        self.uav1.initDyn(Jx=0.0023, Jy=0.0023, Jz=0.004, mass=0.5, l=0.35, c=0.0245)
        self.uav1.initCost(wrt=5, wqt=80, wthrust=0.1, wrf=5, wvf=5, wqf=0, wwf=3, goal_pos=self.goal_pos)
        self.uav1.init_TraCost()
        self.dt = 0.1
        self.uavoc1 = OCSys()
        self.uavoc1.setAuxvarVariable()
        sc, wc, tw, t2w = 1e20, math.pi / 2, 1.22, 2
        self.uavoc1.setStateVariable(self.uav1.X, state_lb=[-sc] * 10 + [-wc] * 3, state_ub=[sc] * 10 + [wc] * 3)
        self.uavoc1.setControlVariable(self.uav1.U, control_lb=[0, 0, 0, 0], control_ub=[t2w * tw] * 4)
        self.uavoc1.setDyn(self.uav1.f, self.dt)
        self.uavoc1.setthrustcost(self.uav1.thrust_cost)
        self.uavoc1.setPathCost(self.uav1.goal_cost)
        self.uavoc1.setTraCost(self.uav1.tra_cost)
        self.uavoc1.setFinalCost(self.uav1.final_cost)
        self._gate12 = None
        self._last_X = None
        self.last_status = None

    def params(self):
        p = self.uavoc1.params(self.horizon, self.dt)
        p.wing_len = self.winglen
        return p

    def init_obstacle(self, gate_point):
        g = np.asarray(gate_point, dtype=float).reshape(12)
        self.point1, self.point2, self.point3, self.point4 = g[0:3], g[3:6], g[6:9], g[9:12]
        self.obstacle1 = obstacle(self.point1, self.point2, self.point3, self.point4)
        self._gate12 = g

    # ---- quad_policy.py:67-91 -------------------------------------------------------------------------------
    def objective(self, ini_state=None, tra_pos=None, tra_ang=None, t=3, Ulast=None):
        if ini_state is None:
            ini_state = self.ini_state
        r = self.objective_batch(np.asarray(ini_state, float).reshape(1, 13), np.asarray(tra_pos, float).reshape(1, 3),
                                 np.asarray(tra_ang, float).reshape(1, 3), np.array([t], float),
                                 None if Ulast is None else np.asarray(Ulast, float).reshape(1, 4), want_traj=True)
        self._last_X = r["state_traj_opt"][0]
        self.last_status = int(r["status"][0])
        self.traj = self.uav1.get_quadrotor_position(wing_len=self.winglen, state_traj=self._last_X)
        self._details = None
        return float(r["reward"][0])

    def _reward_details(self):
        if self._details is None:
            self._details = ops.reward_host(self._last_X[None], self._gate12[None], np.asarray(self.goal_pos, float)[None],
                                            params=self.params())
        return self._details

    @property
    def collision(self):
        return float(self._reward_details()["collision"][0])

    @property
    def path(self):
        return float(self._reward_details()["path"][0])

    @property
    def co(self):
        return int(self._reward_details()["co"][0])

    def objective_batch(self, ini_state, tra_pos, tra_ang, t, Ulast=None, goal=None, gate=None, want_traj=False):
        """B objectives in one launch.  goal [B,3] / gate [B,12] default to this object's."""
        x0 = np.asarray(ini_state, float).reshape(-1, 13)
        B = x0.shape[0]
        goal = np.broadcast_to(np.asarray(self.goal_pos if goal is None else goal, float).reshape(-1, 3), (B, 3))
        if gate is None:
            if self._gate12 is None:
                raise RuntimeError("call init_obstacle(gate_point) first")
            gate = self._gate12
        gate = np.broadcast_to(np.asarray(gate, float).reshape(-1, 12), (B, 12))
        return ops.objective_host(x0, goal, gate, tra_pos, tra_ang, t, Ulast, params=self.params(), want_traj=want_traj)

    # ---- quad_policy.py:94-112 ------------------------------------------------------------------------------
    def sol_gradient(self, ini_state=None, tra_pos=None, tra_ang=None, t=None, Ulast=None):
        if ini_state is None:
            ini_state = self.ini_state
        out, st9 = self.sol_gradient_batch(np.asarray(ini_state, float).reshape(1, 13), np.asarray(tra_pos, float).reshape(1, 3),
                                           np.asarray(tra_ang, float).reshape(1, 3), np.array([t], float),
                                           None if Ulast is None else np.asarray(Ulast, float).reshape(1, 4))
        self.last_status = st9[0]
        return out[0]

    def sol_gradient_batch(self, ini_state, tra_pos, tra_ang, t, Ulast=None, goal=None, gate=None):
        x0 = np.asarray(ini_state, float).reshape(-1, 13)
        B = x0.shape[0]
        goal = np.broadcast_to(np.asarray(self.goal_pos if goal is None else goal, float).reshape(-1, 3), (B, 3))
        if gate is None:
            if self._gate12 is None:
                raise RuntimeError("call init_obstacle(gate_point) first")
            gate = self._gate12
        gate = np.broadcast_to(np.asarray(gate, float).reshape(-1, 12), (B, 12))
        return ops.sol_gradient_host(x0, goal, gate, tra_pos, tra_ang, t, Ulast, params=self.params())

    # ---- quad_policy.py:202-211: t is NOT rounded here -----------------------------------------------------------
    def get_input(self, ini_state, Ulast, tra_pos, tra_ang, t):
        sol = self.get_input_batch(np.asarray(ini_state, float).reshape(1, 13), None if Ulast is None else np.asarray(Ulast, float).reshape(1, 4),
                                   np.asarray(tra_pos, float).reshape(1, 3), np.asarray(tra_ang, float).reshape(1, 3), np.array([t], float))
        self.sol1 = {"state_traj_opt": sol["state_traj_opt"][0], "control_traj_opt": sol["control_traj_opt"][0],
                     "costate_traj_opt": sol["costate_traj_opt"][0], "auxvar_value": 1,
                     "time": np.array([k for k in range(self.horizon + 1)]), "horizon": self.horizon,
                     "cost": np.array([[sol["cost"][0]]]), "status": int(sol["status"][0])}
        return self.sol1["control_traj_opt"][0, :]

    def get_input_batch(self, ini_state, Ulast, tra_pos, tra_ang, t, goal=None):
        x0 = np.asarray(ini_state, float).reshape(-1, 13)
        B = x0.shape[0]
        goal = np.broadcast_to(np.asarray(self.goal_pos if goal is None else goal, float).reshape(-1, 3), (B, 3))
        return ops.solve_host(x0, goal, tra_pos, np.asarray(tra_ang, float).reshape(B, 3), t, Ulast, round_t=False, params=self.params())
