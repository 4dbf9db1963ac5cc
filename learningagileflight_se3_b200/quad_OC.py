"""`OCSys` with the reference's call signatures (quad_OC.py:8-212) on top of the batched CUDA solver.

The reference builds a casadi NLP and calls IPOPT per solve; here `ocSolver` packs the numbers that define the
same NLP into `se3mpc_params` + per-problem inputs and calls the C ABI (`se3mpc_solve_host`: host numpy buffers
in, host numpy buffers out, exactly what a caller of the reference passes and receives).  `ocSolverBatch` is the
batched variant (leading dimension B).  There is no CPU path: without the CUDA library the call raises.
"""
from __future__ import annotations

import numpy as np

from . import _lib, ops
from .quad_model import _Token


class OCSys:
    def __init__(self, project_name="my optimal control system"):
        self.project_name = project_name
        self.t = 3.0
        self._model = None
        self.last_stats = None

    # -- the reference's builder API; symbolic arguments are replaced by the Quadrotor tokens -------------------
    def setAuxvarVariable(self, auxvar=None):
        self.auxvar = auxvar
        self.n_auxvar = 1

    def setStateVariable(self, state, state_lb=[], state_ub=[]):
        self.state = state
        self.n_state = state.numel()
        self._bind(state)
        self.state_lb = list(state_lb) if len(state_lb) == self.n_state else self.n_state * [-1e20]
        self.state_ub = list(state_ub) if len(state_ub) == self.n_state else self.n_state * [1e20]

    def setControlVariable(self, control, control_lb=[], control_ub=[]):
        self.control = control
        self.n_control = control.numel()
        self._bind(control)
        self.control_lb = list(control_lb) if len(control_lb) == self.n_control else self.n_control * [-1e20]
        self.control_ub = list(control_ub) if len(control_ub) == self.n_control else self.n_control * [1e20]

    def setDyn(self, f, dt):
        if not hasattr(self, "auxvar"):
            self.setAuxvarVariable()
        self._bind(f)
        self.dyn = f
        self.dt = float(dt)

    def setthrustcost(self, thrust_cost):
        self._bind(thrust_cost)
        self.thrust_cost = thrust_cost

    def setPathCost(self, path_cost):
        self._bind(path_cost)
        self.path_cost = path_cost

    def setFinalCost(self, final_cost):
        self._bind(final_cost)
        self.final_cost = final_cost

    def setTraCost(self, tra_cost, t=3.0):
        self._bind(tra_cost)
        self.t = t
        self.tra_cost = tra_cost

    def _bind(self, token):
        if not isinstance(token, _Token):
            raise TypeError("OCSys (se3mpc) takes the attributes of learningagileflight_se3_b200.quad_model.Quadrotor, "
                            "not casadi expressions")
        if self._model is None:
            self._model = token.owner
        elif self._model is not token.owner:
            raise ValueError("all OCSys pieces must come from the same Quadrotor")

    # -- parameters of the NLP this object currently describes ---------------------------------------------------
    def params(self, horizon, dt) -> _lib.Params:
        m = self._model
        lb, ub = np.asarray(self.control_lb, float), np.asarray(self.control_ub, float)
        slb, sub = np.asarray(self.state_lb, float), np.asarray(self.state_ub, float)
        if not (np.all(lb == lb[0]) and np.all(ub == ub[0])):
            raise NotImplementedError("per-rotor thrust bounds are not part of the accelerated path")
        if not (np.all(slb[:10] <= -1e19) and np.all(sub[:10] >= 1e19) and np.all(sub[10:] == sub[10]) and np.all(slb[10:] == -sub[10])):
            raise NotImplementedError("only the reference's state bounds (free r,v,q; symmetric box on omega) are supported")
        return _lib.default_params(Jx=m.Jx, Jy=m.Jy, Jz=m.Jz, mass=m.mass, arm_l=m.l, c_tau=m.c, grav=m.g,
                                   wrt=m.wrt, wqt=m.wqt, wthrust=m.wthrust, wrf=m.wrf, wvf=m.wvf, wwf=m.wwf,
                                   dt=float(dt), u_lb=float(lb[0]), u_ub=float(ub[0]), w_bound=float(sub[10]),
                                   N=int(horizon), att_power=int(m.att_power))

    def _check_ready(self):
        assert hasattr(self, "state"), "Define the state variable first!"
        assert hasattr(self, "control"), "Define the control variable first!"
        assert hasattr(self, "dyn"), "Define the system dynamics first!"
        assert hasattr(self, "path_cost"), "Define the running cost function first!"
        assert hasattr(self, "final_cost"), "Define the final cost function first!"

    def ocSolver(self, ini_state, Ulast=None, horizon=None, auxvar_value=1, print_level=0, dt=0.1, costate_option=0):
        """One solve; same return dict as quad_OC.py:204-210 (plus `status`, which the reference never reads)."""
        self._check_ready()
        if costate_option != 0:
            raise NotImplementedError("costate_option=1 (PMP recursion) is not used by any reference caller")
        sol = self.ocSolverBatch(np.asarray(ini_state, dtype=float).reshape(1, 13),
                                 None if Ulast is None else np.asarray(Ulast, dtype=float).reshape(1, 4),
                                 horizon=horizon, dt=dt)
        return {"state_traj_opt": sol["state_traj_opt"][0], "control_traj_opt": sol["control_traj_opt"][0],
                "costate_traj_opt": sol["costate_traj_opt"][0], "auxvar_value": auxvar_value,
                "time": np.array([k for k in range(int(horizon) + 1)]), "horizon": horizon,
                "cost": np.array([[sol["cost"][0]]]), "status": int(sol["status"][0])}

    def ocSolverBatch(self, ini_state, Ulast=None, horizon=None, dt=0.1, tra_pos=None, tra_q=None, goal=None, t=None):
        """B solves in one launch.  ini_state [B,13]; Ulast [B,4] or None.  tra_pos/tra_q/goal/t default to the model's."""
        self._check_ready()
        m = self._model
        x0 = np.asarray(ini_state, dtype=float).reshape(-1, 13)
        B = x0.shape[0]
        p = self.params(horizon, dt)
        goal = np.broadcast_to(np.asarray(m.goal_r_I if goal is None else goal, float).reshape(-1, 3), (B, 3))
        tra_pos = np.broadcast_to(np.asarray(m.tra_r_I if tra_pos is None else tra_pos, float).reshape(-1, 3), (B, 3))
        tt = np.broadcast_to(np.asarray(self.t if t is None else t, float).reshape(-1), (B,))
        q = np.ascontiguousarray(np.broadcast_to(np.asarray(m.tra_q if tra_q is None else tra_q, float).reshape(-1, 4), (B, 4)))
        sol = ops.solve_host(x0, goal, tra_pos, q, tt, Ulast, round_t=False, params=p)
        self.last_stats = {k: sol[k] for k in ("status", "iters", "nfact", "nls", "kkt_error")}
        return sol
