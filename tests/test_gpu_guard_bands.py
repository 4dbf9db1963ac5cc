"""GPU tier: no device entry point writes outside the buffers it was given.

compute-sanitizer is not available on the GPU pool, so the bounds check is done by hand: while an operator runs, every
buffer it allocates (outputs, statistics, the solver workspace) is carved out of a larger allocation with 1 KiB guard bands
filled with 0xA5 on both sides; afterwards the bands must be untouched.  Batch sizes are ragged (not multiples of the warp /
CTA size), and one launch has more problems than resident warps (ticket queue + longest-first ordering kernels)."""
import types

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD_BYTES = 1024


class _Guarded:
    """Stand-in for the `torch` module inside ops.py: empty/empty_like return guarded views."""

    def __init__(self, torch):
        self._t, self.blocks = torch, []

    def __getattr__(self, k):
        return getattr(self._t, k)

    def empty(self, *shape, dtype=None, device=None, **kw):
        t = self._t
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, t.Size)):
            shape = tuple(shape[0])
        dtype = dtype or t.float32
        n = int(np.prod(shape)) if len(shape) else 1
        g = GUARD_BYTES // t.empty(0, dtype=dtype).element_size()
        raw = t.empty(n + 2 * g, dtype=dtype, device=device)
        raw.view(t.uint8).fill_(0xA5)
        self.blocks.append((raw, g, n))
        return raw[g:g + n].view(shape)

    def empty_like(self, x):
        return self.empty(tuple(x.shape), dtype=x.dtype, device=x.device)

    def check(self):
        t = self._t
        t.cuda.synchronize()
        assert self.blocks
        for raw, g, n in self.blocks:
            b = raw.view(t.uint8)
            e = raw.element_size()
            assert bool((b[:g * e] == 0xA5).all()), "write below a buffer of %d elements" % n
            assert bool((b[(g + n) * e:] == 0xA5).all()), "write above a buffer of %d elements" % n
        return len(self.blocks)


@pytest.fixture()
def guarded(monkeypatch):
    import torch
    from learningagileflight_se3_b200 import ops
    g = _Guarded(torch)
    monkeypatch.setattr(ops, "torch", g)
    monkeypatch.setattr(ops, "_workspaces", {})       # the cached workspace would bypass the guarded allocator
    return types.SimpleNamespace(ops=ops, g=g, torch=torch)


def _scen(torch, B, seed):
    from tests.common import sample_scenarios
    sc = sample_scenarios(B, seed)
    return {k: torch.as_tensor(sc[k], dtype=torch.float64, device="cuda:0") for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")}


@pytest.mark.parametrize("B", [1, 37, 1300])
def test_objective_stays_inside_its_buffers(guarded, B):
    ops, torch = guarded.ops, guarded.torch
    s = _scen(torch, B, 5)
    r = ops.objective(s["x0"], s["goal"], s["gate"], s["tra_pos"], s["tra_ang"], s["t"], want_traj=True)
    assert guarded.g.check() >= 9
    assert int((r["status"] <= 1).sum()) >= int(0.99 * B)
    assert bool(torch.isfinite(r["state_traj_opt"]).all() and torch.isfinite(r["control_traj_opt"]).all())


def test_solve_warm_reward_gradient_stay_inside_their_buffers(guarded):
    ops, torch = guarded.ops, guarded.torch
    s = _scen(torch, 37, 3)
    q = ops.solve(s["x0"], s["goal"], s["tra_pos"], s["tra_ang"], s["t"], round_t=True)
    X, U = q["state_traj_opt"], q["control_traj_opt"]
    xi = torch.cat([X[:, 1:], X[:, -1:]], 1).contiguous()
    ui = torch.cat([U[:, 1:], U[:, -1:]], 1).contiguous()
    w = ops.solve(X[:, 1].contiguous(), s["goal"], s["tra_pos"], s["tra_ang"], (s["t"] - 0.1).clamp_min(0.0), ulast=U[:, 0].contiguous(),
                  x_init=xi, u_init=ui)
    rw = ops.reward(X, s["gate"], s["goal"])
    s5 = _scen(torch, 5, 4)
    grad, st9 = ops.sol_gradient(s5["x0"], s5["goal"], s5["gate"], s5["tra_pos"], s5["tra_ang"], s5["t"])
    guarded.g.check()
    assert int((q["status"] <= 1).sum()) == 37 and int((w["status"] <= 1).sum()) == 37 and int((st9 <= 1).sum()) == 45
    assert bool(torch.isfinite(rw["reward"]).all() and torch.isfinite(grad).all())


def test_episode_kernels_stay_inside_their_buffers(guarded):
    ops, torch = guarded.ops, guarded.torch
    B = 37
    s = _scen(torch, B, 5)
    rng = np.random.default_rng(0)
    velo = torch.as_tensor(rng.normal(size=(B, 3)), device="cuda:0")
    wz = torch.as_tensor(rng.normal(size=B), device="cuda:0")
    gf, gp = ops.gate_frame(s["gate"], velo, wz, s["t"], s["x0"], s["goal"], want_gate=True)
    V = torch.as_tensor(rng.normal(size=(B, 51, 3)), device="cuda:0")
    score, co = ops.collis_det(V, s["gate"], 50)
    guarded.g.check()
    assert bool(torch.isfinite(gf).all() and torch.isfinite(gp).all() and torch.isfinite(score).all())
