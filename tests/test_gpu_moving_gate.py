"""GPU tier, SURVEY.md 8(f) rank 1-2: the moving-gate closed loop of main.py / quad_moving.py through the C ABI, against the
reference-generated golden fixture (tests/golden/gate_frame.npz) and the CPU oracle (oracle/gate_oracle.py, oracle/oracle.py)."""
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gate_frame.npz")


def _torch_net(d, dev):
    import torch
    W1, b1, W2, b2 = (torch.as_tensor(d[k], dtype=torch.float64, device=dev) for k in ("W1", "b1", "W2", "b2"))

    def f(x):
        y = torch.tanh(x @ W1.T + b1) @ W2.T + b2
        y = y.clone()
        y[:, 6] = 2.0 + 1.5 * torch.tanh(0.5 * y[:, 6])
        return y
    return f


def _np_net(d):
    def f(x):
        y = d["W2"] @ np.tanh(d["W1"] @ np.asarray(x, float) + d["b1"]) + d["b2"]
        y[6] = 2.0 + 1.5 * math.tanh(0.5 * y[6])
        return y
    return f


def test_gate_frame_matches_reference_golden_host_and_device():
    import torch
    from learningagileflight_se3_b200 import ops
    d = np.load(GOLD)
    out, gp = ops.gate_frame_host(d["gates"], d["velo"], d["w"], d["t"], d["state"], d["final"], want_gate=True)
    assert np.abs(out - d["inputs"]).max() <= 1e-12                     # bit-level agreement up to libm vs CUDA sin/cos/atan
    dev = torch.device("cuda", 0)
    tt = lambda k: torch.as_tensor(d[k], dtype=torch.float64, device=dev)  # noqa: E731
    out_d, gp_d = ops.gate_frame(tt("gates"), tt("velo"), tt("w"), tt("t"), tt("state"), tt("final"), want_gate=True)
    assert np.array_equal(out_d.cpu().numpy(), out) and np.array_equal(gp_d.cpu().numpy(), gp)      # host ABI == device ABI
    from oracle import gate_oracle as G
    for i in range(8):
        g, _, _ = G.predict_gate(d["gates"][i], d["velo"][i], d["w"][i], d["t"][i])
        assert np.abs(gp[i].reshape(4, 3) - g).max() <= 1e-12
    assert ops.gate_frame_host(np.zeros((0, 12)), np.zeros((0, 3)), np.zeros(0), np.zeros(0), np.zeros((0, 13)), np.zeros((0, 3))).shape == (0, 18)


def test_plant_step_matches_oracle():
    from learningagileflight_se3_b200 import ops
    from oracle import gate_oracle as G
    rng = np.random.default_rng(5)
    B = 37
    x = rng.normal(0, 1, (B, 13)); u = rng.uniform(0, 2.44, (B, 4))
    for n in (1, 10):
        got = ops.plant_step_host(x, u, 0.01, n)
        ref = np.stack([G.plant_steps(x[i], u[i], 0.01, n) for i in range(B)])
        assert np.abs(got - ref).max() <= 1e-12
    assert np.array_equal(ops.plant_step_host(x, u, 0.01, 0), x)


def test_traversal_time_fixed_point_batched_and_reference_signature():
    import torch
    from learningagileflight_se3_b200 import quad_moving
    from learningagileflight_se3_b200.quad_model import gate
    from oracle import gate_oracle as G
    d = np.load(GOLD)
    dev = torch.device("cuda", 0)
    m = d["t_fix"].shape[0]
    tt = lambda k: torch.as_tensor(d[k][:m], dtype=torch.float64, device=dev)  # noqa: E731
    t1, iters = quad_moving.traversal_time(_torch_net(d, dev), tt("state"), tt("final"), tt("gates"), tt("velo"), tt("w"))
    assert np.abs(t1.cpu().numpy() - d["t_fix"]).max() <= 1e-9           # golden: the reference's own quad_moving.solver
    net = _np_net(d)
    for i in range(4):
        _, it = G.traversal_time(net, d["state"][i], d["final"][i], d["gates"][i], d["velo"][i], d["w"][i])
        assert int(iters[i]) == it                                        # each episode stops where its own while-loop would
        t = quad_moving.solver(net, d["state"][i], d["final"][i], gate(d["gates"][i].reshape(4, 3)), d["velo"][i], d["w"][i])
        assert abs(t - d["t_fix"][i]) <= 1e-9


def test_gate_mirror_class_transform_and_t_final():
    from learningagileflight_se3_b200.quad_model import gate
    d = np.load(GOLD)
    for i in range(6):
        g = gate(d["gates"][i].reshape(4, 3))
        g.translate(d["t"][i] * d["velo"][i])                             # main.py:86-87
        g.rotate_y(d["t"][i] * d["w"][i])
        assert np.abs(g.transform(d["state"][i]) - d["inputs"][i, 0:13]).max() <= 1e-12
        assert np.abs(g.t_final(d["final"][i]) - d["inputs"][i, 13:16]).max() <= 1e-12


def test_closed_loop_episodes_match_tick_by_tick_oracle(oracle):
    import torch
    from learningagileflight_se3_b200 import quad_moving
    from oracle import gate_oracle as G
    O = oracle
    d = np.load(GOLD)
    dev = torch.device("cuda", 0)
    B, ticks = 3, 2
    rng = np.random.default_rng(11)
    state0 = np.zeros((B, 13)); state0[:, 0:3] = rng.uniform(-2, 2, (B, 3)) + np.array([0, -6, 0]); state0[:, 6] = 1.0
    final = rng.uniform(-1, 1, (B, 3)) + np.array([0, 6, 0])
    gates = d["gates"][:B] * 1.0
    gates = gates.reshape(B, 4, 3) - gates.reshape(B, 4, 3).mean(1, keepdims=True)     # gates around the origin
    gates = gates.reshape(B, 12)
    velo = np.tile(np.array([0.3, 0.1, 0.1]), (B, 1)); w = 0.5
    net_t, net = _torch_net(d, dev), _np_net(d)
    # DNN-2 stand-in with moderate outputs: traversal position/attitude small, time from the fixture network
    def model_t(x):
        y = net_t(x)
        return torch.cat([0.1 * torch.tanh(y[:, 0:6]), y[:, 6:7]], 1)
    def model_n(x):
        y = net(x)
        return np.concatenate([0.1 * np.tanh(y[0:6]), y[6:7]])
    tt = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)  # noqa: E731
    res = quad_moving.run_episodes(model_t, tt(state0), tt(final), tt(gates), tt(velo), w, n_ticks=ticks)
    Ug = res["controls"].cpu().numpy()
    p = O.default_params()
    for b in range(B):
        x, g, u = state0[b].copy(), gates[b].copy(), np.zeros(4)
        for k in range(ticks):
            t, _ = G.traversal_time(model_n, x, final[b], g, velo[b], w)
            inp = G.gate_frame_inputs(g, velo[b], w, t, x, final[b])
            out = model_n(inp)
            pb = O.make_problem(inp[0:13], inp[13:16], out[0:3], O.rd2quat(out[3:6]), out[6], u)
            sol = O.solve(p, pb)
            u = sol["control_traj_opt"][0]
            assert np.abs(Ug[b, k] - u).max() <= 1e-5, (b, k)            # north_star: controls within 1e-5
            for _ in range(10):
                x = G.plant_steps(x, u, 0.01, 1)
                g = G.predict_gate(g, velo[b], w, 0.01)[0].reshape(12)
        assert np.abs(res["state"][b].cpu().numpy() - x).max() <= 1e-6


def test_warm_started_solve_matches_oracle_from_the_same_point(oracle):
    """se3mpc_solve_warm_device vs the oracle started from the same primal point: a perturbed copy of the optimum (the receding-horizon
    situation) -- same optimum, same tolerances as the cold start, fewer iterations."""
    import torch
    from learningagileflight_se3_b200 import ops
    from tests.common import sample_scenarios
    O = oracle
    p = O.default_params()
    B = 6
    sc = sample_scenarios(B, 41)
    dev = torch.device("cuda", 0)
    tt = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)  # noqa: E731
    cold = ops.solve(tt(sc["x0"]), tt(sc["goal"]), tt(sc["tra_pos"]), tt(sc["tra_ang"]), tt(sc["t"]), round_t=True)
    rng = np.random.default_rng(2)
    Xi = cold["state_traj_opt"].cpu().numpy() + rng.normal(0, 0.02, (B, 51, 13))
    Ui = cold["control_traj_opt"].cpu().numpy() + rng.normal(0, 0.02, (B, 50, 4))
    Ui[0, 3] = 5.0; Xi[1, 7, 10:13] = 9.0                                    # outside the bounds: must be pushed inside, not rejected
    warm = ops.solve(tt(sc["x0"]), tt(sc["goal"]), tt(sc["tra_pos"]), tt(sc["tra_ang"]), tt(sc["t"]), round_t=True, x_init=tt(Xi), u_init=tt(Ui))
    for i in range(B):
        pb = O.make_problem(sc["x0"][i], sc["goal"][i], sc["tra_pos"][i], O.rd2quat(sc["tra_ang"][i]), O.round1(sc["t"][i]))
        so = O.solve(p, pb, x_init=Xi[i], u_init=Ui[i])
        assert so["status"] == 0 and int(warm["status"][i]) == 0
        assert abs(float(warm["cost"][i]) - so["cost"]) <= 1e-6 * abs(so["cost"])
        assert np.abs(warm["state_traj_opt"][i].cpu().numpy() - so["state_traj_opt"]).max() <= 1e-5
        assert np.abs(warm["control_traj_opt"][i].cpu().numpy() - so["control_traj_opt"]).max() <= 1e-5
        assert abs(int(warm["iters"][i]) - so["iters"]) <= 3
    assert float(warm["iters"].double().mean()) < float(cold["iters"].double().mean())


def test_closed_loop_warm_start_reduces_iterations():
    import torch
    from learningagileflight_se3_b200 import quad_moving
    from tests.common import sample_scenarios
    d = np.load(GOLD)
    dev = torch.device("cuda", 0)
    B = 64
    sc = sample_scenarios(B, 43)
    net_t = _torch_net(d, dev)

    def model(x):
        y = net_t(x)
        return torch.cat([0.1 * torch.tanh(y[:, 0:6]), y[:, 6:7]], 1)
    tt = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)  # noqa: E731
    velo = tt(np.tile(np.array([0.3, 0.1, 0.1]), (B, 1)))
    cold = quad_moving.run_episodes(model, tt(sc["x0"]), tt(sc["goal"]), tt(sc["gate"]), velo, 0.5, n_ticks=4)
    warm = quad_moving.run_episodes(model, tt(sc["x0"]), tt(sc["goal"]), tt(sc["gate"]), velo, 0.5, n_ticks=4, warm_start=True)
    assert (warm["status"] <= 1).float().mean() >= 0.95 and (cold["status"] <= 1).float().mean() >= 0.95
    assert torch.equal(cold["controls"][:, 0], warm["controls"][:, 0])                 # tick 0 is a cold start in both
    ic, iw = cold["iters"][:, 1:].double().mean().item(), warm["iters"][:, 1:].double().mean().item()
    print(f"closed loop mean IPM iterations per re-solve: cold {ic:.1f}, warm {iw:.1f}")
    assert iw < ic


def test_quadrotor_shim_plant_and_rotor_positions_run_on_the_gpu(oracle):
    """Quadrotor.dyn_fn (main.py:108) and get_quadrotor_position (quad_model.py:239-276) of the mirror class go through the C ABI."""
    from learningagileflight_se3_b200 import ops
    from learningagileflight_se3_b200.quad_model import Quadrotor
    from oracle import gate_oracle as G
    rng = np.random.default_rng(8)
    uav = Quadrotor()
    uav.initDyn(Jx=0.0023, Jy=0.0023, Jz=0.004, mass=0.5, l=0.35, c=0.0245)
    uav.setDyn(0.01)
    x, u = rng.normal(0, 1, 13), rng.uniform(0, 2.4, 4)
    assert np.abs(uav.dyn_fn(x, u) - G.plant_steps(x, u, 0.01, 1)).max() <= 1e-13
    X = rng.normal(0, 1, (51, 13))
    p = oracle.default_params()
    assert np.abs(uav.get_quadrotor_position(1.5, X) - oracle.rotor_positions(p, X)).max() <= 1e-13
    assert ops.rotor_positions_host(np.zeros((0, 13)), 1.5).shape == (0, 15)


def test_episode_advance_is_the_fused_form_of_plant_step_and_gate_move():
    """ONE launch for the 10 sub-steps between two re-solves == 10 x (plant_step ; gate_frame(t = dt) -> predicted corners), bit for bit,
    with and without velocity noise; ragged batch."""
    import torch
    from learningagileflight_se3_b200 import ops
    dev = torch.device("cuda", 0)
    d = np.load(GOLD)
    rng = np.random.default_rng(9)
    B, n = 37, 10
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)  # noqa: E731
    x = rng.normal(0, 1, (B, 13)); u = rng.uniform(0, 2.44, (B, 4))
    gates = np.resize(d["gates"], (B, 12)) + rng.normal(0, 0.1, (B, 12))
    velo = rng.normal(0, 1, (B, 3)); w = rng.uniform(-1, 1, B); fin = rng.normal(0, 1, (B, 3))
    for noise in (None, rng.normal(0, 0.3, (n, B, 3))):
        xs, gs = tt(x), tt(gates)
        for j in range(n):
            xs = ops.plant_step(xs, tt(u), 0.01, 1)
            v = tt(velo) if noise is None else tt(velo + noise[j])
            _, gs = ops.gate_frame(gs, v, tt(w), torch.full((B,), 0.01, dtype=torch.float64, device=dev), xs, tt(fin), want_gate=True)
        xf, gf = tt(x), tt(gates)
        ops.episode_advance(xf, gf, tt(u), tt(velo), tt(w), 0.01, n, noise=None if noise is None else tt(noise))
        assert torch.equal(xf, xs) and torch.equal(gf, gs)


def test_closed_loop_is_sync_free_and_invariant_to_stream_groups():
    """run_episodes with 1 and 3 stream groups (and the fixed-point with / without host checks): identical controls, states, status."""
    import torch
    from learningagileflight_se3_b200 import quad_moving
    d = np.load(GOLD)
    dev = torch.device("cuda", 0)
    B, ticks = 10, 2
    rng = np.random.default_rng(12)
    state0 = np.zeros((B, 13)); state0[:, 0:3] = rng.uniform(-2, 2, (B, 3)) + np.array([0, -6, 0]); state0[:, 6] = 1.0
    final = rng.uniform(-1, 1, (B, 3)) + np.array([0, 6, 0])
    gates = np.resize(d["gates"], (B, 12)).reshape(B, 4, 3)
    gates = (gates - gates.mean(1, keepdims=True)).reshape(B, 12)
    velo = np.tile(np.array([0.3, 0.1, 0.1]), (B, 1)); w = 0.5
    net_t = _torch_net(d, dev)

    def model_t(x):
        y = net_t(x)
        return torch.cat([0.1 * torch.tanh(y[:, 0:6]), y[:, 6:7]], 1)
    tt = lambda a: torch.as_tensor(a, dtype=torch.float64, device=dev)  # noqa: E731
    noise = tt(rng.normal(0, 0.05, (ticks * 10, B, 3)))
    t_a, it_a = quad_moving.traversal_time(model_t, tt(state0), tt(final), tt(gates), tt(velo), w)                    # host check every iteration
    t_b, it_b, act = quad_moving.traversal_time(model_t, tt(state0), tt(final), tt(gates), tt(velo), w, max_iter=24, check_every=0, return_active=True)
    assert torch.equal(t_a, t_b) and torch.equal(it_a, it_b) and int(act.sum()) == 0 and int(it_a.max()) < 24
    r1 = quad_moving.run_episodes(model_t, tt(state0), tt(final), tt(gates), tt(velo), w, n_ticks=ticks, gate_noise=noise, record=True)
    r3 = quad_moving.run_episodes(model_t, tt(state0), tt(final), tt(gates), tt(velo), w, n_ticks=ticks, gate_noise=noise, record=True, groups=3)
    torch.cuda.synchronize()
    assert int(r1["tt_unconverged"].sum()) == 0
    for k in ("state", "gate", "controls", "status", "t_tra", "iters", "states"):
        assert r1[k].shape == r3[k].shape and torch.equal(r1[k], r3[k]), k
    assert r1["controls"].shape == (B, ticks, 4) and (r1["status"] <= 1).all()
