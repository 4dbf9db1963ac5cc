"""GPU tier (-m gpu): the CUDA kernels, called through the C ABI, against the CPU oracle on identical inputs and
against the golden fixtures generated from the reference's own code.

Tolerances (BASELINE.json north_star): optimal cost 1e-6 relative, state/control trajectories 1e-5, reward 1e-6
(relative to max(1,|reward|) on solver outputs; 1e-9 absolute on identical trajectories).  Solves that end in a
different local minimum (cost differs by more than 1e-4 relative) are counted and bounded, not hidden."""
import numpy as np
import pytest
import torch

from tests.common import flat_w, sample_scenarios

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda")


@pytest.fixture(scope="module")
def se3():
    import learningagileflight_se3_b200 as m
    assert m._lib.lib().se3mpc_device_count() >= 1
    return m


@pytest.fixture(scope="module")
def batch64(oracle, se3):
    sc = sample_scenarios(64, 101)
    p = oracle.default_params()
    ref = oracle.solve_batch(p, sc["x0"], sc["goal"], sc["tra_pos"], sc["tra_ang"], sc["t"], round_t=True)
    got = se3.ops.solve(_t(sc["x0"]), _t(sc["goal"]), _t(sc["tra_pos"]), _t(sc["tra_ang"]), _t(sc["t"]), round_t=True)
    torch.cuda.synchronize()
    return sc, ref, {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in got.items()}


def test_solve_matches_oracle(batch64):
    sc, ref, got = batch64
    assert np.all(ref["status"] <= 1) and np.all(got["status"] <= 1)
    rel = np.abs(got["cost"] - ref["cost"]) / np.abs(ref["cost"])
    different_minimum = rel > 1e-4
    assert different_minimum.sum() <= 1, f"{different_minimum.sum()} of 64 solves ended in a different local minimum"
    ok = ~different_minimum
    assert np.all(rel[ok] <= 1e-6)
    assert np.abs(got["state_traj_opt"] - ref["X"])[ok].max() <= 1e-5
    assert np.abs(got["control_traj_opt"] - ref["U"])[ok].max() <= 1e-5
    lam_scale = np.abs(ref["LAM"]).max(axis=(1, 2), keepdims=True)
    assert (np.abs(got["costate_traj_opt"] - ref["LAM"]) / lam_scale)[ok].max() <= 1e-5
    assert np.abs(got["iters"] - ref["iters"])[ok].max() <= 5


def test_solution_is_a_kkt_point_of_the_nlp(batch64, oracle):
    """Solver-independent certificate on the GPU output: feasibility and objective value re-evaluated by the oracle's
    dense NLP (itself pinned to the reference's NLP by tests/golden)."""
    sc, ref, got = batch64
    p = oracle.default_params()
    for i in range(0, 64, 8):
        pb = oracle.make_problem(sc["x0"][i], sc["goal"][i], sc["tra_pos"][i], oracle.rd2quat(sc["tra_ang"][i]), oracle.round1(sc["t"][i]))
        w = flat_w(got["state_traj_opt"][i], got["control_traj_opt"][i])
        f, grad, c, jac, _ = oracle.nlp_eval_dense(p, pb, w, None, 1.0, want_jac=True, want_hess=False)
        assert abs(f - got["cost"][i]) <= 1e-8 * abs(f)
        assert np.abs(c).max() <= 1e-6                       # projection onto the original box moves c by <= 2e-7
        lam = got["costate_traj_opt"][i].reshape(-1)
        stat = grad + jac.T @ lam
        free_interior = np.ones_like(w, dtype=bool)
        free_interior[:13] = False
        N = p.N
        for k in range(N):
            u = w[k * 17 + 13:k * 17 + 17]
            free_interior[k * 17 + 13:k * 17 + 17] = (u > 1e-4) & (u < p.u_ub - 1e-4)
            om = w[(k + 1) * 17 + 10:(k + 1) * 17 + 13]
            free_interior[(k + 1) * 17 + 10:(k + 1) * 17 + 13] = np.abs(om) < p.w_bound - 1e-4
        assert np.abs(stat[free_interior]).max() <= 1e-3     # bound multipliers ~ mu/slack <= 1e-5 there; scaled tol 1e-8*s_d


def test_host_abi_equals_device_abi_and_is_batch_invariant(batch64, se3):
    sc, ref, got = batch64
    sub = slice(10, 20)
    h = se3.ops.solve_host(sc["x0"][sub], sc["goal"][sub], sc["tra_pos"][sub], sc["tra_ang"][sub], sc["t"][sub], round_t=True)
    # every problem is solved by one warp with a fixed arithmetic order: results do not depend on the batch around it
    assert np.array_equal(h["state_traj_opt"], got["state_traj_opt"][sub])
    assert np.array_equal(h["control_traj_opt"], got["control_traj_opt"][sub])
    assert np.array_equal(h["cost"], got["cost"][sub]) and np.array_equal(h["iters"], got["iters"][sub])


def test_reward_kernel_matches_oracle(oracle, se3):
    from tests.test_oracle_kkt_reward import _random_traj
    rng = np.random.default_rng(9)
    B = 512
    sc = sample_scenarios(B, 102)
    X = np.stack([_random_traj(rng) for _ in range(B)])
    X[::7, :, 1] = -np.abs(X[::7, :, 1]) - 0.5      # never crosses
    X[::11, :, 1] = np.abs(X[::11, :, 1]) + 1.5      # starts behind the plane
    p = oracle.default_params()
    want = np.array([oracle.reward(p, X[i], sc["gate"][i], sc["goal"][i]) for i in range(B)])
    got = se3.ops.reward(_t(X), _t(sc["gate"]), _t(sc["goal"]))
    torch.cuda.synchronize()
    assert np.abs(got["reward"].cpu().numpy() - want[:, 0]).max() <= 1e-9
    assert np.abs(got["collision"].cpu().numpy() - want[:, 1]).max() <= 1e-12
    assert np.abs(got["path"].cpu().numpy() - want[:, 2]).max() <= 1e-10
    assert np.array_equal(got["co"].cpu().numpy(), want[:, 3].astype(np.int32))
    hst = se3.ops.reward_host(X, sc["gate"], sc["goal"])
    assert np.array_equal(hst["reward"], got["reward"].cpu().numpy())
    # polyline-level collis_det (solid_geometry.obstacle.collis_det)
    pos = np.stack([oracle.rotor_positions(p, X[i])[:, 3:6] for i in range(64)])
    sc_o = np.array([oracle.collis_det(sc["gate"][i], pos[i], 50) for i in range(64)])
    s_g, co_g = se3.ops.collis_det_host(pos, sc["gate"][:64], 50)
    assert np.abs(s_g - sc_o[:, 0]).max() <= 1e-12 and np.array_equal(co_g, sc_o[:, 1].astype(np.int32))


def test_reward_kernel_ragged_batches_and_unaligned_trajectories(oracle, se3):
    """The stand-alone reward kernel fetches each trajectory chunk with a 16-byte-granular bulk copy from the aligned-down
    address: odd batch sizes (X ends 8 bytes off a 16-byte boundary: the tail fix-up path), a base pointer that is only 8-byte
    aligned, partially filled warps and a horizon whose last chunk is short must all give the oracle's numbers."""
    from tests.test_oracle_kkt_reward import _random_traj
    rng = np.random.default_rng(10)
    p = oracle.default_params()
    pool = np.stack([_random_traj(rng) for _ in range(131)])
    sc = sample_scenarios(131, 103)
    want = np.array([oracle.reward(p, pool[i], sc["gate"][i], sc["goal"][i]) for i in range(131)])
    for B in (1, 3, 31, 33, 64, 97, 131):
        for shift in (0, 1):
            flat = torch.empty(B * 51 * 13 + shift, dtype=torch.float64, device="cuda")      # exact-size allocation
            Xd = flat[shift:].view(B, 51, 13)
            Xd.copy_(_t(pool[:B]))
            assert Xd.data_ptr() % 16 == 8 * shift
            got = se3.ops.reward(Xd, _t(sc["gate"][:B]), _t(sc["goal"][:B]))
            torch.cuda.synchronize()
            assert np.abs(got["reward"].cpu().numpy() - want[:B, 0]).max() <= 1e-9, (B, shift)
            assert np.abs(got["collision"].cpu().numpy() - want[:B, 1]).max() <= 1e-12, (B, shift)
            assert np.array_equal(got["co"].cpu().numpy(), want[:B, 3].astype(np.int32)), (B, shift)
    # short horizons: N=2 (3 stages: one chunk), N=5 (6 stages: a full chunk of 4 and a short one)
    for N in (2, 5):
        pn = oracle.default_params(N=N)
        ps = se3.default_params(N=N)
        Xs = pool[:37, :N + 1].copy()
        w = np.array([oracle.reward(pn, Xs[i], sc["gate"][i], sc["goal"][i]) for i in range(37)])
        g = se3.ops.reward(_t(Xs), _t(sc["gate"][:37]), _t(sc["goal"][:37]), params=ps)
        torch.cuda.synchronize()
        assert np.abs(g["reward"].cpu().numpy() - w[:, 0]).max() <= 1e-9, N
        assert np.array_equal(g["co"].cpu().numpy(), w[:, 3].astype(np.int32)), N


def test_objective_and_sol_gradient_match_oracle(oracle, se3):
    sc = sample_scenarios(6, 103)
    p = oracle.default_params()
    ro = oracle.objective_batch(p, sc["x0"], sc["goal"], sc["gate"], sc["tra_pos"], sc["tra_ang"], sc["t"] + 0.03)
    rg = se3.ops.objective_host(sc["x0"], sc["goal"], sc["gate"], sc["tra_pos"], sc["tra_ang"], sc["t"] + 0.03)
    assert np.all(rg["status"] <= 1)
    assert np.all(np.abs(rg["reward"] - ro["reward"]) <= 1e-6 * np.maximum(1, np.abs(ro["reward"])))
    assert np.all(np.abs(rg["cost"] - ro["cost"]) <= 1e-6 * np.abs(ro["cost"]))
    ul = np.tile(np.array([1.0, 1.2, 0.8, 1.1]), (6, 1))
    out, st9 = se3.ops.sol_gradient_host(sc["x0"], sc["goal"], sc["gate"], sc["tra_pos"], sc["tra_ang"], sc["t"], ul)
    assert np.all(st9 <= 1)
    for i in range(6):
        want, _ = oracle.sol_gradient(p, sc["x0"][i], sc["goal"][i], sc["gate"][i], sc["tra_pos"][i], sc["tra_ang"][i], sc["t"][i], ul[i])
        assert abs(out[i, 7] - want[7]) <= 1e-6 * max(1, abs(want[7]))
        # the 7 pseudo-gradients are clipped reward DIFFERENCES scaled by <= 0.2 (quad_policy.py:100-110)
        assert np.abs(out[i, :7] - want[:7]).max() <= 2e-6 * max(1, abs(want[7]))


@pytest.mark.parametrize("variant", ["root", "gym"])
def test_golden_cases_from_the_reference_pipeline(se3, golden_root, golden_gym, variant):
    """End to end against tests/golden (reference run_quad.objective with the certified solution)."""
    g = golden_root if variant == "root" else golden_gym
    p = se3.default_params() if variant == "root" else se3.default_params(att_power=1, u_ub=2.4)
    n = int(g["n_cases"])
    x0 = np.stack([g[f"c{i}_x0"] for i in range(n)]); goal = np.stack([g[f"c{i}_goal"] for i in range(n)])
    gate = np.stack([g[f"c{i}_gate"] for i in range(n)]); tp = np.stack([g[f"c{i}_tra_pos"] for i in range(n)])
    ta = np.stack([g[f"c{i}_tra_ang"] for i in range(n)]); t = np.array([float(g[f"c{i}_t"]) for i in range(n)])
    r = se3.ops.objective_host(x0, goal, gate, tp, ta, t, params=p, want_traj=True)
    for i in range(n):
        assert r["status"][i] == 0
        assert abs(r["cost"][i] - float(g[f"c{i}_oracle_cost"])) <= 1e-6 * abs(r["cost"][i])
        assert np.abs(r["state_traj_opt"][i] - g[f"c{i}_oracle_X"]).max() <= 1e-5
        assert np.abs(r["control_traj_opt"][i] - g[f"c{i}_oracle_U"]).max() <= 1e-5
        ref = float(g[f"c{i}_reward_ref"])
        assert abs(r["reward"][i] - ref) <= 1e-6 * max(1, abs(ref))


def test_run_quad_shim_has_the_reference_call_shapes(se3, golden_root):
    from learningagileflight_se3_b200.quad_model import toQuaternion
    from learningagileflight_se3_b200.quad_policy import run_quad
    g, i = golden_root, 0
    inputs = g[f"c{i}_nn_inputs"]
    quad1 = run_quad(goal_pos=inputs[3:6], ini_r=inputs[0:3].tolist(), ini_q=toQuaternion(inputs[6], [0, 0, 1]))   # deep_learning.py:29
    quad1.init_obstacle(g[f"c{i}_gate"])
    assert np.allclose(quad1.ini_state, g[f"c{i}_x0"], rtol=0, atol=1e-15)
    j = quad1.objective(quad1.ini_state, g[f"c{i}_tra_pos"], g[f"c{i}_tra_ang"], float(g[f"c{i}_t"]))
    ref = float(g[f"c{i}_reward_ref"])
    assert isinstance(j, float) and abs(j - ref) <= 1e-6 * max(1, abs(ref))
    assert np.abs(quad1.traj - g[f"c{i}_traj_ref"]).max() <= 1e-5 and quad1.traj.shape == (51, 15)
    assert abs(quad1.collision - float(g[f"c{i}_collision_ref"])) <= 1e-6 and quad1.co == int(g[f"c{i}_co_ref"])
    gra = quad1.sol_gradient(quad1.ini_state, g[f"c{i}_tra_pos"], g[f"c{i}_tra_ang"], float(g[f"c{i}_t"]))
    assert gra.shape == (8,) and abs(gra[7] - ref) <= 1e-6 * max(1, abs(ref))
    u = quad1.get_input(quad1.ini_state, [0, 0, 0, 0], g[f"c{i}_tra_pos"], g[f"c{i}_tra_ang"], float(g[f"c{i}_t"]))
    assert u.shape == (4,) and set(quad1.sol1) >= {"state_traj_opt", "control_traj_opt", "costate_traj_opt", "time", "horizon", "cost"}
    assert quad1.sol1["state_traj_opt"].shape == (51, 13) and quad1.sol1["cost"].shape == (1, 1)
    assert np.abs(u - g[f"c{i}_oracle_U"][0]).max() <= 1e-5      # same t (already a multiple of 0.1 for case 0)
    # OCSys.ocSolver directly, as quad_policy.py:77 calls it
    quad1.uav1.init_TraCost(g[f"c{i}_tra_pos"], [2 * np.arctan(np.linalg.norm(g[f"c{i}_tra_ang"])), g[f"c{i}_tra_ang"] + np.array([1e-8, 0, 0])])
    quad1.uavoc1.setTraCost(quad1.uav1.tra_cost, float(g[f"c{i}_t"]))
    sol = quad1.uavoc1.ocSolver(ini_state=quad1.ini_state, horizon=50, dt=0.1, Ulast=None)
    assert np.abs(sol["state_traj_opt"] - g[f"c{i}_oracle_X"]).max() <= 1e-5


def test_torch_custom_op_and_full_batch_properties(se3, oracle):
    """BASELINE config 2 size (4096): every solve converges; results equal the oracle on a random subset; the
    registered torch op returns the same tensors."""
    se3.ops.register_torch_ops()
    B = 4096
    sc = sample_scenarios(B, 1)
    args = [_t(sc[k]) for k in ("x0", "goal", "gate", "tra_pos", "tra_ang", "t")]
    rew, status = torch.ops.se3mpc.objective(*args, None)
    r = se3.ops.objective(*args, want_traj=True)
    torch.cuda.synchronize()
    assert torch.equal(rew, r["reward"]) and torch.equal(status, r["status"])
    st = r["status"].cpu().numpy()
    assert (st <= 1).mean() >= 0.999, np.bincount(st)
    # dynamics feasibility of every returned trajectory (size-independent property): x_{k+1} = x_k + dt f(x_k,u_k)
    X, U = r["state_traj_opt"], r["control_traj_opt"]
    q, w, v = X[:, :-1, 6:10], X[:, :-1, 10:13], X[:, :-1, 3:6]
    T = U.sum(-1)
    acc = torch.stack([2 * (q[..., 1] * q[..., 3] + q[..., 0] * q[..., 2]), 2 * (q[..., 2] * q[..., 3] - q[..., 0] * q[..., 1]),
                       1 - 2 * (q[..., 1] ** 2 + q[..., 2] ** 2)], -1) * (T / 0.5)[..., None]
    acc[..., 2] -= 9.78
    ok = torch.as_tensor(st <= 1, device="cuda")
    assert ((X[:, 1:, 0:3] - (X[:, :-1, 0:3] + 0.1 * v)).abs().amax((1, 2))[ok] <= 1e-6).all()
    assert ((X[:, 1:, 3:6] - (v + 0.1 * acc)).abs().amax((1, 2))[ok] <= 1e-6).all()
    assert (U.amin((1, 2)) >= 0).all() and (U.amax((1, 2)) <= 2.44).all() and (X[:, 1:, 10:13].abs().amax((1, 2)) <= np.pi / 2).all()
    idx = np.random.default_rng(0).choice(B, 24, replace=False)
    p = oracle.default_params()
    ro = oracle.objective_batch(p, sc["x0"][idx], sc["goal"][idx], sc["gate"][idx], sc["tra_pos"][idx], sc["tra_ang"][idx], sc["t"][idx])
    rg, cg = r["reward"].cpu().numpy()[idx], r["cost"].cpu().numpy()[idx]
    same = np.abs(cg - ro["cost"]) <= 1e-4 * np.abs(ro["cost"])
    assert same.sum() >= 23
    assert np.all(np.abs(cg - ro["cost"])[same] <= 1e-6 * np.abs(ro["cost"])[same])
    assert np.all(np.abs(rg - ro["reward"])[same] <= 1e-6 * np.maximum(1, np.abs(ro["reward"]))[same])


def test_edge_cases_empty_ragged_horizons_and_bad_arguments(oracle, se3):
    """Empty and ragged batches (1, 3, 5, 33 problems: below / across the 4-warp CTA granularity), the shortest and the longest
    supported horizon, order independence, and argument errors (codes, no crash, no CPU fallback)."""
    from learningagileflight_se3_b200 import Se3MpcError
    p = oracle.default_params()
    sc = sample_scenarios(33, 301)
    keys = ("x0", "goal", "tra_pos", "tra_ang", "t")
    full = se3.ops.solve_host(*(sc[k] for k in keys), round_t=True)
    assert (full["status"] == 0).all()
    for B in (0, 1, 3, 5):
        r = se3.ops.solve_host(*(sc[k][:B] for k in keys), round_t=True)
        assert r["state_traj_opt"].shape == (B, 51, 13) and r["control_traj_opt"].shape == (B, 50, 4) and r["cost"].shape == (B,)
        assert np.array_equal(r["state_traj_opt"], full["state_traj_opt"][:B]) and np.array_equal(r["cost"], full["cost"][:B])
    perm = np.random.default_rng(0).permutation(33)                      # results do not depend on the position in the batch
    rp = se3.ops.solve_host(*(sc[k][perm] for k in keys), round_t=True)
    assert np.array_equal(rp["control_traj_opt"], full["control_traj_opt"][perm])
    for N in (2, 63):                                                     # SE3MPC_MAX_HORIZON = 63 (64 stage columns per row)
        pn = se3.default_params(N=N)
        po = oracle.default_params(N=N)
        r = se3.ops.solve_host(*(sc[k][:2] for k in keys), round_t=True, params=pn)
        for i in range(2):
            pb = oracle.make_problem(sc["x0"][i], sc["goal"][i], sc["tra_pos"][i], oracle.rd2quat(sc["tra_ang"][i]), oracle.round1(sc["t"][i]))
            so = oracle.solve(po, pb)
            assert so["status"] == int(r["status"][i]) == 0
            assert abs(r["cost"][i] - so["cost"]) <= 1e-6 * abs(so["cost"])
            assert np.abs(r["state_traj_opt"][i] - so["state_traj_opt"]).max() <= 1e-5
    with pytest.raises(Se3MpcError):
        se3.ops.solve_host(*(sc[k][:1] for k in keys), params=se3.default_params(N=64))          # horizon too long
    with pytest.raises(Se3MpcError):
        se3.ops.solve_host(*(sc[k][:1] for k in keys), params=se3.default_params(att_power=3))   # unknown cost variant
    with pytest.raises(Se3MpcError):
        se3.ops.solve(torch.zeros(1, 13, dtype=torch.float32, device="cuda"), _t(sc["goal"][:1]), _t(sc["tra_pos"][:1]), _t(sc["tra_ang"][:1]), _t(sc["t"][:1]))
