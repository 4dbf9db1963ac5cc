"""GPU tier, SURVEY.md 8(f) rank 3-4: the batched RL step (deep_learning.py) and imitation-data generator (nn_train_2.py)."""
import copy
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class Net:
    """quad_nn.network(9, 16, 16, 7) restated with the reference's call convention: numpy vector in, tensor out (fp32)."""
    def __init__(self, seed, d_in=9):
        import torch
        torch.manual_seed(seed)
        self.m = torch.nn.Sequential(torch.nn.Linear(d_in, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 7))
        with torch.no_grad():
            self.m[4].weight.mul_(0.05); self.m[4].bias.copy_(torch.tensor([0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 2.5]))   # planner-like outputs

    def __call__(self, x):
        import torch
        return self.m(torch.tensor(np.asarray(x), dtype=torch.float))

    def parameters(self):
        return self.m.parameters()


def _nn_inputs(B, seed):
    from tests.common import sample_scenarios
    sc = sample_scenarios(B, seed)
    inp = np.zeros((B, 9))
    inp[:, 0:3] = sc["x0"][:, 0:3]; inp[:, 3:6] = sc["goal"]
    inp[:, 6] = 2 * np.arctan2(sc["x0"][:, 9], sc["x0"][:, 6]); inp[:, 7] = sc["width"]; inp[:, 8] = sc["pitch"]
    return inp, sc


def test_scenario_from_nn_inputs_is_the_reference_construction():
    from learningagileflight_se3_b200 import learning
    inp, sc = _nn_inputs(16, 5)
    s = learning.scenario_from_nn_inputs(inp)
    assert np.abs(s["x0"] - sc["x0"]).max() <= 1e-15 and np.abs(s["gate"] - sc["gate"]).max() <= 1e-15 and np.array_equal(s["goal"], sc["goal"])


def test_rl_step_sequential_reproduces_the_reference_update_rule(oracle):
    import torch
    from learningagileflight_se3_b200 import learning
    O = oracle
    p = O.default_params()
    inp, _ = _nn_inputs(3, 21)
    net = Net(0)
    ref = copy.deepcopy(net)
    opt = torch.optim.Adam(net.parameters(), lr=1e-4)
    rew = learning.rl_step(net, opt, inp, mode="sequential")
    # the reference's loop (deep_learning.py:44-81) with the ORACLE's sol_gradient in place of the 10 worker processes
    opt_r = torch.optim.Adam(ref.parameters(), lr=1e-4)
    outs = [ref(inp[j]).data.numpy().astype(np.float64) for j in range(3)]
    sc = learning.scenario_from_nn_inputs(inp)
    gra = [O.sol_gradient(p, sc["x0"][j], sc["goal"][j], sc["gate"][j], outs[j][0:3], outs[j][3:6], outs[j][6])[0] for j in range(3)]
    for j in range(3):
        outputs = ref(inp[j])
        loss = torch.matmul(torch.tensor(gra[j][0:7], dtype=torch.float), outputs)          # network.myloss, quad_nn.py:141-145
        opt_r.zero_grad(); loss.backward(); opt_r.step()
        assert abs(rew[j] - gra[j][7]) <= 1e-6 * max(1.0, abs(gra[j][7]))
    for a, b in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=0, atol=2e-6)
    # batched mode: one step, parameters move, rewards identical
    net2 = Net(0)
    before = [q.detach().clone() for q in net2.parameters()]
    rew2 = learning.rl_step(net2, torch.optim.Adam(net2.parameters(), lr=1e-4), inp, mode="batched")
    assert np.array_equal(rew, rew2) and any(not torch.equal(a, b) for a, b in zip(before, net2.parameters()))


def test_imitation_dataset_matches_oracle_trajectories(oracle):
    from learningagileflight_se3_b200 import learning
    O = oracle
    p = O.default_params()
    B = 4
    inp, _ = _nn_inputs(B, 22)
    rng = np.random.default_rng(3)
    outs = np.concatenate([rng.normal(0, 0.2, (B, 6)), rng.uniform(2.0, 3.5, (B, 1))], 1)
    X, Y, status = learning.imitation_dataset(inp, outs)
    assert X.shape == (B * 50, 18) and Y.shape == (B * 50, 7) and (status <= 1).all()
    sc = learning.scenario_from_nn_inputs(inp)
    for b in range(B):
        pb = O.make_problem(sc["x0"][b], sc["goal"][b], outs[b, 0:3], O.rd2quat(outs[b, 3:6]), outs[b, 6], np.zeros(4))     # t NOT rounded (get_input)
        sol = O.solve(p, pb)
        assert np.abs(X[b * 50:(b + 1) * 50, 0:13] - sol["state_traj_opt"][0:50]).max() <= 1e-5
        assert np.array_equal(X[b * 50 + 7, 13:16], inp[b, 3:6]) and np.array_equal(X[b * 50 + 7, 16:18], inp[b, 7:9])
        assert np.array_equal(Y[b * 50:(b + 1) * 50, 0:6], np.tile(outs[b, 0:6], (50, 1)))
        assert np.abs(Y[b * 50:(b + 1) * 50, 6] - (outs[b, 6] - 0.1 * np.arange(50))).max() <= 1e-15
