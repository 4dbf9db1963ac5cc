"""State order of the DMMA factor sweep (csrc/se3mpc_core.cuh: xpos / xidx / ab_block) against the oracle's Jacobians: every
(k-step, column tile) product the kernel skips must be structurally zero in A = I + dt df/dx and B = dt df/du (reference dynamics:
quad_model.py:35-213, Euler step quad_OC.py:97-101).  If the model gains a coupling that the skip list does not know -- drag, say --
this fails on the CPU tier, before any GPU parity test has to notice wrong Newton steps."""
import numpy as np
import pytest

from tests.emu import emu


@pytest.fixture(scope="module")
def maps():
    L = emu.lib()
    xpos = [L.se3emu_xpos(i) for i in range(13)]
    xidx = [L.se3emu_xidx(p) for p in range(16)]
    return L, xpos, xidx


def test_slot_map_is_a_bijection_with_three_padding_slots(maps):
    L, xpos, xidx = maps
    assert sorted(xpos) == sorted(set(xpos)) and all(0 <= p < 16 for p in xpos)
    assert all(xidx[xpos[i]] == i for i in range(13))
    assert sorted(p for p in range(16) if xidx[p] < 0) == sorted(set(range(16)) - set(xpos))
    if L.se3emu_perm():
        # k-step (kt, ks) of the register-resident contraction covers the slots 8 kt + ks + {0, 2, 4, 6}: one component group each
        groups = [range(0, 3), range(3, 6), range(6, 10), range(10, 13)]
        for ks4, comp in enumerate(groups):
            kt, ks = divmod(ks4, 2)
            slots = {8 * kt + ks + 2 * t for t in range(4)}
            assert {xpos[i] for i in comp} <= slots
            assert {xidx[p] for p in slots if xidx[p] >= 0} == set(comp)


def test_skipped_products_are_structurally_zero(maps, oracle):
    L, xpos, xidx = maps
    if not L.se3emu_perm():
        pytest.skip("natural state order: no product is skipped")
    O, p, rng = oracle, oracle.default_params(), np.random.default_rng(7)
    kstep = lambda slot: 2 * (slot // 8) + (slot & 1)          # noqa: E731  k-step that covers a slot
    nz = np.zeros((4, 3), dtype=bool)                            # (k-step, column tile) blocks of [A | B] with a non-zero entry somewhere
    for _ in range(50):
        x = rng.normal(size=13) * 2.0
        x[6:10] /= np.linalg.norm(x[6:10])
        u = rng.uniform(0.0, 4.0, 4)
        A, B = O.dyn_jac(p, x, u)          # Jacobians of the Euler step x + dt f(x, u): A = I + dt df/dx, B = dt df/du
        for j in range(13):                                      # row j of A: the equation of component j (contraction index of the products)
            for i in range(13):
                if A[j, i] != 0.0:
                    nz[kstep(xpos[j]), xpos[i] // 8] = True
            if np.any(B[j] != 0.0):
                nz[kstep(xpos[j]), 2] = True
    for ks4 in range(4):
        for tile in range(3):
            if nz[ks4, tile]:
                assert L.se3emu_ab_block(tile, ks4), f"the kernel skips the non-zero block k-step {ks4} x column tile {tile}"
    # and the list is tight: the kernel executes no block that is always zero
    assert sum(L.se3emu_ab_block(t, k) for t in range(3) for k in range(4)) == int(nz.sum())


def test_rv_columns_of_A_hold_only_one_and_dt(maps, oracle):
    """SE3_RV_COLS / SE3_RV_ROWS replace eight DMMAs per stage by register arithmetic; they rely on column r_t of A = I + dt df/dx being the
    unit vector e_{r_t}, column v_t being dt e_{r_t} + e_{v_t}, and on (r_t, v_t) sitting in neighbouring slots (2 t, 2 t + 1) of tile 0."""
    L, xpos, xidx = maps
    if not L.se3emu_perm():
        pytest.skip("natural state order")
    O, p, rng = oracle, oracle.default_params(), np.random.default_rng(11)
    for t in range(3):
        assert xpos[t] == 2 * t and xpos[3 + t] == 2 * t + 1
    for _ in range(20):
        x = rng.normal(size=13) * 3.0
        x[6:10] /= np.linalg.norm(x[6:10])
        u = rng.uniform(0.0, 4.0, 4)
        A, _ = O.dyn_jac(p, x, u)
        for t in range(3):
            er, ev = np.zeros(13), np.zeros(13)
            er[t] = 1.0
            ev[t], ev[3 + t] = p.dt, 1.0
            assert np.array_equal(A[:, t], er)
            assert np.array_equal(A[:, 3 + t], ev)
