"""The CPU SIMT emulator's own semantics (tests/warpsim): warp collectives, block barriers with a lock-step loop, the mbarrier / bulk-copy
emulation, and the property the kernel tests rely on to catch missing barriers — a result that depends on the lane order."""
import ctypes

import numpy as np
import pytest

P = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731


@pytest.fixture(scope="module")
def sim():
    from tests import warpsim

    return warpsim.lib()


@pytest.mark.parametrize("reverse", [0, 1])
def test_warp_collectives(sim, reverse):
    rng = np.random.default_rng(0)
    nb, nw = 3, 2
    x = rng.random((nb * nw, 32))
    s, b, sc, al = np.zeros(nb * nw), np.zeros(nb * nw, dtype=np.uint32), np.zeros((nb * nw, 32)), np.zeros(nb * nw, dtype=np.int32)
    sim.sim_selftest_warp(P(x), nb, nw, P(s), P(b), P(sc), P(al), reverse)
    np.testing.assert_allclose(s, x.sum(1), rtol=1e-14)
    np.testing.assert_allclose(sc, np.cumsum(x, 1), rtol=1e-14)
    for w in range(nb * nw):
        mask = sum(1 << i for i in range(32) if x[w, i] > 0.5)
        assert b[w] == mask and al[w] == bin(mask).count("1") + 1000


@pytest.mark.parametrize("reverse", [0, 1])
def test_block_barriers_and_lockstep_loop(sim, reverse):
    nb, nw = 2, 4
    x = np.arange(nb * nw * 32, dtype=np.float64)
    out, rounds = np.zeros(nb), np.zeros(nb * nw, dtype=np.int32)
    sim.sim_selftest_block(P(x), nb, nw, P(out), P(rounds), reverse)
    np.testing.assert_array_equal(out, x.reshape(nb, -1).sum(1))
    # warp w needs w + 1 rounds of its own; every warp stays in the loop until the slowest (nw rounds) is done
    for blk in range(nb):
        for w in range(nw):
            assert rounds[blk * nw + w] == 100 * nw + (w + 1)


def test_missing_barrier_shows_up_as_lane_order_dependence(sim):
    want = np.roll(np.arange(32.0), -1)
    ok_f, ok_r, bad_f, bad_r = (np.zeros(32) for _ in range(4))
    sim.sim_selftest_neighbour(P(ok_f), 1, 0)
    sim.sim_selftest_neighbour(P(ok_r), 1, 1)
    np.testing.assert_array_equal(ok_f, want)
    np.testing.assert_array_equal(ok_r, want)
    sim.sim_selftest_neighbour(P(bad_f), 0, 0)
    sim.sim_selftest_neighbour(P(bad_r), 0, 1)
    assert not np.array_equal(bad_f, bad_r)  # the race is visible: forward and reverse lane order disagree


@pytest.mark.parametrize("reverse", [0, 1])
def test_bulk_copy_and_mbarrier_emulation(sim, reverse):
    src, out = np.arange(40.0), np.zeros(40)
    sim.sim_selftest_tma(P(src), P(out), 40, reverse)
    np.testing.assert_array_equal(out, 2 * src)
