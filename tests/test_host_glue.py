"""CPU tests of the C-side host glue behind b200mpc_controller_step (judo_b200/csrc/host_glue.cpp): the sampler must reproduce
NumPy's legacy global stream bit for bit (the reference draws its candidates from it: judo/optimizers/mppi.py:58), the spline basis
must equal judo_b200.spline (pinned to scipy / the reference's goldens in test_host.py)."""
import numpy as np
import pytest

from judo_b200 import _lib
from judo_b200.engine import LegacyStream, legacy_stream
from judo_b200.spline import spline_basis


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 127, 128, 4095 * 4, 16383, 70001])
@pytest.mark.parametrize("predraw", [0, 1, 2, 3])
def test_legacy_stream_randn_is_numpy_randn(n, predraw):
    """Same values, same generator state afterwards (incl. the cached second gaussian), whatever the cache held on entry."""
    np.random.seed(1234 + n)
    np.random.randn(predraw)
    ref = np.random.randn(n)
    ref_after = np.random.randn(5), np.random.rand(3)
    np.random.seed(1234 + n)
    np.random.randn(predraw)
    got = legacy_stream().randn(n)
    got_after = np.random.randn(5), np.random.rand(3)
    assert np.array_equal(ref, got)
    assert np.array_equal(ref_after[0], got_after[0]) and np.array_equal(ref_after[1], got_after[1])


def test_legacy_normals_long_run_and_reseed():
    s = legacy_stream()
    for seed in (0, 42, 2**32 - 1):
        np.random.seed(seed)
        a = np.concatenate([np.random.randn(1000) for _ in range(7)])
        np.random.seed(seed)
        b = np.concatenate([s.randn(1000) for _ in range(7)])
        assert np.array_equal(a, b)
    np.random.seed(7); a = np.random.randn(2_000_000)
    np.random.seed(7); b = s.randn(2_000_000)
    assert np.array_equal(a, b)


def test_stream_view_detects_a_replaced_generator():
    s = LegacyStream()
    assert s.ok and s.current()
    st = np.random.get_state()
    np.random.set_state(st)      # in place: the view stays valid
    assert s.current()


@pytest.mark.parametrize("order", ["zero", "linear", "cubic"])
@pytest.mark.parametrize("K", [4, 5, 12])
def test_c_spline_basis_matches_python(order, K):
    lib = _lib.load()
    rng = np.random.default_rng(K)
    t = np.cumsum(0.05 + rng.random(K))
    q = np.concatenate([[t[0] - 0.3, t[0], t[-1], t[-1] + 0.2], t[1:-1], t[0] + (t[-1] - t[0]) * rng.random(60)])
    B = np.empty((len(q), K))
    assert lib.b200mpc_spline_basis({"zero": 0, "linear": 1, "cubic": 2}[order], t.ctypes.data, K, q.ctypes.data, len(q), B.ctypes.data) == 0
    ref = spline_basis(t, q, order)
    if order == "cubic":
        np.testing.assert_allclose(B, ref, rtol=0, atol=5e-14)
    else:
        assert np.array_equal(B, ref)


def test_c_spline_basis_rejects_bad_requests():
    lib = _lib.load()
    t = np.linspace(0, 1, 3); q = np.zeros(2); B = np.empty((2, 3))
    assert lib.b200mpc_spline_basis(2, t.ctypes.data, 3, q.ctypes.data, 2, B.ctypes.data) != 0   # cubic needs 4 knots
    assert lib.b200mpc_spline_basis(5, t.ctypes.data, 3, q.ctypes.data, 2, B.ctypes.data) != 0
