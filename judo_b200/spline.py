"""Spline basis matrices: the reference's make_spline (scipy interp1d, judo/controller/controller.py:382-401) is
LINEAR in the knots, so   controls[n, t, :] = sum_k B[t, k] * knots[n, k, :]   with one (H, K) matrix shared by every
rollout.  The kernels consume B; this module builds it on the host (tiny) in closed form:

  zero    previous-knot hold, right-continuous (value at a knot is that knot), as interp1d(kind="zero")
  linear  piecewise linear
  cubic   not-a-knot cubic spline (interp1d(kind="cubic") == make_interp_spline(k=3))

Outside [t_0, t_{K-1}] the value is the first / last knot (fill_value=(first, last), bounds_error=False).
tests/test_spline.py pins this against scipy and against the golden vectors produced by the reference.
"""

from __future__ import annotations

import numpy as np

SPLINE_ORDERS = ("zero", "linear", "cubic")


def _cubic_not_a_knot_matrix(t: np.ndarray) -> np.ndarray:
    """Second-derivative operator: returns C (K,K) with m = C @ y the knot second derivatives."""
    K = len(t)
    h = np.diff(t)
    A = np.zeros((K, K))
    R = np.zeros((K, K))
    for i in range(1, K - 1):
        A[i, i - 1], A[i, i], A[i, i + 1] = h[i - 1], 2 * (h[i - 1] + h[i]), h[i]
        R[i, i - 1], R[i, i], R[i, i + 1] = 6 / h[i - 1], -6 / h[i - 1] - 6 / h[i], 6 / h[i]
    # not-a-knot: third derivative continuous across t_1 and t_{K-2}
    A[0, 0], A[0, 1], A[0, 2] = h[1], -(h[0] + h[1]), h[0]
    A[-1, -3], A[-1, -2], A[-1, -1] = h[-1], -(h[-2] + h[-1]), h[-2]
    return np.linalg.solve(A, R)


def spline_basis(times: np.ndarray, query: np.ndarray, order: str) -> np.ndarray:
    """B (len(query), K) such that interp1d(times, y, kind=order, ...)(query) == B @ y along the knot axis."""
    t = np.asarray(times, dtype=np.float64)
    q = np.asarray(query, dtype=np.float64)
    K, H = len(t), len(q)
    if order not in SPLINE_ORDERS:
        raise ValueError(f"spline_order must be one of {SPLINE_ORDERS}, got {order!r}")
    if order == "cubic" and K < 4:
        raise ValueError("cubic splines need at least 4 knots")
    # segment of every query: index of the last knot <= q, clamped to a valid interval; queries outside the knot span are pinned to
    # the first / last knot afterwards.  (Few, flat NumPy calls: at the reference's sizes this function is Python-overhead bound and
    # it sits twice on the plan step's critical path.)
    B = np.zeros((H, K))
    rows = np.arange(H)
    seg = np.searchsorted(t, q, side="right") - 1   # -1 below the first knot, K - 1 from the last knot on
    if order == "zero":
        B[rows, np.maximum(seg, 0)] = 1.0
        return B
    below, above = seg < 0, q > t[-1]
    seg = np.minimum(np.maximum(seg, 0), K - 2)
    t0, t1 = t[seg], t[seg + 1]
    h = t1 - t0
    a = (t1 - q) / h
    b = (q - t0) / h
    out = below | above
    if out.any():
        a[below], b[below] = 1.0, 0.0
        a[above], b[above] = 0.0, 1.0
    if order == "linear":
        B[rows, seg] = a
        B[rows, seg + 1] = b
        return B
    C = _cubic_not_a_knot_matrix(t)
    ca = (a**3 - a) * h * h / 6.0
    cb = (b**3 - b) * h * h / 6.0
    B[rows, seg] = a
    B[rows, seg + 1] = b
    B += ca[:, None] * C[seg] + cb[:, None] * C[seg + 1]
    return B
