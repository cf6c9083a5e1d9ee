"""Synthetic BoxQP instances and LP points for benchmarks (CPLEX is not needed for the bench).

Instance: symmetric integer Q_file_ij in U{-50..50} kept with probability `density` (the ranges of the
spar*.in files); arrays as the BoxQP reader builds them (cut_select_qp.py:313-326).
LP point: x ~ U(0,1), X_ij ~ U(max(0, x_i+x_j-1), min(x_i, x_j)) -- inside the McCormick box
(cut_select_qp.py:357-373); vars_values = [X upper-tri row-major | x] (cut_select_qp.py:547).
"""
import numpy as np


def instance(n, density, seed=7):
    rng = np.random.default_rng(seed)
    vals = rng.integers(-50, 51, size=(n, n))
    keep = rng.random((n, n)) < density
    U = np.triu(vals * keep)
    return (U + np.triu(U, 1).T).astype(np.float64)


def boxqp_arrays(Qf):
    """(Q_arr, Q_adj): Q = -Q_file, diagonal halved, upper triangle row-major; adjacency of non-zeros."""
    Qf = np.asarray(Qf, dtype=np.float64)
    n = Qf.shape[0]
    Qa = -Qf.copy()
    Qa[np.arange(n), np.arange(n)] /= 2.0
    return Qa[np.triu_indices(n, k=0)], (Qf != 0).astype(np.uint8)


def lp_point(n, seed=8):
    rng = np.random.default_rng(seed)
    x = rng.random(n)
    iu = np.triu_indices(n)
    xi, xj = x[iu[0]], x[iu[1]]
    lo = np.maximum(0.0, xi + xj - 1.0)
    hi = np.minimum(xi, xj)
    return np.concatenate([lo + rng.random(lo.size) * (hi - lo), x])
