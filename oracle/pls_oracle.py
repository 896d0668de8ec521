"""CPU oracle for the PLS rotations of the KPLS option  --  TEST INFRASTRUCTURE ONLY.

Reference call site: ``PlsRegression::params(n_components).fit(&ds)`` followed by
``.rotations().0`` at crates/gp/src/algorithm.rs:843-855 and
crates/gp/src/sparse_algorithm.rs:442-455.  The arithmetic lives in the third-party
crate ``linfa-pls 0.8.0`` (Cargo.lock:1250-1251), which is NOT vendored under
/root/reference; it is a port of scikit-learn's ``PLSRegression`` (NIPALS, scale=true,
max_iter 500, tol 1e-6, deflation mode "regression").  This file restates that published
algorithm in numpy; it is pinned by ``tests/test_pls.py`` against scikit-learn's own
``PLSRegression.x_rotations_`` (fixtures ``tests/golden/pls_rotations.json`` written by
``tests/golden/make_pls_golden.py``).  Only ``tests/`` may import it.
"""
from __future__ import annotations

import numpy as np

EPS = np.finfo(np.float64).eps


class ConstantResidualError(Exception):
    """linfa_pls::PlsError::PowerMethodConstantResidualError -- the reference maps it to an
    all-zero w_star (algorithm.rs:846-851)."""


def center_scale(a):
    """center_scale_dataset with scale = true: column mean, std with ddof = 1, zero std -> 1."""
    a = np.asarray(a, dtype=np.float64)
    mean = a.mean(axis=0)
    std = a.std(axis=0, ddof=1)
    std = np.where(std == 0.0, 1.0, std)
    return (a - mean) / std, mean, std


def _first_singular_vectors_power_method(x, y, max_iter=500, tol=1e-6):
    y_score = None
    for j in range(y.shape[1]):
        if np.abs(y[:, j]).max() > EPS:
            y_score = y[:, j].copy()
            break
    if y_score is None:
        raise ConstantResidualError()
    x_weights_old = np.full(x.shape[1], 100.0)
    for _ in range(max_iter):
        x_weights = x.T @ y_score / (y_score @ y_score)
        x_weights = x_weights / (np.sqrt(x_weights @ x_weights) + EPS)
        x_score = x @ x_weights
        y_weights = y.T @ x_score / (x_score @ x_score)
        y_score = (y @ y_weights) / (y_weights @ y_weights + EPS)
        diff = x_weights - x_weights_old
        if diff @ diff < tol or y.shape[1] == 1:
            break
        x_weights_old = x_weights
    return x_weights, y_weights


def _svd_flip_1d(u, v):
    """sign convention of the weight vectors: the largest |u| component is positive."""
    k = int(np.argmax(np.abs(u)))
    s = np.sign(u[k])
    if s == 0.0:
        s = 1.0
    return u * s, v * s


def pls_rotations(x, y, n_components):
    """x_rotations (nx x n_components) of a NIPALS PLS regression of y (n, or n x 1) on x (n x nx).
    Raises ConstantResidualError where linfa-pls returns PowerMethodConstantResidualError."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y[:, None]
    nx = x.shape[1]
    if not 1 <= n_components <= nx:
        raise ValueError("n_components")
    xk, _, _ = center_scale(x)
    yk, _, _ = center_scale(y)
    xw = np.zeros((nx, n_components))
    xl = np.zeros((nx, n_components))
    for k in range(n_components):
        # columns of the y residual that are numerically zero are set to zero
        small = np.abs(yk).max(axis=0) < 10 * EPS
        yk[:, small] = 0.0
        w, c = _first_singular_vectors_power_method(xk, yk)
        w, c = _svd_flip_1d(w, c)
        t = xk @ w
        tt = t @ t
        p = (t @ xk) / tt
        xk = xk - np.outer(t, p)
        q = (t @ yk) / tt
        yk = yk - np.outer(t, q)
        xw[:, k] = w
        xl[:, k] = p
    return xw @ np.linalg.pinv(xl.T @ xw)


def kpls_w_star(x, y, n_components):
    """What gp `fit` stores as w_star (algorithm.rs:843-855): the rotations, or zeros when the
    power method reports a constant residual."""
    try:
        return pls_rotations(x, y, n_components)
    except ConstantResidualError:
        return np.zeros((np.asarray(x).shape[1], n_components))
