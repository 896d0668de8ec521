"""PLS rotations of the KPLS option (gp/src/algorithm.rs:843-855 -> linfa-pls 0.8.0, a port of
scikit-learn's PLSRegression).  The oracle restatement and the C++ host routine behind the C ABI
(`egx_pls_rotations`) are both pinned on scikit-learn's own x_rotations_ (tests/golden/
pls_rotations.json, written by tests/golden/make_pls_golden.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import pls_oracle as P

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "pls_rotations.json")))
_dp = C.POINTER(C.c_double)


def _c_rotations(x, y, k):
    from egobox_b200 import _lib
    lib = _lib.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    w = np.full((x.shape[1], k), np.nan)
    st = lib.egx_pls_rotations(x.ctypes.data_as(_dp), x.shape[0], x.shape[1], y.ctypes.data_as(_dp), k,
                               w.ctypes.data_as(_dp))
    return st, w


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_oracle_matches_sklearn(case):
    r = P.pls_rotations(np.array(case["x"]), np.array(case["y"]), case["k"])
    np.testing.assert_allclose(r, np.array(case["x_rotations"]), rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_c_abi_matches_sklearn_and_oracle(case):
    x, y, k = np.array(case["x"]), np.array(case["y"]), case["k"]
    st, w = _c_rotations(x, y, k)
    assert st == 0
    np.testing.assert_allclose(w, np.array(case["x_rotations"]), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(w, P.pls_rotations(x, y, k), rtol=1e-10, atol=1e-12)


def test_constant_target_gives_zero_rotations():
    # algorithm.rs:846-851: PowerMethodConstantResidualError -> zeros
    rng = np.random.default_rng(0)
    x = rng.random((20, 4))
    y = np.full(20, 3.25)
    assert np.all(P.kpls_w_star(x, y, 2) == 0.0)
    st, w = _c_rotations(x, y, 2)
    assert st == 0 and np.all(w == 0.0)


def test_exhausted_residual_gives_zero_rotations():
    # y is exactly linear in one input: the residual vanishes after the first component
    rng = np.random.default_rng(1)
    x = rng.random((30, 3))
    y = 2.0 * x[:, 0]
    x[:, 1:] = x[:, 1:] - x[:, 1:].mean(axis=0)
    # make the other columns orthogonal to y so that one component explains y completely
    x0 = x[:, 0] - x[:, 0].mean()
    for j in (1, 2):
        x[:, j] -= x0 * (x0 @ x[:, j]) / (x0 @ x0)
    ref = P.kpls_w_star(x, y, 2)
    st, w = _c_rotations(x, y, 2)
    assert st == 0
    np.testing.assert_allclose(w, ref, rtol=1e-9, atol=1e-12)


def test_invalid_arguments():
    x = np.zeros((5, 2))
    y = np.zeros(5)
    st, _ = _c_rotations(x, y, 3)          # more components than inputs
    assert st == 4                          # EGX_INVALID_VALUE
