"""Fast CPU legs of the oracle (C/OpenMP correlation + LAPACK) -- TEST INFRASTRUCTURE and
bench.py's cpu_baseline / --impl reference only.  Same arithmetic as gp_oracle.py."""
import ctypes as C
import os

import numpy as np

from . import gp_oracle as O

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        from . import build_oracle
        path = build_oracle.OUT
        if not os.path.exists(path):
            build_oracle.build()
        lib = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        lib.egx_oracle_corr_matrix.argtypes = [C.c_int, dp, C.c_int, C.c_int, dp, dp, C.c_int, C.c_double, dp]
        lib.egx_oracle_corr_matrix.restype = None
        lib.egx_oracle_cross_corr.argtypes = [C.c_int, dp, C.c_int, dp, C.c_int, C.c_int, dp, dp, C.c_int, dp]
        lib.egx_oracle_cross_corr.restype = None
        lib.egx_oracle_set_threads.argtypes = [C.c_int]
        lib.egx_oracle_set_threads.restype = None
        lib.egx_oracle_max_threads.restype = C.c_int
        _LIB = lib
    return _LIB


def set_threads(n):
    """Use n OpenMP threads in the C kernels (torchrun exports OMP_NUM_THREADS=1)."""
    _lib().egx_oracle_set_threads(int(n))
    return int(_lib().egx_oracle_max_threads())


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def corr_matrix(kind, xnorm, theta, w, nugget=O.DEFAULT_NUGGET):
    x = np.ascontiguousarray(xnorm, dtype=np.float64)
    th = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1)
    w = np.ascontiguousarray(w, dtype=np.float64)
    n, d = x.shape
    R = np.empty((n, n))
    _lib().egx_oracle_corr_matrix(kind, _p(x), n, d, _p(th), _p(w), th.size, float(nugget), _p(R))
    return R


def cross_corr(kind, xs_norm, xnorm, theta, w):
    xs = np.ascontiguousarray(xs_norm, dtype=np.float64)
    x = np.ascontiguousarray(xnorm, dtype=np.float64)
    th = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1)
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.empty((xs.shape[0], x.shape[0]))
    _lib().egx_oracle_cross_corr(kind, _p(xs), xs.shape[0], _p(x), x.shape[0], x.shape[1], _p(th), _p(w), th.size,
                                 _p(out))
    return out


def reduced_likelihood(kind, xnorm, fx, ynorm, y_std, theta, w, nugget=O.DEFAULT_NUGGET):
    """objfn body + reduced_likelihood (algorithm.rs:892-893, 989-1056) with the C correlation build."""
    R = corr_matrix(kind, xnorm, theta, w, nugget)
    return O.reduced_likelihood_from_R(fx, R, np.asarray(ynorm).reshape(-1, 1), y_std, copy=False)


def predict_valvar(gp, x, chunk=1024):
    """GaussianProcess.predict_valvar (algorithm.rs:282-307) with the C cross-correlation."""
    import scipy.linalg as sla
    ys, vs = [], []
    x = np.asarray(x, dtype=np.float64)
    for i0 in range(0, x.shape[0], chunk):
        xn = gp._xnorm(x[i0:i0 + chunk])
        f = O.mean_value(gp.mean, xn)
        corr = cross_corr(gp.corr, xn, gp.xt_norm, gp.theta, gp.w_star)
        y_ = f.dot(gp.inner.beta) + corr.dot(gp.inner.gamma)
        ys.append((y_ * gp.y_std + gp.y_mean)[:, 0])
        rt = sla.solve_triangular(gp.inner.r_chol, corr.T, lower=True, check_finite=False)
        rhs = gp.inner.ft.T.dot(rt) - f.T
        u = sla.solve_triangular(gp.inner.ft_qr_r.T, rhs, lower=True, check_finite=False)
        mse = gp.inner.sigma2 * (1.0 - (rt * rt).sum(axis=0) + (u * u).sum(axis=0))
        vs.append(np.where(mse < 0.0, 0.0, mse))
    return np.concatenate(ys), np.concatenate(vs)
