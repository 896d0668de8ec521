"""Thin object wrapper over the egx_gp_* C ABI (device-resident GP context)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import (EGX_OK, EGX_CUDA_ERROR, EGX_INVALID_VALUE, GpuError, NUM_STAGES, STAGE_NAMES)

SQUARED_EXPONENTIAL, ABSOLUTE_EXPONENTIAL, MATERN32, MATERN52 = 0, 1, 2, 3
CONSTANT, LINEAR, QUADRATIC = 0, 1, 2
DEFAULT_NUGGET = 100.0 * np.finfo(np.float64).eps      # gp/src/parameters.rs:118


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class GpContext:
    """One training set resident on one GPU (egx_gp_ctx)."""

    def __init__(self, xnorm, ynorm, x_mean, x_std, y_mean, y_std, corr, mean, w_star=None,
                 nugget=DEFAULT_NUGGET, device=0):
        self._lib = _lib.load()
        xnorm = _f64(xnorm)
        ynorm = _f64(ynorm).reshape(-1)
        n, d = xnorm.shape
        w = np.eye(d) if w_star is None else _f64(w_star)
        h = w.shape[1]
        self.n, self.d, self.h = n, d, h
        self._h = C.c_void_p()
        st = self._lib.egx_gp_create(C.byref(self._h), device, corr, mean, _ptr(xnorm), n, d, _ptr(ynorm),
                                     _ptr(_f64(x_mean)), _ptr(_f64(x_std)), float(y_mean), float(y_std),
                                     _ptr(_f64(w)), h, float(nugget))
        if st != EGX_OK:
            raise GpuError(st, _lib.last_error())
        pp = C.c_int()
        self._lib.egx_gp_dims(self._h, None, None, None, C.byref(pp))
        self.p = pp.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.egx_gp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st in (EGX_CUDA_ERROR, EGX_INVALID_VALUE):
            raise GpuError(st, _lib.last_error())
        return st

    def _check_all(self, st):
        if st != 0:
            raise GpuError(st, _lib.last_error())
        return st

    def reduced_likelihood(self, theta):
        """-> (status, rlf); numerical failures are statuses (rlf = NaN)."""
        th = _f64(theta).reshape(-1)
        assert th.size == self.h
        out = C.c_double()
        st = self._check(self._lib.egx_gp_reduced_likelihood(self._h, _ptr(th), C.byref(out)))
        return st, out.value

    def reduced_likelihood_batch(self, thetas):
        th = _f64(thetas).reshape(-1, self.h)
        B = th.shape[0]
        rlf = np.empty(B)
        status = np.empty(B, dtype=np.int32)
        self._check(self._lib.egx_gp_reduced_likelihood_batch(
            self._h, _ptr(th), B, _ptr(rlf), status.ctypes.data_as(C.POINTER(C.c_int))))
        return status, rlf

    # asynchronous seam: independent chains, one slot each (egx_gp_async_slots / eval_begin / eval_end)
    def async_slots(self, wanted):
        return int(self._lib.egx_gp_async_slots(self._h, int(wanted)))

    def eval_begin(self, slot, theta):
        th = _f64(theta).reshape(-1)
        return int(self._lib.egx_gp_eval_begin(self._h, int(slot), _ptr(th)))

    def eval_end(self, slot):
        out = C.c_double()
        st = int(self._lib.egx_gp_eval_end(self._h, int(slot), C.byref(out)))
        return st, out.value

    def reduced_likelihood_grad(self, theta, rel_step=1e-6):
        """-> (status, rlf, d rlf / d theta) by one batch of 2h+1 evaluations (central differences)."""
        th = _f64(theta).reshape(-1)
        out = C.c_double()
        g = np.empty(self.h)
        st = self._check(self._lib.egx_gp_reduced_likelihood_grad(self._h, _ptr(th), float(rel_step), C.byref(out), _ptr(g)))
        return st, out.value, g

    def reduced_likelihood_grad_analytic(self, theta):
        """-> (status, rlf, d rlf / d theta) in closed form from one factorisation (h <= 32)."""
        th = _f64(theta).reshape(-1)
        out = C.c_double()
        g = np.empty(self.h)
        st = self._check(self._lib.egx_gp_reduced_likelihood_grad_analytic(self._h, _ptr(th), C.byref(out), _ptr(g)))
        return st, out.value, g

    def finalize(self, theta, want_ft=True):
        th = _f64(theta).reshape(-1)
        rlf, s2 = C.c_double(), C.c_double()
        beta = np.empty(self.p)
        gamma = np.empty(self.n)
        ft = np.empty((self.n, self.p)) if want_ft else None
        g = np.empty((self.p, self.p))
        st = self._check(self._lib.egx_gp_finalize(self._h, _ptr(th), C.byref(rlf), C.byref(s2), _ptr(beta),
                                                   _ptr(gamma), _ptr(ft) if want_ft else None, _ptr(g)))
        return st, dict(rlf=rlf.value, sigma2=s2.value, beta=beta, gamma=gamma, ft=ft, ft_qr_r=g)

    def download_chol(self):
        out = np.empty((self.n, self.n))
        self._check(self._lib.egx_gp_download_chol(self._h, _ptr(out)))
        return out

    def predict(self, x):
        x = _f64(x).reshape(-1, self.d)
        y = np.empty(x.shape[0])
        self._check(self._lib.egx_gp_predict(self._h, _ptr(x), x.shape[0], _ptr(y)))
        return y

    def predict_var(self, x):
        x = _f64(x).reshape(-1, self.d)
        v = np.empty(x.shape[0])
        self._check(self._lib.egx_gp_predict_var(self._h, _ptr(x), x.shape[0], _ptr(v)))
        return v

    def predict_valvar(self, x):
        x = _f64(x).reshape(-1, self.d)
        y = np.empty(x.shape[0])
        v = np.empty(x.shape[0])
        self._check(self._lib.egx_gp_predict_valvar(self._h, _ptr(x), x.shape[0], _ptr(y), _ptr(v)))
        return y, v

    def predict_gradients(self, x):
        x = _f64(x).reshape(-1, self.d)
        g = np.empty((x.shape[0], self.d))
        self._check(self._lib.egx_gp_predict_gradients(self._h, _ptr(x), x.shape[0], _ptr(g)))
        return g

    def predict_var_gradients(self, x):
        x = _f64(x).reshape(-1, self.d)
        g = np.empty((x.shape[0], self.d))
        self._check(self._lib.egx_gp_predict_var_gradients(self._h, _ptr(x), x.shape[0], _ptr(g)))
        return g

    def covariance(self, x):
        """Conditional covariance at the rows of x (algorithm.rs:310-326), (m, m)."""
        x = _f64(x).reshape(-1, self.d)
        cov = np.empty((x.shape[0], x.shape[0]))
        self._check_all(self._lib.egx_gp_covariance(self._h, _ptr(x), x.shape[0], _ptr(cov)))
        return cov

    def sample(self, x, z, method=0):
        """mean + C z with C C^T the conditional covariance (algorithm.rs:1153-1194); z is (m, n_traj)
        standard-normal draws, method 0 = Cholesky, 1 = eigenvalues."""
        x = _f64(x).reshape(-1, self.d)
        z = _f64(z).reshape(x.shape[0], -1)
        out = np.empty_like(z)
        self._check_all(self._lib.egx_gp_sample(self._h, _ptr(x), x.shape[0], _ptr(z), z.shape[1], int(method),
                                                _ptr(out)))
        return out

    def predict_valvar_dev(self, x_ptr, m, y_ptr, v_ptr):
        """x/y/var are raw device addresses (ints) on this context's GPU."""
        self._check(self._lib.egx_gp_predict_valvar_dev(self._h, C.c_void_p(x_ptr), m,
                                                        C.c_void_p(y_ptr) if y_ptr else None,
                                                        C.c_void_p(v_ptr) if v_ptr else None))

    def correlation_matrix(self, theta):
        th = _f64(theta).reshape(-1)
        out = np.empty((self.n, self.n))
        self._check(self._lib.egx_gp_correlation_matrix(self._h, _ptr(th), _ptr(out)))
        return out

    def cross_correlation(self, x):
        x = _f64(x).reshape(-1, self.d)
        out = np.empty((x.shape[0], self.n))
        self._check(self._lib.egx_gp_cross_correlation(self._h, _ptr(x), x.shape[0], _ptr(out)))
        return out

    def set_profiling(self, on=True):
        self._lib.egx_gp_set_profiling(self._h, int(bool(on)))

    def reset_profile(self):
        self._lib.egx_gp_reset_profile(self._h)

    def profile(self):
        ms = np.zeros(NUM_STAGES)
        ln = np.zeros(NUM_STAGES, dtype=np.int64)
        self._lib.egx_gp_get_profile(self._h, _ptr(ms), ln.ctypes.data_as(C.POINTER(C.c_longlong)))
        return {STAGE_NAMES[i]: (float(ms[i]), int(ln[i])) for i in range(NUM_STAGES)}

    def timer_start(self):
        self._check(self._lib.egx_gp_timer_start(self._h))

    def timer_stop(self):
        """Elapsed device milliseconds on the context's stream since timer_start()."""
        ms = C.c_double()
        self._check(self._lib.egx_gp_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_lookahead(self, on=True):
        self._lib.egx_gp_set_lookahead(self._h, int(bool(on)))

    def set_force_blocked(self, on=True):
        self._lib.egx_gp_set_force_blocked(self._h, int(bool(on)))
