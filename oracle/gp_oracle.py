"""CPU oracle for the egobox-gp kriging hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a numpy/scipy (fp64) restatement of the reference algorithm
(relf/egobox @ be16128, crates/gp).  It is the *checker* for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package ``egobox_b200``
never imports anything from ``oracle/``.

Pinning: validated by ``tests/test_oracle_golden.py`` against every tight
fixture the reference holds for this path (kernel known-answers
``correlation_models.rs:597-641,718-726``; ``utils.rs:150-242``;
``mean_models.rs:169-213``; 5-point kriging ``test_gpmix.py:37-53,137-142``;
the 16-digit serialized model in ``doc/Gpx_Tutorial.ipynb:165-167,420-421``).
Values at n > 300, predict_var beyond toy sizes and FITC/VFE likelihoods are
NOT pinned by any reference fixture ("parity unpinned" there).

The dense factorizations live in third-party crates that are not vendored in
/root/reference (linfa-linalg 0.2.1 default / ndarray-linalg 0.17 + LAPACK
with feature ``blas``); they are textbook Cholesky / triangular solve / thin
QR / SVD and are restated with LAPACK through scipy.  QR sign convention
follows linfa-linalg (diag(R) > 0), which is what the notebook fixture shows.

Citations are file:line under /root/reference/crates/gp/src/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla

SQEXP, ABSEXP, MATERN32, MATERN52 = 0, 1, 2, 3
CORR_NAMES = {SQEXP: "SquaredExponential", ABSEXP: "AbsoluteExponential",
              MATERN32: "Matern32", MATERN52: "Matern52"}
CONSTANT, LINEAR, QUADRATIC = 0, 1, 2
MEAN_NAMES = {CONSTANT: "Constant", LINEAR: "Linear", QUADRATIC: "Quadratic"}

DEFAULT_NUGGET = 100.0 * np.finfo(np.float64).eps      # parameters.rs:118
THETA_DEFAULT_INIT = 1e-1                                # parameters.rs:49
THETA_DEFAULT_BOUNDS = (1e-2, 1e1)                       # parameters.rs:51
GP_OPTIM_N_START = 10                                    # algorithm.rs:33
GP_COBYLA_MIN_EVAL = 25                                  # algorithm.rs:35
GP_COBYLA_MAX_EVAL = 1000                                # algorithm.rs:37


class LinalgError(Exception):
    """GpError::LinalgError (errors.rs:19): Cholesky of a non-PD matrix."""


class LikelihoodComputationError(Exception):
    """GpError::LikelihoodComputationError (errors.rs:11)."""


# --------------------------------------------------------------------------
# utils.rs
# --------------------------------------------------------------------------
def normalize(x):
    """utils.rs:45-54: column mean, std (ddof=1, 0 -> 1), (x - mean) / std."""
    x = np.asarray(x, dtype=np.float64)
    mean = x.mean(axis=0)
    std = x.std(axis=0, ddof=1)
    std = np.where(std == 0.0, 1.0, std)
    return (x - mean) / std, mean, std


def diff_matrix(x):
    """utils.rs:80-104: |x_k - x_i| for all k < i, row order (0,1),(0,2)...,
    plus the (k, i) index table."""
    x = np.asarray(x, dtype=np.float64)
    n, nx = x.shape
    npairs = n * (n - 1) // 2
    d = np.zeros((npairs, nx))
    idx = np.zeros((npairs, 2), dtype=np.int64)
    pos = 0
    for k in range(n - 1):
        cnt = n - k - 1
        d[pos:pos + cnt] = x[k] - x[k + 1:]
        idx[pos:pos + cnt, 0] = k
        idx[pos:pos + cnt, 1] = np.arange(k + 1, n)
        pos += cnt
    return np.abs(d), idx


def pairwise_differences(x, y):
    """utils.rs:110-131: (nx*ny, d) array, row i*ny+j = x_i - y_j."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    assert x.shape[1] == y.shape[1]
    return (x[:, None, :] - y[None, :, :]).reshape(-1, x.shape[1])


# --------------------------------------------------------------------------
# mean_models.rs
# --------------------------------------------------------------------------
def mean_value(kind, x):
    """mean_models.rs:42-44 (constant), :68-71 (linear), :97-104 (quadratic:
    [1, x_i, {x_k * x_j, j >= k} for k = 0..nx-1])."""
    x = np.asarray(x, dtype=np.float64)
    n, nx = x.shape
    if kind == CONSTANT:
        return np.ones((n, 1))
    res = np.concatenate([np.ones((n, 1)), x], axis=1)
    if kind == LINEAR:
        return res
    parts = [res]
    for k in range(nx):
        parts.append(x[:, k:] * x[:, k:k + 1])
    return np.concatenate(parts, axis=1)


def mean_nbasis(kind, nx):
    return {CONSTANT: 1, LINEAR: nx + 1, QUADRATIC: (nx + 1) * (nx + 2) // 2}[kind]


# --------------------------------------------------------------------------
# correlation_models.rs  (value only)
# --------------------------------------------------------------------------
def corr_value(kind, d, theta, w):
    """r(d; theta, W) for a (P, nx) array of component differences.

    SqExp   correlation_models.rs:91-104
    AbsExp  :185-196
    Matern32 :277-286, 326-353
    Matern52 :446-455, 497-523
    Returns shape (P,)."""
    d = np.asarray(d, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    w = np.asarray(w, dtype=np.float64)
    assert w.shape == (d.shape[1], theta.shape[0]), (w.shape, d.shape, theta.shape)
    if kind == SQEXP:
        theta_w = ((theta * w) ** 2).sum(axis=1)
        r = (d ** 2).dot(theta_w)
        return np.exp(-0.5 * r)
    if kind == ABSEXP:
        theta_w = np.abs(w).dot(theta)
        r = np.abs(d).dot(theta_w)
        return np.exp(-r)
    theta_w = theta * np.abs(w)                       # (nx, h)
    abs_d = np.abs(d)
    a = np.ones(d.shape[0])
    if kind == MATERN32:
        s = math.sqrt(3.0)
        for j in range(d.shape[1]):
            for l in range(theta_w.shape[1]):
                a *= 1.0 + s * theta_w[j, l] * abs_d[:, j]
        b = np.exp(-s * abs_d.dot(theta_w).sum(axis=1))
        return a * b
    if kind == MATERN52:
        s = math.sqrt(5.0)
        c = 5.0 / 3.0
        for j in range(d.shape[1]):
            for l in range(theta_w.shape[1]):
                v = theta_w[j, l]
                a *= 1.0 + s * v * abs_d[:, j] + c * (v * v * d[:, j] * d[:, j])
        b = np.exp(-s * abs_d.dot(theta_w).sum(axis=1))
        return a * b
    raise ValueError(kind)


def mean_jacobian(kind, x):
    """mean_models.rs:50-53 (constant), :75-82 (linear), :110-128 (quadratic): (p, nx) at one point x."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    nx = x.size
    if kind == CONSTANT:
        return np.zeros((1, nx))
    if kind == LINEAR:
        jac = np.zeros((nx + 1, nx))
        jac[1:, :] = np.eye(nx)
        return jac
    jac = np.zeros((1 + nx + nx * (nx + 1) // 2, nx))
    jac[1:nx + 1, :] = np.eye(nx)
    o, p = 1 + nx, nx
    for i in range(nx):
        part = np.zeros((p, p))
        part[:, 0] = x[i:]
        part = part + np.eye(p) * x[i]
        jac[o:o + nx - i, i:nx] = part
        o += p
        p -= 1
    return jac


def corr_jacobian(kind, x, xtrain, theta, w):
    """d r(x, xtrain_j) / d x_k as an (n, nx) array.
    SqExp correlation_models.rs:106-122, AbsExp :198-214, Matern32 :288-298 + 355-412,
    Matern52 :457-468 + 525-586."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    xtrain = np.asarray(xtrain, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    w = np.asarray(w, dtype=np.float64)
    d = x[None, :] - xtrain                                  # utils.rs:136-142 `differences`
    if kind == SQEXP:
        r = corr_value(kind, d, theta, w)
        dtheta_w = -((theta * w) ** 2).sum(axis=1)
        return d * dtheta_w * r[:, None]
    if kind == ABSEXP:
        r = corr_value(kind, d, theta, w)
        dtheta_w = np.copysign(1.0, d) * (-(theta * np.abs(w)).sum(axis=1))    # Rust f64::signum(+0.0) = 1
        return dtheta_w * r[:, None]
    s = math.sqrt(3.0) if kind == MATERN32 else math.sqrt(5.0)
    theta_w = theta * np.abs(w)                              # (nx, h)
    abs_d, sign_d = np.abs(d), np.copysign(1.0, d)
    a = np.ones(d.shape[0])
    fac = np.empty((d.shape[0],) + theta_w.shape)            # f_jl per pair
    for j in range(theta_w.shape[0]):
        for l in range(theta_w.shape[1]):
            v = theta_w[j, l] * abs_d[:, j]
            fac[:, j, l] = 1.0 + s * v if kind == MATERN32 else 1.0 + s * v + (5.0 / 3.0) * v * v
            a *= fac[:, j, l]
    b = np.exp(-s * abs_d.dot(theta_w).sum(axis=1))
    tw = np.abs(w).dot(theta)                                # (nx,)
    db = -s * tw[None, :] * sign_d * (a * b)[:, None]
    da = np.zeros_like(d)
    for j in range(theta_w.shape[0]):
        for k in range(theta_w.shape[1]):
            if kind == MATERN32:
                deriv = s * theta_w[j, k] * sign_d[:, j]
            else:
                deriv = s * theta_w[j, k] * sign_d[:, j] + (10.0 / 3.0) * theta_w[j, k] ** 2 * sign_d[:, j] * abs_d[:, j]
            term = np.ones(d.shape[0])
            for p_ in range(theta_w.shape[0]):
                for l in range(theta_w.shape[1]):
                    if l != k or p_ != j:
                        term = term * fac[:, p_, l]
            da[:, j] += deriv * term
    return db + b[:, None] * da


def corr_matrix(kind, xnorm, theta, w, nugget=DEFAULT_NUGGET, chunk=256):
    """R = (1+nugget) I + symmetric scatter of r over pairs.

    Same values as DiffMatrix + value + the scatter at algorithm.rs:997-1001
    but built row-block by row-block so that the (P, nx) difference table
    (2.68 GB at n=8192, d=10) is never materialised."""
    x = np.asarray(xnorm, dtype=np.float64)
    n = x.shape[0]
    R = np.empty((n, n))
    for i0 in range(0, n, chunk):
        i1 = min(n, i0 + chunk)
        dx = pairwise_differences(x[i0:i1], x)
        R[i0:i1] = corr_value(kind, np.abs(dx), theta, w).reshape(i1 - i0, n)
    R[np.diag_indices(n)] = 1.0 + nugget
    return R


def corr_theta_log_derivatives(kind, d, theta, w):
    """d ln r / d theta_l for a (P, nx) array of component differences -> (P, h).  Not in the reference
    (`CorrelationModel` has no theta derivative; algorithm.rs:880 ignores `_gradient`): differentiates the four
    `value` formulas restated in corr_value, term by term."""
    d = np.asarray(d, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    w = np.asarray(w, dtype=np.float64)
    abs_d = np.abs(d)
    if kind == SQEXP:                                   # ln r = -1/2 sum_j d_j^2 sum_l (theta_l W_jl)^2
        return -(d ** 2).dot(w ** 2) * theta[None, :]
    if kind == ABSEXP:                                  # ln r = -sum_j |d_j| sum_l |W_jl| theta_l
        return -abs_d.dot(np.abs(w))
    s = math.sqrt(3.0) if kind == MATERN32 else math.sqrt(5.0)
    out = np.zeros((d.shape[0], theta.size))
    for j in range(d.shape[1]):
        for l in range(theta.size):
            a = abs(w[j, l]) * abs_d[:, j]              # dv/dtheta_l with v = theta_l |W_jl| |d_j|
            v = theta[l] * a
            if kind == MATERN32:                        # d/dv [ln(1 + s v) - s v] = -3 v / (1 + s v)
                out[:, l] += a * (-3.0 * v / (1.0 + s * v))
            else:                                       # d/dv [ln(1 + s v + 5/3 v^2) - s v]
                out[:, l] += a * (-(5.0 / 3.0) * v * (1.0 + s * v) / (1.0 + s * v + (5.0 / 3.0) * v * v))
    return out


def reduced_likelihood_grad(kind, xnorm, fx, ynorm, y_std, theta, w, nugget=DEFAULT_NUGGET, chunk=256):
    """(rlf, d rlf / d theta) in closed form.  With rlf = -n log10 sigma2 - log10 det R (algorithm.rs:1039-1043), beta
    the generalised least-squares minimiser (so its own derivative drops out) and gamma = R^-1 (y - F beta) (:1034):
        d rlf / d theta_l = [ gamma^T (dR/dtheta_l) gamma / sigma2 - tr(R^-1 dR/dtheta_l) ] / ln 10.
    The diagonal of R is the constant 1 + nugget, so only pairs i != j contribute."""
    x = np.asarray(xnorm, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64).reshape(-1)
    n = x.shape[0]
    rlf, inner = reduced_likelihood(kind, x, fx, ynorm, y_std, theta, w, nugget)
    gamma = inner.gamma.reshape(-1)
    sigma2 = inner.sigma2 / (y_std * y_std)
    rinv = sla.cho_solve((inner.r_chol, True), np.eye(n), check_finite=False)
    weight = np.outer(gamma, gamma) / sigma2 - rinv
    np.fill_diagonal(weight, 0.0)
    grad = np.zeros(theta.size)
    for i0 in range(0, n, chunk):
        i1 = min(n, i0 + chunk)
        dx = pairwise_differences(x[i0:i1], x)
        r = corr_value(kind, np.abs(dx), theta, w)
        dlog = corr_theta_log_derivatives(kind, dx, theta, w)
        grad += (weight[i0:i1].reshape(-1) * r).dot(dlog)
    return rlf, grad / math.log(10.0)


# --------------------------------------------------------------------------
# algorithm.rs: reduced likelihood
# --------------------------------------------------------------------------
@dataclass
class InnerParams:
    """GpInnerParams, algorithm.rs:47-60."""
    sigma2: float
    beta: np.ndarray      # (p, 1)
    gamma: np.ndarray     # (n, 1)
    r_chol: np.ndarray    # (n, n) lower
    ft: np.ndarray        # (n, p)
    ft_qr_r: np.ndarray   # (p, p) upper, diag > 0


def _qr_pos(a):
    """Thin QR with diag(R) > 0 (linfa-linalg / nalgebra convention)."""
    q, r = np.linalg.qr(a, mode="reduced")
    sgn = np.sign(np.diag(r))
    sgn[sgn == 0] = 1.0
    return q * sgn, (r.T * sgn).T


def reduced_likelihood_from_R(fx, R, ynorm, y_std, copy=True):
    """algorithm.rs:1002-1055 given the assembled correlation matrix R."""
    n = R.shape[0]
    try:
        r_chol = sla.cholesky(R, lower=True, overwrite_a=not copy, check_finite=False)
    except sla.LinAlgError as e:                               # :1004 `?`
        raise LinalgError(str(e))
    if not np.all(np.isfinite(np.diag(r_chol))):
        raise LinalgError("non finite Cholesky factor")
    ft = sla.solve_triangular(r_chol, fx, lower=True, check_finite=False)      # :1006
    q, g = _qr_pos(ft)                                                          # :1007
    sv = np.linalg.svd(g, compute_uv=False)                                     # :1010
    cond_ft = sv[-1] / sv[0]
    if cond_ft < 1e-10:                                                         # :1012
        sv_f = np.linalg.svd(fx, compute_uv=False)
        cond_fx = sv_f[0] / sv_f[-1]
        if cond_fx > 1e15:
            raise LikelihoodComputationError("F is too ill conditioned")
        raise LikelihoodComputationError("ft is too ill conditioned")
    yt = sla.solve_triangular(r_chol, ynorm, lower=True, check_finite=False)    # :1028
    beta = sla.solve_triangular(g, q.T.dot(yt), lower=False, check_finite=False)  # :1030
    rho = yt - ft.dot(beta)                                                     # :1031
    rho_sqr = (rho * rho).sum(axis=0)
    gamma = sla.solve_triangular(r_chol.T, rho, lower=False, check_finite=False)  # :1034
    logdet = np.log10(np.diag(r_chol)).sum() * 2.0 / n                          # :1039
    sigma2 = rho_sqr / n                                                        # :1042
    with np.errstate(divide="ignore", invalid="ignore"):
        rlf = -n * (np.log10(sigma2.sum()) + logdet)                            # :1043
    inner = InnerParams(sigma2=float(sigma2[0] * y_std * y_std), beta=beta, gamma=gamma,
                        r_chol=np.tril(r_chol), ft=ft, ft_qr_r=g)
    return float(rlf), inner


def reduced_likelihood(kind, xnorm, fx, ynorm, y_std, theta, w, nugget=DEFAULT_NUGGET):
    """objfn body (algorithm.rs:892-893) + reduced_likelihood (:989-1056)."""
    R = corr_matrix(kind, xnorm, theta, w, nugget)
    return reduced_likelihood_from_R(fx, R, np.asarray(ynorm).reshape(-1, 1), y_std, copy=False)


def objective(kind, xnorm, fx, ynorm, y_std, theta, w, nugget=DEFAULT_NUGGET):
    """-rlf with the Err -> +inf, NaN -> +inf mapping of algorithm.rs:880-897."""
    theta = np.asarray(theta, dtype=np.float64)
    if np.any(np.isnan(theta)):
        return math.inf
    try:
        rlf, _ = reduced_likelihood(kind, xnorm, fx, ynorm, y_std, theta, w, nugget)
    except (LinalgError, LikelihoodComputationError):
        return math.inf
    return -rlf


# --------------------------------------------------------------------------
# optimization.rs
# --------------------------------------------------------------------------
def lhs_classic(xlimits, ns, rng):
    """LHS *construction* of crates/doe/src/lhs.rs (one uniform draw per
    stratum, independent permutation per column).  The reference's maximin
    optimisation and its Xoshiro256Plus stream are not reproduced."""
    xlimits = np.asarray(xlimits, dtype=np.float64)
    nx = xlimits.shape[0]
    cut = np.linspace(0.0, 1.0, ns + 1)
    u = rng.random((ns, nx))
    pts = cut[:ns, None] + u * (cut[1:, None] - cut[:ns, None])
    out = np.empty_like(pts)
    for j in range(nx):
        out[:, j] = pts[rng.permutation(ns), j]
    return xlimits[:, 0] + out * (xlimits[:, 1] - xlimits[:, 0])


def prepare_multistart(n_start, theta0, bounds, rng=None):
    """optimization.rs:26-71: log10 bounds; row 0 = log10(theta0); rows 1.. =
    LHS points in the log10 box (reference: maximin LHS, Xoshiro seed 42)."""
    bounds = [(math.log10(lo), math.log10(hi)) for lo, hi in bounds]
    theta0 = np.asarray(theta0, dtype=np.float64)
    theta0s = np.zeros((n_start + 1, theta0.size))
    theta0s[0] = np.log10(theta0)
    rng = np.random.default_rng(42) if rng is None else rng
    if n_start == 1:
        theta0s[1] = [rng.uniform(a, b) for a, b in bounds]
    elif n_start > 1:
        theta0s[1:] = lhs_classic(np.array(bounds), n_start, rng)
    return theta0s, bounds


# --------------------------------------------------------------------------
# algorithm.rs: GaussianProcess
# --------------------------------------------------------------------------
@dataclass
class GaussianProcess:
    """Trained state, algorithm.rs:174-192."""
    corr: int
    mean: int
    theta: np.ndarray
    likelihood: float
    inner: InnerParams
    w_star: np.ndarray
    xt_norm: np.ndarray
    x_mean: np.ndarray
    x_std: np.ndarray
    yt_norm: np.ndarray
    y_mean: float
    y_std: float
    training_data: tuple = field(default=None, repr=False)

    # algorithm.rs:372-380
    def _compute_correlation(self, xnorm):
        dx = pairwise_differences(xnorm, self.xt_norm)
        r = corr_value(self.corr, dx, self.theta, self.w_star)
        return r.reshape(xnorm.shape[0], self.xt_norm.shape[0])

    # algorithm.rs:330-369
    def _compute_rt_u(self, xnorm, corr):
        rt = sla.solve_triangular(self.inner.r_chol, corr.T, lower=True, check_finite=False)
        rhs = self.inner.ft.T.dot(rt) - mean_value(self.mean, xnorm).T
        u = sla.solve_triangular(self.inner.ft_qr_r.T, rhs, lower=True, check_finite=False)
        return rt, u

    def _xnorm(self, x):
        x = np.asarray(x, dtype=np.float64)
        return (x - self.x_mean) / self.x_std

    # algorithm.rs:253-263
    def predict(self, x, chunk=1024):
        out = []
        x = np.asarray(x, dtype=np.float64)
        for i0 in range(0, x.shape[0], chunk):
            xn = self._xnorm(x[i0:i0 + chunk])
            f = mean_value(self.mean, xn)
            corr = self._compute_correlation(xn)
            y_ = f.dot(self.inner.beta) + corr.dot(self.inner.gamma)
            out.append((y_ * self.y_std + self.y_mean)[:, 0])
        return np.concatenate(out)

    # algorithm.rs:518-550 (predict_gradients via predict_jacobian)
    def predict_gradients(self, x):
        x = np.asarray(x, dtype=np.float64)
        out = np.zeros((x.shape[0], self.xt_norm.shape[1]))
        for i in range(x.shape[0]):
            xn = (x[i] - self.x_mean) / self.x_std
            df_dx = mean_jacobian(self.mean, xn).T.dot(self.inner.beta)            # (nx, 1)
            dr = corr_jacobian(self.corr, xn, self.xt_norm, self.theta, self.w_star)
            dr_dx = df_dx + dr.T.dot(self.inner.gamma)
            out[i] = dr_dx[:, 0] * self.y_std / self.x_std
        return out

    # algorithm.rs:554-616 predict_var_gradients_single, :697-704 predict_var_gradients
    def predict_var_gradients(self, x):
        x = np.asarray(x, dtype=np.float64)
        out = np.zeros((x.shape[0], self.xt_norm.shape[1]))
        L = self.inner.r_chol
        f_mean = mean_value(self.mean, self.xt_norm)
        rho2 = sla.solve_triangular(L, f_mean, lower=True, check_finite=False)
        inv_kf = sla.solve_triangular(L.T, rho2, lower=False, check_finite=False)            # R^-1 F
        b_mat = f_mean.T.dot(inv_kf)
        rho3 = sla.cholesky(b_mat, lower=True, check_finite=False)
        for i in range(x.shape[0]):
            xn = ((x[i] - self.x_mean) / self.x_std)[None, :]
            r = corr_value(self.corr, xn - self.xt_norm, self.theta, self.w_star)[:, None]   # (n, 1)
            dr = corr_jacobian(self.corr, xn[0], self.xt_norm, self.theta, self.w_star)      # (n, nx)
            rho1 = sla.solve_triangular(L, r, lower=True, check_finite=False)
            inv_kr = sla.solve_triangular(L.T, rho1, lower=False, check_finite=False)        # R^-1 r
            p2 = inv_kr.T.dot(dr)                                                            # (1, nx)
            f_x = mean_value(self.mean, xn).T                                                # (p, 1)
            a_mat = f_x.T - r.T.dot(inv_kf)                                                  # (1, p)
            inv_bat = sla.solve_triangular(rho3, a_mat.T, lower=True, check_finite=False)
            d_mat = sla.solve_triangular(rho3.T, inv_bat, lower=False, check_finite=False)   # B^-1 A^T
            df = mean_jacobian(self.mean, xn[0])                                             # (p, nx)
            d_a = df.T - dr.T.dot(inv_kf)                                                    # (nx, p)
            p4 = d_mat.T.dot(d_a.T)                                                          # (1, nx)
            prime = 2.0 * (p4 - p2)
            out[i] = (prime / self.x_std * self.inner.sigma2)[0]
        return out

    # algorithm.rs:310-326 _compute_covariance: conditional covariance of the GP at the rows of x
    def compute_covariance(self, x):
        xn = self._xnorm(x)
        corr = self._compute_correlation(xn)
        rt, u = self._compute_rt_u(xn, corr)
        k = corr_value(self.corr, pairwise_differences(xn, xn), self.theta, self.w_star)
        k = k.reshape(xn.shape[0], xn.shape[0])
        return self.inner.sigma2 * (k - rt.T.dot(rt) + u.T.dot(u))

    # algorithm.rs:383-410 sample_chol / sample_eig / sample, :1153-1194 `sample`.  The reference draws
    # the standard-normal matrix itself (ndarray-rand, unseeded); here it is an argument (n_eval x n_traj)
    # so that the arithmetic can be compared.
    def sample(self, x, z, method="eig"):
        x = np.asarray(x, dtype=np.float64)
        mean = self.predict(x)[:, None]
        cov = self.compute_covariance(x)
        if method == "chol":
            c = sla.cholesky(cov, lower=True, check_finite=False)
        else:
            v, w = np.linalg.eigh(cov)
            v = np.where(v < 1e-9, 0.0, np.sqrt(np.where(v < 1e-9, 1.0, v)))
            c = w.dot(np.diag(v))
        return mean + c.dot(np.asarray(z, dtype=np.float64))

    # algorithm.rs:267-279
    def predict_var(self, x, chunk=1024):
        return self.predict_valvar(x, chunk)[1]

    # algorithm.rs:282-307
    def predict_valvar(self, x, chunk=1024):
        ys, vs = [], []
        x = np.asarray(x, dtype=np.float64)
        for i0 in range(0, x.shape[0], chunk):
            xn = self._xnorm(x[i0:i0 + chunk])
            f = mean_value(self.mean, xn)
            corr = self._compute_correlation(xn)
            y_ = f.dot(self.inner.beta) + corr.dot(self.inner.gamma)
            ys.append((y_ * self.y_std + self.y_mean)[:, 0])
            rt, u = self._compute_rt_u(xn, corr)
            mse = 1.0 - (rt * rt).sum(axis=0) + (u * u).sum(axis=0)
            mse = self.inner.sigma2 * mse
            vs.append(np.where(mse < 0.0, 0.0, mse))
        return np.concatenate(ys), np.concatenate(vs)


def _theta0(theta_init, dim):
    theta_init = np.atleast_1d(np.asarray(theta_init, dtype=np.float64))
    if theta_init.size == 1:
        return np.full(dim, theta_init[0])
    assert theta_init.size == dim          # algorithm.rs:835 panics otherwise
    return theta_init.copy()


def fit(x, y, corr=SQEXP, mean=CONSTANT, theta_init=THETA_DEFAULT_INIT,
        theta_bounds=THETA_DEFAULT_BOUNDS, fixed=False, n_start=GP_OPTIM_N_START,
        max_eval=GP_COBYLA_MAX_EVAL, nugget=DEFAULT_NUGGET, w_star=None, rng=None,
        trace=None):
    """impl Fit for GpValidParams::fit, algorithm.rs:791-979 (no KPLS: the PLS
    rotations come from linfa-pls; pass ``w_star`` explicitly to exercise the
    weighted kernels).  Optimiser: scipy's COBYLA with rhobeg 0.5
    (optimization.rs:16-24); the reference's cobyla 0.8 crate trajectory is
    not reproduced, only its objective, bounds, multistart and reduction."""
    from scipy.optimize import minimize

    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    nx = x.shape[1]
    w = np.eye(nx) if w_star is None else np.asarray(w_star, dtype=np.float64)
    dim = w.shape[1]
    theta0 = _theta0(theta_init, dim)
    xn, xm, xs = normalize(x)
    yn, ym, ys = normalize(y)
    fx = mean_value(mean, xn)
    y_std = float(ys[0])

    if fixed:
        theta = theta0
    else:
        b = np.atleast_2d(np.asarray(theta_bounds, dtype=np.float64))
        bounds = [tuple(b[0])] * dim if b.shape[0] == 1 else [tuple(r) for r in b]
        inits, lbounds = prepare_multistart(n_start, theta0, bounds, rng)
        maxeval = min(max(10 * dim, GP_COBYLA_MIN_EVAL), max_eval)       # :936-937
        lo = np.array([p[0] for p in lbounds])
        hi = np.array([p[1] for p in lbounds])

        def objfn(z):
            th = 10.0 ** np.clip(z, lo, hi)
            v = objective(corr, xn, fx, yn, y_std, th, w, nugget)
            if trace is not None:
                trace.append((th.copy(), v))
            return v if np.isfinite(v) else 1e300

        best = (math.inf, np.zeros(dim))
        for i in range(inits.shape[0]):
            res = minimize(objfn, inits[i], method="COBYLA", bounds=list(zip(lo, hi)),
                           options=dict(rhobeg=0.5, maxiter=maxeval, tol=1e-4))
            f = res.fun if np.isfinite(res.fun) and res.fun < 1e299 else math.inf
            if f < best[0]:
                best = (f, np.clip(res.x, lo, hi))
        theta = 10.0 ** best[1]

    rlf, inner = reduced_likelihood(corr, xn, fx, yn, y_std, theta, w, nugget)   # :966-968
    return GaussianProcess(corr=corr, mean=mean, theta=np.asarray(theta), likelihood=rlf,
                           inner=inner, w_star=w, xt_norm=xn, x_mean=xm, x_std=xs,
                           yt_norm=yn, y_mean=float(ym[0]), y_std=y_std,
                           training_data=(x, y[:, 0]))
