"""Shared helpers for the -m gpu parity tests (CUDA path vs oracle on the same inputs)."""
import numpy as np

from oracle import gp_oracle as O


def lhs(n, d, seed):
    rng = np.random.default_rng(seed)
    return O.lhs_classic(np.array([[0.0, 1.0]] * d), n, rng)


def rosenbrock(x):
    z = 4.0 * x - 2.0
    return np.sum(100.0 * (z[:, 1:] - z[:, :-1] ** 2) ** 2 + (1.0 - z[:, :-1]) ** 2, axis=1)


def sphere_sin(x):
    return np.sum(np.sin(3.0 * x) + x * x, axis=1)


def make_problem(n, d, seed=42, fn=None):
    x = lhs(n, d, seed)
    if fn is None:
        fn = rosenbrock if d > 1 else (lambda a: (a[:, 0] * 25 - 3.5) * np.sin((a[:, 0] * 25 - 3.5) / np.pi))
    y = fn(x)
    return x, y


def make_context(x, y, corr, mean, w_star=None, nugget=O.DEFAULT_NUGGET):
    import egobox_b200 as eg
    xn, xm, xs = O.normalize(x)
    yn, ym, ys = O.normalize(y.reshape(-1, 1))
    ctx = eg.GpContext(xn, yn[:, 0], xm, xs, float(ym[0]), float(ys[0]), corr, mean, w_star=w_star, nugget=nugget)
    return ctx, (xn, xm, xs, yn, float(ym[0]), float(ys[0]))


def oracle_gp(x, y, corr, mean, theta, w_star=None, nugget=O.DEFAULT_NUGGET):
    return O.fit(x, y, corr=corr, mean=mean, theta_init=theta, fixed=True, w_star=w_star, nugget=nugget)
