"""Input generation for the dev probes (no oracle import: tools/ only drive the product path)."""
import numpy as np

import egobox_b200 as eg


def lhs(n, d, seed):
    rng = np.random.default_rng(seed)
    u = rng.random((n, d))
    pts = (np.arange(n)[:, None] + u) / n
    for j in range(d):
        pts[:, j] = pts[rng.permutation(n), j]
    return pts


def rosenbrock(x):
    z = 4.0 * x - 2.0
    return np.sum(100.0 * (z[:, 1:] - z[:, :-1] ** 2) ** 2 + (1.0 - z[:, :-1]) ** 2, axis=1)


def make_problem(n, d, seed=42):
    x = lhs(n, d, seed)
    y = rosenbrock(x) if d > 1 else (x[:, 0] * 25 - 3.5) * np.sin((x[:, 0] * 25 - 3.5) / np.pi)
    return x, y


def normalize(a):
    mean = a.mean(axis=0)
    std = a.std(axis=0, ddof=1)
    std = np.where(std == 0.0, 1.0, std)
    return (a - mean) / std, mean, std


def make_context(x, y, corr, mean):
    xn, xm, xs = normalize(x)
    yn, ym, ys = normalize(y.reshape(-1, 1))
    return eg.GpContext(xn, yn[:, 0], xm, xs, float(ym[0]), float(ys[0]), corr, mean)
