"""Writes tests/golden/pls_rotations.json: x_rotations_ of scikit-learn's PLSRegression (the
algorithm linfa-pls 0.8.0 ports; reference call site crates/gp/src/algorithm.rs:843-855) on
small seeded datasets.  Run once in the build container (needs scikit-learn); the JSON is what
the tests read.

    python tests/golden/make_pls_golden.py
"""
import json
import os

import numpy as np
from sklearn.cross_decomposition import PLSRegression
import sklearn


def griewank(x):
    d = x.shape[1]
    return (x ** 2).sum(axis=1) / 4000.0 - np.prod(np.cos(x / np.sqrt(np.arange(1, d + 1))), axis=1) + 1.0


def rosenbrock(x):
    return (100.0 * (x[:, 1:] - x[:, :-1] ** 2) ** 2 + (1.0 - x[:, :-1]) ** 2).sum(axis=1)


CASES = [
    # name, n, d, k, limits, function   (shapes of algorithm.rs:1326-1440's kpls tests)
    ("griewank_100x5_k3", 100, 5, 3, (-600.0, 600.0), griewank),
    ("tp_exp_300x3_k1", 300, 3, 1, (-1.0, 1.0), lambda x: np.exp(x).prod(axis=1)),
    ("rosenb_30x20_k1", 30, 20, 1, (-1.0, 1.0), rosenbrock),
    ("rosenb_60x10_k4", 60, 10, 4, (-2.0, 2.0), rosenbrock),
    ("full_rank_40x4_k4", 40, 4, 4, (0.0, 1.0), lambda x: np.sin(3 * x).sum(axis=1) + x[:, 0] * x[:, 1]),
]


def main():
    out = {"sklearn": sklearn.__version__, "cases": []}
    for i, (name, n, d, k, lim, f) in enumerate(CASES):
        rng = np.random.default_rng(100 + i)
        x = lim[0] + (lim[1] - lim[0]) * rng.random((n, d))
        y = f(x)
        pls = PLSRegression(n_components=k, scale=True, max_iter=500, tol=1e-6).fit(x, y)
        out["cases"].append({"name": name, "seed": 100 + i, "n": n, "d": d, "k": k,
                             "x": x.tolist(), "y": y.tolist(),
                             "x_rotations": pls.x_rotations_.tolist()})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pls_rotations.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path)


if __name__ == "__main__":
    main()
