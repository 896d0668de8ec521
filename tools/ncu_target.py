"""Small driver for ncu captures: one likelihood evaluation + finalize at n=8192 (Matern-5/2, d=10)
and predict_valvar on 2048 points -- every kernel of the hot path appears at its bench shape."""
import sys

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg                                # noqa: E402
from tools._util import make_problem, make_context      # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
x, y = make_problem(n, 10, seed=42)
ctx = make_context(x, y, eg.MATERN52, eg.CONSTANT)
theta = np.full(10, 1.0)
st, rlf = ctx.reduced_likelihood(theta)
st, res = ctx.finalize(theta, want_ft=False)
xs = np.random.default_rng(43).random((m, 10))
yv = ctx.predict_valvar(xs)
print("rlf", rlf, "var[0]", yv[1][0])
ctx.close()
