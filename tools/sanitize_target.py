"""Small end-to-end run of every kernel of the path for compute-sanitizer (memcheck / racecheck):
dense GP n=300 (blocked path, 3 block columns -> look-ahead active) + K8 batch n=60 + sparse FITC N=500 M=140."""
import sys

import numpy as np

sys.path.insert(0, ".")
import egobox_b200 as eg   # noqa: E402

rng = np.random.default_rng(0)
x = rng.random((300, 3))
y = np.sum(np.sin(3 * x), axis=1)
gp = eg.GaussianProcess.params(eg.LinearMean, eg.Matern52Corr).n_start(1).max_eval(25).fit(x, y)
yv = gp.predict_valvar(rng.random((200, 3)))
print("dense ok", gp.likelihood(), float(yv[1].mean()))
xs = rng.random((60, 2))
gp2 = eg.Kriging.params().n_start(2).fit(xs, np.cos(4 * xs[:, 0]) + xs[:, 1])
print("small ok", gp2.likelihood())
X = 2 * rng.random((500, 2)) - 1
Y = np.sin(3 * X[:, 0]) + 0.1 * rng.normal(size=500)
ctx = eg.SgpContext(X, Y, X[:140].copy(), corr=eg.MATERN32, method=eg.SparseMethod.FITC, nugget=1e-8)
st, res = ctx.finalize([1.5, 1.0], 0.8, 0.02, want_inv=True)
v = ctx.predict_var(X[:100])
print("sparse ok", st, res["likelihood"], float(v.mean()))
