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
# tcgen05 trailing update (13 block columns -> sliced launches in the look-ahead and in the batched form)
from tools._util import make_problem, make_context   # noqa: E402
xo, yo = make_problem(1664, 3, seed=1)
cto = make_context(xo, yo, eg.MATERN52, eg.CONSTANT)
cto.set_profiling(True)
st, rl = cto.reduced_likelihood(np.full(3, 1.0))
stb, rlb = cto.reduced_likelihood_batch(np.full((3, 3), 1.0) * np.array([[0.9], [1.0], [1.1]]))
print("tcgen05 ok", st, rl, stb.tolist(), cto.profile()["ozaki_syrk"][1])
cto.close()
# mixture of experts on the device (responsibilities, hard / smooth recombination, gradients) + KPLS fit
xm = rng.random((160, 2))
ym = np.where(xm[:, 0] < 0.5, np.sin(4 * xm).sum(axis=1), 3.0 + (xm ** 2).sum(axis=1))
for rec in (eg.Recombination.HARD, eg.Recombination.SMOOTH):
    gpx = eg.Gpx.builder(n_clusters=2, recombination=rec, n_start=1, seed=1).fit(xm, ym)
    xq = rng.random((50, 2))
    print("moe ok", rec, float(gpx.predict(xq).mean()), float(gpx.predict_var(xq).mean()),
          float(np.abs(gpx.predict_gradients(xq)).mean()), float(np.abs(gpx.predict_var_gradients(xq)).mean()))
gk = eg.GaussianProcess.params(eg.ConstantMean, eg.SquaredExponentialCorr).kpls_dim(2).n_start(1).fit(rng.random((80, 5)), rng.random(80))
print("kpls ok", gk.theta().tolist())
