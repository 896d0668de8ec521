"""Cross-validation scores of a fitted surrogate (Marrel & Iooss 2024): Q2, predictive variance adequacy, integrated absolute
error on alpha -- gp/src/metrics.rs:35-58 (`PredictScore`) and moe/src/metrics.rs:31-215 (`CrossValScore`).

Every score is a fan-out of `kfold` refits with the model's own parameters followed by host arithmetic on (m,) vectors; the
refits are the GPU work.  `fit(x, y)` returns a model with `predict` / `predict_valvar` (and optionally `close`)."""
from __future__ import annotations

import math
from statistics import NormalDist

import numpy as np


def folds(n, kfold):
    """linfa `Dataset::fold(k)`: validation rows [i fs, (i + 1) fs) with fs = n / k; the remainder rows always train."""
    if kfold < 1 or kfold > n:
        raise ValueError("kfold should be in 1..%d, got %d" % (n, kfold))
    fs = n // kfold
    for i in range(kfold):
        yield (np.concatenate([np.arange(0, i * fs), np.arange((i + 1) * fs, n)]), np.arange(i * fs, (i + 1) * fs))


def _each_fold(training_data, kfold, fit):
    x, y = training_data
    for tr, va in folds(x.shape[0], kfold):
        model = fit(x[tr], y[tr])            # `.expect("cross-validation: sub model fitted")`: a failure propagates
        try:
            yield model, x[va], y[va]
        finally:
            if hasattr(model, "close"):
                model.close()


def q2_k_score(training_data, kfold, fit):
    """1 - PRESS / TSS (TSS around the mean of ALL targets), moe/src/metrics.rs:32-50."""
    y_mean = training_data[1].mean()
    press = tss = 0.0
    for model, xv, yv in _each_fold(training_data, kfold, fit):
        press += float(((yv - model.predict(xv)) ** 2).sum())
        tss += float(((yv - y_mean) ** 2).sum())
    return 1.0 - press / tss


def pva_k_score(training_data, kfold, fit):
    """|ln( mean (y - yhat)^2 / var )|, moe/src/metrics.rs:58-75."""
    varss, n = 0.0, 0
    for model, xv, yv in _each_fold(training_data, kfold, fit):
        pred, var = model.predict_valvar(xv)
        varss += float((((yv - pred) ** 2) / var).sum())
        n += yv.shape[0]
    return abs(math.log(varss / n))


def iae_alpha(model, xv, yv, alphas):
    """moe/src/metrics.rs:146-194: empirical coverage of the (1 - alpha) prediction intervals against 1 - alpha."""
    pred, var = model.predict_valvar(xv)
    sigma = np.sqrt(var)
    q = np.array([NormalDist().inv_cdf(1.0 - a / 2.0) for a in alphas])
    lo = pred[:, None] - sigma[:, None] * q[None, :]
    hi = pred[:, None] + sigma[:, None] * q[None, :]
    inside = (yv[:, None] >= lo) & (yv[:, None] <= hi)
    deltas = inside.sum(axis=0) / float(xv.shape[0])
    return float(np.abs(deltas - (1.0 - alphas)).sum() / alphas.size), deltas


def iae_alpha_k_score(training_data, kfold, fit, n_alpha=20):
    """moe/src/metrics.rs:83-138 -> (score, alphas, mean coverage per alpha)."""
    alphas = np.linspace(0.02, 0.98, n_alpha)
    scores, deltas = [], np.zeros(n_alpha)
    for model, xv, yv in _each_fold(training_data, kfold, fit):
        s, d = iae_alpha(model, xv, yv, alphas)
        scores.append(s)
        deltas += d
    return sum(scores) / len(scores), alphas, deltas / len(scores)
