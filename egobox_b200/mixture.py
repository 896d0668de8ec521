"""Mixture-of-experts fan-out over the kriging path.

The reference (`egobox-moe`) clusters the training set with a Gaussian mixture, fits one GP expert per cluster
(sequential loop, moe/src/algorithm.rs:165-177) and recombines the experts' predictions with the mixture's
responsibilities (`predict_smooth` :411-423, `predict_var_smooth` :670-685, `predict_hard` :879-888,
`predict_var_hard` :894-910).  The clustering itself is the reference's control plane and stays there: this
class takes the cluster labels of the training rows and a responsibility function `probas(x) -> (m, k)` (the
reference's `gmx.predict_probas`) and owns what runs on the GPUs -- the k independent fits (sharded over
ranks through `parallel.fit_experts` when torch.distributed is initialised) and batched prediction with
hard / smooth recombination.  Hard mode groups the points by cluster and calls each expert once, where the
reference predicts one point at a time."""
from __future__ import annotations

import numpy as np

from . import parallel
from .gp import GaussianProcess, GpParams


from .gpx import Recombination          # HARD = 0, SMOOTH = 1 (python/src/types.rs Recombination)


def recombine_smooth(preds, variances, probas):
    """y = sum_k p_k y_k (algorithm.rs:417-421) ; var = sum_k p_k^2 var_k (:675-683).
    preds / variances: (k, m) expert outputs (either may be None) ; probas: (m, k)."""
    p = np.asarray(probas, dtype=np.float64)
    y = None if preds is None else np.einsum("km,mk->m", np.asarray(preds, dtype=np.float64), p)
    v = None if variances is None else np.einsum("km,mk->m", np.asarray(variances, dtype=np.float64), p * p)
    return y, v


def hard_clusters(probas):
    """`gmx.predict`: index of the largest responsibility per point (algorithm.rs:880)."""
    return np.argmax(np.asarray(probas, dtype=np.float64), axis=1)


class ExpertMixture:
    """k GP experts + a responsibility function; `fit` is the expert loop of moe/src/algorithm.rs:165-177."""

    def __init__(self, experts, probas, recombination=Recombination.HARD, owner=None):
        self.experts = experts                    # dict or list: expert index -> GaussianProcess (local ones)
        self.probas = probas
        self.recombination = recombination
        self.n_clusters = len(owner) if owner is not None else len(experts)
        self.owner = owner                        # expert index -> rank holding its device state (None: all local)
        self.table = None

    @classmethod
    def fit(cls, xt, yt, labels, probas, params: GpParams | None = None, recombination=Recombination.HARD,
            min_points=None, expert_fit=None):
        """labels[i] = cluster of training row i (0..k-1) ; one expert per cluster, on the rank
        `cluster % world_size` when a process group is up (weak scaling, no data-path collective).
        expert_fit(x_c, y_c) -> model replaces `params.fit` (the gloo CPU tests inject a stand-in)."""
        xt = np.ascontiguousarray(xt, dtype=np.float64)
        yt = np.ascontiguousarray(yt, dtype=np.float64).reshape(-1)
        labels = np.asarray(labels).reshape(-1)
        if labels.shape[0] != xt.shape[0]:
            raise ValueError("one cluster label per training row expected")
        k = int(labels.max()) + 1
        params = params if params is not None else GaussianProcess.params()
        need = xt.shape[1] + 1 if min_points is None else min_points
        for c in range(k):
            if int(np.sum(labels == c)) < max(need, 2):
                raise ValueError("cluster %d has too few points for a GP expert" % c)
        theta_dim = params._kpls_dim or xt.shape[1]
        fit_fn = expert_fit if expert_fit is not None else params.fit

        def fit_one(c):
            rows = labels == c
            gp = fit_fn(xt[rows], yt[rows])
            return gp, gp.likelihood(), gp.variance(), gp.theta()

        models, table = parallel.fit_experts(fit_one, k, theta_dim)
        dist = parallel._dist()
        world = dist.get_world_size() if dist else 1
        mix = cls(models, probas, recombination, owner=[c % world for c in range(k)])
        mix.table = table                          # (k, 2 + theta_dim): likelihood, variance, theta of every expert
        return mix

    # -- prediction ---------------------------------------------------------------------------
    def _local(self, c):
        return self.experts[c] if isinstance(self.experts, dict) else self.experts[c]

    def _have(self, c):
        return (c in self.experts) if isinstance(self.experts, dict) else c < len(self.experts)

    def predict_valvar(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        p = np.asarray(self.probas(x), dtype=np.float64)
        m, k = x.shape[0], self.n_clusters
        if p.shape != (m, k):
            raise ValueError("probas(x) must be (m, %d)" % k)
        y = np.zeros(m)
        v = np.zeros(m)
        if self.recombination == Recombination.SMOOTH:
            for c in range(k):
                if not self._have(c):
                    continue
                yc, vc = self._local(c).predict_valvar(x)
                y += p[:, c] * yc
                v += p[:, c] * p[:, c] * vc
        else:
            cl = hard_clusters(p)
            for c in range(k):
                idx = np.nonzero(cl == c)[0]
                if idx.size == 0 or not self._have(c):
                    continue
                yc, vc = self._local(c).predict_valvar(x[idx])
                y[idx] = yc
                v[idx] = vc
        return self._reduce(y), self._reduce(v)

    def predict(self, x):
        return self.predict_valvar(x)[0]

    def predict_var(self, x):
        return self.predict_valvar(x)[1]

    def _reduce(self, a):
        """Experts living on other ranks contribute through one sum all-reduce of the (m,) partial result."""
        dist = parallel._dist()
        if dist is None or dist.get_world_size() == 1:
            return a
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a)).to(parallel._device_for_backend())
        dist.all_reduce(t)
        return t.cpu().numpy()

    def close(self):
        for gp in (self.experts.values() if isinstance(self.experts, dict) else self.experts):
            gp.close()
