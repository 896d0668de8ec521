"""-m gpu: the serial chain of the factorisation (r02 forms of K3 / K5, the diagonal-tile-first look-ahead schedule, the
one-launch back substitution) against the CPU oracle, at sizes that reach every branch:
  T = 2 (no look-ahead), T = 6 / 11 (look-ahead, odd and even block-column counts, 32-row slabs), n = 8192 elsewhere
  (tests/test_gpu_fullsize.py: 64-row slabs in the first half of the columns).
Reference being replaced: `r_mx.cholesky()` + solve_triangular + gamma, gp/src/algorithm.rs:1004-1034."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import gp_oracle as O
from tests.gpu_util import make_problem, make_context, oracle_gp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,d,mean", [(250, 3, O.CONSTANT), (700, 5, O.LINEAR), (1300, 8, O.CONSTANT), (1408, 4, O.QUADRATIC)])
def test_factor_and_gamma_against_oracle(n, d, mean):
    x, y = make_problem(n, d, seed=n)
    theta = np.full(d, 1.5)
    ctx, _ = make_context(x, y, O.MATERN52, mean)
    gp = oracle_gp(x, y, O.MATERN52, mean, theta)
    st, res = ctx.finalize(theta)
    assert st == 0
    assert res["rlf"] == pytest.approx(gp.likelihood, rel=1e-9)            # north-star bound: 1e-6
    L = ctx.download_chol()
    np.testing.assert_allclose(L, gp.inner.r_chol, rtol=0, atol=1e-10)
    gs = np.abs(gp.inner.gamma).max()
    np.testing.assert_allclose(res["gamma"], gp.inner.gamma[:, 0], rtol=0, atol=1e-7 * gs)
    # a second evaluation on the same workspace (CUDA-graph replay for npad <= 4096) gives the same bits
    st2, res2 = ctx.finalize(theta)
    assert st2 == 0 and res2["rlf"] == res["rlf"]
    np.testing.assert_array_equal(res2["gamma"], res["gamma"])
    ctx.close()


@pytest.mark.parametrize("n", [300, 900])
def test_not_positive_definite_is_an_error_not_a_number(n):
    # diagonal 1 + nugget = 0.5 under off-diagonal correlations close to 1: the second pivot is negative whatever the
    # rounding (dpotrf semantics: algorithm.rs:1004 -> Err -> the objective returns +inf, :893-896)
    x, y = make_problem(n, 2, seed=11)
    ctx, _ = make_context(x, y, O.SQEXP, O.CONSTANT, nugget=-0.5)
    st, rlf = ctx.reduced_likelihood(np.full(2, 1e-3))
    assert st != 0 and not np.isfinite(rlf)
    ctx.close()


def test_correlation_matrix_at_the_dimension_of_the_reference_bench():
    # crates/gp/benches/corr.rs runs dim = 100: the coordinate tiles still fit in shared memory
    n, d = 300, 100
    x, y = make_problem(n, d, seed=5)
    ctx, (xn, *_rest) = make_context(x, y, O.MATERN52, O.CONSTANT)
    theta = np.linspace(0.05, 0.2, d)
    R = ctx.correlation_matrix(theta)
    Ro = O.corr_matrix(O.MATERN52, xn, theta, np.eye(d))
    np.testing.assert_allclose(R, Ro, rtol=1e-12, atol=1e-300)
    ctx.close()


def test_diagonal_block_probe():
    """tools/micro/potrf_probe.cu: K3 against a long-double host Cholesky, its inverted 32 x 32 blocks, the LAPACK failure
    index, and that nothing outside the lower triangle of the tile is touched."""
    exe = os.path.join(ROOT, "tools", "micro", "potrf_probe")
    if not os.path.exists(exe):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            pytest.skip("no nvcc to build the probe")
        subprocess.run([nvcc, "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                        os.path.join(ROOT, "tools", "micro", "potrf_probe.cu")], check=True, timeout=600)
    for version in ("2", "1"):
        out = subprocess.run([exe], env=dict(os.environ, PROBE_V=version), capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and "potrf probe: ok" in out.stdout, out.stdout + out.stderr
