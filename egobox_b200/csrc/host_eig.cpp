// Host eigen-decomposition of a small symmetric matrix, used only by the eigenvalue variant of the
// conditional sampler (gp/src/algorithm.rs:1169-1187: `cov_x.eigh()` on the m x m covariance of the
// sampled locations, CPU in the reference as well).  Householder tridiagonalisation followed by the
// implicit-shift QL iteration (the classical EISPACK tred2 / tql2 pair).
#include <cmath>
#include <vector>

#include "common.cuh"
#include "../../include/egobox_gpu.h"
#include "abi_guard.h"

// a: n x n symmetric, row-major.  On return the COLUMNS of a are the orthonormal eigenvectors and
// w[j] the matching eigenvalues (unsorted).  Returns 0, or 1 if the QL iteration did not converge.
int egx_host_symmetric_eig(int n, double* a, double* w) {
    if (n <= 0) return 0;
    std::vector<double> e(n, 0.0);
    double* d = w;
    auto V = [&](int i, int j) -> double& { return a[static_cast<size_t>(i) * n + j]; };

    // ---- reduction to tridiagonal form ----------------------------------------------------------
    for (int j = 0; j < n; ++j) d[j] = V(n - 1, j);
    for (int i = n - 1; i > 0; --i) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; ++j) {
                d[j] = V(i - 1, j);
                V(i, j) = 0.0;
                V(j, i) = 0.0;
            }
        } else {
            for (int k = 0; k < i; ++k) {
                d[k] /= scale;
                h += d[k] * d[k];
            }
            double f = d[i - 1];
            double g = std::sqrt(h);
            if (f > 0.0) g = -g;
            e[i] = scale * g;
            h -= f * g;
            d[i - 1] = f - g;
            for (int j = 0; j < i; ++j) e[j] = 0.0;
            for (int j = 0; j < i; ++j) {
                f = d[j];
                V(j, i) = f;
                g = e[j] + V(j, j) * f;
                for (int k = j + 1; k <= i - 1; ++k) {
                    g += V(k, j) * d[k];
                    e[k] += V(k, j) * f;
                }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; ++j) {
                e[j] /= h;
                f += e[j] * d[j];
            }
            const double hh = f / (h + h);
            for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
            for (int j = 0; j < i; ++j) {
                f = d[j];
                g = e[j];
                for (int k = j; k <= i - 1; ++k) V(k, j) -= (f * e[k] + g * d[k]);
                d[j] = V(i - 1, j);
                V(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    // accumulate the Householder reflectors
    for (int i = 0; i < n - 1; ++i) {
        V(n - 1, i) = V(i, i);
        V(i, i) = 1.0;
        const double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; ++k) d[k] = V(k, i + 1) / h;
            for (int j = 0; j <= i; ++j) {
                double g = 0.0;
                for (int k = 0; k <= i; ++k) g += V(k, i + 1) * V(k, j);
                for (int k = 0; k <= i; ++k) V(k, j) -= g * d[k];
            }
        }
        for (int k = 0; k <= i; ++k) V(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; ++j) {
        d[j] = V(n - 1, j);
        V(n - 1, j) = 0.0;
    }
    V(n - 1, n - 1) = 1.0;
    e[0] = 0.0;

    // ---- implicit QL on the tridiagonal matrix ---------------------------------------------------
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = std::ldexp(1.0, -52);
    int status = 0;
    for (int l = 0; l < n; ++l) {
        tst1 = std::fmax(tst1, std::fabs(d[l]) + std::fabs(e[l]));
        int m = l;
        while (m < n - 1 && std::fabs(e[m]) > eps * tst1) ++m;
        if (m > l) {
            int iter = 0;
            do {
                if (++iter > 60) {
                    status = 1;
                    break;
                }
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = std::hypot(p, 1.0);
                if (p < 0.0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                const double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; ++i) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = 1.0, c3 = 1.0;
                const double el1 = e[l + 1];
                double s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; --i) {
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = std::hypot(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; ++k) {
                        h = V(k, i + 1);
                        V(k, i + 1) = s * V(k, i) + c * h;
                        V(k, i) = c * V(k, i) - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
            } while (std::fabs(e[l]) > eps * tst1);
        }
        d[l] += f;
        e[l] = 0.0;
    }
    return status;
}

extern "C" int egx_symmetric_eig(int n, double* a, double* w) try {
    if (n < 0 || (n > 0 && (!a || !w))) return EGX_INVALID_VALUE;
    return egx_host_symmetric_eig(n, a, w) == 0 ? EGX_OK : EGX_INVALID_VALUE;
}
EGX_ABI_CATCH
