// PLS rotations for the KPLS option of the fit drivers (host, O(n d k) -- once per fit).
//
// Reference call site: `PlsRegression::params(n_components).fit(&ds)` + `.rotations().0`
// (gp/src/algorithm.rs:843-855, gp/src/sparse_algorithm.rs:442-455).  The arithmetic is the
// third-party crate linfa-pls 0.8.0 (a port of scikit-learn's PLSRegression): NIPALS with
// centred / scaled (ddof = 1) x and y, regression-mode deflation, rotations = W (P^T W)^+.
// A y residual that is numerically constant makes linfa-pls fail with
// PowerMethodConstantResidualError, which the reference turns into an all-zero w_star.
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "../../include/egobox_gpu.h"
#include "abi_guard.h"

namespace {

constexpr double kEps = 2.220446049250313e-16;

void center_scale(const double* a, int n, int c, std::vector<double>& out) {
    out.resize(static_cast<size_t>(n) * c);
    for (int j = 0; j < c; ++j) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += a[static_cast<size_t>(i) * c + j];
        const double mean = s / n;
        double v = 0.0;
        for (int i = 0; i < n; ++i) {
            const double t = a[static_cast<size_t>(i) * c + j] - mean;
            v += t * t;
        }
        double sd = n > 1 ? std::sqrt(v / (n - 1)) : 0.0;
        if (sd == 0.0 || std::isnan(sd)) sd = 1.0;
        for (int i = 0; i < n; ++i) out[static_cast<size_t>(i) * c + j] = (a[static_cast<size_t>(i) * c + j] - mean) / sd;
    }
}

// inverse of the k x k matrix a (row-major) by Gauss-Jordan with partial pivoting; false if singular
bool invert(std::vector<double>& a, int k, std::vector<double>& inv) {
    inv.assign(static_cast<size_t>(k) * k, 0.0);
    for (int i = 0; i < k; ++i) inv[static_cast<size_t>(i) * k + i] = 1.0;
    for (int c = 0; c < k; ++c) {
        int piv = c;
        for (int r = c + 1; r < k; ++r)
            if (std::fabs(a[static_cast<size_t>(r) * k + c]) > std::fabs(a[static_cast<size_t>(piv) * k + c])) piv = r;
        if (!(std::fabs(a[static_cast<size_t>(piv) * k + c]) > 1e-300)) return false;
        if (piv != c)
            for (int j = 0; j < k; ++j) {
                std::swap(a[static_cast<size_t>(piv) * k + j], a[static_cast<size_t>(c) * k + j]);
                std::swap(inv[static_cast<size_t>(piv) * k + j], inv[static_cast<size_t>(c) * k + j]);
            }
        const double ip = 1.0 / a[static_cast<size_t>(c) * k + c];
        for (int j = 0; j < k; ++j) {
            a[static_cast<size_t>(c) * k + j] *= ip;
            inv[static_cast<size_t>(c) * k + j] *= ip;
        }
        for (int r = 0; r < k; ++r) {
            if (r == c) continue;
            const double f = a[static_cast<size_t>(r) * k + c];
            if (f == 0.0) continue;
            for (int j = 0; j < k; ++j) {
                a[static_cast<size_t>(r) * k + j] -= f * a[static_cast<size_t>(c) * k + j];
                inv[static_cast<size_t>(r) * k + j] -= f * inv[static_cast<size_t>(c) * k + j];
            }
        }
    }
    return true;
}

}  // namespace

// x: n x d raw inputs, y: n raw outputs (single target, as in gp `fit`), k components.
// w_star: d x k, row-major.  Zeros (and EGX_OK) on a constant residual, like the reference.
extern "C" int egx_pls_rotations(const double* x, int n, int d, const double* y, int k, double* w_star) try {
    if (!x || !y || !w_star || n < 2 || d < 1 || k < 1 || k > d) {
        egx_set_error("egx_pls_rotations: invalid argument (n=%d d=%d k=%d)", n, d, k);
        return EGX_INVALID_VALUE;
    }
    std::vector<double> xk, yk;
    center_scale(x, n, d, xk);
    center_scale(y, n, 1, yk);
    std::vector<double> W(static_cast<size_t>(d) * k, 0.0), P(static_cast<size_t>(d) * k, 0.0);
    std::vector<double> w(d), t(n), p(d);
    for (int c = 0; c < k; ++c) {
        // y residual numerically zero -> constant-residual error of the power method
        double ymax = 0.0;
        for (int i = 0; i < n; ++i) ymax = std::fmax(ymax, std::fabs(yk[i]));
        if (ymax < 10.0 * kEps) {
            std::memset(w_star, 0, sizeof(double) * d * k);
            return EGX_OK;
        }
        // one power iteration is exact for a single target: w = X^T y / (y^T y), normalised
        double yy = 0.0;
        for (int i = 0; i < n; ++i) yy += yk[i] * yk[i];
        for (int j = 0; j < d; ++j) w[j] = 0.0;
        for (int i = 0; i < n; ++i) {
            const double yi = yk[i];
            const double* xr = &xk[static_cast<size_t>(i) * d];
            for (int j = 0; j < d; ++j) w[j] += xr[j] * yi;
        }
        double nw = 0.0;
        for (int j = 0; j < d; ++j) {
            w[j] /= yy;
            nw += w[j] * w[j];
        }
        nw = std::sqrt(nw) + kEps;
        int jmax = 0;
        for (int j = 0; j < d; ++j) {
            w[j] /= nw;
            if (std::fabs(w[j]) > std::fabs(w[jmax])) jmax = j;
        }
        if (w[jmax] < 0.0)                       // svd_flip_1d: largest |w| component positive
            for (int j = 0; j < d; ++j) w[j] = -w[j];
        // scores, loadings, deflation
        double tt = 0.0;
        for (int i = 0; i < n; ++i) {
            const double* xr = &xk[static_cast<size_t>(i) * d];
            double s = 0.0;
            for (int j = 0; j < d; ++j) s += xr[j] * w[j];
            t[i] = s;
            tt += s * s;
        }
        for (int j = 0; j < d; ++j) p[j] = 0.0;
        double ty = 0.0;
        for (int i = 0; i < n; ++i) {
            const double* xr = &xk[static_cast<size_t>(i) * d];
            for (int j = 0; j < d; ++j) p[j] += t[i] * xr[j];
            ty += t[i] * yk[i];
        }
        for (int j = 0; j < d; ++j) p[j] /= tt;
        const double q = ty / tt;
        for (int i = 0; i < n; ++i) {
            double* xr = &xk[static_cast<size_t>(i) * d];
            for (int j = 0; j < d; ++j) xr[j] -= t[i] * p[j];
            yk[i] -= t[i] * q;
        }
        for (int j = 0; j < d; ++j) {
            W[static_cast<size_t>(j) * k + c] = w[j];
            P[static_cast<size_t>(j) * k + c] = p[j];
        }
    }
    // rotations = W (P^T W)^-1
    std::vector<double> ptw(static_cast<size_t>(k) * k, 0.0), inv;
    for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b) {
            double s = 0.0;
            for (int j = 0; j < d; ++j) s += P[static_cast<size_t>(j) * k + a] * W[static_cast<size_t>(j) * k + b];
            ptw[static_cast<size_t>(a) * k + b] = s;
        }
    if (!invert(ptw, k, inv)) {
        egx_set_error("egx_pls_rotations: singular loadings/weights product");
        return EGX_INVALID_VALUE;
    }
    for (int j = 0; j < d; ++j)
        for (int b = 0; b < k; ++b) {
            double s = 0.0;
            for (int a = 0; a < k; ++a) s += W[static_cast<size_t>(j) * k + a] * inv[static_cast<size_t>(a) * k + b];
            w_star[static_cast<size_t>(j) * k + b] = s;
        }
    return EGX_OK;
}
EGX_ABI_CATCH
