/* CPU restatement (plain C, OpenMP over rows) of the reference correlation kernels --
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY, never linked into the product library.
 *
 * Follows crates/gp/src/correlation_models.rs value():
 *   SquaredExponential :91-104   r = exp(-1/2 sum_j d_j^2 * sum_l (theta_l W_jl)^2)
 *   AbsoluteExponential:185-196  r = exp(-sum_j |d_j| * sum_l |W_jl| theta_l)
 *   Matern32 :326-353            prod_j prod_l (1 + sqrt3 tw_jl |d_j|) * exp(-sqrt3 sum_jl tw_jl |d_j|)
 *   Matern52 :497-523            prod_j prod_l (1 + sqrt5 tw |d| + 5/3 tw^2 d^2) * exp(-sqrt5 sum ...)
 * with tw_jl = theta_l |W_jl|, evaluated pair by pair in the reference's loop order
 * (j outer, l inner) but without materialising the difference table
 * (utils.rs:80-104 / :110-131).  It is validated against oracle/gp_oracle.py (numpy) in
 * tests/test_oracle_golden.py and is the kernel of bench.py's cpu_baseline.
 */
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline sets its thread count explicitly */
void egx_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int egx_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static double pair_value(int kind, const double* a, const double* b, int d, int h, const double* theta,
                         const double* w) {
    const double sqrt3 = sqrt(3.0), sqrt5 = sqrt(5.0), c53 = 5.0 / 3.0;
    if (kind == 0) {
        double r = 0.0;
        for (int j = 0; j < d; ++j) {
            double tw = 0.0;
            for (int l = 0; l < h; ++l) { const double v = theta[l] * w[j * h + l]; tw += v * v; }
            const double dj = a[j] - b[j];
            r += dj * dj * tw;
        }
        return exp(-0.5 * r);
    }
    if (kind == 1) {
        double r = 0.0;
        for (int j = 0; j < d; ++j) {
            double tw = 0.0;
            for (int l = 0; l < h; ++l) tw += fabs(w[j * h + l]) * theta[l];
            r += fabs(a[j] - b[j]) * tw;
        }
        return exp(-r);
    }
    double prod = 1.0, s = 0.0;
    for (int j = 0; j < d; ++j) {
        const double dj = a[j] - b[j], ad = fabs(dj);
        for (int l = 0; l < h; ++l) {
            const double v = theta[l] * fabs(w[j * h + l]);
            if (kind == 2) prod *= 1.0 + sqrt3 * v * ad;
            else prod *= 1.0 + sqrt5 * v * ad + c53 * (v * v * dj * dj);
            s += ad * v;
        }
    }
    return prod * exp(-(kind == 2 ? sqrt3 : sqrt5) * s);
}

/* full symmetric R (n x n), (1 + nugget) on the diagonal: algorithm.rs:997-1001 */
void egx_oracle_corr_matrix(int kind, const double* x, int n, int d, const double* theta, const double* w, int h,
                            double nugget, double* R) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < i; ++j) {
            const double r = pair_value(kind, x + (size_t)i * d, x + (size_t)j * d, d, h, theta, w);
            R[(size_t)i * n + j] = r;
            R[(size_t)j * n + i] = r;
        }
        R[(size_t)i * n + i] = 1.0 + nugget;
    }
}

/* c(x*, X): m x n, algorithm.rs:372-380 */
void egx_oracle_cross_corr(int kind, const double* xs, int m, const double* x, int n, int d, const double* theta,
                           const double* w, int h, double* C) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j)
            C[(size_t)i * n + j] = pair_value(kind, xs + (size_t)i * d, x + (size_t)j * d, d, h, theta, w);
}
