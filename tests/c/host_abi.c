/* Host-only entry points of the C ABI called from plain C99 (no GPU, no Python): the optimisers behind the fit drivers,
 * the multistart seeds, the symmetric eigen-solver of the eigenvalue sampler and the PLS rotations.  Built and run by
 * tests/test_cabi_exports.py; exit code = number of failed checks. */
#include <math.h>
#include <stdio.h>

#include "egobox_gpu.h"

static int failures = 0;
#define CHECK(cond, what)                                   \
    do {                                                    \
        if (!(cond)) {                                      \
            fprintf(stderr, "FAILED: %s (%s)\n", what, #cond); \
            ++failures;                                     \
        }                                                   \
    } while (0)

static double quad(const double* x, int n, void* user) {
    const double* c = (const double*)user;
    double s = 0.0;
    int i;
    for (i = 0; i < n; ++i) s += (x[i] - c[i]) * (x[i] - c[i]);
    return s;
}
static double quad_grad(const double* x, int n, double* g, void* user) {
    const double* c = (const double*)user;
    int i;
    for (i = 0; i < n; ++i) g[i] = 2.0 * (x[i] - c[i]);
    return quad(x, n, user);
}

int main(void) {
    /* minimum of |x - c|^2 over [0, 1]^3 with c = (0.3, 1.7, -0.4): (0.3, 1, 0) */
    double c[3] = {0.3, 1.7, -0.4}, x0[3] = {0.5, 0.5, 0.5}, lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
    double x[3], f = 0.0;
    int nev = 0, st;

    st = egx_bound_cobyla_minimize(quad, c, 3, x0, lo, hi, 0.5, 1e-10, 300, x, &f, &nev);
    CHECK(st == EGX_OK && nev > 3 && nev <= 300, "cobyla status / budget");
    CHECK(fabs(x[0] - 0.3) < 1e-3 && fabs(x[1] - 1.0) < 1e-9 && fabs(x[2]) < 1e-9, "cobyla optimum");

    st = egx_bound_lbfgs_minimize(quad_grad, c, 3, x0, lo, hi, 1e-15, 1e-10, 100, x, &f, &nev);
    CHECK(st == EGX_OK && nev <= 100, "lbfgs status / budget");
    CHECK(fabs(x[0] - 0.3) < 1e-8 && x[1] == 1.0 && x[2] == 0.0, "lbfgs optimum");
    CHECK(fabs(f - (0.49 + 0.16)) < 1e-12, "lbfgs value");
    CHECK(egx_bound_lbfgs_minimize(0, c, 3, x0, lo, hi, 1e-9, 1e-7, 10, x, &f, &nev) == EGX_INVALID_VALUE, "lbfgs null fn");

    {   /* prepare_multistart: row 0 = log10 theta0, rows 1.. inside the log10 box (optimization.rs:26-71) */
        double theta0[2] = {0.1, 0.1}, bounds[4] = {1e-2, 1e1, 1e-2, 1e1}, starts[11 * 2];
        int i, inside = 1;
        st = egx_prepare_multistart(10, theta0, bounds, 2, 42ULL, starts);
        CHECK(st == EGX_OK && fabs(starts[0] + 1.0) < 1e-15 && fabs(starts[1] + 1.0) < 1e-15, "multistart row 0");
        for (i = 2; i < 22; ++i) inside = inside && starts[i] >= -2.0 && starts[i] <= 1.0;
        CHECK(inside, "multistart rows inside the box");
    }
    {   /* eigenvalues of [[2, 1], [1, 2]] are 1 and 3 */
        double a[4] = {2, 1, 1, 2}, w[2];
        st = egx_symmetric_eig(2, a, w);
        CHECK(st == EGX_OK, "eig status");
        CHECK(fabs((w[0] < w[1] ? w[0] : w[1]) - 1.0) < 1e-14 && fabs((w[0] < w[1] ? w[1] : w[0]) - 3.0) < 1e-14, "eig values");
    }
    {   /* PLS with one component on y = 2 x0 - x1: the rotation is proportional to X^T y of the centred, scaled data */
        double xs[8] = {0, 0, 1, 0, 0, 1, 1, 1}, ys[4] = {0, 2, -1, 1}, w[2];
        st = egx_pls_rotations(xs, 4, 2, ys, 1, w);
        CHECK(st == EGX_OK, "pls status");
        CHECK(fabs(w[0] / w[1] + 2.0) < 1e-12 && fabs(w[0] * w[0] + w[1] * w[1] - 1.0) < 1e-12, "pls rotation");
    }
    {   /* the exchange of the sharded path, world size 1 (tests/test_host_comm.py runs three processes) */
        egx_comm* comm = 0;
        double v = 2.5, pay[2] = {1.0, 2.0}, all[2];
        int winner = -1;
        CHECK(egx_comm_init(&comm, 1, 0, 0, 0, 0) == EGX_OK && egx_comm_size(comm) == 1 && egx_comm_rank(comm) == 0, "comm init");
        CHECK(egx_comm_allgather(comm, pay, 2, all) == EGX_OK && all[0] == 1.0 && all[1] == 2.0, "comm allgather");
        CHECK(egx_argmin_allreduce(comm, &v, pay, 2, &winner) == EGX_OK && winner == 0 && v == 2.5, "comm argmin");
        egx_comm_destroy(comm);
        CHECK(egx_comm_init(&comm, 2, 5, "127.0.0.1", 29999, 100) == EGX_INVALID_VALUE && comm == 0, "comm bad rank");
    }
    CHECK(egx_device_count() >= 0 && egx_version() != 0, "version / device count");
    if (failures == 0) printf("host ABI ok\n");
    return failures;
}
