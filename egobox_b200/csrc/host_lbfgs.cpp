// Bound-constrained limited-memory BFGS for the gradient-based multistart (the reference
// optimises with derivative-free COBYLA, gp/src/optimization.rs:122-169, because it has no theta gradient).
// Host only.  Projected-gradient form: variables sitting on a bound with the gradient pushing outward are frozen, the
// two-loop recursion (8 pairs) runs on the free ones, the step is projected back into the box and accepted by an Armijo
// backtracking search; a non-descent direction or a failed search drops the curvature pairs and restarts from steepest
// descent.  An objective value of +inf / NaN (the likelihood's Err(_) -> +inf mapping, algorithm.rs:893-896) counts as
// a rejected trial point.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <vector>

#include "../../include/egobox_gpu.h"
#include "abi_guard.h"

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();

double dot(const std::vector<double>& a, const std::vector<double>& b) {
    double s = 0.0;
    for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
    return s;
}

}  // namespace

extern "C" int egx_bound_lbfgs_minimize(egx_objective_grad_fn fg, void* user, int n, const double* x0, const double* lo,
                                        const double* hi, double ftol_rel, double gtol, int maxeval, double* x_opt,
                                        double* f_opt, int* n_evals) try {
    if (!fg || !x0 || !lo || !hi || n < 1 || !x_opt || !f_opt) return EGX_INVALID_VALUE;
    for (int i = 0; i < n; ++i)
        if (!(lo[i] <= hi[i])) return EGX_INVALID_VALUE;
    const int memory = 8;
    std::vector<double> x(n), g(n), xn(n), gn(n), d(n), pg(n);
    for (int i = 0; i < n; ++i) x[i] = std::min(std::max(x0[i], lo[i]), hi[i]);
    int nfev = 0;
    auto eval = [&](const std::vector<double>& z, std::vector<double>& grad) {
        ++nfev;
        double v = fg(z.data(), n, grad.data(), user);
        if (std::isnan(v)) v = kInf;
        if (std::isfinite(v))
            for (int i = 0; i < n; ++i)
                if (!std::isfinite(grad[i])) v = kInf;
        return v;
    };
    double f = eval(x, g);
    // a start that fails (non positive definite R, ill-conditioned regression: +inf, algorithm.rs:893-896) is not given up at
    // once: the COBYLA chains treat +inf as an ordinary bad vertex and keep exploring, so a few points on the way to the centre
    // of the box are tried before this start is abandoned
    for (int probe = 1; probe <= 4 && !std::isfinite(f) && nfev < maxeval; ++probe) {
        const double w = 0.25 * probe;
        for (int i = 0; i < n; ++i) x[i] = (1.0 - w) * x0[i] + w * 0.5 * (lo[i] + hi[i]);
        for (int i = 0; i < n; ++i) x[i] = std::min(std::max(x[i], lo[i]), hi[i]);
        f = eval(x, g);
    }
    std::memcpy(x_opt, x.data(), sizeof(double) * n);
    *f_opt = f;
    std::deque<std::vector<double>> S, Y;
    std::deque<double> RHO;
    bool first = true;
    while (std::isfinite(f) && nfev < maxeval) {
        // projected gradient
        double pgmax = 0.0;
        for (int i = 0; i < n; ++i) {
            const bool frozen = (x[i] <= lo[i] && g[i] > 0.0) || (x[i] >= hi[i] && g[i] < 0.0);
            pg[i] = frozen ? 0.0 : g[i];
            pgmax = std::max(pgmax, std::fabs(pg[i]));
        }
        if (pgmax <= gtol * std::max(1.0, std::fabs(f))) break;
        // two-loop recursion on the free variables
        for (int i = 0; i < n; ++i) d[i] = pg[i];
        std::vector<double> alpha(S.size());
        for (int k = static_cast<int>(S.size()) - 1; k >= 0; --k) {
            alpha[k] = RHO[k] * dot(S[k], d);
            for (int i = 0; i < n; ++i) d[i] -= alpha[k] * Y[k][i];
        }
        if (!S.empty()) {
            const double gamma = dot(S.back(), Y.back()) / dot(Y.back(), Y.back());
            for (int i = 0; i < n; ++i) d[i] *= gamma;
        }
        for (size_t k = 0; k < S.size(); ++k) {
            const double beta = RHO[k] * dot(Y[k], d);
            for (int i = 0; i < n; ++i) d[i] += (alpha[k] - beta) * S[k][i];
        }
        for (int i = 0; i < n; ++i) d[i] = (pg[i] == 0.0) ? 0.0 : -d[i];
        if (!(dot(d, g) < 0.0)) {                       // not a descent direction: restart from steepest descent
            S.clear();
            Y.clear();
            RHO.clear();
            for (int i = 0; i < n; ++i) d[i] = -pg[i];
        }
        // projected Armijo backtracking
        double t = 1.0;
        if (first || S.empty()) {
            double dn = 0.0;
            for (int i = 0; i < n; ++i) dn = std::max(dn, std::fabs(d[i]));
            double box = 0.0;
            for (int i = 0; i < n; ++i) box = std::max(box, hi[i] - lo[i]);
            t = std::min(1.0, 0.25 * (box > 0.0 ? box : 1.0) / dn);     // first trial: a quarter of the box at most
        }
        first = false;
        double fn = kInf;
        bool accepted = false;
        for (int ls = 0; ls < 25 && nfev < maxeval; ++ls, t *= 0.5) {
            double decrease = 0.0, moved = 0.0;
            for (int i = 0; i < n; ++i) {
                xn[i] = std::min(std::max(x[i] + t * d[i], lo[i]), hi[i]);
                decrease += g[i] * (xn[i] - x[i]);
                moved = std::max(moved, std::fabs(xn[i] - x[i]));
            }
            if (moved == 0.0) break;
            fn = eval(xn, gn);
            if (std::isfinite(fn) && fn <= f + 1e-4 * decrease) {
                accepted = true;
                break;
            }
        }
        if (!accepted) {
            if (S.empty()) break;                        // steepest descent failed too: converged to working precision
            S.clear();
            Y.clear();
            RHO.clear();
            continue;
        }
        std::vector<double> s(n), y(n);
        for (int i = 0; i < n; ++i) {
            s[i] = xn[i] - x[i];
            y[i] = gn[i] - g[i];
        }
        const double sy = dot(s, y);
        if (sy > 1e-10 * std::sqrt(dot(s, s) * dot(y, y))) {
            S.push_back(s);
            Y.push_back(y);
            RHO.push_back(1.0 / sy);
            if (static_cast<int>(S.size()) > memory) {
                S.pop_front();
                Y.pop_front();
                RHO.pop_front();
            }
        }
        const double fprev = f;
        x = xn;
        g = gn;
        f = fn;
        std::memcpy(x_opt, x.data(), sizeof(double) * n);
        *f_opt = f;
        if (fprev - f <= ftol_rel * std::max(std::max(std::fabs(fprev), std::fabs(f)), 1.0)) break;
    }
    if (n_evals) *n_evals = nfev;
    return EGX_OK;
}
EGX_ABI_CATCH
