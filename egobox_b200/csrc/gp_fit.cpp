// Host-side GP model: the fit driver of gp/src/algorithm.rs:791-979 above the device seam.
//
//   normalise (utils.rs:45-54) -> device context -> theta0 / bounds / multistart seeds
//   (algorithm.rs:815-838, 899-925; optimization.rs:26-71) -> n_start+1 derivative-free
//   chains advanced in LOCK STEP, one egx_gp_reduced_likelihood_batch call per optimiser
//   iteration (replaces the rayon `into_par_iter` at algorithm.rs:928-945) -> min reduction
//   -> final evaluation (algorithm.rs:966-968).
//
// The optimiser is a linear-interpolation trust-region method in the COBYLA family
// (Powell 1994: simplex of n+1 points, linear model, trust radius rho halved when the
// model stops paying, geometry-improvement steps with alpha = 1/4, beta = 2.1, gamma = 1/2),
// specialised to the only constraints this path has -- simple bounds in log10(theta) -- so
// the trust-region subproblem is solved exactly (clipped steepest descent) instead of
// through COBYLA's general LP.  It keeps the reference's settings (rhobeg 0.5, ftol_rel
// 1e-4, maxeval = clamp(10 dim, 25, max_eval)); the trajectory of the third-party
// `cobyla 0.8.0` crate is not reproduced (it is not in /root/reference, and the reference
// pins only the optimum, test_gpmix.py:37-53).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "../../include/egobox_gpu.h"
#include "abi_guard.h"

void egx_set_error(const char* fmt, ...);

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();
constexpr int GP_COBYLA_MIN_EVAL = 25;   // gp/src/algorithm.rs:35

// ---- xoshiro256+ (the generator family the reference seeds its LHS with) -------------
struct Xoshiro256Plus {
    uint64_t s[4];
    explicit Xoshiro256Plus(uint64_t seed) {
        uint64_t z = seed;
        for (auto& v : s) {   // splitmix64 seeding
            z += 0x9e3779b97f4a7c15ULL;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
            x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
            v = x ^ (x >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = s[0] + s[3];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    // The draws below restate rand 0.8.5 / rand_xoshiro 0.6.0 (the versions Cargo.lock pins for crates/gp, crates/doe), so that
    // a seed gives the SAME LHS multistart points and the SAME inducing points as the reference (pinned on the fixture of
    // crates/doe/src/lhs.rs:332-347 in tests/test_host_rng.py):
    //   RngCore::next_u32 = upper half of next_u64; Uniform::new(0., 1.) = 52 mantissa bits of next_u64 (UniformFloat::sample);
    //   SliceRandom::shuffle = for i in (1..len).rev(): swap(i, gen_index(i + 1)); gen_index = UniformInt<u32>::sample_single
    //   (widening multiply, rejection zone (range << lz) - 1).
    uint32_t next_u32() { return static_cast<uint32_t>(next() >> 32); }
    double uniform() { return static_cast<double>(next() >> 12) * (1.0 / 4503599627370496.0); }
    uint64_t below(uint64_t n) {
        if (n > 0xffffffffULL) {                       // usize path of gen_index (never reached by the sizes of this path)
            const uint64_t zone = (n << __builtin_clzll(n)) - 1;
            for (;;) {
                const unsigned __int128 m = static_cast<unsigned __int128>(next()) * n;
                if (static_cast<uint64_t>(m) <= zone) return static_cast<uint64_t>(m >> 64);
            }
        }
        const uint32_t range = static_cast<uint32_t>(n);
        const uint32_t zone = (range << __builtin_clz(range)) - 1;
        for (;;) {
            const uint64_t m = static_cast<uint64_t>(next_u32()) * range;
            if (static_cast<uint32_t>(m) <= zone) return m >> 32;
        }
    }
    template <class T>
    void shuffle(std::vector<T>& v) {
        for (size_t i = v.size(); i-- > 1;) std::swap(v[i], v[below(i + 1)]);
    }
};

// crates/doe/src/lhs.rs:235-258 (classic: ALL uniforms first, column by column, then one shuffle per column) and :283-304
// (maximin = best of 5 classic designs by smallest pair distance)
std::vector<double> lhs_classic(int ns, int nx, Xoshiro256Plus& rng) {
    std::vector<double> pts(static_cast<size_t>(ns) * nx);
    const double step = 1.0 / ns;                      // Array::linspace(0., 1., ns + 1): cut[i] = 0 + step * i
    std::vector<std::vector<double>> cols(nx, std::vector<double>(ns));
    for (int j = 0; j < nx; ++j)
        for (int i = 0; i < ns; ++i) cols[j][i] = rng.uniform();
    for (int j = 0; j < nx; ++j) {
        for (int i = 0; i < ns; ++i) {
            const double a = step * i, b = step * (i + 1);
            cols[j][i] = cols[j][i] * (b - a) + a;
        }
        rng.shuffle(cols[j]);
        for (int i = 0; i < ns; ++i) pts[static_cast<size_t>(i) * nx + j] = cols[j][i];
    }
    return pts;
}
double min_pdist(const std::vector<double>& p, int ns, int nx) {
    double best = kInf;
    for (int a = 0; a < ns; ++a)
        for (int b = a + 1; b < ns; ++b) {
            double s = 0.0;
            for (int j = 0; j < nx; ++j) {
                const double t = p[static_cast<size_t>(a) * nx + j] - p[static_cast<size_t>(b) * nx + j];
                s += t * t;
            }
            best = std::min(best, std::sqrt(s));
        }
    return best;
}
std::vector<double> lhs_maximin(int ns, int nx, Xoshiro256Plus& rng) {
    std::vector<double> best = lhs_classic(ns, nx, rng);
    double dbest = min_pdist(best, ns, nx);
    for (int it = 0; it < 4; ++it) {
        std::vector<double> cand = lhs_classic(ns, nx, rng);
        const double dm = min_pdist(cand, ns, nx);
        if (dbest < dm) {
            dbest = dm;
            best.swap(cand);
        }
    }
    return best;
}

// ---- bound-constrained linear-model trust-region optimiser (ask / tell) ---------------
class BoundCobyla {
   public:
    BoundCobyla(const std::vector<double>& x0, const std::vector<double>& lo, const std::vector<double>& hi,
                double rhobeg, double ftol_rel, int maxfun)
        : n_(static_cast<int>(x0.size())), lo_(lo), hi_(hi), rho_(rhobeg), ftol_rel_(ftol_rel), maxfun_(maxfun) {
        V_.assign(n_ + 1, std::vector<double>(n_));
        F_.assign(n_ + 1, kInf);
        for (int i = 0; i < n_; ++i) V_[0][i] = std::min(std::max(x0[i], lo_[i]), hi_[i]);
        pending_ = V_[0];
        phase_ = INIT;
        init_idx_ = 0;
        if (maxfun_ < 1) phase_ = DONE;
    }
    bool done() const { return phase_ == DONE; }
    const std::vector<double>& ask() const { return pending_; }
    double best_f() const { return fbest_; }
    const std::vector<double>& best_x() const { return xbest_; }
    int nfev() const { return nfev_; }

    void tell(double f_raw) {
        // failed evaluations come back as +inf (algorithm.rs:893-896); keep the linear algebra finite
        const double f = std::isnan(f_raw) ? kBig : std::min(f_raw, kBig);
        ++nfev_;
        bool ftol_hit = false;
        if (f_raw < fbest_) {   // NaN compares false
            if (ftol_rel_ > 0.0 && std::isfinite(fbest_) &&
                std::fabs(f_raw - fbest_) < ftol_rel_ * (std::fabs(f_raw) + std::fabs(fbest_)) * 0.5)
                ftol_hit = true;
            fbest_ = f_raw;
            xbest_ = pending_;
        } else if (xbest_.empty()) {
            xbest_ = pending_;
        }
        if (phase_ == INIT) {
            V_[init_idx_] = pending_;
            F_[init_idx_] = f;
            ++init_idx_;
            if (nfev_ >= maxfun_) {
                phase_ = DONE;
                return;
            }
            if (init_idx_ <= n_) {
                // Powell's initial simplex is greedy: the next coordinate step starts from the best
                // point found so far (the better vertex becomes the base), not from x0
                const int j = init_idx_ - 1;
                int jb = 0;
                for (int k = 1; k < init_idx_; ++k)
                    if (F_[k] < F_[jb]) jb = k;
                pending_ = V_[jb];
                double step = rho_;
                if (pending_[j] + step > hi_[j]) step = -step;
                pending_[j] = std::min(std::max(pending_[j] + step, lo_[j]), hi_[j]);
                return;
            }
            iterate(false);
            return;
        }
        if (ftol_hit || nfev_ >= maxfun_) {
            phase_ = DONE;
            return;
        }
        if (phase_ == TRUST) {
            const double actual = F_[0] - f;
            insert_after_trust(f);
            // Powell: a step that earns >= 10 % of the predicted reduction is followed by another
            // trust-region step straight away (no geometry check in between)
            if (replaced_ && actual > 0.0 && actual >= 0.1 * predicted_) iterate(true);
            else after_failure();
            return;
        }
        if (phase_ == GEOM) {
            V_[jdrop_] = pending_;
            F_[jdrop_] = f;
            iterate(true);
            return;
        }
    }

   private:
    enum Phase { INIT, TRUST, GEOM, DONE };
    static constexpr double kBig = 1e100;
    static constexpr double kAlpha = 0.25, kBeta = 2.1, kGamma = 0.5;

    void best_to_front() {
        int jb = 0;
        for (int j = 1; j <= n_; ++j)
            if (F_[j] < F_[jb]) jb = j;
        if (jb != 0) {
            std::swap(V_[0], V_[jb]);
            std::swap(F_[0], F_[jb]);
        }
    }

    // simi = inverse of the edge matrix S (columns s_j = V_j - V_0); false if singular
    bool compute_simi() {
        const int n = n_;
        std::vector<double> a(static_cast<size_t>(n) * 2 * n, 0.0);
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) a[static_cast<size_t>(i) * 2 * n + j] = V_[j + 1][i] - V_[0][i];
            a[static_cast<size_t>(i) * 2 * n + n + i] = 1.0;
        }
        for (int c = 0; c < n; ++c) {
            int piv = c;
            for (int r = c + 1; r < n; ++r)
                if (std::fabs(a[static_cast<size_t>(r) * 2 * n + c]) > std::fabs(a[static_cast<size_t>(piv) * 2 * n + c])) piv = r;
            if (std::fabs(a[static_cast<size_t>(piv) * 2 * n + c]) < 1e-300) return false;
            if (piv != c)
                for (int k = 0; k < 2 * n; ++k) std::swap(a[static_cast<size_t>(piv) * 2 * n + k], a[static_cast<size_t>(c) * 2 * n + k]);
            const double inv = 1.0 / a[static_cast<size_t>(c) * 2 * n + c];
            for (int k = 0; k < 2 * n; ++k) a[static_cast<size_t>(c) * 2 * n + k] *= inv;
            for (int r = 0; r < n; ++r) {
                if (r == c) continue;
                const double f = a[static_cast<size_t>(r) * 2 * n + c];
                if (f == 0.0) continue;
                for (int k = 0; k < 2 * n; ++k) a[static_cast<size_t>(r) * 2 * n + k] -= f * a[static_cast<size_t>(c) * 2 * n + k];
            }
        }
        simi_.assign(static_cast<size_t>(n) * n, 0.0);
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) simi_[static_cast<size_t>(j) * n + i] = a[static_cast<size_t>(j) * 2 * n + n + i];
        sigma_.assign(n, 0.0);
        eta_.assign(n, 0.0);
        acceptable_ = true;
        for (int j = 0; j < n; ++j) {
            double rs = 0.0, es = 0.0;
            for (int i = 0; i < n; ++i) {
                rs += simi_[static_cast<size_t>(j) * n + i] * simi_[static_cast<size_t>(j) * n + i];
                const double e = V_[j + 1][i] - V_[0][i];
                es += e * e;
            }
            sigma_[j] = 1.0 / std::sqrt(rs);
            eta_[j] = std::sqrt(es);
            if (sigma_[j] < kAlpha * rho_ || eta_[j] > kBeta * rho_) acceptable_ = false;
        }
        return true;
    }

    void gradient() {
        g_.assign(n_, 0.0);
        for (int j = 0; j < n_; ++j) {
            const double df = F_[j + 1] - F_[0];
            for (int i = 0; i < n_; ++i) g_[i] += simi_[static_cast<size_t>(j) * n_ + i] * df;
        }
    }

    // argmin g.d  s.t. |d| <= rho, lo <= V0 + d <= hi  : d = clip(-t g) with |d| = rho
    double trust_step(std::vector<double>& d) const {
        const int n = n_;
        std::vector<double> a(n), b(n);
        for (int i = 0; i < n; ++i) {
            a[i] = lo_[i] - V_[0][i];
            b[i] = hi_[i] - V_[0][i];
        }
        auto eval = [&](double t, std::vector<double>& out) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) {
                double v = -t * g_[i];
                v = std::min(std::max(v, a[i]), b[i]);
                out[i] = v;
                s += v * v;
            }
            return std::sqrt(s);
        };
        d.assign(n, 0.0);
        double gmax = 0.0;
        for (int i = 0; i < n; ++i) gmax = std::max(gmax, std::fabs(g_[i]));
        if (gmax == 0.0 || !std::isfinite(gmax)) return 0.0;
        // corner reached when t -> inf
        std::vector<double> dinf(n);
        for (int i = 0; i < n; ++i) dinf[i] = g_[i] < 0.0 ? b[i] : (g_[i] > 0.0 ? a[i] : 0.0);
        double ninf = 0.0;
        for (double v : dinf) ninf += v * v;
        ninf = std::sqrt(ninf);
        if (ninf <= rho_) {
            d = dinf;
            return ninf;
        }
        double tlo = 0.0, thi = rho_ / gmax;
        while (eval(thi, d) < rho_) thi *= 2.0;
        for (int it = 0; it < 200; ++it) {
            const double tm = 0.5 * (tlo + thi);
            if (eval(tm, d) < rho_) tlo = tm;
            else thi = tm;
            if (thi - tlo <= 1e-15 * thi) break;
        }
        return eval(thi, d);
    }

    void reinit_simplex() {
        // degenerate simplex: rebuild around the best vertex with the current radius
        best_to_front();
        phase_ = INIT;
        init_idx_ = 1;
        pending_ = V_[0];
        double step = rho_;
        if (pending_[0] + step > hi_[0]) step = -step;
        pending_[0] = std::min(std::max(pending_[0] + step, lo_[0]), hi_[0]);
    }

    void iterate(bool force_trust) {
        for (;;) {
            best_to_front();
            if (!compute_simi()) {
                reinit_simplex();
                return;
            }
            if (force_trust || acceptable_) {
                gradient();
                std::vector<double> d;
                const double dn = trust_step(d);
                if (dn >= 0.5 * rho_) {
                    predicted_ = 0.0;
                    for (int i = 0; i < n_; ++i) predicted_ -= g_[i] * d[i];
                    d_ = d;
                    pending_ = V_[0];
                    for (int i = 0; i < n_; ++i) pending_[i] = std::min(std::max(pending_[i] + d[i], lo_[i]), hi_[i]);
                    phase_ = TRUST;
                    return;
                }
                // step too short to be worth an evaluation
                if (!acceptable_) {
                    geometry_step();
                    return;
                }
                if (!reduce_rho()) return;
                force_trust = false;
                continue;
            }
            geometry_step();
            return;
        }
    }

    void after_failure() {
        best_to_front();
        if (!compute_simi()) {
            reinit_simplex();
            return;
        }
        if (!acceptable_) {
            geometry_step();
            return;
        }
        if (!reduce_rho()) return;
        iterate(false);
    }

    bool reduce_rho() {
        if (rho_ <= rhoend_) {
            phase_ = DONE;
            return false;
        }
        rho_ *= 0.5;
        if (rho_ <= 1.5 * rhoend_) rho_ = rhoend_;
        return true;
    }

    // V_/F_ are ordered (best first) and simi_ is current
    void geometry_step() {
        int jd = -1;
        double worst = kBeta * rho_;
        for (int j = 0; j < n_; ++j)
            if (eta_[j] > worst) {
                worst = eta_[j];
                jd = j;
            }
        if (jd < 0) {
            double smin = kInf;
            for (int j = 0; j < n_; ++j)
                if (sigma_[j] < smin) {
                    smin = sigma_[j];
                    jd = j;
                }
        }
        gradient();
        std::vector<double> u(n_);
        for (int i = 0; i < n_; ++i) u[i] = simi_[static_cast<size_t>(jd) * n_ + i] * sigma_[jd];   // unit normal
        double gu = 0.0;
        for (int i = 0; i < n_; ++i) gu += g_[i] * u[i];
        double sgn = (gu > 0.0) ? -1.0 : 1.0;
        auto make = [&](double s, std::vector<double>& out) {
            double len = 0.0;
            out = V_[0];
            for (int i = 0; i < n_; ++i) {
                const double t = std::min(std::max(out[i] + s * kGamma * rho_ * u[i], lo_[i]), hi_[i]);
                len += (t - out[i]) * (t - out[i]);
                out[i] = t;
            }
            return std::sqrt(len);
        };
        std::vector<double> c1, c2;
        const double l1 = make(sgn, c1);
        if (l1 < 0.5 * kGamma * rho_) {
            const double l2 = make(-sgn, c2);
            if (l2 > l1) c1 = c2;
        }
        pending_ = c1;
        jdrop_ = jd + 1;
        phase_ = GEOM;
    }

    // Decide whether the trust-region point replaces a vertex (Powell 1994, COBYLA "L440"):
    //   temp_j = |simi_j . d| is the ratio new/old simplex volume if vertex j is replaced; the change is
    //   mandatory when the point improves on the best vertex, otherwise the volume must grow (temp_j > 1).
    //   Among replacements that keep the simplex acceptable, prefer dropping the vertex farthest from the
    //   new point when it lies beyond 1.1 rho.
    void insert_after_trust(double f) {
        const int n = n_;
        const bool improved = f < F_[0];
        double ratio = improved ? 0.0 : 1.0;
        int jd = -1;
        std::vector<double> sigbar(n);
        for (int j = 0; j < n; ++j) {
            double t = 0.0;
            for (int i = 0; i < n; ++i) t += simi_[static_cast<size_t>(j) * n + i] * d_[i];
            t = std::fabs(t);
            if (t > ratio) {
                jd = j;
                ratio = t;
            }
            sigbar[j] = t * sigma_[j];
        }
        double edgmax = 1.1 * rho_;
        int l = -1;
        const double parsig = kAlpha * rho_;
        for (int j = 0; j < n; ++j) {
            if (sigbar[j] >= parsig || sigbar[j] >= sigma_[j]) {
                double t = eta_[j];
                if (predicted_ > 0.0) {
                    t = 0.0;
                    for (int i = 0; i < n; ++i) {
                        const double e = d_[i] - (V_[j + 1][i] - V_[0][i]);
                        t += e * e;
                    }
                    t = std::sqrt(t);
                }
                if (t > edgmax) {
                    l = j;
                    edgmax = t;
                }
            }
        }
        if (l >= 0) jd = l;
        replaced_ = jd >= 0;
        if (!replaced_) return;
        V_[jd + 1] = pending_;
        F_[jd + 1] = f;
    }

    int n_;
    std::vector<double> lo_, hi_;
    double rho_, rhoend_ = 1e-8, ftol_rel_;
    int maxfun_, nfev_ = 0;
    std::vector<std::vector<double>> V_;
    std::vector<double> F_, simi_, sigma_, eta_, g_, d_, pending_, xbest_;
    bool acceptable_ = false, replaced_ = false;
    double predicted_ = 0.0, fbest_ = kInf;
    int init_idx_ = 0, jdrop_ = 0;
    Phase phase_;
};

void normalize_cols(const double* x, int n, int d, std::vector<double>& xn, std::vector<double>& mean,
                    std::vector<double>& sd) {
    // gp/src/utils.rs:45-54: mean, std with ddof = 1, zero std -> 1
    xn.assign(static_cast<size_t>(n) * d, 0.0);
    mean.assign(d, 0.0);
    sd.assign(d, 1.0);
    for (int j = 0; j < d; ++j) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += x[static_cast<size_t>(i) * d + j];
        const double m = s / n;
        double v = 0.0;
        for (int i = 0; i < n; ++i) {
            const double t = x[static_cast<size_t>(i) * d + j] - m;
            v += t * t;
        }
        double st = (n > 1) ? std::sqrt(v / (n - 1)) : 0.0;
        if (st == 0.0 || std::isnan(st)) st = 1.0;
        mean[j] = m;
        sd[j] = st;
        for (int i = 0; i < n; ++i) xn[static_cast<size_t>(i) * d + j] = (x[static_cast<size_t>(i) * d + j] - m) / st;
    }
}

}  // namespace

struct egx_gp_model {
    egx_gp_ctx* ctx = nullptr;
    int n = 0, d = 0, h = 0, p = 0;
    std::vector<double> theta, x_mean, x_std, w_star;
    double y_mean = 0.0, y_std = 1.0, likelihood = NAN, sigma2 = NAN;
    long long n_evals = 0;
};

// Stand-alone access to the optimiser (host only, no GPU): used by the CPU tests and by
// callers that want the reference's `optimize_params` (optimization.rs:122-169) semantics.
extern "C" int egx_bound_cobyla_minimize(egx_objective_fn f, void* user, int n, const double* x0, const double* lo,
                                         const double* hi, double rhobeg, double ftol_rel, int maxeval,
                                         double* x_opt, double* f_opt, int* n_evals) try {
    if (!f || !x0 || !lo || !hi || n < 1 || !x_opt || !f_opt) return EGX_INVALID_VALUE;
    BoundCobyla opt(std::vector<double>(x0, x0 + n), std::vector<double>(lo, lo + n), std::vector<double>(hi, hi + n),
                    rhobeg, ftol_rel, maxeval);
    while (!opt.done()) {
        const std::vector<double>& z = opt.ask();
        double v = f(z.data(), n, user);
        if (std::isnan(v)) v = kInf;
        opt.tell(v);
    }
    *f_opt = opt.best_f();
    if (opt.best_x().empty()) std::memcpy(x_opt, x0, sizeof(double) * n);
    else std::memcpy(x_opt, opt.best_x().data(), sizeof(double) * n);
    if (n_evals) *n_evals = opt.nfev();
    return EGX_OK;
}
EGX_ABI_CATCH

// prepare_multistart seeds (optimization.rs:26-71): (n_start + 1) x dim log10 starts
extern "C" int egx_prepare_multistart(int n_start, const double* theta0, const double* bounds, int dim,
                                      unsigned long long seed, double* starts_out) try {
    if (!theta0 || !bounds || dim < 1 || n_start < 0 || !starts_out) return EGX_INVALID_VALUE;
    for (int i = 0; i < dim; ++i) starts_out[i] = std::log10(theta0[i]);
    Xoshiro256Plus rng(seed);
    if (n_start == 1) {
        for (int i = 0; i < dim; ++i) {
            const double a = std::log10(bounds[2 * i]), b = std::log10(bounds[2 * i + 1]);
            starts_out[dim + i] = a + (b - a) * rng.uniform();
        }
    } else if (n_start > 1) {
        std::vector<double> pts = lhs_maximin(n_start, dim, rng);
        for (int s = 0; s < n_start; ++s)
            for (int i = 0; i < dim; ++i) {
                const double a = std::log10(bounds[2 * i]), b = std::log10(bounds[2 * i + 1]);
                starts_out[static_cast<size_t>(s + 1) * dim + i] = a + (b - a) * pts[static_cast<size_t>(s) * dim + i];
            }
    }
    return EGX_OK;
}
EGX_ABI_CATCH

// doe/src/lhs.rs:67-88 (SamplingMethod::sample = normalized_sample * (upper - lower) + lower)
extern "C" int egx_lhs_sample(int kind, int ns, int nx, const double* xlimits, unsigned long long seed, double* out) try {
    if (!xlimits || !out || ns < 1 || nx < 1 || kind < 0 || kind > 1) return EGX_INVALID_VALUE;
    Xoshiro256Plus rng(seed);
    const std::vector<double> pts = kind == 0 ? lhs_classic(ns, nx, rng) : lhs_maximin(ns, nx, rng);
    for (int i = 0; i < ns; ++i)
        for (int j = 0; j < nx; ++j) {
            const double lo = xlimits[2 * j], hi = xlimits[2 * j + 1];
            out[static_cast<size_t>(i) * nx + j] = pts[static_cast<size_t>(i) * nx + j] * (hi - lo) + lo;
        }
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_shuffled_indices(int n, unsigned long long seed, int* out) try {
    if (!out || n < 0) return EGX_INVALID_VALUE;
    Xoshiro256Plus rng(seed);
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    rng.shuffle(idx);
    std::copy(idx.begin(), idx.end(), out);
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" void egx_gp_params_default(egx_gp_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->corr = EGX_CORR_SQUARED_EXPONENTIAL;
    p->mean = EGX_MEAN_CONSTANT;
    p->theta_tuning = EGX_THETA_FULL;
    p->n_start = 10;                       // GP_OPTIM_N_START
    p->max_eval = 1000;                    // GP_COBYLA_MAX_EVAL
    p->nugget = 100.0 * 2.220446049250313e-16;
    p->seed = 42;
    p->cobyla_rhobeg = 0.5;
    p->cobyla_ftol_rel = 1e-4;
}

extern "C" void egx_gp_model_destroy(egx_gp_model* m) {
    if (!m) return;
    if (m->ctx) egx_gp_destroy(m->ctx);
    delete m;
}

extern "C" int egx_gp_fit(const egx_gp_params* prm, const double* x, int n, int d, const double* y,
                          egx_gp_model** out) try {
    if (!out) return EGX_INVALID_VALUE;
    *out = nullptr;
    if (!prm || !x || !y || n < 1 || d < 1) {
        egx_set_error("egx_gp_fit: invalid argument");
        return EGX_INVALID_VALUE;
    }
    const bool kpls = prm->w_star != nullptr || prm->kpls_dim > 0;
    if (kpls && (prm->kpls_dim < 1 || prm->kpls_dim > d)) {
        // algorithm.rs:798-807
        egx_set_error("Dimension reduction %d should be smaller than actual training input dimensions %d",
                      prm->kpls_dim, d);
        return EGX_INVALID_VALUE;
    }
    const int h = kpls ? prm->kpls_dim : d;

    // theta0 (algorithm.rs:828-838)
    static const double kDefaultInit = 0.1;
    std::vector<double> theta0(h);
    if (prm->theta_init == nullptr || prm->n_theta_init == 0) {
        std::fill(theta0.begin(), theta0.end(), kDefaultInit);
    } else if (prm->n_theta_init == 1) {
        std::fill(theta0.begin(), theta0.end(), prm->theta_init[0]);
    } else if (prm->n_theta_init == h) {
        theta0.assign(prm->theta_init, prm->theta_init + h);
    } else {
        egx_set_error("Initial guess for theta should be either 1-dim or dim of xtrain (w_star.ncols()), got %d",
                      prm->n_theta_init);
        return EGX_INVALID_VALUE;
    }

    std::unique_ptr<egx_gp_model, void (*)(egx_gp_model*)> m(new egx_gp_model(), egx_gp_model_destroy);
    m->n = n;
    m->d = d;
    m->h = h;
    std::vector<double> xn, yn, ym, ys;
    normalize_cols(x, n, d, xn, m->x_mean, m->x_std);
    normalize_cols(y, n, 1, yn, ym, ys);
    m->y_mean = ym[0];
    m->y_std = ys[0];
    if (kpls && prm->w_star) m->w_star.assign(prm->w_star, prm->w_star + static_cast<size_t>(d) * h);
    else if (kpls) {
        // PLS rotations of the RAW data (algorithm.rs:843-855)
        m->w_star.assign(static_cast<size_t>(d) * h, 0.0);
        const int stp = egx_pls_rotations(x, n, d, y, h, m->w_star.data());
        if (stp != EGX_OK) return stp;
    } else {
        m->w_star.assign(static_cast<size_t>(d) * d, 0.0);
        for (int j = 0; j < d; ++j) m->w_star[static_cast<size_t>(j) * d + j] = 1.0;
    }
    int st = egx_gp_create(&m->ctx, prm->device, prm->corr, prm->mean, xn.data(), n, d, yn.data(), m->x_mean.data(),
                           m->x_std.data(), m->y_mean, m->y_std, m->w_star.data(), h, prm->nugget);
    if (st != EGX_OK) return st;
    egx_gp_dims(m->ctx, nullptr, nullptr, nullptr, &m->p);

    std::vector<double> theta_opt = theta0;
    if (prm->theta_tuning != EGX_THETA_FIXED) {
        // active set and bounds (algorithm.rs:815-827, 899-920)
        std::vector<int> active;
        if (prm->theta_tuning == EGX_THETA_PARTIAL) {
            for (int i = 0; i < prm->n_active; ++i) {
                if (prm->active[i] < 0 || prm->active[i] >= h) {
                    egx_set_error("active theta component %d out of range", prm->active[i]);
                    return EGX_INVALID_VALUE;
                }
                active.push_back(prm->active[i]);
            }
        } else {
            for (int i = 0; i < h; ++i) active.push_back(i);
        }
        std::vector<double> blo(h, 1e-2), bhi(h, 1e1);     // ThetaTuning::DEFAULT_BOUNDS
        if (prm->theta_bounds != nullptr && prm->n_theta_bounds > 0) {
            if (prm->n_theta_bounds == 1) {
                std::fill(blo.begin(), blo.end(), prm->theta_bounds[0]);
                std::fill(bhi.begin(), bhi.end(), prm->theta_bounds[1]);
            } else if (prm->n_theta_bounds == h) {
                for (int i = 0; i < h; ++i) {
                    blo[i] = prm->theta_bounds[2 * i];
                    bhi[i] = prm->theta_bounds[2 * i + 1];
                }
            } else {
                egx_set_error("Bounds for theta should be either 1-dim or dim of xtrain (%d), got %d", h,
                              prm->n_theta_bounds);
                return EGX_INVALID_VALUE;
            }
        }
        const int na = static_cast<int>(active.size());
        if (na > 0) {
            std::vector<double> lo(na), hi(na), z0(na);
            for (int i = 0; i < na; ++i) {
                lo[i] = std::log10(blo[active[i]]);
                hi[i] = std::log10(bhi[active[i]]);
                z0[i] = std::log10(theta0[active[i]]);
            }
            // prepare_multistart (optimization.rs:26-71)
            const int n_start = std::max(prm->n_start, 0);
            std::vector<std::vector<double>> starts;
            starts.push_back(z0);
            Xoshiro256Plus rng(prm->seed);
            if (n_start == 1) {
                std::vector<double> v(na);
                for (int i = 0; i < na; ++i) v[i] = lo[i] + (hi[i] - lo[i]) * rng.uniform();
                starts.push_back(v);
            } else if (n_start > 1) {
                std::vector<double> pts = lhs_maximin(n_start, na, rng);
                for (int s = 0; s < n_start; ++s) {
                    std::vector<double> v(na);
                    for (int i = 0; i < na; ++i) v[i] = lo[i] + (hi[i] - lo[i]) * pts[static_cast<size_t>(s) * na + i];
                    starts.push_back(v);
                }
            }
            const int maxeval = std::min(std::max(10 * na, GP_COBYLA_MIN_EVAL), std::max(prm->max_eval, 1));
            // chains sharded over processes (one per GPU): chain c belongs to rank c % world; the reduce by min of
            // algorithm.rs:942-945 is completed across ranks by prm->exchange below
            const int cworld = std::max(prm->chain_world, 1), crank = prm->chain_rank;
            if (crank < 0 || crank >= cworld) {
                egx_set_error("egx_gp_fit: chain_rank %d outside [0, %d)", crank, cworld);
                return EGX_INVALID_VALUE;
            }
            if (cworld > 1) {
                std::vector<std::vector<double>> mine;
                for (size_t c = 0; c < starts.size(); ++c)
                    if (static_cast<int>(c % cworld) == crank) mine.push_back(starts[c]);
                starts.swap(mine);
            }
            std::vector<BoundCobyla> chains;
            chains.reserve(starts.size());
            for (auto& s0 : starts) chains.emplace_back(s0, lo, hi, prm->cobyla_rhobeg, prm->cobyla_ftol_rel, maxeval);

            auto theta_of = [&](const std::vector<double>& z) {
                std::vector<double> th = theta0;
                for (int i = 0; i < na; ++i) th[active[i]] = std::pow(10.0, z[i]);
                return th;
            };
            // reduce (algorithm.rs:942-945): first strictly smaller wins, default theta = 1 (log10 = 0)
            double fbest = kInf;
            std::vector<double> zbest(na, 0.0);
            if (prm->optimizer == EGX_OPT_LBFGSB && h > 32) {
                egx_set_error("EGX_OPT_LBFGSB needs the closed-form theta gradient (at most 32 components, got %d)", h);
                return EGX_INVALID_VALUE;
            }
            if (prm->optimizer == EGX_OPT_LBFGSB) {
                // gradient-based multistart: one projected L-BFGS run per start on
                // f(z) = -rlf(10^z), df/dz_i = -ln(10) theta_i d rlf / d theta_i, same starts and budget as the chains
                struct GradUser {
                    egx_gp_ctx* ctx;
                    const std::vector<int>* active;
                    const std::vector<double>* theta0;
                    long long* n_evals;
                    int h;
                    int fatal;
                } gu{m->ctx, &active, &theta0, &m->n_evals, h, EGX_OK};
                auto fg = [](const double* z, int nz, double* grad, void* user) -> double {
                    GradUser* u = static_cast<GradUser*>(user);
                    std::vector<double> th = *u->theta0, gth(u->h, 0.0);
                    for (int i = 0; i < nz; ++i) th[(*u->active)[i]] = std::pow(10.0, z[i]);
                    double v = NAN;
                    const int s1 = egx_gp_reduced_likelihood_grad_analytic(u->ctx, th.data(), &v, gth.data());
                    *u->n_evals += 1;
                    if (s1 == EGX_CUDA_ERROR) u->fatal = s1;
                    for (int i = 0; i < nz; ++i) {
                        const int a = (*u->active)[i];
                        grad[i] = -2.302585092994046 * th[a] * gth[a];
                    }
                    return (s1 == EGX_OK && !std::isnan(v)) ? -v : kInf;      // Err(_) -> +inf, algorithm.rs:893-896
                };
                for (auto& s0 : starts) {
                    std::vector<double> zopt(na);
                    double fopt = kInf;
                    int nev = 0;
                    st = egx_bound_lbfgs_minimize(fg, &gu, na, s0.data(), lo.data(), hi.data(), 1e-9, 1e-7, maxeval,
                                                  zopt.data(), &fopt, &nev);
                    if (st != EGX_OK) return st;
                    if (gu.fatal != EGX_OK) return gu.fatal;
                    if (fopt < fbest) {
                        fbest = fopt;
                        zbest = zopt;
                    }
                }
                if (!std::isfinite(fbest) && prm->exchange == nullptr) {
                    egx_set_error("egx_gp_fit (EGX_OPT_LBFGSB): the likelihood could not be evaluated at any multistart point");
                    return EGX_NOT_POSITIVE_DEFINITE;
                }
            } else {
            static const bool lockstep_forced = getenv("EGX_FIT_LOCKSTEP") != nullptr && atoi(getenv("EGX_FIT_LOCKSTEP")) != 0;
            const int slots = lockstep_forced ? 0 : egx_gp_async_slots(m->ctx, static_cast<int>(chains.size()));
            if (slots >= 2) {
                // Independent chains, as in the reference (one rayon task per start): `slots` evaluations stay in
                // flight; whenever the oldest one is back its chain is told the value and asks for the next theta,
                // which goes out on the same slot at once.  Every chain sees exactly the sequence it would see alone.
                std::vector<int> chain_of(slots, -1);          // slot -> chain with an evaluation in flight
                std::vector<char> busy(chains.size(), 0);
                size_t next_chain = 0;
                int head = 0, in_flight = 0;
                auto feed = [&](int slot) -> int {
                    for (size_t tries = 0; tries < chains.size(); ++tries) {
                        const size_t ci = (next_chain + tries) % chains.size();
                        if (busy[ci] || chains[ci].done()) continue;
                        const std::vector<double> th = theta_of(chains[ci].ask());
                        const int s1 = egx_gp_eval_begin(m->ctx, slot, th.data());
                        if (s1 == EGX_CUDA_ERROR) return s1;
                        m->n_evals += 1;
                        if (s1 != EGX_OK) {                      // rejected before launch (invalid theta): Err(_) -> +inf
                            chains[ci].tell(kInf);
                            --tries;
                            continue;
                        }
                        busy[ci] = 1;
                        chain_of[slot] = static_cast<int>(ci);
                        next_chain = ci + 1;
                        ++in_flight;
                        return EGX_OK;
                    }
                    chain_of[slot] = -1;
                    return EGX_OK;
                };
                // (measured and dropped, profiles/r02/y16_stagger.txt: starting the chains 1-4 ms apart, so that the bulk-heavy
                //  first half of one evaluation would meet the chain-bound second half of another, changes nothing: 3.88-3.90 ms
                //  per evaluation with two chains at n = 8192 whatever the offset)
                for (int sl = 0; sl < slots; ++sl) {
                    st = feed(sl);
                    if (st != EGX_OK) return st;
                }
                while (in_flight > 0) {
                    while (chain_of[head] < 0) head = (head + 1) % slots;
                    const int ci = chain_of[head];
                    double v = NAN;
                    const int s2 = egx_gp_eval_end(m->ctx, head, &v);
                    if (s2 == EGX_CUDA_ERROR) return s2;
                    --in_flight;
                    busy[ci] = 0;
                    double f = (s2 == EGX_OK) ? -v : kInf;      // Err(_) -> +inf, algorithm.rs:893-896
                    if (std::isnan(f)) f = kInf;                // optimization.rs:157-161
                    chains[ci].tell(f);
                    st = feed(head);                            // the slot is free again: next chain in line
                    if (st != EGX_OK) return st;
                    head = (head + 1) % slots;
                }
            } else {
            std::vector<double> thetas, rlf;
            std::vector<int> status, who;
            for (;;) {
                thetas.clear();
                who.clear();
                for (size_t c = 0; c < chains.size(); ++c) {
                    if (chains[c].done()) continue;
                    const std::vector<double> th = theta_of(chains[c].ask());
                    thetas.insert(thetas.end(), th.begin(), th.end());
                    who.push_back(static_cast<int>(c));
                }
                if (who.empty()) break;
                const int B = static_cast<int>(who.size());
                rlf.assign(B, NAN);
                status.assign(B, 0);
                st = egx_gp_reduced_likelihood_batch(m->ctx, thetas.data(), B, rlf.data(), status.data());
                if (st != EGX_OK) return st;
                m->n_evals += B;
                for (int b = 0; b < B; ++b) {
                    // Err(_) -> +inf, algorithm.rs:893-896 ; NaN -> +inf, optimization.rs:157-161
                    double f = (status[b] == EGX_OK) ? -rlf[b] : kInf;
                    if (std::isnan(f)) f = kInf;
                    chains[who[b]].tell(f);
                }
            }
            }
            for (auto& ch : chains)
                if (ch.best_f() < fbest) {
                    fbest = ch.best_f();
                    zbest = ch.best_x();
                }
            }
            if (prm->exchange != nullptr) {
                if (prm->exchange(&fbest, zbest.data(), na, prm->exchange_user) != 0) {
                    egx_set_error("egx_gp_fit: the exchange callback of the sharded multistart failed");
                    return EGX_CUDA_ERROR;
                }
            }
            // algorithm.rs:947-964
            if (prm->theta_tuning == EGX_THETA_PARTIAL) {
                theta_opt = theta0;
                for (int i = 0; i < na; ++i) theta_opt[active[i]] = std::pow(10.0, zbest[i]);
            } else {
                for (int i = 0; i < na; ++i) theta_opt[i] = std::pow(10.0, zbest[i]);
            }
        }
    }

    double rlf = NAN, s2 = NAN;
    st = egx_gp_finalize(m->ctx, theta_opt.data(), &rlf, &s2, nullptr, nullptr, nullptr, nullptr);
    m->n_evals += 1;
    egx_gp_release_workspaces(m->ctx);  // the multistart workspaces (up to 11 x 0.6 GB at n = 8192) are not needed by predict*
    if (st != EGX_OK) return st;        // `?` at algorithm.rs:967-968
    m->theta = theta_opt;
    m->likelihood = rlf;
    m->sigma2 = s2;
    *out = m.release();
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_model_dims(const egx_gp_model* m, int* n, int* d, int* h, int* p) try {
    if (!m) return EGX_INVALID_VALUE;
    if (n) *n = m->n;
    if (d) *d = m->d;
    if (h) *h = m->h;
    if (p) *p = m->p;
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" int egx_gp_model_theta(const egx_gp_model* m, double* theta) try {
    if (!m || !theta) return EGX_INVALID_VALUE;
    std::memcpy(theta, m->theta.data(), sizeof(double) * m->h);
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" double egx_gp_model_variance(const egx_gp_model* m) { return m ? m->sigma2 : NAN; }
extern "C" double egx_gp_model_likelihood(const egx_gp_model* m) { return m ? m->likelihood : NAN; }
extern "C" long long egx_gp_model_n_evals(const egx_gp_model* m) { return m ? m->n_evals : 0; }
extern "C" egx_gp_ctx* egx_gp_model_context(egx_gp_model* m) { return m ? m->ctx : nullptr; }

extern "C" int egx_gp_model_inner_params(egx_gp_model* m, double* beta, double* gamma, double* r_chol, double* ft,
                                         double* ft_qr_r) try {
    if (!m) return EGX_INVALID_VALUE;
    // re-run the final evaluation to fetch the requested pieces (the factor stays on the device)
    double rlf, s2;
    int st = egx_gp_finalize(m->ctx, m->theta.data(), &rlf, &s2, beta, gamma, ft, ft_qr_r);
    if (st != EGX_OK) return st;
    if (r_chol) st = egx_gp_download_chol(m->ctx, r_chol);
    return st;
}
EGX_ABI_CATCH

extern "C" int egx_gp_model_normalization(const egx_gp_model* m, double* x_mean, double* x_std, double* y_mean,
                                          double* y_std, double* w_star) try {
    if (!m) return EGX_INVALID_VALUE;
    if (x_mean) std::memcpy(x_mean, m->x_mean.data(), sizeof(double) * m->d);
    if (x_std) std::memcpy(x_std, m->x_std.data(), sizeof(double) * m->d);
    if (y_mean) *y_mean = m->y_mean;
    if (y_std) *y_std = m->y_std;
    if (w_star) std::memcpy(w_star, m->w_star.data(), sizeof(double) * m->d * m->h);
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_model_predict(egx_gp_model* m, const double* x, int npts, double* y) try {
    if (!m) return EGX_INVALID_VALUE;
    return egx_gp_predict(m->ctx, x, npts, y);
}
EGX_ABI_CATCH
extern "C" int egx_gp_model_predict_var(egx_gp_model* m, const double* x, int npts, double* var) try {
    if (!m) return EGX_INVALID_VALUE;
    return egx_gp_predict_var(m->ctx, x, npts, var);
}
EGX_ABI_CATCH
extern "C" int egx_gp_model_predict_valvar(egx_gp_model* m, const double* x, int npts, double* y, double* var) try {
    if (!m) return EGX_INVALID_VALUE;
    return egx_gp_predict_valvar(m->ctx, x, npts, y, var);
}
EGX_ABI_CATCH

extern "C" int egx_gp_model_predict_var_gradients(egx_gp_model* m, const double* x, int npts, double* grad) try {
    if (!m) return EGX_INVALID_VALUE;
    return egx_gp_predict_var_gradients(m->ctx, x, npts, grad);
}
EGX_ABI_CATCH
extern "C" int egx_gp_model_covariance(egx_gp_model* m, const double* x, int npts, double* cov) try {
    if (!m) return EGX_INVALID_VALUE;
    return egx_gp_covariance(m->ctx, x, npts, cov);
}
EGX_ABI_CATCH
extern "C" int egx_gp_model_sample(egx_gp_model* m, const double* x, int npts, const double* z, int n_traj, int method,
                                   double* out) try {
    if (!m) return EGX_INVALID_VALUE;
    return egx_gp_sample(m->ctx, x, npts, z, n_traj, method, out);
}
EGX_ABI_CATCH
extern "C" int egx_gp_model_predict_gradients(egx_gp_model* m, const double* x, int npts, double* grad) try {
    if (!m) return EGX_INVALID_VALUE;
    return egx_gp_predict_gradients(m->ctx, x, npts, grad);
}
EGX_ABI_CATCH

// ============================================================================================
// Sparse GP fit driver: impl Fit for SgpValidParams, sparse_algorithm.rs:416-648.
// Optimisation variables = log10 of [theta_1..theta_h, sigma2, (noise)].
// ============================================================================================
struct egx_sgp_model {
    egx_sgp_ctx* ctx = nullptr;
    int n = 0, d = 0, h = 0, m = 0;
    std::vector<double> theta, z;
    double sigma2 = NAN, noise = NAN, likelihood = NAN;
    long long n_evals = 0;
};

extern "C" void egx_sgp_params_default(egx_sgp_params* p) {
    if (!p) return;
    std::memset(p, 0, sizeof(*p));
    p->corr = EGX_CORR_SQUARED_EXPONENTIAL;
    p->method = EGX_SGP_FITC;
    p->noise_init = 1e-2;                                   // sparse_parameters.rs:25-32
    p->noise_lo = 100.0 * 2.220446049250313e-16;
    p->noise_hi = 1e10;
    p->n_inducings = 10;                                    // Inducings::Randomized(10), :44-48
    p->n_start = 10;
    p->max_eval = 1000;
    p->nugget = 100.0 * 2.220446049250313e-16;
    p->seed = 42;
    p->cobyla_rhobeg = 0.5;
    p->cobyla_ftol_rel = 1e-4;
}

extern "C" void egx_sgp_model_destroy(egx_sgp_model* m) {
    if (!m) return;
    if (m->ctx) egx_sgp_destroy(m->ctx);
    delete m;
}

extern "C" int egx_sgp_fit(const egx_sgp_params* prm, const double* x, int n, int d, const double* y,
                           egx_sgp_model** out) try {
    if (!out) return EGX_INVALID_VALUE;
    *out = nullptr;
    if (!prm || !x || !y || n < 2 || d < 1) {
        egx_set_error("egx_sgp_fit: invalid argument");
        return EGX_INVALID_VALUE;
    }
    const bool kpls = prm->w_star != nullptr || prm->kpls_dim > 0;
    if (kpls && (prm->kpls_dim < 1 || prm->kpls_dim > d)) {
        egx_set_error("Dimension reduction %d should be smaller than actual training input dimensions %d",
                      prm->kpls_dim, d);
        return EGX_INVALID_VALUE;
    }
    const int h = kpls ? prm->kpls_dim : d;
    std::unique_ptr<egx_sgp_model, void (*)(egx_sgp_model*)> m(new egx_sgp_model(), egx_sgp_model_destroy);
    m->n = n;
    m->d = d;
    m->h = h;
    std::vector<double> w;
    if (kpls && prm->w_star) w.assign(prm->w_star, prm->w_star + static_cast<size_t>(d) * h);
    else if (kpls) {
        // sparse_algorithm.rs:442-455
        w.assign(static_cast<size_t>(d) * h, 0.0);
        const int stp = egx_pls_rotations(x, n, d, y, h, w.data());
        if (stp != EGX_OK) return stp;
    } else {
        w.assign(static_cast<size_t>(d) * d, 0.0);
        for (int j = 0; j < d; ++j) w[static_cast<size_t>(j) * d + j] = 1.0;
    }
    Xoshiro256Plus rng(prm->seed);
    // inducing points (:455-458, make_inducings :833-847)
    if (prm->z != nullptr) {
        m->m = prm->n_inducings;
        m->z.assign(prm->z, prm->z + static_cast<size_t>(m->m) * d);
    } else {
        std::vector<int> idx(n);
        for (int i = 0; i < n; ++i) idx[i] = i;
        rng.shuffle(idx);
        m->m = std::min(prm->n_inducings, n);
        m->z.resize(static_cast<size_t>(m->m) * d);
        for (int r = 0; r < m->m; ++r)
            for (int j = 0; j < d; ++j) m->z[static_cast<size_t>(r) * d + j] = x[static_cast<size_t>(idx[r]) * d + j];
    }
    if (m->m < 1) {
        egx_set_error("egx_sgp_fit: no inducing points");
        return EGX_INVALID_VALUE;
    }
    int st = egx_sgp_create(&m->ctx, prm->device, prm->corr, prm->method, x, n, d, y, m->z.data(), m->m, w.data(), h,
                            prm->nugget);
    if (st != EGX_OK) return st;

    const bool noise_est = !prm->noise_fixed;
    // theta0 (:476-488)
    std::vector<double> theta0(h);
    const int n_init = (prm->theta_init != nullptr) ? prm->n_theta_init : 0;
    if (n_init == 0) std::fill(theta0.begin(), theta0.end(), 0.1);
    else if (n_init == 1) std::fill(theta0.begin(), theta0.end(), prm->theta_init[0]);
    else if (n_init == h) theta0.assign(prm->theta_init, prm->theta_init + h);
    else {
        egx_set_error("Initial guess for theta should be either 1-dim or dim of xtrain (w_star.ncols()), got %d", n_init);
        return EGX_INVALID_VALUE;
    }
    // sigma2_0 = var(y, ddof = 1)  (:491-492)
    double mean = 0.0;
    for (int i = 0; i < n; ++i) mean += y[i];
    mean /= n;
    double var = 0.0;
    for (int i = 0; i < n; ++i) var += (y[i] - mean) * (y[i] - mean);
    const double sigma2_0 = var / (n - 1);
    const int np = h + 1 + (noise_est ? 1 : 0);
    std::vector<double> p0(np);
    for (int i = 0; i < h; ++i) p0[i] = theta0[i];
    p0[h] = sigma2_0;
    if (noise_est) p0[np - 1] = prm->noise_init;
    // bounds (:544-577): theta bounds broadcast to every parameter, then variance and noise overrides
    std::vector<double> lo(np), hi(np);
    {
        double blo = 1e-2, bhi = 1e2;                       // sparse_parameters.rs:160-163
        const int nb = (prm->theta_bounds != nullptr) ? prm->n_theta_bounds : 0;
        if (prm->theta_fixed) {
            for (int i = 0; i < np; ++i) {
                const double v = (i < h) ? theta0[i] : theta0[0];
                lo[i] = hi[i] = std::log10(v);
            }
        } else if (nb <= 1) {
            if (nb == 1) {
                blo = prm->theta_bounds[0];
                bhi = prm->theta_bounds[1];
            }
            for (int i = 0; i < np; ++i) {
                lo[i] = std::log10(blo);
                hi[i] = std::log10(bhi);
            }
        } else if (nb == np) {
            for (int i = 0; i < np; ++i) {
                lo[i] = std::log10(prm->theta_bounds[2 * i]);
                hi[i] = std::log10(prm->theta_bounds[2 * i + 1]);
            }
        } else {
            egx_set_error("Bounds for theta should be either 1-dim or dim of the parameter vector (%d), got %d", np, nb);
            return EGX_INVALID_VALUE;
        }
    }
    // multistart seeds (:566) are drawn in the box BEFORE the variance / noise overrides below, i.e. every
    // parameter (theta, sigma2, noise) starts inside the theta bounds -- exactly like the reference, which calls
    // prepare_multistart first and only then rewrites bounds[sigma2] and bounds[noise] (:568-584)
    const int n_start = std::max(prm->n_start, 0);
    std::vector<std::vector<double>> starts;
    {
        std::vector<double> z0(np);
        for (int i = 0; i < np; ++i) z0[i] = std::log10(p0[i]);
        starts.push_back(z0);
        if (n_start == 1) {
            std::vector<double> v(np);
            for (int i = 0; i < np; ++i) v[i] = lo[i] + (hi[i] - lo[i]) * rng.uniform();
            starts.push_back(v);
        } else if (n_start > 1) {
            std::vector<double> pts = lhs_maximin(n_start, np, rng);
            for (int s = 0; s < n_start; ++s) {
                std::vector<double> v(np);
                for (int i = 0; i < np; ++i) v[i] = lo[i] + (hi[i] - lo[i]) * pts[static_cast<size_t>(s) * np + i];
                starts.push_back(v);
            }
        }
    }
    {
        lo[h] = std::log10(1e-12);
        hi[h] = std::log10(9.0 * sigma2_0);
        if (noise_est) {
            lo[np - 1] = std::log10(prm->noise_lo);
            hi[np - 1] = std::log10(prm->noise_hi);
        }
    }
    const int maxeval = std::min(std::max(10 * std::max(n_init, 1), GP_COBYLA_MIN_EVAL), std::max(prm->max_eval, 1));  // :598-600
    std::vector<BoundCobyla> chains;
    for (auto& s0 : starts) chains.emplace_back(s0, lo, hi, prm->cobyla_rhobeg, prm->cobyla_ftol_rel, maxeval);
    std::vector<double> th(h);
    auto eval = [&](const std::vector<double>& z, double* lik) {
        for (int i = 0; i < h; ++i) th[i] = std::pow(10.0, z[i]);
        const double s2 = std::pow(10.0, z[h]);
        const double nz = noise_est ? std::pow(10.0, z[np - 1]) : prm->noise_init;
        return egx_sgp_reduced_likelihood(m->ctx, th.data(), s2, nz, lik);
    };
    for (;;) {
        bool any = false;
        for (auto& ch : chains) {
            if (ch.done()) continue;
            any = true;
            double lik = NAN;
            st = eval(ch.ask(), &lik);
            if (st == EGX_CUDA_ERROR) return st;
            m->n_evals += 1;
            double f = (st == EGX_OK) ? -lik : kInf;
            if (std::isnan(f)) f = kInf;
            ch.tell(f);
        }
        if (!any) break;
    }
    double fbest = kInf;
    std::vector<double> zbest(np, 0.0);
    for (auto& ch : chains)
        if (ch.best_f() < fbest) {
            fbest = ch.best_f();
            zbest = ch.best_x();
        }
    m->theta.resize(h);
    for (int i = 0; i < h; ++i) m->theta[i] = std::pow(10.0, zbest[i]);
    m->sigma2 = std::pow(10.0, zbest[h]);
    m->noise = noise_est ? std::pow(10.0, zbest[np - 1]) : prm->noise_init;
    double lik = NAN;
    st = egx_sgp_finalize(m->ctx, m->theta.data(), m->sigma2, m->noise, &lik, nullptr, nullptr);
    m->n_evals += 1;
    if (st != EGX_OK) return st;
    m->likelihood = lik;
    *out = m.release();
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_sgp_model_dims(const egx_sgp_model* m, int* n, int* d, int* h, int* nz) try {
    if (!m) return EGX_INVALID_VALUE;
    if (n) *n = m->n;
    if (d) *d = m->d;
    if (h) *h = m->h;
    if (nz) *nz = m->m;
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" int egx_sgp_model_theta(const egx_sgp_model* m, double* theta) try {
    if (!m || !theta) return EGX_INVALID_VALUE;
    std::memcpy(theta, m->theta.data(), sizeof(double) * m->h);
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" double egx_sgp_model_variance(const egx_sgp_model* m) { return m ? m->sigma2 : NAN; }
extern "C" double egx_sgp_model_noise_variance(const egx_sgp_model* m) { return m ? m->noise : NAN; }
extern "C" double egx_sgp_model_likelihood(const egx_sgp_model* m) { return m ? m->likelihood : NAN; }
extern "C" long long egx_sgp_model_n_evals(const egx_sgp_model* m) { return m ? m->n_evals : 0; }
extern "C" int egx_sgp_model_inducings(const egx_sgp_model* m, double* z) try {
    if (!m || !z) return EGX_INVALID_VALUE;
    std::memcpy(z, m->z.data(), sizeof(double) * m->z.size());
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" int egx_sgp_model_woodbury(egx_sgp_model* m, double* w_vec, double* w_inv) try {
    if (!m) return EGX_INVALID_VALUE;
    double lik;
    return egx_sgp_finalize(m->ctx, m->theta.data(), m->sigma2, m->noise, &lik, w_vec, w_inv);
}
EGX_ABI_CATCH
extern "C" egx_sgp_ctx* egx_sgp_model_context(egx_sgp_model* m) { return m ? m->ctx : nullptr; }
extern "C" int egx_sgp_model_predict(egx_sgp_model* m, const double* x, int npts, double* y) try {
    if (!m) return EGX_INVALID_VALUE;
    return egx_sgp_predict(m->ctx, x, npts, y);
}
EGX_ABI_CATCH
extern "C" int egx_sgp_model_sample(egx_sgp_model* m, const double* x, int npts, const double* z, int n_traj, int method,
                                    double* out) try {
    if (!m) return EGX_INVALID_VALUE;
    return egx_sgp_sample(m->ctx, x, npts, z, n_traj, method, out);
}
EGX_ABI_CATCH
extern "C" int egx_sgp_model_predict_var(egx_sgp_model* m, const double* x, int npts, double* var) try {
    if (!m) return EGX_INVALID_VALUE;
    return egx_sgp_predict_var(m->ctx, x, npts, var);
}
EGX_ABI_CATCH
