/* egobox_gpu.h -- C ABI of the B200-native kriging hot path (libegobox_gpu.so).
 *
 * This is the drop-in boundary for egobox-gp's kriging training / prediction
 * path.  The reference (relf/egobox @ be16128) has no FFI for this path; its
 * seams are Rust traits and one free function, and each entry point below
 * names the reference interface it replaces (file:line under
 * /root/reference/crates/).  INTEGRATION.md shows the Rust `extern "C"`
 * binding a maintainer would add under `cfg(feature = "cuda")`.
 *
 * Conventions
 *  - f64 everywhere, arrays row-major (ndarray C order), caller owns every
 *    host buffer, the opaque handle owns all device state.
 *  - Every function returns an `int` status (EGX_*).  Numerical failures are
 *    statuses, never aborts, so a caller can map them to `+inf` exactly like
 *    gp/src/algorithm.rs:893-896 does with `Err(_)`.
 *  - Calls on one handle are serialised by a handle-level mutex (the
 *    reference calls `reduced_likelihood` from rayon workers,
 *    gp/src/algorithm.rs:928-945); different handles are independent.
 *  - There is no CPU fallback: without a CUDA device every call that needs
 *    one returns EGX_CUDA_ERROR.
 */
#ifndef EGOBOX_GPU_H
#define EGOBOX_GPU_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes --------------------------------------------------------
 * 1 <-> GpError::LinalgError (gp/src/errors.rs:19, Cholesky of a non-PD R,
 *       gp/src/algorithm.rs:1004)
 * 2,3 <-> GpError::LikelihoodComputationError (gp/src/algorithm.rs:1012-1026)
 * 4 <-> GpError::InvalidValueError (gp/src/errors.rs:38)                     */
#define EGX_OK                    0
#define EGX_NOT_POSITIVE_DEFINITE 1
#define EGX_ILL_CONDITIONED_FT    2
#define EGX_ILL_CONDITIONED_F     3
#define EGX_INVALID_VALUE         4
#define EGX_CUDA_ERROR            5

/* correlation models, gp/src/correlation_models.rs:87,181,273,442 */
#define EGX_CORR_SQUARED_EXPONENTIAL  0
#define EGX_CORR_ABSOLUTE_EXPONENTIAL 1
#define EGX_CORR_MATERN32             2
#define EGX_CORR_MATERN52             3

/* regression (mean) models, gp/src/mean_models.rs:39,65,94 */
#define EGX_MEAN_CONSTANT  0
#define EGX_MEAN_LINEAR    1
#define EGX_MEAN_QUADRATIC 2

typedef struct egx_gp_ctx egx_gp_ctx;

/* Library / device probe.  Returns the number of visible CUDA devices (0 if
 * none); never fails. */
int egx_device_count(void);
/* Thread-local description of the last non-OK status on this thread. */
const char* egx_last_error(void);
const char* egx_version(void);

/* ---- model context --------------------------------------------------------
 * Holds the device-resident training set of one GP: normalised inputs,
 * normalised output, the regression basis F = mean.value(xnorm), PLS weights.
 * Replaces the state captured by the `objfn` closure in
 * gp/src/algorithm.rs:856-897 (xtrain, ytrain, x_distances, fx, w_star) --
 * the DiffMatrix (gp/src/utils.rs:58-105) is never materialised.
 *
 *   xnorm  n x d   normalised training inputs   (NormalizedData, utils.rs:28-54)
 *   ynorm  n       normalised training output
 *   x_mean,x_std d ; y_mean,y_std  -- kept to (de)normalise in predict*
 *   w_star d x h   identity (h == d) or PLS rotations (algorithm.rs:843-855)
 *   nugget         (1+nugget) on the diagonal of R (algorithm.rs:997)
 */
int egx_gp_create(egx_gp_ctx** out, int device, int corr, int mean,
                  const double* xnorm, int n, int d, const double* ynorm,
                  const double* x_mean, const double* x_std,
                  double y_mean, double y_std,
                  const double* w_star, int h, double nugget);
void egx_gp_destroy(egx_gp_ctx* ctx);

/* sizes: n, d, h (theta length), p (regression basis size) */
int egx_gp_dims(const egx_gp_ctx* ctx, int* n, int* d, int* h, int* p);

/* Reduced likelihood at theta (length h).
 * Replaces `corr.value(D, theta, W)` + `reduced_likelihood(..)`:
 * gp/src/algorithm.rs:892-893 and :989-1056 (value `.0` only).
 * On status != EGX_OK, *rlf is NaN and the caller maps it to +inf. */
int egx_gp_reduced_likelihood(egx_gp_ctx* ctx, const double* theta, double* rlf);

/* Batched form of the above for B candidate thetas (row-major B x h):
 * multistart chains advanced in lock step (gp/src/algorithm.rs:928-945) and
 * theta sweeps.  status[b] holds the per-candidate status; the return value
 * is EGX_OK unless the call itself failed (bad argument, CUDA error). */
int egx_gp_reduced_likelihood_batch(egx_gp_ctx* ctx, const double* thetas, int B,
                                    double* rlf, int* status);

/* Reduced likelihood and its theta-gradient by central differences, d rlf / d theta_k ~
 * (rlf(theta + delta_k e_k) - rlf(theta - delta_k e_k)) / (2 delta_k), delta_k = rel_step * theta_k, evaluated as ONE
 * batch of 2h+1 likelihoods (kept in flight together).  The reference has no theta-gradient
 * (gp/src/algorithm.rs:880 ignores `_gradient`; COBYLA is derivative free); this is the batched building block
 * for gradient-based callers.  Returns the status of the centre point; grad[k] is NaN where a side point failed. */
int egx_gp_reduced_likelihood_grad(egx_gp_ctx* ctx, const double* theta, double rel_step, double* rlf,
                                   double* grad);

/* The same gradient in closed form (h <= 32):
 *   d rlf / d theta_l = [ gamma^T (dR/dtheta_l) gamma / sigma2 - tr(R^-1 dR/dtheta_l) ] / ln 10
 * from ONE factorisation: R^-1 = W W^T with W = L^-T (multi-RHS sweep on the identity + SYRK), then one pass over the
 * pairs.  About three factorisations' worth of work instead of 2h+1, and exact to rounding.  grad is d/d theta (not
 * d/d log10 theta); leaves the context untrained, like egx_gp_reduced_likelihood.  Uses 2 n^2 doubles of extra memory. */
int egx_gp_reduced_likelihood_grad_analytic(egx_gp_ctx* ctx, const double* theta, double* rlf, double* grad);

/* Final evaluation at the selected theta (gp/src/algorithm.rs:966-968): keeps
 * the Cholesky factor, gamma, beta, Ft, G on the device for predict*, and
 * optionally returns GpInnerParams (algorithm.rs:47-60) -- any output pointer
 * may be NULL.  sigma2 is already multiplied by y_std^2 (algorithm.rs:1048).
 *   beta p ; gamma n ; ft n x p ; ft_qr_r p x p (upper, diag > 0)            */
int egx_gp_finalize(egx_gp_ctx* ctx, const double* theta, double* rlf, double* sigma2,
                    double* beta, double* gamma, double* ft, double* ft_qr_r);

/* r_chol (n x n, lower, upper part zeroed) for serde save
 * (GpInnerParams.r_chol, gp/src/algorithm.rs:54-55). */
int egx_gp_download_chol(egx_gp_ctx* ctx, double* r_chol);

/* Prediction at m raw (un-normalised) points x (m x d).  Require a prior
 * successful egx_gp_finalize.
 *   predict        gp/src/algorithm.rs:253-263
 *   predict_var    gp/src/algorithm.rs:267-279  (+ _compute_rt_u :330-369)
 *   predict_valvar gp/src/algorithm.rs:282-307                               */
int egx_gp_predict(egx_gp_ctx* ctx, const double* x, int m, double* y);
int egx_gp_predict_var(egx_gp_ctx* ctx, const double* x, int m, double* var);
int egx_gp_predict_valvar(egx_gp_ctx* ctx, const double* x, int m, double* y, double* var);

/* Batched prediction gradients d yhat / d x at m raw points: grad is m x d (row-major).
 *   predict_gradients gp/src/algorithm.rs:518-529 (one `predict_jacobian` :533-566 per point in the
 *   reference; one warp per point here).  Supports d <= 32. */
int egx_gp_predict_gradients(egx_gp_ctx* ctx, const double* x, int m, double* grad);
/* Batched variance gradients d var / d x (m x d): predict_var_gradients gp/src/algorithm.rs:697-704, one
 * `predict_var_gradients_single` :554-616 (four n x n triangular solves) per point in the reference; here one
 * forward and one backward multi-RHS sweep per chunk of points.  Supports d <= 32 and d * p <= 2048. */
int egx_gp_predict_var_gradients(egx_gp_ctx* ctx, const double* x, int m, double* grad);

/* Conditional covariance of the trained GP at m locations, cov (m x m, row-major) =
 *   sigma2 * (K(x, x) - rt^T rt + u^T u)
 * `_compute_covariance` gp/src/algorithm.rs:310-326 (rt, u as in predict_var, :330-369).  1 <= m <= 8192. */
int egx_gp_covariance(egx_gp_ctx* ctx, const double* x, int m, double* cov);

/* Trajectory sampling, `sample_chol` / `sample_eig` / `sample` gp/src/algorithm.rs:383-410 and the free function
 * `sample` :1153-1194:  out (m x n_traj) = predict(x) 1^T + C z, where C C^T is the conditional covariance, factored
 * by Cholesky (on the device, same blocked factorisation as the likelihood) or by eigen-decomposition with
 * eigenvalues below 1e-9 dropped (:1175-1181; m x m symmetric eigenproblem on the host, as in the reference).
 * The reference draws z ~ N(0, 1) itself (ndarray-rand, :1191-1192); here z (m x n_traj, row-major) is an argument
 * so that callers own the random stream and results are reproducible.
 * EGX_NOT_POSITIVE_DEFINITE when the Cholesky variant meets a non-positive pivot (the reference panics, :1164). */
#define EGX_SAMPLE_CHOLESKY 0
#define EGX_SAMPLE_EIGENVALUES 1
int egx_gp_sample(egx_gp_ctx* ctx, const double* x, int m, const double* z, int n_traj, int method, double* out);

/* Asynchronous form for INDEPENDENT optimiser chains -- the reference runs its n_start + 1 COBYLA chains on rayon
 * threads, each calling reduced_likelihood on its own (gp/src/algorithm.rs:928-945).  egx_gp_async_slots prepares up
 * to `wanted` workspaces and returns how many there are (0: use the batched call -- small problems run a whole batch
 * as one launch); a slot holds at most one evaluation in flight: egx_gp_eval_begin enqueues it and returns,
 * egx_gp_eval_end waits for it and returns the status / value of egx_gp_reduced_likelihood.  Different slots may be
 * driven from different threads concurrently (one rayon worker per slot); one slot, one caller at a time. */
int egx_gp_async_slots(egx_gp_ctx* ctx, int wanted);
/* Returns the extra workspaces created by egx_gp_reduced_likelihood_batch / egx_gp_async_slots to the block cache (they
 * are re-created on demand).  egx_gp_fit calls it once the optimum is found, so a trained model holds ONE workspace. */
int egx_gp_release_workspaces(egx_gp_ctx* ctx);
int egx_gp_eval_begin(egx_gp_ctx* ctx, int slot, const double* theta);
int egx_gp_eval_end(egx_gp_ctx* ctx, int slot, double* rlf);

/* Same, with x / y / var already resident on the context's device (device
 * pointers).  Used to time the kernels without the PCIe copies. */
int egx_gp_predict_valvar_dev(egx_gp_ctx* ctx, const double* x_dev, int m,
                              double* y_dev, double* var_dev);
/* Device-pointer variants of the prediction gradients (x_dev m x d, grad_dev m x d), used by the mixture. */
int egx_gp_predict_gradients_dev(egx_gp_ctx* ctx, const double* x_dev, int m, double* grad_dev);
int egx_gp_predict_var_gradients_dev(egx_gp_ctx* ctx, const double* x_dev, int m, double* grad_dev);

/* ---- building blocks exposed for parity tests ------------------------------
 * R(theta) as assembled at gp/src/algorithm.rs:997-1001 (full symmetric n x n,
 * (1+nugget) on the diagonal), computed by the fused pairwise-distance +
 * correlation kernel. */
int egx_gp_correlation_matrix(egx_gp_ctx* ctx, const double* theta, double* r);
/* c(x*, X) as in `_compute_correlation`, gp/src/algorithm.rs:372-380 (m x n),
 * for the theta of the last finalize. */
int egx_gp_cross_correlation(egx_gp_ctx* ctx, const double* x, int m, double* c);

/* ---- instrumentation --------------------------------------------------------
 * Per-stage device timings (CUDA events on the context's stream) accumulated
 * since the last reset, for bench.py's roofline block.  Stage ids: */
#define EGX_STAGE_CORR_BUILD   0  /* fused distance + correlation, R lower tiles */
#define EGX_STAGE_POTRF_DIAG   1  /* diagonal-panel Cholesky                    */
#define EGX_STAGE_TRSM_PANEL   2  /* panel / row triangular solves               */
#define EGX_STAGE_SYRK_GEMM    3  /* trailing SYRK / GEMM updates (DMMA)         */
#define EGX_STAGE_GLS          4  /* thin QR, beta, rho, sigma2, logdet          */
#define EGX_STAGE_BACKSOLVE    5  /* gamma = L^-T rho                            */
#define EGX_STAGE_CROSS_CORR   6  /* c(x*, X) (+ fused mean / gamma GEMV)        */
#define EGX_STAGE_VAR_FINISH   7  /* row norms, u, variance                      */
#define EGX_STAGE_SMALL_BATCH  8  /* one-CTA-per-theta small-n likelihood        */
#define EGX_STAGE_GEMM_LOOKAHEAD 9 /* updates issued ahead on the panel stream (partner column, next pair's two columns) */
#define EGX_STAGE_OZAKI_SLICE  10 /* fp64 panel pair -> row scales + 8 int8 slices (tcgen05 path)           */
#define EGX_STAGE_OZAKI_SYRK   11 /* trailing SYRK update on tcgen05 (int8-sliced, UTCIMMA, TMEM accumulators) */
#define EGX_STAGE_THETA_GRAD   12 /* closed-form theta gradient: pair kernel + partial-sum reduction */
#define EGX_NUM_STAGES         13
int egx_gp_set_profiling(egx_gp_ctx* ctx, int enabled);
int egx_gp_reset_profile(egx_gp_ctx* ctx);
/* ms[EGX_NUM_STAGES], launches[EGX_NUM_STAGES] */
int egx_gp_get_profile(egx_gp_ctx* ctx, double* ms, long long* launches);
/* Device timer on the context's own stream (CUDA events): start records an event, stop
 * records a second one, waits for it and returns the elapsed milliseconds in between
 * (includes any GPU idle time while the host prepares the next launch). */
int egx_gp_timer_start(egx_gp_ctx* ctx);
int egx_gp_timer_stop(egx_gp_ctx* ctx, double* elapsed_ms);
/* Enable / disable the two-stream look-ahead of the factorisation (default on; EGX_LOOKAHEAD=0 also disables).
 * With look-ahead off every kernel of an evaluation runs back to back on one stream, which is what the
 * per-kernel roofline pass of bench.py times. */
int egx_gp_set_lookahead(egx_gp_ctx* ctx, int enabled);
/* Force the blocked large-n path even when n is small enough for the
 * one-CTA-per-theta kernel (tests exercise both on the same inputs). */
int egx_gp_set_force_blocked(egx_gp_ctx* ctx, int enabled);


/* ============================================================================
 * Host-level model API (C++ host code above the device seam, same library).
 * Mirrors `GpParams::fit(&Dataset) -> GaussianProcess` and the inherent
 * accessors / predict* of gp/src/algorithm.rs:242-439, 785-979 so that the
 * moe `GpSurrogate` wrappers (moe/src/surrogates.rs:147-155) and the PyO3
 * `Gpx` class (python/src/gp_mix.rs:242-496) can bind it one to one.
 * ========================================================================== */
#define EGX_THETA_FIXED   0   /* ThetaTuning::Fixed   gp/src/parameters.rs:17 */
#define EGX_THETA_FULL    1   /* ThetaTuning::Full    gp/src/parameters.rs:19 */
#define EGX_THETA_PARTIAL 2   /* ThetaTuning::Partial gp/src/parameters.rs:26 */

/* GpValidParams, gp/src/parameters.rs:90-120.  Use egx_gp_params_default() to
 * get the reference defaults (theta init 0.1, bounds [1e-2, 10], n_start 10,
 * max_eval 1000, nugget 100*eps, constant mean, squared exponential). */
#define EGX_OPT_COBYLA 0
#define EGX_OPT_LBFGSB 1
/* Exchange step of a multistart fit whose chains are sharded over several processes / GPUs: called ONCE, after the
 * local chains have finished, with the best objective (-likelihood) and its log10 theta (n values) of this process; must
 * overwrite both with those of the process holding the smallest objective (the `reduce` by min of
 * gp/src/algorithm.rs:942-945 across ranks: an NCCL / gloo all-gather in egobox_b200/parallel.py, egx_argmin_allreduce for
 * C callers) and return 0. */
typedef int (*egx_exchange_fn)(double* f_best, double* z_best, int n, void* user);

typedef struct egx_gp_params {
    int corr;                    /* EGX_CORR_*  */
    int mean;                    /* EGX_MEAN_*  */
    int theta_tuning;            /* EGX_THETA_* */
    const double* theta_init;    /* n_theta_init values: 1 (broadcast) or theta dimension */
    int n_theta_init;
    const double* theta_bounds;  /* n_theta_bounds (lo, hi) pairs: 1 (broadcast) or theta dimension */
    int n_theta_bounds;
    const int* active;           /* EGX_THETA_PARTIAL: optimised component indices */
    int n_active;
    int n_start;                 /* multistart count (n_start + 1 chains), algorithm.rs:33 */
    int max_eval;                /* per-chain budget = clamp(10*dim, 25, max_eval), algorithm.rs:936-937 */
    double nugget;
    const double* w_star;        /* optional d x kpls_dim PLS rotations supplied by the caller; NULL = computed */
    int kpls_dim;                /* KPLS components (algorithm.rs:798-813, 843-855): 0 = none (identity w_star);
                                    > 0 with w_star == NULL: rotations by egx_pls_rotations (linfa-pls NIPALS) */
    int device;                  /* CUDA device ordinal */
    unsigned long long seed;     /* multistart LHS seed (reference: 42, optimization.rs:60-63) */
    double cobyla_rhobeg;        /* 0.5   optimization.rs:19 */
    double cobyla_ftol_rel;      /* 1e-4  optimization.rs:20; <= 0 disables the early stop */
    int optimizer;               /* EGX_OPT_COBYLA (default, the reference's optimiser) or EGX_OPT_LBFGSB: projected L-BFGS
                                    per start on the closed-form theta gradient (egx_gp_reduced_likelihood_grad_analytic),
                                    same starts, same evaluation budget -- not in the reference, opt-in */
    int chain_rank;              /* multi-GPU multistart: this process runs the chains c with c % chain_world == chain_rank */
    int chain_world;             /* (0 or 1: all chains here) ... */
    egx_exchange_fn exchange;    /* ... and `exchange` makes the best (objective, theta) of all processes known to each */
    void* exchange_user;
} egx_gp_params;

typedef struct egx_gp_model egx_gp_model;

void egx_gp_params_default(egx_gp_params* p);

/* impl Fit for GpValidParams::fit, gp/src/algorithm.rs:791-979.
 * x: n x d raw inputs, y: n raw outputs.  The n_start+1 COBYLA chains are independent, as in the
 * reference (the rayon fan-out of :928-945): their evaluations overlap on the workspace slots of the
 * asynchronous seam (egx_gp_eval_begin / egx_gp_eval_end). */
int egx_gp_fit(const egx_gp_params* params, const double* x, int n, int d, const double* y,
               egx_gp_model** out);
void egx_gp_model_destroy(egx_gp_model* m);

/* accessors: theta() :413, variance() :418, likelihood() :423, dims() :437 */
int egx_gp_model_dims(const egx_gp_model* m, int* n, int* d, int* h, int* p);
int egx_gp_model_theta(const egx_gp_model* m, double* theta /* h */);
double egx_gp_model_variance(const egx_gp_model* m);
double egx_gp_model_likelihood(const egx_gp_model* m);
long long egx_gp_model_n_evals(const egx_gp_model* m);   /* likelihood evaluations spent by fit */
/* GpInnerParams (algorithm.rs:47-60) and normalisation data for serialisation; any pointer may be NULL */
int egx_gp_model_inner_params(egx_gp_model* m, double* beta, double* gamma, double* r_chol, double* ft,
                              double* ft_qr_r);
int egx_gp_model_normalization(const egx_gp_model* m, double* x_mean, double* x_std, double* y_mean,
                               double* y_std, double* w_star /* d x h */);
/* the device context of a fitted model (owned by the model) */
egx_gp_ctx* egx_gp_model_context(egx_gp_model* m);

/* Contexts draw device / pinned memory from a size-keyed cache (a fit creates and destroys ~12 workspaces and the
 * EGO loop refits every iteration -- the reference allocates its ndarray temporaries per call likewise, through the
 * system allocator).  Idle blocks are capped by EGX_CACHE_MB (default 4096 per kind); this returns them to the
 * driver. */
void egx_release_cached_memory(void);

/* Host-only utilities (no GPU needed).
 * egx_bound_cobyla_minimize: the derivative-free optimiser `optimize_params` runs per
 * chain (gp/src/optimization.rs:122-169): minimise f over the box [lo, hi] from x0 with
 * initial trust radius rhobeg, NLopt-style ftol_rel stop and an evaluation budget.
 * egx_prepare_multistart: gp/src/optimization.rs:26-71, (n_start+1) x dim log10 starts
 * (row 0 = log10 theta0, rows 1.. = maximin LHS in the log10 box, bounds = dim (lo,hi) pairs). */
typedef double (*egx_objective_fn)(const double* x, int n, void* user);
int egx_bound_cobyla_minimize(egx_objective_fn f, void* user, int n, const double* x0, const double* lo,
                              const double* hi, double rhobeg, double ftol_rel, int maxeval,
                              double* x_opt, double* f_opt, int* n_evals);
int egx_prepare_multistart(int n_start, const double* theta0, const double* bounds, int dim,
                           unsigned long long seed, double* starts_out);
/* The seeded random streams of the path, restated from rand 0.8.5 / rand_xoshiro 0.6.0 (Cargo.lock) so that a seed
 * gives the reference's own points:
 * egx_lhs_sample: `Lhs::new(xlimits).kind(..).with_rng(Xoshiro256Plus::seed_from_u64(seed)).sample(ns)`,
 *   doe/src/lhs.rs:67-88, 235-258 (kind 0 = Classic), 283-304 (kind 1 = Maximin, what prepare_multistart uses);
 *   xlimits = nx (lower, upper) pairs, out = ns x nx row-major.  Pinned on the fixture of doe/src/lhs.rs:332-347.
 * egx_shuffled_indices: `(0..n).collect::<Vec<_>>().shuffle(&mut Xoshiro256Plus::seed_from_u64(seed))`, the index
 *   stream of make_inducings (gp/src/sparse_algorithm.rs:833-847): the inducing points are rows out[0..nz) of xt. */
int egx_lhs_sample(int kind, int ns, int nx, const double* xlimits, unsigned long long seed, double* out);
int egx_shuffled_indices(int n, unsigned long long seed, int* out);
/* egx_bound_lbfgs_minimize: the per-start optimiser of EGX_OPT_LBFGSB -- projected limited-memory BFGS (8 pairs) with
 * Armijo backtracking over the box [lo, hi]; fg returns f and writes the gradient.  Stops on
 * f_prev - f <= ftol_rel * max(|f_prev|, |f|, 1), on |projected gradient|_inf <= gtol * max(1, |f|), or on the budget. */
typedef double (*egx_objective_grad_fn)(const double* x, int n, double* grad, void* user);
int egx_bound_lbfgs_minimize(egx_objective_grad_fn fg, void* user, int n, const double* x0, const double* lo,
                             const double* hi, double ftol_rel, double gtol, int maxeval,
                             double* x_opt, double* f_opt, int* n_evals);
/* egx_pls_rotations: `PlsRegression::params(k).fit(&ds)?.rotations().0` of gp/src/algorithm.rs:843-855 and
 * gp/src/sparse_algorithm.rs:442-455 (linfa-pls 0.8.0 = scikit-learn's NIPALS PLSRegression, scale = true):
 * x n x d raw, y n raw -> w_star d x k.  A numerically constant y residual yields zeros, like the reference. */
int egx_pls_rotations(const double* x, int n, int d, const double* y, int k, double* w_star);
/* ---- the exchange of the sharded path (multi-process, one rank per GPU) -----------------------------------------------
 * Multistart chains, theta candidates and experts shard across ranks with NO data-path collective; at the end the ranks
 * exchange (value, payload[h]) -- 8 (h + 2) bytes each -- and keep the best: the reduction of gp/src/algorithm.rs:942-945
 * across processes.  A TCP star (rank 0 listens on addr:port, the MASTER_ADDR convention of torchrun; use a port of your
 * own, torchrun's MASTER_PORT is taken by its store) so that a caller needs no NCCL binding for a message of a few dozen
 * bytes; Python callers with a torch process group use egobox_b200/parallel.py (NCCL all_gather) instead.
 *   egx_comm_init       blocks until all nranks ranks have joined (timeout_ms <= 0: 60 s); nranks == 1 needs no address
 *   egx_comm_allgather  every rank contributes `count` doubles and receives nranks * count doubles in rank order
 *   egx_argmin_allreduce  in place: the pair of the rank with the smallest value (ties: lowest rank, NaN never wins) */
typedef struct egx_comm egx_comm;
int egx_comm_init(egx_comm** out, int nranks, int rank, const char* addr, int port, int timeout_ms);
void egx_comm_destroy(egx_comm* comm);
int egx_comm_rank(const egx_comm* comm);
int egx_comm_size(const egx_comm* comm);
int egx_comm_allgather(egx_comm* comm, const double* send, int count, double* recv);
int egx_argmin_allreduce(egx_comm* comm, double* value, double* payload, int h, int* winner_rank);

/* egx_symmetric_eig: eigen-decomposition of a symmetric n x n matrix (row-major, overwritten by the
 * eigenvectors as COLUMNS; w = eigenvalues, unsorted) -- the host half of the eigenvalue sampler,
 * `cov_x.eigh()` gp/src/algorithm.rs:1171-1173.  Returns EGX_OK or EGX_INVALID_VALUE (no convergence). */
int egx_symmetric_eig(int n, double* a, double* w);

/* predict :253, predict_var :267, predict_valvar :282 (raw x, m x d) */
int egx_gp_model_predict(egx_gp_model* m, const double* x, int npts, double* y);
int egx_gp_model_predict_var(egx_gp_model* m, const double* x, int npts, double* var);
int egx_gp_model_predict_valvar(egx_gp_model* m, const double* x, int npts, double* y, double* var);
int egx_gp_model_predict_gradients(egx_gp_model* m, const double* x, int npts, double* grad /* npts x d */);
int egx_gp_model_predict_var_gradients(egx_gp_model* m, const double* x, int npts, double* grad /* npts x d */);
/* GaussianProcess::sample_chol / sample_eig (gp/src/algorithm.rs:383-395) and the covariance they factor (:310-326) */
int egx_gp_model_covariance(egx_gp_model* m, const double* x, int npts, double* cov /* npts x npts */);
int egx_gp_model_sample(egx_gp_model* m, const double* x, int npts, const double* z /* npts x n_traj */, int n_traj,
                        int method, double* out /* npts x n_traj */);

/* ============================================================================
 * Sparse GP (FITC / VFE) -- crates/gp/src/sparse_algorithm.rs.
 * No input normalisation, zero mean (as in the reference).  x: N x d raw, y: N,
 * z: M x d inducing points (Inducings::Located, or drawn by egx_sgp_fit for
 * Inducings::Randomized :833-847), w_star d x h (identity when h == d).
 * ========================================================================== */
#define EGX_SGP_FITC 0   /* SparseMethod::Fitc, sparse_parameters.rs:56 */
#define EGX_SGP_VFE  1   /* SparseMethod::Vfe */
typedef struct egx_sgp_ctx egx_sgp_ctx;
int egx_sgp_create(egx_sgp_ctx** out, int device, int corr, int method, const double* x, int n, int d,
                   const double* y, const double* z, int m, const double* w_star, int h, double nugget);
void egx_sgp_destroy(egx_sgp_ctx* ctx);
/* SgpValidParams::reduced_likelihood (sparse_algorithm.rs:654-673 -> fitc :695-765 / vfe :769-830),
 * value only.  A failed Cholesky (where the reference `.unwrap()`s) is status EGX_NOT_POSITIVE_DEFINITE. */
int egx_sgp_reduced_likelihood(egx_sgp_ctx* ctx, const double* theta, double sigma2, double noise, double* lik);
/* Final evaluation: keeps U = chol(Kmm), L = chol(A) and the Woodbury vector on the device;
 * w_vec (M) and w_inv (M x M, WoodburyData.inv -- computed only when non-NULL) may be NULL. */
int egx_sgp_finalize(egx_sgp_ctx* ctx, const double* theta, double sigma2, double noise, double* lik,
                     double* w_vec, double* w_inv);
/* predict :237-241, predict_var :245-257 (k^T inv k evaluated by two triangular sweeps, inv never formed) */
int egx_sgp_predict(egx_sgp_ctx* ctx, const double* x, int m, double* y);
int egx_sgp_predict_var(egx_sgp_ctx* ctx, const double* x, int m, double* var);
/* sample_chol / sample_eig / sample, sparse_algorithm.rs:338-364: out (m x n_traj) = predict(x) + C z with C C^T = sigma2 r(x, x)
 * (the reference's `_sample` takes the PRIOR covariance `compute_k(x, x, ..)`); z: m x n_traj standard normal draws of the caller,
 * method: EGX_SAMPLE_CHOLESKY | EGX_SAMPLE_EIGENVALUES (decomposition as in gp/src/algorithm.rs:1153-1194); m <= 8192 */
int egx_sgp_sample(egx_sgp_ctx* ctx, const double* x, int m, const double* z, int n_traj, int method, double* out);
int egx_sgp_set_profiling(egx_sgp_ctx* ctx, int enabled);
int egx_sgp_get_profile(egx_sgp_ctx* ctx, double* ms, long long* launches);

/* SgpValidParams (sparse_parameters.rs:71-82, 151-293) + impl Fit (sparse_algorithm.rs:416-648). */
typedef struct egx_sgp_params {
    int corr;
    int method;                  /* EGX_SGP_FITC | EGX_SGP_VFE */
    int theta_fixed;             /* ThetaTuning::Fixed -> bounds collapse onto init (:470) */
    const double* theta_init;    /* 1 or theta-dimension values (default 0.1) */
    int n_theta_init;
    const double* theta_bounds;  /* 1 (broadcast) or n_params (lo, hi) pairs; default (1e-2, 1e2) */
    int n_theta_bounds;
    int noise_fixed;             /* ParamTuning::Fixed(noise_init) vs Optimized{init, bounds} */
    double noise_init;           /* default 1e-2 */
    double noise_lo, noise_hi;   /* default (100 eps, 1e10) */
    const double* z;             /* Inducings::Located (n_inducings x d) or NULL */
    int n_inducings;             /* Inducings::Randomized(n) when z == NULL (default 10) */
    int n_start, max_eval;
    double nugget;
    const double* w_star;        /* as in egx_gp_params: NULL + kpls_dim > 0 = PLS rotations computed here */
    int kpls_dim;
    int device;
    unsigned long long seed;
    double cobyla_rhobeg, cobyla_ftol_rel;
} egx_sgp_params;
typedef struct egx_sgp_model egx_sgp_model;
void egx_sgp_params_default(egx_sgp_params* p);
int egx_sgp_fit(const egx_sgp_params* params, const double* x, int n, int d, const double* y, egx_sgp_model** out);
void egx_sgp_model_destroy(egx_sgp_model* m);
int egx_sgp_model_dims(const egx_sgp_model* m, int* n, int* d, int* h, int* n_inducings);
int egx_sgp_model_theta(const egx_sgp_model* m, double* theta);
double egx_sgp_model_variance(const egx_sgp_model* m);        /* sigma2        :262 */
double egx_sgp_model_noise_variance(const egx_sgp_model* m);  /* noise         :267 */
double egx_sgp_model_likelihood(const egx_sgp_model* m);
long long egx_sgp_model_n_evals(const egx_sgp_model* m);
int egx_sgp_model_inducings(const egx_sgp_model* m, double* z /* n_inducings x d */);
int egx_sgp_model_woodbury(egx_sgp_model* m, double* w_vec, double* w_inv);
egx_sgp_ctx* egx_sgp_model_context(egx_sgp_model* m);
int egx_sgp_model_predict(egx_sgp_model* m, const double* x, int npts, double* y);
int egx_sgp_model_predict_var(egx_sgp_model* m, const double* x, int npts, double* var);
int egx_sgp_model_sample(egx_sgp_model* m, const double* x, int npts, const double* z, int n_traj, int method, double* out);

/* ============================================================================
 * Mixture of experts -- crates/moe/src/gaussian_mixture.rs + the recombination of
 * crates/moe/src/algorithm.rs.
 * An egx_moe holds the predict-side Gaussian mixture `gmx` (weights k, means k x d,
 * covariances k x d x d, heaviside factor: GaussianMixture::new :62-83 +
 * heaviside_factor :101-106) and BORROWS one fitted expert context per cluster
 * (egx_gp_model_context of the expert trained on the rows of that cluster,
 * algorithm.rs:165-177).  The points stay on the device between the
 * responsibilities, the experts' batched predictions and the recombination.
 * ========================================================================== */
#define EGX_RECOMB_HARD   0   /* Recombination::Hard      moe/src/types.rs */
#define EGX_RECOMB_SMOOTH 1   /* Recombination::Smooth(f) (f = the mixture's heaviside factor) */
typedef struct egx_moe egx_moe;
int egx_moe_create(egx_moe** out, int device, int k, int d, const double* weights, const double* means,
                   const double* covariances, double heaviside_factor);
void egx_moe_destroy(egx_moe* moe);
int egx_moe_set_heaviside_factor(egx_moe* moe, double factor);           /* gaussian_mixture.rs:101-106 */
int egx_moe_set_expert(egx_moe* moe, int cluster, egx_gp_ctx* expert);   /* borrowed, must outlive the calls */
/* precisions / precisions_chol (k x d x d) and log_det (k) as serialised in the `gmx` block; any may be NULL */
int egx_moe_parameters(const egx_moe* moe, double* precisions, double* precisions_chol, double* log_det);
/* predict_probas :109-116 (m x k) and predict :306-318 (arg-max cluster per point); either output may be NULL */
int egx_moe_predict_probas(egx_moe* moe, const double* x, int m, double* probas, int* clusters);
/* predict_probas_derivatives :158-170 (m x k x d) */
int egx_moe_predict_probas_derivatives(egx_moe* moe, const double* x, int m, double* dprobas);
/* GpMixture::predict / predict_var / predict_valvar / predict_gradients / predict_var_gradients /
 * predict_valvar_gradients (algorithm.rs:455-541 dispatching to :411-423, 670-1010): any of y (m),
 * var (m), grad_y (m x d), grad_var (m x d) may be NULL. */
int egx_moe_predict(egx_moe* moe, int recombination, const double* x, int m, double* y, double* var,
                    double* grad_y, double* grad_var);

#ifdef __cplusplus
}
#endif
#endif /* EGOBOX_GPU_H */
