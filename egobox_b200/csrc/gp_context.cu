// Device-resident GP context and the C ABI declared in include/egobox_gpu.h.
//
// One context = one training set on one GPU.  A likelihood evaluation is
//   K1 corr_build -> [F|y]^T rows appended under R -> blocked Cholesky (the appended rows
//   come out as (L^-1 [F|y])^T, i.e. the forward solves of algorithm.rs:1006,1028 are fused
//   into the factorisation) -> GLS kernel -> one small D2H of (rlf, sigma2, info, G, beta).
// Nothing here falls back to the CPU: the only host arithmetic is the O(p^3) condition
// number test on the p x p factor G (algorithm.rs:1010-1027).
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "sweep.cuh"
#include "../../include/egobox_gpu.h"
#include "abi_guard.h"

// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void egx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* egx_last_error(void) { return g_err; }
extern "C" const char* egx_version(void) { return "egobox_b200 0.1 (sm_100a)"; }
extern "C" int egx_device_count(void) try {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return c;
}
EGX_ABI_CATCH_COUNT

namespace {

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// One-sided Jacobi (Hestenes) singular values of a (rows x cols) row-major matrix, rows >= cols.
std::vector<double> singular_values(const double* a, int rows, int cols) {
    std::vector<double> u(static_cast<size_t>(rows) * cols);
    // column-major working copy
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) u[static_cast<size_t>(c) * rows + r] = a[static_cast<size_t>(r) * cols + c];
    const double eps = 2.220446049250313e-16;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int i = 0; i < cols - 1; ++i) {
            double* ui = &u[static_cast<size_t>(i) * rows];
            for (int j = i + 1; j < cols; ++j) {
                double* uj = &u[static_cast<size_t>(j) * rows];
                double al = 0.0, be = 0.0, ga = 0.0;
                for (int k = 0; k < rows; ++k) {
                    al += ui[k] * ui[k];
                    be += uj[k] * uj[k];
                    ga += ui[k] * uj[k];
                }
                if (ga == 0.0 || std::fabs(ga) <= eps * std::sqrt(al * be)) continue;
                rotated = true;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < rows; ++k) {
                    const double x = ui[k], y = uj[k];
                    ui[k] = c * x - s * y;
                    uj[k] = s * x + c * y;
                }
            }
        }
        if (!rotated) break;
    }
    std::vector<double> sv(cols);
    for (int c = 0; c < cols; ++c) {
        double s = 0.0;
        for (int k = 0; k < rows; ++k) s += u[static_cast<size_t>(c) * rows + k] * u[static_cast<size_t>(c) * rows + k];
        sv[c] = std::sqrt(s);
    }
    std::sort(sv.begin(), sv.end(), [](double x, double y) { return x > y; });
    return sv;
}

}  // namespace

struct egx_gp_ctx {
    int device = 0, corr = 0, mean = 0;
    int n = 0, d = 0, h = 0, p = 0, q = 0;
    int npad = 0, qpad = 0, rows_total = 0;
    long ld = 0;
    double nugget = 0.0, y_mean = 0.0, y_std = 1.0;
    std::vector<double> w_star, xnorm_h, ynorm_h, x_mean_h, x_std_h;
    // replicas of this context (own R/L workspace, streams, events) used to keep several independent
    // likelihood evaluations of a batch in flight: one evaluation's serial panel chain overlaps the
    // bulk trailing updates of another
    std::vector<egx_gp_ctx*> replicas;
    bool pending_eval = false;
    std::vector<int> basis_i_h, basis_j_h;

    SweepEnv env;                         // streams, look-ahead events, panel buffers, profiler
    cudaStream_t stream = nullptr;        // alias of env.sb
    double *X = nullptr, *ynorm = nullptr, *x_mean = nullptr, *x_std = nullptr, *FyT = nullptr;
    int *basis_i = nullptr, *basis_j = nullptr;
    CorrTerm* terms = nullptr;
    CorrTerm* terms_h = nullptr;   // pinned
    int max_terms = 0, nterms = 0;

    double* Dinv = nullptr;   // [npad/128][4][32][32] inverted diagonal sub-blocks of L
    double *M = nullptr, *glswork = nullptr, *G = nullptr, *beta = nullptr, *rho = nullptr;
    EvalResult* res = nullptr;
    int* info = nullptr;
    // pinned host mirrors
    EvalResult* res_h = nullptr;
    double *G_h = nullptr, *beta_h = nullptr;

    // small-n batched path (K8)
    double *W_dev = nullptr, *sb_thetas = nullptr, *sb_G = nullptr;
    void* sb_out = nullptr;
    SmallOutHost* sb_out_h = nullptr;   // pinned
    double *sb_G_h = nullptr, *sb_thetas_h = nullptr;
    int sb_cap = 0;

    // trained state (after finalize)
    bool trained = false;
    std::vector<double> theta;
    double sigma2_scaled = 0.0;

    // variance-gradient state (built lazily after a finalize): transposed factor and K_F = R^-1 F
    double *LT = nullptr, *KF = nullptr;
    bool grad_ready = false;

    // predict buffers
    double *Y = nullptr, *xchunk = nullptr, *ychunk = nullptr, *vchunk = nullptr;
    int mb_alloc = 0;

    cudaEvent_t timer_a = nullptr, timer_b = nullptr;
    bool force_blocked = false;

    // One likelihood evaluation captured as a CUDA graph (both streams of the look-ahead schedule): a blocked
    // evaluation is 5-7 launches per block column, host launch-bound below n ~ 4096.  Re-captured when the
    // number of kernel terms, the look-ahead setting or a captured buffer changes; bypassed while profiling.
    cudaGraphExec_t eval_graph = nullptr;
    int graph_nterms = -1, graph_generation = -1, graph_persist = -1;
    bool graph_lookahead = false, use_graphs = true, use_graphs_lookahead = true;
    long long graph_launches[EGX_NUM_STAGES] = {0};
    int async_slots = 0;         // workspaces handed out by egx_gp_async_slots
    // int8 digit slices of L for the multi-RHS solve of predict_var on tcgen05 (built lazily after a finalize)
    int8_t* Lsl = nullptr;
    double* Lsc = nullptr;
    std::vector<long> Lsl_off, Lsc_off;
    bool Lslices_ready = false;
    int Lsl_pairs_needed = 0;    // pairs a solve reads slices of (trailing rows >= ozaki_min_tri_solve tile rows)
    int Lsl_pairs_written = 0;   // leading pairs the last factorisation wrote itself (look-ahead schedule, FactorRef::Lsl_w)
    int graph_Lsl_pairs = 0;     // the same for the captured evaluation
    long long direct_evals = 0;
    // closed-form theta gradient (built lazily): W = L^-T, -R^-1, per-CTA partial sums, term list
    double *tgW = nullptr, *tgRinv = nullptr, *tgPartial = nullptr, *tgGrad = nullptr, *tgGrad_h = nullptr, *tgGamma = nullptr;
    ThetaGradTerm *tgTerms = nullptr, *tgTerms_h = nullptr;
    std::mutex mu;
};

namespace {

void resolve_profile(egx_gp_ctx* c) { c->env.prof.resolve(); }

// host half of the per-theta setup: kernel weights into the pinned staging buffer
int build_terms_host(egx_gp_ctx* c, const double* theta) {
    for (int l = 0; l < c->h; ++l)
        if (std::isnan(theta[l])) {
            egx_set_error("theta[%d] is NaN", l);
            return EGX_INVALID_VALUE;
        }
    c->nterms = egx_fill_terms(c->corr, c->d, c->h, c->w_star.data(), theta, c->terms_h);
    return EGX_OK;
}

int upload_terms(egx_gp_ctx* c) {
    if (c->nterms > 0)
        EGX_CUDA_TRY(cudaMemcpyAsync(c->terms, c->terms_h, c->nterms * sizeof(CorrTerm), cudaMemcpyHostToDevice,
                                     c->stream));
    return EGX_OK;
}

// R(theta) lower block-triangle into M, then the RHS rows (terms already in the pinned staging buffer).
int assemble_staged(egx_gp_ctx* c) {
    int st = upload_terms(c);
    if (st != EGX_OK) return st;
    EGX_CUDA_TRY(cudaMemsetAsync(c->info, 0, sizeof(int), c->stream));
    {
        StageScope sc(c->env.prof, EGX_STAGE_CORR_BUILD, 1, c->stream);
        launch_corr_build(c->corr, c->X, c->n, c->npad, c->d, c->terms, c->nterms, c->M, c->ld, 1.0 + c->nugget,
                          c->stream);
    }
    EGX_CUDA_TRY(cudaMemcpyAsync(c->M + static_cast<long>(c->npad) * c->ld, c->FyT,
                                 static_cast<size_t>(c->q) * c->ld * sizeof(double), cudaMemcpyDeviceToDevice,
                                 c->stream));
    return EGX_OK;
}

int assemble(egx_gp_ctx* c, const double* theta) {
    int st = build_terms_host(c, theta);
    if (st != EGX_OK) return st;
    return assemble_staged(c);
}

FactorRef factor_ref(egx_gp_ctx* c) {
    FactorRef f;
    f.M = c->M;
    f.ld = c->ld;
    f.T = c->npad / EGX_NB;
    f.qpad = c->qpad;
    f.Dinv = c->Dinv;
    f.info = c->info;
    return f;
}

// Once a model has been asked for a solve on tcgen05 (predict_var, the closed-form gradient) its slice storage exists, and every
// later factorisation under the look-ahead schedule writes the slices of its panel rows THERE instead of the rotating buffers:
// they are the slices of the block rows of L the solves need (1 ms to rebuild at n = 8192 otherwise).
void cholesky(egx_gp_ctx* c) {
    FactorRef f = factor_ref(c);
    c->Lsl_pairs_written = 0;
    if (c->Lsl != nullptr && c->Lsc != nullptr) {
        f.Lsl_w = c->Lsl;
        f.Lsc_w = c->Lsc;
        f.Lsl_off_w = c->Lsl_off.data();
        f.Lsc_off_w = c->Lsc_off.data();
        f.Lsl_w_pairs = &c->Lsl_pairs_written;
    }
    blocked_sweep(c->env, f, true, nullptr, 0, 0, 0);
}

// Condition-number test of gp/src/algorithm.rs:1010-1027 on the p x p factor G (host, O(p^3)).
int cond_status(egx_gp_ctx* c, const double* G) {
    std::vector<double> sv = singular_values(G, c->p, c->p);
    const double cond_ft = sv.back() / sv.front();
    if (!(cond_ft >= 1e-10)) {
        // fx = mean.value(xnorm) rebuilt on the host for the (rare) diagnostic branch
        std::vector<double> F(static_cast<size_t>(c->n) * c->p);
        for (int i = 0; i < c->n; ++i)
            for (int l = 0; l < c->p; ++l) {
                const int bi = c->basis_i_h[l], bj = c->basis_j_h[l];
                const double* x = &c->xnorm_h[static_cast<size_t>(i) * c->d];
                F[static_cast<size_t>(i) * c->p + l] = (bi < 0 ? 1.0 : x[bi]) * (bj < 0 ? 1.0 : x[bj]);
            }
        std::vector<double> svf = singular_values(F.data(), c->n, c->p);
        const double cond_fx = svf.front() / svf.back();
        if (cond_fx > 1e15) {
            egx_set_error("F is too ill conditioned. Poor combination of regression model and observations.");
            return EGX_ILL_CONDITIONED_F;
        }
        egx_set_error("ft is too ill conditioned, try another theta again");
        return EGX_ILL_CONDITIONED_FT;
    }
    return EGX_OK;
}

bool small_path_ok(const egx_gp_ctx* c) {
    return !c->force_blocked && small_batch_smem_bytes(c->n, c->d, c->h, c->p) <= 225 * 1024;
}

// B candidate thetas, one CTA each (K8).  thetas: host, B x h.
int evaluate_small_batch(egx_gp_ctx* c, const double* thetas, int B, double* rlf, int* status) {
    if (B <= 0) return EGX_OK;
    c->trained = false;
    if (B > c->sb_cap) {
        egx_dev_free(c->sb_thetas);
        egx_dev_free(c->sb_G);
        egx_dev_free(c->sb_out);
        egx_host_free(c->sb_out_h);
        egx_host_free(c->sb_G_h);
        egx_host_free(c->sb_thetas_h);
        c->sb_cap = 0;
        const int cap = std::max(B, 64);
        EGX_CUDA_TRY(egx_dev_malloc(&c->sb_thetas, static_cast<size_t>(cap) * c->h * sizeof(double)));
        EGX_CUDA_TRY(egx_dev_malloc(&c->sb_G, static_cast<size_t>(cap) * c->p * c->p * sizeof(double)));
        EGX_CUDA_TRY(egx_dev_malloc(&c->sb_out, static_cast<size_t>(cap) * sizeof(SmallOutHost)));
        EGX_CUDA_TRY(egx_host_malloc(&c->sb_out_h, static_cast<size_t>(cap) * sizeof(SmallOutHost)));
        EGX_CUDA_TRY(egx_host_malloc(&c->sb_G_h, static_cast<size_t>(cap) * c->p * c->p * sizeof(double)));
        EGX_CUDA_TRY(egx_host_malloc(&c->sb_thetas_h, static_cast<size_t>(cap) * c->h * sizeof(double)));
        c->sb_cap = cap;
    }
    std::vector<char> bad(B, 0);
    for (int b = 0; b < B; ++b)
        for (int l = 0; l < c->h; ++l) {
            const double v = thetas[static_cast<long>(b) * c->h + l];
            if (std::isnan(v)) bad[b] = 1;
            c->sb_thetas_h[static_cast<long>(b) * c->h + l] = std::isnan(v) ? 1.0 : v;
        }
    EGX_CUDA_TRY(cudaMemcpyAsync(c->sb_thetas, c->sb_thetas_h, static_cast<size_t>(B) * c->h * sizeof(double),
                                 cudaMemcpyHostToDevice, c->stream));
    {
        StageScope sc(c->env.prof, EGX_STAGE_SMALL_BATCH, 1, c->stream);
        launch_small_batch(c->corr, c->X, c->n, c->d, c->W_dev, c->h, c->sb_thetas, B, c->FyT, c->ld, c->p,
                           1.0 + c->nugget, c->sb_out, c->sb_G, c->stream);
    }
    EGX_CUDA_TRY(cudaMemcpyAsync(c->sb_out_h, c->sb_out, static_cast<size_t>(B) * sizeof(SmallOutHost),
                                 cudaMemcpyDeviceToHost, c->stream));
    EGX_CUDA_TRY(cudaMemcpyAsync(c->sb_G_h, c->sb_G, static_cast<size_t>(B) * c->p * c->p * sizeof(double),
                                 cudaMemcpyDeviceToHost, c->stream));
    EGX_CUDA_TRY(cudaStreamSynchronize(c->stream));
    EGX_CUDA_TRY(cudaGetLastError());
    resolve_profile(c);
    for (int b = 0; b < B; ++b) {
        rlf[b] = NAN;
        if (bad[b]) {
            egx_set_error("theta is NaN");
            status[b] = EGX_INVALID_VALUE;
            continue;
        }
        if (c->sb_out_h[b].info != 0) {
            egx_set_error("correlation matrix is not positive definite (pivot %d)", c->sb_out_h[b].info);
            status[b] = EGX_NOT_POSITIVE_DEFINITE;
            continue;
        }
        status[b] = cond_status(c, c->sb_G_h + static_cast<long>(b) * c->p * c->p);
        if (status[b] == EGX_OK) rlf[b] = c->sb_out_h[b].rlf;
    }
    return EGX_OK;
}

// Full likelihood evaluation; leaves L, (L^-1[F|y])^T, beta, G, rho on the device.
// evaluate_launch enqueues everything (no host synchronisation); evaluate_collect waits for the
// result block and applies the host-side status logic.
int enqueue_eval(egx_gp_ctx* c) {
    int st = assemble_staged(c);
    if (st != EGX_OK) return st;
    cholesky(c);
    {
        StageScope sc(c->env.prof, EGX_STAGE_GLS, 1, c->stream);
        launch_gls(c->M, c->ld, c->n, c->npad, c->p, c->glswork, c->G, c->beta, c->rho, c->res, c->info, c->stream);
    }
    EGX_CUDA_TRY(cudaMemcpyAsync(c->res_h, c->res, sizeof(EvalResult), cudaMemcpyDeviceToHost, c->stream));
    EGX_CUDA_TRY(cudaMemcpyAsync(c->G_h, c->G, sizeof(double) * c->p * c->p, cudaMemcpyDeviceToHost, c->stream));
    EGX_CUDA_TRY(cudaMemcpyAsync(c->beta_h, c->beta, sizeof(double) * c->p, cudaMemcpyDeviceToHost, c->stream));
    return EGX_OK;
}

void drop_graph(egx_gp_ctx* c) {
    if (c->eval_graph) cudaGraphExecDestroy(c->eval_graph);
    c->eval_graph = nullptr;
}

// Capture one evaluation (terms upload ... result download) on the context's streams.
int capture_eval_graph(egx_gp_ctx* c) {
    drop_graph(c);
    long long before[EGX_NUM_STAGES];
    for (int i = 0; i < EGX_NUM_STAGES; ++i) before[i] = c->env.prof.launches[i];
    EGX_CUDA_TRY(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const int st = enqueue_eval(c);
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    for (int i = 0; i < EGX_NUM_STAGES; ++i) {
        c->graph_launches[i] = c->env.prof.launches[i] - before[i];
        c->env.prof.launches[i] = before[i];
    }
    if (st != EGX_OK || e != cudaSuccess || g == nullptr) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        egx_set_error("CUDA graph capture of the likelihood evaluation failed: %s", cudaGetErrorString(e));
        return EGX_CUDA_ERROR;
    }
    const cudaError_t ei = cudaGraphInstantiate(&c->eval_graph, g, 0);
    cudaGraphDestroy(g);
    if (ei != cudaSuccess) {
        c->eval_graph = nullptr;
        egx_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
        return EGX_CUDA_ERROR;
    }
    c->graph_Lsl_pairs = c->Lsl_pairs_written;
    c->graph_nterms = c->nterms;
    c->graph_lookahead = c->env.lookahead;
    c->graph_generation = c->env.generation;
    c->graph_persist = c->env.oz_persist;
    return EGX_OK;
}

int evaluate_launch(egx_gp_ctx* c, const double* theta) {
    c->trained = false;
    c->grad_ready = false;
    c->Lslices_ready = false;
    c->pending_eval = false;
    int st = build_terms_host(c, theta);
    if (st != EGX_OK) return st;
    // the first evaluation of a context always launches directly (lazy module loading, function attributes)
    if ((c->use_graphs || (c->use_graphs_lookahead && c->env.lookahead)) && !c->env.prof.on && c->direct_evals > 0) {
        if (c->eval_graph == nullptr || c->graph_nterms != c->nterms || c->graph_lookahead != c->env.lookahead ||
            c->graph_generation != c->env.generation || c->graph_persist != c->env.oz_persist) {
            st = capture_eval_graph(c);
            if (st != EGX_OK) return st;
        }
        for (int i = 0; i < EGX_NUM_STAGES; ++i) c->env.prof.launches[i] += c->graph_launches[i];
        c->Lsl_pairs_written = c->graph_Lsl_pairs;
        EGX_CUDA_TRY(cudaGraphLaunch(c->eval_graph, c->stream));
    } else {
        st = enqueue_eval(c);
        if (st != EGX_OK) return st;
        ++c->direct_evals;
    }
    c->pending_eval = true;
    return EGX_OK;
}

int evaluate_collect(egx_gp_ctx* c, double* rlf_out) {
    *rlf_out = NAN;
    c->pending_eval = false;
    EGX_CUDA_TRY(cudaStreamSynchronize(c->stream));
    EGX_CUDA_TRY(cudaGetLastError());
    resolve_profile(c);
    if (c->res_h->info != 0) {
        egx_set_error("correlation matrix is not positive definite (pivot %d)", c->res_h->info);
        return EGX_NOT_POSITIVE_DEFINITE;
    }
    {
        const int cst = cond_status(c, c->G_h);
        if (cst != EGX_OK) return cst;
    }
    *rlf_out = c->res_h->rlf;
    return EGX_OK;
}

int evaluate(egx_gp_ctx* c, const double* theta, double* rlf_out) {
    *rlf_out = NAN;
    c->env.oz_persist = 0;
    int st = evaluate_launch(c, theta);
    if (st != EGX_OK) return st;
    return evaluate_collect(c, rlf_out);
}

int ensure_predict_buffers(egx_gp_ctx* c, int mb) {
    if (mb <= c->mb_alloc) return EGX_OK;
    egx_dev_free(c->Y);
    egx_dev_free(c->xchunk);
    egx_dev_free(c->ychunk);
    egx_dev_free(c->vchunk);
    c->Y = c->xchunk = c->ychunk = c->vchunk = nullptr;
    c->mb_alloc = 0;
    EGX_CUDA_TRY(egx_dev_malloc(&c->Y, static_cast<size_t>(mb) * c->npad * sizeof(double)));
    EGX_CUDA_TRY(egx_dev_malloc(&c->xchunk, static_cast<size_t>(mb) * c->d * sizeof(double)));
    EGX_CUDA_TRY(egx_dev_malloc(&c->ychunk, static_cast<size_t>(mb) * sizeof(double)));
    EGX_CUDA_TRY(egx_dev_malloc(&c->vchunk, static_cast<size_t>(mb) * sizeof(double)));
    if (c->env.ensure_panel_rows(mb) != EGX_OK) return EGX_CUDA_ERROR;
    c->mb_alloc = mb;
    return EGX_OK;
}

constexpr int PREDICT_CHUNK = 8192;

// Points per chunk of predict / predict_var: whole waves of the 64-row solve slabs (K5) -- two waves (18 944 points on 148 SMs)
// while the chunk x npad buffer stays under 2 GB, else one wave, else 8192.  Measured on the sparse GP, whose chunks are the
// same sweeps (profiles/r02/y7_sgp.txt): 8192 -> 18 944 points per chunk = 8.56 -> 7.05 ms.  EGX_PREDICT_CHUNK overrides.
int predict_chunk_points(const egx_gp_ctx* c) {
    static const int env_pts = getenv("EGX_PREDICT_CHUNK") != nullptr ? std::max(EGX_NB, atoi(getenv("EGX_PREDICT_CHUNK")) / EGX_NB * EGX_NB) : 0;
    if (env_pts > 0) return env_pts;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device) != cudaSuccess || sms <= 0) return PREDICT_CHUNK;
    const long budget = 2L << 30;
    for (int waves = 2; waves >= 1; --waves) {
        const long pts = static_cast<long>(waves) * sms * 64 / EGX_NB * EGX_NB;
        if (pts >= PREDICT_CHUNK && pts * c->npad * static_cast<long>(sizeof(double)) <= budget) return static_cast<int>(pts);
    }
    return PREDICT_CHUNK;
}

// Slices of the block rows of L below every column pair (the B operand of the solve updates), once per trained model:
// as many bytes as the lower triangle of L itself (268 MB at n = 8192), ~1 ms to build.
int ensure_L_slices(egx_gp_ctx* c) {
    if (c->Lslices_ready) return EGX_OK;
    const int T = c->npad / EGX_NB;
    if (!c->env.ozaki || T < c->env.ozaki_min_T || T - 2 < c->env.ozaki_min_tri_solve) return EGX_OK;
    if (c->Lsl == nullptr) {
        // room per pair: its T - k - 2 block rows of L and the appended rows (a factorisation that writes the slices itself
        // slices the whole panel, appended rows included; the solves read the first T - k - 2 row blocks)
        long bytes = 0, rows = 0;
        c->Lsl_off.assign((T + 1) / 2, 0);
        c->Lsc_off.assign((T + 1) / 2, 0);
        for (int k = 0; k + 2 < T; k += 2) {
            c->Lsl_off[k >> 1] = bytes;
            c->Lsc_off[k >> 1] = rows;
            const long rows_k = static_cast<long>(T - k - 2) * EGX_NB + c->qpad;
            bytes += static_cast<long>(ozaki_slice_bytes(rows_k));
            rows += rows_k;
        }
        EGX_CUDA_TRY(egx_dev_malloc(&c->Lsl, static_cast<size_t>(bytes)));
        EGX_CUDA_TRY(egx_dev_malloc(&c->Lsc, static_cast<size_t>(rows) * sizeof(double)));
        c->Lsl_pairs_written = 0;
        ++c->env.generation;          // a captured evaluation still writes to the rotating buffers: capture again
    }
    for (int k = 0; k + 2 < T; k += 2) {
        const int rows_k = (T - k - 2) * EGX_NB;
        if (rows_k / EGX_NB < c->env.ozaki_min_tri_solve) break;
        if ((k >> 1) < c->Lsl_pairs_written) continue;          // the factorisation left them there (cholesky())
        StageScope sc(c->env.prof, EGX_STAGE_OZAKI_SLICE, 2, c->stream);
        launch_ozaki_slice(c->M + static_cast<long>(k + 2) * EGX_NB * c->ld + static_cast<long>(k) * EGX_NB, c->ld, rows_k,
                           c->Lsc + c->Lsc_off[k >> 1], c->Lsl + c->Lsl_off[k >> 1], c->stream);
    }
    c->Lslices_ready = true;
    return EGX_OK;
}

// One chunk of <= mb_alloc points whose raw inputs are at x_dev (device).
int predict_chunk_dev(egx_gp_ctx* c, const double* x_dev, int m, double* y_dev, double* var_dev, double* c_out_dev) {
    const int mpad = round_up(m, EGX_NB);
    const bool want_var = (var_dev != nullptr);
    double* Ybuf = (want_var || c_out_dev != nullptr) ? c->Y : nullptr;
    {
        StageScope sc(c->env.prof, EGX_STAGE_CROSS_CORR, 1, c->stream);
        launch_cross_corr(c->corr, x_dev, m, mpad, c->x_mean, c->x_std, c->X, c->n, c->npad, c->d, c->terms,
                          c->nterms, c->rho /* = gamma after finalize */, c->beta, c->basis_i, c->basis_j, c->p,
                          c->y_mean, c->y_std, Ybuf, c->npad, y_dev, c->stream);
    }
    if (c_out_dev != nullptr)
        EGX_CUDA_TRY(cudaMemcpy2DAsync(c_out_dev, static_cast<size_t>(c->n) * sizeof(double), c->Y,
                                       static_cast<size_t>(c->npad) * sizeof(double),
                                       static_cast<size_t>(c->n) * sizeof(double), m, cudaMemcpyDeviceToDevice,
                                       c->stream));
    if (!want_var) return EGX_OK;
    FactorRef fr = factor_ref(c);
    if (mpad / EGX_NB >= 8) {                      // enough rows for full waves of 128 x 128 tiles
        const int stl = ensure_L_slices(c);
        if (stl != EGX_OK) return stl;
        if (c->Lslices_ready) {
            fr.Lsl = c->Lsl;
            fr.Lsc = c->Lsc;
            fr.Lsl_off = c->Lsl_off.data();
            fr.Lsc_off = c->Lsc_off.data();
        }
    }
    blocked_sweep(c->env, fr, false, c->Y, c->npad, mpad / EGX_NB, mpad / 64);
    {
        StageScope sc(c->env.prof, EGX_STAGE_VAR_FINISH, 1, c->stream);
        launch_var_finish(c->Y, c->npad, m, c->npad, x_dev, c->x_mean, c->x_std, c->d,
                          c->M + static_cast<long>(c->npad) * c->ld, c->ld, c->G, c->p, c->basis_i, c->basis_j,
                          c->sigma2_scaled, var_dev, c->stream);
    }
    return EGX_OK;
}

int predict_impl(egx_gp_ctx* c, const double* x, int m, double* y, double* var, bool device_ptrs) {
    if (!c->trained) {
        egx_set_error("predict* called before a successful egx_gp_finalize");
        return EGX_INVALID_VALUE;
    }
    if (m < 0 || (m > 0 && x == nullptr)) {
        egx_set_error("bad prediction input");
        return EGX_INVALID_VALUE;
    }
    if (m == 0) return EGX_OK;
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    const int mb = std::min(round_up(m, EGX_NB), predict_chunk_points(c));
    int st = ensure_predict_buffers(c, mb);
    if (st != EGX_OK) return st;
    for (int i0 = 0; i0 < m; i0 += mb) {
        const int mc = std::min(mb, m - i0);
        const double* xd;
        double *yd = nullptr, *vd = nullptr;
        if (device_ptrs) {
            xd = x + static_cast<long>(i0) * c->d;
            if (y) yd = y + i0;
            if (var) vd = var + i0;
        } else {
            EGX_CUDA_TRY(cudaMemcpyAsync(c->xchunk, x + static_cast<long>(i0) * c->d,
                                         static_cast<size_t>(mc) * c->d * sizeof(double), cudaMemcpyHostToDevice,
                                         c->stream));
            xd = c->xchunk;
            if (y) yd = c->ychunk;
            if (var) vd = c->vchunk;
        }
        st = predict_chunk_dev(c, xd, mc, yd, vd, nullptr);
        if (st != EGX_OK) return st;
        if (!device_ptrs) {
            if (y) EGX_CUDA_TRY(cudaMemcpyAsync(y + i0, c->ychunk, mc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            if (var) EGX_CUDA_TRY(cudaMemcpyAsync(var + i0, c->vchunk, mc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            // the staging buffers are reused by the next chunk
            EGX_CUDA_TRY(cudaStreamSynchronize(c->stream));
        }
    }
    EGX_CUDA_TRY(cudaStreamSynchronize(c->stream));
    EGX_CUDA_TRY(cudaGetLastError());
    resolve_profile(c);
    return EGX_OK;
}

void free_ctx(egx_gp_ctx* c) {
    if (!c) return;
    egx_dev_free(c->tgW);
    egx_dev_free(c->tgRinv);
    egx_dev_free(c->tgPartial);
    egx_dev_free(c->tgGrad);
    egx_dev_free(c->tgGamma);
    egx_dev_free(c->tgTerms);
    egx_host_free(c->tgGrad_h);
    egx_host_free(c->tgTerms_h);
    egx_dev_free(c->Lsl);
    egx_dev_free(c->Lsc);
    c->Lsl = nullptr;
    c->Lsc = nullptr;
    for (egx_gp_ctx* r : c->replicas) free_ctx(r);
    c->replicas.clear();
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->eval_graph) cudaGraphExecDestroy(c->eval_graph);
    c->eval_graph = nullptr;
    if (c->timer_a) cudaEventDestroy(c->timer_a);
    if (c->timer_b) cudaEventDestroy(c->timer_b);
    egx_dev_free(c->X);
    egx_dev_free(c->ynorm);
    egx_dev_free(c->x_mean);
    egx_dev_free(c->x_std);
    egx_dev_free(c->FyT);
    egx_dev_free(c->basis_i);
    egx_dev_free(c->basis_j);
    egx_dev_free(c->terms);
    egx_dev_free(c->M);
    egx_dev_free(c->LT);
    egx_dev_free(c->KF);
    egx_dev_free(c->W_dev);
    egx_dev_free(c->sb_thetas);
    egx_dev_free(c->sb_G);
    egx_dev_free(c->sb_out);
    egx_host_free(c->sb_out_h);
    egx_host_free(c->sb_G_h);
    egx_host_free(c->sb_thetas_h);
    egx_dev_free(c->Dinv);
    egx_dev_free(c->glswork);
    egx_dev_free(c->G);
    egx_dev_free(c->beta);
    egx_dev_free(c->rho);
    egx_dev_free(c->res);
    egx_dev_free(c->info);
    egx_dev_free(c->Y);
    egx_dev_free(c->xchunk);
    egx_dev_free(c->ychunk);
    egx_dev_free(c->vchunk);
    egx_host_free(c->terms_h);
    egx_host_free(c->res_h);
    egx_host_free(c->G_h);
    egx_host_free(c->beta_h);
    c->env.destroy();
    delete c;
}

}  // namespace

// ---------------------------------------------------------------------------
extern "C" int egx_gp_create(egx_gp_ctx** out, int device, int corr, int mean, const double* xnorm, int n, int d,
                             const double* ynorm, const double* x_mean, const double* x_std, double y_mean,
                             double y_std, const double* w_star, int h, double nugget) try {
    if (!out) return EGX_INVALID_VALUE;
    *out = nullptr;
    if (n < 1 || d < 1 || h < 1 || h > d || !xnorm || !ynorm || !x_mean || !x_std || !w_star || corr < 0 ||
        corr > 3 || mean < 0 || mean > 2) {
        egx_set_error("egx_gp_create: invalid argument (n=%d d=%d h=%d corr=%d mean=%d)", n, d, h, corr, mean);
        return EGX_INVALID_VALUE;
    }
    const int p = (mean == EGX_MEAN_CONSTANT) ? 1 : (mean == EGX_MEAN_LINEAR ? d + 1 : (d + 1) * (d + 2) / 2);
    if (p > 255 || p > n) {
        egx_set_error("egx_gp_create: regression basis size p=%d unsupported (need p <= min(255, n=%d))", p, n);
        return EGX_INVALID_VALUE;
    }
    {
        // the correlation kernels stage two raw 64 x d coordinate tiles, two scaled term-major tiles [terms][64] and the
        // (dimension, component) term list in shared memory (kernels_corr.cu): refuse shapes that cannot fit instead of failing
        // at the first launch
        // (Matern: one term per NON-ZERO weight W_jl -- d for the identity, up to d * h with KPLS rotations; the exponential
        // kernels fold the components of a dimension into one term)
        size_t terms = d;
        if (corr == EGX_CORR_MATERN32 || corr == EGX_CORR_MATERN52) {
            terms = 0;
            for (long e = 0; e < static_cast<long>(d) * h; ++e) terms += (w_star[e] != 0.0) ? 1 : 0;
        }
        const size_t smem = 16 + (2 * static_cast<size_t>(EGX_CT) * (d + terms) + 2 * EGX_CT) * sizeof(double) + terms * sizeof(CorrTerm);
        if (smem > 227 * 1024) {
            egx_set_error("egx_gp_create: d = %d (h = %d) needs %zu bytes of shared memory per CTA in the correlation kernels "
                          "(limit 232448); reduce the input dimension (KPLS does not shrink the coordinate tiles)", d, h, smem);
            return EGX_INVALID_VALUE;
        }
    }
    if (egx_device_count() <= device || device < 0) {
        egx_set_error("egx_gp_create: CUDA device %d not available (no CPU fallback exists)", device);
        return EGX_CUDA_ERROR;
    }
    EGX_CUDA_TRY(cudaSetDevice(device));
    egx_gp_ctx* c = new egx_gp_ctx();
    c->device = device;
    c->corr = corr;
    c->mean = mean;
    c->n = n;
    c->d = d;
    c->h = h;
    c->p = p;
    c->q = p + 1;
    c->npad = round_up(n, EGX_NB);
    c->qpad = round_up(c->q, EGX_NB);
    c->rows_total = c->npad + c->qpad;
    c->ld = c->npad;
    c->nugget = nugget;
    c->y_mean = y_mean;
    c->y_std = y_std;
    c->w_star.assign(w_star, w_star + static_cast<size_t>(d) * h);
    c->xnorm_h.assign(xnorm, xnorm + static_cast<size_t>(n) * d);
    c->ynorm_h.assign(ynorm, ynorm + n);
    c->x_mean_h.assign(x_mean, x_mean + d);
    c->x_std_h.assign(x_std, x_std + d);
    // regression basis f_l(x) = v(bi) * v(bj), v(-1) = 1  (mean_models.rs:42-44, 68-71, 97-104)
    c->basis_i_h.push_back(-1);
    c->basis_j_h.push_back(-1);
    if (mean >= EGX_MEAN_LINEAR)
        for (int j = 0; j < d; ++j) {
            c->basis_i_h.push_back(j);
            c->basis_j_h.push_back(-1);
        }
    if (mean == EGX_MEAN_QUADRATIC)
        for (int k = 0; k < d; ++k)
            for (int j = k; j < d; ++j) {
                c->basis_i_h.push_back(j);
                c->basis_j_h.push_back(k);
            }

#define EGX_CREATE_TRY(expr)                                                          \
    do {                                                                              \
        cudaError_t e__ = (expr);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            egx_set_error("%s failed: %s", #expr, cudaGetErrorString(e__));           \
            free_ctx(c);                                                              \
            return EGX_CUDA_ERROR;                                                    \
        }                                                                             \
    } while (0)

    if (c->env.init(c->npad / EGX_NB) != EGX_OK) {
        free_ctx(c);
        return EGX_CUDA_ERROR;
    }
    c->stream = c->env.sb;
    // measured (tools/midsize_probe.py, configs_probe.py): replay wins 1.5-2.7x for n <= 2048 and 6 % at n = 4096,
    // and loses 2.5 % at n = 8192 where the evaluation is throughput- not launch-bound
    // r02: above 4096 the replay is used while the look-ahead schedule is on, i.e. with fewer than 6 evaluations in flight
    // (single evaluations, the chains of one rank of a sharded fit): 5.24 -> 4.86 ms at n = 8192 (profiles/r02/y5_single.txt)
    c->use_graphs = c->npad <= 4096;
    c->use_graphs_lookahead = true;
    if (const char* e = getenv("EGX_GRAPHS")) c->use_graphs = c->use_graphs_lookahead = atoi(e) != 0;
    const size_t xbytes = static_cast<size_t>(c->npad) * d * sizeof(double);
    EGX_CREATE_TRY(egx_dev_malloc(&c->X, xbytes));
    EGX_CREATE_TRY(cudaMemsetAsync(c->X, 0, xbytes, c->stream));
    EGX_CREATE_TRY(cudaMemcpyAsync(c->X, xnorm, static_cast<size_t>(n) * d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    EGX_CREATE_TRY(egx_dev_malloc(&c->ynorm, c->npad * sizeof(double)));
    EGX_CREATE_TRY(cudaMemsetAsync(c->ynorm, 0, c->npad * sizeof(double), c->stream));
    EGX_CREATE_TRY(cudaMemcpyAsync(c->ynorm, ynorm, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    EGX_CREATE_TRY(egx_dev_malloc(&c->x_mean, d * sizeof(double)));
    EGX_CREATE_TRY(egx_dev_malloc(&c->x_std, d * sizeof(double)));
    EGX_CREATE_TRY(cudaMemcpyAsync(c->x_mean, x_mean, d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    EGX_CREATE_TRY(cudaMemcpyAsync(c->x_std, x_std, d * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    EGX_CREATE_TRY(egx_dev_malloc(&c->basis_i, p * sizeof(int)));
    EGX_CREATE_TRY(egx_dev_malloc(&c->basis_j, p * sizeof(int)));
    EGX_CREATE_TRY(cudaMemcpyAsync(c->basis_i, c->basis_i_h.data(), p * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    EGX_CREATE_TRY(cudaMemcpyAsync(c->basis_j, c->basis_j_h.data(), p * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    EGX_CREATE_TRY(egx_dev_malloc(&c->W_dev, static_cast<size_t>(d) * h * sizeof(double)));
    EGX_CREATE_TRY(cudaMemcpyAsync(c->W_dev, w_star, static_cast<size_t>(d) * h * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    c->max_terms = d * h;
    EGX_CREATE_TRY(egx_dev_malloc(&c->terms, c->max_terms * sizeof(CorrTerm)));
    EGX_CREATE_TRY(egx_host_malloc(&c->terms_h, c->max_terms * sizeof(CorrTerm)));
    EGX_CREATE_TRY(egx_dev_malloc(&c->FyT, static_cast<size_t>(c->q) * c->ld * sizeof(double)));
    const size_t mbytes = static_cast<size_t>(c->rows_total) * c->ld * sizeof(double);
    EGX_CREATE_TRY(egx_dev_malloc(&c->M, mbytes));
    EGX_CREATE_TRY(cudaMemsetAsync(c->M, 0, mbytes, c->stream));
    EGX_CREATE_TRY(egx_dev_malloc(&c->Dinv, static_cast<size_t>(c->npad / EGX_NB) * 4096 * sizeof(double)));
    if (c->env.ensure_panel_rows(c->rows_total) != EGX_OK) {
        free_ctx(c);
        return EGX_CUDA_ERROR;
    }
    EGX_CREATE_TRY(egx_dev_malloc(&c->glswork, static_cast<size_t>(c->q) * c->npad * sizeof(double)));
    EGX_CREATE_TRY(egx_dev_malloc(&c->G, static_cast<size_t>(p) * p * sizeof(double)));
    EGX_CREATE_TRY(egx_dev_malloc(&c->beta, p * sizeof(double)));
    EGX_CREATE_TRY(egx_dev_malloc(&c->rho, c->npad * sizeof(double)));
    EGX_CREATE_TRY(egx_dev_malloc(&c->res, sizeof(EvalResult)));
    EGX_CREATE_TRY(egx_dev_malloc(&c->info, sizeof(int)));
    EGX_CREATE_TRY(egx_host_malloc(&c->res_h, sizeof(EvalResult)));
    EGX_CREATE_TRY(egx_host_malloc(&c->G_h, static_cast<size_t>(p) * p * sizeof(double)));
    EGX_CREATE_TRY(egx_host_malloc(&c->beta_h, p * sizeof(double)));
    launch_mean_basis_rows(c->X, n, c->npad, d, c->basis_i, c->basis_j, p, c->ynorm, c->FyT, c->ld, c->stream);
    EGX_CREATE_TRY(cudaStreamSynchronize(c->stream));
    EGX_CREATE_TRY(cudaGetLastError());
#undef EGX_CREATE_TRY
    *out = c;
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" void egx_gp_destroy(egx_gp_ctx* ctx) { free_ctx(ctx); }

extern "C" int egx_gp_dims(const egx_gp_ctx* c, int* n, int* d, int* h, int* p) try {
    if (!c) return EGX_INVALID_VALUE;
    if (n) *n = c->n;
    if (d) *d = c->d;
    if (h) *h = c->h;
    if (p) *p = c->p;
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_reduced_likelihood(egx_gp_ctx* c, const double* theta, double* rlf) try {
    if (!c || !theta || !rlf) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    if (small_path_ok(c)) {
        int st = EGX_OK;
        const int rc = evaluate_small_batch(c, theta, 1, rlf, &st);
        return rc != EGX_OK ? rc : st;
    }
    return evaluate(c, theta, rlf);
}
EGX_ABI_CATCH

extern "C" int egx_gp_reduced_likelihood_batch(egx_gp_ctx* c, const double* thetas, int B, double* rlf, int* status) try {
    if (!c || !thetas || !rlf || !status || B < 0) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    if (small_path_ok(c)) return evaluate_small_batch(c, thetas, B, rlf, status);
    // Large n: keep `W` independent evaluations in flight on W replicas of the workspace (own streams):
    // the serial diagonal-block / panel chain of one factorisation overlaps the bulk updates of the others.
    // how many: the smaller the matrix, the more an evaluation is a latency chain (block columns x ~80 us)
    // rather than throughput work, and the cheaper a replica is (npad^2 x 8 bytes)
    // measured (tools/batch_sweep.py): n = 8192 plateaus at 4 (6.68 ms per evaluation; 7.23 at 2, 6.69 at 8),
    // n = 4096 keeps improving to 8-12 (1.25 ms at 4, 1.08 at 6, 1.03 at 8 and 12); 12 also takes the 11
    // chains of a fit in one wave
    // with the tcgen05 trailing update an evaluation is ~40 % less GPU work while its serial panel chain is unchanged:
    // more evaluations in flight pay (n = 8192: 4.04 / 4.05 / 4.02 ms at W = 4 / 6 / 8 with look-ahead, 4.08 / 3.74 /
    // 3.70 ms without -- the other evaluations hide the chain better than look-ahead does, and the whole K = 256
    // update then goes through the tcgen05 kernel in one launch)
    int W = (c->npad <= 4096) ? 16 : (c->env.ozaki ? 8 : 4);
    if (const char* e = getenv("EGX_BATCH_STREAMS")) W = std::max(1, atoi(e));
    static const int batch_la = getenv("EGX_BATCH_LOOKAHEAD") != nullptr ? atoi(getenv("EGX_BATCH_LOOKAHEAD")) : -1;
    W = std::min(W, B);
    while (static_cast<int>(c->replicas.size()) < W - 1) {
        egx_gp_ctx* r = nullptr;
        const int st = egx_gp_create(&r, c->device, c->corr, c->mean, c->xnorm_h.data(), c->n, c->d, c->ynorm_h.data(),
                                     c->x_mean_h.data(), c->x_std_h.data(), c->y_mean, c->y_std, c->w_star.data(),
                                     c->h, c->nugget);
        if (st != EGX_OK) return st;
        r->env.prof.on = c->env.prof.on;
        c->replicas.push_back(r);
    }
    std::vector<egx_gp_ctx*> ws;
    ws.push_back(c);
    for (int i = 0; i < W - 1; ++i) ws.push_back(c->replicas[i]);
    std::vector<int> owner(W, -1);     // which candidate each workspace is working on
    for (int b = 0; b < B + W; ++b) {
        egx_gp_ctx* w = ws[b % W];
        const int prev = owner[b % W];
        if (prev >= 0) {
            if (status[prev] == EGX_OK) status[prev] = evaluate_collect(w, &rlf[prev]);
            if (status[prev] == EGX_CUDA_ERROR) return EGX_CUDA_ERROR;
            owner[b % W] = -1;
        }
        if (b < B) {
            rlf[b] = NAN;
            w->env.oz_persist = (W > 1) ? 1 : 0;
            const bool la_saved = w->env.lookahead;
            if (batch_la >= 0) w->env.lookahead = la_saved && batch_la != 0;
            else if (W >= 6 && w->env.ozaki && c->npad > 4096) w->env.lookahead = false;   // re-measured with the look-ahead columns on tcgen05 (profiles/r02/y19_*.txt): look-ahead ON wins up to n = 4096 (C5 n = 2048: 70.9 -> 62.0 ms, n = 4096: 0.570 -> 0.541 ms per evaluation), ties at n = 8192 (3.20 vs 3.22)
            status[b] = evaluate_launch(w, thetas + static_cast<long>(b) * c->h);
            w->env.lookahead = la_saved;
            if (status[b] == EGX_CUDA_ERROR) return EGX_CUDA_ERROR;
            owner[b % W] = b;
        }
    }
    return EGX_OK;
}
EGX_ABI_CATCH

// ---------------------------------------------------------------------------------------------------------------
// Asynchronous seam for independent optimiser chains (the rayon fan-out of gp/src/algorithm.rs:928-945 runs every
// chain on its own): `slots` workspaces, each holding at most one evaluation in flight; a chain whose evaluation is
// back proposes its next theta while the evaluations of the other chains are still running, so there is no
// per-iteration barrier across chains.
// ---------------------------------------------------------------------------------------------------------------
extern "C" int egx_gp_async_slots(egx_gp_ctx* c, int wanted) try {
    if (!c || wanted < 1) return 0;
    std::lock_guard<std::mutex> lk(c->mu);
    if (cudaSetDevice(c->device) != cudaSuccess) return 0;
    if (small_path_ok(c)) return 0;               // one CTA per theta: a lock-step batch is one launch, keep it
    int W = (c->npad <= 4096) ? 16 : (c->env.ozaki ? 8 : 4);
    if (const char* e = getenv("EGX_BATCH_STREAMS")) W = std::max(1, atoi(e));
    W = std::min(W, wanted);
    while (static_cast<int>(c->replicas.size()) < W - 1) {
        egx_gp_ctx* r = nullptr;
        const int st = egx_gp_create(&r, c->device, c->corr, c->mean, c->xnorm_h.data(), c->n, c->d, c->ynorm_h.data(),
                                     c->x_mean_h.data(), c->x_std_h.data(), c->y_mean, c->y_std, c->w_star.data(),
                                     c->h, c->nugget);
        if (st != EGX_OK) return static_cast<int>(c->replicas.size()) + 1;
        r->env.prof.on = c->env.prof.on;
        c->replicas.push_back(r);
    }
    c->async_slots = W;
    return W;
}
EGX_ABI_CATCH_COUNT

// Give the replica workspaces of the batched / asynchronous entry points back to the block cache (each holds a full
// (npad + 128) x npad matrix, panels, slices, streams: ~0.6 GB at n = 8192).  A trained model only needs the primary
// workspace for predict*; the next batch or fit re-creates replicas from the cache.
extern "C" int egx_gp_release_workspaces(egx_gp_ctx* c) try {
    if (!c) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    if (cudaSetDevice(c->device) != cudaSuccess) return EGX_CUDA_ERROR;
    for (egx_gp_ctx* r : c->replicas) free_ctx(r);
    c->replicas.clear();
    c->async_slots = 0;
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_eval_begin(egx_gp_ctx* c, int slot, const double* theta) try {
    if (!c || !theta || slot < 0 || slot >= c->async_slots) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    egx_gp_ctx* w = slot == 0 ? c : c->replicas[slot - 1];
    static const int batch_la = getenv("EGX_BATCH_LOOKAHEAD") != nullptr ? atoi(getenv("EGX_BATCH_LOOKAHEAD")) : -1;
    const int W = c->async_slots;
    w->env.oz_persist = (W > 1) ? 1 : 0;
    const bool la_saved = w->env.lookahead;
    if (batch_la >= 0) w->env.lookahead = la_saved && batch_la != 0;
    else if (W >= 6 && w->env.ozaki && c->npad > 4096) w->env.lookahead = false;   // re-measured with the look-ahead columns on tcgen05 (profiles/r02/y19_*.txt): look-ahead ON wins up to n = 4096 (C5 n = 2048: 70.9 -> 62.0 ms, n = 4096: 0.570 -> 0.541 ms per evaluation), ties at n = 8192 (3.20 vs 3.22)
    const int st = evaluate_launch(w, theta);
    w->env.lookahead = la_saved;
    return st;
}
EGX_ABI_CATCH

extern "C" int egx_gp_eval_end(egx_gp_ctx* c, int slot, double* rlf) try {
    if (!c || !rlf || slot < 0 || slot >= c->async_slots) return EGX_INVALID_VALUE;
    egx_gp_ctx* w = nullptr;
    {
        std::lock_guard<std::mutex> lk(c->mu);
        w = slot == 0 ? c : c->replicas[slot - 1];
        if (!w->pending_eval) {
            egx_set_error("egx_gp_eval_end: no evaluation in flight on slot %d", slot);
            return EGX_INVALID_VALUE;
        }
    }
    // the wait happens OUTSIDE the context lock: a slot is touched by one caller at a time (contract), and the
    // threads that own the other slots must be able to enqueue meanwhile (one rayon worker per slot)
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    return evaluate_collect(w, rlf);
}
EGX_ABI_CATCH

extern "C" int egx_gp_reduced_likelihood_grad(egx_gp_ctx* c, const double* theta, double rel_step, double* rlf,
                                              double* grad) try {
    if (!c || !theta || !rlf || !grad || !(rel_step > 0.0)) return EGX_INVALID_VALUE;
    const int h = c->h, B = 2 * h + 1;
    std::vector<double> th(static_cast<size_t>(B) * h), val(B);
    std::vector<int> st(B);
    for (int b = 0; b < B; ++b)
        for (int l = 0; l < h; ++l) th[static_cast<size_t>(b) * h + l] = theta[l];
    for (int k = 0; k < h; ++k) {
        const double dk = rel_step * theta[k];
        th[static_cast<size_t>(1 + 2 * k) * h + k] = theta[k] + dk;
        th[static_cast<size_t>(2 + 2 * k) * h + k] = theta[k] - dk;
    }
    const int rc = egx_gp_reduced_likelihood_batch(c, th.data(), B, val.data(), st.data());
    if (rc != EGX_OK) return rc;
    *rlf = val[0];
    for (int k = 0; k < h; ++k) {
        const double hi = th[static_cast<size_t>(1 + 2 * k) * h + k], lo = th[static_cast<size_t>(2 + 2 * k) * h + k];
        grad[k] = (st[1 + 2 * k] == EGX_OK && st[2 + 2 * k] == EGX_OK) ? (val[1 + 2 * k] - val[2 + 2 * k]) / (hi - lo) : NAN;
    }
    return st[0];
}
EGX_ABI_CATCH

// d rlf / d theta in closed form (kernels_thetagrad.cu):  one evaluation, W = L^-T by the multi-RHS sweep on the identity
// (only the row tiles above each column pair: a third of the full sweep), gamma = W rho, -R^-1 = -W W^T (W is upper triangular: column panel kp only touches the rows above its end, so the SYRK is a
// sum of growing triangles -- on tcgen05 where the triangle is large enough, else DMMA), then the pair kernel.
extern "C" int egx_gp_reduced_likelihood_grad_analytic(egx_gp_ctx* c, const double* theta, double* rlf, double* grad) try {
    if (!c || !theta || !rlf || !grad) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    const int h = c->h, npad = c->npad, T = npad / EGX_NB;
    for (int l = 0; l < h; ++l) grad[l] = NAN;
    *rlf = NAN;
    if (h > 32) {
        egx_set_error("closed-form theta gradient supports up to 32 components (got %d): use egx_gp_reduced_likelihood_grad", h);
        return EGX_INVALID_VALUE;
    }
    {
        // shared memory of the pair kernel: two 64 x d coordinate tiles, gamma, the warp partials and up to d * h terms
        const int hmax = h <= 8 ? 8 : (h <= 16 ? 16 : 32);
        const size_t smem = (2 * static_cast<size_t>(EGX_CT) * c->d + 2 * EGX_CT + 8 * hmax) * sizeof(double) +
                            static_cast<size_t>(c->d) * h * sizeof(ThetaGradTerm);
        if (smem > 227 * 1024) {
            egx_set_error("closed-form theta gradient: d = %d with %d components needs %zu bytes of shared memory per CTA "
                          "(limit 232448): use egx_gp_reduced_likelihood_grad", c->d, h, smem);
            return EGX_INVALID_VALUE;
        }
    }
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    int st = evaluate(c, theta, rlf);
    if (st != EGX_OK) return st;
    const int nblocks = theta_grad_blocks(npad);
    if (c->tgW == nullptr) {
        const size_t sq = static_cast<size_t>(npad) * npad * sizeof(double);
        EGX_CUDA_TRY(egx_dev_malloc(&c->tgW, sq));
        EGX_CUDA_TRY(egx_dev_malloc(&c->tgRinv, sq));
        EGX_CUDA_TRY(egx_dev_malloc(&c->tgPartial, static_cast<size_t>(nblocks) * h * sizeof(double)));
        EGX_CUDA_TRY(egx_dev_malloc(&c->tgGrad, h * sizeof(double)));
        EGX_CUDA_TRY(egx_dev_malloc(&c->tgGamma, static_cast<size_t>(npad) * sizeof(double)));
        EGX_CUDA_TRY(egx_dev_malloc(&c->tgTerms, static_cast<size_t>(c->d) * h * sizeof(ThetaGradTerm)));
        EGX_CUDA_TRY(egx_host_malloc(&c->tgGrad_h, h * sizeof(double)));
        EGX_CUDA_TRY(egx_host_malloc(&c->tgTerms_h, static_cast<size_t>(c->d) * h * sizeof(ThetaGradTerm)));
    }
    if (c->env.ensure_panel_rows(npad) != EGX_OK) return EGX_CUDA_ERROR;
    const int nt = egx_fill_theta_grad_terms(c->corr, c->d, h, c->w_star.data(), theta, c->tgTerms_h);
    if (nt > 0)
        EGX_CUDA_TRY(cudaMemcpyAsync(c->tgTerms, c->tgTerms_h, nt * sizeof(ThetaGradTerm), cudaMemcpyHostToDevice, c->stream));
    // W = I L^-T
    launch_set_identity(c->tgW, npad, npad, c->stream);
    FactorRef fr = factor_ref(c);
    if (T >= 8) {
        st = ensure_L_slices(c);
        if (st != EGX_OK) return st;
        if (c->Lslices_ready) {
            fr.Lsl = c->Lsl;
            fr.Lsc = c->Lsc;
            fr.Lsl_off = c->Lsl_off.data();
            fr.Lsc_off = c->Lsc_off.data();
        }
    }
    // r02b: the column panels of W W^T do not wait for the whole sweep -- panel kp only needs the two block columns of pair kp,
    // final as soon as the sweep has solved them.  The sweep records one event per pair; -R^-1 is accumulated on the side stream
    // `sq` from its own slice buffers, so its tcgen05 launches fill the gaps the sweep's small kernels (two solves, the K = 128
    // partner update, slicing: ~60 us per pair) leave on the GPU.  EGX_GRAD_PIPELINE=0 restores the serial order.
    static const int pipeline = getenv("EGX_GRAD_PIPELINE") != nullptr ? atoi(getenv("EGX_GRAD_PIPELINE")) : 1;
    const bool oz = c->env.ozaki && c->env.oz_S != nullptr && T >= c->env.ozaki_min_T;
    const bool piped = pipeline && c->env.sq != nullptr && c->env.oz_S2 != nullptr && static_cast<int>(c->env.ev_partner.size()) >= (T + 1) / 2;
    cudaStream_t sw = piped ? c->env.sq : c->stream;              // stream of the W W^T accumulation
    int8_t* wS = piped ? c->env.oz_S2 : c->env.oz_S;
    double* wScale = piped ? c->env.oz_scale2 : c->env.oz_scale;
    if (piped) {
        cudaEventRecord(c->env.ev_fork, c->stream);              // the evaluation (and whatever used tgRinv before) is ahead of sq
        cudaStreamWaitEvent(sw, c->env.ev_fork, 0);
        EGX_CUDA_TRY(cudaMemsetAsync(c->tgRinv, 0, static_cast<size_t>(npad) * npad * sizeof(double), sw));
        c->env.solve_pair_events = &c->env.ev_partner;
    }
    blocked_sweep(c->env, fr, false, c->tgW, npad, T, npad / 64, true);
    c->env.solve_pair_events = nullptr;
    {
        StageScope sc(c->env.prof, EGX_STAGE_BACKSOLVE, 1, c->stream);   // gamma = L^-T rho = W rho (algorithm.rs:1034)
        launch_upper_gemv(c->tgW, npad, c->n, c->rho, c->tgGamma, c->stream);
    }
    // -R^-1 = 0 - W W^T, lower block triangle
    if (!piped) EGX_CUDA_TRY(cudaMemsetAsync(c->tgRinv, 0, static_cast<size_t>(npad) * npad * sizeof(double), c->stream));
    for (int k = 0; k < T; k += 2) {
        const int kw = (k + 1 < T) ? 2 : 1;                   // block columns in this panel
        const int rt = k + kw;                                // tile rows of W that are non-zero in it
        const double* P = c->tgW + static_cast<long>(k) * EGX_NB;
        if (piped) cudaStreamWaitEvent(sw, c->env.ev_partner[k >> 1], 0);
        if (oz && kw == 2 && rt >= c->env.ozaki_min_tri) {
            {
                StageScope sc(c->env.prof, EGX_STAGE_OZAKI_SLICE, 2, sw);
                launch_ozaki_slice(P, npad, rt * EGX_NB, wScale, wS, sw);
            }
            StageScope sc(c->env.prof, EGX_STAGE_OZAKI_SYRK, 1, sw);
            launch_ozaki_syrk(c->tgRinv, npad, wS, wScale, rt, rt, sw);
        } else {
            GemmArgs g;
            g.C = c->tgRinv;
            g.ldc = npad;
            g.A = P;
            g.lda = npad;
            g.B = P;
            g.ldb = npad;
            g.K = kw * EGX_NB;
            g.tri = rt;
            g.Mt = g.Nt = rt;
            StageScope sc(c->env.prof, EGX_STAGE_SYRK_GEMM, 1, sw);
            launch_gemm_nt_sub(g, sw);
        }
    }
    if (piped) {
        cudaEventRecord(c->env.ev_join_q, sw);
        cudaStreamWaitEvent(c->stream, c->env.ev_join_q, 0);
    }
    {
        StageScope sc(c->env.prof, EGX_STAGE_THETA_GRAD, 2, c->stream);
        launch_theta_grad(c->corr, c->X, c->n, npad, c->d, c->tgTerms, nt, c->tgRinv, npad, c->tgGamma, c->res, h, c->tgPartial,
                          c->tgGrad, c->stream);
    }
    EGX_CUDA_TRY(cudaMemcpyAsync(c->tgGrad_h, c->tgGrad, h * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    EGX_CUDA_TRY(cudaStreamSynchronize(c->stream));
    EGX_CUDA_TRY(cudaGetLastError());
    resolve_profile(c);
    std::memcpy(grad, c->tgGrad_h, h * sizeof(double));
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_finalize(egx_gp_ctx* c, const double* theta, double* rlf, double* sigma2, double* beta,
                               double* gamma, double* ft, double* ft_qr_r) try {
    if (!c || !theta) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    double v = NAN;
    int st = evaluate(c, theta, &v);
    if (rlf) *rlf = v;
    if (st != EGX_OK) return st;
    // gamma = L^-T rho, blocked back substitution (algorithm.rs:1034)
    backsolve_vector(c->env, factor_ref(c), c->rho);
    if (gamma) EGX_CUDA_TRY(cudaMemcpyAsync(gamma, c->rho, c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    std::vector<double> ftT;
    if (ft) {
        ftT.resize(static_cast<size_t>(c->p) * c->n);
        EGX_CUDA_TRY(cudaMemcpy2DAsync(ftT.data(), c->n * sizeof(double), c->M + static_cast<long>(c->npad) * c->ld,
                                       c->ld * sizeof(double), c->n * sizeof(double), c->p, cudaMemcpyDeviceToHost,
                                       c->stream));
    }
    EGX_CUDA_TRY(cudaStreamSynchronize(c->stream));
    EGX_CUDA_TRY(cudaGetLastError());
    resolve_profile(c);
    if (ft)
        for (int i = 0; i < c->n; ++i)
            for (int l = 0; l < c->p; ++l) ft[static_cast<size_t>(i) * c->p + l] = ftT[static_cast<size_t>(l) * c->n + i];
    c->sigma2_scaled = c->res_h->sigma2 * c->y_std * c->y_std;   // algorithm.rs:1048
    if (sigma2) *sigma2 = c->sigma2_scaled;
    if (beta) std::memcpy(beta, c->beta_h, c->p * sizeof(double));
    if (ft_qr_r) std::memcpy(ft_qr_r, c->G_h, sizeof(double) * c->p * c->p);
    c->theta.assign(theta, theta + c->h);
    c->trained = true;
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_download_chol(egx_gp_ctx* c, double* r_chol) try {
    if (!c || !r_chol) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->trained) {
        egx_set_error("egx_gp_download_chol before finalize");
        return EGX_INVALID_VALUE;
    }
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    EGX_CUDA_TRY(cudaMemcpy2D(r_chol, c->n * sizeof(double), c->M, c->ld * sizeof(double), c->n * sizeof(double), c->n,
                              cudaMemcpyDeviceToHost));
    for (int i = 0; i < c->n; ++i)
        for (int j = i + 1; j < c->n; ++j) r_chol[static_cast<size_t>(i) * c->n + j] = 0.0;
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_predict(egx_gp_ctx* c, const double* x, int m, double* y) try {
    if (!c || !y) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    return predict_impl(c, x, m, y, nullptr, false);
}
EGX_ABI_CATCH
extern "C" int egx_gp_predict_var(egx_gp_ctx* c, const double* x, int m, double* var) try {
    if (!c || !var) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    return predict_impl(c, x, m, nullptr, var, false);
}
EGX_ABI_CATCH
extern "C" int egx_gp_predict_valvar(egx_gp_ctx* c, const double* x, int m, double* y, double* var) try {
    if (!c || !y || !var) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    return predict_impl(c, x, m, y, var, false);
}
EGX_ABI_CATCH
namespace {
int predict_gradients_impl(egx_gp_ctx* c, const double* x, int m, double* grad, bool device_ptrs);
int predict_var_gradients_impl(egx_gp_ctx* c, const double* x, int m, double* grad, bool device_ptrs);
}  // namespace
extern "C" int egx_gp_predict_gradients(egx_gp_ctx* c, const double* x, int m, double* grad) try {
    if (!c || !x || !grad || m < 0) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    return predict_gradients_impl(c, x, m, grad, false);
}
EGX_ABI_CATCH
extern "C" int egx_gp_predict_var_gradients(egx_gp_ctx* c, const double* x, int m, double* grad) try {
    if (!c || !x || !grad || m < 0) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    return predict_var_gradients_impl(c, x, m, grad, false);
}
EGX_ABI_CATCH
/* device-pointer variants: x_dev (m x d) and grad_dev (m x d) live on the context's device */
extern "C" int egx_gp_predict_gradients_dev(egx_gp_ctx* c, const double* x_dev, int m, double* grad_dev) try {
    if (!c || !x_dev || !grad_dev || m < 0) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    return predict_gradients_impl(c, x_dev, m, grad_dev, true);
}
EGX_ABI_CATCH
extern "C" int egx_gp_predict_var_gradients_dev(egx_gp_ctx* c, const double* x_dev, int m, double* grad_dev) try {
    if (!c || !x_dev || !grad_dev || m < 0) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    return predict_var_gradients_impl(c, x_dev, m, grad_dev, true);
}
EGX_ABI_CATCH
namespace {
int predict_gradients_impl(egx_gp_ctx* c, const double* x, int m, double* grad, bool device_ptrs) {
    if (!c->trained) {
        egx_set_error("predict_gradients called before a successful egx_gp_finalize");
        return EGX_INVALID_VALUE;
    }
    if (c->d > 32) {
        egx_set_error("predict_gradients supports input dimension <= 32 (got %d)", c->d);
        return EGX_INVALID_VALUE;
    }
    if (m == 0) return EGX_OK;
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    const int mb = std::min(round_up(m, EGX_NB), PREDICT_CHUNK);
    double *xd = nullptr, *gd = nullptr;
    if (!device_ptrs) {
        EGX_CUDA_TRY(egx_dev_malloc(&xd, static_cast<size_t>(mb) * c->d * sizeof(double)));
        EGX_CUDA_TRY(egx_dev_malloc(&gd, static_cast<size_t>(mb) * c->d * sizeof(double)));
    }
    int st = EGX_OK;
    for (int i0 = 0; i0 < m && st == EGX_OK; i0 += mb) {
        const int mc = std::min(mb, m - i0);
        const double* xin = x + static_cast<long>(i0) * c->d;
        double* gout = grad + static_cast<long>(i0) * c->d;
        if (!device_ptrs) {
            cudaMemcpyAsync(xd, xin, static_cast<size_t>(mc) * c->d * sizeof(double), cudaMemcpyHostToDevice,
                            c->stream);
            xin = xd;
            gout = gd;
        }
        {
            StageScope sc(c->env.prof, EGX_STAGE_CROSS_CORR, 1, c->stream);
            launch_predict_grad(c->corr, xin, mc, c->x_mean, c->x_std, c->X, c->n, c->npad, c->d, c->terms, c->nterms,
                                c->rho, c->beta, c->basis_i, c->basis_j, c->p, c->y_std, gout, c->stream);
        }
        if (!device_ptrs)
            cudaMemcpyAsync(grad + static_cast<long>(i0) * c->d, gd, static_cast<size_t>(mc) * c->d * sizeof(double),
                            cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) st = EGX_CUDA_ERROR;
    }
    egx_dev_free(xd);
    egx_dev_free(gd);
    if (st != EGX_OK || cudaGetLastError() != cudaSuccess) {
        egx_set_error("predict_gradients: CUDA failure");
        return EGX_CUDA_ERROR;
    }
    resolve_profile(c);
    return EGX_OK;
}
int predict_var_gradients_impl(egx_gp_ctx* c, const double* x, int m, double* grad, bool device_ptrs) {
    if (!c->trained) {
        egx_set_error("predict_var_gradients called before a successful egx_gp_finalize");
        return EGX_INVALID_VALUE;
    }
    if (c->d > 32 || static_cast<long>(c->d) * c->p > 2048) {
        egx_set_error("predict_var_gradients supports d <= 32 and d * p <= 2048 (got d=%d, p=%d)", c->d, c->p);
        return EGX_INVALID_VALUE;
    }
    if (m == 0) return EGX_OK;
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    const int T = c->npad / EGX_NB;
    const long ld = c->ld;
    if (!c->grad_ready) {
        if (!c->LT) EGX_CUDA_TRY(egx_dev_malloc(&c->LT, static_cast<size_t>(c->npad) * ld * sizeof(double)));
        if (!c->KF) EGX_CUDA_TRY(egx_dev_malloc(&c->KF, static_cast<size_t>(c->p) * c->npad * sizeof(double)));
        launch_transpose_lower(c->M, ld, c->LT, T, c->stream);
        // K_F = R^-1 F = L^-T (L^-1 F): rows npad.. of M hold (L^-1 F)^T, one back substitution per basis function
        EGX_CUDA_TRY(cudaMemcpy2DAsync(c->KF, c->npad * sizeof(double), c->M + static_cast<long>(c->npad) * ld,
                                       ld * sizeof(double), c->npad * sizeof(double), c->p, cudaMemcpyDeviceToDevice,
                                       c->stream));
        for (int l = 0; l < c->p; ++l) backsolve_vector(c->env, factor_ref(c), c->KF + static_cast<long>(l) * c->npad);
        c->grad_ready = true;
    }
    const int mb = std::min(round_up(m, EGX_NB), PREDICT_CHUNK);
    int st = ensure_predict_buffers(c, mb);
    if (st != EGX_OK) return st;
    double* gd = nullptr;
    if (!device_ptrs) EGX_CUDA_TRY(egx_dev_malloc(&gd, static_cast<size_t>(mb) * c->d * sizeof(double)));
    for (int i0 = 0; i0 < m && st == EGX_OK; i0 += mb) {
        const int mc = std::min(mb, m - i0);
        const int mpad = round_up(mc, EGX_NB);
        const double* xin = x + static_cast<long>(i0) * c->d;
        double* gout = grad + static_cast<long>(i0) * c->d;
        if (!device_ptrs) {
            cudaMemcpyAsync(c->xchunk, xin, static_cast<size_t>(mc) * c->d * sizeof(double), cudaMemcpyHostToDevice,
                            c->stream);
            xin = c->xchunk;
            gout = gd;
        }
        {
            StageScope sc(c->env.prof, EGX_STAGE_CROSS_CORR, 1, c->stream);
            launch_cross_corr(c->corr, xin, mc, mpad, c->x_mean, c->x_std, c->X, c->n, c->npad, c->d, c->terms,
                              c->nterms, nullptr, nullptr, c->basis_i, c->basis_j, 0, 0.0, 1.0, c->Y, c->npad, nullptr,
                              c->stream);
        }
        blocked_sweep(c->env, factor_ref(c), false, c->Y, c->npad, mpad / EGX_NB, mpad / 64);     // Y <- C L^-T
        // backward sweep Y <- Y L^-1 on the transposed factor (W = C R^-1)
        for (int k = T - 1; k >= 0; --k) {
            double* Pk = c->env.P2[k & 1];
            {
                StageScope sc(c->env.prof, EGX_STAGE_TRSM_PANEL, 1, c->stream);
                launch_trsm_rows_upper(c->Y + static_cast<long>(k) * EGX_NB, c->npad,
                                       c->LT + static_cast<long>(k) * EGX_NB * ld + static_cast<long>(k) * EGX_NB, ld,
                                       c->Dinv + static_cast<long>(k) * 4096, Pk, mpad / 64, c->stream);
            }
            if (k > 0) {
                GemmArgs g;
                g.C = c->Y;
                g.ldc = c->npad;
                g.A = Pk;
                g.lda = EGX_NB;
                g.B = c->LT + static_cast<long>(k) * EGX_NB;      // rows 0 .. k*128-1 of LT, columns of block k
                g.ldb = ld;
                g.tri = 0;
                g.Mt = mpad / EGX_NB;
                g.Nt = k;
                StageScope sc(c->env.prof, EGX_STAGE_SYRK_GEMM, 1, c->stream);
                launch_gemm_nt_sub(g, c->stream);
            }
        }
        {
            StageScope sc(c->env.prof, EGX_STAGE_VAR_FINISH, 1, c->stream);
            launch_var_grad(c->corr, c->Y, c->npad, xin, mc, c->x_mean, c->x_std, c->X, c->n, c->npad, c->d,
                            c->terms, c->nterms, c->KF, c->npad, c->G, c->p, c->basis_i, c->basis_j, c->sigma2_scaled,
                            gout, c->stream);
        }
        if (!device_ptrs)
            cudaMemcpyAsync(grad + static_cast<long>(i0) * c->d, gd, static_cast<size_t>(mc) * c->d * sizeof(double),
                            cudaMemcpyDeviceToHost, c->stream);
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) st = EGX_CUDA_ERROR;
    }
    egx_dev_free(gd);
    if (st != EGX_OK || cudaGetLastError() != cudaSuccess) {
        egx_set_error("predict_var_gradients: CUDA failure (%s)", cudaGetErrorString(cudaGetLastError()));
        return EGX_CUDA_ERROR;
    }
    resolve_profile(c);
    return EGX_OK;
}
}  // namespace
// ---------------------------------------------------------------------------------------------
// Conditional covariance and trajectory sampling (gp/src/algorithm.rs:310-326, 383-410, 1153-1194).
// ---------------------------------------------------------------------------------------------
namespace {

struct CovWork {
    double *K = nullptr, *xn = nullptr, *U = nullptr, *xraw = nullptr, *mean = nullptr;
    int mpad = 0;
    ~CovWork() {
        egx_dev_free(K);
        egx_dev_free(xn);
        egx_dev_free(U);
        egx_dev_free(xraw);
        egx_dev_free(mean);
    }
};

constexpr int COV_MAX_POINTS = 8192;

// w.K <- sigma2 (K(x, x) - rt^T rt + u^T u), padded with the identity to mpad x mpad ; w.mean <- predict(x)
int conditional_cov_dev(egx_gp_ctx* c, const double* x, int m, CovWork& w) {
    if (!c->trained) {
        egx_set_error("covariance / sample called before a successful egx_gp_finalize");
        return EGX_INVALID_VALUE;
    }
    if (m < 1 || m > COV_MAX_POINTS) {
        egx_set_error("covariance / sample: number of locations must be in 1..%d (got %d)", COV_MAX_POINTS, m);
        return EGX_INVALID_VALUE;
    }
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    const int mpad = round_up(m, EGX_NB);
    w.mpad = mpad;
    int st = ensure_predict_buffers(c, mpad);
    if (st != EGX_OK) return st;
    EGX_CUDA_TRY(egx_dev_malloc(&w.K, static_cast<size_t>(mpad) * mpad * sizeof(double)));
    EGX_CUDA_TRY(egx_dev_malloc(&w.xn, static_cast<size_t>(mpad) * c->d * sizeof(double)));
    EGX_CUDA_TRY(egx_dev_malloc(&w.U, static_cast<size_t>(m) * c->p * sizeof(double)));
    EGX_CUDA_TRY(egx_dev_malloc(&w.xraw, static_cast<size_t>(m) * c->d * sizeof(double)));
    EGX_CUDA_TRY(egx_dev_malloc(&w.mean, static_cast<size_t>(m) * sizeof(double)));
    EGX_CUDA_TRY(cudaMemcpyAsync(w.xraw, x, static_cast<size_t>(m) * c->d * sizeof(double), cudaMemcpyHostToDevice,
                                 c->stream));
    // c(x, X) rows + the predicted mean, then rt = L^-1 c^T as rows of Y
    {
        StageScope sc(c->env.prof, EGX_STAGE_CROSS_CORR, 1, c->stream);
        launch_cross_corr(c->corr, w.xraw, m, mpad, c->x_mean, c->x_std, c->X, c->n, c->npad, c->d, c->terms, c->nterms,
                          c->rho, c->beta, c->basis_i, c->basis_j, c->p, c->y_mean, c->y_std, c->Y, c->npad, w.mean,
                          c->stream);
    }
    blocked_sweep(c->env, factor_ref(c), false, c->Y, c->npad, mpad / EGX_NB, mpad / 64);
    {
        StageScope sc(c->env.prof, EGX_STAGE_VAR_FINISH, 1, c->stream);
        launch_var_finish(c->Y, c->npad, m, c->npad, w.xraw, c->x_mean, c->x_std, c->d,
                          c->M + static_cast<long>(c->npad) * c->ld, c->ld, c->G, c->p, c->basis_i, c->basis_j,
                          c->sigma2_scaled, nullptr, c->stream, w.U);
    }
    // K(x, x): the cross-correlation kernel with the (normalised) locations as the "training" set
    launch_normalize_rows(w.xraw, m, mpad, c->d, c->x_mean, c->x_std, w.xn, c->stream);
    {
        StageScope sc(c->env.prof, EGX_STAGE_CROSS_CORR, 1, c->stream);
        launch_cross_corr(c->corr, w.xraw, m, mpad, c->x_mean, c->x_std, w.xn, m, mpad, c->d, c->terms, c->nterms, nullptr,
                          nullptr, c->basis_i, c->basis_j, 0, 0.0, 1.0, w.K, mpad, nullptr, c->stream);
    }
    {
        GemmArgs g;                       // K -= rt^T rt  (rows of Y are the columns of rt)
        g.C = w.K;
        g.ldc = mpad;
        g.A = c->Y;
        g.lda = c->npad;
        g.B = c->Y;
        g.ldb = c->npad;
        g.K = c->npad;
        g.tri = 0;
        g.Mt = g.Nt = mpad / EGX_NB;
        StageScope sc(c->env.prof, EGX_STAGE_SYRK_GEMM, 1, c->stream);
        launch_gemm_nt_sub(g, c->stream);
    }
    launch_cov_finish(w.K, mpad, m, mpad, w.U, c->p, c->sigma2_scaled, c->stream);
    return EGX_OK;
}

}  // namespace

extern "C" int egx_gp_covariance(egx_gp_ctx* c, const double* x, int m, double* cov) try {
    if (!c || !x || !cov) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    CovWork w;
    int st = conditional_cov_dev(c, x, m, w);
    if (st != EGX_OK) return st;
    EGX_CUDA_TRY(cudaMemcpy2DAsync(cov, static_cast<size_t>(m) * sizeof(double), w.K,
                                   static_cast<size_t>(w.mpad) * sizeof(double), static_cast<size_t>(m) * sizeof(double), m,
                                   cudaMemcpyDeviceToHost, c->stream));
    EGX_CUDA_TRY(cudaStreamSynchronize(c->stream));
    EGX_CUDA_TRY(cudaGetLastError());
    resolve_profile(c);
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_sample(egx_gp_ctx* c, const double* x, int m, const double* z, int n_traj, int method,
                             double* out) try {
    if (!c || !x || !z || !out || n_traj < 1) return EGX_INVALID_VALUE;
    if (method != EGX_SAMPLE_CHOLESKY && method != EGX_SAMPLE_EIGENVALUES) {
        egx_set_error("unknown sampling method %d", method);
        return EGX_INVALID_VALUE;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    CovWork w;
    int st = conditional_cov_dev(c, x, m, w);
    if (st != EGX_OK) return st;
    st = sample_from_covariance(c->env, c->stream, w.K, m, w.mpad, w.mean, z, n_traj, method, out);
    if (st != EGX_OK) return st;
    EGX_CUDA_TRY(cudaGetLastError());
    resolve_profile(c);
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_predict_valvar_dev(egx_gp_ctx* c, const double* x_dev, int m, double* y_dev, double* var_dev) try {
    if (!c) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    return predict_impl(c, x_dev, m, y_dev, var_dev, true);
}
EGX_ABI_CATCH

extern "C" int egx_gp_correlation_matrix(egx_gp_ctx* c, const double* theta, double* r) try {
    if (!c || !theta || !r) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    c->trained = false;
    int st = assemble(c, theta);
    if (st != EGX_OK) return st;
    EGX_CUDA_TRY(cudaStreamSynchronize(c->stream));
    EGX_CUDA_TRY(cudaGetLastError());
    resolve_profile(c);
    EGX_CUDA_TRY(cudaMemcpy2D(r, c->n * sizeof(double), c->M, c->ld * sizeof(double), c->n * sizeof(double), c->n,
                              cudaMemcpyDeviceToHost));
    for (int i = 0; i < c->n; ++i)
        for (int j = i + 1; j < c->n; ++j) r[static_cast<size_t>(i) * c->n + j] = r[static_cast<size_t>(j) * c->n + i];
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_cross_correlation(egx_gp_ctx* c, const double* x, int m, double* out) try {
    if (!c || !x || !out || m < 1) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->trained) {
        egx_set_error("egx_gp_cross_correlation before finalize");
        return EGX_INVALID_VALUE;
    }
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    const int mb = std::min(round_up(m, EGX_NB), PREDICT_CHUNK);
    int st = ensure_predict_buffers(c, mb);
    if (st != EGX_OK) return st;
    double* cdev = nullptr;
    EGX_CUDA_TRY(egx_dev_malloc(&cdev, static_cast<size_t>(mb) * c->n * sizeof(double)));
    for (int i0 = 0; i0 < m; i0 += mb) {
        const int mc = std::min(mb, m - i0);
        cudaMemcpyAsync(c->xchunk, x + static_cast<long>(i0) * c->d, static_cast<size_t>(mc) * c->d * sizeof(double),
                        cudaMemcpyHostToDevice, c->stream);
        st = predict_chunk_dev(c, c->xchunk, mc, nullptr, nullptr, cdev);
        if (st != EGX_OK) break;
        cudaMemcpyAsync(out + static_cast<long>(i0) * c->n, cdev, static_cast<size_t>(mc) * c->n * sizeof(double),
                        cudaMemcpyDeviceToHost, c->stream);
        cudaStreamSynchronize(c->stream);
    }
    egx_dev_free(cdev);
    if (st != EGX_OK) return st;
    EGX_CUDA_TRY(cudaGetLastError());
    resolve_profile(c);
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_gp_set_profiling(egx_gp_ctx* c, int enabled) try {
    if (!c) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    c->env.prof.on = enabled != 0;
    for (egx_gp_ctx* r : c->replicas) r->env.prof.on = c->env.prof.on;
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" int egx_gp_reset_profile(egx_gp_ctx* c) try {
    if (!c) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    c->env.prof.reset();
    for (egx_gp_ctx* r : c->replicas) r->env.prof.reset();
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" int egx_gp_get_profile(egx_gp_ctx* c, double* ms, long long* launches) try {
    if (!c) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    for (int i = 0; i < EGX_NUM_STAGES; ++i) {
        if (ms) ms[i] = c->env.prof.ms[i];
        if (launches) launches[i] = c->env.prof.launches[i];
        for (egx_gp_ctx* r : c->replicas) {
            if (ms) ms[i] += r->env.prof.ms[i];
            if (launches) launches[i] += r->env.prof.launches[i];
        }
    }
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" int egx_gp_timer_start(egx_gp_ctx* c) try {
    if (!c) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    if (!c->timer_a) {
        EGX_CUDA_TRY(cudaEventCreate(&c->timer_a));
        EGX_CUDA_TRY(cudaEventCreate(&c->timer_b));
    }
    EGX_CUDA_TRY(cudaEventRecord(c->timer_a, c->stream));
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" int egx_gp_timer_stop(egx_gp_ctx* c, double* elapsed_ms) try {
    if (!c || !elapsed_ms || !c->timer_a) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    for (egx_gp_ctx* r : c->replicas) {   // the replicas' streams end before the stop event
        EGX_CUDA_TRY(cudaEventRecord(r->env.ev_join, r->stream));
        EGX_CUDA_TRY(cudaStreamWaitEvent(c->stream, r->env.ev_join, 0));
    }
    EGX_CUDA_TRY(cudaEventRecord(c->timer_b, c->stream));
    EGX_CUDA_TRY(cudaEventSynchronize(c->timer_b));
    float ms = 0.f;
    EGX_CUDA_TRY(cudaEventElapsedTime(&ms, c->timer_a, c->timer_b));
    *elapsed_ms = ms;
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" int egx_gp_set_lookahead(egx_gp_ctx* c, int enabled) try {
    if (!c) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    c->env.lookahead = enabled != 0;
    for (egx_gp_ctx* r : c->replicas) r->env.lookahead = c->env.lookahead;
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" int egx_gp_set_force_blocked(egx_gp_ctx* c, int enabled) try {
    if (!c) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    c->force_blocked = enabled != 0;
    return EGX_OK;
}
EGX_ABI_CATCH
