// Sparse GP (FITC / VFE) device context and its C ABI (egx_sgp_*).
//
// Reference being replaced: crates/gp/src/sparse_algorithm.rs
//   compute_k :676-691, fitc :695-765, vfe :769-830, predict :237-241, predict_var :245-257.
// No input normalisation, zero mean (as in the reference).  One likelihood evaluation:
//   K1  Kmm = sigma2 r(Z,Z) + nugget I                      (M x M)
//   sweep: U = chol(Kmm)
//   per chunk of <= 8192 points:  K2  Knm = sigma2 r(X, Z)  (chunk x M)
//                                 sweep: Vt = Knm U^-T       (rows = points)
//                                 row stats: nu_i, beta_i, sum ln nu, sum beta y^2 ...
//                                 W = (sqrt(beta) . Vt)^T ; t += Vt^T (beta y)
//                                 split-K SYRK: partial[s] += W_s W_s^T          (DMMA)
//   A = I + sum_s partial[s] ; row M = t ; sweep: L = chol(A), b = L^-1 t (fused as appended row)
//   likelihood from (sum ln nu, sum ln L_ii, ...) -- natural logs.
// The explicit triangular inverses Ui, Li of the reference are never formed on the hot path;
// predict_var evaluates k^T inv k as |U^-1 k|^2 -/+ |L^-1 U^-1 k|^2 with two multi-RHS sweeps.
#include <cmath>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "sweep.cuh"
#include "../../include/egobox_gpu.h"
#include "abi_guard.h"

void launch_sgp_rowstats(const double* Y, long ldy, int mc, int mpad_rows, int Mpad, const double* yv, int method,
                         double sigma2, double noise, double beta_const, double* sqrtb, double* by, double* scal,
                         cudaStream_t s);
void launch_sgp_scale_transpose(const double* Y, long ldy, int rows, int Mpad, const double* sqrtb, const double* by,
                                double* W, long ldw, double* tvec, cudaStream_t s);
void launch_sgp_reduce_partials(const double* partial, int splits, long stride, double* A, long ld, int Mpad,
                                const double* tvec, cudaStream_t s);
void launch_sgp_final(const double* A, long ld, int M, int Mpad, const double* scal, int method, int N, double sigma2,
                      double beta_const, const int* info_u, const int* info_l, double* out, cudaStream_t s);
void launch_sgp_row_sumsq(const double* Y, long ldy, int m, int Mpad, double* out, cudaStream_t s);
void launch_sgp_var(const double* s1, const double* s2, int m, int method, double sigma2, double noise, double* var,
                    cudaStream_t s);

namespace {
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
constexpr int SGP_CHUNK = 8192;
constexpr int SGP_SPLITS = 4;

__global__ void set_identity_kernel(double* A, long ld, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[static_cast<long>(i) * ld + i] = 1.0;
}
}  // namespace

struct egx_sgp_ctx {
    int device = 0, corr = 0, method = 0;
    int N = 0, d = 0, M = 0, h = 0, Mpad = 0, chunk = 0;
    double nugget = 0.0;
    std::vector<double> w_star;
    SweepEnv env;
    cudaStream_t stream = nullptr;
    double *X = nullptr, *Z = nullptr, *yv = nullptr, *zeros_d = nullptr, *ones_d = nullptr;
    CorrTerm *terms = nullptr, *terms_h = nullptr;
    int nterms = 0;
    double *F1 = nullptr, *F2 = nullptr, *Dinv1 = nullptr, *Dinv2 = nullptr;
    int *info1 = nullptr, *info2 = nullptr;
    double *Y = nullptr, *W = nullptr, *sqrtb = nullptr, *by = nullptr, *partial = nullptr, *tvec = nullptr;
    double *scal = nullptr, *out = nullptr, *out_h = nullptr;
    int8_t* Lsl1 = nullptr;       // tcgen05 path of the solves Vt = Knm U^-T: digit slices / row scales of the block rows of U below every
    double* Lsc1 = nullptr;       // column pair (rebuilt after every factorisation of Kmm), offsets per pair
    std::vector<long> Lsl1_off, Lsc1_off;
    bool Lsl1_ready = false;
    int8_t* ozS = nullptr;        // tcgen05 path of the split-K W W^T: int8 digit slices of every K panel of 256 points of a chunk ...
    double* ozScale = nullptr;    // ... and their row scales (kernels_ozaki.cu, launch_ozaki_slice_panels)
    double *vec = nullptr, *s1 = nullptr, *s2 = nullptr, *xchunk = nullptr, *ychunk = nullptr, *vchunk = nullptr;
    bool trained = false;
    double sigma2 = 0.0, noise = 0.0;
    std::mutex mu;
};

namespace {

FactorRef fref(egx_sgp_ctx* c, int which) {
    FactorRef f;
    f.M = which == 1 ? c->F1 : c->F2;
    f.ld = c->Mpad;
    f.T = c->Mpad / EGX_NB;
    f.qpad = which == 1 ? 0 : EGX_NB;
    f.Dinv = which == 1 ? c->Dinv1 : c->Dinv2;
    f.info = which == 1 ? c->info1 : c->info2;
    if (which == 1 && c->Lsl1_ready) {
        f.Lsl = c->Lsl1;
        f.Lsc = c->Lsc1;
        f.Lsl_off = c->Lsl1_off.data();
        f.Lsc_off = c->Lsc1_off.data();
    }
    return f;
}

void free_sgp(egx_sgp_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (double* p : {c->X, c->Z, c->yv, c->zeros_d, c->ones_d, c->F1, c->F2, c->Dinv1, c->Dinv2, c->Y, c->W, c->sqrtb,
                      c->by, c->partial, c->tvec, c->scal, c->out, c->vec, c->s1, c->s2, c->xchunk, c->ychunk,
                      c->vchunk})
        cudaFree(p);
    cudaFree(c->ozS);
    cudaFree(c->ozScale);
    cudaFree(c->Lsl1);
    cudaFree(c->Lsc1);
    cudaFree(c->terms);
    cudaFree(c->info1);
    cudaFree(c->info2);
    cudaFreeHost(c->terms_h);
    cudaFreeHost(c->out_h);
    c->env.destroy();
    delete c;
}

int sgp_evaluate(egx_sgp_ctx* c, const double* theta, double sigma2, double noise, double* lik) {
    *lik = NAN;
    c->trained = false;
    for (int l = 0; l < c->h; ++l)
        if (std::isnan(theta[l])) {
            egx_set_error("theta[%d] is NaN", l);
            return EGX_INVALID_VALUE;
        }
    if (std::isnan(sigma2) || std::isnan(noise)) {
        egx_set_error("sigma2 / noise is NaN");
        return EGX_INVALID_VALUE;
    }
    cudaStream_t s = c->stream;
    const int Mpad = c->Mpad;
    c->nterms = egx_fill_terms(c->corr, c->d, c->h, c->w_star.data(), theta, c->terms_h);
    if (c->nterms > 0)
        EGX_CUDA_TRY(cudaMemcpyAsync(c->terms, c->terms_h, c->nterms * sizeof(CorrTerm), cudaMemcpyHostToDevice, s));
    EGX_CUDA_TRY(cudaMemsetAsync(c->info1, 0, sizeof(int), s));
    EGX_CUDA_TRY(cudaMemsetAsync(c->info2, 0, sizeof(int), s));
    EGX_CUDA_TRY(cudaMemsetAsync(c->scal, 0, 4 * sizeof(double), s));
    EGX_CUDA_TRY(cudaMemsetAsync(c->tvec, 0, Mpad * sizeof(double), s));
    EGX_CUDA_TRY(cudaMemsetAsync(c->partial, 0, static_cast<size_t>(SGP_SPLITS) * Mpad * Mpad * sizeof(double), s));
    // Kmm = sigma2 r(Z, Z) + nugget I      (fitc :706 / vfe :783)
    {
        StageScope sc(c->env.prof, EGX_STAGE_CORR_BUILD, 1, s);
        launch_corr_build(c->corr, c->Z, c->M, Mpad, c->d, c->terms, c->nterms, c->F1, Mpad, sigma2 + c->nugget, s,
                          sigma2);
    }
    c->Lsl1_ready = false;
    blocked_sweep(c->env, fref(c, 1), true, nullptr, 0, 0, 0);          // U = chol(Kmm)
    if (c->Lsl1 != nullptr) {
        // slices of the block rows of U below every column pair: the B operand of the tcgen05 solve updates of all chunks
        const int T = Mpad / EGX_NB;
        for (int k = 0; k + 2 < T; k += 2) {
            StageScope sc(c->env.prof, EGX_STAGE_OZAKI_SLICE, 2, s);
            launch_ozaki_slice(c->F1 + static_cast<long>(k + 2) * EGX_NB * Mpad + static_cast<long>(k) * EGX_NB, Mpad, (T - k - 2) * EGX_NB,
                               c->Lsc1 + c->Lsc1_off[k >> 1], c->Lsl1 + c->Lsl1_off[k >> 1], s);
        }
        c->Lsl1_ready = true;
    }
    const double beta_const = 1.0 / std::max(noise, c->nugget);           // vfe :796
    for (int i0 = 0; i0 < c->N; i0 += c->chunk) {
        const int mc = std::min(c->chunk, c->N - i0);
        const int mpad = round_up(mc, EGX_NB);
        {
            StageScope sc(c->env.prof, EGX_STAGE_CROSS_CORR, 1, s);
            launch_cross_corr(c->corr, c->X + static_cast<long>(i0) * c->d, mc, mpad, c->zeros_d, c->ones_d, c->Z, c->M,
                              Mpad, c->d, c->terms, c->nterms, nullptr, nullptr, nullptr, nullptr, 0, 0.0, 1.0, c->Y,
                              Mpad, nullptr, s, sigma2);
        }
        blocked_sweep(c->env, fref(c, 1), false, c->Y, Mpad, mpad / EGX_NB, mpad / 64);   // Vt = Knm U^-T
        {
            StageScope sc(c->env.prof, EGX_STAGE_VAR_FINISH, 2, s);
            launch_sgp_rowstats(c->Y, Mpad, mc, mpad, Mpad, c->yv + i0, c->method, sigma2, noise, beta_const, c->sqrtb,
                                c->by, c->scal, s);
            launch_sgp_scale_transpose(c->Y, Mpad, mpad, Mpad, c->sqrtb, c->by, c->W, c->chunk, c->tvec, s);
        }
        // A += W W^T over the mpad points of the chunk (fitc :727-731 / vfe :798).  tcgen05: every K panel of 256 points is
        // sliced into int8 digits on its own and ONE launch runs (tile, panel) tasks whose fp64 results meet in L2 through the
        // TMA reduction (kernels_ozaki.cu, launch_ozaki_syrk_add_panels) -- no split-K partial buffers; the DMMA kernel with
        // 4-way split-K stays for small problems and for EGX_OZAKI=0.
        const int tri = Mpad / EGX_NB, kp = (mpad + 255) / 256;
        bool done = false;
        if (c->ozS != nullptr && static_cast<long>(tri) * (tri + 1) / 2 * kp >= 64) {
            if (mpad % 256 != 0)        // the last panel is half empty: its 128 extra columns of W must be zeros
                EGX_CUDA_TRY(cudaMemset2DAsync(c->W + mpad, static_cast<size_t>(c->chunk) * sizeof(double), 0, 128 * sizeof(double),
                                               Mpad, s));
            {
                StageScope sc(c->env.prof, EGX_STAGE_OZAKI_SLICE, 2, s);
                launch_ozaki_slice_panels(c->W, c->chunk, Mpad, kp, c->ozScale, c->ozS, s);
            }
            StageScope sc(c->env.prof, EGX_STAGE_OZAKI_SYRK, 1, s);
            done = launch_ozaki_syrk_add_panels(c->partial, Mpad, c->ozS, c->ozScale, tri, kp, s);
        }
        if (!done) {
            GemmArgs g;
            g.C = c->partial;
            g.ldc = Mpad;
            g.A = c->W;
            g.lda = c->chunk;
            g.B = c->W;
            g.ldb = c->chunk;
            g.tri = Mpad / EGX_NB;
            g.Mt = g.tri;
            g.Nt = g.tri;
            g.add = 1;
            g.splits = SGP_SPLITS;
            g.K = mpad / SGP_SPLITS;
            g.split_c_stride = static_cast<long>(Mpad) * Mpad;
            StageScope sc(c->env.prof, EGX_STAGE_SYRK_GEMM, 1, s);
            launch_gemm_nt_sub(g, s);
        }
    }
    {
        StageScope sc(c->env.prof, EGX_STAGE_GLS, 1, s);
        launch_sgp_reduce_partials(c->partial, SGP_SPLITS, static_cast<long>(Mpad) * Mpad, c->F2, Mpad, Mpad, c->tvec, s);
    }
    blocked_sweep(c->env, fref(c, 2), true, nullptr, 0, 0, 0);          // L = chol(A), row Mpad -> b
    {
        StageScope sc(c->env.prof, EGX_STAGE_GLS, 1, s);
        launch_sgp_final(c->F2, Mpad, c->M, Mpad, c->scal, c->method, c->N, sigma2, beta_const, c->info1, c->info2,
                         c->out, s);
    }
    EGX_CUDA_TRY(cudaMemcpyAsync(c->out_h, c->out, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    EGX_CUDA_TRY(cudaStreamSynchronize(s));
    EGX_CUDA_TRY(cudaGetLastError());
    c->env.prof.resolve();
    if (c->out_h[1] != 0.0 || c->out_h[2] != 0.0) {
        egx_set_error("sparse GP: %s is not positive definite", c->out_h[1] != 0.0 ? "Kmm" : "I + V diag(beta) V^T");
        return EGX_NOT_POSITIVE_DEFINITE;
    }
    *lik = c->out_h[0];
    return EGX_OK;
}

int sgp_predict_impl(egx_sgp_ctx* c, const double* x, int m, double* y, double* var) {
    if (!c->trained) {
        egx_set_error("sparse GP predict* before egx_sgp_finalize");
        return EGX_INVALID_VALUE;
    }
    if (m <= 0) return EGX_OK;
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int Mpad = c->Mpad;
    for (int i0 = 0; i0 < m; i0 += c->chunk) {
        const int mc = std::min(c->chunk, m - i0);
        const int mpad = round_up(mc, EGX_NB);
        EGX_CUDA_TRY(cudaMemcpyAsync(c->xchunk, x + static_cast<long>(i0) * c->d,
                                     static_cast<size_t>(mc) * c->d * sizeof(double), cudaMemcpyHostToDevice, s));
        {
            StageScope sc(c->env.prof, EGX_STAGE_CROSS_CORR, 1, s);
            launch_cross_corr(c->corr, c->xchunk, mc, mpad, c->zeros_d, c->ones_d, c->Z, c->M, Mpad, c->d, c->terms,
                              c->nterms, c->vec, nullptr, nullptr, nullptr, 0, 0.0, 1.0, var ? c->Y : nullptr, Mpad,
                              y ? c->ychunk : nullptr, s, c->sigma2);
        }
        if (var) {
            blocked_sweep(c->env, fref(c, 1), false, c->Y, Mpad, mpad / EGX_NB, mpad / 64);
            launch_sgp_row_sumsq(c->Y, Mpad, mc, Mpad, c->s1, s);
            blocked_sweep(c->env, fref(c, 2), false, c->Y, Mpad, mpad / EGX_NB, mpad / 64);
            launch_sgp_row_sumsq(c->Y, Mpad, mc, Mpad, c->s2, s);
            StageScope sc(c->env.prof, EGX_STAGE_VAR_FINISH, 3, s);
            launch_sgp_var(c->s1, c->s2, mc, c->method, c->sigma2, c->noise, c->vchunk, s);
            EGX_CUDA_TRY(cudaMemcpyAsync(var + i0, c->vchunk, mc * sizeof(double), cudaMemcpyDeviceToHost, s));
        }
        if (y) EGX_CUDA_TRY(cudaMemcpyAsync(y + i0, c->ychunk, mc * sizeof(double), cudaMemcpyDeviceToHost, s));
        EGX_CUDA_TRY(cudaStreamSynchronize(s));
    }
    EGX_CUDA_TRY(cudaGetLastError());
    c->env.prof.resolve();
    return EGX_OK;
}

}  // namespace

extern "C" int egx_sgp_create(egx_sgp_ctx** out, int device, int corr, int method, const double* x, int n, int d,
                              const double* y, const double* z, int m, const double* w_star, int h, double nugget) try {
    if (!out) return EGX_INVALID_VALUE;
    *out = nullptr;
    if (!x || !y || !z || !w_star || n < 1 || d < 1 || m < 1 || h < 1 || h > d || corr < 0 || corr > 3 || method < 0 ||
        method > 1) {
        egx_set_error("egx_sgp_create: invalid argument");
        return EGX_INVALID_VALUE;
    }
    if (egx_device_count() <= device || device < 0) {
        egx_set_error("egx_sgp_create: CUDA device %d not available (no CPU fallback exists)", device);
        return EGX_CUDA_ERROR;
    }
    EGX_CUDA_TRY(cudaSetDevice(device));
    egx_sgp_ctx* c = new egx_sgp_ctx();
    c->device = device;
    c->corr = corr;
    c->method = method;
    c->N = n;
    c->d = d;
    c->M = m;
    c->h = h;
    c->nugget = nugget;
    c->Mpad = round_up(m, EGX_NB);
    // points per chunk: a multiple of 256 (the K panels of the split-K product); EGX_SGP_CHUNK overrides
    // default: TWO full waves of 64-row solve slabs (148 SMs: 18 944 points).  Measured at N = 1e5, M = 1024
    // (profiles/r02/y7_sgp.txt): 8192 -> 8.56 ms per evaluation (128 slabs on 148 SMs, 13 chunks), 9472 -> 7.81, 18944 -> 7.05
    int chunk_pts = SGP_CHUNK;
    {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0)
            chunk_pts = std::max(SGP_CHUNK, 2 * sms * 64 / 256 * 256);
    }
    if (const char* e = getenv("EGX_SGP_CHUNK")) chunk_pts = std::max(256, atoi(e) / 256 * 256);
    c->chunk = std::min(round_up(n, EGX_NB), chunk_pts);
    c->w_star.assign(w_star, w_star + static_cast<size_t>(d) * h);
    const int Mpad = c->Mpad, CH = c->chunk;
#define SGP_TRY(expr)                                                        \
    do {                                                                     \
        cudaError_t e__ = (expr);                                            \
        if (e__ != cudaSuccess) {                                            \
            egx_set_error("%s failed: %s", #expr, cudaGetErrorString(e__));  \
            free_sgp(c);                                                     \
            return EGX_CUDA_ERROR;                                           \
        }                                                                    \
    } while (0)
    if (c->env.init(Mpad / EGX_NB) != EGX_OK || c->env.ensure_panel_rows(std::max(CH, Mpad + EGX_NB)) != EGX_OK) {
        free_sgp(c);
        return EGX_CUDA_ERROR;
    }
    c->stream = c->env.sb;
    cudaStream_t s = c->stream;
    SGP_TRY(cudaMalloc(&c->X, static_cast<size_t>(n) * d * sizeof(double)));
    SGP_TRY(cudaMemcpyAsync(c->X, x, static_cast<size_t>(n) * d * sizeof(double), cudaMemcpyHostToDevice, s));
    SGP_TRY(cudaMalloc(&c->Z, static_cast<size_t>(Mpad) * d * sizeof(double)));
    SGP_TRY(cudaMemsetAsync(c->Z, 0, static_cast<size_t>(Mpad) * d * sizeof(double), s));
    SGP_TRY(cudaMemcpyAsync(c->Z, z, static_cast<size_t>(m) * d * sizeof(double), cudaMemcpyHostToDevice, s));
    SGP_TRY(cudaMalloc(&c->yv, n * sizeof(double)));
    SGP_TRY(cudaMemcpyAsync(c->yv, y, n * sizeof(double), cudaMemcpyHostToDevice, s));
    std::vector<double> ones(d, 1.0);
    SGP_TRY(cudaMalloc(&c->zeros_d, d * sizeof(double)));
    SGP_TRY(cudaMemsetAsync(c->zeros_d, 0, d * sizeof(double), s));
    SGP_TRY(cudaMalloc(&c->ones_d, d * sizeof(double)));
    SGP_TRY(cudaMemcpyAsync(c->ones_d, ones.data(), d * sizeof(double), cudaMemcpyHostToDevice, s));
    SGP_TRY(cudaMalloc(&c->terms, static_cast<size_t>(d) * h * sizeof(CorrTerm)));
    SGP_TRY(cudaMallocHost(&c->terms_h, static_cast<size_t>(d) * h * sizeof(CorrTerm)));
    SGP_TRY(cudaMalloc(&c->F1, static_cast<size_t>(Mpad) * Mpad * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->F2, static_cast<size_t>(Mpad + EGX_NB) * Mpad * sizeof(double)));
    SGP_TRY(cudaMemsetAsync(c->F2, 0, static_cast<size_t>(Mpad + EGX_NB) * Mpad * sizeof(double), s));
    SGP_TRY(cudaMalloc(&c->Dinv1, static_cast<size_t>(Mpad / EGX_NB) * 4096 * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->Dinv2, static_cast<size_t>(Mpad / EGX_NB) * 4096 * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->info1, sizeof(int)));
    SGP_TRY(cudaMalloc(&c->info2, sizeof(int)));
    SGP_TRY(cudaMalloc(&c->Y, static_cast<size_t>(CH) * Mpad * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->W, static_cast<size_t>(Mpad) * CH * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->sqrtb, CH * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->by, CH * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->partial, static_cast<size_t>(SGP_SPLITS) * Mpad * Mpad * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->tvec, Mpad * sizeof(double)));
    if (c->env.ozaki && Mpad / EGX_NB >= 4) {
        const int T = Mpad / EGX_NB;
        long bytes = 0, rows = 0;
        c->Lsl1_off.assign((T + 1) / 2, 0);
        c->Lsc1_off.assign((T + 1) / 2, 0);
        for (int k = 0; k + 2 < T; k += 2) {
            c->Lsl1_off[k >> 1] = bytes;
            c->Lsc1_off[k >> 1] = rows;
            bytes += static_cast<long>(ozaki_slice_bytes(static_cast<long>(T - k - 2) * EGX_NB));
            rows += static_cast<long>(T - k - 2) * EGX_NB;
        }
        SGP_TRY(cudaMalloc(&c->Lsl1, static_cast<size_t>(bytes)));
        SGP_TRY(cudaMalloc(&c->Lsc1, static_cast<size_t>(rows) * sizeof(double)));
        c->env.ozaki_min_tri_solve = 2;
    }
    if (c->env.ozaki) {
        const int kp_max = (CH + 255) / 256;
        SGP_TRY(cudaMalloc(&c->ozS, ozaki_slice_bytes(Mpad) * kp_max));
        SGP_TRY(cudaMalloc(&c->ozScale, static_cast<size_t>(kp_max) * Mpad * sizeof(double)));
    }
    SGP_TRY(cudaMalloc(&c->scal, 4 * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->out, 3 * sizeof(double)));
    SGP_TRY(cudaMallocHost(&c->out_h, 3 * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->vec, Mpad * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->s1, CH * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->s2, CH * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->xchunk, static_cast<size_t>(CH) * d * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->ychunk, CH * sizeof(double)));
    SGP_TRY(cudaMalloc(&c->vchunk, CH * sizeof(double)));
    SGP_TRY(cudaStreamSynchronize(s));
#undef SGP_TRY
    *out = c;
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" void egx_sgp_destroy(egx_sgp_ctx* c) { free_sgp(c); }

extern "C" int egx_sgp_reduced_likelihood(egx_sgp_ctx* c, const double* theta, double sigma2, double noise,
                                          double* lik) try {
    if (!c || !theta || !lik) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    return sgp_evaluate(c, theta, sigma2, noise, lik);
}
EGX_ABI_CATCH

extern "C" int egx_sgp_finalize(egx_sgp_ctx* c, const double* theta, double sigma2, double noise, double* lik,
                                double* w_vec, double* w_inv) try {
    if (!c || !theta) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    double v = NAN;
    int st = sgp_evaluate(c, theta, sigma2, noise, &v);
    if (lik) *lik = v;
    if (st != EGX_OK) return st;
    cudaStream_t s = c->stream;
    const int Mpad = c->Mpad;
    // w_vec = (Li Ui)^T b = U^-T L^-T b          (fitc :760-762, vfe :824-826)
    EGX_CUDA_TRY(cudaMemcpyAsync(c->vec, c->F2 + static_cast<long>(Mpad) * Mpad, Mpad * sizeof(double),
                                 cudaMemcpyDeviceToDevice, s));
    backsolve_vector(c->env, fref(c, 2), c->vec);
    backsolve_vector(c->env, fref(c, 1), c->vec);
    if (w_vec) EGX_CUDA_TRY(cudaMemcpyAsync(w_vec, c->vec, c->M * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (w_inv) {
        // explicit Woodbury inverse, only on request (serialisation / parity):
        //   Uit = U^-T (rows of I solved against U), Wm = Uit L^-T ; inv = Uit Uit^T -/+ Wm Wm^T
        double *Uit = nullptr, *Wm = nullptr, *Cinv = nullptr;
        const size_t mm = static_cast<size_t>(Mpad) * Mpad * sizeof(double);
        EGX_CUDA_TRY(cudaMalloc(&Uit, mm));
        EGX_CUDA_TRY(cudaMalloc(&Wm, mm));
        EGX_CUDA_TRY(cudaMalloc(&Cinv, mm));
        cudaMemsetAsync(Uit, 0, mm, s);
        cudaMemsetAsync(Cinv, 0, mm, s);
        set_identity_kernel<<<(Mpad + 255) / 256, 256, 0, s>>>(Uit, Mpad, Mpad);
        blocked_sweep(c->env, fref(c, 1), false, Uit, Mpad, Mpad / EGX_NB, Mpad / 64);
        cudaMemcpyAsync(Wm, Uit, mm, cudaMemcpyDeviceToDevice, s);
        blocked_sweep(c->env, fref(c, 2), false, Wm, Mpad, Mpad / EGX_NB, Mpad / 64);
        GemmArgs g;
        g.C = Cinv;
        g.ldc = Mpad;
        g.lda = g.ldb = Mpad;
        g.tri = 0;
        g.Mt = g.Nt = Mpad / EGX_NB;
        g.K = Mpad;
        g.A = g.B = Uit;
        g.add = 1;
        launch_gemm_nt_sub(g, s);
        g.A = g.B = Wm;
        g.add = (c->method == 0) ? 0 : 1;
        launch_gemm_nt_sub(g, s);
        cudaMemcpy2DAsync(w_inv, c->M * sizeof(double), Cinv, Mpad * sizeof(double), c->M * sizeof(double), c->M,
                          cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        cudaFree(Uit);
        cudaFree(Wm);
        cudaFree(Cinv);
    }
    EGX_CUDA_TRY(cudaStreamSynchronize(s));
    EGX_CUDA_TRY(cudaGetLastError());
    c->env.prof.resolve();
    c->sigma2 = sigma2;
    c->noise = noise;
    c->trained = true;
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_sgp_predict(egx_sgp_ctx* c, const double* x, int m, double* y) try {
    if (!c || !x || !y) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    return sgp_predict_impl(c, x, m, y, nullptr);
}
EGX_ABI_CATCH
extern "C" int egx_sgp_predict_var(egx_sgp_ctx* c, const double* x, int m, double* var) try {
    if (!c || !x || !var) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    return sgp_predict_impl(c, x, m, nullptr, var);
}
EGX_ABI_CATCH
// Trajectories of the sparse GP (sparse_algorithm.rs:338-364): mean = predict(x), covariance = sigma2 r(x, x) -- the PRIOR
// covariance, as the reference's `_sample` takes it from `compute_k(x, x, ..)` --, then the decomposition shared with the dense GP.
extern "C" int egx_sgp_sample(egx_sgp_ctx* c, const double* x, int m, const double* z, int n_traj, int method, double* out) try {
    if (!c || !x || !z || !out || n_traj < 1) return EGX_INVALID_VALUE;
    if (method != EGX_SAMPLE_CHOLESKY && method != EGX_SAMPLE_EIGENVALUES) {
        egx_set_error("unknown sampling method %d", method);
        return EGX_INVALID_VALUE;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->trained) {
        egx_set_error("sparse GP sample before egx_sgp_finalize");
        return EGX_INVALID_VALUE;
    }
    if (m < 1 || m > 8192) {
        egx_set_error("sparse GP sample: 1 <= number of points <= 8192 (got %d)", m);
        return EGX_INVALID_VALUE;
    }
    std::vector<double> mean_h(m);
    int st = sgp_predict_impl(c, x, m, mean_h.data(), nullptr);
    if (st != EGX_OK) return st;
    EGX_CUDA_TRY(cudaSetDevice(c->device));
    cudaStream_t s = c->stream;
    const int mpad = round_up(m, EGX_NB);
    struct Tmp {
        double *xs = nullptr, *K = nullptr, *mean = nullptr;
        ~Tmp() {
            egx_dev_free(xs);
            egx_dev_free(K);
            egx_dev_free(mean);
        }
    } t;
    EGX_CUDA_TRY(egx_dev_malloc(&t.xs, static_cast<size_t>(mpad) * c->d * sizeof(double)));
    EGX_CUDA_TRY(egx_dev_malloc(&t.K, static_cast<size_t>(mpad) * mpad * sizeof(double)));
    EGX_CUDA_TRY(egx_dev_malloc(&t.mean, static_cast<size_t>(mpad) * sizeof(double)));
    EGX_CUDA_TRY(cudaMemsetAsync(t.xs, 0, static_cast<size_t>(mpad) * c->d * sizeof(double), s));
    EGX_CUDA_TRY(cudaMemcpyAsync(t.xs, x, static_cast<size_t>(m) * c->d * sizeof(double), cudaMemcpyHostToDevice, s));
    EGX_CUDA_TRY(cudaMemsetAsync(t.mean, 0, static_cast<size_t>(mpad) * sizeof(double), s));
    EGX_CUDA_TRY(cudaMemcpyAsync(t.mean, mean_h.data(), static_cast<size_t>(m) * sizeof(double), cudaMemcpyHostToDevice, s));
    if (c->env.ensure_panel_rows(mpad) != EGX_OK) return EGX_CUDA_ERROR;
    {
        // K = sigma2 r(x, x) with the kernel weights of the trained theta (still in c->terms after the finalize)
        StageScope sc(c->env.prof, EGX_STAGE_CROSS_CORR, 1, s);
        launch_cross_corr(c->corr, t.xs, m, mpad, c->zeros_d, c->ones_d, t.xs, m, mpad, c->d, c->terms, c->nterms, nullptr, nullptr,
                          nullptr, nullptr, 0, 0.0, 1.0, t.K, mpad, nullptr, s, c->sigma2);
    }
    launch_cov_finish(t.K, mpad, m, mpad, nullptr, 0, 1.0, s);      // identity on the padding
    st = sample_from_covariance(c->env, s, t.K, m, mpad, t.mean, z, n_traj, method, out);
    if (st != EGX_OK) return st;
    EGX_CUDA_TRY(cudaGetLastError());
    c->env.prof.resolve();
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_sgp_set_profiling(egx_sgp_ctx* c, int enabled) try {
    if (!c) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    c->env.prof.on = enabled != 0;
    c->env.prof.reset();
    return EGX_OK;
}
EGX_ABI_CATCH
extern "C" int egx_sgp_get_profile(egx_sgp_ctx* c, double* ms, long long* launches) try {
    if (!c) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(c->mu);
    for (int i = 0; i < EGX_NUM_STAGES; ++i) {
        if (ms) ms[i] = c->env.prof.ms[i];
        if (launches) launches[i] = c->env.prof.launches[i];
    }
    return EGX_OK;
}
EGX_ABI_CATCH
