// Mixture-of-experts prediction on the device: Gaussian-mixture responsibilities fused with the experts'
// batched predictions and the hard / smooth recombination.
//
// Reference being replaced:
//   moe/src/gaussian_mixture.rs:62-83 (new: precisions by Cholesky), :109-116 predict_probas, :122-170
//   probability derivatives, :221-299 log-determinants / log densities / responsibilities, :306-318 predict;
//   moe/src/algorithm.rs:411-423 predict_smooth, :670-685 predict_var_smooth, :691-783 gradients (smooth),
//   :785-877 valvar(+gradients) smooth, :879-1010 the *_hard variants (one expert call PER POINT there).
// Here the points stay on the device: x is uploaded once, the responsibilities are one kernel, every expert
// predicts the whole batch (smooth) or its own compacted subset (hard) through the device-pointer entry points
// of the GP context, and the recombination is one kernel per expert -- one download at the end.
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "../../include/egobox_gpu.h"
#include "abi_guard.h"

extern "C" int egx_gp_predict_valvar_dev(egx_gp_ctx* c, const double* x_dev, int m, double* y_dev, double* var_dev);
extern "C" int egx_gp_predict_gradients_dev(egx_gp_ctx* c, const double* x_dev, int m, double* grad_dev);
extern "C" int egx_gp_predict_var_gradients_dev(egx_gp_ctx* c, const double* x_dev, int m, double* grad_dev);

namespace {

constexpr double kMin10Exp = -307.0;                 // f64::MIN_10_EXP (gaussian_mixture.rs:246)
constexpr double kEps = 2.220446049250313e-16;
constexpr int kChunk = 65536;                        // points per device pass

// log N(x; mu_c, cov_c / heaviside) for cluster c (gaussian_mixture.rs:260-283):
// z = (x - mu) PCs,  PCs = precisions_chol * factor^-1/2 ;  -0.5 (|z|^2 + d ln 2pi) + log_det
__device__ __forceinline__ double log_gauss(const double* __restrict__ x, const double* __restrict__ mu,
                                            const double* __restrict__ pcs, double log_det, int d) {
    double q = 0.0;
    for (int j = 0; j < d; ++j) {
        double z = 0.0;
        for (int i = 0; i <= j; ++i) z += (x[i] - mu[i]) * pcs[i * d + j];     // PCs is upper triangular
        q += z * z;
    }
    return -0.5 * (q + d * 1.8378770664093453) + log_det;
}

// One thread per point.  probas (m x k) receives the responsibilities, labels (m) the arg-max cluster
// (first maximum, like ndarray-stats argmax).  compute_log_prob_resp :231-256 with its clamps.
__global__ void gmx_probas_kernel(const double* __restrict__ x, int m, int d, int k,
                                  const double* __restrict__ logw, const double* __restrict__ means,
                                  const double* __restrict__ pcs, const double* __restrict__ log_det,
                                  double* __restrict__ probas, int* __restrict__ labels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double* xi = x + static_cast<long>(i) * d;
    double* pi = probas + static_cast<long>(i) * k;
    if (k == 1) {                                    // :110-111
        pi[0] = 1.0;
        if (labels) labels[i] = 0;
        return;
    }
    double s = 0.0;
    for (int c = 0; c < k; ++c) {
        const double lp = log_gauss(xi, means + c * d, pcs + static_cast<long>(c) * d * d, log_det[c], d) + logw[c];
        pi[c] = lp;
        s += (lp <= kMin10Exp) ? 0.0 : exp(lp);
    }
    const double lpn = (fabs(s) < kEps) ? 0.0 : log(s);
    int best = 0;
    double bestv = -1.0;
    for (int c = 0; c < k; ++c) {
        const double r = exp(pi[c] - lpn);
        pi[c] = r;
        if (r > bestv) {
            bestv = r;
            best = c;
        }
    }
    if (labels) labels[i] = best;
}

// predict_single_probas_derivatives :122-152 for every point: out (m x k x d).
// u_c = w_c pdf_c, v = sum_c u_c, deriv_c = (x - mu_c) prec_c / heaviside,
// uprime_c = -deriv_c u_c, vprime = sum_c uprime_c, out_c = (uprime_c v - u_c vprime) / v^2.
// work (m x k) holds u.
__global__ void gmx_probas_deriv_kernel(const double* __restrict__ x, int m, int d, int k,
                                        const double* __restrict__ w, const double* __restrict__ means,
                                        const double* __restrict__ pcs, const double* __restrict__ log_det,
                                        const double* __restrict__ precs_h, double* __restrict__ work,
                                        double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double* xi = x + static_cast<long>(i) * d;
    double* ui = work + static_cast<long>(i) * k;
    double* oi = out + static_cast<long>(i) * k * d;
    double v = 0.0;
    for (int c = 0; c < k; ++c) {
        const double* mu = means + c * d;
        const double u = w[c] * exp(log_gauss(xi, mu, pcs + static_cast<long>(c) * d * d, log_det[c], d));
        ui[c] = u;
        v += u;
        const double* pr = precs_h + static_cast<long>(c) * d * d;
        for (int j = 0; j < d; ++j) {
            double dj = 0.0;
            for (int a = 0; a < d; ++a) dj += (xi[a] - mu[a]) * pr[a * d + j];
            oi[c * d + j] = -dj * u;                 // uprime
        }
    }
    const double v2 = v * v;
    for (int j = 0; j < d; ++j) {
        double vp = 0.0;
        for (int c = 0; c < k; ++c) vp += oi[c * d + j];
        for (int c = 0; c < k; ++c) oi[c * d + j] = (oi[c * d + j] * v - ui[c] * vp) / v2;
    }
}

// smooth recombination, one expert at a time in cluster order (algorithm.rs:417-421, 675-683, 785-806):
// y += p_c y_c ; var += p_c^2 var_c
__global__ void moe_accumulate_kernel(const double* __restrict__ probas, int k, int c, int m,
                                      const double* __restrict__ yc, const double* __restrict__ vc,
                                      double* __restrict__ y, double* __restrict__ v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double p = probas[static_cast<long>(i) * k + c];
    if (y) y[i] = __dadd_rn(y[i], __dmul_rn(yc[i], p));
    if (v) v[i] = __dadd_rn(v[i], __dmul_rn(__dmul_rn(vc[i], p), p));
}

// smooth gradients (algorithm.rs:691-783): dy += p_c grad y_c + dp_c y_c ; dvar += p_c^2 grad v_c + 2 p_c dp_c v_c
__global__ void moe_accumulate_grad_kernel(const double* __restrict__ probas, const double* __restrict__ dprobas,
                                           int k, int c, int m, int d, const double* __restrict__ yc,
                                           const double* __restrict__ vc, const double* __restrict__ gyc,
                                           const double* __restrict__ gvc, double* __restrict__ gy,
                                           double* __restrict__ gv) {
    const long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<long>(m) * d) return;
    const int i = static_cast<int>(e / d), j = static_cast<int>(e - static_cast<long>(i) * d);
    const double p = probas[static_cast<long>(i) * k + c];
    const double dp = dprobas[(static_cast<long>(i) * k + c) * d + j];
    if (gy) gy[e] += gyc[e] * p + dp * yc[i];
    if (gv) gv[e] += gvc[e] * p * p + 2.0 * p * dp * vc[i];
}

__global__ void gather_rows_kernel(const double* __restrict__ src, const int* __restrict__ idx, int cnt, int d,
                                   double* __restrict__ dst) {
    const long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<long>(cnt) * d) return;
    const int r = static_cast<int>(e / d), j = static_cast<int>(e - static_cast<long>(r) * d);
    dst[e] = src[static_cast<long>(idx[r]) * d + j];
}
__global__ void scatter_rows_kernel(const double* __restrict__ src, const int* __restrict__ idx, int cnt, int d,
                                    double* __restrict__ dst) {
    const long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<long>(cnt) * d) return;
    const int r = static_cast<int>(e / d), j = static_cast<int>(e - static_cast<long>(r) * d);
    dst[static_cast<long>(idx[r]) * d + j] = src[e];
}

inline int blocks_for(long n, int t = 256) { return static_cast<int>((n + t - 1) / t); }

// lower Cholesky of a d x d SPD matrix (row-major); false if not positive definite
bool host_cholesky(const double* a, int d, std::vector<double>& l) {
    l.assign(static_cast<size_t>(d) * d, 0.0);
    for (int j = 0; j < d; ++j) {
        double s = a[j * d + j];
        for (int t = 0; t < j; ++t) s -= l[j * d + t] * l[j * d + t];
        if (!(s > 0.0)) return false;
        const double ljj = std::sqrt(s);
        l[j * d + j] = ljj;
        for (int i = j + 1; i < d; ++i) {
            double r = a[i * d + j];
            for (int t = 0; t < j; ++t) r -= l[i * d + t] * l[j * d + t];
            l[i * d + j] = r / ljj;
        }
    }
    return true;
}

}  // namespace

struct egx_moe {
    int device = 0, k = 0, d = 0;
    double heaviside = 1.0;
    std::vector<double> weights, means, covariances;
    std::vector<double> prec_chol;          // k x d x d, (L^-1)^T (upper triangular)   :182-205
    std::vector<double> precisions;         // k x d x d, prec_chol prec_chol^T            :208-216
    std::vector<egx_gp_ctx*> experts;       // borrowed
    // device copies (refreshed by upload())
    double *logw = nullptr, *w = nullptr, *mu = nullptr, *pcs = nullptr, *logdet = nullptr, *precs_h = nullptr;
    cudaStream_t stream = nullptr;
    std::mutex mtx;
};

namespace {

int upload(egx_moe* q) {
    const int k = q->k, d = q->d;
    const double f = std::pow(q->heaviside, -0.5);               // :221-227, :266-268
    std::vector<double> logw(k), pcs(q->prec_chol), logdet(k, 0.0), ph(q->precisions);
    for (int c = 0; c < k; ++c) {
        logw[c] = std::log(q->weights[c]);
        for (int i = 0; i < d * d; ++i) pcs[static_cast<size_t>(c) * d * d + i] *= f;
        for (int i = 0; i < d; ++i) logdet[c] += std::log(pcs[static_cast<size_t>(c) * d * d + i * d + i]);
        for (int i = 0; i < d * d; ++i) ph[static_cast<size_t>(c) * d * d + i] /= q->heaviside;   // :129
    }
    EGX_CUDA_TRY(cudaMemcpyAsync(q->logw, logw.data(), k * sizeof(double), cudaMemcpyHostToDevice, q->stream));
    EGX_CUDA_TRY(cudaMemcpyAsync(q->w, q->weights.data(), k * sizeof(double), cudaMemcpyHostToDevice, q->stream));
    EGX_CUDA_TRY(cudaMemcpyAsync(q->mu, q->means.data(), sizeof(double) * k * d, cudaMemcpyHostToDevice, q->stream));
    EGX_CUDA_TRY(cudaMemcpyAsync(q->pcs, pcs.data(), sizeof(double) * k * d * d, cudaMemcpyHostToDevice, q->stream));
    EGX_CUDA_TRY(cudaMemcpyAsync(q->logdet, logdet.data(), k * sizeof(double), cudaMemcpyHostToDevice, q->stream));
    EGX_CUDA_TRY(cudaMemcpyAsync(q->precs_h, ph.data(), sizeof(double) * k * d * d, cudaMemcpyHostToDevice, q->stream));
    EGX_CUDA_TRY(cudaStreamSynchronize(q->stream));
    return EGX_OK;
}

struct DevBuf {
    double* p = nullptr;
    ~DevBuf() { egx_dev_free(p); }
    cudaError_t alloc(size_t n) { return egx_dev_malloc(&p, n * sizeof(double)); }
};
struct DevIdx {
    int* p = nullptr;
    ~DevIdx() { egx_dev_free(p); }
    cudaError_t alloc(size_t n) { return egx_dev_malloc(&p, n * sizeof(int)); }
};

int check_experts(egx_moe* q) {
    for (int c = 0; c < q->k; ++c)
        if (q->experts[c] == nullptr) {
            egx_set_error("egx_moe: expert %d not set", c);
            return EGX_INVALID_VALUE;
        }
    return EGX_OK;
}

// responsibilities (+ labels) of one chunk already on the device
int probas_dev(egx_moe* q, const double* x_dev, int m, double* probas_dev_out, int* labels_dev) {
    gmx_probas_kernel<<<blocks_for(m, 128), 128, 0, q->stream>>>(x_dev, m, q->d, q->k, q->logw, q->mu, q->pcs,
                                                                  q->logdet, probas_dev_out, labels_dev);
    EGX_CUDA_TRY(cudaGetLastError());
    return EGX_OK;
}

// One chunk.  want: bit 0 values, bit 1 variances, bit 2 value gradients, bit 3 variance gradients.
int predict_chunk(egx_moe* q, int recomb, const double* x, int m, double* y, double* var, double* gy, double* gv) {
    const int k = q->k, d = q->d;
    const bool grads = (gy != nullptr) || (gv != nullptr);
    const bool need_y = (y != nullptr) || (grads && recomb == EGX_RECOMB_SMOOTH && gy != nullptr);
    const bool need_v = (var != nullptr) || (grads && recomb == EGX_RECOMB_SMOOTH && gv != nullptr);
    DevBuf X, P, DP, W, Y, V, GY, GV, Yc, Vc, GYc, GVc, Xg;
    DevIdx L, I;
    EGX_CUDA_TRY(X.alloc(static_cast<size_t>(m) * d));
    EGX_CUDA_TRY(P.alloc(static_cast<size_t>(m) * k));
    EGX_CUDA_TRY(L.alloc(m));
    EGX_CUDA_TRY(Yc.alloc(m));
    EGX_CUDA_TRY(Vc.alloc(m));
    if (y || need_y) EGX_CUDA_TRY(Y.alloc(m));
    if (var || need_v) EGX_CUDA_TRY(V.alloc(m));
    if (gy) {
        EGX_CUDA_TRY(GY.alloc(static_cast<size_t>(m) * d));
        EGX_CUDA_TRY(GYc.alloc(static_cast<size_t>(m) * d));
    }
    if (gv) {
        EGX_CUDA_TRY(GV.alloc(static_cast<size_t>(m) * d));
        EGX_CUDA_TRY(GVc.alloc(static_cast<size_t>(m) * d));
    }
    EGX_CUDA_TRY(cudaMemcpyAsync(X.p, x, sizeof(double) * m * d, cudaMemcpyHostToDevice, q->stream));
    int st = probas_dev(q, X.p, m, P.p, L.p);
    if (st != EGX_OK) return st;
    if (Y.p) EGX_CUDA_TRY(cudaMemsetAsync(Y.p, 0, sizeof(double) * m, q->stream));
    if (V.p) EGX_CUDA_TRY(cudaMemsetAsync(V.p, 0, sizeof(double) * m, q->stream));
    if (GY.p) EGX_CUDA_TRY(cudaMemsetAsync(GY.p, 0, sizeof(double) * m * d, q->stream));
    if (GV.p) EGX_CUDA_TRY(cudaMemsetAsync(GV.p, 0, sizeof(double) * m * d, q->stream));

    if (recomb == EGX_RECOMB_SMOOTH) {
        if (grads) {
            EGX_CUDA_TRY(DP.alloc(static_cast<size_t>(m) * k * d));
            EGX_CUDA_TRY(W.alloc(static_cast<size_t>(m) * k));
            gmx_probas_deriv_kernel<<<blocks_for(m, 128), 128, 0, q->stream>>>(X.p, m, d, k, q->w, q->mu, q->pcs,
                                                                               q->logdet, q->precs_h, W.p, DP.p);
            EGX_CUDA_TRY(cudaGetLastError());
        }
        EGX_CUDA_TRY(cudaStreamSynchronize(q->stream));
        for (int c = 0; c < k; ++c) {
            if (need_y || need_v) {
                st = egx_gp_predict_valvar_dev(q->experts[c], X.p, m, need_y ? Yc.p : nullptr, need_v ? Vc.p : nullptr);
                if (st != EGX_OK) return st;
                if (y || var)
                    moe_accumulate_kernel<<<blocks_for(m), 256, 0, q->stream>>>(P.p, k, c, m, Yc.p, Vc.p,
                                                                                y ? Y.p : nullptr, var ? V.p : nullptr);
            }
            if (gy) {
                st = egx_gp_predict_gradients_dev(q->experts[c], X.p, m, GYc.p);
                if (st != EGX_OK) return st;
            }
            if (gv) {
                st = egx_gp_predict_var_gradients_dev(q->experts[c], X.p, m, GVc.p);
                if (st != EGX_OK) return st;
            }
            if (grads)
                moe_accumulate_grad_kernel<<<blocks_for(static_cast<long>(m) * d), 256, 0, q->stream>>>(
                    P.p, DP.p, k, c, m, d, Yc.p, Vc.p, GYc.p, GVc.p, GY.p, GV.p);
            EGX_CUDA_TRY(cudaGetLastError());
            EGX_CUDA_TRY(cudaStreamSynchronize(q->stream));      // Yc / Vc / G*c are reused by the next expert
        }
    } else {
        // hard: every expert predicts the compacted subset of the points it owns (algorithm.rs:879-1010)
        std::vector<int> labels(m);
        EGX_CUDA_TRY(cudaMemcpyAsync(labels.data(), L.p, sizeof(int) * m, cudaMemcpyDeviceToHost, q->stream));
        EGX_CUDA_TRY(cudaStreamSynchronize(q->stream));
        std::vector<std::vector<int>> rows(k);
        for (int i = 0; i < m; ++i) rows[labels[i]].push_back(i);
        EGX_CUDA_TRY(I.alloc(m));
        EGX_CUDA_TRY(Xg.alloc(static_cast<size_t>(m) * d));
        for (int c = 0; c < k; ++c) {
            const int cnt = static_cast<int>(rows[c].size());
            if (cnt == 0) continue;
            EGX_CUDA_TRY(cudaMemcpyAsync(I.p, rows[c].data(), sizeof(int) * cnt, cudaMemcpyHostToDevice, q->stream));
            gather_rows_kernel<<<blocks_for(static_cast<long>(cnt) * d), 256, 0, q->stream>>>(X.p, I.p, cnt, d, Xg.p);
            EGX_CUDA_TRY(cudaGetLastError());
            EGX_CUDA_TRY(cudaStreamSynchronize(q->stream));
            if (y || var) {
                st = egx_gp_predict_valvar_dev(q->experts[c], Xg.p, cnt, y ? Yc.p : nullptr, var ? Vc.p : nullptr);
                if (st != EGX_OK) return st;
                if (y) scatter_rows_kernel<<<blocks_for(cnt), 256, 0, q->stream>>>(Yc.p, I.p, cnt, 1, Y.p);
                if (var) scatter_rows_kernel<<<blocks_for(cnt), 256, 0, q->stream>>>(Vc.p, I.p, cnt, 1, V.p);
            }
            if (gy) {
                st = egx_gp_predict_gradients_dev(q->experts[c], Xg.p, cnt, GYc.p);
                if (st != EGX_OK) return st;
                scatter_rows_kernel<<<blocks_for(static_cast<long>(cnt) * d), 256, 0, q->stream>>>(GYc.p, I.p, cnt, d, GY.p);
            }
            if (gv) {
                st = egx_gp_predict_var_gradients_dev(q->experts[c], Xg.p, cnt, GVc.p);
                if (st != EGX_OK) return st;
                scatter_rows_kernel<<<blocks_for(static_cast<long>(cnt) * d), 256, 0, q->stream>>>(GVc.p, I.p, cnt, d, GV.p);
            }
            EGX_CUDA_TRY(cudaGetLastError());
            EGX_CUDA_TRY(cudaStreamSynchronize(q->stream));
        }
    }
    if (y) EGX_CUDA_TRY(cudaMemcpyAsync(y, Y.p, sizeof(double) * m, cudaMemcpyDeviceToHost, q->stream));
    if (var) EGX_CUDA_TRY(cudaMemcpyAsync(var, V.p, sizeof(double) * m, cudaMemcpyDeviceToHost, q->stream));
    if (gy) EGX_CUDA_TRY(cudaMemcpyAsync(gy, GY.p, sizeof(double) * m * d, cudaMemcpyDeviceToHost, q->stream));
    if (gv) EGX_CUDA_TRY(cudaMemcpyAsync(gv, GV.p, sizeof(double) * m * d, cudaMemcpyDeviceToHost, q->stream));
    EGX_CUDA_TRY(cudaStreamSynchronize(q->stream));
    return EGX_OK;
}

}  // namespace

extern "C" int egx_moe_create(egx_moe** out, int device, int k, int d, const double* weights, const double* means,
                              const double* covariances, double heaviside_factor) try {
    if (!out) return EGX_INVALID_VALUE;
    *out = nullptr;
    if (k < 1 || d < 1 || !weights || !means || !covariances || !(heaviside_factor > 0.0)) {
        egx_set_error("egx_moe_create: invalid argument");
        return EGX_INVALID_VALUE;
    }
    egx_moe* q = new egx_moe();
    q->device = device;
    q->k = k;
    q->d = d;
    q->heaviside = heaviside_factor;
    q->weights.assign(weights, weights + k);
    q->means.assign(means, means + static_cast<size_t>(k) * d);
    q->covariances.assign(covariances, covariances + static_cast<size_t>(k) * d * d);
    q->prec_chol.assign(static_cast<size_t>(k) * d * d, 0.0);
    q->precisions.assign(static_cast<size_t>(k) * d * d, 0.0);
    q->experts.assign(k, nullptr);
    std::vector<double> l, li(static_cast<size_t>(d) * d);
    for (int c = 0; c < k; ++c) {
        if (!host_cholesky(covariances + static_cast<size_t>(c) * d * d, d, l)) {
            egx_set_error("egx_moe_create: covariance %d is not positive definite", c);
            delete q;
            return EGX_NOT_POSITIVE_DEFINITE;
        }
        // li = L^-1 (forward substitution on the identity), prec_chol = li^T
        std::fill(li.begin(), li.end(), 0.0);
        for (int col = 0; col < d; ++col)
            for (int i = col; i < d; ++i) {
                double s = (i == col) ? 1.0 : 0.0;
                for (int t = col; t < i; ++t) s -= l[i * d + t] * li[t * d + col];
                li[i * d + col] = s / l[i * d + i];
            }
        double* pc = &q->prec_chol[static_cast<size_t>(c) * d * d];
        double* pr = &q->precisions[static_cast<size_t>(c) * d * d];
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) pc[i * d + j] = li[j * d + i];
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                double s = 0.0;
                for (int t = 0; t < d; ++t) s += pc[i * d + t] * pc[j * d + t];
                pr[i * d + j] = s;
            }
    }
    auto fail = [&](const char* what) {
        egx_set_error("egx_moe_create: %s failed: %s", what, cudaGetErrorString(cudaGetLastError()));
        egx_moe_destroy(q);
        return EGX_CUDA_ERROR;
    };
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice");
    if (cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate");
    if (egx_dev_malloc(&q->logw, k * sizeof(double)) != cudaSuccess || egx_dev_malloc(&q->w, k * sizeof(double)) != cudaSuccess ||
        egx_dev_malloc(&q->mu, sizeof(double) * k * d) != cudaSuccess ||
        egx_dev_malloc(&q->pcs, sizeof(double) * k * d * d) != cudaSuccess ||
        egx_dev_malloc(&q->logdet, k * sizeof(double)) != cudaSuccess ||
        egx_dev_malloc(&q->precs_h, sizeof(double) * k * d * d) != cudaSuccess)
        return fail("device allocation");
    const int st = upload(q);
    if (st != EGX_OK) {
        egx_moe_destroy(q);
        return st;
    }
    *out = q;
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" void egx_moe_destroy(egx_moe* q) {
    if (!q) return;
    cudaSetDevice(q->device);
    egx_dev_free(q->logw);
    egx_dev_free(q->w);
    egx_dev_free(q->mu);
    egx_dev_free(q->pcs);
    egx_dev_free(q->logdet);
    egx_dev_free(q->precs_h);
    if (q->stream) cudaStreamDestroy(q->stream);
    delete q;
}

extern "C" int egx_moe_set_heaviside_factor(egx_moe* q, double factor) try {
    if (!q || !(factor > 0.0)) return EGX_INVALID_VALUE;
    std::lock_guard<std::mutex> lk(q->mtx);
    EGX_CUDA_TRY(cudaSetDevice(q->device));
    q->heaviside = factor;
    return upload(q);
}
EGX_ABI_CATCH

extern "C" int egx_moe_set_expert(egx_moe* q, int cluster, egx_gp_ctx* ctx) try {
    if (!q || cluster < 0 || cluster >= q->k) return EGX_INVALID_VALUE;
    if (ctx) {
        int d = 0;
        egx_gp_dims(ctx, nullptr, &d, nullptr, nullptr);
        if (d != q->d) {
            egx_set_error("egx_moe_set_expert: expert input dimension %d != mixture dimension %d", d, q->d);
            return EGX_INVALID_VALUE;
        }
    }
    std::lock_guard<std::mutex> lk(q->mtx);
    q->experts[cluster] = ctx;
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_moe_parameters(const egx_moe* q, double* precisions, double* precisions_chol, double* log_det) try {
    if (!q) return EGX_INVALID_VALUE;
    const size_t kdd = static_cast<size_t>(q->k) * q->d * q->d;
    if (precisions) std::memcpy(precisions, q->precisions.data(), kdd * sizeof(double));
    if (precisions_chol) std::memcpy(precisions_chol, q->prec_chol.data(), kdd * sizeof(double));
    if (log_det) {
        const double f = std::pow(q->heaviside, -0.5);
        for (int c = 0; c < q->k; ++c) {
            double s = 0.0;
            for (int i = 0; i < q->d; ++i) s += std::log(q->prec_chol[static_cast<size_t>(c) * q->d * q->d + i * q->d + i] * f);
            log_det[c] = s;
        }
    }
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_moe_predict_probas(egx_moe* q, const double* x, int m, double* probas, int* clusters) try {
    if (!q || m < 0 || (m > 0 && !x) || (!probas && !clusters)) return EGX_INVALID_VALUE;
    if (m == 0) return EGX_OK;
    std::lock_guard<std::mutex> lk(q->mtx);
    EGX_CUDA_TRY(cudaSetDevice(q->device));
    const int k = q->k, d = q->d;
    const int mb = std::min(m, kChunk);
    DevBuf X, P;
    DevIdx L;
    EGX_CUDA_TRY(X.alloc(static_cast<size_t>(mb) * d));
    EGX_CUDA_TRY(P.alloc(static_cast<size_t>(mb) * k));
    EGX_CUDA_TRY(L.alloc(mb));
    for (int i0 = 0; i0 < m; i0 += mb) {
        const int mc = std::min(mb, m - i0);
        EGX_CUDA_TRY(cudaMemcpyAsync(X.p, x + static_cast<long>(i0) * d, sizeof(double) * mc * d, cudaMemcpyHostToDevice,
                                     q->stream));
        const int st = probas_dev(q, X.p, mc, P.p, L.p);
        if (st != EGX_OK) return st;
        if (probas)
            EGX_CUDA_TRY(cudaMemcpyAsync(probas + static_cast<long>(i0) * k, P.p, sizeof(double) * mc * k,
                                         cudaMemcpyDeviceToHost, q->stream));
        if (clusters)
            EGX_CUDA_TRY(cudaMemcpyAsync(clusters + i0, L.p, sizeof(int) * mc, cudaMemcpyDeviceToHost, q->stream));
        EGX_CUDA_TRY(cudaStreamSynchronize(q->stream));
    }
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_moe_predict_probas_derivatives(egx_moe* q, const double* x, int m, double* dprobas) try {
    if (!q || m < 0 || (m > 0 && (!x || !dprobas))) return EGX_INVALID_VALUE;
    if (m == 0) return EGX_OK;
    std::lock_guard<std::mutex> lk(q->mtx);
    EGX_CUDA_TRY(cudaSetDevice(q->device));
    const int k = q->k, d = q->d;
    const int mb = std::min(m, kChunk);
    DevBuf X, W, DP;
    EGX_CUDA_TRY(X.alloc(static_cast<size_t>(mb) * d));
    EGX_CUDA_TRY(W.alloc(static_cast<size_t>(mb) * k));
    EGX_CUDA_TRY(DP.alloc(static_cast<size_t>(mb) * k * d));
    for (int i0 = 0; i0 < m; i0 += mb) {
        const int mc = std::min(mb, m - i0);
        EGX_CUDA_TRY(cudaMemcpyAsync(X.p, x + static_cast<long>(i0) * d, sizeof(double) * mc * d, cudaMemcpyHostToDevice,
                                     q->stream));
        gmx_probas_deriv_kernel<<<blocks_for(mc, 128), 128, 0, q->stream>>>(X.p, mc, d, k, q->w, q->mu, q->pcs, q->logdet,
                                                                            q->precs_h, W.p, DP.p);
        EGX_CUDA_TRY(cudaGetLastError());
        EGX_CUDA_TRY(cudaMemcpyAsync(dprobas + static_cast<long>(i0) * k * d, DP.p, sizeof(double) * mc * k * d,
                                     cudaMemcpyDeviceToHost, q->stream));
        EGX_CUDA_TRY(cudaStreamSynchronize(q->stream));
    }
    return EGX_OK;
}
EGX_ABI_CATCH

extern "C" int egx_moe_predict(egx_moe* q, int recombination, const double* x, int m, double* y, double* var,
                               double* grad_y, double* grad_var) try {
    if (!q || m < 0 || (m > 0 && !x) || (recombination != EGX_RECOMB_HARD && recombination != EGX_RECOMB_SMOOTH))
        return EGX_INVALID_VALUE;
    if (!y && !var && !grad_y && !grad_var) return EGX_INVALID_VALUE;
    if (m == 0) return EGX_OK;
    std::lock_guard<std::mutex> lk(q->mtx);
    int st = check_experts(q);
    if (st != EGX_OK) return st;
    EGX_CUDA_TRY(cudaSetDevice(q->device));
    const int d = q->d;
    for (int i0 = 0; i0 < m; i0 += kChunk) {
        const int mc = std::min(kChunk, m - i0);
        st = predict_chunk(q, recombination, x + static_cast<long>(i0) * d, mc, y ? y + i0 : nullptr,
                           var ? var + i0 : nullptr, grad_y ? grad_y + static_cast<long>(i0) * d : nullptr,
                           grad_var ? grad_var + static_cast<long>(i0) * d : nullptr);
        if (st != EGX_OK) return st;
    }
    return EGX_OK;
}
EGX_ABI_CATCH
