// K8: whole reduced-likelihood evaluation for small training sets, ONE CTA PER THETA.
//
// For n up to ~160 the matrix R(theta), the appended right-hand sides and all work vectors fit
// in the 227 KB of shared memory of one SM, so a batch of B candidate thetas (multistart chains
// advanced in lock step, or an EGO theta sweep: BASELINE config 5, 512 candidates) is evaluated by
// B independent CTAs with no global-memory traffic besides reading X (n x d) -- the batched form of
// the `objfn` closure of gp/src/algorithm.rs:880-897 -> reduced_likelihood :989-1056.
//
// Inside the CTA: fused distance+correlation build of the lower triangle of R (never through HBM),
// right-looking Cholesky with [F|y]^T carried as extra rows (forward solves fused, as in the blocked
// path), Householder thin QR of Ft, beta, rho, sigma2, log10-det, rlf.  The p x p factor G goes back
// to the host for the condition-number test of :1010-1027.
#include "common.cuh"
#include "../../include/egobox_gpu.h"

namespace {

constexpr int SB_THREADS = 256;
constexpr int SB_WARPS = SB_THREADS / 32;

__device__ __forceinline__ double sb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ double sb_block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = sb_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = (lane < SB_WARPS) ? red[lane] : 0.0;
    return sb_warp_sum(t);
}

template <int CORR>
__device__ __forceinline__ double sb_pair(const double* __restrict__ xi, const double* __restrict__ xj, int d, int h,
                                          const double* __restrict__ tw) {
    // tw: SqExp/AbsExp: d summed weights; Matern: d*h theta_l |W_jl|
    if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) {
        double acc = 0.0;
        for (int j = 0; j < d; ++j) {
            const double dx = xi[j] - xj[j];
            acc += tw[j] * (dx * dx);
        }
        return exp(-0.5 * acc);
    }
    if (CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) {
        double acc = 0.0;
        for (int j = 0; j < d; ++j) acc += tw[j] * fabs(xi[j] - xj[j]);
        return exp(-acc);
    }
    double prod = 1.0, acc = 0.0;
    for (int j = 0; j < d; ++j) {
        const double dx = xi[j] - xj[j], ad = fabs(dx);
        for (int l = 0; l < h; ++l) {
            const double v = tw[j * h + l];
            if (CORR == EGX_CORR_MATERN32) prod *= 1.0 + (1.7320508075688772 * v) * ad;
            else prod *= (1.0 + (2.23606797749979 * v) * ad) + (5.0 / 3.0) * (((v * v) * dx) * dx);
            acc += v * ad;
        }
    }
    return prod * exp(-(CORR == EGX_CORR_MATERN32 ? 1.7320508075688772 : 2.23606797749979) * acc);
}

struct SmallOut {
    double rlf, sigma2;
    int info, pad_;
};

template <int CORR>
__global__ void __launch_bounds__(SB_THREADS)
    small_batch_kernel(const double* __restrict__ X, int n, int d, const double* __restrict__ W, int h,
                       const double* __restrict__ thetas, const double* __restrict__ FyT, long ldf, int p,
                       double diag_value, SmallOut* __restrict__ out, double* __restrict__ out_G) {
    extern __shared__ double sh[];
    const int q = p + 1, rows = n + q;
    const int ldS = (n & 1) ? n + 2 : n + 1;          // odd leading dimension: conflict-free column walks
    double* S = sh;                                    // [rows][ldS]
    double* Wq = S + static_cast<long>(rows) * ldS;    // [q][n] Householder work copy
    double* Gs = Wq + static_cast<long>(q) * n;        // [p][p]
    double* alphas = Gs + p * p;
    double* ytil = alphas + p;
    double* betas = ytil + p;
    double* red = betas + p;                           // [32]
    double* tw = red + 32;                             // [d*h]
    __shared__ double s_vtv;
    __shared__ int s_info;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* theta = thetas + static_cast<long>(blockIdx.x) * h;
    if (tid == 0) s_info = 0;

    // per-theta kernel weights (correlation_models.rs:97-100, 191, 333, 505)
    if (CORR == EGX_CORR_SQUARED_EXPONENTIAL || CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) {
        for (int j = tid; j < d; j += SB_THREADS) {
            double s = 0.0;
            for (int l = 0; l < h; ++l) {
                if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) {
                    const double v = theta[l] * W[j * h + l];
                    s += v * v;
                } else {
                    s += fabs(W[j * h + l]) * theta[l];
                }
            }
            tw[j] = s;
        }
    } else {
        for (int t = tid; t < d * h; t += SB_THREADS) tw[t] = theta[t % h] * fabs(W[t]);
    }
    __syncthreads();

    // R(theta): lower triangle + diagonal, straight into shared memory
    const int npairs = n * (n - 1) / 2;
    for (int idx = tid; idx < npairs; idx += SB_THREADS) {
        int i = static_cast<int>((1.0 + sqrt(1.0 + 8.0 * static_cast<double>(idx))) * 0.5);
        while (i * (i - 1) / 2 > idx) --i;
        while ((i + 1) * i / 2 <= idx) ++i;
        const int j = idx - i * (i - 1) / 2;
        S[i * ldS + j] = sb_pair<CORR>(X + static_cast<long>(i) * d, X + static_cast<long>(j) * d, d, h, tw);
    }
    for (int i = tid; i < n; i += SB_THREADS) S[i * ldS + i] = diag_value;
    for (int idx = tid; idx < q * n; idx += SB_THREADS) {
        const int r = idx / n, c = idx - r * n;
        S[(n + r) * ldS + c] = FyT[static_cast<long>(r) * ldf + c];
    }

    // Cholesky with the RHS rows carried along
    for (int j = 0; j < n; ++j) {
        __syncthreads();
        const double dj = S[j * ldS + j];
        if (!(dj > 0.0) && tid == 0 && s_info == 0) s_info = j + 1;
        const double ljj = sqrt(dj);
        __syncthreads();
        if (tid == 0) S[j * ldS + j] = ljj;
        for (int i = j + 1 + tid; i < rows; i += SB_THREADS) S[i * ldS + j] /= ljj;
        __syncthreads();
        for (int i = j + 1 + warp; i < rows; i += SB_WARPS) {
            const double lij = S[i * ldS + j];
            const int cmax = (i < n) ? i : n - 1;
            for (int c = j + 1 + lane; c <= cmax; c += 32) S[i * ldS + c] -= lij * S[c * ldS + j];
        }
    }
    __syncthreads();

    // thin QR of Ft (rows n..n+p-1 of S, each of length n) applied to yt (row n+p)
    for (int idx = tid; idx < q * n; idx += SB_THREADS) {
        const int r = idx / n, c = idx - r * n;
        Wq[idx] = S[(n + r) * ldS + c];
    }
    __syncthreads();
    for (int j = 0; j < p; ++j) {
        double* vj = Wq + j * n;
        double part = 0.0;
        for (int i = j + tid; i < n; i += SB_THREADS) part += vj[i] * vj[i];
        const double nrm2 = sb_block_sum(part, red);
        if (tid == 0) {
            const double x0 = vj[j], nrm = sqrt(nrm2);
            double alpha = 0.0, vtv = 0.0;
            if (nrm > 0.0) {
                alpha = (x0 > 0.0) ? -nrm : nrm;
                const double v0 = x0 - alpha;
                vtv = (nrm2 - x0 * x0) + v0 * v0;
                vj[j] = v0;
            }
            alphas[j] = alpha;
            s_vtv = vtv;
        }
        __syncthreads();
        const double vtv = s_vtv;
        if (vtv > 0.0) {
            for (int k = j + 1 + warp; k <= p; k += SB_WARPS) {
                double* wk = Wq + k * n;
                double dot = 0.0;
                for (int i = j + lane; i < n; i += 32) dot += vj[i] * wk[i];
                dot = sb_warp_sum(dot);
                const double f = 2.0 * dot / vtv;
                for (int i = j + lane; i < n; i += 32) wk[i] -= f * vj[i];
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < p * p; idx += SB_THREADS) {
        const int i = idx / p, j = idx - i * p;
        double v = 0.0;
        if (i < j) v = Wq[j * n + i];
        else if (i == j) v = alphas[i];
        if (alphas[i] < 0.0) v = -v;
        Gs[idx] = v;
        out_G[static_cast<long>(blockIdx.x) * p * p + idx] = v;
    }
    if (tid < p) {
        const double v = Wq[p * n + tid];
        ytil[tid] = (alphas[tid] < 0.0) ? -v : v;
    }
    __syncthreads();
    if (warp == 0) {
        for (int i = p - 1; i >= 0; --i) {
            double s = 0.0;
            for (int j = i + 1 + lane; j < p; j += 32) s += Gs[i * p + j] * betas[j];
            s = sb_warp_sum(s);
            if (lane == 0) betas[i] = (ytil[i] - s) / Gs[i * p + i];
            __syncwarp();
        }
    }
    __syncthreads();

    double part = 0.0;
    for (int i = tid; i < n; i += SB_THREADS) {
        double r = S[(n + p) * ldS + i];
        for (int l = 0; l < p; ++l) r -= S[(n + l) * ldS + i] * betas[l];
        part += r * r;
    }
    const double rho_sqr = sb_block_sum(part, red);
    part = 0.0;
    for (int i = tid; i < n; i += SB_THREADS) part += log10(S[i * ldS + i]);
    const double slog = sb_block_sum(part, red);
    if (tid == 0) {
        const double nd = static_cast<double>(n);
        const double sigma2 = rho_sqr / nd;
        SmallOut o;
        o.sigma2 = sigma2;
        o.rlf = -nd * (log10(sigma2) + slog * 2.0 / nd);
        o.info = s_info;
        o.pad_ = 0;
        out[blockIdx.x] = o;
    }
}

}  // namespace

size_t small_batch_smem_bytes(int n, int d, int h, int p) {
    const int q = p + 1, rows = n + q;
    const int ldS = (n & 1) ? n + 2 : n + 1;
    const size_t doubles = static_cast<size_t>(rows) * ldS + static_cast<size_t>(q) * n + static_cast<size_t>(p) * p +
                           3 * static_cast<size_t>(p) + 32 + static_cast<size_t>(d) * h;
    return doubles * sizeof(double);
}

void launch_small_batch(int corr, const double* X, int n, int d, const double* W, int h, const double* thetas, int B,
                        const double* FyT, long ldf, int p, double diag_value, void* out, double* out_G,
                        cudaStream_t s) {
    const size_t smem = small_batch_smem_bytes(n, d, h, p);
#define EGX_LAUNCH_SB(CK)                                                                                         \
    cudaFuncSetAttribute(small_batch_kernel<CK>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
    small_batch_kernel<CK><<<B, SB_THREADS, smem, s>>>(X, n, d, W, h, thetas, FyT, ldf, p, diag_value,              \
                                                       static_cast<SmallOut*>(out), out_G);
    switch (corr) {
        case EGX_CORR_SQUARED_EXPONENTIAL: EGX_LAUNCH_SB(EGX_CORR_SQUARED_EXPONENTIAL) break;
        case EGX_CORR_ABSOLUTE_EXPONENTIAL: EGX_LAUNCH_SB(EGX_CORR_ABSOLUTE_EXPONENTIAL) break;
        case EGX_CORR_MATERN32: EGX_LAUNCH_SB(EGX_CORR_MATERN32) break;
        default: EGX_LAUNCH_SB(EGX_CORR_MATERN52) break;
    }
#undef EGX_LAUNCH_SB
}
