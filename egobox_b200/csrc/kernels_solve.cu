// Generalised-least-squares epilogue of the reduced likelihood, the single-RHS back solve
// for gamma and the variance epilogue of predict_var.
//
// Reference being replaced: gp/src/algorithm.rs:1007-1043 (thin QR of Ft, beta, rho,
// sigma2, gamma, log10-det, reduced likelihood) and :272-278, 352-367 (u, mse clamp).
// The condition-number test on the p x p factor G (:1010-1027) runs on the host from the
// G this kernel emits (O(p^3), microseconds) -- see gp_context.cu.
#include "common.cuh"
#include "../../include/egobox_gpu.h"

namespace {

constexpr int GLS_THREADS = 1024;
constexpr int GLS_MAXP = 256;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the whole block; result valid in every thread
__device__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
    t = warp_sum(t);
    return t;
}

// Rows npad .. npad+p of M hold (L^-1 [F | y])^T after the factorisation.
__global__ void __launch_bounds__(GLS_THREADS)
    gls_kernel(const double* __restrict__ M, long ld, int n, int npad, int p, double* __restrict__ work,
               double* __restrict__ G, double* __restrict__ beta, double* __restrict__ rho,
               EvalResult* __restrict__ res, const int* __restrict__ info) {
    __shared__ double red[32];
    __shared__ double alphas[GLS_MAXP];
    __shared__ double ytil[GLS_MAXP];
    __shared__ double betas[GLS_MAXP];
    __shared__ double s_vtv;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = p + 1;
    const double* rows = M + static_cast<long>(npad) * ld;

    for (long idx = tid; idx < static_cast<long>(q) * npad; idx += GLS_THREADS) {
        const int r = static_cast<int>(idx / npad), c = static_cast<int>(idx - static_cast<long>(r) * npad);
        work[idx] = rows[static_cast<long>(r) * ld + c];
    }
    __syncthreads();

    // Householder QR of Ft (columns = rows 0..p-1 of work), applied to yt (row p) as well
    for (int j = 0; j < p; ++j) {
        double* vj = work + static_cast<long>(j) * npad;
        double part = 0.0;
        for (int i = j + tid; i < npad; i += GLS_THREADS) part += vj[i] * vj[i];
        const double nrm2 = block_sum(part, red);
        if (tid == 0) {
            const double x0 = vj[j];
            const double nrm = sqrt(nrm2);
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
            for (int k = j + 1 + warp; k <= p; k += GLS_THREADS / 32) {
                double* wk = work + static_cast<long>(k) * npad;
                double dot = 0.0;
                for (int i = j + lane; i < npad; i += 32) dot += vj[i] * wk[i];
                dot = warp_sum(dot);
                const double f = 2.0 * dot / vtv;
                for (int i = j + lane; i < npad; i += 32) wk[i] -= f * vj[i];
            }
        }
        __syncthreads();
    }

    // G (upper, diag > 0 like linfa-linalg / nalgebra `r()`), and Q^T yt
    for (int idx = tid; idx < p * p; idx += GLS_THREADS) {
        const int i = idx / p, j = idx - i * p;
        double v = 0.0;
        if (i < j) v = work[static_cast<long>(j) * npad + i];
        else if (i == j) v = alphas[i];
        if (alphas[i] < 0.0) v = -v;
        G[idx] = v;
    }
    if (tid < p) {
        const double v = work[static_cast<long>(p) * npad + tid];
        ytil[tid] = (alphas[tid] < 0.0) ? -v : v;
    }
    __syncthreads();

    // beta: G beta = Q^T yt   (algorithm.rs:1030)
    if (warp == 0) {
        for (int i = p - 1; i >= 0; --i) {
            double s = 0.0;
            for (int j = i + 1 + lane; j < p; j += 32) s += G[i * p + j] * betas[j];
            s = warp_sum(s);
            if (lane == 0) betas[i] = (ytil[i] - s) / G[i * p + i];
            __syncwarp();
        }
    }
    __syncthreads();
    if (tid < p) beta[tid] = betas[tid];

    // rho = yt - Ft beta (original, untransformed rows), sigma2 = rho^T rho / n
    double part = 0.0;
    for (int i = tid; i < npad; i += GLS_THREADS) {
        double r = rows[static_cast<long>(p) * ld + i];
        for (int l = 0; l < p; ++l) r -= rows[static_cast<long>(l) * ld + i] * betas[l];
        rho[i] = r;
        part += r * r;
    }
    const double rho_sqr = block_sum(part, red);

    // log10 det:  2/n * sum log10 L_ii   (algorithm.rs:1039)
    part = 0.0;
    for (int i = tid; i < n; i += GLS_THREADS) part += log10(M[static_cast<long>(i) * ld + i]);
    const double slog = block_sum(part, red);

    if (tid == 0) {
        const double nd = static_cast<double>(n);
        const double logdet = slog * 2.0 / nd;
        const double sigma2 = rho_sqr / nd;
        res->rho_sqr = rho_sqr;
        res->sigma2 = sigma2;
        res->logdet = logdet;
        res->rlf = -nd * (log10(sigma2) + logdet);
        res->info = *info;
    }
}

// gamma_k = L_kk^-T rho_k   (one 128-block of the back substitution, algorithm.rs:1034)
__global__ void __launch_bounds__(128) backsolve_diag_kernel(const double* __restrict__ L, long ld,
                                                             double* __restrict__ b_io) {
    extern __shared__ double Ls[];   // [128][128]
    __shared__ double b[EGX_NB];
    const int tid = threadIdx.x;
    for (int r = 0; r < EGX_NB; ++r) Ls[r * EGX_NB + tid] = (tid <= r) ? L[static_cast<long>(r) * ld + tid] : 0.0;
    b[tid] = b_io[tid];
    __syncthreads();
    for (int c = EGX_NB - 1; c >= 0; --c) {
        if (tid == c) b[c] = b[c] / Ls[c * EGX_NB + c];
        __syncthreads();
        if (tid < c) b[tid] -= Ls[c * EGX_NB + tid] * b[c];
        __syncthreads();
    }
    b_io[tid] = b[tid];
}

// rho_c -= L[k, c]^T gamma_k for every column block c < k
__global__ void __launch_bounds__(128) backsolve_update_kernel(const double* __restrict__ Lrow, long ld,
                                                               const double* __restrict__ gamma_k,
                                                               double* __restrict__ rho) {
    __shared__ double g[EGX_NB];
    const int tid = threadIdx.x;
    g[tid] = gamma_k[tid];
    __syncthreads();
    const double* Lc = Lrow + static_cast<long>(blockIdx.x) * EGX_NB + tid;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
    for (int i = 0; i < EGX_NB; i += 4) {
        s0 += Lc[static_cast<long>(i) * ld] * g[i];
        s1 += Lc[static_cast<long>(i + 1) * ld] * g[i + 1];
        s2 += Lc[static_cast<long>(i + 2) * ld] * g[i + 2];
        s3 += Lc[static_cast<long>(i + 3) * ld] * g[i + 3];
    }
    rho[blockIdx.x * EGX_NB + tid] -= (s0 + s1) + (s2 + s3);
}

// ---------------------------------------------------------------------------
// v <- L^-T v in ONE launch (r02; the r01 form was 2 T launches of one-CTA kernels, 5.2 ms at T = 64).
// CTA q owns block c = T-1-q of the vector.  It walks k = T-1 .. c+1: the rows of L[k, c] it needs are loaded into
// registers BEFORE it waits for gamma_k (they do not depend on it), so the wait hides the loads; then
// b_c -= L[k, c]^T gamma_k.  When gamma_{c+1} has been folded in it solves its own diagonal block with the inverted
// 32 x 32 sub-blocks of K3 (a blocked back substitution of four 32-wide steps out of shared memory), publishes
// gamma_c and raises flags[c].  A CTA only ever waits for CTAs with a smaller blockIdx, which the hardware has
// dispatched before it, so the chain cannot dead-lock whatever the number of resident CTAs.
// ---------------------------------------------------------------------------
constexpr int BSC_THREADS = 512;

__device__ __forceinline__ int bsc_ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void bsc_st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(BSC_THREADS) backsolve_chain_kernel(const double* __restrict__ L, long ld,
                                                                      const double* __restrict__ Dinv, int T,
                                                                      double* v, int* flags) {
    extern __shared__ __align__(16) double bsm[];
    double* Lsub = bsm;                   // the six 32 x 32 blocks below the diagonal of L_cc: block (s', s'') at s'(s'-1)/2 + s''
    double* Dis = Lsub + 6 * 1024;        // inverted diagonal 32 x 32 blocks
    double* g = Dis + 4 * 1024;           // gamma_k of the step / x of the diagonal solve
    double* b = g + EGX_NB;
    double (*part)[EGX_NB] = reinterpret_cast<double (*)[EGX_NB]>(b + EGX_NB);
    const int tid = threadIdx.x;
    const int col = tid & 127, rg = tid >> 7;      // column of the block, group of 32 rows
    const int c = T - 1 - static_cast<int>(blockIdx.x);
    const double* Lcc = L + static_cast<long>(c) * EGX_NB * ld + static_cast<long>(c) * EGX_NB;
    for (int e = tid; e < 6 * 1024; e += BSC_THREADS) {
        const int blk = e >> 10, k = (e >> 5) & 31, i = e & 31;
        const int sp = blk == 0 ? 1 : (blk < 3 ? 2 : 3);
        const int spp = blk - sp * (sp - 1) / 2;
        Lsub[e] = Lcc[static_cast<long>(32 * sp + k) * ld + 32 * spp + i];
    }
    for (int e = tid; e < 4096; e += BSC_THREADS) Dis[e] = Dinv[static_cast<long>(c) * 4096 + e];
    double acc = 0.0;     // sum over k of (L[k, c]^T gamma_k)[col], rows of group rg
    for (int k = T - 1; k > c; --k) {
        const double* Lk = L + (static_cast<long>(k) * EGX_NB + rg * 32) * ld + static_cast<long>(c) * EGX_NB + col;
        double lv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) lv[i] = __ldcs(Lk + static_cast<long>(i) * ld);
        if (tid == 0)
            while (bsc_ld_acquire(flags + k) == 0) {
            }
        __syncthreads();
        if (tid < EGX_NB) g[tid] = __ldcg(v + static_cast<long>(k) * EGX_NB + tid);
        __syncthreads();
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            s0 = fma(lv[i], g[rg * 32 + i], s0);
            s1 = fma(lv[i + 1], g[rg * 32 + i + 1], s1);
        }
        acc += s0 + s1;
    }
    part[rg][col] = acc;
    __syncthreads();
    if (tid < EGX_NB) b[tid] = v[static_cast<long>(c) * EGX_NB + tid] - ((part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]));
    __syncthreads();
    // x = L_cc^-T b:  for s = 3 .. 0:  x_s = Dinv_s^T b_s ;  b_s'' -= L[s, s'']^T x_s (s'' < s)
    for (int s = 3; s >= 0; --s) {
        if (tid < 32) {
            double x0 = 0.0, x1 = 0.0;
#pragma unroll
            for (int k = 0; k < 32; k += 2) {       // Dinv is lower triangular: rows k < tid hold zeros above the diagonal
                x0 = fma(Dis[s * 1024 + k * 32 + tid], b[32 * s + k], x0);
                x1 = fma(Dis[s * 1024 + (k + 1) * 32 + tid], b[32 * s + k + 1], x1);
            }
            g[32 * s + tid] = x0 + x1;
        }
        __syncthreads();
        if (tid < 32 * s) {
            const int spp = tid >> 5, i = tid & 31;
            const double* Lb = Lsub + (s * (s - 1) / 2 + spp) * 1024 + i;
            double u0 = 0.0, u1 = 0.0;
#pragma unroll
            for (int k = 0; k < 32; k += 2) {
                u0 = fma(Lb[k * 32], g[32 * s + k], u0);
                u1 = fma(Lb[(k + 1) * 32], g[32 * s + k + 1], u1);
            }
            b[tid] -= u0 + u1;
        }
        __syncthreads();
    }
    if (tid < EGX_NB) v[static_cast<long>(c) * EGX_NB + tid] = g[tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) bsc_st_release(flags + c, 1);
}

// variance epilogue: one warp per prediction point, 8 points per CTA.
//   s1 = sum_j rt_j^2 ; z = Ft^T rt - f(x) ; u = G^-T z ; var = sigma2 * max(0, 1 - s1 + |u|^2)
__global__ void __launch_bounds__(256)
    var_finish_kernel(const double* __restrict__ Y, long ldy, int m, int npad, const double* __restrict__ xraw,
                      const double* __restrict__ x_mean, const double* __restrict__ x_std, int d,
                      const double* __restrict__ FtT, long ldf, const double* __restrict__ G, int p,
                      const int* __restrict__ basis_i, const int* __restrict__ basis_j, double sigma2,
                      double* __restrict__ var, double* __restrict__ U) {
    extern __shared__ double vsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* xs = vsm + warp * (d + p);
    double* z = xs + d;
    const int i = blockIdx.x * 8 + warp;
    if (i >= m) return;
    const double* y = Y + static_cast<long>(i) * ldy;
    for (int c = lane; c < d; c += 32) xs[c] = (xraw[static_cast<long>(i) * d + c] - x_mean[c]) / x_std[c];
    __syncwarp();

    double s1 = 0.0;
    for (int j = 2 * lane; j < npad; j += 64) {
        const double2 v = *reinterpret_cast<const double2*>(y + j);
        s1 += v.x * v.x + v.y * v.y;
    }
    s1 = warp_sum(s1);

    for (int l0 = 0; l0 < p; l0 += 4) {
        double a[4] = {0.0, 0.0, 0.0, 0.0};
        const int nl = min(4, p - l0);
        for (int j = 2 * lane; j < npad; j += 64) {
            const double2 v = *reinterpret_cast<const double2*>(y + j);
#pragma unroll
            for (int t = 0; t < 4; ++t)
                if (t < nl) {
                    const double2 f = *reinterpret_cast<const double2*>(FtT + static_cast<long>(l0 + t) * ldf + j);
                    a[t] += v.x * f.x + v.y * f.y;
                }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const double s = warp_sum(a[t]);
            if (t < nl && lane == 0) {
                const int bi = basis_i[l0 + t], bj = basis_j[l0 + t];
                const double f = (bi < 0 ? 1.0 : xs[bi]) * (bj < 0 ? 1.0 : xs[bj]);
                z[l0 + t] = s - f;
            }
        }
    }
    __syncwarp();
    // u = G^-T z  (forward substitution with the lower-triangular G^T), in place
    double su = 0.0;
    for (int a = 0; a < p; ++a) {
        double s = 0.0;
        for (int b = lane; b < a; b += 32) s += G[b * p + a] * z[b];
        s = warp_sum(s);
        const double u = (z[a] - s) / G[a * p + a];
        __syncwarp();
        if (lane == 0) {
            z[a] = u;
            if (U != nullptr) U[static_cast<long>(i) * p + a] = u;
        }
        __syncwarp();
        su += u * u;
    }
    if (lane == 0 && var != nullptr) {
        double mse = (1.0 - s1) + su;
        mse = sigma2 * mse;
        var[i] = (mse < 0.0) ? 0.0 : mse;
    }
}

// Conditional covariance epilogue (gp/src/algorithm.rs:323-324): C holds K(x, x) - rt^T rt on entry;
//   cov[i][j] = sigma2 * (C[i][j] + u_i . u_j)   for i, j < m ;  identity on the padding (so that the padded
// matrix can be factorised as it is).
__global__ void __launch_bounds__(256)
    cov_finish_kernel(double* __restrict__ C, long ld, int m, int mpad, const double* __restrict__ U, int p,
                      double sigma2) {
    const int j = blockIdx.x * 64 + (threadIdx.x & 63);
    const int i = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i >= mpad || j >= mpad) return;
    double v;
    if (i < m && j < m) {
        double s = C[static_cast<long>(i) * ld + j];
        for (int l = 0; l < p; ++l) s += U[static_cast<long>(i) * p + l] * U[static_cast<long>(j) * p + l];
        v = sigma2 * s;
    } else {
        v = (i == j) ? 1.0 : 0.0;
    }
    C[static_cast<long>(i) * ld + j] = v;
}

// out[i][:] = (x[i][:] - mean) / std for i < m, zero rows up to mpad
__global__ void normalize_rows_kernel(const double* __restrict__ x, int m, int mpad, int d,
                                      const double* __restrict__ mean, const double* __restrict__ sd,
                                      double* __restrict__ out) {
    const long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<long>(mpad) * d) return;
    const int i = static_cast<int>(e / d), c = static_cast<int>(e - static_cast<long>(i) * d);
    out[e] = (i < m) ? (x[e] - mean[c]) / sd[c] : 0.0;
}

// strictly-upper part of an (npad x npad) matrix <- 0 (a factor used as a dense GEMM operand)
__global__ void zero_upper_kernel(double* __restrict__ A, long ld, int npad) {
    const int j = blockIdx.x * 64 + (threadIdx.x & 63);
    const int i = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (i < npad && j < npad && j > i) A[static_cast<long>(i) * ld + j] = 0.0;
}

// out[i][c] = mean[i] for i < m (else 0): the trajectories start from the predicted mean
__global__ void bcast_rows_kernel(double* __restrict__ out, long ld, int m, int mpad, int cols,
                                  const double* __restrict__ mean) {
    const long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= static_cast<long>(mpad) * cols) return;
    const int i = static_cast<int>(e / cols), c = static_cast<int>(e - static_cast<long>(i) * cols);
    out[static_cast<long>(i) * ld + c] = (i < m) ? mean[i] : 0.0;
}

}  // namespace

void launch_cov_finish(double* C, long ld, int m, int mpad, const double* U, int p, double sigma2, cudaStream_t s) {
    cov_finish_kernel<<<dim3((mpad + 63) / 64, (mpad + 3) / 4), 256, 0, s>>>(C, ld, m, mpad, U, p, sigma2);
}
void launch_normalize_rows(const double* x, int m, int mpad, int d, const double* mean, const double* sd, double* out,
                           cudaStream_t s) {
    const long tot = static_cast<long>(mpad) * d;
    normalize_rows_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(x, m, mpad, d, mean, sd, out);
}
void launch_zero_upper(double* A, long ld, int npad, cudaStream_t s) {
    zero_upper_kernel<<<dim3((npad + 63) / 64, (npad + 3) / 4), 256, 0, s>>>(A, ld, npad);
}
void launch_bcast_rows(double* out, long ld, int m, int mpad, int cols, const double* mean, cudaStream_t s) {
    const long tot = static_cast<long>(mpad) * cols;
    bcast_rows_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(out, ld, m, mpad, cols, mean);
}

void launch_gls(const double* M, long ld, int n, int npad, int p, double* work, double* G, double* beta, double* rho,
                EvalResult* res, const int* info, cudaStream_t s) {
    gls_kernel<<<1, GLS_THREADS, 0, s>>>(M, ld, n, npad, p, work, G, beta, rho, res, info);
}

void launch_backsolve_chain(const double* L, long ld, const double* Dinv, int T, double* v, int* flags, cudaStream_t s) {
    if (T <= 0) return;
    static bool configured_dev[64] = {false};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& configured = configured_dev[dev_ & 63];
    const int smem = (6 * 1024 + 4 * 1024 + 6 * EGX_NB) * static_cast<int>(sizeof(double));
    if (!configured) {
        cudaFuncSetAttribute(backsolve_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    cudaMemsetAsync(flags, 0, static_cast<size_t>(T) * sizeof(int), s);
    backsolve_chain_kernel<<<T, BSC_THREADS, smem, s>>>(L, ld, Dinv, T, v, flags);
}

void launch_backsolve_diag(const double* Lkk, long ld, double* rho_k, cudaStream_t s) {
    static bool configured_dev[64] = {false};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& configured = configured_dev[dev_ & 63];   // the attribute is per device (one process may drive several)
    const int smem = EGX_NB * EGX_NB * sizeof(double);
    if (!configured) {
        cudaFuncSetAttribute(backsolve_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    backsolve_diag_kernel<<<1, 128, smem, s>>>(Lkk, ld, rho_k);
}

void launch_backsolve_update(const double* Lrow, long ld, const double* gamma_k, double* rho, int ncolblocks,
                             cudaStream_t s) {
    if (ncolblocks <= 0) return;
    backsolve_update_kernel<<<ncolblocks, 128, 0, s>>>(Lrow, ld, gamma_k, rho);
}

void launch_var_finish(const double* Y, long ldy, int m, int npad, const double* xraw, const double* x_mean,
                       const double* x_std, int d, const double* FtT, long ldf, const double* G, int p,
                       const int* basis_i, const int* basis_j, double sigma2, double* var, cudaStream_t s,
                       double* U) {
    const size_t smem = 8 * static_cast<size_t>(d + p) * sizeof(double);
    var_finish_kernel<<<(m + 7) / 8, 256, smem, s>>>(Y, ldy, m, npad, xraw, x_mean, x_std, d, FtT, ldf, G, p, basis_i,
                                                     basis_j, sigma2, var, U);
}
