// K12: gradient of the reduced likelihood with respect to theta, in closed form.
//
// The reference has no such function (`objfn` ignores its `_gradient` argument, gp/src/algorithm.rs:880, COBYLA is
// derivative free); SURVEY.md section 8 (f)-4 lists it as the next building block.  With
//   rlf = -n log10 sigma2 - log10 det R                       (algorithm.rs:1039-1043)
// beta the generalised least-squares minimiser (its own derivative drops out) and gamma = R^-1 (y - F beta) (:1034):
//   d rlf / d theta_l = [ gamma^T (dR/dtheta_l) gamma / sigma2 - tr(R^-1 dR/dtheta_l) ] / ln 10
//                     = (2 / ln 10) sum_{i > j} r_ij (d ln r_ij / d theta_l) (gamma_i gamma_j / sigma2 - (R^-1)_ij)
// (the diagonal of R is the constant 1 + nugget).  R^-1 = W W^T with W = L^-T comes from the blocked multi-RHS sweep
// applied to the identity and one SYRK (gp_context.cu); this file holds the pair kernel: like K1 it walks the 64 x 64
// tiles of the lower block triangle, recomputes r_ij from the coordinates (R itself was overwritten by L), forms the
// h logarithmic derivatives of the pair in one pass over the (dimension, component) term list, and reduces
// h partial sums per CTA; a second kernel adds the per-CTA partials in a fixed order, so the result is reproducible.
#include "common.cuh"
#include "../../include/egobox_gpu.h"

namespace {

__device__ __forceinline__ void tg_tri_decode(int t, int& r, int& c) {
    int rr = static_cast<int>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
    while ((rr + 1) * (rr + 2) / 2 <= t) ++rr;
    while (rr * (rr + 1) / 2 > t) --rr;
    r = rr;
    c = t - rr * (rr + 1) / 2;
}

// C holds -R^-1 (the SYRK kernels subtract), lower block triangle.
template <int CORR, int HMAX>
__global__ void __launch_bounds__(256)
    theta_grad_kernel(const double* __restrict__ X, int n, int d, const ThetaGradTerm* __restrict__ gterms, int nterms,
                      const double* __restrict__ C, long ldc, const double* __restrict__ gamma,
                      const EvalResult* __restrict__ res, int h, double* __restrict__ partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* Xi = reinterpret_cast<double*>(smem_raw);          // [64][d]
    double* XjT = Xi + EGX_CT * d;                             // [d][64]
    double* gi = XjT + EGX_CT * d;                             // [64]
    double* gj = gi + EGX_CT;                                  // [64]
    double* red = gj + EGX_CT;                                 // [8][HMAX]
    ThetaGradTerm* terms = reinterpret_cast<ThetaGradTerm*>(red + 8 * HMAX);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pair = blockIdx.x >> 2, sub = blockIdx.x & 3;
    int R, Cc;
    tg_tri_decode(pair, R, Cc);
    const int i0 = (2 * R + (sub >> 1)) * EGX_CT;
    const int j0 = (2 * Cc + (sub & 1)) * EGX_CT;
    double* out = partial + static_cast<long>(blockIdx.x) * h;
    if (i0 < j0 || j0 >= n) {                                  // upper 64-tile of a diagonal block / padding only
        if (tid < h) out[tid] = 0.0;
        return;
    }
    for (int e = tid; e < EGX_CT * d; e += 256) {
        const int r = e / d, c = e - r * d;
        Xi[e] = X[static_cast<long>(i0) * d + e];
        XjT[c * EGX_CT + r] = X[static_cast<long>(j0) * d + e];
    }
    if (tid < EGX_CT) {
        gi[tid] = (i0 + tid < n) ? gamma[i0 + tid] : 0.0;
        gj[tid] = (j0 + tid < n) ? gamma[j0 + tid] : 0.0;
    }
    for (int t = tid; t < nterms; t += 256) terms[t] = gterms[t];
    __syncthreads();

    const double inv_s2 = 1.0 / res->sigma2;
    const double sq = (CORR == EGX_CORR_MATERN32) ? 1.7320508075688772 : 2.23606797749979;
    double g[HMAX];
#pragma unroll
    for (int k = 0; k < HMAX; ++k) g[k] = 0.0;

#pragma unroll 1
    for (int it = 0; it < (EGX_CT * EGX_CT) / 256; ++it) {
        const int e = tid + 256 * it;
        const int r = e >> 6, cc = e & 63;
        const int i = i0 + r, j = j0 + cc;
        if (i >= n || j >= i) continue;
        double s[HMAX];
#pragma unroll
        for (int k = 0; k < HMAX; ++k) s[k] = 0.0;
        double acc = 0.0, prod = 1.0;
        const double* xi = Xi + r * d;
        for (int t = 0; t < nterms; ++t) {
            const ThetaGradTerm tm = terms[t];
            const double dx = xi[tm.dim] - XjT[tm.dim * EGX_CT + cc];
            double sk;
            if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) {
                const double d2 = dx * dx;
                acc += tm.tw * d2;                 // tw = (theta_l W_jl)^2
                sk = tm.a * d2;                    // a  = -theta_l W_jl^2
            } else if (CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) {
                const double ad = fabs(dx);
                acc += tm.tw * ad;                 // tw = theta_l |W_jl|
                sk = tm.a * ad;                    // a  = -|W_jl|
            } else if (CORR == EGX_CORR_MATERN32) {
                const double ad = fabs(dx);
                const double v = tm.tw * ad;       // tw = theta_l |W_jl|, a = |W_jl|
                const double f = 1.0 + sq * v;
                prod *= f;
                acc += v;
                sk = (tm.a * ad) * (-3.0 * v / f);
            } else {
                const double ad = fabs(dx);
                const double v = tm.tw * ad;
                const double f = (1.0 + sq * v) + (5.0 / 3.0) * (v * v);
                prod *= f;
                acc += v;
                sk = (tm.a * ad) * (-(5.0 / 3.0) * v * (1.0 + sq * v) / f);
            }
#pragma unroll
            for (int k = 0; k < HMAX; ++k)
                if (k == tm.comp) s[k] += sk;
        }
        double rv;
        if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) rv = exp(-0.5 * acc);
        else if (CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) rv = exp(-acc);
        else rv = prod * exp(-sq * acc);
        const double wgt = rv * (gi[r] * gj[cc] * inv_s2 + C[static_cast<long>(i) * ldc + j]);
#pragma unroll
        for (int k = 0; k < HMAX; ++k) g[k] += wgt * s[k];
    }

#pragma unroll
    for (int k = 0; k < HMAX; ++k) {
        double v = g[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp * HMAX + k] = v;
    }
    __syncthreads();
    if (tid < h) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w * HMAX + tid];
        out[tid] = v;
    }
}

// grad[k] = (2 / ln 10) * sum over CTAs of partial[cta][k]; one CTA per component, fixed summation order
__global__ void __launch_bounds__(256)
    theta_grad_reduce_kernel(const double* __restrict__ partial, int nblocks, int h, double* __restrict__ grad) {
    __shared__ double sm[256];
    const int k = blockIdx.x, tid = threadIdx.x;
    double v = 0.0;
    for (int b = tid; b < nblocks; b += 256) v += partial[static_cast<long>(b) * h + k];
    sm[tid] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) sm[tid] += sm[tid + o];
        __syncthreads();
    }
    if (tid == 0) grad[k] = sm[0] * (2.0 / 2.302585092994046);
}

// out = W rho for the upper-triangular W = L^-T (row i: columns i .. n-1): gamma = L^-T rho (algorithm.rs:1034) without the
// 2 T launches of the blocked back substitution, W being at hand.  One warp per row.
__global__ void __launch_bounds__(256)
    upper_gemv_kernel(const double* __restrict__ W, long ld, int n, const double* __restrict__ rho, double* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const double* w = W + static_cast<long>(row) * ld;
    double acc = 0.0;
    for (int k = (row & ~31) + lane; k < n; k += 32)
        if (k >= row) acc += w[k] * rho[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[row] = acc;
}

__global__ void set_identity_kernel(double* __restrict__ A, long ld, int npad) {
    const long e = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long total = static_cast<long>(npad) * npad;
    if (e >= total) return;
    const long i = e / npad, j = e - i * npad;
    A[i * ld + j] = (i == j) ? 1.0 : 0.0;
}

size_t theta_grad_smem(int d, int nterms, int hmax) {
    return (2 * static_cast<size_t>(EGX_CT) * d + 2 * EGX_CT + 8 * hmax) * sizeof(double) + nterms * sizeof(ThetaGradTerm);
}

template <int CORR, int HMAX>
void launch_tg(int grid, cudaStream_t s, const double* X, int n, int d, const ThetaGradTerm* terms, int nterms,
               const double* C, long ldc, const double* gamma, const EvalResult* res, int h, double* partial) {
    const size_t smem = theta_grad_smem(d, nterms, HMAX);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(theta_grad_kernel<CORR, HMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    theta_grad_kernel<CORR, HMAX><<<grid, 256, smem, s>>>(X, n, d, terms, nterms, C, ldc, gamma, res, h, partial);
}

template <int CORR>
void launch_tg_h(int grid, cudaStream_t s, const double* X, int n, int d, const ThetaGradTerm* terms, int nterms,
                 const double* C, long ldc, const double* gamma, const EvalResult* res, int h, double* partial) {
    if (h <= 8) launch_tg<CORR, 8>(grid, s, X, n, d, terms, nterms, C, ldc, gamma, res, h, partial);
    else if (h <= 16) launch_tg<CORR, 16>(grid, s, X, n, d, terms, nterms, C, ldc, gamma, res, h, partial);
    else launch_tg<CORR, 32>(grid, s, X, n, d, terms, nterms, C, ldc, gamma, res, h, partial);
}

}  // namespace

// One term per (dimension j, component l) with W_jl != 0 -- also where theta_l = 0, the derivative does not vanish there.
int egx_fill_theta_grad_terms(int corr, int d, int h, const double* w, const double* theta, ThetaGradTerm* t) {
    int nt = 0;
    for (int j = 0; j < d; ++j)
        for (int l = 0; l < h; ++l) {
            const double wjl = w[j * h + l];
            if (wjl == 0.0) continue;
            t[nt].dim = j;
            t[nt].comp = l;
            if (corr == EGX_CORR_SQUARED_EXPONENTIAL) {
                t[nt].a = -theta[l] * wjl * wjl;
                t[nt].tw = (theta[l] * wjl) * (theta[l] * wjl);
            } else if (corr == EGX_CORR_ABSOLUTE_EXPONENTIAL) {
                t[nt].a = -std::fabs(wjl);
                t[nt].tw = theta[l] * std::fabs(wjl);
            } else {
                t[nt].a = std::fabs(wjl);
                t[nt].tw = theta[l] * std::fabs(wjl);
            }
            ++nt;
        }
    return nt;
}

int theta_grad_blocks(int npad) {
    const int T = npad / EGX_NB;
    return 4 * (T * (T + 1) / 2);
}

// h <= 32 (caller checks).  partial: theta_grad_blocks(npad) x h doubles; grad: h doubles (device).
void launch_theta_grad(int corr, const double* X, int n, int npad, int d, const ThetaGradTerm* terms, int nterms,
                       const double* Cneg_rinv, long ldc, const double* gamma, const EvalResult* res, int h,
                       double* partial, double* grad, cudaStream_t s) {
    const int grid = theta_grad_blocks(npad);
    switch (corr) {
        case EGX_CORR_SQUARED_EXPONENTIAL:
            launch_tg_h<EGX_CORR_SQUARED_EXPONENTIAL>(grid, s, X, n, d, terms, nterms, Cneg_rinv, ldc, gamma, res, h, partial);
            break;
        case EGX_CORR_ABSOLUTE_EXPONENTIAL:
            launch_tg_h<EGX_CORR_ABSOLUTE_EXPONENTIAL>(grid, s, X, n, d, terms, nterms, Cneg_rinv, ldc, gamma, res, h, partial);
            break;
        case EGX_CORR_MATERN32:
            launch_tg_h<EGX_CORR_MATERN32>(grid, s, X, n, d, terms, nterms, Cneg_rinv, ldc, gamma, res, h, partial);
            break;
        default:
            launch_tg_h<EGX_CORR_MATERN52>(grid, s, X, n, d, terms, nterms, Cneg_rinv, ldc, gamma, res, h, partial);
            break;
    }
    theta_grad_reduce_kernel<<<h, 256, 0, s>>>(partial, grid, h, grad);
}

void launch_set_identity(double* A, long ld, int npad, cudaStream_t s) {
    const long total = static_cast<long>(npad) * npad;
    set_identity_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(A, ld, npad);
}

void launch_upper_gemv(const double* W, long ld, int n, const double* rho, double* out, cudaStream_t s) {
    upper_gemv_kernel<<<(n + 7) / 8, 256, 0, s>>>(W, ld, n, rho, out);
}
