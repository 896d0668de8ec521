// Fused pairwise-distance + correlation kernels (K1: R(theta) lower block-triangle,
// K2: cross-correlation c(x*, X) fused with the mean / gamma GEMV of predict).
//
// Reference being replaced: DiffMatrix::new (gp/src/utils.rs:80-104) +
// CorrelationModel::value (gp/src/correlation_models.rs:91-104, 185-196, 277-353,
// 446-523) + the scatter into R (gp/src/algorithm.rs:997-1001), and for K2
// pairwise_differences (utils.rs:110-131) + value + `.dot(gamma)`
// (algorithm.rs:253-263, 372-380).  The (P x d) difference table is never formed.
//
// Row tiles of X are contiguous in HBM (row-major n x d), so they are staged
// into shared memory with one 1-D TMA bulk copy each (cp.async.bulk ->
// SASS UBLKCP) completing on an mbarrier; the column tile is then transposed
// in shared memory so that the 16 lanes of a half-warp read 16 consecutive
// double2 (conflict free), and results leave as coalesced 16-byte stores.
#include <cstdlib>

#include "common.cuh"
#include "../../include/egobox_gpu.h"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

template <int CORR>
__device__ __forceinline__ void pair_term(const CorrTerm& t, double dx, double& acc, double& prod) {
    if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) {
        acc += t.k1 * (dx * dx);
    } else if (CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) {
        acc += t.k1 * fabs(dx);
    } else if (CORR == EGX_CORR_MATERN32) {
        const double ad = fabs(dx);
        prod *= 1.0 + t.k2 * ad;
        acc += t.k1 * ad;
    } else {
        const double ad = fabs(dx);
        prod *= (1.0 + t.k2 * ad) + (5.0 / 3.0) * ((t.k3 * dx) * dx);
        acc += t.k1 * ad;
    }
}
template <int CORR>
__device__ __forceinline__ double pair_finish(double acc, double prod) {
    if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) return exp(-0.5 * acc);
    if (CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) return exp(-acc);
    if (CORR == EGX_CORR_MATERN32) return prod * exp(-1.7320508075688772 * acc);
    return prod * exp(-2.23606797749979 * acc);
}

// Pre-scaled coordinates (r02): every (dimension, component) term t gets its own copy of the coordinate, multiplied by the
// term's weight -- u_t = c_t x_dim(t), c_t = sqrt(k1) (squared exponential), k1 (absolute exponential), sqrt(3) tw / sqrt(5) tw
// (Matern) -- so that a = |u_i - u_j| is already the argument of the kernel:
//   squared exponential  acc += a^2                      (DADD + DFMA instead of DADD + DMUL + DFMA)
//   Matern-5/2           prod *= 1 + a (1 + a / 3), acc += a   (5 fp64 instructions per pair and term instead of 7)
// and the square roots 3 / 5 leave pair_finish.  The kernels are bound by fp64 issue, so this is
// the lever; the results differ from the unscaled form by rounding only (|u| ulps instead of |x| ulps in the difference).
template <int CORR>
__device__ __forceinline__ double term_coef(const CorrTerm& t) {
    if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) return sqrt(t.k1);
    if (CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) return t.k1;
    return t.k2;
}
template <int CORR>
__device__ __forceinline__ void pair_term_s(double du, double& acc, double& prod) {
    if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) {
        acc = fma(du, du, acc);
    } else if (CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) {
        acc += fabs(du);
    } else if (CORR == EGX_CORR_MATERN32) {
        const double a = fabs(du);
        prod *= 1.0 + a;
        acc += a;
    } else {
        const double a = fabs(du);
        prod *= fma(a, fma(a, 1.0 / 3.0, 1.0), 1.0);
        acc += a;
    }
}
template <int CORR>
__device__ __forceinline__ double pair_finish_s(double acc, double prod) {
    if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) return exp(-0.5 * acc);
    if (CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) return exp(-acc);
    return prod * exp(-acc);
}
// scaled, term-major copy of a 64-row coordinate tile: out[t][r] = c_t X[r][dim(t)]
template <int CORR, int ROWS = EGX_CT>
__device__ __forceinline__ void scale_tile(const double* __restrict__ Xrow, const CorrTerm* __restrict__ terms, int nterms, int d,
                                           double* __restrict__ out, int tid) {
    for (int e = tid; e < nterms * ROWS; e += 256) {
        const int t = e / ROWS, r = e - t * ROWS;
        out[e] = term_coef<CORR>(terms[t]) * Xrow[r * d + terms[t].dim];
    }
}

// Each of the 256 threads owns a 4 x 4 patch of the 64 x 64 tile:
// rows ty + 16*ri, columns 2*tx + 32*cj + {0,1}.  XiS / XjS: scaled term-major tiles [nterms][64].
template <int CORR, int STRIDE = EGX_CT>
__device__ __forceinline__ void tile_values(const double* __restrict__ XiS, const double* __restrict__ XjS, int nterms, int ty,
                                            int tx, double (&out)[4][4]) {
    double acc[4][4], prod[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            acc[a][b] = 0.0;
            prod[a][b] = 1.0;
        }
    for (int t = 0; t < nterms; ++t) {
        double xi[4];
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) xi[ri] = XiS[t * STRIDE + ty + 16 * ri];
        const double2 xa = *reinterpret_cast<const double2*>(&XjS[t * STRIDE + 2 * tx]);
        const double2 xb = *reinterpret_cast<const double2*>(&XjS[t * STRIDE + 2 * tx + 32]);
        const double xj[4] = {xa.x, xa.y, xb.x, xb.y};
#pragma unroll
        for (int ri = 0; ri < 4; ++ri)
#pragma unroll
            for (int c = 0; c < 4; ++c) pair_term_s<CORR>(xi[ri] - xj[c], acc[ri][c], prod[ri][c]);
    }
#pragma unroll
    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
        for (int c = 0; c < 4; ++c) out[ri][c] = pair_finish_s<CORR>(acc[ri][c], prod[ri][c]);
}

__device__ __forceinline__ void tri_decode(int t, int& r, int& c) {
    int rr = static_cast<int>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
    while ((rr + 1) * (rr + 2) / 2 <= t) ++rr;
    while (rr * (rr + 1) / 2 > t) --rr;
    r = rr;
    c = t - rr * (rr + 1) / 2;
}

// ---------------------------------------------------------------------------
// K1: R(theta), every 64 x 64 tile inside the lower 128-block triangle.
// X is the zero-padded (npad x d) normalised training set.
// ---------------------------------------------------------------------------
template <int CORR>
__global__ void __launch_bounds__(256)
    corr_build_kernel(const double* __restrict__ X, int n, int d, const CorrTerm* __restrict__ gterms,
                      int nterms, double* __restrict__ M, long ld, double diag_value, double scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Xi = reinterpret_cast<double*>(smem_raw + 16);
    double* Xj = Xi + EGX_CT * d;
    double* XiS = Xj + EGX_CT * d;
    double* XjS = XiS + EGX_CT * nterms;
    CorrTerm* terms = reinterpret_cast<CorrTerm*>(XjS + EGX_CT * nterms);

    const int tid = threadIdx.x;
    const int pair = blockIdx.x >> 2, sub = blockIdx.x & 3;
    int R, C;
    tri_decode(pair, R, C);
    const int i0 = (2 * R + (sub >> 1)) * EGX_CT;
    const int j0 = (2 * C + (sub & 1)) * EGX_CT;
    const uint32_t bytes = static_cast<uint32_t>(EGX_CT * d * sizeof(double));

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, 2 * bytes);
        bulk_g2s(Xi, X + static_cast<long>(i0) * d, bytes, bar);
        bulk_g2s(Xj, X + static_cast<long>(j0) * d, bytes, bar);
    }
    for (int t = tid; t < nterms; t += 256) terms[t] = gterms[t];
    __syncthreads();                 // the term list is read by every thread below
    mbar_wait(bar, 0);
    scale_tile<CORR>(Xi, terms, nterms, d, XiS, tid);
    scale_tile<CORR>(Xj, terms, nterms, d, XjS, tid);
    __syncthreads();

    const int ty = tid >> 4, tx = tid & 15;
    double v[4][4];
    tile_values<CORR>(XiS, XjS, nterms, ty, tx, v);

#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {
        const int i = i0 + ty + 16 * ri;
#pragma unroll
        for (int cj = 0; cj < 2; ++cj) {
            const int j = j0 + 2 * tx + 32 * cj;
            double a = scale * v[ri][2 * cj], b = scale * v[ri][2 * cj + 1];
            if (i >= n || j >= n) a = 0.0;
            if (i >= n || j + 1 >= n) b = 0.0;
            if (i == j) a = (i < n) ? diag_value : 1.0;
            if (i == j + 1) b = (i < n) ? diag_value : 1.0;
            *reinterpret_cast<double2*>(&M[static_cast<long>(i) * ld + j]) = make_double2(a, b);
        }
    }
}

// K1, one CTA per 128 x 128 block (r02): the two 128-row coordinate strips are loaded and scaled ONCE for the four 64 x 64
// sub-tiles.  With a CTA per sub-tile the prologue (bulk copy + scaling + three barriers, ~2.5 k clk) stood against 1.9 k clk
// of arithmetic for the squared exponential at d = 6 and 4.3 k for Matern-5/2 at d = 10 (profiles/r02/y12_corr_*.txt).  Used
// while the strips fit beside a second CTA (d + nterms <= 46); larger dimensions keep the per-sub-tile kernel.
template <int CORR>
__global__ void __launch_bounds__(256)
    corr_build128_kernel(const double* __restrict__ X, int n, int d, const CorrTerm* __restrict__ gterms,
                         int nterms, double* __restrict__ M, long ld, double diag_value, double scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Xi = reinterpret_cast<double*>(smem_raw + 16);
    double* Xj = Xi + EGX_NB * d;
    double* XiS = Xj + EGX_NB * d;
    double* XjS = XiS + EGX_NB * nterms;
    CorrTerm* terms = reinterpret_cast<CorrTerm*>(XjS + EGX_NB * nterms);

    const int tid = threadIdx.x;
    int R, C;
    tri_decode(blockIdx.x, R, C);
    const int I0 = R * EGX_NB, J0 = C * EGX_NB;
    const uint32_t bytes = static_cast<uint32_t>(EGX_NB * d * sizeof(double));

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, 2 * bytes);
        bulk_g2s(Xi, X + static_cast<long>(I0) * d, bytes, bar);
        bulk_g2s(Xj, X + static_cast<long>(J0) * d, bytes, bar);
    }
    for (int t = tid; t < nterms; t += 256) terms[t] = gterms[t];
    __syncthreads();                 // the term list is read by every thread below
    mbar_wait(bar, 0);
    scale_tile<CORR, EGX_NB>(Xi, terms, nterms, d, XiS, tid);
    scale_tile<CORR, EGX_NB>(Xj, terms, nterms, d, XjS, tid);
    __syncthreads();

    const int ty = tid >> 4, tx = tid & 15;
#pragma unroll 1
    for (int sub = 0; sub < 4; ++sub) {
        const int i0 = I0 + (sub >> 1) * EGX_CT, j0 = J0 + (sub & 1) * EGX_CT;
        double v[4][4];
        tile_values<CORR, EGX_NB>(XiS + (sub >> 1) * EGX_CT, XjS + (sub & 1) * EGX_CT, nterms, ty, tx, v);
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
            const int i = i0 + ty + 16 * ri;
#pragma unroll
            for (int cj = 0; cj < 2; ++cj) {
                const int j = j0 + 2 * tx + 32 * cj;
                double a = scale * v[ri][2 * cj], b = scale * v[ri][2 * cj + 1];
                if (i >= n || j >= n) a = 0.0;
                if (i >= n || j + 1 >= n) b = 0.0;
                if (i == j) a = (i < n) ? diag_value : 1.0;
                if (i == j + 1) b = (i < n) ? diag_value : 1.0;
                *reinterpret_cast<double2*>(&M[static_cast<long>(i) * ld + j]) = make_double2(a, b);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K2: cross-correlation of a block of 64 prediction points against all
// training points, fused with yhat = f(x) beta + c(x, X) gamma.  Optionally
// stores c into Y (row-major, ldy) for the variance TRSM.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double basis_value(const double* x, int bi, int bj) {
    const double a = bi < 0 ? 1.0 : x[bi];
    const double b = bj < 0 ? 1.0 : x[bj];
    return a * b;
}

template <int CORR>
__global__ void __launch_bounds__(256)
    cross_corr_kernel(const double* __restrict__ xraw, int m, const double* __restrict__ x_mean,
                      const double* __restrict__ x_std, const double* __restrict__ X, int n, int npad, int d,
                      const CorrTerm* __restrict__ gterms, int nterms, const double* __restrict__ gamma,
                      const double* __restrict__ beta, const int* __restrict__ basis_i,
                      const int* __restrict__ basis_j, int p, double y_mean, double y_std,
                      double* __restrict__ Y, long ldy, double* __restrict__ yout, double scale) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    double* Xp = reinterpret_cast<double*>(smem_raw + 16);
    double* Xj = Xp + EGX_CT * d;
    double* XpS = Xj + EGX_CT * d;
    double* XjS = XpS + EGX_CT * nterms;
    double* gam = XjS + EGX_CT * nterms;
    double* ysum = gam + EGX_CT;
    CorrTerm* terms = reinterpret_cast<CorrTerm*>(ysum + EGX_CT);

    const int tid = threadIdx.x;
    const int i0 = blockIdx.x * EGX_CT;
    const uint32_t bytes = static_cast<uint32_t>(EGX_CT * d * sizeof(double));

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    for (int e = tid; e < EGX_CT * d; e += 256) {
        const int r = e / d, c = e - r * d;
        const int i = i0 + r;
        Xp[e] = (i < m) ? (xraw[static_cast<long>(i) * d + c] - x_mean[c]) / x_std[c] : 0.0;
    }
    for (int t = tid; t < nterms; t += 256) terms[t] = gterms[t];
    __syncthreads();
    scale_tile<CORR>(Xp, terms, nterms, d, XpS, tid);

    const int ty = tid >> 4, tx = tid & 15;
    double yacc[4] = {0.0, 0.0, 0.0, 0.0};
    const int ntiles = npad / EGX_CT;
    for (int jt = 0; jt < ntiles; ++jt) {
        const int j0 = jt * EGX_CT;
        if (tid == 0) {
            mbar_expect_tx(bar, bytes);
            bulk_g2s(Xj, X + static_cast<long>(j0) * d, bytes, bar);
        }
        if (tid < EGX_CT) gam[tid] = (gamma != nullptr && j0 + tid < n) ? gamma[j0 + tid] : 0.0;
        mbar_wait(bar, jt & 1);
        scale_tile<CORR>(Xj, terms, nterms, d, XjS, tid);
        __syncthreads();

        double v[4][4];
        tile_values<CORR>(XpS, XjS, nterms, ty, tx, v);
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
            const int i = i0 + ty + 16 * ri;
#pragma unroll
            for (int cj = 0; cj < 2; ++cj) {
                const int jl = 2 * tx + 32 * cj;
                double a = scale * v[ri][2 * cj], b = scale * v[ri][2 * cj + 1];
                if (i >= m || j0 + jl >= n) a = 0.0;
                if (i >= m || j0 + jl + 1 >= n) b = 0.0;
                yacc[ri] += a * gam[jl] + b * gam[jl + 1];
                if (Y != nullptr)
                    *reinterpret_cast<double2*>(&Y[static_cast<long>(i) * ldy + j0 + jl]) = make_double2(a, b);
            }
        }
        __syncthreads();   // Xj / XjS / gam are overwritten by the next tile
    }

    if (yout != nullptr) {
        // reduce the 16 column-lanes of each row group (lanes tx = 0..15 of a half-warp)
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
            double s = yacc[ri];
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (tx == 0) ysum[ty + 16 * ri] = s;
        }
        __syncthreads();
        if (tid < EGX_CT && i0 + tid < m) {
            const double* x = Xp + tid * d;
            double f = 0.0;
            for (int l = 0; l < p; ++l) f += basis_value(x, basis_i[l], basis_j[l]) * beta[l];
            yout[i0 + tid] = (f + ysum[tid]) * y_std + y_mean;
        }
    }
}

// ---------------------------------------------------------------------------
// Batched prediction gradients  d yhat / d x  (gp/src/algorithm.rs:510-550 `predict_gradients` via
// `predict_jacobian`; kernels' jacobians correlation_models.rs:106-122, 198-214, 288-298+355-412,
// 457-468+525-586).  One warp per prediction point, lanes over training points; per pair the kernel
// value r and, per input dimension k, the logarithmic derivative  s_k = (dr/dx_k) / r  are formed in
// one pass over the term list (for the Matern models: sum_l f'_kl / f_kl - sqrt(nu') tw_k sign, every
// factor f >= 1 so the division is safe), and  g_k += gamma_j r s_k  is accumulated in registers.
// ---------------------------------------------------------------------------
template <int CORR, int DMAX>
__global__ void __launch_bounds__(256)
    predict_grad_kernel(const double* __restrict__ xraw, int m, const double* __restrict__ x_mean,
                        const double* __restrict__ x_std, const double* __restrict__ X, int n, int npad, int d,
                        const CorrTerm* __restrict__ gterms, int nterms, const double* __restrict__ gamma,
                        const double* __restrict__ beta, const int* __restrict__ basis_i,
                        const int* __restrict__ basis_j, int p, double y_std, double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* XjT = reinterpret_cast<double*>(smem_raw);          // [d][64]
    double* gam = XjT + EGX_CT * d;                               // [64]
    double* xp = gam + EGX_CT;                                    // [8][d] normalised prediction points
    CorrTerm* terms = reinterpret_cast<CorrTerm*>(xp + 8 * d);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * 8 + warp;
    for (int t = tid; t < nterms; t += 256) terms[t] = gterms[t];
    for (int e = tid; e < 8 * d; e += 256) {
        const int w = e / d, c = e - w * d;
        const int ii = blockIdx.x * 8 + w;
        xp[e] = (ii < m) ? (xraw[static_cast<long>(ii) * d + c] - x_mean[c]) / x_std[c] : 0.0;
    }
    double g[DMAX];
#pragma unroll
    for (int k = 0; k < DMAX; ++k) g[k] = 0.0;
    const double* xi = xp + warp * d;
    const double sq = (CORR == EGX_CORR_MATERN32) ? 1.7320508075688772 : 2.23606797749979;
    for (int j0 = 0; j0 < npad; j0 += EGX_CT) {
        __syncthreads();     // previous tile fully consumed (also covers the terms / xp staging the first time)
        for (int e = tid; e < EGX_CT * d; e += 256) {
            const int r = e / d, c = e - r * d;
            XjT[c * EGX_CT + r] = X[static_cast<long>(j0 + r) * d + c];
        }
        if (tid < EGX_CT) gam[tid] = (j0 + tid < n) ? gamma[j0 + tid] : 0.0;
        __syncthreads();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int jl = lane + 32 * h;
            const double gj = gam[jl];
            double s[DMAX];
#pragma unroll
            for (int k = 0; k < DMAX; ++k) s[k] = 0.0;
            double acc = 0.0, prod = 1.0;
            for (int t = 0; t < nterms; ++t) {
                const CorrTerm tm = terms[t];
                const double dx = xi[tm.dim] - XjT[tm.dim * EGX_CT + jl];
                const double ad = fabs(dx), sg = copysign(1.0, dx);
                double sk;
                if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) {
                    acc += tm.k1 * (dx * dx);
                    sk = -tm.k1 * dx;
                } else if (CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) {
                    acc += tm.k1 * ad;
                    sk = -tm.k1 * sg;
                } else if (CORR == EGX_CORR_MATERN32) {
                    const double f = 1.0 + tm.k2 * ad;
                    prod *= f;
                    acc += tm.k1 * ad;
                    sk = sg * (tm.k2 / f - sq * tm.k1);
                } else {
                    const double f = (1.0 + tm.k2 * ad) + (5.0 / 3.0) * ((tm.k3 * dx) * dx);
                    prod *= f;
                    acc += tm.k1 * ad;
                    sk = sg * ((tm.k2 + (10.0 / 3.0) * tm.k3 * ad) / f - sq * tm.k1);
                }
#pragma unroll
                for (int k = 0; k < DMAX; ++k)
                    if (k == tm.dim) s[k] += sk;
            }
            const double r = gj * pair_finish<CORR>(acc, prod);
#pragma unroll
            for (int k = 0; k < DMAX; ++k) g[k] += r * s[k];
        }
    }
    if (i >= m) return;
#pragma unroll
    for (int k = 0; k < DMAX; ++k) {
        double v = g[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        g[k] = v;
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < DMAX; ++k) {
            if (k < d) {
                // (d f / d x_k)^T beta with f_l = v(bi) v(bj)  (mean_models.rs jacobians)
                double df = 0.0;
                for (int l = 0; l < p; ++l) {
                    const int bi = basis_i[l], bj = basis_j[l];
                    double t = 0.0;
                    if (bi == k) t += (bj < 0) ? 1.0 : xi[bj];
                    if (bj == k) t += (bi < 0) ? 1.0 : xi[bi];
                    df += t * beta[l];
                }
                out[static_cast<long>(i) * d + k] = (df + g[k]) * y_std / x_std[k];
            }
        }
    }
}

// F^T rows (p x npad) followed by ynorm as row p: the right-hand sides that are
// carried through the Cholesky as extra rows of the matrix (forward solves fused).
__global__ void mean_basis_rows_kernel(const double* __restrict__ X, int n, int npad, int d,
                                       const int* __restrict__ basis_i, const int* __restrict__ basis_j, int p,
                                       const double* __restrict__ ynorm, double* __restrict__ FyT, long ld) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npad) return;
    const double* x = X + static_cast<long>(j) * d;
    for (int l = 0; l < p; ++l) FyT[static_cast<long>(l) * ld + j] = (j < n) ? basis_value(x, basis_i[l], basis_j[l]) : 0.0;
    FyT[static_cast<long>(p) * ld + j] = (j < n) ? ynorm[j] : 0.0;
}

// two raw 64 x d tiles (TMA destinations) + two scaled term-major tiles [nterms][64] + the term list
size_t corr_build_smem(int d, int nterms) {
    return 16 + 2 * EGX_CT * (static_cast<size_t>(d) + nterms) * sizeof(double) + nterms * sizeof(CorrTerm);
}
size_t cross_corr_smem(int d, int nterms) {
    return 16 + 2 * EGX_CT * (static_cast<size_t>(d) + nterms) * sizeof(double) + 2 * EGX_CT * sizeof(double) + nterms * sizeof(CorrTerm);
}

template <typename K>
void set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
}

}  // namespace

void launch_corr_build(int corr, const double* X, int n, int npad, int d, const CorrTerm* terms, int nterms,
                       double* M, long ld, double diag_value, cudaStream_t s, double scale) {
    const int T = npad / EGX_NB;
    // one CTA per 128 x 128 block while the two 128-row strips (raw + scaled) leave room for a second CTA on the SM
    // (EGX_CORR_TILE=64 keeps the r01 form: one CTA per 64 x 64 sub-tile)
    static const int force64 = getenv("EGX_CORR_TILE") != nullptr && atoi(getenv("EGX_CORR_TILE")) == 64;
    const size_t smem128 = 16 + 2 * EGX_NB * (static_cast<size_t>(d) + nterms) * sizeof(double) + nterms * sizeof(CorrTerm);
    const bool big = !force64 && smem128 <= 96 * 1024;
    const int grid = (big ? 1 : 4) * (T * (T + 1) / 2);
    const size_t smem = big ? smem128 : corr_build_smem(d, nterms);
#define EGX_LAUNCH_K1(CK)                                                                         \
    if (big) {                                                                                    \
        set_smem(corr_build128_kernel<CK>, smem);                                                 \
        corr_build128_kernel<CK><<<grid, 256, smem, s>>>(X, n, d, terms, nterms, M, ld, diag_value, scale); \
    } else {                                                                                      \
        set_smem(corr_build_kernel<CK>, smem);                                                    \
        corr_build_kernel<CK><<<grid, 256, smem, s>>>(X, n, d, terms, nterms, M, ld, diag_value, scale); \
    }
    switch (corr) {
        case EGX_CORR_SQUARED_EXPONENTIAL: EGX_LAUNCH_K1(EGX_CORR_SQUARED_EXPONENTIAL) break;
        case EGX_CORR_ABSOLUTE_EXPONENTIAL: EGX_LAUNCH_K1(EGX_CORR_ABSOLUTE_EXPONENTIAL) break;
        case EGX_CORR_MATERN32: EGX_LAUNCH_K1(EGX_CORR_MATERN32) break;
        default: EGX_LAUNCH_K1(EGX_CORR_MATERN52) break;
    }
#undef EGX_LAUNCH_K1
}

void launch_cross_corr(int corr, const double* xraw, int m, int mpad, const double* x_mean, const double* x_std,
                       const double* X, int n, int npad, int d, const CorrTerm* terms, int nterms,
                       const double* gamma, const double* beta, const int* basis_i, const int* basis_j, int p,
                       double y_mean, double y_std, double* Y, long ldy, double* yout, cudaStream_t s, double scale) {
    const int grid = mpad / EGX_CT;
    const size_t smem = cross_corr_smem(d, nterms);
#define EGX_LAUNCH_K2(CK)                                                                               \
    set_smem(cross_corr_kernel<CK>, smem);                                                              \
    cross_corr_kernel<CK><<<grid, 256, smem, s>>>(xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms, \
                                                  gamma, beta, basis_i, basis_j, p, y_mean, y_std, Y, ldy, yout, scale);
    switch (corr) {
        case EGX_CORR_SQUARED_EXPONENTIAL: EGX_LAUNCH_K2(EGX_CORR_SQUARED_EXPONENTIAL) break;
        case EGX_CORR_ABSOLUTE_EXPONENTIAL: EGX_LAUNCH_K2(EGX_CORR_ABSOLUTE_EXPONENTIAL) break;
        case EGX_CORR_MATERN32: EGX_LAUNCH_K2(EGX_CORR_MATERN32) break;
        default: EGX_LAUNCH_K2(EGX_CORR_MATERN52) break;
    }
#undef EGX_LAUNCH_K2
}

void launch_mean_basis_rows(const double* X, int n, int npad, int d, const int* basis_i, const int* basis_j, int p,
                            const double* ynorm_dev, double* FyT, long ld, cudaStream_t s) {
    mean_basis_rows_kernel<<<(npad + 255) / 256, 256, 0, s>>>(X, n, npad, d, basis_i, basis_j, p, ynorm_dev, FyT, ld);
}

template <int CORR>
static void launch_pg(int dmax, int grid, size_t smem, cudaStream_t s, const double* xraw, int m, const double* x_mean,
                      const double* x_std, const double* X, int n, int npad, int d, const CorrTerm* terms, int nterms,
                      const double* gamma, const double* beta, const int* bi, const int* bj, int p, double y_std,
                      double* out) {
    if (dmax <= 8) {
        set_smem(predict_grad_kernel<CORR, 8>, smem);
        predict_grad_kernel<CORR, 8><<<grid, 256, smem, s>>>(xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms, gamma,
                                                            beta, bi, bj, p, y_std, out);
    } else if (dmax <= 16) {
        set_smem(predict_grad_kernel<CORR, 16>, smem);
        predict_grad_kernel<CORR, 16><<<grid, 256, smem, s>>>(xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms, gamma,
                                                             beta, bi, bj, p, y_std, out);
    } else {
        set_smem(predict_grad_kernel<CORR, 32>, smem);
        predict_grad_kernel<CORR, 32><<<grid, 256, smem, s>>>(xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms, gamma,
                                                             beta, bi, bj, p, y_std, out);
    }
}

// d <= 32 (caller checks)
void launch_predict_grad(int corr, const double* xraw, int m, const double* x_mean, const double* x_std, const double* X,
                         int n, int npad, int d, const CorrTerm* terms, int nterms, const double* gamma,
                         const double* beta, const int* basis_i, const int* basis_j, int p, double y_std, double* out,
                         cudaStream_t s) {
    const int grid = (m + 7) / 8;
    const size_t smem = (static_cast<size_t>(EGX_CT) * d + EGX_CT + 8 * d) * sizeof(double) + nterms * sizeof(CorrTerm);
    switch (corr) {
        case EGX_CORR_SQUARED_EXPONENTIAL:
            launch_pg<EGX_CORR_SQUARED_EXPONENTIAL>(d, grid, smem, s, xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms,
                                                    gamma, beta, basis_i, basis_j, p, y_std, out);
            break;
        case EGX_CORR_ABSOLUTE_EXPONENTIAL:
            launch_pg<EGX_CORR_ABSOLUTE_EXPONENTIAL>(d, grid, smem, s, xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms,
                                                     gamma, beta, basis_i, basis_j, p, y_std, out);
            break;
        case EGX_CORR_MATERN32:
            launch_pg<EGX_CORR_MATERN32>(d, grid, smem, s, xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms, gamma,
                                         beta, basis_i, basis_j, p, y_std, out);
            break;
        default:
            launch_pg<EGX_CORR_MATERN52>(d, grid, smem, s, xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms, gamma,
                                         beta, basis_i, basis_j, p, y_std, out);
            break;
    }
}
