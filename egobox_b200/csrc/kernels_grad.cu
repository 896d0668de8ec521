// Batched variance gradients  d var / d x  (gp/src/algorithm.rs:554-616 `predict_var_gradients_single`,
// :697-704 `predict_var_gradients`):
//     dvar/dx = 2 sigma2 (p4 - p2) / x_std ,   p2 = (R^-1 r)^T dr/dx ,   p4 = (B^-1 A^T)^T dA/dx^T
//     A = f(x)^T - r^T R^-1 F ,  B = F^T R^-1 F = G^T G ,  dA/dx = df/dx^T - dr/dx^T R^-1 F .
// The reference does four n x n triangular solves PER POINT; here W = C R^-1 for a whole chunk of points is
// one forward multi-RHS sweep (Y L^-T, shared with predict_var) plus one backward sweep (Y L^-1) that runs on
// the transposed factor, and the per-point reductions re-derive r and dr/dx on the fly (never stored).
#include "common.cuh"
#include "../../include/egobox_gpu.h"

namespace {

__device__ __forceinline__ double vg_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// LT block (c, r) = L block (r, c)^T for every 128-block in the lower block triangle (r >= c)
__global__ void __launch_bounds__(256) transpose_lower_kernel(const double* __restrict__ L, long ld,
                                                              double* __restrict__ LT, int T) {
    __shared__ double tile[32][33];
    // blockIdx.x enumerates 32x32 sub-tiles of the lower block triangle: pair index * 16 + sub
    const int pair = blockIdx.x >> 4, sub = blockIdx.x & 15;
    int r = static_cast<int>((sqrt(8.0 * static_cast<double>(pair) + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= pair) ++r;
    while (r * (r + 1) / 2 > pair) --r;
    const int c = pair - r * (r + 1) / 2;
    if (r >= T) return;
    const int r0 = r * EGX_NB + (sub >> 2) * 32, c0 = c * EGX_NB + (sub & 3) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int i = ty; i < 32; i += 8) tile[i][tx] = L[static_cast<long>(r0 + i) * ld + c0 + tx];
    __syncthreads();
#pragma unroll
    for (int i = ty; i < 32; i += 8) LT[static_cast<long>(c0 + i) * ld + r0 + tx] = tile[tx][i];
}

constexpr int TU_ROWS = 64;
constexpr int TU_LDX = 132;
constexpr int TU_LDL = 100;
constexpr int TU_LDD = 36;

__device__ __forceinline__ void dmma884_u(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void tu_cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// X (64 x 128 slab, in place) <- X * L_kk^-1, using the TRANSPOSED diagonal block LT_kk (upper triangular,
// row-major) and the inverted 32 x 32 diagonal sub-blocks:  for b = 3..0
//     T = X_b - sum_{b' > b} X_b' * L[b', b]        (B operand [c][k] = LT[32b + c][32b' + k])
//     X_b = T * Dinv_b                               (B operand [c][k] = Dinv_b[k][c])
__global__ void __launch_bounds__(256) trsm_rows_upper_kernel(double* __restrict__ X, long ldx,
                                                              const double* __restrict__ LT, long ldl,
                                                              const double* __restrict__ Dinv,
                                                              double* __restrict__ P) {
    extern __shared__ __align__(16) double sm[];
    double* Xs = sm;                              // [64][132]
    double* DsT = Xs + TU_ROWS * TU_LDX;          // [4][32][36], transposed inverses
    double* Lb = DsT + 4 * 32 * TU_LDD;           // [32][100]: rows 32b.. of LT, columns 32(b+1)..127
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    double* Xg = X + static_cast<long>(blockIdx.x) * TU_ROWS * ldx;
    for (int e = tid; e < TU_ROWS * 64; e += 256) {
        const int r = e >> 6, ch = e & 63;
        tu_cp_async16(&Xs[r * TU_LDX + ch * 2], Xg + static_cast<long>(r) * ldx + ch * 2);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int e = tid; e < 4096; e += 256) {
        const int b = e >> 10, k = (e >> 5) & 31, c = e & 31;
        DsT[(b * 32 + c) * TU_LDD + k] = Dinv[e];      // Dinv[b][k][c] -> DsT[b][c][k]
    }
    const int wm = warp >> 1, wn = warp & 1;
    const int gid = lane >> 2, tig = lane & 3;
    const int row0 = wm * 16;
#pragma unroll 1
    for (int b = 3; b >= 0; --b) {
        const int ncol = 32 * (3 - b);                // columns of L beyond block b
        for (int e = tid; e < 32 * (ncol / 2); e += 256) {
            const int r = e / (ncol / 2), ch = e - r * (ncol / 2);
            tu_cp_async16(&Lb[r * TU_LDL + ch * 2], LT + static_cast<long>(32 * b + r) * ldl + 32 * (b + 1) + ch * 2);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int col0 = b * 32 + wn * 16;
        double acc[2][2][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const double2 v = *reinterpret_cast<const double2*>(&Xs[(row0 + mi * 8 + gid) * TU_LDX + col0 + ni * 8 + 2 * tig]);
                acc[mi][ni][0] = v.x;
                acc[mi][ni][1] = v.y;
            }
        for (int k0 = 0; k0 < ncol; k0 += 4) {
            double af[2], bf[2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) af[mi] = -Xs[(row0 + mi * 8 + gid) * TU_LDX + 32 * (b + 1) + k0 + tig];
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) bf[ni] = Lb[(wn * 16 + ni * 8 + gid) * TU_LDL + k0 + tig];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) dmma884_u(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
                *reinterpret_cast<double2*>(&Xs[(row0 + mi * 8 + gid) * TU_LDX + col0 + ni * 8 + 2 * tig]) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        __syncthreads();
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 4) {
            double af[2], bf[2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) af[mi] = Xs[(row0 + mi * 8 + gid) * TU_LDX + b * 32 + k0 + tig];
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) bf[ni] = DsT[(b * 32 + wn * 16 + ni * 8 + gid) * TU_LDD + k0 + tig];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) dmma884_u(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
        __syncthreads();
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
                *reinterpret_cast<double2*>(&Xs[(row0 + mi * 8 + gid) * TU_LDX + col0 + ni * 8 + 2 * tig]) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        __syncthreads();
    }
    double* Pg = (P != nullptr) ? P + static_cast<long>(blockIdx.x) * TU_ROWS * EGX_NB : nullptr;
    for (int e = tid; e < TU_ROWS * 64; e += 256) {
        const int r = e >> 6, ch = e & 63;
        const double2 v = *reinterpret_cast<const double2*>(&Xs[r * TU_LDX + ch * 2]);
        *reinterpret_cast<double2*>(Xg + static_cast<long>(r) * ldx + ch * 2) = v;
        if (Pg != nullptr) *reinterpret_cast<double2*>(Pg + r * EGX_NB + ch * 2) = v;
    }
}

template <int CORR>
__device__ __forceinline__ double vg_finish(double acc, double prod) {
    if (CORR == EGX_CORR_SQUARED_EXPONENTIAL) return exp(-0.5 * acc);
    if (CORR == EGX_CORR_ABSOLUTE_EXPONENTIAL) return exp(-acc);
    if (CORR == EGX_CORR_MATERN32) return prod * exp(-1.7320508075688772 * acc);
    return prod * exp(-2.23606797749979 * acc);
}

// r and the logarithmic derivatives s_k = (dr/dx_k)/r of one (point, training point) pair
template <int CORR, int DMAX>
__device__ __forceinline__ double pair_r_s(const double* __restrict__ xi, const double* __restrict__ XjT, int jl,
                                           const CorrTerm* __restrict__ terms, int nterms, double (&s)[DMAX]) {
    const double sq = (CORR == EGX_CORR_MATERN32) ? 1.7320508075688772 : 2.23606797749979;
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
    return vg_finish<CORR>(acc, prod);
}

// One warp per point.  Pass 0: p2_k = sum_j w_j r_j s_jk.  Pass 1+l: A_l = f_l - sum_j r_j KF[l][j],
// T2[k][l] = sum_j r_j s_jk KF[l][j].  Then the p x p solves with G and the final combination.
template <int CORR, int DMAX>
__global__ void __launch_bounds__(256)
    var_grad_kernel(const double* __restrict__ W, long ldw, const double* __restrict__ xraw, int m,
                    const double* __restrict__ x_mean, const double* __restrict__ x_std, const double* __restrict__ X,
                    int n, int npad, int d, const CorrTerm* __restrict__ gterms, int nterms,
                    const double* __restrict__ KF, long ldk, const double* __restrict__ G, int p,
                    const int* __restrict__ basis_i, const int* __restrict__ basis_j, double sigma2,
                    double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* XjT = reinterpret_cast<double*>(smem_raw);        // [d][64]
    double* colv = XjT + EGX_CT * d;                           // [64] w_j or KF[l][j] of the current tile, per warp -> [8][64]
    double* xp = colv + 8 * EGX_CT;                            // [8][d]
    double* scr = xp + 8 * d;                                  // [8][d*p + 2p + d]  T2, A, dmat, p2
    CorrTerm* terms = reinterpret_cast<CorrTerm*>(scr + 8 * (static_cast<long>(d) * p + 2 * p + d));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i = blockIdx.x * 8 + warp;
    const bool active = i < m;
    for (int t = tid; t < nterms; t += 256) terms[t] = gterms[t];
    for (int e = tid; e < 8 * d; e += 256) {
        const int w = e / d, c = e - w * d;
        const int ii = blockIdx.x * 8 + w;
        xp[e] = (ii < m) ? (xraw[static_cast<long>(ii) * d + c] - x_mean[c]) / x_std[c] : 0.0;
    }
    const double* xi = xp + warp * d;
    double* T2 = scr + warp * (static_cast<long>(d) * p + 2 * p + d);   // [p][d]
    double* Av = T2 + static_cast<long>(d) * p;                         // [p]
    double* dm = Av + p;                                                // [p]
    double* p2 = dm + p;                                                // [d]
    double* cw = colv + warp * EGX_CT;

    for (int pass = 0; pass <= p; ++pass) {
        double g[DMAX];
#pragma unroll
        for (int k = 0; k < DMAX; ++k) g[k] = 0.0;
        double asum = 0.0;
        const double* colsrc = (pass == 0) ? (active ? W + static_cast<long>(i) * ldw : nullptr)
                                           : KF + static_cast<long>(pass - 1) * ldk;
        for (int j0 = 0; j0 < npad; j0 += EGX_CT) {
            __syncthreads();
            for (int e = tid; e < EGX_CT * d; e += 256) {
                const int r = e / d, c = e - r * d;
                XjT[c * EGX_CT + r] = X[static_cast<long>(j0 + r) * d + c];
            }
            // per-warp column values (w_j differs per point; KF is shared but staged per warp for simplicity)
            for (int e = lane; e < EGX_CT; e += 32) cw[e] = (colsrc != nullptr && j0 + e < n) ? colsrc[j0 + e] : 0.0;
            __syncthreads();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int jl = lane + 32 * h;
                double s[DMAX];
                const double r = pair_r_s<CORR, DMAX>(xi, XjT, jl, terms, nterms, s);
                const double rc = (j0 + jl < n) ? r * cw[jl] : 0.0;
                asum += rc;
#pragma unroll
                for (int k = 0; k < DMAX; ++k) g[k] += rc * s[k];
            }
        }
        asum = vg_warp_sum(asum);
#pragma unroll
        for (int k = 0; k < DMAX; ++k) g[k] = vg_warp_sum(g[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < DMAX; ++k)
                if (k < d) {
                    if (pass == 0) p2[k] = g[k];
                    else T2[static_cast<long>(pass - 1) * d + k] = g[k];
                }
            if (pass > 0) {
                const int l = pass - 1, bi = basis_i[l], bj = basis_j[l];
                const double f = (bi < 0 ? 1.0 : xi[bi]) * (bj < 0 ? 1.0 : xi[bj]);
                Av[l] = f - asum;
            }
        }
        __syncwarp();
    }
    if (!active) return;
    // dmat = (G^T G)^-1 A^T : forward with G^T (lower), backward with G (upper)
    for (int a = 0; a < p; ++a) {
        double s = 0.0;
        for (int b = lane; b < a; b += 32) s += G[b * p + a] * dm[b];
        s = vg_warp_sum(s);
        if (lane == 0) dm[a] = (Av[a] - s) / G[a * p + a];
        __syncwarp();
    }
    for (int a = p - 1; a >= 0; --a) {
        double s = 0.0;
        for (int b = a + 1 + lane; b < p; b += 32) s += G[a * p + b] * dm[b];
        s = vg_warp_sum(s);
        if (lane == 0) dm[a] = (dm[a] - s) / G[a * p + a];
        __syncwarp();
    }
    for (int k = lane; k < d; k += 32) {
        double p4 = 0.0;
        for (int l = 0; l < p; ++l) {
            const int bi = basis_i[l], bj = basis_j[l];
            double df = 0.0;
            if (bi == k) df += (bj < 0) ? 1.0 : xi[bj];
            if (bj == k) df += (bi < 0) ? 1.0 : xi[bi];
            p4 += dm[l] * (df - T2[static_cast<long>(l) * d + k]);
        }
        out[static_cast<long>(i) * d + k] = 2.0 * (p4 - p2[k]) / x_std[k] * sigma2;
    }
}

template <typename K>
void vg_set_smem(K kernel, size_t bytes) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
}

}  // namespace

void launch_transpose_lower(const double* L, long ld, double* LT, int T, cudaStream_t s) {
    const int pairs = T * (T + 1) / 2;
    transpose_lower_kernel<<<pairs * 16, 256, 0, s>>>(L, ld, LT, T);
}

void launch_trsm_rows_upper(double* X, long ldx, const double* LTkk, long ldl, const double* Dinv, double* P,
                            int nblocks64, cudaStream_t s) {
    const int smem = (TU_ROWS * TU_LDX + 4 * 32 * TU_LDD + 32 * TU_LDL) * sizeof(double);
    vg_set_smem(trsm_rows_upper_kernel, smem);
    if (nblocks64 <= 0) return;
    trsm_rows_upper_kernel<<<nblocks64, 256, smem, s>>>(X, ldx, LTkk, ldl, Dinv, P);
}

size_t var_grad_smem_bytes(int d, int p, int nterms) {
    return (static_cast<size_t>(EGX_CT) * d + 8 * EGX_CT + 8 * d + 8 * (static_cast<size_t>(d) * p + 2 * p + d)) *
               sizeof(double) +
           nterms * sizeof(CorrTerm);
}

template <int CORR>
static void launch_vg(int d, int grid, size_t smem, cudaStream_t s, const double* W, long ldw, const double* xraw, int m,
                      const double* x_mean, const double* x_std, const double* X, int n, int npad,
                      const CorrTerm* terms, int nterms, const double* KF, long ldk, const double* G, int p,
                      const int* bi, const int* bj, double sigma2, double* out) {
    if (d <= 8) {
        vg_set_smem(var_grad_kernel<CORR, 8>, smem);
        var_grad_kernel<CORR, 8><<<grid, 256, smem, s>>>(W, ldw, xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms, KF,
                                                         ldk, G, p, bi, bj, sigma2, out);
    } else if (d <= 16) {
        vg_set_smem(var_grad_kernel<CORR, 16>, smem);
        var_grad_kernel<CORR, 16><<<grid, 256, smem, s>>>(W, ldw, xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms, KF,
                                                          ldk, G, p, bi, bj, sigma2, out);
    } else {
        vg_set_smem(var_grad_kernel<CORR, 32>, smem);
        var_grad_kernel<CORR, 32><<<grid, 256, smem, s>>>(W, ldw, xraw, m, x_mean, x_std, X, n, npad, d, terms, nterms, KF,
                                                          ldk, G, p, bi, bj, sigma2, out);
    }
}

void launch_var_grad(int corr, const double* W, long ldw, const double* xraw, int m, const double* x_mean,
                     const double* x_std, const double* X, int n, int npad, int d, const CorrTerm* terms, int nterms,
                     const double* KF, long ldk, const double* G, int p, const int* basis_i, const int* basis_j,
                     double sigma2, double* out, cudaStream_t s) {
    const int grid = (m + 7) / 8;
    const size_t smem = var_grad_smem_bytes(d, p, nterms);
    switch (corr) {
        case EGX_CORR_SQUARED_EXPONENTIAL:
            launch_vg<EGX_CORR_SQUARED_EXPONENTIAL>(d, grid, smem, s, W, ldw, xraw, m, x_mean, x_std, X, n, npad, terms,
                                                    nterms, KF, ldk, G, p, basis_i, basis_j, sigma2, out);
            break;
        case EGX_CORR_ABSOLUTE_EXPONENTIAL:
            launch_vg<EGX_CORR_ABSOLUTE_EXPONENTIAL>(d, grid, smem, s, W, ldw, xraw, m, x_mean, x_std, X, n, npad, terms,
                                                     nterms, KF, ldk, G, p, basis_i, basis_j, sigma2, out);
            break;
        case EGX_CORR_MATERN32:
            launch_vg<EGX_CORR_MATERN32>(d, grid, smem, s, W, ldw, xraw, m, x_mean, x_std, X, n, npad, terms, nterms, KF,
                                         ldk, G, p, basis_i, basis_j, sigma2, out);
            break;
        default:
            launch_vg<EGX_CORR_MATERN52>(d, grid, smem, s, W, ldw, xraw, m, x_mean, x_std, X, n, npad, terms, nterms, KF,
                                         ldk, G, p, basis_i, basis_j, sigma2, out);
            break;
    }
}
