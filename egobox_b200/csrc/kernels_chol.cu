// Right-looking blocked Cholesky building blocks (fp64), block size 128:
//   K3  potrf_diag   : 128 x 128 diagonal block factorised in shared memory
//   K5  trsm_rows    : X <- X * L_kk^-T for 64-row slabs (true triangular solve, no
//                      explicit inverse), also emits the contiguous panel copy P
//   K4  gemm_nt_sub  : C -= A * B^T on 128 x 128 tiles with FP64 tensor-core MMA
//                      (mma.sync.m8n8k4.f64 -> SASS DMMA), used for the trailing SYRK
//                      update and for the multi-RHS TRSM of predict_var.
//
// Reference being replaced: `r_mx.cholesky()` gp/src/algorithm.rs:1004 (linfa-linalg)
// / :1077 (LAPACK dpotrf Lower) and `solve_triangular` :343-350, 1006, 1028.
//
// tcgen05 has no f64 kind (kinds: f16, tf32, f8f6f4, i8, mxf8f6f4, mxf4, mxf4nvf4), so the
// fp64-exact contraction runs on the DMMA pipe; see DESIGN.md for the int8-sliced tcgen05
// variant and which one each profile measures.
#include <cstdlib>

#include "common.cuh"
#include "../../include/egobox_gpu.h"

namespace {

// ---------------------------------------------------------------------------
// K3: diagonal block, register-resident right-looking Cholesky.
// 256 threads = 16 x 16 grid; thread (ty, tx) owns S[ty + 16 a][tx + 16 b], a, b = 0..7
// (2-D cyclic, so the shrinking trailing matrix stays balanced).  Per column j the owners of
// column j publish it through a double-buffered 128-entry shared vector (ONE barrier per
// column), every thread scales with rsqrt(pivot) and applies the rank-1 update to its
// registers.  Afterwards the four 32 x 32 diagonal sub-blocks are inverted (one warp each,
// one lane per column) for the blocked triangular solves of K5.
// ---------------------------------------------------------------------------
constexpr int PD_DLD = 33;   // smem leading dimension of the 32x32 diagonal sub-blocks

// The column owners publish column j with ZEROS in rows <= j (and the pivot in slot 128), so that
// every thread can apply the rank-1 update unconditionally: rows / columns that are already final
// see a zero multiplier.  This keeps the per-column instruction count (the kernel is issue bound,
// not latency bound: ~8 warps x 128 columns on ONE SM) close to the 36 DFMA it needs.
template <int JA>
__device__ __forceinline__ void potrf_block_columns(double (&a)[8][8], double (*colbuf)[EGX_NB + 8], int ty, int tx,
                                                    int* info, int base_index, bool& failed) {
#pragma unroll 1
    for (int jr = 0; jr < 16; ++jr) {
        const int j = JA * 16 + jr;
        double* cb = colbuf[j & 1];
        if (tx == jr) {
#pragma unroll
            for (int ai = JA; ai < 8; ++ai) {
                const int i = ty + 16 * ai;
                cb[i] = (i > j) ? a[ai][JA] : 0.0;
                if (i == j) cb[EGX_NB] = a[ai][JA];
            }
        }
        __syncthreads();
        const double dj = cb[EGX_NB];
        if (!(dj > 0.0)) {                       // LAPACK dpotrf: non-positive or NaN pivot
            if (ty == 0 && tx == 0 && !failed) atomicCAS(info, 0, base_index + j + 1);
            failed = true;
        }
        const double rs = rsqrt(dj);
        double lr[8], lc[8];
#pragma unroll
        for (int ai = JA; ai < 8; ++ai) lr[ai] = cb[ty + 16 * ai] * rs;
#pragma unroll
        for (int bi = JA; bi < 8; ++bi) lc[bi] = cb[tx + 16 * bi] * rs;
        // rank-1 update of the trailing block (upper halves of diagonal 16x16 register blocks are
        // updated too: they are never read)
#pragma unroll
        for (int ai = JA; ai < 8; ++ai)
#pragma unroll
            for (int bi = JA; bi <= ai; ++bi) a[ai][bi] -= lr[ai] * lc[bi];
        // final values of column j (owners only): L[i][j] = S[i][j] / sqrt(d), L[j][j] = sqrt(d)
        if (tx == jr) {
#pragma unroll
            for (int ai = JA; ai < 8; ++ai) {
                const int i = ty + 16 * ai;
                if (i > j) a[ai][JA] = lr[ai];
                else if (i == j) a[ai][JA] = dj * rs;
            }
        }
    }
}

__global__ void __launch_bounds__(256) potrf_diag_kernel(double* __restrict__ A, long ld, int* __restrict__ info,
                                                         int base_index, double* __restrict__ Dinv) {
    __shared__ double colbuf[2][EGX_NB + 8];
    __shared__ double Dg[4][32 * PD_DLD];
    __shared__ double rdiag[4][32];
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    double a[8][8];
#pragma unroll
    for (int ai = 0; ai < 8; ++ai)
#pragma unroll
        for (int bi = 0; bi < 8; ++bi) {
            const int r = ty + 16 * ai, c = tx + 16 * bi;
            a[ai][bi] = (c <= r) ? A[static_cast<long>(r) * ld + c] : 0.0;
        }
    bool failed = false;
    potrf_block_columns<0>(a, colbuf, ty, tx, info, base_index, failed);
    potrf_block_columns<1>(a, colbuf, ty, tx, info, base_index, failed);
    potrf_block_columns<2>(a, colbuf, ty, tx, info, base_index, failed);
    potrf_block_columns<3>(a, colbuf, ty, tx, info, base_index, failed);
    potrf_block_columns<4>(a, colbuf, ty, tx, info, base_index, failed);
    potrf_block_columns<5>(a, colbuf, ty, tx, info, base_index, failed);
    potrf_block_columns<6>(a, colbuf, ty, tx, info, base_index, failed);
    potrf_block_columns<7>(a, colbuf, ty, tx, info, base_index, failed);

#pragma unroll
    for (int ai = 0; ai < 8; ++ai)
#pragma unroll
        for (int bi = 0; bi < 8; ++bi) {
            const int r = ty + 16 * ai, c = tx + 16 * bi;
            if (c <= r) {
                A[static_cast<long>(r) * ld + c] = a[ai][bi];
                if ((r >> 5) == (c >> 5)) Dg[r >> 5][(r & 31) * PD_DLD + (c & 31)] = a[ai][bi];
            }
        }
    __syncthreads();
    // inverse of each 32 x 32 diagonal sub-block: lane c solves L x = e_c by forward substitution
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < 4) {
        const double* Lb = Dg[warp];
        rdiag[warp][lane] = 1.0 / Lb[lane * PD_DLD + lane];
        __syncwarp();
        double x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            double sacc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int j = 0; j < i; ++j) sacc -= Lb[i * PD_DLD + j] * x[j];
            x[i] = sacc * rdiag[warp][i];
        }
        double* Do = Dinv + warp * 1024;
#pragma unroll
        for (int i = 0; i < 32; ++i) Do[i * 32 + lane] = x[i];
    }
}

__device__ __forceinline__ void dmma884_t(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void tr_cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// ---------------------------------------------------------------------------
// K3 (r02 form): the same 128 x 128 factorisation with the latency chain taken out of the CTA-wide barriers.
// The tile sits in shared memory; it is factorised in four 32-column block steps:
//   column phase (no CTA barrier): warp w < 3 - b holds, one row per lane, the 32 x 32 diagonal block (every warp
//     its own copy, symmetric storage) AND the 32 rows of panel block w below it.  Per column: the pivot is
//     shuffled from its lane, every lane scales its entries with rsqrt(pivot), the scaled column goes through a
//     warp-private 32-entry shared vector and the rank-1 update of diagonal block and panel rows runs out of
//     registers.  The dependent chain of a column is shuffle + rsqrt + DMUL + DFMA (~120 clk against ~650 clk of the
//     barrier form), the panel rows ride in its issue bubbles, so the in-tile panel solve needs no inverse.
//   update phase: the trailing (96 - 32 b)^2 lower triangle -= X X^T (K = 32) on the FP64 tensor pipe, 16 x 16
//     macro tiles over the 8 warps.
// Warp 4 inverts the previous 32 x 32 diagonal block (for K5) while the other warps are in the column phase.
// ---------------------------------------------------------------------------
constexpr int P2_LD = 132;     // (4 g + t) mod 16 distinct: conflict-free DMMA fragment loads
#ifdef POTRF_TIMING             // tools/micro/potrf_probe.cu: clock stamps of the phases of one launch
__device__ long long potrf_dbg[32];
#define POTRF_STAMP(i) do { if ((threadIdx.x & 31) == 0) potrf_dbg[i] = clock64(); } while (0)
#else
#define POTRF_STAMP(i) do { } while (0)
#endif

// y ~ d^-1/2 to fp64 accuracy: MUFU.RSQ64H seed (2^-22) + one third-order step (e^3 term ~ 2^-67), 4 dependent
// fp64 operations instead of the ~7 of rsqrt()
__device__ __forceinline__ double potrf_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);              // 1 - d y^2
    const double pc = fma(0.375, e, 0.5);          // 1/2 + 3/8 e
    const double ye = y * e;
    return fma(ye, pc, y);                         // y (1 + e/2 + 3 e^2/8)
}

template <bool PANEL>
__device__ __forceinline__ void potrf32_columns(double (&a)[32], double (&p)[32], double dg, double* lb, int lane,
                                                int& fail_col) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const double d = __shfl_sync(0xffffffffu, dg, j);
        if (!(d > 0.0) && fail_col < 0) fail_col = j;       // uniform over the warp
#ifdef POTRF_LIBRSQRT
        const double rs = rsqrt(d);
#else
        const double rs = potrf_rsqrt(d);
#endif
        const double l = a[j] * rs;                          // lane j: d * rs = sqrt(d)
        a[j] = l;
        double lp = 0.0;
        if (PANEL) {
            lp = p[j] * rs;
            p[j] = lp;
        }
        dg = fma(-l, l, dg);                                 // the lane's own diagonal entry: no round trip on the chain
        if (j < 31) {
            double* buf = lb + (j & 1) * 32;
            buf[lane] = l;
            __syncwarp();
            if (((j + 1) & 1) != 0) {
                const double lk = buf[j + 1];
                a[j + 1] = fma(-l, lk, a[j + 1]);
                if (PANEL) p[j + 1] = fma(-lp, lk, p[j + 1]);
            }
#pragma unroll
            for (int k = (j + 2) & ~1; k < 32; k += 2) {
                const double2 lk = *reinterpret_cast<const double2*>(buf + k);
                a[k] = fma(-l, lk.x, a[k]);
                a[k + 1] = fma(-l, lk.y, a[k + 1]);
                if (PANEL) {
                    p[k] = fma(-lp, lk.x, p[k]);
                    p[k + 1] = fma(-lp, lk.y, p[k + 1]);
                }
            }
        }
    }
}

// inverse of the 32 x 32 lower block at Sb (leading dimension P2_LD): lane c solves L x = e_c by forward substitution
// (componentwise accurate; a blocked form X21 = -C^-1 B A^-1 on four warps was 2x faster and lost the ill-conditioned
// band test -- cond(R) ~ 1e15 -- to its larger error, profiles/r02/y3_potrf_probe.txt).  Row i of L is read as
// 16-byte broadcast loads; two partial sums, the term of x[i-1] -- the only one on the dependent chain of the rows -- last.
__device__ __forceinline__ void potrf_invert32(const double* Sb, double* __restrict__ Do, int lane) {
    const double myrd = 1.0 / Sb[lane * P2_LD + lane];
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
        for (int j = 0; j + 1 <= i - 2; j += 2) {       // pairs (j, j+1) below the last term
            const double2 lv = *reinterpret_cast<const double2*>(&Sb[i * P2_LD + j]);
            s0 = fma(-lv.x, x[j], s0);
            s1 = fma(-lv.y, x[j + 1], s1);
        }
        if (i >= 2 && ((i - 1) & 1)) s0 = fma(-Sb[i * P2_LD + i - 2], x[i - 2], s0);   // odd count of terms below i-1: the last single
        double sacc = s0 + s1;
        if (i >= 1) sacc = fma(-Sb[i * P2_LD + i - 1], x[i - 1], sacc);
        x[i] = sacc * __shfl_sync(0xffffffffu, myrd, i);
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) Do[i * 32 + lane] = x[i];
}

__global__ void __launch_bounds__(256, 1) potrf_diag2_kernel(double* __restrict__ A, long ld, int* __restrict__ info,
                                                             int base_index, double* __restrict__ Dinv) {
    extern __shared__ __align__(16) double psm[];
    double* S = psm;                                   // [128][P2_LD]
    double* lbuf = psm + EGX_NB * P2_LD;               // [4 warps][2][32]
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    // lower triangle of the tile -> shared memory (16-byte chunks that touch columns <= row)
    for (int e = tid; e < EGX_NB * 64; e += 256) {
        const int r = e >> 6, ch = e & 63;
        if (ch * 2 <= r) tr_cp_async16(&S[r * P2_LD + ch * 2], A + static_cast<long>(r) * ld + ch * 2);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (tid == 0) POTRF_STAMP(0);
    __syncthreads();
    if (tid == 0) POTRF_STAMP(1);

    int fail_col = -1;
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        const int c0 = 32 * b;
        const int npanel = 3 - b;
        if (warp < npanel || warp == 0) {
            double a[32], p[32];
            // diagonal block, symmetric: entry (lane, k) from the stored lower triangle
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const int hi = lane > k ? lane : k, lo = lane > k ? k : lane;
                a[k] = S[(c0 + hi) * P2_LD + c0 + lo];
            }
            const double dg = S[(c0 + lane) * P2_LD + c0 + lane];
            double* lb = lbuf + warp * 64;
            // every warp of the phase has its copy of the diagonal block before warp 0 may write the factor over it (it does so
            // ~6 k clk later; compute-sanitizer racecheck rightly wants the order stated): named barrier of the 32 * npanel threads
            if (npanel >= 2) asm volatile("bar.sync 1, %0;" ::"r"(32 * npanel) : "memory");
            if (tid == 0) POTRF_STAMP(2 + 4 * b);     // registers loaded (diagonal block)
#ifdef POTRF_ONE_INST
            if (true) {
                const int prow = warp < npanel ? c0 + 32 * (warp + 1) + lane : c0 + lane;   // b = 3: a second copy of the diagonal rows
                const double* pr = &S[prow * P2_LD + c0];
#else
            if (warp < npanel) {
                const int prow = c0 + 32 * (warp + 1) + lane;
                const double* pr = &S[prow * P2_LD + c0];
#endif
#pragma unroll
                for (int k = 0; k < 32; k += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(pr + k);
                    p[k] = v.x;
                    p[k + 1] = v.y;
                }
                int fc = -1;
                potrf32_columns<true>(a, p, dg, lb, lane, fc);
                if (fail_col < 0 && fc >= 0) fail_col = c0 + fc;
                if (tid == 0) POTRF_STAMP(3 + 4 * b);  // columns done
                if (warp < npanel) {
                    double* pw = &S[prow * P2_LD + c0];
#pragma unroll
                    for (int k = 0; k < 32; k += 2) *reinterpret_cast<double2*>(pw + k) = make_double2(p[k], p[k + 1]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) p[k] = 0.0;
                int fc = -1;
                potrf32_columns<false>(a, p, dg, lb, lane, fc);
                if (fail_col < 0 && fc >= 0) fail_col = c0 + fc;
                if (tid == 0) POTRF_STAMP(3 + 4 * b);
            }
            if (warp == 0) {
                // entries right of the diagonal are scratch: nothing reads the upper triangle of S
                double* dw = &S[(c0 + lane) * P2_LD + c0];
#pragma unroll
                for (int k = 0; k < 32; k += 2) *reinterpret_cast<double2*>(dw + k) = make_double2(a[k], a[k + 1]);
            }
        } else if (warp == 4 && b > 0) {
            potrf_invert32(&S[(c0 - 32) * P2_LD + c0 - 32], Dinv + (b - 1) * 1024, lane);
        }
        __syncthreads();
        if (tid == 0) POTRF_STAMP(4 + 4 * b);          // column phase over for the CTA (incl. the inverse of the previous block)
        if (b == 3) break;
        // trailing update: rows / columns c0 + 32 .. 127, lower triangle in 16 x 16 macro tiles
        const int t0 = c0 + 32;
        const int nt = (EGX_NB - t0) / 16;
        const int ntile = nt * (nt + 1) / 2;
        const int gid = lane >> 2, tig = lane & 3;
        for (int t = warp; t < ntile; t += 8) {
            int tr = 0;
            while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
            const int tc = t - tr * (tr + 1) / 2;
            const int r0 = t0 + 16 * tr, q0 = t0 + 16 * tc;
            double acc[2][2][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) {
                    const double2 v = *reinterpret_cast<const double2*>(&S[(r0 + mi * 8 + gid) * P2_LD + q0 + ni * 8 + 2 * tig]);
                    acc[mi][ni][0] = v.x;
                    acc[mi][ni][1] = v.y;
                }
#pragma unroll
            for (int k0 = 0; k0 < 32; k0 += 4) {
                double af[2], bf[2];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) af[mi] = -S[(r0 + mi * 8 + gid) * P2_LD + c0 + k0 + tig];
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) bf[ni] = S[(q0 + ni * 8 + gid) * P2_LD + c0 + k0 + tig];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) dmma884_t(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni)
                    *reinterpret_cast<double2*>(&S[(r0 + mi * 8 + gid) * P2_LD + q0 + ni * 8 + 2 * tig]) =
                        make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
        __syncthreads();
        if (tid == 0) POTRF_STAMP(5 + 4 * b);          // trailing update done
    }
    // LAPACK dpotrf: 1-based index of the first non-positive (or NaN) pivot; warp 0 saw every pivot
    if (tid == 0 && fail_col >= 0) atomicCAS(info, 0, base_index + fail_col + 1);
    if (warp == 4) {
        potrf_invert32(&S[96 * P2_LD + 96], Dinv + 3 * 1024, lane);
        POTRF_STAMP(20);
        return;
    }
    // lower triangle back to the matrix: 7 warps, 8 chunks of 16 bytes in flight per thread
    {
        const int wt = warp < 4 ? tid : tid - 32;       // 0 .. 223
        for (int e0 = 0; e0 < EGX_NB * 64; e0 += 224 * 8) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * 224 + wt;
                const int r = e >> 6, c = (e & 63) * 2;
                if (e < EGX_NB * 64 && c <= r) v[u] = *reinterpret_cast<const double2*>(&S[r * P2_LD + c]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * 224 + wt;
                const int r = e >> 6, c = (e & 63) * 2;
                if (e < EGX_NB * 64) {
                    if (c < r) *reinterpret_cast<double2*>(A + static_cast<long>(r) * ld + c) = v[u];
                    else if (c == r) A[static_cast<long>(r) * ld + c] = v[u].x;
                }
            }
        }
    }
    if (tid == 0) POTRF_STAMP(21);
}

// ---------------------------------------------------------------------------
// K5: X (64 x 128 slab, in place) <- X * L^-T with L = 128 x 128 lower block, blocked by
// 32 columns on the FP64 tensor pipe:
//   for b = 0..3:  T   = X_b - sum_{b' < b} X_b' * L[b, b']^T      (DMMA, K = 32 b)
//                  X_b = T * Dinv_b^T                               (DMMA, K = 32)
// Dinv_b are the inverted 32 x 32 diagonal sub-blocks produced by K3 (the standard blocked
// TRSM of GPU BLAS libraries).  Also emits the contiguous panel copy P used by K4.
// 256 threads = 8 warps as 4 (rows) x 2 (cols), warp tile 16 x 16 of the 64 x 32 output.
// ---------------------------------------------------------------------------
constexpr int TR_LDX = 132;   // (4 g + t) mod 16 distinct -> conflict-free fragment loads
constexpr int TR_LDL = 100;   // 96 columns of one 32-row block of L (+4 pad)
constexpr int TR_LDD = 36;

// Shared memory: X slab ROWS x 132 + the four Dinv blocks (36.9 KB) + the 96 rows of L below its first diagonal
// sub-block (96 x 100, 76.8 KB) -- everything is staged ONCE at the start (r01 re-staged one 32-row block of L per
// block step to stay under 130 KB beside a 92 KB update CTA; the r02 update kernel fills an SM on its own, and the three
// exposed global round trips cost ~3 us of a kernel on the serial chain).
// ROWS = 64: 8 warps as 4 x 2, warp tile 16 x 16.  ROWS = 32 (chosen when the 32-row slabs still fit in one wave):
// 8 warps as 2 x 4, warp tile 16 x 8 -- half the FP64 work per CTA on the chain.
template <int ROWS>
__global__ void __launch_bounds__(256) trsm_rows_kernel(double* __restrict__ X, long ldx,
                                                        const double* __restrict__ L, long ldl,
                                                        const double* __restrict__ Dinv,
                                                        double* __restrict__ P, long ldp,
                                                        double* __restrict__ rmaxq) {
    constexpr int WN = (ROWS == 64) ? 2 : 4;      // warps along the 32 output columns of a block step
    constexpr int NI = 4 / WN;                    // n8 tiles per warp
    constexpr int NW = 8 * NI;                    // output columns per warp
    extern __shared__ __align__(16) double sm[];
    double* Xs = sm;                              // [ROWS][132]
    double* Ds = Xs + ROWS * TR_LDX;              // [4][32][36]
    double* Lb = Ds + 4 * 32 * TR_LDD;            // [96][100]: rows 32 .. 127 of L, columns 0 .. 32 (row / 32) - 1
    double* Ts = Lb + 96 * TR_LDL;                // [ROWS][36]: T of the current block step (its own buffer: one barrier less per step)
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    double* Xg = X + static_cast<long>(blockIdx.x) * ROWS * ldx;

    for (int e = tid; e < ROWS * 64; e += 256) {
        const int r = e >> 6, ch = e & 63;
        tr_cp_async16(&Xs[r * TR_LDX + ch * 2], Xg + static_cast<long>(r) * ldx + ch * 2);
    }
    for (int e = tid; e < 2048; e += 256) {
        const int b = e >> 9, r = (e >> 4) & 31, ch = e & 15;
        tr_cp_async16(&Ds[(b * 32 + r) * TR_LDD + ch * 2], Dinv + e * 2);
    }
    for (int e = tid; e < 96 * 48; e += 256) {
        const int r = e / 48, ch = e - r * 48;     // row 32 + r of L, 16-byte chunk ch
        if (ch < 16 * (1 + (r >> 5))) tr_cp_async16(&Lb[r * TR_LDL + ch * 2], L + static_cast<long>(32 + r) * ldl + ch * 2);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    const int wm = warp / WN, wn = warp % WN;
    const int gid = lane >> 2, tig = lane & 3;
    const int row0 = wm * 16;        // rows of this warp inside the slab
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
        const int col0 = b * 32 + wn * NW;          // output columns of this warp
        double acc[2][NI][2];
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const double2 v = *reinterpret_cast<const double2*>(&Xs[(row0 + mi * 8 + gid) * TR_LDX + col0 + ni * 8 + 2 * tig]);
                acc[mi][ni][0] = v.x;
                acc[mi][ni][1] = v.y;
            }
        // T = X_b - X[:, 0:32b] * L[b-block rows, 0:32b]^T   (8 b k-steps: unrolled by 8 so that the fragment loads run ahead)
        const double* Lrow = Lb + ((b - 1) * 32 + wn * NW + gid) * TR_LDL + tig;
        const double* Xrow = Xs + (row0 + gid) * TR_LDX + tig;
#pragma unroll 1
        for (int k1 = 0; k1 < 32 * b; k1 += 32) {
#pragma unroll
            for (int k0 = k1; k0 < k1 + 32; k0 += 4) {
                double af[2], bf[NI];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) af[mi] = -Xrow[mi * 8 * TR_LDX + k0];
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) bf[ni] = Lrow[ni * 8 * TR_LDL + k0];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < NI; ++ni) dmma884_t(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
            }
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni)
                *reinterpret_cast<double2*>(&Ts[(row0 + mi * 8 + gid) * TR_LDD + wn * NW + ni * 8 + 2 * tig]) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        __syncthreads();        // T is visible
        // X_b = T * Dinv_b^T
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll
        for (int k0 = 0; k0 < 32; k0 += 4) {
            double af[2], bf[NI];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) af[mi] = Ts[(row0 + mi * 8 + gid) * TR_LDD + k0 + tig];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) bf[ni] = Ds[(b * 32 + wn * NW + ni * 8 + gid) * TR_LDD + k0 + tig];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) dmma884_t(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni)
                *reinterpret_cast<double2*>(&Xs[(row0 + mi * 8 + gid) * TR_LDX + col0 + ni * 8 + 2 * tig]) =
                    make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        __syncthreads();        // X_b visible to the next block step; every warp is done with T
    }
    // panel copy for the trailing update: 128 columns of a (rows x ldp) buffer (ldp = 256: two panels side by side)
    // rmaxq (optional): max |x| of every 64-column quarter of the solved rows, [row][2] at the caller's offset -- the
    // row scales of the int8 slicing (kernels_ozaki.cu) then need no extra pass over the panel
    double* Pg = (P != nullptr) ? P + static_cast<long>(blockIdx.x) * ROWS * ldp : nullptr;
    for (int e = tid; e < ROWS * 64; e += 256) {
        const int r = e >> 6, ch = e & 63;        // a warp covers 64 consecutive columns of one row
        const double2 v = *reinterpret_cast<const double2*>(&Xs[r * TR_LDX + ch * 2]);
        *reinterpret_cast<double2*>(Xg + static_cast<long>(r) * ldx + ch * 2) = v;
        if (Pg != nullptr) *reinterpret_cast<double2*>(Pg + static_cast<long>(r) * ldp + ch * 2) = v;
        if (rmaxq != nullptr) {
            double mx = fmax(fabs(v.x), fabs(v.y));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            if (lane == 0) rmaxq[(static_cast<long>(blockIdx.x) * ROWS + r) * 4 + (ch >> 5)] = mx;
        }
    }
}

// ---------------------------------------------------------------------------
// K4: C(128x128 tile) -= A(128 x K) * B(128 x K)^T,  K = 128, all row-major fp64.
// 256 threads = 8 warps (2 x 4), warp tile 64 x 32, mma.m8n8k4.f64, 3-stage
// cp.async pipeline over K in steps of 16.  The accumulators are initialised
// with the C tile and A fragments are negated, so the epilogue is a pure store.
// ---------------------------------------------------------------------------

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// tile t -> (row tile r of 128 rows, column tile c of BN columns)
template <int BN>
__device__ __forceinline__ void gemm_tile_decode(const GemmArgs& g, int t, int& r, int& c) {
    constexpr int S = EGX_NB / BN;          // column tiles per 128-block (1 or 2)
    if (g.tri > 0) {
        const int ntri = S * g.tri * (g.tri + 1) / 2;
        if (t < ntri) {
            // rows r = 0..tri-1 hold S*(r+1) tiles each; prefix(r) = S r (r+1) / 2
            int rr = static_cast<int>((sqrt(8.0 * static_cast<double>(t) / S + 1.0) - 1.0) * 0.5);
            while (S * (rr + 1) * (rr + 2) / 2 <= t) ++rr;
            while (S * rr * (rr + 1) / 2 > t) --rr;
            r = rr;
            c = t - S * rr * (rr + 1) / 2;
        } else {
            const int u = t - ntri, w = S * g.tri;
            r = g.tri + u / w;
            c = u % w;
        }
    } else {
        const int w = S * g.Nt;
        r = t / w;
        c = t % w;
    }
}

// ADD = false: C -= A B^T (the negation folds into the DMMA operand, SASS `DMMA R, -R, R, R`); ADD = true: C += A B^T
// BK x STAGES: 16 x 3 (default) or 32 x 2 (half the barriers per tile, same shared-memory footprint class).
// mbarrier helpers of the MB variant (no CTA-wide barrier in the main loop)
__device__ __forceinline__ void mb_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
                 "r"(count));
}
__device__ __forceinline__ void mb_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MB_DONE_%=;\n"
        "bra MB_WAIT_%=;\n"
        "MB_DONE_%=:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mb_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
                 : "memory");
}
// the calling thread's earlier cp.async copies arrive on the barrier when they have landed
__device__ __forceinline__ void mb_cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(bar)))
                 : "memory");
}

// MB = true: the per-k-block __syncthreads is replaced by two mbarrier arrays -- full[s] (256 arrivals, one per
// thread, triggered by the completion of its cp.async copies into stage s) and empty[s] (8 arrivals, one per warp,
// after its last fragment load from stage s).  A warp starts a k-block as soon as the data has landed, whatever the
// progress of the other warps; only the refill of a stage waits for every warp to have left it, and that wait sits
// behind 16 queued MMAs.
template <int BN, bool ADD, int BK, int STAGES, bool MB = false>
__global__ void __launch_bounds__(256, (BN == 128) ? 1 : 2) gemm_nt_sub_kernel(const GemmArgs g) {
    constexpr int WARPS_N = (BN == 128) ? 4 : 2;
    constexpr int WARPS_M = 8 / WARPS_N;
    constexpr int MI = EGX_NB / WARPS_M / 8;      // m8 tiles per warp
    constexpr int NI = BN / WARPS_N / 8;          // n8 tiles per warp
    constexpr int LDS = BK + 4;                   // (4 g + t) mod 16 distinct: conflict-free fragment loads
    constexpr int A_ELEMS = EGX_NB * LDS, B_ELEMS = BN * LDS;
    constexpr int CH = BK / 2;                    // 16-byte chunks per row
    extern __shared__ __align__(16) double gsm[];
    double* As = gsm;                               // [STAGES][128][LDS]
    double* Bs = gsm + STAGES * A_ELEMS;            // [STAGES][BN][LDS]
    __shared__ uint64_t mb_full[STAGES], mb_empty[STAGES];
    if constexpr (MB) {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
                mb_init(&mb_full[s], 256);
                mb_init(&mb_empty[s], 8);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    int tr, tc;
    gemm_tile_decode<BN>(g, blockIdx.x, tr, tc);
    const long koff = static_cast<long>(blockIdx.y) * g.K;      // split-K slice
    const double* Ag = g.A + static_cast<long>(tr) * EGX_NB * g.lda + koff;
    const double* Bg = g.B + static_cast<long>(tc) * BN * g.ldb + koff;
    double* Cg = g.C + static_cast<long>(blockIdx.y) * g.split_c_stride + static_cast<long>(tr) * EGX_NB * g.ldc +
                 static_cast<long>(tc) * BN;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int gid = lane >> 2, tig = lane & 3;

    // per-thread copy slots: chunk `ch` (16 bytes) of rows row0, row0 + 256/CH, ... ; the global pointers are formed
    // once, a k-block step only adds kb * BK to them
    constexpr int RSTEP = 256 / CH;                 // rows covered by one pass of the 256 threads
    const int row0 = tid / CH, ch0 = tid % CH;
    const double* a_src = Ag + static_cast<long>(row0) * g.lda + ch0 * 2;
    const double* b_src = Bg + static_cast<long>(row0) * g.ldb + ch0 * 2;
    const long a_step = static_cast<long>(RSTEP) * g.lda, b_step = static_cast<long>(RSTEP) * g.ldb;
    const int s_off = row0 * LDS + ch0 * 2;
    auto load_stage = [&](int stage, int kb) {
        double* as = As + stage * A_ELEMS + s_off;
        double* bs = Bs + stage * B_ELEMS + s_off;
        const double* ap = a_src + kb * BK;
        const double* bp = b_src + kb * BK;
#pragma unroll
        for (int i = 0; i < EGX_NB / RSTEP; ++i) cp_async16(as + i * RSTEP * LDS, ap + i * a_step);
#pragma unroll
        for (int i = 0; i < BN / RSTEP; ++i) cp_async16(bs + i * RSTEP * LDS, bp + i * b_step);
    };

    const int KB = g.K / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KB) {
            load_stage(s, s);
            if constexpr (MB) mb_cp_async_arrive(&mb_full[s]);
        }
        cp_async_commit();
    }

    // late_c: accumulate the product from zero and fold the C tile in at the end, so that the first DMMA does not
    // wait for the C tile from HBM (it is prefetched into L2 here and read back as L2 hits in the epilogue);
    // otherwise the accumulators start as the C tile
    double acc[MI][NI][2];
    if (g.late_c) {
        // 128 rows x BN doubles = BN/16 128-byte lines per row
        for (int i = tid; i < EGX_NB * (BN / 16); i += 256) {
            const int row = i / (BN / 16), ln = i % (BN / 16);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(Cg + static_cast<long>(row) * g.ldc + ln * 16));
        }
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    } else {
#pragma unroll
        for (int mi = 0; mi < MI; ++mi)
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const int row = wm * (MI * 8) + mi * 8 + gid;
                const int col = wn * (NI * 8) + ni * 8 + 2 * tig;
                const double2 c = *reinterpret_cast<const double2*>(Cg + static_cast<long>(row) * g.ldc + col);
                acc[mi][ni][0] = c.x;
                acc[mi][ni][1] = c.y;
            }
    }

#pragma unroll 1
    for (int kb = 0; kb < KB; ++kb) {
        if constexpr (MB) {
            mb_wait(&mb_full[kb % STAGES], (kb / STAGES) & 1);
        } else {
            cp_async_wait<STAGES - 2>();
            __syncthreads();
        }
        const double* as = As + (kb % STAGES) * A_ELEMS + (wm * (MI * 8) + gid) * LDS + tig;
        const double* bs = Bs + (kb % STAGES) * B_ELEMS + (wn * (NI * 8) + gid) * LDS + tig;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double a[MI], b[NI];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) a[mi] = ADD ? as[mi * 8 * LDS + kk * 4] : -as[mi * 8 * LDS + kk * 4];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) b[ni] = bs[ni * 8 * LDS + kk * 4];
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
            if (kk == 0) {
                // the refill of the stage read in the previous iteration is issued AFTER the first 16 MMAs of this
                // one are queued, so that the copy issue (address arithmetic + 6 LDGSTS) overlaps tensor work
                // instead of standing between the barrier and the first MMA
                const int r = kb + STAGES - 1;
                if (r < KB) {
                    if constexpr (MB) {
                        // stage r % STAGES was last read in iteration kb - 1: wait until all 8 warps have left it
                        if (r >= STAGES) mb_wait(&mb_empty[r % STAGES], ((r / STAGES) - 1) & 1);
                    }
                    load_stage(r % STAGES, r);
                    if constexpr (MB) mb_cp_async_arrive(&mb_full[r % STAGES]);
                }
                cp_async_commit();
            }
        }
        if constexpr (MB) {
            __syncwarp();
            if (lane == 0) mb_arrive(&mb_empty[kb % STAGES]);
        }
    }
    cp_async_wait<0>();

    if (g.late_c) {
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            double2 c[NI];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const int row = wm * (MI * 8) + mi * 8 + gid;
                const int col = wn * (NI * 8) + ni * 8 + 2 * tig;
                c[ni] = *reinterpret_cast<const double2*>(Cg + static_cast<long>(row) * g.ldc + col);
            }
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) {
                const int row = wm * (MI * 8) + mi * 8 + gid;
                const int col = wn * (NI * 8) + ni * 8 + 2 * tig;
                *reinterpret_cast<double2*>(Cg + static_cast<long>(row) * g.ldc + col) =
                    make_double2(c[ni].x + acc[mi][ni][0], c[ni].y + acc[mi][ni][1]);
            }
        }
        return;
    }
#pragma unroll
    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) {
            const int row = wm * (MI * 8) + mi * 8 + gid;
            const int col = wn * (NI * 8) + ni * 8 + 2 * tig;
            *reinterpret_cast<double2*>(Cg + static_cast<long>(row) * g.ldc + col) =
                make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
}

// ---------------------------------------------------------------------------
// Diagonal tile ahead of its column (look-ahead schedule of sweep.cu): C(128 x 128, lower 32 x 32 sub-tiles)
// -= A A^T with A = 128 x K panel rows (K = 128 or 256).  A whole 128 x 64 tile of K4 occupies one SM for
// 128*64*K / 64 clk (17 us at K = 256) -- too long for the serial panel chain -- so the tile is cut into its ten
// lower 32 x 32 sub-tiles, one CTA each, the 8 warps of a CTA splitting K (fixed-order reduction through shared
// memory: bit-reproducible).
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256) diag_tile_update_kernel(double* __restrict__ C, long ldc,
                                                               const double* __restrict__ A, long lda) {
    extern __shared__ __align__(16) double dsm[];      // [8 warps][1024]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int gid = lane >> 2, tig = lane & 3;
    int sr = 0;
    while ((sr + 1) * (sr + 2) / 2 <= static_cast<int>(blockIdx.x)) ++sr;
    const int sc = static_cast<int>(blockIdx.x) - sr * (sr + 1) / 2;
    constexpr int kw = K / 8;                           // K range of this warp
    const double* Ar = A + static_cast<long>(32 * sr + gid) * lda + warp * kw + tig;
    const double* Ac = A + static_cast<long>(32 * sc + gid) * lda + warp * kw + tig;
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    // all fragment loads of the warp's K range are issued before the first MMA: ONE L2 round trip on the serial chain
    double af[kw / 4][4], bf[kw / 4][4];
#pragma unroll
    for (int ks = 0; ks < kw / 4; ++ks) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) af[ks][mi] = Ar[static_cast<long>(mi) * 8 * lda + 4 * ks];
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) bf[ks][ni] = Ac[static_cast<long>(ni) * 8 * lda + 4 * ks];
    }
#pragma unroll
    for (int ks = 0; ks < kw / 4; ++ks)
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) dmma884_t(acc[mi][ni][0], acc[mi][ni][1], af[ks][mi], bf[ks][ni]);
    double* mine = dsm + warp * 1024;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
            *reinterpret_cast<double2*>(mine + ((mi * 4 + ni) * 32 + lane) * 2) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
    __syncthreads();
    for (int q = tid; q < 512; q += 256) {
        const int t = q >> 5, l = q & 31;
        const int row = 32 * sr + (t >> 2) * 8 + (l >> 2), col = 32 * sc + (t & 3) * 8 + 2 * (l & 3);
        double2 sum = make_double2(0.0, 0.0);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const double2 v = *reinterpret_cast<const double2*>(dsm + w * 1024 + q * 2);
            sum.x += v.x;
            sum.y += v.y;
        }
        double2* cp = reinterpret_cast<double2*>(C + static_cast<long>(row) * ldc + col);
        double2 c = *cp;
        c.x -= sum.x;
        c.y -= sum.y;
        *cp = c;
    }
}

}  // namespace

int gemm_smem_bytes() { return 3 * (EGX_NB + 128) * 20 * static_cast<int>(sizeof(double)); }

void launch_potrf_diag(double* Akk, long ld, int* info, int base_index, double* Dinv, cudaStream_t s) {
    // EGX_POTRF_V=1: the r01 kernel (CTA-wide barrier per column), kept for A/B
    static const int version = getenv("EGX_POTRF_V") != nullptr ? atoi(getenv("EGX_POTRF_V")) : 2;
    if (version == 1) {
        potrf_diag_kernel<<<1, 256, 0, s>>>(Akk, ld, info, base_index, Dinv);
        return;
    }
    static bool configured_dev[64] = {false};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& configured = configured_dev[dev_ & 63];
    const int smem = (EGX_NB * P2_LD + 4 * 64) * static_cast<int>(sizeof(double));
    if (!configured) {
        cudaFuncSetAttribute(potrf_diag2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    potrf_diag2_kernel<<<1, 256, smem, s>>>(Akk, ld, info, base_index, Dinv);
}

void launch_diag_tile_update(double* C, long ldc, const double* A, long lda, int K, cudaStream_t s) {
    static bool configured_dev[64] = {false};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& configured = configured_dev[dev_ & 63];
    const int smem = 8 * 1024 * static_cast<int>(sizeof(double));
    if (!configured) {
        cudaFuncSetAttribute(diag_tile_update_kernel<EGX_NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(diag_tile_update_kernel<2 * EGX_NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    if (K == EGX_NB) diag_tile_update_kernel<EGX_NB><<<10, 256, smem, s>>>(C, ldc, A, lda);
    else diag_tile_update_kernel<2 * EGX_NB><<<10, 256, smem, s>>>(C, ldc, A, lda);
}

void launch_trsm_rows(double* X, long ldx, const double* Lkk, long ldl, const double* Dinv, double* P, long ldp,
                      int nblocks64, cudaStream_t s, double* rmaxq) {
    static bool configured_dev[64] = {false};
    static int sms_dev[64] = {0};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& configured = configured_dev[dev_ & 63];   // the attribute is per device (one process may drive several)
    constexpr int smem_fixed = (4 * 32 * TR_LDD + 96 * TR_LDL) * static_cast<int>(sizeof(double));
    constexpr int smem64 = 64 * (TR_LDX + TR_LDD) * static_cast<int>(sizeof(double)) + smem_fixed;
    constexpr int smem32 = 32 * (TR_LDX + TR_LDD) * static_cast<int>(sizeof(double)) + smem_fixed;
    if (!configured) {
        cudaFuncSetAttribute(trsm_rows_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64);
        cudaFuncSetAttribute(trsm_rows_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem32);
        cudaDeviceGetAttribute(&sms_dev[dev_ & 63], cudaDevAttrMultiProcessorCount, dev_);
        configured = true;
    }
    if (nblocks64 <= 0) return;
    // 32-row slabs while they still make one wave (one CTA per SM): half the FP64 work per CTA of a kernel that sits on
    // the serial chain of the factorisation (EGX_TRSM_ROWS=64 keeps the 64-row slabs)
    static const int force64 = getenv("EGX_TRSM_ROWS") != nullptr && atoi(getenv("EGX_TRSM_ROWS")) == 64;
    if (!force64 && 2 * nblocks64 <= sms_dev[dev_ & 63])
        trsm_rows_kernel<32><<<2 * nblocks64, 256, smem32, s>>>(X, ldx, Lkk, ldl, Dinv, P, ldp, rmaxq);
    else
        trsm_rows_kernel<64><<<nblocks64, 256, smem64, s>>>(X, ldx, Lkk, ldl, Dinv, P, ldp, rmaxq);
}

static int gemm_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e != nullptr ? atoi(e) : dflt;
}
static int gemm_bn() {
    static int bn = 0;
    if (bn == 0) bn = (gemm_env("EGX_GEMM_BN", 64) == 128) ? 128 : 64;
    return bn;
}
static int gemm_bk() {
    static int bk = 0;
    if (bk == 0) bk = (gemm_env("EGX_GEMM_BK", 16) == 32) ? 32 : 16;
    return bk;
}

template <int BN, int BK, int STAGES, bool MB = false>
static void gemm_launch_variant(const GemmArgs& g, dim3 grid, cudaStream_t s) {
    constexpr int smem = STAGES * (EGX_NB + BN) * (BK + 4) * static_cast<int>(sizeof(double));
    static bool configured_dev[64] = {false};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& configured = configured_dev[dev_ & 63];   // the attribute is per device (one process may drive several)
    if (!configured) {
        cudaFuncSetAttribute(gemm_nt_sub_kernel<BN, false, BK, STAGES, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(gemm_nt_sub_kernel<BN, true, BK, STAGES, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    if (g.add) gemm_nt_sub_kernel<BN, true, BK, STAGES, MB><<<grid, 256, smem, s>>>(g);
    else gemm_nt_sub_kernel<BN, false, BK, STAGES, MB><<<grid, 256, smem, s>>>(g);
}

void launch_gemm_nt_sub(const GemmArgs& g_in, cudaStream_t s) {
    // measured at n = 8192: 7.03 (late) vs 7.00 ms (early) per batched evaluation -- the C-tile latency is already
    // hidden by the co-resident CTA; kept as a switch
    static const int late_c = gemm_env("EGX_GEMM_LATEC", 0);
    GemmArgs g = g_in;
    g.late_c = late_c;
    const int S = (gemm_bn() == 128) ? 1 : 2;
    int tiles;
    if (g.tri > 0) tiles = S * g.tri * (g.tri + 1) / 2 + (g.Mt - g.tri) * S * g.tri;
    else tiles = g.Mt * S * g.Nt;
    if (tiles <= 0) return;
    const dim3 grid(tiles, g.splits > 1 ? g.splits : 1);
    const bool bk32 = gemm_bk() == 32 && (g.K % 32 == 0);
    if (S == 1) {
        if (bk32) gemm_launch_variant<128, 32, 2>(g, grid, s);
        else gemm_launch_variant<128, 16, 3>(g, grid, s);
    } else {
        // measured at n = 8192, 48 evaluations in a batch: 6.68 (bar.sync) vs 6.56 ms (mbarrier pipeline) per evaluation
        static const int use_mb = gemm_env("EGX_GEMM_MB", 1);
        if (bk32) gemm_launch_variant<64, 32, 2>(g, grid, s);
        else if (use_mb) gemm_launch_variant<64, 16, 3, true>(g, grid, s);
        else gemm_launch_variant<64, 16, 3>(g, grid, s);
    }
}
