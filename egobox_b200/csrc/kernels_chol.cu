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
#include "common.cuh"
#include "../../include/egobox_gpu.h"

namespace {

// ---------------------------------------------------------------------------
// K3: diagonal block.  512 threads, S[128][129] in shared memory.
// ---------------------------------------------------------------------------
constexpr int PD_LD = 129;

__global__ void __launch_bounds__(512) potrf_diag_kernel(double* __restrict__ A, long ld, int* __restrict__ info,
                                                         int base_index) {
    extern __shared__ double S[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    // load the lower triangle (rows are contiguous in HBM)
    for (int r = warp; r < EGX_NB; r += 16)
        for (int c = lane; c <= r; c += 32) S[r * PD_LD + c] = A[static_cast<long>(r) * ld + c];
    __syncthreads();

    for (int j = 0; j < EGX_NB; ++j) {
        const double dj = S[j * PD_LD + j];
        // LAPACK dpotrf semantics: fail on a non-positive or NaN pivot
        if (!(dj > 0.0)) {
            if (tid == 0) atomicCAS(info, 0, base_index + j + 1);
        }
        const double ljj = sqrt(dj);
        __syncthreads();   // everyone has read the pivot before it is overwritten
        if (tid == 0) S[j * PD_LD + j] = ljj;
        for (int i = j + 1 + tid; i < EGX_NB; i += 512) S[i * PD_LD + j] /= ljj;
        __syncthreads();
        // trailing update of the lower triangle: S[i][c] -= S[i][j] * S[c][j],  j < c <= i
        for (int i = j + 1 + warp; i < EGX_NB; i += 16) {
            const double lij = S[i * PD_LD + j];
            for (int c = j + 1 + lane; c <= i; c += 32) S[i * PD_LD + c] -= lij * S[c * PD_LD + j];
        }
        // the next iteration's pivot read happens after this barrier
        __syncthreads();
    }
    for (int r = warp; r < EGX_NB; r += 16)
        for (int c = lane; c <= r; c += 32) A[static_cast<long>(r) * ld + c] = S[r * PD_LD + c];
}

// ---------------------------------------------------------------------------
// K5: X (64 x 128 slab, in place) <- X * L^-T,  L = 128 x 128 lower block.
//   x[r][c] = (a[r][c] - sum_{j<c} x[r][j] L[c][j]) / L[c][c]
// 256 threads: 4 lanes per row (same warp), columns processed in order, the
// j-sum is split over the 4 lanes (j = q mod 4) and combined with two shuffles.
// ---------------------------------------------------------------------------
constexpr int TR_ROWS = 64;
constexpr int TR_LDX = 132;

__global__ void __launch_bounds__(256) trsm_rows_kernel(double* __restrict__ X, long ldx,
                                                        const double* __restrict__ L, long ldl,
                                                        double* __restrict__ P) {
    extern __shared__ double sm[];
    double* Ls = sm;                       // [128][128]
    double* Xs = sm + EGX_NB * EGX_NB;     // [64][132]
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    double* Xg = X + static_cast<long>(blockIdx.x) * TR_ROWS * ldx;

    for (int r = warp; r < EGX_NB; r += 8) {
        for (int c = lane; c < EGX_NB; c += 32) Ls[r * EGX_NB + c] = (c <= r) ? L[static_cast<long>(r) * ldl + c] : 0.0;
    }
    for (int r = warp; r < TR_ROWS; r += 8) {
        for (int c = lane; c < EGX_NB; c += 32) Xs[r * TR_LDX + c] = Xg[static_cast<long>(r) * ldx + c];
    }
    __syncthreads();

    const int r = tid >> 2, q = tid & 3;
    double* xr = Xs + r * TR_LDX;
    for (int c = 0; c < EGX_NB; ++c) {
        const double* lc = Ls + c * EGX_NB;
        double s0 = 0.0, s1 = 0.0;
        int j = q;
        for (; j + 4 < c; j += 8) {
            s0 += xr[j] * lc[j];
            s1 += xr[j + 4] * lc[j + 4];
        }
        if (j < c) s0 += xr[j] * lc[j];
        double s = s0 + s1;
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const double x = (xr[c] - s) / lc[c];
        __syncwarp();
        if (q == 0) xr[c] = x;
        __syncwarp();
    }
    __syncthreads();
    double* Pg = (P != nullptr) ? P + static_cast<long>(blockIdx.x) * TR_ROWS * EGX_NB : nullptr;
    for (int rr = warp; rr < TR_ROWS; rr += 8) {
        for (int c = lane; c < EGX_NB; c += 32) {
            const double v = Xs[rr * TR_LDX + c];
            Xg[static_cast<long>(rr) * ldx + c] = v;
            if (Pg != nullptr) Pg[rr * EGX_NB + c] = v;
        }
    }
}

// ---------------------------------------------------------------------------
// K4: C(128x128 tile) -= A(128 x K) * B(128 x K)^T,  K = 128, all row-major fp64.
// 256 threads = 8 warps (2 x 4), warp tile 64 x 32, mma.m8n8k4.f64, 3-stage
// cp.async pipeline over K in steps of 16.  The accumulators are initialised
// with the C tile and A fragments are negated, so the epilogue is a pure store.
// ---------------------------------------------------------------------------
constexpr int GM_BK = 16;
constexpr int GM_LDS = 20;                        // 16 + 4 pad doubles: conflict-free fragment loads
constexpr int GM_STAGES = 3;
constexpr int GM_TILE_ELEMS = EGX_NB * GM_LDS;    // per operand per stage
constexpr int GM_K = EGX_NB;

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

__device__ __forceinline__ void gemm_tile_decode(const GemmArgs& g, int t, int& r, int& c) {
    if (g.tri > 0) {
        const int ntri = g.tri * (g.tri + 1) / 2;
        if (t < ntri) {
            int rr = static_cast<int>((sqrt(8.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
            while ((rr + 1) * (rr + 2) / 2 <= t) ++rr;
            while (rr * (rr + 1) / 2 > t) --rr;
            r = rr;
            c = t - rr * (rr + 1) / 2;
        } else {
            const int u = t - ntri;
            r = g.tri + u / g.tri;
            c = u % g.tri;
        }
    } else {
        r = t / g.Nt;
        c = t % g.Nt;
    }
}

__global__ void __launch_bounds__(256, 1) gemm_nt_sub_kernel(const GemmArgs g) {
    extern __shared__ __align__(16) double gsm[];
    double* As = gsm;                                   // [STAGES][128][20]
    double* Bs = gsm + GM_STAGES * GM_TILE_ELEMS;       // [STAGES][128][20]

    int tr, tc;
    gemm_tile_decode(g, blockIdx.x, tr, tc);
    const double* Ag = g.A + static_cast<long>(tr) * EGX_NB * g.lda;
    const double* Bg = g.B + static_cast<long>(tc) * EGX_NB * g.ldb;
    double* Cg = g.C + static_cast<long>(tr) * EGX_NB * g.ldc + static_cast<long>(tc) * EGX_NB;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 2, wn = warp & 3;
    const int gid = lane >> 2, tig = lane & 3;

    auto load_stage = [&](int stage, int kb) {
        double* as = As + stage * GM_TILE_ELEMS;
        double* bs = Bs + stage * GM_TILE_ELEMS;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + i * 256;          // 1024 16-byte chunks per operand
            const int row = idx >> 3, ch = idx & 7;
            cp_async16(as + row * GM_LDS + ch * 2, Ag + static_cast<long>(row) * g.lda + kb * GM_BK + ch * 2);
            cp_async16(bs + row * GM_LDS + ch * 2, Bg + static_cast<long>(row) * g.ldb + kb * GM_BK + ch * 2);
        }
    };

    constexpr int KB = GM_K / GM_BK;   // 8
#pragma unroll
    for (int s = 0; s < GM_STAGES - 1; ++s) {
        load_stage(s, s);
        cp_async_commit();
    }

    // accumulators start as the C tile
    double acc[8][4][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int row = wm * 64 + mi * 8 + gid;
            const int col = wn * 32 + ni * 8 + 2 * tig;
            const double2 c = *reinterpret_cast<const double2*>(Cg + static_cast<long>(row) * g.ldc + col);
            acc[mi][ni][0] = c.x;
            acc[mi][ni][1] = c.y;
        }

#pragma unroll 1
    for (int kb = 0; kb < KB; ++kb) {
        cp_async_wait<GM_STAGES - 2>();
        __syncthreads();
        if (kb + GM_STAGES - 1 < KB) load_stage((kb + GM_STAGES - 1) % GM_STAGES, kb + GM_STAGES - 1);
        cp_async_commit();
        const double* as = As + (kb % GM_STAGES) * GM_TILE_ELEMS + (wm * 64 + gid) * GM_LDS + tig;
        const double* bs = Bs + (kb % GM_STAGES) * GM_TILE_ELEMS + (wn * 32 + gid) * GM_LDS + tig;
#pragma unroll
        for (int kk = 0; kk < GM_BK / 4; ++kk) {
            double a[8], b[4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) a[mi] = -as[mi * 8 * GM_LDS + kk * 4];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) b[ni] = bs[ni * 8 * GM_LDS + kk * 4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int row = wm * 64 + mi * 8 + gid;
            const int col = wn * 32 + ni * 8 + 2 * tig;
            *reinterpret_cast<double2*>(Cg + static_cast<long>(row) * g.ldc + col) =
                make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
}

}  // namespace

int gemm_smem_bytes() { return 2 * GM_STAGES * GM_TILE_ELEMS * static_cast<int>(sizeof(double)); }

void launch_potrf_diag(double* Akk, long ld, int* info, int base_index, cudaStream_t s) {
    static bool configured = false;
    const int smem = EGX_NB * PD_LD * sizeof(double);
    if (!configured) {
        cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    potrf_diag_kernel<<<1, 512, smem, s>>>(Akk, ld, info, base_index);
}

void launch_trsm_rows(double* X, long ldx, const double* Lkk, long ldl, double* P, int nblocks64, cudaStream_t s) {
    static bool configured = false;
    const int smem = (EGX_NB * EGX_NB + TR_ROWS * TR_LDX) * sizeof(double);
    if (!configured) {
        cudaFuncSetAttribute(trsm_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    if (nblocks64 <= 0) return;
    trsm_rows_kernel<<<nblocks64, 256, smem, s>>>(X, ldx, Lkk, ldl, P);
}

void launch_gemm_nt_sub(const GemmArgs& g, cudaStream_t s) {
    static bool configured = false;
    const int smem = gemm_smem_bytes();
    if (!configured) {
        cudaFuncSetAttribute(gemm_nt_sub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        configured = true;
    }
    int tiles;
    if (g.tri > 0) tiles = g.tri * (g.tri + 1) / 2 + (g.Mt - g.tri) * g.tri;
    else tiles = g.Mt * g.Nt;
    if (tiles <= 0) return;
    gemm_nt_sub_kernel<<<tiles, 256, smem, s>>>(g);
}
