// K4o: the trailing SYRK update  C -= P P^T  (K = 256, fp64 result) on the 5th-generation tensor cores.
//
// tcgen05 has no f64 kind, so the fp64 panel P is split ONCE per panel pair into 8 signed 7-bit slices per entry
// (int8), relative to a power-of-two scale per ROW (Ozaki splitting):
//     P[i][k] = 2^e_i * sum_{s=0..7} S_s[i][k] * 2^-7(s+1)   (+ truncation below 2^(e_i-56)),   |S_s| <= 127.
// A product of two slices is exact in int32 (127^2 * 256 < 2^22; the <= 8 slice pairs that share a weight sum to
// < 2^25), so   (P P^T)[i][j] = 2^(e_i+e_j) * sum_g 2^-7(g+2) * ACC_g[i][j],   ACC_g = sum_{p+q=g} S_p S_q^T
// with every ACC_g an exact integer matrix produced by `tcgen05.mma.kind::i8` (SASS UTCIMMA) into tensor memory.
// Slice pairs with p + q > 7 are dropped: they sit below 2^-56 * K * 7 of |row_i|_max |row_j|_max, i.e. at the level
// of the rounding of the fp64 dot product itself.  36 int8 MMAs replace one fp64 MMA: 8192 int8 MAC/clk/SM
// (measured, tools/micro/i8mma_probe.cu) against 64 fp64 FMA/clk/SM on the DMMA pipe.
//
// Reference being replaced: the trailing update inside `r_mx.cholesky()` gp/src/algorithm.rs:1004 / :1077.
//
// Data layout.  The slices of a (rows x 256) panel live in HBM as
//     S[row block rb of 128][K step ks of 32][slice s][K chunk kc of 16][row r of 128][16 bytes]
// so that (i) the 128 x 32 operand of one MMA is the canonical K-major / no-swizzle shared-memory layout of the
// tensor core (core matrix = 8 rows x 16 bytes contiguous; 128 bytes between row groups, 2048 bytes between the two
// K chunks) and (ii) all slices a CTA needs for one K step are ONE contiguous run -> one 1-D TMA bulk copy per
// operand and stage (cp.async.bulk, SASS UBLKCP), no tensor map.
//
// Kernel.  One CTA per 128 x 128 tile of C, 10 warps: warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
// warps 2..9 = epilogue.  Tensor memory holds 4 accumulators of 128 columns (all 512 columns), so the 8 weights are
// done in two passes over K: pass 0 = weights 0..3 (slices 0..3, 10 pairs), pass 1 = weights 4..7 (slices 0..7, 26
// pairs).  After a pass the epilogue reads the 4 accumulators (tcgen05.ld 16x256b: a quad of lanes holds 8
// consecutive columns of a row, so the read-modify-write of C uses full 32-byte sectors), folds them exactly
// into one int64  T = (a0<<21) + (a1<<14) + (a2<<7) + a3 , converts once and applies
//     C[i][j] -= T * 2^(e_i + e_j - 35 - 28 pass).
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "../../include/egobox_gpu.h"

namespace {

constexpr int OZ_SLICES = 8;
constexpr int OZ_K = 256;                        // contraction length (a panel pair)
constexpr int OZ_KSTEPS = OZ_K / 32;             // MMA K = 32 bytes
constexpr int OZ_SLICE_STEP_BYTES = 128 * 32;    // one slice, one K step, 128 rows
constexpr int OZ_STAGE_OPERAND = OZ_SLICES * OZ_SLICE_STEP_BYTES;   // 32 KB
constexpr int OZ_STAGE_BYTES = 2 * OZ_STAGE_OPERAND;                // A + B
constexpr int OZ_STAGES = 3;
constexpr int OZ_THREADS = 320;
constexpr long OZ_RB_BYTES = static_cast<long>(OZ_KSTEPS) * OZ_STAGE_OPERAND;   // slices of one 128-row block: 256 KB

__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void oz_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void oz_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(oz_smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void oz_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(oz_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(oz_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     oz_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(oz_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void oz_umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(oz_smem_u32(bar))
                 : "memory");
}
// K-major, no swizzle: LBO = byte distance of the two 16-byte K chunks, SBO = byte distance of 8-row groups
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
    return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (static_cast<uint64_t>(2048 >> 4) << 16) |
           (static_cast<uint64_t>(128 >> 4) << 32) | (static_cast<uint64_t>(1) << 46);
}
// instruction descriptor: D = s32, A = B = signed int8, both K-major, M = 128, N = 128
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(128 >> 3) << 17) |
                              (static_cast<uint32_t>(128 >> 4) << 24);
__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(OZ_IDESC), "r"(accumulate)
        : "memory");
}
// 16 lanes x 256 bit, 4 repetitions along the columns: 32 columns of 16 rows; thread t holds, for repetition j,
// v[4j+0..1] = row (t / 4), columns 8 j + 2 (t % 4) + {0, 1} and v[4j+2..3] = row (t / 4) + 8, same columns
__device__ __forceinline__ void oz_tmem_ld(uint32_t addr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr));
}

__device__ __forceinline__ void oz_tile_decode(int t, int tri, int& r, int& c) {
    const int ntri = tri * (tri + 1) / 2;
    if (t < ntri) {
        int rr = static_cast<int>((sqrtf(8.0f * static_cast<float>(t) + 1.0f) - 1.0f) * 0.5f);
        while ((rr + 1) * (rr + 2) / 2 <= t) ++rr;
        while (rr * (rr + 1) / 2 > t) --rr;
        r = rr;
        c = t - rr * (rr + 1) / 2;
    } else {
        const int u = t - ntri;
        r = tri + u / tri;
        c = u % tri;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// slicing: (rows x 256) fp64 panel -> row scales 2^e_i and the int8 slices
// ---------------------------------------------------------------------------------------------------------------
// one warp per row: 2^e with |row|_max < 2^e (0 for an all-zero row)
__global__ void __launch_bounds__(256) ozaki_rowscale_kernel(const double* __restrict__ P, long ldp, int rows,
                                                             double* __restrict__ rscale) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const double* p = P + static_cast<long>(row) * ldp;
    double m = 0.0;
#pragma unroll
    for (int j = 0; j < OZ_K / 64; ++j) {
        const double2 v = *reinterpret_cast<const double2*>(p + 64 * j + 2 * lane);
        m = fmax(m, fmax(fabs(v.x), fabs(v.y)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        double s = 0.0;
        if (m > 0.0 && m < 1.0e300) s = scalbn(1.0, ilogb(m) + 1);
        rscale[row] = s;
    }
}

// thread = (row, 16-entry K chunk): 16 doubles in, 8 x 16 bytes out
__global__ void __launch_bounds__(128) ozaki_slice_kernel(const double* __restrict__ P, long ldp,
                                                          const double* __restrict__ rscale, int8_t* __restrict__ S) {
    const int rb = blockIdx.x, chunk = blockIdx.y, rr = threadIdx.x;
    const long row = static_cast<long>(rb) * 128 + rr;
    const double sc = rscale[row];
    const double inv = sc > 0.0 ? 72057594037927936.0 / sc : 0.0;           // 2^56 / 2^e
    const double* p = P + row * ldp + chunk * 16;
    uint32_t w[OZ_SLICES][4];
#pragma unroll
    for (int s = 0; s < OZ_SLICES; ++s)
#pragma unroll
        for (int q = 0; q < 4; ++q) w[s][q] = 0u;
#pragma unroll
    for (int e2 = 0; e2 < 8; ++e2) {
        const double2 v = *reinterpret_cast<const double2*>(p + 2 * e2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int e = 2 * e2 + h;
            const long long t = __double2ll_rz((h ? v.y : v.x) * inv);     // |t| < 2^56, exact scaling
            const unsigned long long mag = static_cast<unsigned long long>(t < 0 ? -t : t);
#pragma unroll
            for (int s = 0; s < OZ_SLICES; ++s) {
                int d = static_cast<int>((mag >> (7 * (7 - s))) & 127ull);
                if (t < 0) d = -d;
                w[s][e >> 2] |= (static_cast<uint32_t>(d) & 0xffu) << (8 * (e & 3));
            }
        }
    }
    const int ks = chunk >> 1, kc = chunk & 1;
#pragma unroll
    for (int s = 0; s < OZ_SLICES; ++s) {
        int8_t* dst = S + static_cast<long>(rb) * OZ_RB_BYTES +
                      ((static_cast<long>(ks) * OZ_SLICES + s) * 2 + kc) * 2048 + rr * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the update kernel
// ---------------------------------------------------------------------------------------------------------------
struct OzakiArgs {
    double* C;
    long ldc;
    const int8_t* S;          // slices of the panel rows; row block 0 = first tile row / column of C
    const double* rscale;     // 2^e per panel row
    int Mt, tri;              // tile rows; the first `tri` rows are triangular (c <= r), the others full (c < tri)
};

struct __align__(8) OzBarriers {
    uint64_t full[OZ_STAGES], empty[OZ_STAGES], acc_full, acc_empty;
    uint32_t tmem_base, pad_;
};

__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_syrk_kernel(const OzakiArgs g) {
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    OzBarriers* bars = reinterpret_cast<OzBarriers*>(oz_smem + OZ_STAGES * OZ_STAGE_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int tr, tc;
    oz_tile_decode(blockIdx.x, g.tri, tr, tc);

    if (tid == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) {
            oz_mbar_init(&bars->full[s], 1);
            oz_mbar_init(&bars->empty[s], 1);
        }
        oz_mbar_init(&bars->acc_full, 1);
        oz_mbar_init(&bars->acc_empty, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&bars->tmem_base)),
                     "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0) {
        // ===== producer: one elected lane streams the K steps of both passes through the stage ring =====
        if (lane == 0) {
            const int8_t* Ag = g.S + static_cast<long>(tr) * OZ_RB_BYTES;
            const int8_t* Bg = g.S + static_cast<long>(tc) * OZ_RB_BYTES;
            for (int it = 0; it < 2 * OZ_KSTEPS; ++it) {
                const int stage = it % OZ_STAGES, pass = it / OZ_KSTEPS, ks = it % OZ_KSTEPS;
                const uint32_t bytes = (pass == 0 ? 4 : 8) * OZ_SLICE_STEP_BYTES;
                oz_mbar_wait(&bars->empty[stage], ((it / OZ_STAGES) & 1) ^ 1);
                oz_mbar_expect_tx(&bars->full[stage], 2 * bytes);
                unsigned char* st = oz_smem + stage * OZ_STAGE_BYTES;
                oz_bulk_g2s(st, Ag + static_cast<long>(ks) * OZ_STAGE_OPERAND, bytes, &bars->full[stage]);
                oz_bulk_g2s(st + OZ_STAGE_OPERAND, Bg + static_cast<long>(ks) * OZ_STAGE_OPERAND, bytes, &bars->full[stage]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (int it = 0; it < 2 * OZ_KSTEPS; ++it) {
                const int stage = it % OZ_STAGES, pass = it / OZ_KSTEPS, ks = it % OZ_KSTEPS;
                if (it == OZ_KSTEPS) {                     // pass 1 reuses the accumulators: wait for the drain
                    oz_mbar_wait(&bars->acc_empty, 0);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                oz_mbar_wait(&bars->full[stage], (it / OZ_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = oz_smem_u32(oz_smem + stage * OZ_STAGE_BYTES);
                const uint32_t sb = sa + OZ_STAGE_OPERAND;
                const int g0 = pass * 4;
#pragma unroll
                for (int gg = 0; gg < 4; ++gg) {
                    const int w = g0 + gg;                 // weight p + q
                    for (int p = 0; p <= w; ++p) {
                        const int q = w - p;
                        if (p > 7 || q > 7) continue;
                        oz_mma(tmem + gg * 128, oz_desc(sa + p * OZ_SLICE_STEP_BYTES), oz_desc(sb + q * OZ_SLICE_STEP_BYTES),
                               (ks > 0 || p > 0) ? 1u : 0u);
                    }
                }
                oz_umma_commit(&bars->empty[stage]);       // frees the stage when these MMAs have read it
                if (ks == OZ_KSTEPS - 1) oz_umma_commit(&bars->acc_full);
            }
        }
    } else {
        // ===== epilogue: 8 warps; lane quarter = warp % 4, column half = (warp - 2) / 4 =====
        const int quarter = warp & 3, chalf = (warp - 2) >> 2;
        const int r_in = lane >> 2, cq = 2 * (lane & 3);
        double* Cb = g.C + static_cast<long>(tr) * 128 * g.ldc + static_cast<long>(tc) * 128;
        const double* rsA = g.rscale + static_cast<long>(tr) * 128;
        const double* rsB = g.rscale + static_cast<long>(tc) * 128;
        for (int pass = 0; pass < 2; ++pass) {
            oz_mbar_wait(&bars->acc_full, pass);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const double wscale = pass == 0 ? 2.9103830456733704e-11 /* 2^-35 */ : 1.0842021724855044e-19 /* 2^-63 */;
#pragma unroll 1
            for (int rh = 0; rh < 2; ++rh) {
                const int row0 = 32 * quarter + 16 * rh;          // TMEM lane of row r_in = 0 of this block
                const double sr0 = rsA[row0 + r_in] * wscale, sr1 = rsA[row0 + r_in + 8] * wscale;
#pragma unroll 1
                for (int cc = 0; cc < 2; ++cc) {
                    const int col0 = 64 * chalf + 32 * cc;
                    uint32_t a[4][16];
#pragma unroll
                    for (int gg = 0; gg < 4; ++gg)
                        oz_tmem_ld(tmem + (static_cast<uint32_t>(row0) << 16) + gg * 128 + col0, a[gg]);
                    // the C values of this thread: rows row0 + r_in (+8), columns col0 + 8 j + cq + {0,1}
                    double2 c0[4], c1[4];
                    double* p0 = Cb + static_cast<long>(row0 + r_in) * g.ldc + col0 + cq;
                    double* p1 = p0 + 8 * g.ldc;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        c0[j] = *reinterpret_cast<const double2*>(p0 + 8 * j);
                        c1[j] = *reinterpret_cast<const double2*>(p1 + 8 * j);
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double sc0 = rsB[col0 + 8 * j + cq], sc1 = rsB[col0 + 8 * j + cq + 1];
                        long long t[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            t[e] = static_cast<long long>(static_cast<int32_t>(a[0][4 * j + e])) * 2097152LL +
                                   static_cast<long long>(static_cast<int32_t>(a[1][4 * j + e])) * 16384LL +
                                   static_cast<long long>(static_cast<int32_t>(a[2][4 * j + e])) * 128LL +
                                   static_cast<long long>(static_cast<int32_t>(a[3][4 * j + e]));
                        c0[j].x = fma(-static_cast<double>(t[0]), sr0 * sc0, c0[j].x);
                        c0[j].y = fma(-static_cast<double>(t[1]), sr0 * sc1, c0[j].y);
                        c1[j].x = fma(-static_cast<double>(t[2]), sr1 * sc0, c1[j].x);
                        c1[j].y = fma(-static_cast<double>(t[3]), sr1 * sc1, c1[j].y);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        *reinterpret_cast<double2*>(p0 + 8 * j) = c0[j];
                        *reinterpret_cast<double2*>(p1 + 8 * j) = c1[j];
                    }
                }
            }
            if (pass == 0) {
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) oz_mbar_arrive(&bars->acc_empty);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

constexpr int OZ_SMEM_BYTES = OZ_STAGES * OZ_STAGE_BYTES + static_cast<int>(sizeof(OzBarriers));

}  // namespace

// bytes of slice storage / scale storage for a panel of `rows` rows (multiple of 128)
size_t ozaki_slice_bytes(long rows) { return static_cast<size_t>(rows / 128) * OZ_RB_BYTES; }

// P: rows x 256 (ldp), rows a multiple of 128 -> rscale[rows], S
void launch_ozaki_slice(const double* P, long ldp, int rows, double* rscale, int8_t* S, cudaStream_t s) {
    if (rows <= 0) return;
    ozaki_rowscale_kernel<<<(rows + 7) / 8, 256, 0, s>>>(P, ldp, rows, rscale);
    ozaki_slice_kernel<<<dim3(rows / 128, OZ_K / 16), 128, 0, s>>>(P, ldp, rscale, S);
}

// C (tile rows Mt, first `tri` triangular) -= P P^T from the slices of P
void launch_ozaki_syrk(double* C, long ldc, const int8_t* S, const double* rscale, int Mt, int tri, cudaStream_t s) {
    static bool configured_dev[64] = {false};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& configured = configured_dev[dev_ & 63];
    if (!configured) {
        cudaFuncSetAttribute(ozaki_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES);
        configured = true;
    }
    const int tiles = tri * (tri + 1) / 2 + (Mt - tri) * tri;
    if (tiles <= 0) return;
    OzakiArgs g{C, ldc, S, rscale, Mt, tri};
    ozaki_syrk_kernel<<<tiles, OZ_THREADS, OZ_SMEM_BYTES, s>>>(g);
}
