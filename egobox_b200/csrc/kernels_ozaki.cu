// K4o: the trailing SYRK update  C -= P P^T  (K = 256, fp64 result) on the 5th-generation tensor cores.
//
// tcgen05 has no f64 kind, so the fp64 panel P is split ONCE per panel pair into 7 balanced base-256 digits per entry
// (int8, the full range -128..127), relative to a power-of-two scale per ROW (Ozaki splitting):
//     P[i][k] = s_i * sum_{d=0..6} S_d[i][k] * 256^-(d+1)   (rounded at s_i 2^-57),   s_i = 2^e >= 4 |row i|_max.
// A product of two digit slices is exact in int32 (2^14 * 256 = 2^22 per pair; the <= 7 pairs that share a weight stay
// below 2^25), so   (P P^T)[i][j] = s_i s_j * sum_w 256^-(w+2) * ACC_w[i][j],   ACC_w = sum_{p+q=w} S_p S_q^T
// with every ACC_w an exact integer matrix produced by `tcgen05.mma.kind::i8` (SASS UTCIMMA) into tensor memory.
// Digit pairs with p + q > 6 are dropped: they sit below 2^-47 of s_i s_j, i.e. at the level of the rounding of the
// fp64 dot product itself (measured against a long-double loop in tools/micro/ozaki_probe.cu).  28 int8 MMAs replace
// one fp64 MMA: 8190 int8 MAC/clk/SM (measured, tools/micro/i8mma_probe.cu) against 64 fp64 FMA/clk/SM on the DMMA pipe.
//
// Reference being replaced: the trailing update inside `r_mx.cholesky()` gp/src/algorithm.rs:1004 / :1077.
//
// Data layout.  The slices of a (rows x 256) panel live in HBM as
//     S[row block rb of 128][K step ks of 32][slot s of 8][row group of 8][K chunk kc of 16][row in group][16 bytes]
// (slots 0..6 = the digits, slot 7 unused) so that (i) the 128 x 32 operand of one MMA is the canonical K-major /
// no-swizzle shared-memory layout of the tensor core (core matrix = 8 rows x 16 bytes contiguous; 128 bytes between
// the two K chunks, 256 bytes between row groups), (ii) all slices a CTA needs for one K step are ONE contiguous run
// -> one 1-D TMA bulk copy per operand and stage (cp.async.bulk, SASS UBLKCP), no tensor map, and (iii) any 64-row
// half of a slice is contiguous too (the two-CTA kernel splits the B rows between the CTAs of a pair).
//
// Kernel.  One CTA per 128 x 128 tile of C, 10 warps: warp 0 = TMA producer, warp 1 = MMA issuer (one thread),
// warps 2..9 = epilogue.  Tensor memory holds 4 accumulators of 128 columns (all 512 columns), so the 7 weights are
// done in two passes over K: pass 0 = weights 0..3 (digits 0..3, 10 pairs), pass 1 = weights 4..6 (digits 0..6, 18
// pairs).  After a pass the epilogue reads the accumulators (tcgen05.ld 16x256b: a quad of lanes holds 8 consecutive
// columns of a row, so C moves in full 32-byte sectors), folds them exactly into one int64
//     T = ((a0 * 256 + a1) * 256 + a2) * 256 + a3 ,  converts once and applies  C[i][j] -= T * s_i s_j 2^-40
// (pass 1: three accumulators, 2^-64).  An epilogue thread keeps its 64 entries of C in registers across both passes.
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include <cuda.h>              // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"
#include "../../include/egobox_gpu.h"

namespace {

constexpr int OZ_SLICES = 7;                     // balanced base-256 digits per entry
constexpr int OZ_SLOTS = 8;                      // slice slots per K step in the layout (slot 7 unused)
constexpr int OZ_K = 256;                        // contraction length (a panel pair)
constexpr int OZ_KSTEPS = OZ_K / 32;             // MMA K = 32 bytes
constexpr int OZ_SLICE_STEP_BYTES = 128 * 32;    // one slice, one K step, 128 rows
constexpr int OZ_STAGE_OPERAND = OZ_SLOTS * OZ_SLICE_STEP_BYTES;    // 32 KB
constexpr int OZ_STAGE_BYTES = 2 * OZ_STAGE_OPERAND;                // A + B
constexpr int OZ_STAGES = 3;
constexpr int OZ_CSTG_ROW = 1088;                 // staging row pitch (128 doubles + 64 B: conflict-free 16-byte stores)
constexpr int OZ_CSTG_BYTES = 32 * OZ_CSTG_ROW;   // 32 rows of the C update staged for cp.reduce.async.bulk
constexpr int OZ_THREADS = 384;                   // warps 0..3: A producer, MMA issuer, B producer, idle; warps 4..11: epilogue
constexpr int OZ_REGS_CTRL = 56, OZ_REGS_EPI = 224;   // setmaxnreg: 128 x 56 + 256 x 224 = 64512 <= 65536
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
// non-blocking probe of a phase: the result comes back ~150 clk later, but nothing waits for it until it is used
__device__ __forceinline__ uint32_t oz_mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(oz_smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok;
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
// K-major, no swizzle: LBO = byte distance of the two 16-byte K chunks (128), SBO = byte distance of 8-row groups (256)
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr) {
    return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (static_cast<uint64_t>(128 >> 4) << 16) |
           (static_cast<uint64_t>(256 >> 4) << 32) | (static_cast<uint64_t>(1) << 46);
}
// instruction descriptor: D = s32, A = B = signed int8, both K-major, M = 128, N = 128
constexpr uint32_t OZ_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(128 >> 3) << 17) |
                              (static_cast<uint32_t>(128 >> 4) << 24);
// same with N = 256: the B descriptor then spans TWO adjacent digit slices (a slice slot is exactly 16 row groups x 256 B, so
// slice q+1 continues the row-group stride of slice q) and the result lands in two adjacent accumulators
constexpr uint32_t OZ_IDESC_N256 = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(256 >> 3) << 17) |
                                   (static_cast<uint32_t>(128 >> 4) << 24);
__device__ __forceinline__ void oz_mma_n256(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(OZ_IDESC_N256), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void oz_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(OZ_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void oz_mma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate);
// 16 lanes x 256 bit, 4 repetitions along the columns: 32 columns of 16 rows; thread t holds, for repetition j,
// v[4j+0..1] = row (t / 4), columns 8 j + 2 (t % 4) + {0, 1} and v[4j+2..3] = row (t / 4) + 8, same columns
__device__ __forceinline__ void oz_tmem_ld(uint32_t addr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr));
}

// tri > 0: lower-triangular tile set (+ full rows below); tri == 0: Mt x Nt rectangle, row index fastest (consecutive
// CTAs share the B slices)
__device__ __forceinline__ void oz_tile_decode(int t, int tri, int Mt, int& r, int& c) {
    if (tri == 0) {
        c = t / Mt;
        r = t - c * Mt;
        return;
    }
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
// one warp per row: s = 2^e with 4 |row|_max <= s (0 for an all-zero row)
__global__ void __launch_bounds__(256) ozaki_rowscale_kernel(const double* __restrict__ P, long ldp, int rows,
                                                             double* __restrict__ rscale) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    P += static_cast<long>(blockIdx.z) * OZ_K;                      // K panel blockIdx.z: columns 256 z .. of the same rows
    rscale += static_cast<long>(blockIdx.z) * rows;
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
        if (m > 0.0 && m < 1.0e300) s = scalbn(1.0, ilogb(m) + 3);
        rscale[row] = s;
    }
}

// thread = (row, 16-entry K chunk): 16 doubles in, 7 x 16 bytes out.
// t = rint(x 2^56 / s) (|t| < 2^54); adding 0x80 to each of its 7 low bytes makes every byte the balanced digit + 128
// with no borrows, so the digits are the bytes of (t + 0x00808080 80808080) XOR 0x80: byte j = digit 6 - j.
// rmaxq != nullptr: the row maxima come as four quarter-row values from the panel solves (trsm_rows_kernel) and the
// scale is formed here (the chunk-0 block also stores it for the update kernel); else rscale is an input.
__global__ void __launch_bounds__(128) ozaki_slice_kernel(const double* __restrict__ P, long ldp, double* __restrict__ rscale,
                                                          const double* __restrict__ rmaxq, int8_t* __restrict__ S) {
    const int rb = blockIdx.x, chunk = blockIdx.y, rr = threadIdx.x;
    const long row = static_cast<long>(rb) * 128 + rr;
    P += static_cast<long>(blockIdx.z) * OZ_K;                      // K panel blockIdx.z (launch_ozaki_slice_panels)
    rscale += static_cast<long>(blockIdx.z) * (static_cast<long>(gridDim.x) * 128);
    S += static_cast<long>(blockIdx.z) * (static_cast<long>(gridDim.x) * OZ_RB_BYTES);
    double sc;
    if (rmaxq != nullptr) {
        const double4 q = *reinterpret_cast<const double4*>(rmaxq + row * 4);
        const double m = fmax(fmax(q.x, q.y), fmax(q.z, q.w));
        sc = (m > 0.0 && m < 1.0e300) ? scalbn(1.0, ilogb(m) + 3) : 0.0;
        if (chunk == 0) rscale[row] = sc;
    } else {
        sc = rscale[row];
    }
    const double inv = sc > 0.0 ? 72057594037927936.0 / sc : 0.0;           // 2^56 / s
    // the 128 rows x 128 bytes of this block through shared memory (r02): a thread reading its own row directly issues 16-byte
    // loads 2 KB apart -- 32 cache lines per warp instruction, the L1 request rate was the limiter (ncu: L1/TEX 62 %, 2.6 TB/s);
    // staged, a warp instruction covers 4 full lines.  Row pitch 144 bytes: the 16-byte reads of 8 consecutive rows hit 32 distinct banks.
    __shared__ __align__(16) unsigned char tile[128 * 144];
    {
        const double* pb = P + static_cast<long>(rb) * 128 * ldp + chunk * 16;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = it * 16 + (rr >> 3), piece = rr & 7;
            const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(tile + r * 144 + piece * 16));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(pb + static_cast<long>(r) * ldp + piece * 2) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    const double* p = reinterpret_cast<const double*>(tile + rr * 144);
    uint32_t w[OZ_SLICES][4];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {                                         // 4 consecutive entries -> one word per slice
        uint32_t lo[4], hi[4];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const double2 v = *reinterpret_cast<const double2*>(p + 4 * q4 + 2 * h2);
            const unsigned long long u0 = static_cast<unsigned long long>(__double2ll_rn(v.x * inv)) + 0x0080808080808080ull;
            const unsigned long long u1 = static_cast<unsigned long long>(__double2ll_rn(v.y * inv)) + 0x0080808080808080ull;
            lo[2 * h2] = static_cast<uint32_t>(u0);
            hi[2 * h2] = static_cast<uint32_t>(u0 >> 32);
            lo[2 * h2 + 1] = static_cast<uint32_t>(u1);
            hi[2 * h2 + 1] = static_cast<uint32_t>(u1 >> 32);
        }
        // 4 x 4 byte transposes: word k of the result = byte k of the four entries
        const uint32_t l01a = __byte_perm(lo[0], lo[1], 0x5140), l01b = __byte_perm(lo[0], lo[1], 0x7362);
        const uint32_t l23a = __byte_perm(lo[2], lo[3], 0x5140), l23b = __byte_perm(lo[2], lo[3], 0x7362);
        const uint32_t h01a = __byte_perm(hi[0], hi[1], 0x5140), h01b = __byte_perm(hi[0], hi[1], 0x7362);
        const uint32_t h23a = __byte_perm(hi[2], hi[3], 0x5140), h23b = __byte_perm(hi[2], hi[3], 0x7362);
        w[6][q4] = __byte_perm(l01a, l23a, 0x5410) ^ 0x80808080u;           // byte 0 = digit 6
        w[5][q4] = __byte_perm(l01a, l23a, 0x7632) ^ 0x80808080u;
        w[4][q4] = __byte_perm(l01b, l23b, 0x5410) ^ 0x80808080u;
        w[3][q4] = __byte_perm(l01b, l23b, 0x7632) ^ 0x80808080u;
        w[2][q4] = __byte_perm(h01a, h23a, 0x5410) ^ 0x80808080u;           // byte 4 = digit 2
        w[1][q4] = __byte_perm(h01a, h23a, 0x7632) ^ 0x80808080u;
        w[0][q4] = __byte_perm(h01b, h23b, 0x5410) ^ 0x80808080u;           // byte 6 = digit 0 (most significant)
    }
    const int ks = chunk >> 1, kc = chunk & 1;
#pragma unroll
    for (int s = 0; s < OZ_SLICES; ++s) {
        int8_t* dst = S + static_cast<long>(rb) * OZ_RB_BYTES + (static_cast<long>(ks) * OZ_SLOTS + s) * OZ_SLICE_STEP_BYTES +
                      (rr >> 3) * 256 + kc * 128 + (rr & 7) * 16;
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
    const int8_t* SB;         // tri == 0 (rectangular Mt x Nt product C -= A B^T): slices / scales of the B rows
    const double* rscaleB;
    int Nt;
    long long* dbg;           // OZ_TIMING builds (tools/micro/ozaki_probe.cu): clock64 stamps of CTA 0
    int c_reduce;             // 1: C += c through shared memory + cp.reduce.async.bulk (no read of C); 0: load / add / store
    int prod3;                // 1: one producer thread per ring stage (both operands); 0: one per operand
    int kpanels;              // v5 only: > 1 = the tile set is repeated for `kpanels` K panels of 256 whose products all go to the SAME C
    long panel_bytes;         //   tiles (split-K through the L2 reduction): slices / scales of panel kp start at S + kp * panel_bytes,
    long panel_rows;          //   rscale + kp * panel_rows
    int add;                  // v5 only: 1 = C += A B^T instead of C -= A B^T
    int xp;                   // OZ_TIMING builds only (tools/micro/ozaki_probe.cu): timing experiments, bit mask --
                              // 1 half of the MMAs, 2 no B loads, 4 one-slice loads, 8 no read of C, 16 no store of C
};
#ifdef OZ_TIMING
#define OZ_STAMP(slot) do { if (blockIdx.x == 0 && g.dbg) g.dbg[slot] = clock64(); } while (0)
#else
#define OZ_STAMP(slot) do { } while (0)
#endif

struct __align__(8) OzBarriers {
    uint64_t full[OZ_STAGES], empty[OZ_STAGES], acc_full, acc_empty;
    uint32_t tmem_base, pad_;
};

// x2: 16 columns of 16 rows; thread t holds, for repetition j = 0, 1,
// v[4j+0..1] = row (t / 4), columns 8 j + 2 (t % 4) + {0, 1} and v[4j+2..3] = row (t / 4) + 8, same columns
__device__ __forceinline__ void oz_tmem_ld2(uint32_t addr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr));
}

__device__ __forceinline__ double2 oz_shfl_xor4(double2 v) {
    v.x = __shfl_xor_sync(0xffffffffu, v.x, 4);
    v.y = __shfl_xor_sync(0xffffffffu, v.y, 4);
    return v;
}

// exact int64 -> double for |t| < 2^51 without the slow 64-bit convert: bias into the mantissa of 2^52 + 2^51
__device__ __forceinline__ double oz_i64_to_double(long long t) {
    return __longlong_as_double(t + 0x4338000000000000LL) - 6755399441055744.0;
}

// Fold the NACC accumulators of a pass (weights w0 .. w0 + NACC - 1, accumulator g at columns 128 g) into the 64
// entries of C this thread owns: rows 32 quarter + 16 rh + r_in (+8), columns 64 chalf + 8 j + cq + {0, 1}.
// rsA / rsB point at the row scales of this thread's first row / column; wscale = 256^-(w_last + 2).
template <int NACC>
__device__ __forceinline__ void oz_fold8(const uint32_t (&a)[NACC][8], int jj, const double* __restrict__ rsB, double sr0,
                                         double sr1, double2 (&c0)[8], double2 (&c1)[8]) {
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
        const int j = 2 * jj + rep;
        const double sc0 = rsB[8 * j], sc1 = rsB[8 * j + 1];
        double d[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            long long t = static_cast<long long>(static_cast<int32_t>(a[0][4 * rep + e]));
#pragma unroll
            for (int gg = 1; gg < NACC; ++gg) t = t * 256LL + static_cast<long long>(static_cast<int32_t>(a[gg][4 * rep + e]));
            d[e] = oz_i64_to_double(t);
        }
        c0[j].x = fma(-d[0], sr0 * sc0, c0[j].x);
        c0[j].y = fma(-d[1], sr0 * sc1, c0[j].y);
        c1[j].x = fma(-d[2], sr1 * sc0, c1[j].x);
        c1[j].y = fma(-d[3], sr1 * sc1, c1[j].y);
    }
}
// The TMEM loads of step i + 1 are in flight while step i is folded (two register sets).
template <int NACC>
__device__ __forceinline__ void oz_drain(uint32_t tmem, int quarter, int chalf, const double* __restrict__ rsA,
                                         const double* __restrict__ rsB, double wscale, double2 (&c)[2][2][8]) {
    uint32_t a0[NACC][8], a1[NACC][8];
    const uint32_t tbase = tmem + (static_cast<uint32_t>(32 * quarter) << 16) + 64 * chalf;
    auto issue = [&](int step, uint32_t (&dst)[NACC][8]) {
        const uint32_t taddr = tbase + (static_cast<uint32_t>(16 * (step >> 2)) << 16) + 16 * (step & 3);
#pragma unroll
        for (int gg = 0; gg < NACC; ++gg) oz_tmem_ld2(taddr + gg * 128, dst[gg]);
    };
    issue(0, a0);
#pragma unroll
    for (int step = 0; step < 8; step += 2) {
        const int rh = step >> 2;
        const double sr0 = rsA[16 * rh] * wscale, sr1 = rsA[16 * rh + 8] * wscale;
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        issue(step + 1, a1);
        oz_fold8<NACC>(a0, step & 3, rsB, sr0, sr1, c[rh][0], c[rh][1]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (step + 2 < 8) issue(step + 2, a0);
        oz_fold8<NACC>(a1, (step + 1) & 3, rsB, sr0, sr1, c[rh][0], c[rh][1]);
    }
}

// All MMAs of one K step of a pass, fully unrolled: the issuing thread is a single thread, so every integer
// instruction between two tcgen05.mma counts against the 64-clock budget of an MMA (a run-time (p, q) loop with
// descriptor arithmetic issued one MMA per ~100 clocks).  The descriptors of the digit slices differ only in the
// start-address field (bits 0..13, units of 16 bytes).
template <int W0, int NW, int B_SLICE_BYTES, bool TWO_CTA>
__device__ __forceinline__ void oz_issue_kstep(uint32_t tmem, uint32_t sa, uint32_t sb, uint32_t not_first_ks, int half = 0);

// Persistent: CTA b works on tiles b, b + gridDim.x, ...  The stage ring and the accumulator hand-shake run across
// tiles, so the producer is already streaming the next tile while the epilogue folds the last pass of this one.
// An epilogue thread owns the same 64 entries of C in both passes: they are loaded while the MMAs of pass 0 run,
// updated in registers after each pass and stored once.
// NW0 = number of weights (= digits of each operand) of pass 0: 4 -> passes {0..3}, {4..6} (11 slice loads per operand and
// K step), 3 -> passes {0..2}, {3..6} (10 slice loads)
template <int NW0>
__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_syrk_kernel(const OzakiArgs g, const __grid_constant__ CUtensorMap cmap) {
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    unsigned char* cstg = oz_smem + OZ_STAGES * OZ_STAGE_BYTES;
    OzBarriers* bars = reinterpret_cast<OzBarriers*>(cstg + OZ_CSTG_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ntiles = g.tri > 0 ? g.tri * (g.tri + 1) / 2 + (g.Mt - g.tri) * g.tri : g.Mt * g.Nt;
    if (tid == 0) OZ_STAMP(0);

    if (tid == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) {
            oz_mbar_init(&bars->full[s], g.prod3 == 1 ? 1 : 2);
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
    if (tid == 0) OZ_STAMP(1);

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(OZ_REGS_CTRL));
    if (warp != 1 && g.prod3 != 1) {
        // ===== producers, one per operand: warp 0 streams the A slices, warp 2 the B slices =====
        if (lane == 0 && warp != 3) {
            const bool isB = warp == 2;
            uint32_t n = 0;                                  // global K-step counter
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                int tr, tc;
                oz_tile_decode(tile, g.tri, g.Mt, tr, tc);
                const int8_t* Sg = isB ? g.SB + static_cast<long>(tc) * OZ_RB_BYTES : g.S + static_cast<long>(tr) * OZ_RB_BYTES;
                for (int it = 0; it < 2 * OZ_KSTEPS; ++it, ++n) {
                    const uint32_t stage = n % OZ_STAGES, round = n / OZ_STAGES;
                    const int pass = it / OZ_KSTEPS, ks = it % OZ_KSTEPS;
                    uint32_t bytes = (pass == 0 ? NW0 : OZ_SLICES) * OZ_SLICE_STEP_BYTES;
#ifdef OZ_TIMING
                    if (g.xp & 4) bytes = OZ_SLICE_STEP_BYTES;
                    if ((g.xp & 2) && isB) bytes = 16;
#endif
                    oz_mbar_wait(&bars->empty[stage], (round & 1) ^ 1);
                    oz_mbar_expect_tx(&bars->full[stage], bytes);
                    oz_bulk_g2s(oz_smem + stage * OZ_STAGE_BYTES + (isB ? OZ_STAGE_OPERAND : 0),
                                Sg + static_cast<long>(ks) * OZ_STAGE_OPERAND, bytes, &bars->full[stage]);
                }
            }
        }
    } else if (warp != 1) {
        // ===== producers: warps 0, 2, 3 (lane 0).  Bulk copies issued by ONE thread are served one after the other
        // (600 clk each on an idle chip, 1000-1300 in the steady state of this kernel, whatever their size), so every
        // stage of the ring has its own producer thread, which loads both operands of its K steps: three K steps in
        // flight.  (Three lanes of one warp do NOT work: their mbarrier spin loops serialise inside the warp.) =====
        if (lane == 0) {
            const uint32_t stage = warp == 0 ? 0u : static_cast<uint32_t>(warp - 1);     // warps 0, 2, 3 -> stages 0, 1, 2
            const int my_tiles = ntiles > static_cast<int>(blockIdx.x) ? (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
            const uint32_t total = static_cast<uint32_t>(my_tiles) * 2 * OZ_KSTEPS;
            for (uint32_t n = stage; n < total; n += OZ_STAGES) {      // global K-step counter; n % OZ_STAGES == stage
                const int tile = static_cast<int>(blockIdx.x) + static_cast<int>(n / (2 * OZ_KSTEPS)) * static_cast<int>(gridDim.x);
                const int it = static_cast<int>(n % (2 * OZ_KSTEPS));
                int tr, tc;
                oz_tile_decode(tile, g.tri, g.Mt, tr, tc);
                const uint32_t round = n / OZ_STAGES;
                const int pass = it / OZ_KSTEPS, ks = it % OZ_KSTEPS;
                const uint32_t bytes = (pass == 0 ? NW0 : OZ_SLICES) * OZ_SLICE_STEP_BYTES;
                oz_mbar_wait(&bars->empty[stage], (round & 1) ^ 1);
                oz_mbar_expect_tx(&bars->full[stage], 2 * bytes);
                unsigned char* st = oz_smem + stage * OZ_STAGE_BYTES;
                oz_bulk_g2s(st, g.S + static_cast<long>(tr) * OZ_RB_BYTES + static_cast<long>(ks) * OZ_STAGE_OPERAND, bytes,
                            &bars->full[stage]);
                oz_bulk_g2s(st + OZ_STAGE_OPERAND, g.SB + static_cast<long>(tc) * OZ_RB_BYTES + static_cast<long>(ks) * OZ_STAGE_OPERAND,
                            bytes, &bars->full[stage]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            uint32_t n = 0, P = 0;                           // global K-step / pass counters
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int pass = 0; pass < 2; ++pass, ++P) {
                    if (P > 0) {                             // the accumulators are reused: wait for the drain of pass P-1
                        oz_mbar_wait(&bars->acc_empty, (P - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    for (int ks = 0; ks < OZ_KSTEPS; ++ks, ++n) {
                        const uint32_t stage = n % OZ_STAGES, round = n / OZ_STAGES;
                        oz_mbar_wait(&bars->full[stage], round & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (P == 0 && ks == 0) OZ_STAMP(2);
                        if (P == 1 && ks == 0) OZ_STAMP(3);
                        if (P == 2 && ks == 0) OZ_STAMP(15);
                        if (P == 2 && ks == 7) OZ_STAMP(16);
                        if (P == 3 && ks == 0) OZ_STAMP(17);
                        const uint32_t sa = oz_smem_u32(oz_smem + stage * OZ_STAGE_BYTES);
                        const uint32_t sb = sa + OZ_STAGE_OPERAND;
#ifdef OZ_TIMING
                        const int half = g.xp & 1;
#else
                        constexpr int half = 0;
#endif
                        if (pass == 0) oz_issue_kstep<0, NW0, OZ_SLICE_STEP_BYTES, false>(tmem, sa, sb, ks > 0 ? 1u : 0u, half);
                        else oz_issue_kstep<NW0, OZ_SLICES - NW0, OZ_SLICE_STEP_BYTES, false>(tmem, sa, sb, ks > 0 ? 1u : 0u, half);
                        oz_umma_commit(&bars->empty[stage]);   // frees the stage when these MMAs have read it
                        if (ks == OZ_KSTEPS - 1) oz_umma_commit(&bars->acc_full);
                    }
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(OZ_REGS_EPI));
        // ===== epilogue: 8 warps; lane quarter = warp % 4, column half = (warp - 4) / 4 =====
        const int quarter = warp & 3, chalf = (warp - 4) >> 2;
        const int r_in = lane >> 2, cq = 2 * (lane & 3);
        const bool odd = (r_in & 1) != 0;
        uint32_t P = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            int tr, tc;
            oz_tile_decode(tile, g.tri, g.Mt, tr, tc);
            // this thread's entries: rows 32 quarter + 16 rh + r_in (+8), columns 64 chalf + 8 j + cq + {0,1}
            double* Cb = g.C + (static_cast<long>(tr) * 128 + 32 * quarter + r_in) * g.ldc + static_cast<long>(tc) * 128 +
                         64 * chalf + cq;
            const double* rsA = g.rscale + static_cast<long>(tr) * 128 + 32 * quarter + r_in;
            const double* rsB = g.rscaleB + static_cast<long>(tc) * 128 + 64 * chalf + cq;
            double2 c[2][2][8];                              // -(P P^T) of this thread's entries, [rh][row r_in / +8][j]
#pragma unroll
            for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 8; ++j) c[rh][h][j] = make_double2(0.0, 0.0);
            // pull this thread's share of the C tile into L2 now (HBM -> L2 only, no traffic into the SM): the
            // read-modify-write after the second pass then sees L2 latency
            if (!g.c_reduce) {
#pragma unroll
                for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int j = 0; j < 8; j += 2)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(Cb + static_cast<long>(16 * rh + 8 * h) * g.ldc + 8 * j));
            }
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass, ++P) {
                // 256^-(last weight of the pass + 2)
                const double wscale = pass == 0 ? (NW0 == 4 ? 9.094947017729282e-13 /* 256^-5 */ : 2.3283064365386963e-10 /* 256^-4 */)
                                                : 5.421010862427522e-20 /* 256^-8 */;
                oz_mbar_wait(&bars->acc_full, P & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tid == 128 && tile == blockIdx.x) OZ_STAMP(4 + 2 * pass);
                if (tid == 128 && tile == blockIdx.x + gridDim.x) OZ_STAMP(10 + 2 * pass);
                if (pass == 0) oz_drain<NW0>(tmem, quarter, chalf, rsA, rsB, wscale, c);
                else oz_drain<OZ_SLICES - NW0>(tmem, quarter, chalf, rsA, rsB, wscale, c);
                // all TMEM reads of this pass are complete: hand the accumulators back
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) oz_mbar_arrive(&bars->acc_empty);
                if (tid == 128 && tile == blockIdx.x) OZ_STAMP(5 + 2 * pass);
                if (tid == 128 && tile == blockIdx.x + gridDim.x) OZ_STAMP(11 + 2 * pass);
#ifdef OZ_TIMING
                if (g.xp & 8) continue;
#endif
                if (pass == 0 && !g.c_reduce) {
                    // c = C - (pass-0 part): the tile is READ here, while the 144 MMAs of pass 1 run and these warps
                    // would only wait (pass 1 is bound by the shared-memory port, not by the L2 -> SM path); after the
                    // second pass only the store is left.  Read at the start of the tile it would compete with the stage
                    // refills of pass 0, read after pass 1 it lands on the next tile's pass 0 (measured: 12 k clk each).
                    // Two adjacent quads (rows r, r+1 of the fragment) team up so that ONE instruction covers 128 contiguous
                    // bytes of a row (4 full lines per warp instruction instead of 8 half lines): the even quad reads
                    // columns 8j.. of the even row and of the odd row, the odd quad columns 8(j+1).. of both; what belongs
                    // to the partner changes hands through shfl.xor 4.
#pragma unroll
                    for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const double* base = Cb + static_cast<long>(16 * rh + 8 * h) * g.ldc + (odd ? 8 : 0);
                            const double* row_e = base - (odd ? g.ldc : 0);
                            const double* row_o = row_e + g.ldc;
                            double2 ve[4], vo[4];
#pragma unroll
                            for (int jp = 0; jp < 4; ++jp) {
                                ve[jp] = *reinterpret_cast<const double2*>(row_e + 16 * jp);
                                vo[jp] = *reinterpret_cast<const double2*>(row_o + 16 * jp);
                            }
#pragma unroll
                            for (int jp = 0; jp < 4; ++jp) {
                                // even lane: ve = own (j), vo = partner's (j);  odd lane: vo = own (j+1), ve = partner's (j+1)
                                const double2 got = oz_shfl_xor4(odd ? ve[jp] : vo[jp]);
                                const double2 own = odd ? vo[jp] : ve[jp];
                                const double2 add0 = odd ? got : own, add1 = odd ? own : got;     // for columns 8j.. / 8(j+1)..
                                c[rh][h][2 * jp].x += add0.x;
                                c[rh][h][2 * jp].y += add0.y;
                                c[rh][h][2 * jp + 1].x += add1.x;
                                c[rh][h][2 * jp + 1].y += add1.y;
                            }
                        }
                }
            }
            if (g.c_reduce == 2) {
                // C += c without reading C, through the TMA: 32 rows x 128 columns at a time are staged densely in shared
                // memory (quad pairs exchange fragments so that an instruction writes 128-byte row segments) and leave
                // as ONE cp.reduce.async.bulk.tensor (.add, fp64 tensor map of C; SASS UTMAREDG): the adds happen in L2
#pragma unroll 1
                for (int qq = 0; qq < 4; ++qq) {
                    if (quarter == qq) {
#pragma unroll
                        for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                double* base = reinterpret_cast<double*>(cstg) + (16 * rh + 8 * h + r_in) * 128 + 64 * chalf + cq + (odd ? 8 : 0);
                                double* row_e = base - (odd ? 128 : 0);
                                double* row_o = row_e + 128;
#pragma unroll
                                for (int jp = 0; jp < 4; ++jp) {
                                    const double2 own = odd ? c[rh][h][2 * jp + 1] : c[rh][h][2 * jp];
                                    const double2 got = oz_shfl_xor4(odd ? c[rh][h][2 * jp] : c[rh][h][2 * jp + 1]);
                                    *reinterpret_cast<double2*>(row_e + 16 * jp) = odd ? got : own;
                                    *reinterpret_cast<double2*>(row_o + 16 * jp) = odd ? own : got;
                                }
                            }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (warp == 4 && lane == 0) {
                        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                         reinterpret_cast<uint64_t>(&cmap)),
                                     "r"(oz_smem_u32(cstg)), "r"(tc * 128), "r"(tr * 128 + 32 * qq)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            } else if (g.c_reduce) {
                // C += c without reading C: 32 rows at a time go through shared memory and leave as one
                // cp.reduce.async.bulk (.add.f64, SASS UBLKRED) per row -- the adds happen in L2, nothing comes back
#pragma unroll 1
                for (int qq = 0; qq < 4; ++qq) {
                    if (quarter == qq) {
#pragma unroll
                        for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                            for (int h = 0; h < 2; ++h)
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    *reinterpret_cast<double2*>(cstg + (16 * rh + 8 * h + r_in) * OZ_CSTG_ROW +
                                                                (64 * chalf + 8 * j + cq) * 8) = c[rh][h][j];
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (warp == 4) {
                        double* dst = g.C + (static_cast<long>(tr) * 128 + 32 * qq + lane) * g.ldc + static_cast<long>(tc) * 128;
                        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst),
                                     "r"(oz_smem_u32(cstg + lane * OZ_CSTG_ROW)), "r"(1024)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            } else {
#ifdef OZ_TIMING
                if ((g.prod3 == 7 || (g.xp & 16)) && c[0][0][0].x != 1.2345e300) goto skip_store;     // timing experiment: no C store
#endif
#pragma unroll
                for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double* base = Cb + static_cast<long>(16 * rh + 8 * h) * g.ldc + (odd ? 8 : 0);
                        double* row_e = base - (odd ? g.ldc : 0);
                        double* row_o = row_e + g.ldc;
#pragma unroll
                        for (int jp = 0; jp < 4; ++jp) {
                            const double2 own = odd ? c[rh][h][2 * jp + 1] : c[rh][h][2 * jp];
                            const double2 got = oz_shfl_xor4(odd ? c[rh][h][2 * jp] : c[rh][h][2 * jp + 1]);
                            *reinterpret_cast<double2*>(row_e + 16 * jp) = odd ? got : own;     // 128 B of the even row per quad pair
                            *reinterpret_cast<double2*>(row_o + 16 * jp) = odd ? own : got;     // 128 B of the odd row
                        }
                    }
#ifdef OZ_TIMING
            skip_store:;
#endif
            }
            if (tid == 128 && tile == blockIdx.x) OZ_STAMP(9);
            if (tid == 128 && tile == blockIdx.x + gridDim.x) OZ_STAMP(14);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) OZ_STAMP(8);
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

// =================================================================================================================
// Two-CTA variant (thread-block cluster of 2, tcgen05 cta_group::2).
// With both operands read from shared memory an M = 128, N = 128 int8 MMA needs 128 B/clk of shared-memory
// bandwidth -- all there is -- and the TMA refills of the stages come on top: the single-CTA kernel is bound by the
// shared-memory port (83 clk per MMA measured instead of 64).  A CTA pair works on a 256 x 128 tile: each CTA holds the
// slices of ITS 128 rows (A) and HALF of the 128 B rows, the pair-wide MMA (M = 256) reads A 4 KB + B 2 KB per CTA and
// instruction (96 B/clk) and every B slice crosses L2 -> SM once per pair instead of once per CTA.
//   rank 0 (leader): issues the MMAs for the pair; rank 1: its warp 1 relays "my stage has landed" to the leader.
//   stage free / accumulators ready: tcgen05.commit multicast to both CTAs; accumulators drained: both epilogues
//   arrive on the leader's barrier (remote arrive for rank 1).
// =================================================================================================================
constexpr int OZ2_STAGES = 4;
constexpr int OZ2_A_BYTES = OZ_STAGE_OPERAND;                  // 8 slices x 128 rows x 32 B
constexpr int OZ2_BH_SLICE = 64 * 32;                          // one slice of the 64-row half, one K step
constexpr int OZ2_BH_BYTES = OZ_SLOTS * OZ2_BH_SLICE;          // 16 KB
constexpr int OZ2_STAGE_BYTES = OZ2_A_BYTES + OZ2_BH_BYTES;    // 48 KB

struct __align__(8) Oz2Barriers {
    uint64_t full[OZ2_STAGES], peer_full[OZ2_STAGES], empty[OZ2_STAGES], acc_full, acc_empty;
    uint32_t tmem_base, pad_;
};
constexpr int OZ2_SMEM_BYTES = OZ2_STAGES * OZ2_STAGE_BYTES + static_cast<int>(sizeof(Oz2Barriers));

__device__ __forceinline__ uint32_t oz_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void oz_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t oz_mapa(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void oz_remote_arrive(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void oz_mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(oz_smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void oz_umma_commit2(uint64_t* bar) {      // arrive on the same barrier of BOTH CTAs
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            oz_smem_u32(bar)),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}
constexpr uint32_t OZ2_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(128 >> 3) << 17) |
                               (static_cast<uint32_t>(256 >> 4) << 24);
__device__ __forceinline__ void oz_mma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(OZ2_IDESC), "r"(accumulate)
        : "memory");
}

template <int W0, int NW, int B_SLICE_BYTES, bool TWO_CTA>
__device__ __forceinline__ void oz_issue_kstep(uint32_t tmem, uint32_t sa, uint32_t sb, uint32_t not_first_ks, int half) {
    const uint64_t da0 = oz_desc(sa), db0 = oz_desc(sb);
    if constexpr (TWO_CTA) {
#pragma unroll
        for (int gg = 0; gg < NW; ++gg) {
#pragma unroll
            for (int p = 0; p < OZ_SLICES; ++p) {
                const int q = W0 + gg - p;
                if (q < 0 || q >= OZ_SLICES) continue;
                const uint64_t da = da0 + static_cast<uint64_t>(p * (OZ_SLICE_STEP_BYTES >> 4));
                const uint64_t db = db0 + static_cast<uint64_t>(q * (B_SLICE_BYTES >> 4));
                oz_mma2(tmem + gg * 128, da, db, (p == 0) ? not_first_ks : 1u);   // p = 0 is the first pair of every weight
            }
        }
    } else {
        // digit p of A against every digit q of B whose weight p + q belongs to this pass; two consecutive q share ONE
        // N = 256 MMA (the A slice is read once for both, 12 KB instead of 16 KB of shared-memory reads per pair of
        // products): pass 0 = 4 x N256 + 2 x N128, pass 1 = 6 x N256 + 6 x N128.  p = 0 touches every accumulator first.
#pragma unroll
        for (int p = 0; p < OZ_SLICES; ++p) {
#ifdef OZ_TIMING
            if (half && (p & 1)) continue;
#endif
            const int q_lo = (W0 - p) > 0 ? (W0 - p) : 0;
            const int q_hi = (W0 + NW - 1 - p) < (OZ_SLICES - 1) ? (W0 + NW - 1 - p) : (OZ_SLICES - 1);
            const uint64_t da = da0 + static_cast<uint64_t>(p * (OZ_SLICE_STEP_BYTES >> 4));
            const uint32_t acc = (p == 0) ? not_first_ks : 1u;
#pragma unroll
            for (int q = q_lo; q <= q_hi; q += 2) {
                const uint64_t db = db0 + static_cast<uint64_t>(q * (B_SLICE_BYTES >> 4));
                const uint32_t dst = tmem + (p + q - W0) * 128;
                if (q + 1 <= q_hi) oz_mma_n256(dst, da, db, acc);
                else oz_mma(dst, da, db, acc);
            }
        }
    }
}

// work item w -> (pair row a, column block tc): pair a = row blocks 2a, 2a+1 and columns 0 .. min(2a+2, tri) - 1
__device__ __forceinline__ void oz2_decode(int w, int tri, int& a, int& tc) {
    a = 0;
    for (;;) {
        const int nc = (2 * a + 2 < tri) ? 2 * a + 2 : tri;
        if (w < nc) break;
        w -= nc;
        ++a;
    }
    tc = w;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OZ_THREADS, 1) ozaki_syrk2_kernel(const OzakiArgs g, int nwork) {
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    Oz2Barriers* bars = reinterpret_cast<Oz2Barriers*>(oz_smem + OZ2_STAGES * OZ2_STAGE_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = oz_cluster_rank();
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    if (tid == 0) OZ_STAMP(0);

    if (tid == 0) {
        for (int s = 0; s < OZ2_STAGES; ++s) {
            oz_mbar_init(&bars->full[s], 2);
            oz_mbar_init(&bars->peer_full[s], 1);
            oz_mbar_init(&bars->empty[s], 1);
        }
        oz_mbar_init(&bars->acc_full, 1);
        oz_mbar_init(&bars->acc_empty, 16);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&bars->tmem_base)),
                     "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    oz_cluster_sync();                                        // both CTAs: barriers initialised, TMEM allocated
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = bars->tmem_base;
    if (tid == 0) OZ_STAMP(1);

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(OZ_REGS_CTRL));
    if (warp == 0 || warp == 2) {
        // ===== producers (both CTAs): warp 0 = own A rows, warp 2 = own half of the B rows =====
        if (lane == 0) {
            const bool isB = warp == 2;
            uint32_t n = 0;
            for (int w = cluster_id; w < nwork; w += nclusters) {
                int a, tc;
                oz2_decode(w, g.tri, a, tc);
                int tr = 2 * a + static_cast<int>(rank);
                if (tr >= g.Mt) tr = g.Mt - 1;               // odd tile-row count: the partner reloads the last block
                const int8_t* Ag = g.S + static_cast<long>(tr) * OZ_RB_BYTES;
                const int8_t* Bg = g.S + static_cast<long>(tc) * OZ_RB_BYTES + rank * OZ2_BH_SLICE;
                for (int it = 0; it < 2 * OZ_KSTEPS; ++it, ++n) {
                    const uint32_t stage = n % OZ2_STAGES, round = n / OZ2_STAGES;
                    const int pass = it / OZ_KSTEPS, ks = it % OZ_KSTEPS;
                    const int nsl = pass == 0 ? 4 : OZ_SLICES;
                    oz_mbar_wait_cluster(&bars->empty[stage], (round & 1) ^ 1);
                    unsigned char* st = oz_smem + stage * OZ2_STAGE_BYTES;
                    if (!isB) {
                        oz_mbar_expect_tx(&bars->full[stage], nsl * OZ_SLICE_STEP_BYTES);
                        oz_bulk_g2s(st, Ag + static_cast<long>(ks) * OZ_STAGE_OPERAND, nsl * OZ_SLICE_STEP_BYTES, &bars->full[stage]);
                    } else {
                        oz_mbar_expect_tx(&bars->full[stage], nsl * OZ2_BH_SLICE);
                        const int8_t* bsrc = Bg + static_cast<long>(ks) * OZ_STAGE_OPERAND;
                        for (int c = 0; c < nsl; ++c)        // the 64-row half of slice c: 2 KB contiguous
                            oz_bulk_g2s(st + OZ2_A_BYTES + c * OZ2_BH_SLICE, bsrc + c * OZ_SLICE_STEP_BYTES, OZ2_BH_SLICE,
                                        &bars->full[stage]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 1) {
            // ===== relay: tell the leader that this CTA's stage has landed =====
            uint32_t n = 0;
            for (int w = cluster_id; w < nwork; w += nclusters)
                for (int it = 0; it < 2 * OZ_KSTEPS; ++it, ++n) {
                    const uint32_t stage = n % OZ2_STAGES, round = n / OZ2_STAGES;
                    oz_mbar_wait(&bars->full[stage], round & 1);
                    oz_remote_arrive(oz_mapa(oz_smem_u32(&bars->peer_full[stage]), 0));
                }
        } else if (lane == 0) {
            // ===== MMA issuer (leader) =====
            uint32_t n = 0, P = 0;
            for (int w = cluster_id; w < nwork; w += nclusters) {
                for (int pass = 0; pass < 2; ++pass, ++P) {
                    if (P > 0) {
                        oz_mbar_wait_cluster(&bars->acc_empty, (P - 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    for (int ks = 0; ks < OZ_KSTEPS; ++ks, ++n) {
                        const uint32_t stage = n % OZ2_STAGES, round = n / OZ2_STAGES;
                        oz_mbar_wait(&bars->full[stage], round & 1);
                        oz_mbar_wait_cluster(&bars->peer_full[stage], round & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (P == 0 && ks == 0) OZ_STAMP(2);
                        if (P == 1 && ks == 0) OZ_STAMP(3);
                        const uint32_t sa = oz_smem_u32(oz_smem + stage * OZ2_STAGE_BYTES);
                        const uint32_t sb = sa + OZ2_A_BYTES;
                        if (pass == 0) oz_issue_kstep<0, 4, OZ2_BH_SLICE, true>(tmem, sa, sb, ks > 0 ? 1u : 0u);
                        else oz_issue_kstep<4, 3, OZ2_BH_SLICE, true>(tmem, sa, sb, ks > 0 ? 1u : 0u);
                        oz_umma_commit2(&bars->empty[stage]);
                        if (ks == OZ_KSTEPS - 1) oz_umma_commit2(&bars->acc_full);
                    }
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(OZ_REGS_EPI));
        // ===== epilogue (both CTAs): rows of row block 2a + rank =====
        const int quarter = warp & 3, chalf = (warp - 4) >> 2;
        const int r_in = lane >> 2, cq = 2 * (lane & 3);
        const uint32_t acc_empty_leader = oz_mapa(oz_smem_u32(&bars->acc_empty), 0);
        uint32_t P = 0;
        for (int w = cluster_id; w < nwork; w += nclusters) {
            int a, tc;
            oz2_decode(w, g.tri, a, tc);
            const int tr = 2 * a + static_cast<int>(rank);
            const bool live = tr < g.Mt;
            const int trc = live ? tr : g.Mt - 1;
            double* Cb = g.C + (static_cast<long>(trc) * 128 + 32 * quarter + r_in) * g.ldc + static_cast<long>(tc) * 128 +
                         64 * chalf + cq;
            const double* rsA = g.rscale + static_cast<long>(trc) * 128 + 32 * quarter + r_in;
            const double* rsB = g.rscale + static_cast<long>(tc) * 128 + 64 * chalf + cq;
            double2 c[2][2][8];
#pragma unroll
            for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 8; ++j) c[rh][h][j] = make_double2(0.0, 0.0);
            // pull this thread's share of the C tile into L2 now (HBM -> L2 only, no traffic into the SM): the
            // read-modify-write after the second pass then sees L2 latency
#pragma unroll
            for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < 8; j += 2)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(Cb + static_cast<long>(16 * rh + 8 * h) * g.ldc + 8 * j));
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass, ++P) {
                const double wscale = pass == 0 ? 9.094947017729282e-13 /* 256^-5 */ : 5.421010862427522e-20 /* 256^-8 */;
                oz_mbar_wait_cluster(&bars->acc_full, P & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tid == 128 && w == cluster_id) OZ_STAMP(4 + 2 * pass);
                if (pass == 0) oz_drain<4>(tmem, quarter, chalf, rsA, rsB, wscale, c);
                else oz_drain<3>(tmem, quarter, chalf, rsA, rsB, wscale, c);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) oz_remote_arrive(acc_empty_leader);
                if (tid == 128 && w == cluster_id) OZ_STAMP(5 + 2 * pass);
            }
            if (live) {
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
                    double2 v[2][8];
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            v[h][j] = *reinterpret_cast<const double2*>(Cb + static_cast<long>(16 * rh + 8 * h) * g.ldc + 8 * j);
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<double2*>(Cb + static_cast<long>(16 * rh + 8 * h) * g.ldc + 8 * j) =
                                make_double2(v[h][j].x + c[rh][h][j].x, v[h][j].y + c[rh][h][j].y);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    oz_cluster_sync();                                        // nobody signals a departed CTA; all MMAs / reads done
    if (tid == 0) OZ_STAMP(8);
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}


// =================================================================================================================
// v5 (r02): four weight groups, two accumulator buffers -- the accumulator drains run UNDER the MMAs.
// What the r02 probes showed (profiles/r02/pattern_probe.txt, x1_*, x2_*):
//   * an N = 256 MMA (two adjacent B digits) runs at the nominal 128 clk, an N = 128 MMA takes 100 clk, not 64 (operand
//     fetch: 8 KB of shared memory per MMA); every mbarrier wait of the issuing thread costs ~130 clk that the tensor
//     pipe does NOT hide (the MMA queue is shallow), and so does anything else that slows that one thread down;
//   * operand loads are far from any limit (one slice instead of 4 / 7 per load: 3 % faster), L2 -> SM sustains 68 B/clk/SM
//     with all SMs pulling (bulk_probe), the two-pass kernels need ~25;
//   * in the two-pass kernels the tensor pipe idles through both drains (all 512 TMEM columns are live): 7.4 k of 31 k clk.
// Schedule: weights {5,6}, {3,4}, {1,2}, {0}, alternating between the TMEM halves X (columns 0..255) and Y (256..511):
// while the MMAs of a group fill one half, the epilogue drains the other.  In this order a group needs digits 0..6, 0..4,
// 0..2, 0 of both operands: 16 slice loads per operand and K step (10 in the two-pass form -- affordable, see above), and
// 19 of the 28 products still pair up into N = 256 MMAs (6 + 4 + 2 N256, 1 + 1 + 1 + 1 N128: 1936 clk per K step at the
// measured rates, the same as the two-pass form).  Ring: 3 stages x 56 KB (A half | B half) -- the rest of shared
// memory stages the C update, see the epilogue; a stage holds 1 / 1 / 2 / 7 K steps of group 0 / 1 / 2 / 3, so the issuing
// thread waits 8 + 8 + 4 + 2 = 22 times per tile.  (Measured and dropped: a byte-granular ring of 192 KB with 80 KB blocks for
// group 1 -- four of its K steps in flight instead of three, 18 waits -- was 5 % SLOWER, 0.184 vs 0.176 ms on 1830 tiles,
// profiles/r02/x7_v5_byte_ring.txt: the extra index arithmetic sits in the producer / issuer threads whose latency is what the
// ring exists to hide.)
// Epilogue: NO fp64 instruction at all (while int8 MMAs stream, DADD / DFMA / I2F.F64 of any warp of the SM run 7 - 130 x
// slower: profiles/r02/alu_probe.txt, drain_probe.txt).  The groups are accumulated into one int64 per entry, converted to
// -T 2^E with integer instructions (the row scales are powers of two: an exponent-field add), staged in shared memory and
// ADDED to C by the L2 through cp.reduce.async.bulk.tensor (.add, fp64 tensor map of C): C is never read by the SM.
// =================================================================================================================
constexpr int OZ5_STAGES = 3;
constexpr int OZ5_HALF = OZ_SLICES * OZ_SLICE_STEP_BYTES;                // 28 KB: one operand
constexpr int OZ5_STAGE_BYTES = 2 * OZ5_HALF;                            // 56 KB
constexpr int OZ5_GROUPS = 4;
constexpr int OZ5_PIECE_ROWS = 8;                                        // C leaves in pieces of 8 rows x 128 columns (8 KB) ...
constexpr int OZ5_PIECE_BYTES = OZ5_PIECE_ROWS * 128 * 8;                // ... one staging buffer per lane quarter (warp pair)
constexpr int OZ5_STAGING_BYTES = 4 * OZ5_PIECE_BYTES;                   // 32 KB

struct __align__(8) Oz5Barriers {
    uint64_t full[OZ5_STAGES], empty[OZ5_STAGES], acc_full[2], acc_empty[2];
    uint32_t tmem_base, pad_;
};
constexpr int OZ5_SMEM_BYTES = OZ5_STAGES * OZ5_STAGE_BYTES + OZ5_STAGING_BYTES + static_cast<int>(sizeof(Oz5Barriers));
static_assert(OZ5_SMEM_BYTES <= 232448, "227 KB of shared memory per CTA");

// group g: weights W0 .. W0 + NW - 1, digits 0 .. D - 1 of both operands, KPS K steps per ring stage
template <int G> struct Oz5Group;
template <> struct Oz5Group<0> { static constexpr int W0 = 5, NW = 2, D = 7, KPS = 1; };
template <> struct Oz5Group<1> { static constexpr int W0 = 3, NW = 2, D = 5, KPS = 1; };
template <> struct Oz5Group<2> { static constexpr int W0 = 1, NW = 2, D = 3, KPS = 2; };
template <> struct Oz5Group<3> { static constexpr int W0 = 0, NW = 1, D = 1, KPS = 7; };

// producer of ONE operand: all stage uses of group G of the current tile
template <int G>
__device__ __forceinline__ void oz5_produce(unsigned char* smem, Oz5Barriers* bars, const int8_t* Sg, int operand, uint32_t& n,
                                            int tiny) {
    using Gp = Oz5Group<G>;
#pragma unroll 1
    for (int ks0 = 0; ks0 < OZ_KSTEPS; ks0 += Gp::KPS, ++n) {
        const int nk = (OZ_KSTEPS - ks0) < Gp::KPS ? (OZ_KSTEPS - ks0) : Gp::KPS;
        const uint32_t stage = n % OZ5_STAGES, round = n / OZ5_STAGES;
        uint32_t bytes = Gp::D * OZ_SLICE_STEP_BYTES;
        if (tiny) bytes = 1024;
        unsigned char* dst = smem + stage * OZ5_STAGE_BYTES + operand * OZ5_HALF;
        oz_mbar_wait(&bars->empty[stage], (round & 1) ^ 1);
        oz_mbar_expect_tx(&bars->full[stage], static_cast<uint32_t>(nk) * bytes);
        for (int j = 0; j < nk; ++j)
            oz_bulk_g2s(dst + j * (Gp::D * OZ_SLICE_STEP_BYTES), Sg + static_cast<long>(ks0 + j) * OZ_STAGE_OPERAND, bytes,
                        &bars->full[stage]);
    }
}

// MMA issue of group G of the current tile into the accumulator buffer at tmem_buf.
// Every mbarrier wait of this thread costs ~130 clk that the tensor pipe does not hide (its queue is shallow: pattern_probe).  So
// the full barrier of the NEXT stage use is probed (test_wait, non-blocking) BEFORE the MMAs of this one are issued -- the probe's
// latency passes under their issue -- and the blocking wait is skipped when the probe has already seen the data (`ready`): with
// the producers two to three stages ahead that is the normal case.
template <int G>
__device__ __forceinline__ void oz5_issue(unsigned char* smem, Oz5Barriers* bars, uint32_t tmem_buf, uint32_t& n, uint32_t& ready,
                                          int half) {
    using Gp = Oz5Group<G>;
#pragma unroll 1
    for (int ks0 = 0; ks0 < OZ_KSTEPS; ks0 += Gp::KPS, ++n) {
        const int nk = (OZ_KSTEPS - ks0) < Gp::KPS ? (OZ_KSTEPS - ks0) : Gp::KPS;
        const uint32_t stage = n % OZ5_STAGES, round = n / OZ5_STAGES;
        const uint32_t sa = oz_smem_u32(smem + stage * OZ5_STAGE_BYTES);
        if (!ready) oz_mbar_wait(&bars->full[stage], round & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t next_ok = oz_mbar_test(&bars->full[(n + 1) % OZ5_STAGES], ((n + 1) / OZ5_STAGES) & 1);
#pragma unroll 1
        for (int j = 0; j < nk; ++j) {
            const uint32_t a = sa + j * (Gp::D * OZ_SLICE_STEP_BYTES);
            oz_issue_kstep<Gp::W0, Gp::NW, OZ_SLICE_STEP_BYTES, false>(tmem_buf, a, a + OZ5_HALF, (ks0 + j) > 0 ? 1u : 0u, half);
        }
        oz_umma_commit(&bars->empty[stage]);
        ready = next_ok;
    }
}

// Integer accumulation of the groups.  Measured (profiles/r02/alu_probe.txt, drain_probe.txt): while int8 MMAs stream, fp64
// instructions (DADD / DFMA) of ANY warp of the SM run 7 - 8 x slower, integer and fp32 instructions are unaffected -- a
// drain that folds in fp64 takes 68 k instead of 1.8 k clk under a saturated MMA stream.  So an entry is kept as ONE int64
//     T = a0 2^40 + a1 2^32 + a2 2^24 + a3 2^16 + a4 2^8 + a5 + round(a6 / 256)      (units of 256^-7 s_i s_j)
// (|a_w| <= (w + 1) 2^22: |T| < 2^63; dropping the low 8 bits of the last weight is 2^-57 s_i s_j, far below the digit pairs the
// scheme drops anyway), accumulated with IMAD.WIDE / IADD3 as the groups complete, and meets fp64 once per tile:
// one int64 -> double conversion, the power-of-two scales applied to its exponent field, one DADD with C.
// thread t holds, for repetition jr of a 16x256b.x4 load, v[4jr+0..1] = row (t/4), columns 8 jr + 2 (t%4) + {0,1}, v[4jr+2..3] = row + 8
template <int G, int NACC>
__device__ __forceinline__ void oz5_fold(const uint32_t (&a)[NACC][16], int j0, long long (&T0)[8][2], long long (&T1)[8][2]) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
        const int j = j0 + rep;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            long long& t = (e & 2) ? T1[j][e & 1] : T0[j][e & 1];
            const long long x0 = static_cast<long long>(static_cast<int32_t>(a[0][4 * rep + e]));
            if constexpr (G == 0) {          // weights 5, 6 -- the first group of a tile initialises T
                const int32_t x1 = static_cast<int32_t>(a[1][4 * rep + e]);
                t = x0 + static_cast<long long>((x1 + 128) >> 8);
            } else if constexpr (G == 1) {   // weights 3, 4
                const long long x1 = static_cast<long long>(static_cast<int32_t>(a[1][4 * rep + e]));
                t += x0 * 65536LL + x1 * 256LL;
            } else if constexpr (G == 2) {   // weights 1, 2
                const long long x1 = static_cast<long long>(static_cast<int32_t>(a[1][4 * rep + e]));
                t += x0 * 4294967296LL + x1 * 16777216LL;
            } else {                         // weight 0
                t += x0 * 1099511627776LL;
            }
        }
    }
}
// 4 steps of 16 rows x 32 columns (tcgen05.ld.16x256b.x4), the loads of step i + 1 in flight while step i is folded
template <int G, int NACC>
__device__ __forceinline__ void oz5_drain(uint32_t tmem_buf, int quarter, int chalf, long long (&T)[2][2][8][2]) {
    uint32_t a0[NACC][16], a1[NACC][16];
    const uint32_t tbase = tmem_buf + (static_cast<uint32_t>(32 * quarter) << 16) + 64 * chalf;
    auto issue = [&](int step, uint32_t (&dst)[NACC][16]) {
        const uint32_t taddr = tbase + (static_cast<uint32_t>(16 * (step >> 1)) << 16) + 32 * (step & 1);
#pragma unroll
        for (int gg = 0; gg < NACC; ++gg) oz_tmem_ld(taddr + gg * 128, dst[gg]);
    };
    issue(0, a0);
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        issue(2 * rh + 1, a1);
        oz5_fold<G, NACC>(a0, 0, T[rh][0], T[rh][1]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (rh == 0) issue(2, a0);
        oz5_fold<G, NACC>(a1, 4, T[rh][0], T[rh][1]);
    }
}
// exponent of a power-of-two scale as e << 20 (the position of the exponent field in the high word of a double), made safe for the
// exponent arithmetic: 0 (an all-zero row: its T is 0 whatever E) -> 2^0, below 2^-480 -> 2^-480 (such a row contributes < 2^-900 to C)
__device__ __forceinline__ int oz5_scale_hi(const double* p) {
    const int hi = __double2hiint(*p);
    return hi == 0 ? 0 : ((hi < ((1023 - 480) << 20) ? ((1023 - 480) << 20) : hi) - (1023 << 20));
}
// T 2^E, E = e_i + e_j - 56 (the unit of T is 256^-7 s_i s_j): (e_i << 20) + (e_j << 20) + OZ5_ECONST = E << 20 is ADDED to the high
// word of (double) T -- an exponent shift, exact; T == 0 stays 0
constexpr int OZ5_ECONST = -56 * 1048576;
// (double) t * 2^E with INTEGER instructions only (I2F.F64.S64 sits on the fp64 pipe too: 64 conversions per thread took 15 k clk
// under the MMA stream): normalise |t| (bit 63 set), keep 53 bits rounded half-up (1 ulp of 2^-53 relative against RN at worst,
// far below the scheme's own truncation), assemble sign | exponent | mantissa.  eshift = E << 20.
// Returns -t 2^E (the update is ADDED to C by the TMA reduction); sign_flip = 0x80000000 returns +t 2^E.
__device__ __forceinline__ double oz5_scaled(long long t, int eshift, int sign_flip = 0) {
    const unsigned long long a = t < 0 ? 0ULL - static_cast<unsigned long long>(t) : static_cast<unsigned long long>(t);
    const int lz = __clzll(static_cast<long long>(a | 1ULL));
    const unsigned long long nrm = a << lz;
    const unsigned long long mant = (nrm >> 11) + ((nrm >> 10) & 1ULL);           // 2^52 .. 2^53 (the implicit bit included)
    int hi = ((1085 - lz) << 20) + eshift + static_cast<int>(mant >> 32);        // (exponent - 1) << 20, + the implicit bit
    hi |= (~static_cast<int>(static_cast<unsigned long long>(t) >> 32) ^ sign_flip) & static_cast<int>(0x80000000u);
    const bool nz = t != 0;
    return __hiloint2double(nz ? hi : 0, nz ? static_cast<int>(static_cast<uint32_t>(mant)) : 0);
}

__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_syrk5_kernel(const OzakiArgs g, const __grid_constant__ CUtensorMap cmap) {
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    unsigned char* staging = oz_smem + OZ5_STAGES * OZ5_STAGE_BYTES;
    Oz5Barriers* bars = reinterpret_cast<Oz5Barriers*>(staging + OZ5_STAGING_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int base_tiles = g.tri > 0 ? g.tri * (g.tri + 1) / 2 + (g.Mt - g.tri) * g.tri : g.Mt * g.Nt;
    const int ntiles = base_tiles * (g.kpanels > 1 ? g.kpanels : 1);     // task t = (K panel t / base_tiles, tile t % base_tiles)
    if (tid == 0) OZ_STAMP(0);

    if (tid == 0) {
        for (int s = 0; s < OZ5_STAGES; ++s) {
            oz_mbar_init(&bars->full[s], 2);
            oz_mbar_init(&bars->empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            oz_mbar_init(&bars->acc_full[b], 1);
            oz_mbar_init(&bars->acc_empty[b], 8);
        }
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
    if (tid == 0) OZ_STAMP(1);
#ifdef OZ_TIMING
    const int xp = g.xp;
#else
    constexpr int xp = 0;
#endif

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(OZ_REGS_CTRL));
        if ((warp == 0 || warp == 2) && lane == 0) {
            // ===== producers: warp 0 = the A halves of the stages, warp 2 = the B halves =====
            const int operand = warp >> 1;
            uint32_t n = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                int tr, tc;
                const int kp = tile / base_tiles;
                oz_tile_decode(tile - kp * base_tiles, g.tri, g.Mt, tr, tc);
                const int8_t* Sg = (operand ? g.SB + static_cast<long>(tc) * OZ_RB_BYTES : g.S + static_cast<long>(tr) * OZ_RB_BYTES) +
                                   static_cast<long>(kp) * g.panel_bytes;
                oz5_produce<0>(oz_smem, bars, Sg, operand, n, xp & 4);
                oz5_produce<1>(oz_smem, bars, Sg, operand, n, xp & 4);
                oz5_produce<2>(oz_smem, bars, Sg, operand, n, xp & 4);
                oz5_produce<3>(oz_smem, bars, Sg, operand, n, xp & 4);
            }
        } else if (warp == 1 && lane == 0) {
            // ===== MMA issuer: groups 0, 2 -> buffer X, groups 1, 3 -> buffer Y; use u of a buffer waits for the drain of use u - 1 =====
            uint32_t n = 0, u = 0, ready = 0;                 // stage uses; uses of EACH buffer so far (both advance together)
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const bool t1 = tile == static_cast<int>(blockIdx.x + gridDim.x);
                if (u > 0) { oz_mbar_wait(&bars->acc_empty[0], (u - 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
                if (t1) OZ_STAMP(20);
                oz5_issue<0>(oz_smem, bars, tmem, n, ready, xp & 1);
                oz_umma_commit(&bars->acc_full[0]);
                if (u > 0) { oz_mbar_wait(&bars->acc_empty[1], (u - 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
                if (t1) OZ_STAMP(21);
                oz5_issue<1>(oz_smem, bars, tmem + 256, n, ready, xp & 1);
                oz_umma_commit(&bars->acc_full[1]);
                ++u;
                oz_mbar_wait(&bars->acc_empty[0], (u - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (t1) OZ_STAMP(22);
                oz5_issue<2>(oz_smem, bars, tmem, n, ready, xp & 1);
                oz_umma_commit(&bars->acc_full[0]);
                oz_mbar_wait(&bars->acc_empty[1], (u - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (t1) OZ_STAMP(23);
                oz5_issue<3>(oz_smem, bars, tmem + 256, n, ready, xp & 1);
                oz_umma_commit(&bars->acc_full[1]);
                ++u;
                if (t1) OZ_STAMP(24);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(OZ_REGS_EPI));
        // ===== epilogue: 8 warps; lane quarter = warp % 4, column half = (warp - 4) / 4 =====
        const int quarter = warp & 3, chalf = (warp - 4) >> 2;
        const int r_in = lane >> 2, cq = 2 * (lane & 3);
        const bool odd = (r_in & 1) != 0;
        uint32_t u = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const bool t0 = tile == static_cast<int>(blockIdx.x), t1 = tile == static_cast<int>(blockIdx.x + gridDim.x);
            int tr, tc;
            const int kp = tile / base_tiles;
            oz_tile_decode(tile - kp * base_tiles, g.tri, g.Mt, tr, tc);
            const int sign_flip = g.add ? static_cast<int>(0x80000000u) : 0;
            // this thread's entries: rows 32 quarter + 16 rh + r_in (+8), columns 64 chalf + 8 j + cq + {0,1}
            double* Cb = g.C + (static_cast<long>(tr) * 128 + 32 * quarter + r_in) * g.ldc + static_cast<long>(tc) * 128 +
                         64 * chalf + cq;
            const double* rsA = g.rscale + static_cast<long>(kp) * g.panel_rows + static_cast<long>(tr) * 128 + 32 * quarter + r_in;
            const double* rsB = g.rscaleB + static_cast<long>(kp) * g.panel_rows + static_cast<long>(tc) * 128 + 64 * chalf + cq;
            int hiA[2][2], hiB[8][2];
#pragma unroll
            for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                for (int h = 0; h < 2; ++h) hiA[rh][h] = oz5_scale_hi(rsA + 16 * rh + 8 * h);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                hiB[j][0] = oz5_scale_hi(rsB + 8 * j);
                hiB[j][1] = oz5_scale_hi(rsB + 8 * j + 1);
            }
            // (prefetching the C tile into L2 at this point was measured: 0.182 vs 0.178 ms -- the reductions do not wait for it)
            long long T[2][2][8][2];                         // this thread's entries, [rh][row r_in / +8][j][column parity]
#pragma unroll
            for (int grp = 0; grp < OZ5_GROUPS; ++grp) {
                const int b = grp & 1;
                if (grp == 2) ++u;
                oz_mbar_wait(&bars->acc_full[b], u & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tid == 128 && t1) OZ_STAMP(25 + 2 * grp);
                if (grp == 0) oz5_drain<0, 2>(tmem, quarter, chalf, T);
                else if (grp == 1) oz5_drain<1, 2>(tmem + 256, quarter, chalf, T);
                else if (grp == 2) oz5_drain<2, 2>(tmem, quarter, chalf, T);
                else oz5_drain<3, 1>(tmem + 256, quarter, chalf, T);
                // all TMEM reads of this group are complete: hand the buffer back
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) oz_mbar_arrive(&bars->acc_empty[b]);
                if (tid == 128 && t1) OZ_STAMP(26 + 2 * grp);
            }
            ++u;
            // C += -T 2^E WITHOUT the fp64 pipe and without reading C: the entries are converted with integer instructions,
            // staged in shared memory 8 rows x 128 columns at a time (one buffer per lane quarter = per pair of warps) and
            // leave as cp.reduce.async.bulk.tensor .add on an fp64 tensor map of C (SASS UTMAREDG): the adds happen in L2.
            // Two adjacent quads (rows r, r+1 of the fragment) team up so that 8 lanes write 128 contiguous bytes of one
            // row (no bank conflicts): the even quad writes columns 8j.. of the even row and of the odd row, the odd quad
            // columns 8(j+1).. of both; what belongs to the partner changes hands through shfl.xor 4.
            {
                double* stg = reinterpret_cast<double*>(staging + quarter * OZ5_PIECE_BYTES);
                double* st_e = stg + (r_in - (odd ? 1 : 0)) * 128 + 64 * chalf + cq + (odd ? 8 : 0);
                double* st_o = st_e + 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rh = q >> 1, h = q & 1;
                    const int er = hiA[rh][h] + OZ5_ECONST;
                    // convert first (registers only): this overlaps the TMA's read of the previous piece of this warp pair
                    double2 se[4], so[4];
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) {
                        double2 c0, c1;                                               // columns 8 (2 jp).. / 8 (2 jp + 1).. of this thread's row
                        c0.x = oz5_scaled(T[rh][h][2 * jp][0], er + hiB[2 * jp][0], sign_flip);
                        c0.y = oz5_scaled(T[rh][h][2 * jp][1], er + hiB[2 * jp][1], sign_flip);
                        c1.x = oz5_scaled(T[rh][h][2 * jp + 1][0], er + hiB[2 * jp + 1][0], sign_flip);
                        c1.y = oz5_scaled(T[rh][h][2 * jp + 1][1], er + hiB[2 * jp + 1][1], sign_flip);
                        const double2 mine = odd ? c1 : c0;
                        const double2 back = oz_shfl_xor4(odd ? c0 : c1);
                        se[jp] = odd ? back : mine;                                   // 128 B of the even row per quad pair
                        so[jp] = odd ? mine : back;                                   // 128 B of the odd row
                    }
                    // the previous piece has been read by the TMA before anybody overwrites the buffer
                    if (chalf == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) {
                        *reinterpret_cast<double2*>(st_e + 16 * jp) = se[jp];
                        *reinterpret_cast<double2*>(st_o + 16 * jp) = so[jp];
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + quarter) : "memory");
                    if (chalf == 0 && lane == 0 && !(xp & 16)) {
                        if (xp & 32) {      // timing experiment (probe builds): a plain TMA store instead of the reduction
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                             reinterpret_cast<uint64_t>(&cmap)),
                                         "r"(oz_smem_u32(stg)), "r"(tc * 128), "r"(tr * 128 + 32 * quarter + 16 * rh + 8 * h)
                                         : "memory");
                        } else {
                            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                                             reinterpret_cast<uint64_t>(&cmap)),
                                         "r"(oz_smem_u32(stg)), "r"(tc * 128), "r"(tr * 128 + 32 * quarter + 16 * rh + 8 * h)
                                         : "memory");
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            if (tid == 128 && t0) OZ_STAMP(9);
            if (tid == 128 && t1) OZ_STAMP(14);
        }
        // the reductions of this thread have been performed (not only read from shared memory) before the CTA retires
        if (chalf == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) OZ_STAMP(8);
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

constexpr int OZ_SMEM_BYTES = OZ_STAGES * OZ_STAGE_BYTES + OZ_CSTG_BYTES + static_cast<int>(sizeof(OzBarriers));

}  // namespace

// bytes of slice storage / scale storage for a panel of `rows` rows (multiple of 128)
size_t ozaki_slice_bytes(long rows) { return static_cast<size_t>(rows / 128) * OZ_RB_BYTES; }

// P: rows x 256 (ldp), rows a multiple of 128 -> rscale[rows], S
void launch_ozaki_slice(const double* P, long ldp, int rows, double* rscale, int8_t* S, cudaStream_t s,
                        const double* rmaxq) {
    if (rows <= 0) return;
    if (rmaxq == nullptr) ozaki_rowscale_kernel<<<(rows + 7) / 8, 256, 0, s>>>(P, ldp, rows, rscale);
    ozaki_slice_kernel<<<dim3(rows / 128, OZ_K / 16), 128, 0, s>>>(P, ldp, rscale, rmaxq, S);
}

// P: rows x (256 kpanels) (ldp): every K panel of 256 columns is sliced on its own (own row scales):
// rscale[kp][rows], S[kp][rows / 128][256 KB]
void launch_ozaki_slice_panels(const double* P, long ldp, int rows, int kpanels, double* rscale, int8_t* S, cudaStream_t s) {
    if (rows <= 0 || kpanels <= 0) return;
    ozaki_rowscale_kernel<<<dim3((rows + 7) / 8, 1, kpanels), 256, 0, s>>>(P, ldp, rows, rscale);
    ozaki_slice_kernel<<<dim3(rows / 128, OZ_K / 16, kpanels), 128, 0, s>>>(P, ldp, rscale, nullptr, S);
}

// C -= A B^T from the slices of A (SA, rsA) and of B (SB, rsB).  tri > 0: SYRK form (B = A), lower-triangular tile set
// with Mt - tri full tile rows below; tri == 0: rectangular Mt x Nt.
static void ozaki_launch(double* C, long ldc, const int8_t* SA, const double* rsA, const int8_t* SB, const double* rsB, int Mt,
                         int Nt, int tri, cudaStream_t s, long long* dbg, int persist_hint, int kpanels = 1, long panel_bytes = 0,
                         long panel_rows = 0, int add = 0) {
    static bool configured_dev[64] = {false};
    int dev_ = 0;
    cudaGetDevice(&dev_);
    bool& configured = configured_dev[dev_ & 63];
    if (!configured) {
        cudaFuncSetAttribute(ozaki_syrk_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES);
        cudaFuncSetAttribute(ozaki_syrk_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES);
        cudaFuncSetAttribute(ozaki_syrk2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ2_SMEM_BYTES);
        cudaFuncSetAttribute(ozaki_syrk5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ5_SMEM_BYTES);
        configured = true;
    }
    static int sm_count_dev[64] = {0};
    int& sms = sm_count_dev[dev_ & 63];
    if (sms == 0) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev_);
    const int tiles = (tri > 0 ? tri * (tri + 1) / 2 + (Mt - tri) * tri : Mt * Nt) * (kpanels > 1 ? kpanels : 1);
    if (tiles <= 0) return;
    static const int c_reduce = getenv("EGX_OZAKI_CRED") != nullptr ? atoi(getenv("EGX_OZAKI_CRED")) : 0;   // measured: same tile rate as load / add / store (the 128 row reductions of a tile serialise in the TMA unit)
    static const int prod3 = getenv("EGX_OZAKI_PROD3") != nullptr ? atoi(getenv("EGX_OZAKI_PROD3")) : 0;
    static const int xp = getenv("EGX_OZAKI_XP") != nullptr ? atoi(getenv("EGX_OZAKI_XP")) : 0;           // probe builds only
    static const int nw0 = getenv("EGX_OZAKI_NW0") != nullptr ? atoi(getenv("EGX_OZAKI_NW0")) : 4;
    OzakiArgs g{C, ldc, SA, rsA, Mt, tri, SB, rsB, Nt, dbg, c_reduce, prod3, kpanels, panel_bytes, panel_rows, add, xp};
    alignas(64) CUtensorMap cmap;
    memset(&cmap, 0, sizeof(cmap));
    static const int version = getenv("EGX_OZAKI_V") != nullptr ? atoi(getenv("EGX_OZAKI_V")) : 5;     // 3: the r01 kernel (3-stage ring, two passes)
    bool have_map = false;
    if (c_reduce == 2 || version == 5) {
        // fp64 tensor map of the C region of this launch: (columns, rows), row pitch ldc, boxes of 128 columns x 32 (v5: 8) rows
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static encode_fn encode = nullptr;
        if (encode == nullptr) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess) encode = reinterpret_cast<encode_fn>(fn);
        }
        const cuuint64_t dims[2] = {static_cast<cuuint64_t>(Nt) * 128, static_cast<cuuint64_t>(Mt) * 128};
        const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ldc) * sizeof(double)};
        const cuuint32_t box[2] = {128, static_cast<cuuint32_t>(version == 5 ? OZ5_PIECE_ROWS : 32)}, estr[2] = {1, 1};
        if (encode == nullptr ||
            encode(&cmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            g.c_reduce = 0;                        // no encoder / unsupported shape: plain load / add / store (the r01 kernel)
        } else {
            have_map = true;
        }
    }
    // persistent: a resident grid loops over the tiles (prefetch across tiles, no per-tile set-up, the C update of a
    // tile overlaps the MMAs of the next) -- but it keeps the high-priority panel / look-ahead kernels of the SAME
    // factorisation waiting for SMs.  The sweep asks for it when several evaluations are in flight (the batched entry
    // point: other evaluations fill the gaps; 4.39 -> 4.05 ms per evaluation at n = 8192) and not for a single one
    // (6.9 vs 7.7 ms).  EGX_OZAKI_PERSIST=0/1 overrides.
    static const int persist_env = getenv("EGX_OZAKI_PERSIST") != nullptr ? atoi(getenv("EGX_OZAKI_PERSIST")) : -1;
    // v5 overlaps the drain / C update of a tile with the MMAs of the next, so a resident grid is 30 % faster than one tile per
    // CTA (0.170 vs 0.241 ms on 1830 tiles) and -- unlike the r01 kernel -- costs a single evaluation with look-ahead nothing
    // (6.78 ms either way, profiles/r02/x6_*): always persistent
    const int persist = persist_env >= 0 ? persist_env : (version == 5 ? 1 : persist_hint);
    static const int two_cta = getenv("EGX_OZAKI_2CTA") != nullptr ? atoi(getenv("EGX_OZAKI_2CTA")) : 0;
    if (two_cta && tri > 0) {
        int nwork = 0;
        for (int a = 0; 2 * a < Mt; ++a) nwork += (2 * a + 2 < tri) ? 2 * a + 2 : tri;
        const int clusters = (persist && nwork > sms / 2) ? sms / 2 : nwork;
        ozaki_syrk2_kernel<<<2 * clusters, OZ_THREADS, OZ2_SMEM_BYTES, s>>>(g, nwork);
        return;
    }
    // persistent grid: as few CTAs as finish in the same number of tile rounds (230 tiles: 115 CTAs x 2 rounds instead of
    // 148 CTAs of which 66 idle through the second round) -- the SMs left over go to the other evaluations in flight
    // EGX_OZAKI_MAXCTAS caps the resident grid: with several evaluations in flight the SMs left over run the panel / solve /
    // correlation kernels of the OTHER evaluations instead of queueing behind this launch
    static const int max_ctas = getenv("EGX_OZAKI_MAXCTAS") != nullptr ? atoi(getenv("EGX_OZAKI_MAXCTAS")) : 0;
    // measured with 8 evaluations in flight at n = 8192 (profiles/r02/x8_batch_sweep.txt): 148 / 132 / 120 / 104 CTAs -> 3.32 / 3.28 /
    // 3.23 / 3.28 ms per evaluation: the batched entry point (persist_hint) leaves 24 SMs to the other evaluations' small kernels
    const int avail = (max_ctas > 0 && max_ctas < sms) ? max_ctas : ((persist_hint && max_ctas == 0 && sms > 48) ? sms - 24 : sms);
    int grid = tiles;
    if (persist && tiles > avail) {
        const int rounds = (tiles + avail - 1) / avail;
        grid = (tiles + rounds - 1) / rounds;
    }
    if (version == 5 && have_map) ozaki_syrk5_kernel<<<grid, OZ_THREADS, OZ5_SMEM_BYTES, s>>>(g, cmap);
    else if (nw0 == 3) ozaki_syrk_kernel<3><<<grid, OZ_THREADS, OZ_SMEM_BYTES, s>>>(g, cmap);
    else ozaki_syrk_kernel<4><<<grid, OZ_THREADS, OZ_SMEM_BYTES, s>>>(g, cmap);
}

void launch_ozaki_syrk(double* C, long ldc, const int8_t* S, const double* rscale, int Mt, int tri, cudaStream_t s,
                       long long* dbg, int persist_hint) {
    ozaki_launch(C, ldc, S, rscale, S, rscale, Mt, tri, tri, s, dbg, persist_hint);
}

// rectangular: C (Mt x Nt tiles) -= A B^T ; the multi-RHS triangular solve of predict_var (A = solved rows of the point
// chunk, B = block rows of L, sliced once per model)
void launch_ozaki_gemm(double* C, long ldc, const int8_t* SA, const double* rsA, const int8_t* SB, const double* rsB, int Mt,
                       int Nt, cudaStream_t s, int persist_hint) {
    ozaki_launch(C, ldc, SA, rsA, SB, rsB, Mt, Nt, 0, s, nullptr, persist_hint);
}

// C (lower tile triangle, tri x tri tiles) += sum over `kpanels` K panels of 256 of W_kp W_kp^T, from the panel-wise slices of
// launch_ozaki_slice_panels: ONE launch of tri (tri + 1) / 2 x kpanels tile tasks whose results meet in L2 (the TMA reduction is
// atomic per element).  The split-K SYRK of the sparse GP (A = I + V diag(beta) V^T, gp/src/sparse_algorithm.rs:727-731, 798).
// Returns false when the v5 kernel is not in use (EGX_OZAKI_V=3) -- the caller keeps its DMMA path.
bool launch_ozaki_syrk_add_panels(double* C, long ldc, const int8_t* S, const double* rscale, int tri, int kpanels, cudaStream_t s) {
    static const int version = getenv("EGX_OZAKI_V") != nullptr ? atoi(getenv("EGX_OZAKI_V")) : 5;
    if (version != 5) return false;
    ozaki_launch(C, ldc, S, rscale, S, rscale, tri, tri, tri, s, nullptr, 1, kpanels, static_cast<long>(ozaki_slice_bytes(tri * 128L)),
                 tri * 128L, 1);
    return true;
}
