// Stand-alone probe of the int8-sliced tcgen05 SYRK update (egobox_b200/csrc/kernels_ozaki.cu):
//  (a) register layout of tcgen05.ld.16x256b.x4, (b) C -= P P^T against a long-double host loop on a small
//  tile set, (c) time of a trailing-update-sized launch.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/micro/ozaki_probe tools/micro/ozaki_probe.cu
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#define OZ_TIMING 1
void egx_set_error(const char*, ...) {}
#include "../../egobox_b200/csrc/kernels_ozaki.cu"

__global__ void layout_probe(uint32_t* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&slot)), "n"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    // lane L of warp w owns TMEM lane 32 w + L: write value = 1000 * tmem_lane + column for 32 columns
    for (int c = 0; c < 32; ++c) {
        const uint32_t v = 1000u * (32 * warp + lane) + c;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + (static_cast<uint32_t>(32 * warp) << 16) + c), "r"(v));
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int rh = 0; rh < 2; ++rh) {
        uint32_t v[16];
        oz_tmem_ld(tmem + (static_cast<uint32_t>(32 * warp + 16 * rh) << 16), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[((warp * 2 + rh) * 32 + lane) * 16 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32));
}

static int check_layout() {
    uint32_t* d;
    cudaMalloc(&d, 4 * 2 * 32 * 16 * 4);
    layout_probe<<<1, 128>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        printf("layout probe: CUDA error %s\n", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    std::vector<uint32_t> h(4 * 2 * 32 * 16);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int w = 0; w < 4; ++w)
        for (int rh = 0; rh < 2; ++rh)
            for (int t = 0; t < 32; ++t)
                for (int j = 0; j < 4; ++j)
                    for (int e = 0; e < 4; ++e) {
                        const int row = 32 * w + 16 * rh + t / 4 + (e >= 2 ? 8 : 0);
                        const int col = 8 * j + 2 * (t % 4) + (e & 1);
                        const uint32_t want = 1000u * row + col, got = h[((w * 2 + rh) * 32 + t) * 16 + 4 * j + e];
                        if (want != got) {
                            if (bad < 8) printf("  layout: warp %d half %d thread %d reg %d: got lane %u col %u, assumed lane %d col %d\n",
                                                w, rh, t, 4 * j + e, got / 1000, got % 1000, row, col);
                            ++bad;
                        }
                    }
    printf("tcgen05.ld.16x256b.x4 layout mismatches: %ld\n", bad);
    return bad != 0;
}

static int check_update(int Mt, int tri, bool timing_only, int reps) {
    const int rows = Mt * 128, cols = tri * 128;
    const long ldp = 256, ldc = cols;
    std::mt19937_64 rng(7 + Mt);
    std::uniform_real_distribution<double> U(-1.0, 1.0), E(0.0, 6.0);
    std::vector<double> P(static_cast<size_t>(rows) * ldp), C(static_cast<size_t>(rows) * ldc);
    for (auto& v : P) v = U(rng) * std::pow(10.0, -E(rng));
    for (auto& v : C) v = U(rng);
    double *dP, *dC, *dR;
    int8_t* dS;
    cudaMalloc(&dP, P.size() * 8);
    cudaMalloc(&dC, C.size() * 8);
    cudaMalloc(&dR, rows * 8);
    cudaMalloc(&dS, ozaki_slice_bytes(rows));
    cudaMemcpy(dP, P.data(), P.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1, e2;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&e2);
    cudaEventRecord(e0);
    launch_ozaki_slice(dP, ldp, rows, dR, dS, 0);
    cudaEventRecord(e1);
    long long* dbg;
    cudaMalloc(&dbg, 48 * 8);
    cudaMemset(dbg, 0, 48 * 8);
    for (int r = 0; r < reps; ++r) launch_ozaki_syrk(dC, ldc, dS, dR, Mt, tri, 0, dbg, 0);
    cudaEventRecord(e2);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        printf("update (Mt=%d tri=%d): CUDA error %s\n", Mt, tri, cudaGetErrorString(err));
        return 1;
    }
    float ms_slice = 0, ms_upd = 0;
    cudaEventElapsedTime(&ms_slice, e0, e1);
    cudaEventElapsedTime(&ms_upd, e1, e2);
    const long tiles = static_cast<long>(tri) * (tri + 1) / 2 + static_cast<long>(Mt - tri) * tri;
    printf("Mt=%d tri=%d: %ld tiles, slice %.3f ms, update %.3f ms per launch -> %.1f fp64-equivalent TFLOP/s\n", Mt, tri, tiles,
           ms_slice, ms_upd / reps, 2.0 * 128 * 128 * 256 * tiles / (ms_upd / reps * 1e-3) / 1e12);
    {
        long long h[48];
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        if (h[20] != 0) {        // v5 stamps (second tile of CTA 0)
            printf("  v5 CTA 0: first tile done %lld, second tile done %lld (tile period %lld clk), CTA end %lld\n", h[9] - h[0], h[14] - h[0],
                   h[14] - h[9], h[8] - h[0]);
            printf("    second tile, MMA thread: group starts %lld %lld %lld %lld, all issued %lld\n", h[20] - h[0], h[21] - h[0], h[22] - h[0],
                   h[23] - h[0], h[24] - h[0]);
            printf("    second tile, epilogue: drains [%lld %lld] [%lld %lld] [%lld %lld] [%lld %lld]\n", h[25] - h[0], h[26] - h[0], h[27] - h[0],
                   h[28] - h[0], h[29] - h[0], h[30] - h[0], h[31] - h[0], h[32] - h[0]);
        } else
        printf("  CTA 0 clocks from start: setup %lld, first stage landed %lld, pass-1 first stage %lld | epilogue: acc0 ready %lld, "
               "drain0 done %lld, acc1 ready %lld, drain1 done %lld | CTA end %lld\n", h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0],
               h[5] - h[0], h[6] - h[0], h[7] - h[0], h[8] - h[0]);
        printf("    first tile C update done %lld | second tile: acc0 ready %lld, drain0 done %lld, acc1 ready %lld, drain1 done %lld, C update done %lld\n",
               h[9] - h[0], h[10] - h[0], h[11] - h[0], h[12] - h[0], h[13] - h[0], h[14] - h[0]);
        printf("    MMA warp, second tile: pass-0 first stage ready %lld, last K step issued %lld, pass-1 first stage %lld\n", h[15] - h[0],
               h[16] - h[0], h[17] - h[0]);
    }
    if (timing_only) return 0;
    std::vector<double> Cg(C.size());
    cudaMemcpy(Cg.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
    double worst = 0.0, worst64 = 0.0;
    for (int tr = 0; tr < Mt; ++tr)
        for (int tc = 0; tc < tri; ++tc) {
            if (tr < tri && tc > tr) continue;
            for (int i = 0; i < 128; ++i)
                for (int j = 0; j < 128; ++j) {
                    const double* a = &P[static_cast<size_t>(tr * 128 + i) * ldp];
                    const double* b = &P[static_cast<size_t>(tc * 128 + j) * ldp];
                    long double s = 0.0L;
                    double s64 = 0.0, amax = 0.0, bmax = 0.0;
                    for (int k = 0; k < 256; ++k) {
                        s += static_cast<long double>(a[k]) * b[k];
                        s64 = std::fma(a[k], b[k], s64);
                        amax = std::fmax(amax, std::fabs(a[k]));
                        bmax = std::fmax(bmax, std::fabs(b[k]));
                    }
                    const size_t idx = static_cast<size_t>(tr * 128 + i) * ldc + tc * 128 + j;
                    const long double want = static_cast<long double>(C[idx]) - s;
                    const double den = amax * bmax * 256.0 + 1e-300;
                    worst = std::fmax(worst, static_cast<double>(fabsl(Cg[idx] - want)) / den);
                    worst64 = std::fmax(worst64, static_cast<double>(fabsl((C[idx] - s64) - want)) / den);
                }
        }
    printf("  max |err| / (K |a|max |b|max): sliced tcgen05 %.3e, plain fp64 fma loop %.3e (2^-53 = 1.1e-16)\n", worst, worst64);
    return worst < 5e-15 ? 0 : 1;
}


// ---- issue-pattern probe: the MMA sequences of the update kernel with NO loads (operands = whatever sits in shared
// memory), to separate the tensor-pipe rate of the instruction mix from everything else.
//   mode 0: pass-0 pattern {0,1,2}; 1: pass-1 pattern {3..6}; 2: only N = 256 MMAs, rotating over 2 accumulator pairs;
//   3: only N = 128 MMAs rotating over 4 accumulators; 4: N = 128 MMAs into ONE accumulator; 5: N = 256 into one pair
//   commit_each: tcgen05.commit to a dummy mbarrier after every K step; vary_addr: K step ks reads slot ks % 4
__global__ void __launch_bounds__(128, 1) pattern_kernel(int mode, int ksteps, int commit_each, int vary_addr, int wait_each, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(oz_smem + 8 * OZ5_HALF - 256);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 10; ++i) oz_mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(tslot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    if (tid == 32) {
        const long long t0 = clock64();
        uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int ks = 0; ks < ksteps; ++ks) {
            const int slot = vary_addr ? (ks & 3) : 0;
            const uint32_t sa = oz_smem_u32(oz_smem + (2 * slot) * OZ5_HALF), sb = sa + OZ5_HALF;
            const uint32_t nf = ks > 0 ? 1u : 0u;
            if (wait_each && ks >= wait_each) {          // wait for the commit of K step ks - wait_each (ring of depth wait_each)
                const int b = (ks - wait_each) & 7;
                oz_mbar_wait(&bars[b], ph[b] & 1);
                ph[b]++;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            if (mode == 0) oz_issue_kstep<0, 3, OZ_SLICE_STEP_BYTES, false>(tmem, sa, sb, nf);
            else if (mode == 1) oz_issue_kstep<3, 4, OZ_SLICE_STEP_BYTES, false>(tmem, sa, sb, nf);
            else if (mode == 2) {
                const uint64_t da = oz_desc(sa), db = oz_desc(sb);
#pragma unroll
                for (int i = 0; i < 8; ++i) oz_mma_n256(tmem + (i & 1) * 256, da + i * 256, db + (i & 1) * 512, nf);
            } else if (mode == 3) {
                const uint64_t da = oz_desc(sa), db = oz_desc(sb);
#pragma unroll
                for (int i = 0; i < 8; ++i) oz_mma(tmem + (i & 3) * 128, da + i * 256, db + (i & 3) * 256, nf);
            } else if (mode == 4) {
                const uint64_t da = oz_desc(sa), db = oz_desc(sb);
#pragma unroll
                for (int i = 0; i < 8; ++i) oz_mma(tmem, da + i * 256, db + (i & 3) * 256, nf);
            } else {
                const uint64_t da = oz_desc(sa), db = oz_desc(sb);
#pragma unroll
                for (int i = 0; i < 8; ++i) oz_mma_n256(tmem, da + i * 256, db + (i & 1) * 512, nf);
            }
            if (commit_each) oz_umma_commit(&bars[ks & 7]);
        }
        oz_umma_commit(&bars[8]);
        oz_mbar_wait(&bars[8], 0);
        cyc[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

static void pattern_probe() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(pattern_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * OZ5_HALF);
    long long* d;
    cudaMalloc(&d, 256 * 8);
    long long h[256];
    const int ideal[6] = {2 * 128 + 2 * 64, 10 * 128 + 2 * 64, 8 * 128, 8 * 64, 8 * 64, 8 * 128};
    const int ninstr[6] = {4, 12, 8, 8, 8, 8};
    for (int grid : {1, sms})
        for (int mode = 0; mode < 6; ++mode)
            for (int cfg = 0; cfg < 5; ++cfg) {
                const int commit_each = cfg >= 1, vary = cfg >= 2, wait_each = cfg == 3 ? 4 : (cfg == 4 ? 2 : 0);
                const int ksteps = 256;
                for (int rep = 0; rep < 2; ++rep) pattern_kernel<<<grid, 128, 8 * OZ5_HALF>>>(mode, ksteps, commit_each, vary, wait_each, d);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("pattern: error %s\n", cudaGetErrorString(cudaGetLastError())); return; }
                cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("pattern grid=%3d mode=%d (%2d MMAs / K step, tensor-pipe floor %4d clk) commit_each=%d vary_addr=%d wait_depth=%d : %.0f clk / K step\n",
                       grid, mode, ninstr[mode], ideal[mode], commit_each, vary, wait_each, double(mx) / ksteps);
            }
}

// ---- TMEM read probe: bytes per clock of tcgen05.ld for a few shapes, with the tensor pipe idle and with a stream of
// N = 256 int8 MMAs accumulating into the OTHER half of tensor memory.  8 reader warps (warps 4..11, as the epilogue of the
// update kernel): warp w reads lanes 32 (w % 4) .., columns 128 (w >= 8) .. + 127 of buffer X = 64 KB per pass over all warps... x2.
template <int SHAPE>
__device__ __forceinline__ uint32_t ldtm_pass(uint32_t tbase) {
    uint32_t acc = 0;
    if (SHAPE == 0) {            // 16x256b.x4: 16 lanes x 32 columns, 16 registers
#pragma unroll
        for (int lh = 0; lh < 2; ++lh)
#pragma unroll
            for (int cb = 0; cb < 4; ++cb) {
                uint32_t v[16];
                oz_tmem_ld(tbase + (static_cast<uint32_t>(16 * lh) << 16) + 32 * cb, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; ++i) acc ^= v[i];
            }
    } else if (SHAPE == 1) {     // 32x32b.x32: 32 lanes x 32 columns, 32 registers
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            uint32_t v[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
                "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(tbase + 32 * cb));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) acc ^= v[i];
        }
    } else if (SHAPE == 2) {     // 16x256b.x4, two loads in flight before the wait
#pragma unroll
        for (int lh = 0; lh < 2; ++lh)
#pragma unroll
            for (int cb = 0; cb < 4; cb += 2) {
                uint32_t v[16], w[16];
                oz_tmem_ld(tbase + (static_cast<uint32_t>(16 * lh) << 16) + 32 * cb, v);
                oz_tmem_ld(tbase + (static_cast<uint32_t>(16 * lh) << 16) + 32 * cb + 32, w);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; ++i) acc ^= v[i] ^ w[i];
            }
    } else {                     // 32x32b.x16, two in flight
#pragma unroll
        for (int cb = 0; cb < 8; cb += 2) {
            uint32_t v[16], w[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(tbase + 16 * cb));
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]),
                  "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                : "r"(tbase + 16 * cb + 16));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; ++i) acc ^= v[i] ^ w[i];
        }
    }
    return acc;
}

// mma_mode: 0 none, 1 N = 256 MMAs into columns 256..511 (the other half), 2 N = 128 MMAs into columns 256..383
__global__ void __launch_bounds__(384, 1) ldtm_kernel(int shape, int passes, int mma_mode, int nreaders, long long* cyc, uint32_t* sink) {
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(oz_smem + 4 * OZ5_HALF);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
    volatile int* stop = reinterpret_cast<volatile int*>(tslot + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        oz_mbar_init(&bars[0], 1);
        oz_mbar_init(&bars[1], 1);
        oz_mbar_init(&bars[2], 1);
        *stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(tslot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    if (tid == 32 && mma_mode) {
        const uint64_t da = oz_desc(oz_smem_u32(oz_smem)), db = oz_desc(oz_smem_u32(oz_smem + OZ5_HALF));
        uint32_t n = 0;
        if (mma_mode <= 2) {
            while (!*stop) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (mma_mode == 1) oz_mma_n256(tmem + 256, da + (i & 3) * 256, db + (i & 1) * 512, 1u);
                    else oz_mma(tmem + 256, da + (i & 3) * 256, db + (i & 3) * 256, 1u);
                }
                ++n;
            }
            oz_umma_commit(&bars[0]);
            oz_mbar_wait(&bars[0], 0);
            cyc[1] = n;
        } else {
            // throttled: groups of G = mma_mode - 2 N = 256 MMAs, each followed by a commit; group k + 1 is issued only when
            // group k - 1 has completed, so at most 2 G MMAs are ever outstanding and the thread SLEEPS on an mbarrier instead
            // of sitting on a full MMA queue
            const int G = mma_mode - 2;
            uint32_t k = 0;
            while (!*stop) {
                for (int i = 0; i < G; ++i) oz_mma_n256(tmem + 256, da + (i & 3) * 256, db + (i & 1) * 512, 1u);
                oz_umma_commit(&bars[1 + (k & 1)]);
                if (k >= 1) oz_mbar_wait(&bars[1 + ((k - 1) & 1)], ((k - 1) >> 1) & 1);
                ++k;
            }
            oz_mbar_wait(&bars[1 + ((k - 1) & 1)], ((k - 1) >> 1) & 1);
            cyc[1] = (static_cast<long long>(k) * G) / 8;        // in units of 8 MMAs, like the free-running modes
        }
    }
    if (warp >= 4 && warp < 4 + nreaders) {
        const uint32_t tbase = tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + 128 * ((warp - 4) >> 2);
        uint32_t acc = 0;
        if (shape >= 4) {            // the real drains of the update kernels on whatever tensor memory holds (|values| < 2^25 not guaranteed: timing only)
            const int quarter = warp & 3, chalf = (warp - 4) >> 2;
            int hiA[2][2] = {{1023 << 20, 1022 << 20}, {1021 << 20, 1020 << 20}}, hiB[8][2];
            for (int j = 0; j < 8; ++j) hiB[j][0] = hiB[j][1] = (1023 - j) << 20;
            double2 c[2][2][8];
            long long TT[2][2][8][2];
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int j = 0; j < 8; ++j) {
                c[a][b][j] = make_double2(1.0 * tid, 2.0 * j);
                TT[a][b][j][0] = tid + j;
                TT[a][b][j][1] = tid - j;
            }
            __shared__ double rs[256];
            rs[tid & 255] = 1.0;
            asm volatile("bar.sync 1, %0;" ::"r"(32 * nreaders) : "memory");
            const long long t0 = clock64();
            for (int it = 0; it < passes; ++it) {
                if (shape == 4) oz5_drain<1, 2>(tmem, quarter, chalf, TT);
                else if (shape == 6) oz5_drain<3, 1>(tmem, quarter, chalf, TT);
                else oz_drain<4>(tmem, quarter, chalf, rs, rs + 128, 1e-10, c);
            }
            asm volatile("bar.sync 1, %0;" ::"r"(32 * nreaders) : "memory");
            const long long t1 = clock64();
            if (tid == 128) {
                cyc[0] = t1 - t0;
                *stop = 1;
            }
            double sum = 0.0;
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) for (int j = 0; j < 8; ++j) sum += c[a][b][j].x + c[a][b][j].y + static_cast<double>(TT[a][b][j][0] ^ TT[a][b][j][1]);
            if (sum == 1.2345) sink[tid] = 1;
            goto done;
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * nreaders) : "memory");
        const long long t0 = clock64();
        for (int it = 0; it < passes; ++it) {
            if (shape == 0) acc ^= ldtm_pass<0>(tbase);
            else if (shape == 1) acc ^= ldtm_pass<1>(tbase);
            else if (shape == 2) acc ^= ldtm_pass<2>(tbase);
            else acc ^= ldtm_pass<3>(tbase);
        }
        asm volatile("bar.sync 1, %0;" ::"r"(32 * nreaders) : "memory");
        const long long t1 = clock64();
        if (tid == 128) {
            cyc[0] = t1 - t0;
            *stop = 1;
        }
        if (acc == 0x12345678u) sink[tid] = acc;
    }
done:
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

static void drain_probe() {
    cudaFuncSetAttribute(ldtm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * OZ5_HALF + 256);
    long long* d;
    uint32_t* sink;
    cudaMalloc(&d, 64);
    cudaMalloc(&sink, 4096);
    const char* names[3] = {"v5 drain, 2 accumulators (int64 accumulate)", "v3/v4 drain, 4 accumulators (fp64 scaling)", "v5 drain, 1 accumulator"};
    const int naccs[3] = {2, 4, 1};
    for (int mma_mode = 0; mma_mode < 7; ++mma_mode)
        for (int shape = 4; shape < 7; ++shape) {
            const int passes = 64;
            cudaMemset(d, 0, 64);
            for (int rep = 0; rep < 2; ++rep) ldtm_kernel<<<1, 384, 4 * OZ5_HALF + 256>>>(shape, passes, mma_mode, 8, d, sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("drain: error %s\n", cudaGetErrorString(cudaGetLastError())); return; }
            long long h[2];
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const char* mm[7] = {"idle", "N=256 free-running", "N=128 free-running", "N=256, <= 2 outstanding", "N=256, <= 4 outstanding",
                                 "N=256, <= 6 outstanding", "N=256, <= 8 outstanding"};
            printf("drain: %-44s, MMA %-26s: %6.0f clk per drain = %5.0f clk per accumulator", names[shape - 4], mm[mma_mode],
                   double(h[0]) / passes, double(h[0]) / passes / naccs[shape - 4]);
            if (mma_mode) printf("   (MMA stream at %.0f %% of its own rate)", 100.0 * double(h[1]) * 8 * (mma_mode == 2 ? 100.0 : 128.0) / double(h[0]));
            printf("\n");
        }
}

static void ldtm_probe() {
    cudaFuncSetAttribute(ldtm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * OZ5_HALF + 256);
    long long* d;
    uint32_t* sink;
    cudaMalloc(&d, 64);
    cudaMalloc(&sink, 4096);
    const char* names[4] = {"16x256b.x4 (1 in flight)", "32x32b.x32 (1 in flight)", "16x256b.x4 (2 in flight)", "32x32b.x16 (2 in flight)"};
    for (int nreaders : {8, 4})
        for (int mma_mode = 0; mma_mode < 3; ++mma_mode)
            for (int shape = 0; shape < 4; ++shape) {
                const int passes = 256;
                cudaMemset(d, 0, 64);
                for (int rep = 0; rep < 2; ++rep) ldtm_kernel<<<1, 384, 4 * OZ5_HALF + 256>>>(shape, passes, mma_mode, nreaders, d, sink);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("ldtm: error %s\n", cudaGetErrorString(cudaGetLastError())); return; }
                long long h[2];
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                const double bytes = double(passes) * nreaders * 32 * 128 * 4;       // each reader warp: 32 lanes x 128 columns per pass
                const double mma_clk = mma_mode == 1 ? 128.0 : 100.0;
                printf("tmem read: %d reader warps, %-26s, MMA %-28s: %6.1f B/clk", nreaders, names[shape],
                       mma_mode == 0 ? "idle" : (mma_mode == 1 ? "N=256 stream (other half)" : "N=128 stream (other half)"), bytes / double(h[0]));
                if (mma_mode) printf("   (MMA stream at %.0f %% of its own rate)", 100.0 * double(h[1]) * 8 * mma_clk / double(h[0]));
                printf("\n");
            }
}

// ---- which instruction types of the epilogue does a running MMA stream slow down?  8 warps run 4096 independent-ish
// instructions of one type each (8 chains per thread); one thread of warp 1 free-runs N = 256 MMAs.
__global__ void __launch_bounds__(384, 1) alu_kernel(int kind, int mma_on, long long* cyc, double* sink) {
    extern __shared__ __align__(1024) unsigned char oz_smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(oz_smem + 4 * OZ5_HALF);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bars + 4);
    volatile int* stop = reinterpret_cast<volatile int*>(tslot + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        oz_mbar_init(&bars[0], 1);
        *stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(tslot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    if (tid == 32 && mma_on) {
        const uint64_t da = oz_desc(oz_smem_u32(oz_smem)), db = oz_desc(oz_smem_u32(oz_smem + OZ5_HALF));
        while (!*stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i) oz_mma_n256(tmem + 256, da + (i & 3) * 256, db + (i & 1) * 512, 1u);
        }
        oz_umma_commit(&bars[0]);
        oz_mbar_wait(&bars[0], 0);
    }
    if (warp >= 4) {
        long long ia[8];
        double da_[8];
        int ib[8];
        float fa[8];
        for (int i = 0; i < 8; ++i) { ia[i] = tid * 7 + i; da_[i] = 1.0 + 1e-9 * (tid + i); ib[i] = tid + 3 * i; fa[i] = 1.0f + 1e-3f * i; }
        const int mul = tid | 1;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const long long t0 = clock64();
        for (int it = 0; it < 512; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (kind == 0) ib[i] = (ib[i] ^ mul) & (ib[(i + 1) & 7] | it);                 // LOP3
                else if (kind == 1) ib[i] = ib[i] + ib[(i + 1) & 7] + it;                       // IADD3
                else if (kind == 2) ib[i] = ib[i] * mul + it;                                   // IMAD
                else if (kind == 3) ia[i] = static_cast<long long>(ib[i]) * mul + ia[i];        // IMAD.WIDE
                else if (kind == 4) da_[i] = da_[i] + 1.000001;                                 // DADD
                else if (kind == 5) da_[i] = fma(da_[i], 1.0000001, 1e-7);                      // DFMA
                else if (kind == 6) fa[i] = fmaf(fa[i], 1.0001f, 1e-3f);                        // FFMA
                else ib[i] = __funnelshift_r(ib[i], ib[(i + 1) & 7], 7) + 1;                    // SHF + IADD
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const long long t1 = clock64();
        if (tid == 128) {
            cyc[0] = t1 - t0;
            *stop = 1;
        }
        double sum = 0.0;
        for (int i = 0; i < 8; ++i) sum += static_cast<double>(ia[i]) + da_[i] + ib[i] + fa[i];
        if (sum == 1.2345) sink[tid] = sum;
        // per-warp time of the warps sharing the issuer's scheduler (warp % 4 == 1) against the others
        if ((tid & 31) == 0) cyc[2 + warp] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

static void alu_probe() {
    cudaFuncSetAttribute(alu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * OZ5_HALF + 256);
    long long* d;
    double* sink;
    cudaMalloc(&d, 16 * 8);
    cudaMalloc(&sink, 4096 * 8);
    const char* names[8] = {"LOP3", "IADD3", "IMAD", "IMAD.WIDE", "DADD", "DFMA", "FFMA", "SHF+IADD"};
    for (int kind = 0; kind < 8; ++kind) {
        double base = 0;
        for (int mma_on = 0; mma_on < 2; ++mma_on) {
            cudaMemset(d, 0, 16 * 8);
            for (int rep = 0; rep < 2; ++rep) alu_kernel<<<1, 384, 4 * OZ5_HALF + 256>>>(kind, mma_on, d, sink);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("alu: error %s\n", cudaGetErrorString(cudaGetLastError())); return; }
            long long h[16];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            if (!mma_on) base = double(h[0]);
            printf("alu: %-10s MMA %-12s: %8lld clk for 4096 instr / thread (x %.2f)", names[kind], mma_on ? "free-running" : "idle", h[0],
                   double(h[0]) / base);
            printf("   per warp 4..11:");
            for (int w = 4; w < 12; ++w) printf(" %lld", h[2 + w]);
            printf("\n");
        }
    }
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "alu") == 0) { alu_probe(); return 0; }
    if (argc > 1 && strcmp(argv[1], "ldtm") == 0) { ldtm_probe(); return 0; }
    if (argc > 1 && strcmp(argv[1], "drain") == 0) { drain_probe(); return 0; }
    if (argc > 1 && strcmp(argv[1], "pattern") == 0) { pattern_probe(); return 0; }
    if (argc > 1 && strcmp(argv[1], "time") == 0) {        // timing only; variants come through EGX_OZAKI_* (one process each)
        check_update(60, 60, true, 3);            // warm-up (clocks, first-launch effects)
        check_update(60, 60, true, 40);
        return 0;
    }
    int rc = check_layout();
    rc |= check_update(3, 2, false, 1);
    rc |= check_update(6, 4, false, 1);
    check_update(60, 60, true, 5);
    check_update(40, 30, true, 5);
    return rc;
}
