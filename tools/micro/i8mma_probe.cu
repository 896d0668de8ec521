// Micro-probe for tcgen05.mma kind::i8 on sm_100a: (1) correctness of one 128 x N x 32 int8 MMA with both
// operands in shared memory (K-major, no swizzle, core matrices of 8 rows x 16 bytes) against a host loop;
// (2) issue-to-completion cycles of a chain of such MMAs (N = 128 / 256), i.e. the int8 rate of one SM.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/i8mma_probe tools/micro/i8mma_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;          // descriptor version (Blackwell)
    return d;                                     // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
    } while (!ok);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// element (r, kb) of an R-row operand tile with K = 32 bytes: two K chunks of 16 bytes, chunk stride R*16
__host__ __device__ inline int tile_off(int r, int kb, int R) { return (kb >> 4) * R * 16 + r * 16 + (kb & 15); }

template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const int8_t* A, const int8_t* B, int32_t* D, int reps, long long* cycles, int nacc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    int8_t* sA = reinterpret_cast<int8_t*>(smem);              // 128 x 32
    int8_t* sB = sA + 128 * 32;                                // N x 32
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + N * 32);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int e = tid; e < 128 * 32; e += 128) sA[e] = A[e];    // inputs are already in tile order
    for (int e = tid; e < N * 32; e += 128) sB[e] = B[e];
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (tensor core)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    const uint64_t da = make_desc(smem_u32(sA), 128 * 16, 128);
    const uint64_t db = make_desc(smem_u32(sB), N * 16, 128);
    const uint32_t idesc = make_idesc(128, N);
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r) mma_i8(tmem + (r % nacc) * N, da, db, idesc, r >= nacc ? 1u : 0u);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    if (tid == 0) {
        t1 = clock64();
        cycles[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // accumulator 0: row = TMEM lane = 32 * warp + lane, N columns
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t addr = tmem + (static_cast<uint32_t>(32 * warp) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) D[(32 * warp + lane) * N + c0 + j] = static_cast<int32_t>(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

template <int N>
int run(int reps_time) {
    std::vector<int8_t> a(128 * 32), b(N * 32), at(128 * 32), bt(N * 32);
    srand(1234 + N);
    for (auto& v : a) v = static_cast<int8_t>(rand() % 255 - 127);
    for (auto& v : b) v = static_cast<int8_t>(rand() % 255 - 127);
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 32; ++k) at[tile_off(r, k, 128)] = a[r * 32 + k];
    for (int r = 0; r < N; ++r)
        for (int k = 0; k < 32; ++k) bt[tile_off(r, k, N)] = b[r * 32 + k];
    int8_t *dA, *dB;
    int32_t* dD;
    long long* dC;
    cudaMalloc(&dA, at.size());
    cudaMalloc(&dB, bt.size());
    cudaMalloc(&dD, 128 * N * 4);
    cudaMalloc(&dC, 8);
    cudaMemcpy(dA, at.data(), at.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, bt.data(), bt.size(), cudaMemcpyHostToDevice);
    const int smem = 128 * 32 + N * 32 + 64;
    // (1) correctness: reps = 1 -> D = A B^T
    probe_kernel<N><<<1, 128, smem>>>(dA, dB, dD, 1, dC, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e));
        return 1;
    }
    std::vector<int32_t> d(128 * N);
    cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < N; ++j) {
            int32_t s = 0;
            for (int k = 0; k < 32; ++k) s += static_cast<int32_t>(a[i * 32 + k]) * b[j * 32 + k];
            if (s != d[i * N + j]) {
                if (bad < 5) printf("  mismatch (%d,%d): got %d want %d\n", i, j, d[i * N + j], s);
                ++bad;
            }
        }
    printf("N=%d: single MMA mismatches = %ld of %d\n", N, bad, 128 * N);
    // (2) timing: a chain of MMAs alternating between two accumulators
    for (int nacc : {1, 2, 512 / N})
    for (int reps : {reps_time}) {
        probe_kernel<N><<<1, 128, smem>>>(dA, dB, dD, reps, dC, nacc);
        cudaDeviceSynchronize();
        long long cyc = 0;
        cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
        printf("N=%d: %d MMAs (128x%dx32 int8) round-robin over %d accumulator(s) in %lld cycles = %.1f cycles/MMA, %.0f MAC/cycle/SM\n", N, reps, N, nacc, cyc,
               double(cyc) / reps, 128.0 * N * 32 * reps / double(cyc));
    }
    return bad != 0;
}

int main() {
    int rc = run<128>(2048);
    rc |= run<256>(2048);
    return rc;
}
