// How fast can ONE SM pull L2-resident data?  (a) 1-D TMA bulk copies (cp.async.bulk) through a ring of smem slots,
// (b) plain 16-byte loads by 512 threads.  Reported in bytes per clock per SM, for 1 CTA and for one CTA on every SM.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/bulk_probe tools/micro/bulk_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// `nprod` producer threads (lane 0 of warps 0..nprod-1), each with its own ring of `slots` slots of `sz` bytes and its
// own source region; reports the clocks of producer 0
__global__ void __launch_bounds__(128) bulk_kernel(const unsigned char* src, long span, int sz, int slots, int ncopies,
                                                   int nprod, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char sm[];
    const int pid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0 && pid < nprod) {
        unsigned char* ring = sm + static_cast<long>(pid) * slots * sz;
        uint64_t* bars = reinterpret_cast<uint64_t*>(sm + static_cast<long>(nprod) * slots * sz) + pid * slots;
        for (int i = 0; i < slots; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const long myspan = span / nprod;
        const unsigned char* base = src + static_cast<long>(blockIdx.x) * span + static_cast<long>(pid) * myspan;
        const long long t0 = clock64();
        for (int n = 0; n < ncopies + slots; ++n) {
            const int slot = n % slots;
            if (n >= slots) mbar_wait(&bars[slot], ((n / slots) - 1) & 1);
            if (n < ncopies) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[slot])), "r"(sz) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 s32(ring + static_cast<long>(slot) * sz)),
                             "l"(base + (static_cast<long>(n) * sz) % myspan), "r"(sz), "r"(s32(&bars[slot]))
                             : "memory");
            }
        }
        if (pid == 0) cycles[blockIdx.x] = clock64() - t0;
    }
}

__global__ void __launch_bounds__(512) ldg_kernel(const uint4* src, long span16, int iters, uint4* sink, long long* cycles) {
    const uint4* base = src + static_cast<long>(blockIdx.x) * span16;
    uint4 acc = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcg(base + ((static_cast<long>(i + u) * 512 + threadIdx.x) % span16));
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
    if (acc.x == 0x12345678u) sink[threadIdx.x] = acc;
}

int main() {
    const long span = 1 << 20;                      // 1 MB per CTA
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned char* d;
    cudaMalloc(&d, span * sms);
    cudaMemset(d, 1, span * sms);
    long long* dc;
    cudaMalloc(&dc, sms * 8);
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[256];
    for (int grid : {1, sms})
      for (int nprod : {1, 2, 4})
        for (int sz : {8192, 16384, 32768, 65536})
            for (int slots : {2}) {
                if (static_cast<long>(nprod) * slots * sz > 196608) continue;
                const int ncopies = (4 << 20) / sz;
                const long span = grid == 1 ? (1 << 20) : (1 << 18);        // keep the all-SM case inside L2
                for (int rep = 0; rep < 2; ++rep)
                    bulk_kernel<<<grid, 128, nprod * slots * sz + 256>>>(d, span, sz, slots, ncopies, nprod, dc);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("bulk: error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                cudaMemcpy(h, dc, grid * 8, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
                printf("TMA bulk  grid=%3d producers=%d size=%5d slots=%d : %.1f B/clk/SM\n", grid, nprod, sz, slots,
                       double(ncopies) * sz * nprod / mx);
            }
    uint4* sink;
    cudaMalloc(&sink, 512 * 16);
    for (int grid : {1, sms}) {
        const int iters = 16384;                    // x 512 threads x 16 B = 128 MB
        for (int rep = 0; rep < 2; ++rep) ldg_kernel<<<grid, 512>>>(reinterpret_cast<const uint4*>(d), span / 16, iters, sink, dc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, dc, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("LDG.128   grid=%3d 512 threads x 4 in flight : %.1f B/clk/SM\n", grid, double(iters) * 512 * 16 / mx);
    }
    return 0;
}
