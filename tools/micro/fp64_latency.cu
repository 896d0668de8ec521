// Micro-benchmark: dependent-chain latency (cycles) of DFMA, DMUL, MUFU.RSQ64H-based rsqrt(), DMMA.8x8x4,
// shared-memory round trip and __syncthreads on sm_100a.  One warp / one CTA, clock64 deltas.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_kernel(double* out, long long* cyc, double seed) {
    const int N = 2048;
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, 1e-30);
    long long t1 = clock64();
    double a = x;
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a * y;
    long long t2 = clock64();
    double r = fabs(a) + 1.5;
#pragma unroll 4
    for (int i = 0; i < 256; ++i) r = rsqrt(r) + 1.5;
    long long t3 = clock64();
    double c0 = r, c1 = x;
#pragma unroll 16
    for (int i = 0; i < N; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(y), "d"(y));
    long long t4 = clock64();
    __shared__ double sh[64];
    sh[threadIdx.x] = c0;
    __syncthreads();
    double s = c1;
    for (int i = 0; i < 256; ++i) {
        sh[threadIdx.x] = s;
        __syncthreads();
        s = sh[(threadIdx.x + 1) & 31] + 1.0;
    }
    long long t5 = clock64();
    double q = s;
#pragma unroll 4
    for (int i = 0; i < 256; ++i) q = 1.0 / (q + 2.0);
    long long t6 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = (t1 - t0); cyc[1] = (t2 - t1); cyc[2] = (t3 - t2); cyc[3] = (t4 - t3); cyc[4] = (t5 - t4); cyc[5] = (t6 - t5);
    }
    out[threadIdx.x] = x + a + r + c0 + c1 + s + q;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 64 * sizeof(double));
    cudaMallocManaged(&cyc, 8 * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) {
        lat_kernel<<<1, 32>>>(out, cyc, 1.0);
        cudaDeviceSynchronize();
    }
    printf("{\"dfma_chain_cycles\": %.2f, \"dmul_chain_cycles\": %.2f, \"rsqrt_plus_add_cycles\": %.2f, \"dmma884_chain_cycles\": %.2f, "
           "\"sts_bar_lds_add_cycles\": %.2f, \"div_plus_add_cycles\": %.2f}\n",
           cyc[0] / 2048.0, cyc[1] / 2048.0, cyc[2] / 256.0, cyc[3] / 2048.0, cyc[4] / 256.0, cyc[5] / 256.0);
    return 0;
}
