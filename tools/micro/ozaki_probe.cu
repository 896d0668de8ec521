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
    cudaMalloc(&dbg, 24 * 8);
    cudaMemset(dbg, 0, 24 * 8);
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
        long long h[24];
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
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

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "time") == 0) {        // timing only; variants come through EGX_OZAKI_* (one process each)
        check_update(60, 60, true, 5);
        return 0;
    }
    int rc = check_layout();
    rc |= check_update(3, 2, false, 1);
    rc |= check_update(6, 4, false, 1);
    check_update(60, 60, true, 5);
    check_update(40, 30, true, 5);
    return rc;
}
