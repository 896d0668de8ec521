// Small kernels of the sparse-GP (FITC / VFE) likelihood around the shared K1/K2/K4/K5 building blocks.
//
// Reference being replaced: crates/gp/src/sparse_algorithm.rs fitc :695-765, vfe :769-830,
// predict_var :245-257.  Data layout: Vt = (U^-1 Kmn)^T is kept as rows = data points
// (chunk x Mpad, row-major), so every per-point quantity (nu, beta) is a row reduction and the
// M x M matrix A = I + V diag(beta) V^T is accumulated as W W^T with W = (sqrt(beta) . Vt)^T.
#include "common.cuh"
#include "../../include/egobox_gpu.h"

namespace {

__device__ __forceinline__ double sg_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace

// scal[0] += sum ln(nu)   (FITC term1)        scal[1] += sum beta_i y_i^2 (term3)
// scal[2] += sum beta_i s1_i (VFE trace(A))   scal[3] += sum y_i^2
// Per row i < mc of the chunk: s1 = |Vt_i|^2 ; FITC: nu = sigma2 - s1 + noise, beta = 1/nu ; VFE: beta = beta_const.
__global__ void __launch_bounds__(256)
    sgp_rowstats_kernel(const double* __restrict__ Y, long ldy, int mc, int mpad_rows, int Mpad,
                        const double* __restrict__ yv, int method, double sigma2, double noise, double beta_const,
                        double* __restrict__ sqrtb, double* __restrict__ by, double* __restrict__ scal) {
    __shared__ double part[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + warp;
    double t1 = 0.0, t3 = 0.0, tA = 0.0, ty = 0.0;
    if (i < mpad_rows) {
        double sb = 0.0, b_y = 0.0;
        if (i < mc) {
            const double* y = Y + static_cast<long>(i) * ldy;
            double s1 = 0.0;
            for (int j = 2 * lane; j < Mpad; j += 64) {
                const double2 v = *reinterpret_cast<const double2*>(y + j);
                s1 += v.x * v.x + v.y * v.y;
            }
            s1 = sg_warp_sum(s1);
            double beta;
            if (method == 0) {
                const double nu = (sigma2 - s1) + noise;     // knn - sum V^2 + eta2, sparse_algorithm.rs:723-726
                beta = 1.0 / nu;
                t1 = log(nu);
            } else {
                beta = beta_const;
            }
            const double yi = yv[i];
            t3 = beta * yi * yi;
            tA = beta * s1;
            ty = yi * yi;
            sb = sqrt(beta);
            b_y = beta * yi;
        }
        if (lane == 0) {
            sqrtb[i] = sb;
            by[i] = b_y;
        }
    }
    if (lane == 0) {
        part[0][warp] = t1;
        part[1][warp] = t3;
        part[2][warp] = tA;
        part[3][warp] = ty;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += part[threadIdx.x][w];
        atomicAdd(&scal[threadIdx.x], s);
    }
}

// W[a][i] = Y[i][a] * sqrtb[i]   (Mpad x rows, ldw)   and   t[a] += sum_i Y[i][a] * by[i]
__global__ void __launch_bounds__(256)
    sgp_scale_transpose_kernel(const double* __restrict__ Y, long ldy, int rows, const double* __restrict__ sqrtb,
                               const double* __restrict__ by, double* __restrict__ W, long ldw,
                               double* __restrict__ tvec) {
    __shared__ double tile[32][33];
    __shared__ double tpart[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int a0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    double tacc = 0.0;
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        const int i = i0 + r;
        const double v = Y[static_cast<long>(i) * ldy + a0 + tx];
        tile[r][tx] = v * sqrtb[i];
        tacc += v * by[i];
    }
    tpart[ty][tx] = tacc;
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) W[static_cast<long>(a0 + r) * ldw + i0 + tx] = tile[tx][r];
    if (ty == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += tpart[w][tx];
        atomicAdd(&tvec[a0 + tx], s);
    }
}

// A (lower 128-block triangle, ld) = I + sum_s partial[s] ; appended RHS row (row Mpad) = tvec
__global__ void __launch_bounds__(256)
    sgp_reduce_partials_kernel(const double* __restrict__ partial, int splits, long stride, double* __restrict__ A,
                               long ld, int Mpad, const double* __restrict__ tvec) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= Mpad) return;
    if (r == Mpad) {
        A[static_cast<long>(r) * ld + c] = tvec[c];
        return;
    }
    if ((c >> 7) > (r >> 7)) return;           // only the lower block triangle is ever read
    double s = (r == c) ? 1.0 : 0.0;
    for (int k = 0; k < splits; ++k) s += partial[static_cast<long>(k) * stride + static_cast<long>(r) * ld + c];
    A[static_cast<long>(r) * ld + c] = s;
}

// likelihood (natural logs).  scal: see sgp_rowstats_kernel.  Row Mpad of A holds b = L^-1 t.
__global__ void __launch_bounds__(256)
    sgp_final_kernel(const double* __restrict__ A, long ld, int M, int Mpad, const double* __restrict__ scal,
                     int method, int N, double sigma2, double beta_const, const int* __restrict__ info_u,
                     const int* __restrict__ info_l, double* __restrict__ out /* lik, info_u, info_l */) {
    __shared__ double red[2][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double t2 = 0.0, t4 = 0.0;
    for (int i = threadIdx.x; i < M; i += 256) {
        t2 += log(A[static_cast<long>(i) * ld + i]);
        const double b = A[static_cast<long>(Mpad) * ld + i];
        t4 += b * b;
    }
    t2 = sg_warp_sum(t2);
    t4 = sg_warp_sum(t4);
    if (lane == 0) {
        red[0][warp] = t2;
        red[1][warp] = t4;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s2 = 0.0, s4 = 0.0;
        for (int w = 0; w < 8; ++w) {
            s2 += red[0][w];
            s4 += red[1][w];
        }
        const double term2 = 2.0 * s2;
        double lik;
        if (method == 0) {
            // fitc, sparse_algorithm.rs:749-756
            lik = -0.5 * (scal[0] + term2 + scal[1] - s4);
        } else {
            // vfe :806-815 ; b = beta * Li V y  ->  |b|^2 = beta^2 |L^-1 V y|^2 (row Mpad holds L^-1 (V beta y))
            const double term1 = -static_cast<double>(N) * log(beta_const);
            const double term3 = beta_const * scal[3];
            const double term5 = static_cast<double>(N) * beta_const * sigma2;
            lik = -0.5 * (term1 + term2 + term3 - s4 + term5 - scal[2]);
        }
        out[0] = lik;
        out[1] = static_cast<double>(*info_u);
        out[2] = static_cast<double>(*info_l);
    }
}

// out[i] = sum_j Y[i][j]^2
__global__ void __launch_bounds__(256) sgp_row_sumsq_kernel(const double* __restrict__ Y, long ldy, int m, int Mpad,
                                                            double* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + warp;
    if (i >= m) return;
    const double* y = Y + static_cast<long>(i) * ldy;
    double s = 0.0;
    for (int j = 2 * lane; j < Mpad; j += 64) {
        const double2 v = *reinterpret_cast<const double2*>(y + j);
        s += v.x * v.x + v.y * v.y;
    }
    s = sg_warp_sum(s);
    if (lane == 0) out[i] = s;
}

// predict_var epilogue, sparse_algorithm.rs:245-257:  var = sigma2 - k^T inv k, floored at 1e-15, + noise
//   FITC: k^T inv k = |U^-1 k|^2 - |L^-1 U^-1 k|^2 ; VFE (reference formula): |U^-1 k|^2 + |L^-1 U^-1 k|^2
__global__ void sgp_var_kernel(const double* __restrict__ s1, const double* __restrict__ s2, int m, int method,
                               double sigma2, double noise, double* __restrict__ var) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double q = (method == 0) ? (s1[i] - s2[i]) : (s1[i] + s2[i]);
    const double v = sigma2 - q;
    var[i] = (v < 1e-15) ? 1e-15 + noise : v + noise;
}

// launchers ----------------------------------------------------------------
void launch_sgp_rowstats(const double* Y, long ldy, int mc, int mpad_rows, int Mpad, const double* yv, int method,
                         double sigma2, double noise, double beta_const, double* sqrtb, double* by, double* scal,
                         cudaStream_t s) {
    sgp_rowstats_kernel<<<(mpad_rows + 7) / 8, 256, 0, s>>>(Y, ldy, mc, mpad_rows, Mpad, yv, method, sigma2, noise,
                                                            beta_const, sqrtb, by, scal);
}
void launch_sgp_scale_transpose(const double* Y, long ldy, int rows, int Mpad, const double* sqrtb, const double* by,
                                double* W, long ldw, double* tvec, cudaStream_t s) {
    dim3 grid(Mpad / 32, rows / 32);
    sgp_scale_transpose_kernel<<<grid, 256, 0, s>>>(Y, ldy, rows, sqrtb, by, W, ldw, tvec);
}
void launch_sgp_reduce_partials(const double* partial, int splits, long stride, double* A, long ld, int Mpad,
                                const double* tvec, cudaStream_t s) {
    dim3 grid((Mpad + 255) / 256, Mpad + 1);
    sgp_reduce_partials_kernel<<<grid, 256, 0, s>>>(partial, splits, stride, A, ld, Mpad, tvec);
}
void launch_sgp_final(const double* A, long ld, int M, int Mpad, const double* scal, int method, int N, double sigma2,
                      double beta_const, const int* info_u, const int* info_l, double* out, cudaStream_t s) {
    sgp_final_kernel<<<1, 256, 0, s>>>(A, ld, M, Mpad, scal, method, N, sigma2, beta_const, info_u, info_l, out);
}
void launch_sgp_row_sumsq(const double* Y, long ldy, int m, int Mpad, double* out, cudaStream_t s) {
    sgp_row_sumsq_kernel<<<(m + 7) / 8, 256, 0, s>>>(Y, ldy, m, Mpad, out);
}
void launch_sgp_var(const double* s1, const double* s2, int m, int method, double sigma2, double noise, double* var,
                    cudaStream_t s) {
    sgp_var_kernel<<<(m + 255) / 256, 256, 0, s>>>(s1, s2, m, method, sigma2, noise, var);
}
