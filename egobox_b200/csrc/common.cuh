// Internal definitions shared by the CUDA translation units of libegobox_gpu.so.
// sm_100a only (compiled with -gencode arch=compute_100a,code=sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define EGX_NB 128            // Cholesky / TRSM block size (rows and columns)
#define EGX_CT 64             // correlation-kernel tile edge

// One multiplicative / additive term of a correlation kernel: a (dimension j,
// PLS component l) pair with theta_w[j,l] != 0 for the Matern models
// (correlation_models.rs:333-351, 505-521), or a dimension j with its summed
// weight for the exponential models (:97-100, :191-192).
struct CorrTerm {
    int dim;
    int pad_;
    double k1;   // SqExp: sum_l (theta_l W_jl)^2 | AbsExp: sum_l |W_jl| theta_l | Matern: tw = theta_l |W_jl|
    double k2;   // Matern32: sqrt(3) tw | Matern52: sqrt(5) tw
    double k3;   // Matern52: tw * tw
};

// One (input dimension j, component l) term of the theta-gradient pair kernel (kernels_thetagrad.cu).
struct ThetaGradTerm {
    int dim;
    int comp;
    double a;    // SqExp: -theta_l W_jl^2 | AbsExp: -|W_jl| | Matern: |W_jl|
    double tw;   // SqExp: (theta_l W_jl)^2 | others: theta_l |W_jl|
};

// Result block written by the GLS kernel (one per likelihood evaluation).
struct EvalResult {
    double rlf;        // reduced likelihood, algorithm.rs:1043
    double sigma2;     // rho^T rho / n  (normalised units)
    double logdet;     // (2/n) sum log10 L_ii
    double rho_sqr;
    int info;          // 0 ok, >0: 1-based index of the first non-positive pivot
    int pad_;
};

struct GemmArgs {
    double* C; long ldc;
    const double* A; long lda;
    const double* B; long ldb;
    int Mt, Nt;      // tile counts (128 x 128 tiles)
    int tri;         // tri > 0: lower-triangular tile set, first `tri` tile rows are
                     // triangular (c <= r), rows tri..Mt-1 are full (c < tri).  tri == 0: full Mt x Nt
    int K = 128;     // contraction length (multiple of 16)
    int add = 0;     // 0: C -= A B^T (factorisation / solve updates) ; 1: C += A B^T
    int splits = 1;  // split-K: blockIdx.y = s works on K-range [s*K, (s+1)*K) of A/B and on C + s*split_c_stride
    long split_c_stride = 0;
    int late_c = 0;  // set by the launcher: accumulate from zero, fold C in at the end (hides the C-tile load)
};

#define EGX_CUDA_TRY(expr)                                                         \
    do {                                                                           \
        cudaError_t e__ = (expr);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            egx_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                          __FILE__, __LINE__);                                     \
            return EGX_CUDA_ERROR;                                                 \
        }                                                                          \
    } while (0)

void egx_set_error(const char* fmt, ...);

// ---- launchers (host) ------------------------------------------------------
// kernels_corr.cu
void launch_corr_build(int corr, const double* X, int n, int npad, int d, const CorrTerm* terms,
                       int nterms, double* M, long ld, double diag_value, cudaStream_t s, double scale = 1.0);
void launch_cross_corr(int corr, const double* xraw, int m, int mpad, const double* x_mean,
                       const double* x_std, const double* X, int n, int npad, int d,
                       const CorrTerm* terms, int nterms, const double* gamma, const double* beta,
                       const int* basis_i, const int* basis_j, int p, double y_mean, double y_std,
                       double* Y, long ldy, double* yout, cudaStream_t s, double scale = 1.0);
void launch_predict_grad(int corr, const double* xraw, int m, const double* x_mean, const double* x_std, const double* X,
                         int n, int npad, int d, const CorrTerm* terms, int nterms, const double* gamma,
                         const double* beta, const int* basis_i, const int* basis_j, int p, double y_std, double* out,
                         cudaStream_t s);
void launch_mean_basis_rows(const double* X, int n, int npad, int d, const int* basis_i,
                            const int* basis_j, int p, const double* ynorm_dev, double* FyT, long ld,
                            cudaStream_t s);
// kernels_chol.cu
void launch_potrf_diag(double* Akk, long ld, int* info, int base_index, double* Dinv, cudaStream_t s);
void launch_diag_tile_update(double* C, long ldc, const double* A, long lda, int K /* 128 or 256 */, cudaStream_t s);
void launch_trsm_rows(double* X, long ldx, const double* Lkk, long ldl, const double* Dinv, double* P, long ldp,
                      int nblocks64, cudaStream_t s, double* rmaxq = nullptr /* optional [row][4] quarter-row max |x|, see kernels_ozaki.cu */);
void launch_gemm_nt_sub(const GemmArgs& g, cudaStream_t s);
int gemm_smem_bytes();
// kernels_ozaki.cu: the trailing SYRK update on tcgen05 (int8-sliced fp64)
size_t ozaki_slice_bytes(long rows);
void launch_ozaki_slice_panels(const double* P, long ldp, int rows, int kpanels, double* rscale, int8_t* S, cudaStream_t s);
bool launch_ozaki_syrk_add_panels(double* C, long ldc, const int8_t* S, const double* rscale, int tri, int kpanels, cudaStream_t s);
void launch_ozaki_slice(const double* P, long ldp, int rows, double* rscale, int8_t* S, cudaStream_t s,
                        const double* rmaxq = nullptr /* [row][4] from the panel solves: no row-maximum pass */);
void launch_ozaki_syrk(double* C, long ldc, const int8_t* S, const double* rscale, int Mt, int tri, cudaStream_t s,
                       long long* dbg = nullptr, int persist_hint = 0);
void launch_ozaki_gemm(double* C, long ldc, const int8_t* SA, const double* rsA, const int8_t* SB, const double* rsB, int Mt,
                       int Nt, cudaStream_t s, int persist_hint = 0);
// kernels_solve.cu
void launch_gls(const double* M, long ld, int n, int npad, int p, double* work, double* G,
                double* beta, double* rho, EvalResult* res, const int* info, cudaStream_t s);
void launch_backsolve_chain(const double* L, long ld, const double* Dinv, int T, double* v, int* flags, cudaStream_t s);
void launch_backsolve_diag(const double* Lkk, long ld, double* rho_k, cudaStream_t s);
void launch_backsolve_update(const double* Lrow, long ld, const double* gamma_k, double* rho, int ncolblocks,
                             cudaStream_t s);
void launch_var_finish(const double* Y, long ldy, int m, int npad, const double* xraw, const double* x_mean,
                       const double* x_std, int d, const double* FtT, long ldf, const double* G, int p,
                       const int* basis_i, const int* basis_j, double sigma2, double* var, cudaStream_t s,
                       double* U = nullptr /* optional m x p: u_i = G^-T (Ft^T rt_i - f(x_i)) */);
void launch_cov_finish(double* C, long ld, int m, int mpad, const double* U, int p, double sigma2, cudaStream_t s);
void launch_normalize_rows(const double* x, int m, int mpad, int d, const double* mean, const double* sd, double* out,
                           cudaStream_t s);
void launch_zero_upper(double* A, long ld, int npad, cudaStream_t s);
void launch_bcast_rows(double* out, long ld, int m, int mpad, int cols, const double* mean, cudaStream_t s);
int egx_host_symmetric_eig(int n, double* a, double* w);
// kernels_thetagrad.cu
int egx_fill_theta_grad_terms(int corr, int d, int h, const double* w, const double* theta, ThetaGradTerm* t);
int theta_grad_blocks(int npad);
void launch_theta_grad(int corr, const double* X, int n, int npad, int d, const ThetaGradTerm* terms, int nterms,
                       const double* Cneg_rinv, long ldc, const double* gamma, const EvalResult* res, int h,
                       double* partial, double* grad, cudaStream_t s);
void launch_set_identity(double* A, long ld, int npad, cudaStream_t s);
void launch_upper_gemv(const double* W, long ld, int n, const double* rho, double* out, cudaStream_t s);

// cached allocators (devmem.cu): same contract as cudaMalloc / cudaFree / cudaMallocHost / cudaFreeHost
cudaError_t egx_dev_malloc_bytes(void** p, size_t bytes);
void egx_dev_free(void* p);
cudaError_t egx_host_malloc_bytes(void** p, size_t bytes);
void egx_host_free(void* p);
void egx_mem_trim();
template <typename T>
inline cudaError_t egx_dev_malloc(T** p, size_t bytes) {
    return egx_dev_malloc_bytes(reinterpret_cast<void**>(p), bytes);
}
template <typename T>
inline cudaError_t egx_host_malloc(T** p, size_t bytes) {
    return egx_host_malloc_bytes(reinterpret_cast<void**>(p), bytes);
}
// small_batch.cu
size_t small_batch_smem_bytes(int n, int d, int h, int p);
void launch_small_batch(int corr, const double* X, int n, int d, const double* W, int h, const double* thetas, int B,
                        const double* FyT, long ldf, int p, double diag_value, void* out, double* out_G,
                        cudaStream_t s);
struct SmallOutHost {
    double rlf, sigma2;
    int info, pad_;
};
// kernels_grad.cu
void launch_transpose_lower(const double* L, long ld, double* LT, int T, cudaStream_t s);
void launch_trsm_rows_upper(double* X, long ldx, const double* LTkk, long ldl, const double* Dinv, double* P,
                            int nblocks64, cudaStream_t s);
size_t var_grad_smem_bytes(int d, int p, int nterms);
void launch_var_grad(int corr, const double* W, long ldw, const double* xraw, int m, const double* x_mean,
                     const double* x_std, const double* X, int n, int npad, int d, const CorrTerm* terms, int nterms,
                     const double* KF, long ldk, const double* G, int p, const int* basis_i, const int* basis_j,
                     double sigma2, double* out, cudaStream_t s);
