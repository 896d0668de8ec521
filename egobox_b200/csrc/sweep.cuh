// Host-side execution helpers shared by the dense-GP and sparse-GP contexts:
// stage profiler (CUDA events on the launching stream), the two-stream sweep environment and the
// blocked "solve block column k, update the trailing columns" sweep (Cholesky factorisation and
// multi-RHS triangular solve).
#pragma once
#include <vector>

#include "common.cuh"
#include "../../include/egobox_gpu.h"

struct ProfEvent {
    cudaEvent_t a, b;
    int stage;
};

struct Profiler {
    bool on = false;
    std::vector<ProfEvent> pending;
    std::vector<cudaEvent_t> pool;
    double ms[EGX_NUM_STAGES] = {0};
    long long launches[EGX_NUM_STAGES] = {0};
    void resolve();
    void reset();
    void destroy();
};

struct StageScope {
    Profiler* p;
    ProfEvent ev;
    bool on;
    cudaStream_t st;
    StageScope(Profiler& prof, int stage, int launches, cudaStream_t stream);
    ~StageScope();
};

struct SweepEnv {
    cudaStream_t sb = nullptr;   // bulk stream
    cudaStream_t sp = nullptr;   // high-priority panel / look-ahead stream
    cudaStream_t sq = nullptr;   // high-priority stream of the column updates that run beside the diagonal-block chain (r02)
    bool lookahead = true;
    std::vector<cudaEvent_t> ev_panel, ev_bulk;
    std::vector<cudaEvent_t> ev_trsm_a, ev_partner, ev_colrest, ev_slice;   // r02 look-ahead: per column pair
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join_q = nullptr;
    int lookahead_v = 2;                  // EGX_LOOKAHEAD_V=1: the r01 schedule (whole columns on the panel stream)
    double* P2[2] = {nullptr, nullptr};   // double-buffered contiguous panel copies (rows x 128)
    long p_rows = 0;
    // tcgen05 path of the trailing update (kernels_ozaki.cu): int8 slices + row scales of the current panel pair
    int8_t* oz_S = nullptr;
    double* oz_scale = nullptr;
    int8_t* oz_S2 = nullptr;              // second slice / scale buffer: the look-ahead schedule slices pair p + 1 on `sq` while the bulk
    double* oz_scale2 = nullptr;          // update of pair p still reads the slices of pair p on `sb` (buffer = pair & 1)
    std::vector<cudaEvent_t>* solve_pair_events = nullptr;   // solves only: event [pair] recorded when the two block columns of a pair are final
    int la_ozaki = 1;                     // EGX_LA_OZAKI=0: look-ahead column updates on the DMMA kernel (r02 first form)
    double* oz_rmaxq[2] = {nullptr, nullptr};   // [row][4] quarter-row maxima written by the panel solves, per P2 buffer
    int ozaki = 1;                        // EGX_OZAKI=0 keeps every update on the DMMA kernel
    int oz_persist = 0;                   // set by the batched entry point: several evaluations share the GPU
    int ozaki_min_tri = 4;                // smallest trailing tile-triangle worth the slicing pass (EGX_OZAKI_MIN_TRI; r02: 8 -> 4, C5 66.1 -> 64.9 ms)
    int ozaki_min_tri_solve = 8;          // the same for the updates of a multi-RHS solve (row_tiles x tri2 tiles: the sparse GP
                                          // solves 64 row tiles against 8 block columns and sets 2)
    int ozaki_min_T = 1;                  // smallest factor (block columns) that uses it at all (EGX_OZAKI_MIN_T); measured
                                          // on batches of 96: ahead of DMMA from n = 2048 (0.165 vs 0.187 ms) upwards
    int* bs_flags = nullptr;              // per-block-column flags of the chained back substitution (max_block_cols ints)
    int bs_flags_n = 0;
    int generation = 0;                   // bumped when a buffer captured in a CUDA graph is reallocated
    Profiler prof;
    int init(int max_block_cols);
    int ensure_panel_rows(long rows);
    void destroy();
};

// A lower-triangular factor stored as 128-blocks in a row-major matrix (ld), T block columns,
// optionally followed by `qpad` appended right-hand-side rows (factor sweeps only).
struct FactorRef {
    double* M;
    long ld;
    int T;
    int qpad;
    double* Dinv;   // [T][4][32][32]
    int* info;
    // optional (solves only): int8 digit slices + row scales of the block rows of L below every column PAIR, in the
    // layout of kernels_ozaki.cu; pair p = block columns 2p, 2p+1 starts at Lsl + Lsl_off[p] bytes / Lsc + Lsc_off[p]
    const int8_t* Lsl = nullptr;
    const double* Lsc = nullptr;
    const long* Lsl_off = nullptr;
    const long* Lsc_off = nullptr;
    // optional (factor sweeps with the look-ahead schedule only): write the slices of the panel rows of pair p -- which ARE the
    // slices of the block rows of L below that pair, the B operand of later solves -- at Lsl_w + Lsl_off_w[p] / Lsc_w + Lsc_off_w[p]
    // instead of the two rotating buffers; *Lsl_w_pairs = number of leading pairs written (room per pair: T - 2p - 2 row blocks of
    // L plus the qpad appended rows)
    int8_t* Lsl_w = nullptr;
    double* Lsc_w = nullptr;
    const long* Lsl_off_w = nullptr;
    const long* Lsc_off_w = nullptr;
    int* Lsl_w_pairs = nullptr;
};

// factor = true : in-place Cholesky of f.M (+ appended rows become (L^-1 B)^T)
// factor = false: rows (row_tiles*128 x T*128, ld_rows) <- rows * L^-T using the factor f
// upper_rows: the rows are upper triangular (factor = false only): step k works on row tiles 0 .. k+1
void blocked_sweep(SweepEnv& env, const FactorRef& f, bool factor, double* rows, long ld_rows, int row_tiles,
                   int slabs64, bool upper_rows = false);

// Per-theta kernel weights: (dimension, component) term list consumed by K1/K2 (correlation_models.rs
// :97-100, 191, 333, 505).  Returns the number of terms written (<= d*h).
int egx_fill_terms(int corr, int d, int h, const double* w_star, const double* theta, CorrTerm* out);

// gamma-style single-vector back substitution  v <- L^-T v  with the blocked factor f (algorithm.rs:1034)
void backsolve_vector(SweepEnv& env, const FactorRef& f, double* v);

// Trajectories from a covariance on the device (gp/src/algorithm.rs:1153-1194): K (mpad x mpad, identity on the padding) is
// factorised in place (method 0: Cholesky with the blocked sweep; 1: host eigen-decomposition, eigenvalues < 1e-9 dropped) and
// out (m x n_traj, host) = mean + C z.  `s` must be env.sb.
int sample_from_covariance(SweepEnv& env, cudaStream_t s, double* K, int m, int mpad, const double* mean_dev, const double* z,
                           int n_traj, int method, double* out);
