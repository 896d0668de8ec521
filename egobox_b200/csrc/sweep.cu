// Implementation of the shared sweep environment (see sweep.cuh).
#include <cmath>
#include <cstdlib>
#include <vector>

#include "sweep.cuh"

void Profiler::resolve() {
    for (auto& ev : pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ev.a, ev.b) == cudaSuccess) ms[ev.stage] += t;
        pool.push_back(ev.a);
        pool.push_back(ev.b);
    }
    pending.clear();
}
void Profiler::reset() {
    for (int i = 0; i < EGX_NUM_STAGES; ++i) {
        ms[i] = 0.0;
        launches[i] = 0;
    }
}
void Profiler::destroy() {
    for (auto& ev : pending) {
        cudaEventDestroy(ev.a);
        cudaEventDestroy(ev.b);
    }
    pending.clear();
    for (auto e : pool) cudaEventDestroy(e);
    pool.clear();
}

StageScope::StageScope(Profiler& prof, int stage, int launches, cudaStream_t stream)
    : p(&prof), on(prof.on), st(stream) {
    p->launches[stage] += launches;
    if (on) {
        ev.stage = stage;
        for (cudaEvent_t* e : {&ev.a, &ev.b}) {
            if (!p->pool.empty()) {
                *e = p->pool.back();
                p->pool.pop_back();
            } else {
                cudaEventCreate(e);
            }
        }
        cudaEventRecord(ev.a, st);
    }
}
StageScope::~StageScope() {
    if (on) {
        cudaEventRecord(ev.b, st);
        p->pending.push_back(ev);
    }
}

int SweepEnv::init(int max_block_cols) {
    int lo = 0, hi = 0;
    EGX_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    EGX_CUDA_TRY(cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, lo));
    EGX_CUDA_TRY(cudaStreamCreateWithPriority(&sp, cudaStreamNonBlocking, hi));
    EGX_CUDA_TRY(cudaStreamCreateWithPriority(&sq, cudaStreamNonBlocking, hi));
    const int npairs = (max_block_cols + 1) / 2;
    ev_panel.assign(max_block_cols, nullptr);
    ev_bulk.assign(max_block_cols, nullptr);
    ev_trsm_a.assign(npairs, nullptr);
    ev_partner.assign(npairs, nullptr);
    ev_colrest.assign(npairs, nullptr);
    ev_slice.assign(npairs, nullptr);
    for (int k = 0; k < max_block_cols; ++k) {
        EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_panel[k], cudaEventDisableTiming));
        EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_bulk[k], cudaEventDisableTiming));
    }
    for (int k = 0; k < npairs; ++k) {
        EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_trsm_a[k], cudaEventDisableTiming));
        EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_partner[k], cudaEventDisableTiming));
        EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_colrest[k], cudaEventDisableTiming));
        EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_slice[k], cudaEventDisableTiming));
    }
    EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_join_q, cudaEventDisableTiming));
    if (const char* v = getenv("EGX_LOOKAHEAD_V")) lookahead_v = atoi(v);
    if (const char* v = getenv("EGX_LA_OZAKI")) la_ozaki = atoi(v);
    bs_flags_n = max_block_cols;
    EGX_CUDA_TRY(egx_dev_malloc(&bs_flags, static_cast<size_t>(max_block_cols > 0 ? max_block_cols : 1) * sizeof(int)));
    const char* e = getenv("EGX_LOOKAHEAD");
    lookahead = !(e != nullptr && atoi(e) == 0);
    if ((e = getenv("EGX_OZAKI")) != nullptr) ozaki = atoi(e);
    if ((e = getenv("EGX_OZAKI_MIN_TRI")) != nullptr) {
        ozaki_min_tri = atoi(e) > 1 ? atoi(e) : 1;
        ozaki_min_tri_solve = ozaki_min_tri;
    }
    if ((e = getenv("EGX_OZAKI_MIN_T")) != nullptr) ozaki_min_T = atoi(e);
    return EGX_OK;
}

int SweepEnv::ensure_panel_rows(long rows) {
    if (rows <= p_rows) return EGX_OK;
    for (int i = 0; i < 2; ++i) {
        egx_dev_free(P2[i]);
        P2[i] = nullptr;
        EGX_CUDA_TRY(egx_dev_malloc(&P2[i], static_cast<size_t>(rows) * 2 * EGX_NB * sizeof(double)));
    }
    egx_dev_free(oz_S);
    egx_dev_free(oz_scale);
    egx_dev_free(oz_S2);
    egx_dev_free(oz_scale2);
    oz_S = oz_S2 = nullptr;
    oz_scale = oz_scale2 = nullptr;
    for (int i = 0; i < 2; ++i) {
        egx_dev_free(oz_rmaxq[i]);
        oz_rmaxq[i] = nullptr;
    }
    if (ozaki) {
        EGX_CUDA_TRY(egx_dev_malloc(&oz_S, ozaki_slice_bytes(rows)));
        EGX_CUDA_TRY(egx_dev_malloc(&oz_scale, static_cast<size_t>(rows) * sizeof(double)));
        if (la_ozaki) {
            EGX_CUDA_TRY(egx_dev_malloc(&oz_S2, ozaki_slice_bytes(rows)));
            EGX_CUDA_TRY(egx_dev_malloc(&oz_scale2, static_cast<size_t>(rows) * sizeof(double)));
        }
        for (int i = 0; i < 2; ++i) EGX_CUDA_TRY(egx_dev_malloc(&oz_rmaxq[i], static_cast<size_t>(rows) * 4 * sizeof(double)));
    }
    p_rows = rows;
    ++generation;
    return EGX_OK;
}

void SweepEnv::destroy() {
    if (sb) cudaStreamSynchronize(sb);
    if (sp) cudaStreamSynchronize(sp);
    if (sq) cudaStreamSynchronize(sq);
    prof.destroy();
    for (auto* vec : {&ev_panel, &ev_bulk, &ev_trsm_a, &ev_partner, &ev_colrest, &ev_slice}) {
        for (auto e : *vec)
            if (e) cudaEventDestroy(e);
        vec->clear();
    }
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (ev_join_q) cudaEventDestroy(ev_join_q);
    ev_fork = ev_join = ev_join_q = nullptr;
    egx_dev_free(bs_flags);
    bs_flags = nullptr;
    egx_dev_free(P2[0]);
    egx_dev_free(P2[1]);
    egx_dev_free(oz_S);
    egx_dev_free(oz_scale);
    egx_dev_free(oz_S2);
    egx_dev_free(oz_scale2);
    oz_S = oz_S2 = nullptr;
    oz_scale = oz_scale2 = nullptr;
    for (int i = 0; i < 2; ++i) {
        egx_dev_free(oz_rmaxq[i]);
        oz_rmaxq[i] = nullptr;
    }
    if (sq) cudaStreamDestroy(sq);
    if (sp) cudaStreamDestroy(sp);
    if (sb) cudaStreamDestroy(sb);
    sb = sp = sq = nullptr;
}

// Two-level blocking: block columns are processed in PAIRS (outer block 256 = two 128-panels).  Panel A, the
// K=128 update of the partner column, panel B, then ONE trailing update with K = 256 over both panels
// (side by side in a rows x 256 buffer): half as many passes over the trailing matrix and half as many tile
// prologues/epilogues as a K = 128 update per panel.
// With look-ahead (factorisation only) the panel chain and the update of the NEXT pair's two block columns run
// on the high-priority stream while the bulk of the trailing update is still in flight on the bulk stream.
// The multi-RHS solve has no serial part and runs as plain back-to-back launches (measured 22.9 vs 28.8 ms
// per 8192-point chunk at n = 8192 with / without stream splitting).
// The K = 256 trailing update of the factorisation, C(tile rows Mt, first `tri` triangular) -= A A^T with A the
// (Mt * 128) x 256 panel-pair rows: on tcgen05 through the int8 slices when the tile set is large enough to pay
// for the slicing pass, else on the DMMA kernel.
static void trailing_syrk(SweepEnv& env, const GemmArgs& g, cudaStream_t st, int T, const double* rmaxq) {
    if (env.ozaki && env.oz_S != nullptr && T >= env.ozaki_min_T && g.tri >= env.ozaki_min_tri && g.K == 2 * EGX_NB && g.A == g.B &&
        g.lda == 2 * EGX_NB && static_cast<long>(g.Mt) * EGX_NB <= env.p_rows) {
        {
            StageScope sc(env.prof, EGX_STAGE_OZAKI_SLICE, rmaxq ? 1 : 2, st);
            launch_ozaki_slice(g.A, g.lda, g.Mt * EGX_NB, env.oz_scale, env.oz_S, st, rmaxq);
        }
        StageScope sc(env.prof, EGX_STAGE_OZAKI_SYRK, 1, st);
        launch_ozaki_syrk(g.C, g.ldc, env.oz_S, env.oz_scale, g.Mt, g.tri, st, nullptr, env.oz_persist);
        return;
    }
    StageScope sc(env.prof, EGX_STAGE_SYRK_GEMM, 1, st);
    launch_gemm_nt_sub(g, st);
}

// Factorisation with the r02 look-ahead schedule: the serial chain of a block column is
//   diagonal-tile update (ten 32 x 32 CTAs) -> K3 -> K5
// on the panel stream `sp`; the rest of the column (everything below the diagonal tile) is updated on `sq` WHILE K3 runs
// and joins before K5.  r01 updated whole columns on the panel stream before K3 (EGX_LOOKAHEAD_V=1): one 128 x 64 tile with
// K = 256 keeps an SM busy for 17 us and the early columns are bound by the FP64 pipe of the whole GPU (28 us at n = 8192).
static void factor_sweep_lookahead(SweepEnv& env, const FactorRef& f) {
    const int T = f.T, Qt = f.qpad / EGX_NB;
    const long ld = f.ld;
    const long LDP = 2 * EGX_NB;
    cudaStream_t sb = env.sb, sp = env.sp, sq = env.sq;
    cudaEventRecord(env.ev_fork, sb);
    cudaStreamWaitEvent(sp, env.ev_fork, 0);
    cudaStreamWaitEvent(sq, env.ev_fork, 0);
    auto blk = [&](int r, int c) { return f.M + static_cast<long>(r) * EGX_NB * ld + static_cast<long>(c) * EGX_NB; };
    bool rest_pending = false;       // sq still updates this pair's columns (ev_colrest of the previous pair)
    int slices_kept = 0;             // leading pairs whose slices went to the caller's persistent buffer (f.Lsl_w)
    for (int k = 0; k < T; k += 2) {
        const int pair = k >> 1;
        double* Pw = env.P2[pair & 1];                 // rows x 256, row 0 = first row of block k+1
        double* Rq = env.ozaki ? env.oz_rmaxq[pair & 1] : nullptr;
        const bool two = (k + 1 < T);
        // ---- panel A ----------------------------------------------------------------------------------
        {
            StageScope sc(env.prof, EGX_STAGE_POTRF_DIAG, 1, sp);
            launch_potrf_diag(blk(k, k), ld, f.info, k * EGX_NB, f.Dinv + static_cast<long>(k) * 4096, sp);
        }
        if (rest_pending) {
            cudaStreamWaitEvent(sp, env.ev_colrest[pair - 1], 0);
            rest_pending = false;
        }
        const int rows_below = (T - k - 1) * EGX_NB + f.qpad;
        if (rows_below > 0) {
            StageScope sc(env.prof, EGX_STAGE_TRSM_PANEL, 1, sp);
            launch_trsm_rows(blk(k + 1, k), ld, blk(k, k), ld, f.Dinv + static_cast<long>(k) * 4096, Pw, LDP, rows_below / 64, sp, Rq);
        }
        if (!two) break;
        const int tri1 = T - k - 1;                    // block columns right of k
        cudaEventRecord(env.ev_trsm_a[pair], sp);
        // ---- partner column k+1: its diagonal tile on the chain, the rows below it beside K3 --------------
        {
            StageScope sc(env.prof, EGX_STAGE_GEMM_LOOKAHEAD, 1, sp);
            launch_diag_tile_update(blk(k + 1, k + 1), ld, Pw, LDP, EGX_NB, sp);
        }
        const int rest1 = tri1 - 1 + Qt;
        if (rest1 > 0) {
            cudaStreamWaitEvent(sq, env.ev_trsm_a[pair], 0);
            GemmArgs g;
            g.C = blk(k + 2, k + 1);
            g.ldc = ld;
            g.A = Pw + static_cast<long>(EGX_NB) * LDP;
            g.lda = LDP;
            g.B = Pw;
            g.ldb = LDP;
            g.K = EGX_NB;
            g.tri = 0;
            g.Mt = rest1;
            g.Nt = 1;
            StageScope sc(env.prof, EGX_STAGE_GEMM_LOOKAHEAD, 1, sq);
            launch_gemm_nt_sub(g, sq);
            cudaEventRecord(env.ev_partner[pair], sq);
        }
        {
            StageScope sc(env.prof, EGX_STAGE_POTRF_DIAG, 1, sp);
            launch_potrf_diag(blk(k + 1, k + 1), ld, f.info, (k + 1) * EGX_NB, f.Dinv + static_cast<long>(k + 1) * 4096, sp);
        }
        if (rest1 > 0) {
            cudaStreamWaitEvent(sp, env.ev_partner[pair], 0);
            StageScope sc(env.prof, EGX_STAGE_TRSM_PANEL, 1, sp);
            launch_trsm_rows(blk(k + 2, k + 1), ld, blk(k + 1, k + 1), ld, f.Dinv + static_cast<long>(k + 1) * 4096,
                             Pw + static_cast<long>(EGX_NB) * LDP + EGX_NB, LDP, rest1 * 2, sp,
                             Rq ? Rq + static_cast<long>(EGX_NB) * 4 + 2 : nullptr);
        }
        const int tri2 = T - k - 2;                    // block columns right of the pair
        cudaEventRecord(env.ev_panel[pair], sp);
        if (tri2 <= 0) continue;
        // ---- next pair's columns, K = 256: diagonal tile on the chain, the rest beside the next K3 -------------
        if (pair > 0) cudaStreamWaitEvent(sp, env.ev_bulk[pair - 1], 0);
        {
            StageScope sc(env.prof, EGX_STAGE_GEMM_LOOKAHEAD, 1, sp);
            launch_diag_tile_update(blk(k + 2, k + 2), ld, Pw + static_cast<long>(EGX_NB) * LDP, LDP, 2 * EGX_NB, sp);
        }
        const int nla = tri2 < 2 ? tri2 : 2;
        const int rest2 = tri2 - 1 + Qt;
        // tcgen05 for the look-ahead columns too (r02b): the panel rows from block k+2 on are sliced ONCE on `sq`, the columns of
        // the next pair are updated from those slices in the rectangular form and the bulk update on `sb` reuses them -- the
        // 2 x rest2 tiles of K = 256 leave the FP64 pipe (0.3 ms of DMMA per evaluation at n = 8192, a sixth of the GPU work of
        // an evaluation at n = 4096).  Slices are double-buffered by pair: the bulk update of pair p may still read its slices
        // when pair p + 1 is sliced.
        const bool oz_la = env.la_ozaki && env.ozaki && env.oz_S != nullptr && env.oz_S2 != nullptr && T >= env.ozaki_min_T &&
                           tri2 - 2 >= env.ozaki_min_tri && static_cast<long>(tri2 + Qt) * EGX_NB <= env.p_rows;
        if (oz_la) {
            int8_t* S = (pair & 1) ? env.oz_S2 : env.oz_S;
            double* sc = (pair & 1) ? env.oz_scale2 : env.oz_scale;
            if (f.Lsl_w != nullptr && slices_kept == pair) {     // kept for the solves that follow (predict_var, W = L^-T)
                S = f.Lsl_w + f.Lsl_off_w[pair];
                sc = f.Lsc_w + f.Lsc_off_w[pair];
                ++slices_kept;
            }
            const size_t rb = ozaki_slice_bytes(EGX_NB);        // bytes of the slices of one 128-row block
            cudaStreamWaitEvent(sq, env.ev_panel[pair], 0);
            if (pair > 0) cudaStreamWaitEvent(sq, env.ev_bulk[pair - 1], 0);
            {
                StageScope sc1(env.prof, EGX_STAGE_OZAKI_SLICE, Rq ? 1 : 2, sq);
                launch_ozaki_slice(Pw + static_cast<long>(EGX_NB) * LDP, LDP, (tri2 + Qt) * EGX_NB, sc, S, sq,
                                   Rq ? Rq + static_cast<long>(EGX_NB) * 4 : nullptr);
            }
            cudaEventRecord(env.ev_slice[pair], sq);
            {
                StageScope sc2(env.prof, EGX_STAGE_GEMM_LOOKAHEAD, 1, sq);
                launch_ozaki_gemm(blk(k + 3, k + 2), ld, S + rb, sc + EGX_NB, S, sc, rest2, nla, sq, env.oz_persist);
            }
            cudaEventRecord(env.ev_colrest[pair], sq);
            rest_pending = true;
            cudaStreamWaitEvent(sb, env.ev_panel[pair], 0);
            cudaStreamWaitEvent(sb, env.ev_slice[pair], 0);
            {
                StageScope sc3(env.prof, EGX_STAGE_OZAKI_SYRK, 1, sb);
                launch_ozaki_syrk(blk(k + 4, k + 4), ld, S + 2 * rb, sc + 2 * EGX_NB, tri2 - 2 + Qt, tri2 - 2, sb, nullptr, env.oz_persist);
            }
            cudaEventRecord(env.ev_bulk[pair], sb);
            continue;
        }
        if (rest2 > 0) {
            cudaStreamWaitEvent(sq, env.ev_panel[pair], 0);
            if (pair > 0) cudaStreamWaitEvent(sq, env.ev_bulk[pair - 1], 0);
            GemmArgs g;
            g.C = blk(k + 3, k + 2);
            g.ldc = ld;
            g.A = Pw + static_cast<long>(2 * EGX_NB) * LDP;
            g.lda = LDP;
            g.B = Pw + static_cast<long>(EGX_NB) * LDP;
            g.ldb = LDP;
            g.K = 2 * EGX_NB;
            g.tri = 0;
            g.Mt = rest2;
            g.Nt = nla;
            StageScope sc(env.prof, EGX_STAGE_GEMM_LOOKAHEAD, 1, sq);
            launch_gemm_nt_sub(g, sq);
            cudaEventRecord(env.ev_colrest[pair], sq);
            rest_pending = true;
        }
        // ---- bulk: block columns k+4 .., on the bulk stream -------------------------------------------------
        cudaStreamWaitEvent(sb, env.ev_panel[pair], 0);
        if (tri2 > 2) {
            GemmArgs gb;
            gb.K = 2 * EGX_NB;
            gb.lda = LDP;
            gb.ldb = LDP;
            gb.ldc = ld;
            gb.C = blk(k + 4, k + 4);
            gb.A = Pw + static_cast<long>(3 * EGX_NB) * LDP;
            gb.B = gb.A;
            gb.tri = tri2 - 2;
            gb.Mt = tri2 - 2 + Qt;
            gb.Nt = tri2 - 2;
            trailing_syrk(env, gb, sb, T, Rq ? Rq + static_cast<long>(3 * EGX_NB) * 4 : nullptr);
        }
        cudaEventRecord(env.ev_bulk[pair], sb);
    }
    cudaEventRecord(env.ev_join, sp);
    cudaStreamWaitEvent(sb, env.ev_join, 0);
    cudaEventRecord(env.ev_join_q, sq);
    cudaStreamWaitEvent(sb, env.ev_join_q, 0);
    if (f.Lsl_w_pairs != nullptr) *f.Lsl_w_pairs = slices_kept;
}

void blocked_sweep(SweepEnv& env, const FactorRef& f, bool factor, double* rows, long ld_rows, int row_tiles_all,
                   int slabs64_all, bool upper_rows) {
    const int T = f.T, Qt = f.qpad / EGX_NB;
    const long ld = f.ld;
    const long LDP = 2 * EGX_NB;
    const bool la = env.lookahead && factor && T > 4 && static_cast<int>(env.ev_panel.size()) >= (T + 1) / 2;
    if (f.Lsl_w_pairs != nullptr) *f.Lsl_w_pairs = 0;
    if (la && env.lookahead_v != 1 && env.sq != nullptr && static_cast<int>(env.ev_colrest.size()) >= (T + 1) / 2) {
        factor_sweep_lookahead(env, f);
        return;
    }
    cudaStream_t sb = env.sb, sp = la ? env.sp : env.sb;
    if (la) {
        cudaEventRecord(env.ev_fork, sb);
        cudaStreamWaitEvent(sp, env.ev_fork, 0);
    }
    auto blk = [&](int r, int c) { return f.M + static_cast<long>(r) * EGX_NB * ld + static_cast<long>(c) * EGX_NB; };
    for (int k = 0; k < T; k += 2) {
        const int pair = k >> 1;
        double* Pw = env.P2[pair & 1];                 // rows x 256; factor: row 0 = first row of block k+1
        double* Rq = (env.ozaki && (factor || f.Lsl != nullptr)) ? env.oz_rmaxq[pair & 1] : nullptr;   // quarter-row maxima, same row origin as Pw
        const bool two = (k + 1 < T);
        // upper_rows (solves only): the rows are upper triangular (the identity on entry), so row tiles below block
        // column k + 1 are still zero in this pair's columns and their updates would subtract zeros
        const int row_tiles = (upper_rows && k + 2 < row_tiles_all) ? k + 2 : row_tiles_all;
        const int slabs64 = (upper_rows && k + 2 < row_tiles_all) ? 2 * (k + 2) : slabs64_all;
        // ---- panel A (block column k) -----------------------------------------------------------------
        if (factor) {
            {
                StageScope sc(env.prof, EGX_STAGE_POTRF_DIAG, 1, sp);
                launch_potrf_diag(blk(k, k), ld, f.info, k * EGX_NB, f.Dinv + static_cast<long>(k) * 4096, sp);
            }
            const int rows_below = (T - k - 1) * EGX_NB + f.qpad;
            if (rows_below > 0) {
                StageScope sc(env.prof, EGX_STAGE_TRSM_PANEL, 1, sp);
                launch_trsm_rows(blk(k + 1, k), ld, blk(k, k), ld, f.Dinv + static_cast<long>(k) * 4096, Pw, LDP,
                                 rows_below / 64, sp, Rq);
            }
        } else {
            StageScope sc(env.prof, EGX_STAGE_TRSM_PANEL, 1, sp);
            launch_trsm_rows(rows + static_cast<long>(k) * EGX_NB, ld_rows, blk(k, k), ld,
                             f.Dinv + static_cast<long>(k) * 4096, Pw, LDP, slabs64, sp, Rq);
        }
        if (!two) {
            if (!factor && env.solve_pair_events != nullptr) cudaEventRecord((*env.solve_pair_events)[pair], sp);
            break;
        }
        const int tri1 = T - k - 1;                    // block columns right of k
        // ---- partner column k+1: K = 128 update with panel A ------------------------------------------
        {
            GemmArgs g;
            g.A = Pw;
            g.lda = LDP;
            g.K = EGX_NB;
            g.tri = 0;
            g.Nt = 1;
            if (factor) {
                g.C = blk(k + 1, k + 1);
                g.ldc = ld;
                g.B = Pw;
                g.ldb = LDP;
                g.Mt = tri1 + Qt;
            } else {
                g.C = rows + static_cast<long>(k + 1) * EGX_NB;
                g.ldc = ld_rows;
                g.B = blk(k + 1, k);
                g.ldb = ld;
                g.Mt = row_tiles;
            }
            StageScope sc(env.prof, la ? EGX_STAGE_GEMM_LOOKAHEAD : EGX_STAGE_SYRK_GEMM, 1, sp);
            launch_gemm_nt_sub(g, sp);
        }
        // ---- panel B (block column k+1) -> columns 128..255 of Pw ---------------------------------------
        if (factor) {
            {
                StageScope sc(env.prof, EGX_STAGE_POTRF_DIAG, 1, sp);
                launch_potrf_diag(blk(k + 1, k + 1), ld, f.info, (k + 1) * EGX_NB, f.Dinv + static_cast<long>(k + 1) * 4096,
                                  sp);
            }
            const int rows_below = (T - k - 2) * EGX_NB + f.qpad;
            if (rows_below > 0) {
                StageScope sc(env.prof, EGX_STAGE_TRSM_PANEL, 1, sp);
                launch_trsm_rows(blk(k + 2, k + 1), ld, blk(k + 1, k + 1), ld, f.Dinv + static_cast<long>(k + 1) * 4096,
                                 Pw + static_cast<long>(EGX_NB) * LDP + EGX_NB, LDP, rows_below / 64, sp,
                                 Rq ? Rq + static_cast<long>(EGX_NB) * 4 + 2 : nullptr);
            }
        } else {
            StageScope sc(env.prof, EGX_STAGE_TRSM_PANEL, 1, sp);
            launch_trsm_rows(rows + static_cast<long>(k + 1) * EGX_NB, ld_rows, blk(k + 1, k + 1), ld,
                             f.Dinv + static_cast<long>(k + 1) * 4096, Pw + EGX_NB, LDP, slabs64, sp, Rq ? Rq + 2 : nullptr);
        }
        const int tri2 = T - k - 2;                    // block columns right of the pair
        if (la) cudaEventRecord(env.ev_panel[pair], sp);
        if (!factor && env.solve_pair_events != nullptr) cudaEventRecord((*env.solve_pair_events)[pair], sp);
        if (tri2 <= 0) continue;                       // nothing right of the pair (appended rows only touch existing columns)
        // ---- trailing update with K = 256 ---------------------------------------------------------------
        GemmArgs g;
        g.K = 2 * EGX_NB;
        g.lda = LDP;
        if (factor) {
            g.A = Pw + static_cast<long>(EGX_NB) * LDP;      // rows from block k+2
            g.B = g.A;
            g.ldb = LDP;
            g.C = blk(k + 2, k + 2);
            g.ldc = ld;
        } else {
            g.A = Pw;
            g.B = blk(k + 2, k);                            // L[(k+2).., 256 columns of the pair]
            g.ldb = ld;
            g.C = rows + static_cast<long>(k + 2) * EGX_NB;
            g.ldc = ld_rows;
        }
        if (!la) {
            g.tri = factor ? tri2 : 0;
            g.Mt = factor ? tri2 + Qt : row_tiles;
            g.Nt = tri2;
            if (factor) {
                trailing_syrk(env, g, sb, T, Rq ? Rq + static_cast<long>(EGX_NB) * 4 : nullptr);
            } else if (f.Lsl != nullptr && env.ozaki && env.oz_S != nullptr && T >= env.ozaki_min_T && tri2 >= env.ozaki_min_tri_solve &&
                       static_cast<long>(row_tiles) * EGX_NB <= env.p_rows) {
                // multi-RHS solve on tcgen05: rows[:, (k+2)..] -= Pw L[(k+2).., pair]^T with the slices of L made once per
                // model and the slices of the freshly solved rows made here
                {
                    StageScope sc(env.prof, EGX_STAGE_OZAKI_SLICE, 1, sb);
                    launch_ozaki_slice(Pw, LDP, row_tiles * EGX_NB, env.oz_scale, env.oz_S, sb, Rq);
                }
                StageScope sc(env.prof, EGX_STAGE_OZAKI_SYRK, 1, sb);
                launch_ozaki_gemm(g.C, g.ldc, env.oz_S, env.oz_scale, f.Lsl + f.Lsl_off[pair], f.Lsc + f.Lsc_off[pair], row_tiles,
                                  tri2, sb, 1);
            } else {
                StageScope sc(env.prof, EGX_STAGE_SYRK_GEMM, 1, sb);
                launch_gemm_nt_sub(g, sb);
            }
            continue;
        }
        // look-ahead part: the next pair's two block columns, on the panel stream
        if (pair > 0) cudaStreamWaitEvent(sp, env.ev_bulk[pair - 1], 0);
        const int nla = tri2 < 2 ? tri2 : 2;
        {
            GemmArgs ga = g;
            ga.tri = 0;
            ga.Mt = tri2 + Qt;
            ga.Nt = nla;
            StageScope sc(env.prof, EGX_STAGE_GEMM_LOOKAHEAD, 1, sp);
            launch_gemm_nt_sub(ga, sp);
        }
        // bulk: block columns k+4 .., on the bulk stream
        cudaStreamWaitEvent(sb, env.ev_panel[pair], 0);
        if (tri2 > 2) {
            GemmArgs gb = g;
            gb.C = g.C + static_cast<long>(2 * EGX_NB) * ld + 2 * EGX_NB;
            gb.A = g.A + static_cast<long>(2 * EGX_NB) * LDP;
            gb.B = gb.A;
            gb.tri = tri2 - 2;
            gb.Mt = tri2 - 2 + Qt;
            gb.Nt = tri2 - 2;
            trailing_syrk(env, gb, sb, T, Rq ? Rq + static_cast<long>(3 * EGX_NB) * 4 : nullptr);
        }
        cudaEventRecord(env.ev_bulk[pair], sb);
    }
    if (la) {
        cudaEventRecord(env.ev_join, sp);
        cudaStreamWaitEvent(sb, env.ev_join, 0);
    }
}

int egx_fill_terms(int corr, int d, int h, const double* w, const double* theta, CorrTerm* t) {
    int nt = 0;
    if (corr == EGX_CORR_SQUARED_EXPONENTIAL || corr == EGX_CORR_ABSOLUTE_EXPONENTIAL) {
        for (int j = 0; j < d; ++j) {
            double s = 0.0;
            for (int l = 0; l < h; ++l) {
                if (corr == EGX_CORR_SQUARED_EXPONENTIAL) {
                    const double v = theta[l] * w[j * h + l];
                    s += v * v;
                } else {
                    s += std::fabs(w[j * h + l]) * theta[l];
                }
            }
            if (s != 0.0) {
                t[nt].dim = j;
                t[nt].pad_ = 0;
                t[nt].k1 = s;
                t[nt].k2 = 0.0;
                t[nt].k3 = 0.0;
                ++nt;
            }
        }
    } else {
        const double sq = (corr == EGX_CORR_MATERN32) ? std::sqrt(3.0) : std::sqrt(5.0);
        for (int j = 0; j < d; ++j)
            for (int l = 0; l < h; ++l) {
                const double tw = theta[l] * std::fabs(w[j * h + l]);
                if (tw != 0.0) {
                    t[nt].dim = j;
                    t[nt].pad_ = 0;
                    t[nt].k1 = tw;
                    t[nt].k2 = sq * tw;
                    t[nt].k3 = tw * tw;
                    ++nt;
                }
            }
    }
    return nt;
}

void backsolve_vector(SweepEnv& env, const FactorRef& f, double* v) {
    // EGX_BACKSOLVE_V=1: the r01 form (2 T one-CTA launches), kept for A/B
    static const int version = getenv("EGX_BACKSOLVE_V") != nullptr ? atoi(getenv("EGX_BACKSOLVE_V")) : 2;
    if (version != 1 && env.bs_flags != nullptr && f.T <= env.bs_flags_n && f.Dinv != nullptr) {
        StageScope sc(env.prof, EGX_STAGE_BACKSOLVE, 1, env.sb);
        launch_backsolve_chain(f.M, f.ld, f.Dinv, f.T, v, env.bs_flags, env.sb);
        return;
    }
    for (int k = f.T - 1; k >= 0; --k) {
        const double* Lkk = f.M + static_cast<long>(k) * EGX_NB * f.ld + static_cast<long>(k) * EGX_NB;
        StageScope sc(env.prof, EGX_STAGE_BACKSOLVE, k > 0 ? 2 : 1, env.sb);
        launch_backsolve_diag(Lkk, f.ld, v + k * EGX_NB, env.sb);
        if (k > 0) launch_backsolve_update(f.M + static_cast<long>(k) * EGX_NB * f.ld, f.ld, v + k * EGX_NB, v, k, env.sb);
    }
}

// Trajectories from a covariance on the device (gp/src/algorithm.rs:1153-1194, shared by the dense and the sparse GP):
// K (mpad x mpad, identity on the padding) is factorised in place -- Cholesky with the blocked sweep, or the host
// eigen-decomposition with eigenvalues below 1e-9 dropped -- and out (m x n_traj, host) = mean + C z.
int sample_from_covariance(SweepEnv& env, cudaStream_t s, double* K, int m, int mpad, const double* mean_dev, const double* z,
                           int n_traj, int method, double* out) {
    const int tpad = (n_traj + EGX_NB - 1) / EGX_NB * EGX_NB;
    struct Tmp {
        double *ZT = nullptr, *OUT = nullptr, *Dinv = nullptr;
        int* info = nullptr;
        ~Tmp() {
            egx_dev_free(ZT);
            egx_dev_free(OUT);
            egx_dev_free(Dinv);
            egx_dev_free(info);
        }
    } t;
    EGX_CUDA_TRY(egx_dev_malloc(&t.ZT, static_cast<size_t>(tpad) * mpad * sizeof(double)));
    EGX_CUDA_TRY(egx_dev_malloc(&t.OUT, static_cast<size_t>(mpad) * tpad * sizeof(double)));
    // the normal draws, one trajectory per row (the B operand of the NT product)
    std::vector<double> zt(static_cast<size_t>(tpad) * mpad, 0.0);
    for (int i = 0; i < m; ++i)
        for (int k = 0; k < n_traj; ++k) zt[static_cast<size_t>(k) * mpad + i] = z[static_cast<size_t>(i) * n_traj + k];
    EGX_CUDA_TRY(cudaMemcpyAsync(t.ZT, zt.data(), zt.size() * sizeof(double), cudaMemcpyHostToDevice, s));

    if (method == EGX_SAMPLE_CHOLESKY) {
        // cov = C C^T with the same blocked factorisation as the likelihood (algorithm.rs:1162-1168)
        EGX_CUDA_TRY(egx_dev_malloc(&t.Dinv, static_cast<size_t>(mpad / EGX_NB) * 4096 * sizeof(double)));
        EGX_CUDA_TRY(egx_dev_malloc(&t.info, sizeof(int)));
        EGX_CUDA_TRY(cudaMemsetAsync(t.info, 0, sizeof(int), s));
        FactorRef f;
        f.M = K;
        f.ld = mpad;
        f.T = mpad / EGX_NB;
        f.qpad = 0;
        f.Dinv = t.Dinv;
        f.info = t.info;
        blocked_sweep(env, f, true, nullptr, 0, 0, 0);
        int info_h = 0;
        EGX_CUDA_TRY(cudaMemcpyAsync(&info_h, t.info, sizeof(int), cudaMemcpyDeviceToHost, s));
        EGX_CUDA_TRY(cudaStreamSynchronize(s));
        if (info_h != 0) {
            egx_set_error("conditional covariance is not positive definite (pivot %d); use the eigenvalue method",
                          info_h);
            return EGX_NOT_POSITIVE_DEFINITE;
        }
        launch_zero_upper(K, mpad, mpad, s);
    } else {
        // C = W diag(sqrt(max(v, 0))) with eigenvalues below 1e-9 dropped (algorithm.rs:1169-1187); the m x m
        // eigen-decomposition runs on the host, as in the reference
        std::vector<double> a(static_cast<size_t>(m) * m), ev(m);
        EGX_CUDA_TRY(cudaMemcpy2DAsync(a.data(), static_cast<size_t>(m) * sizeof(double), K,
                                       static_cast<size_t>(mpad) * sizeof(double), static_cast<size_t>(m) * sizeof(double),
                                       m, cudaMemcpyDeviceToHost, s));
        EGX_CUDA_TRY(cudaStreamSynchronize(s));
        if (egx_host_symmetric_eig(m, a.data(), ev.data()) != 0) {
            egx_set_error("eigen-decomposition of the conditional covariance did not converge");
            return EGX_INVALID_VALUE;
        }
        for (int j = 0; j < m; ++j) ev[j] = (ev[j] < 1e-9) ? 0.0 : std::sqrt(ev[j]);
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) a[static_cast<size_t>(i) * m + j] *= ev[j];
        EGX_CUDA_TRY(cudaMemsetAsync(K, 0, static_cast<size_t>(mpad) * mpad * sizeof(double), s));
        EGX_CUDA_TRY(cudaMemcpy2DAsync(K, static_cast<size_t>(mpad) * sizeof(double), a.data(),
                                       static_cast<size_t>(m) * sizeof(double), static_cast<size_t>(m) * sizeof(double), m,
                                       cudaMemcpyHostToDevice, s));
        EGX_CUDA_TRY(cudaStreamSynchronize(s));      // `a` goes out of scope
    }
    // trajectories = mean + C Z   (algorithm.rs:1191-1193)
    launch_bcast_rows(t.OUT, tpad, m, mpad, tpad, mean_dev, s);
    {
        GemmArgs g;
        g.C = t.OUT;
        g.ldc = tpad;
        g.A = K;
        g.lda = mpad;
        g.B = t.ZT;
        g.ldb = mpad;
        g.K = mpad;
        g.add = 1;
        g.tri = 0;
        g.Mt = mpad / EGX_NB;
        g.Nt = tpad / EGX_NB;
        StageScope sc(env.prof, EGX_STAGE_SYRK_GEMM, 1, s);
        launch_gemm_nt_sub(g, s);
    }
    EGX_CUDA_TRY(cudaMemcpy2DAsync(out, static_cast<size_t>(n_traj) * sizeof(double), t.OUT,
                                   static_cast<size_t>(tpad) * sizeof(double), static_cast<size_t>(n_traj) * sizeof(double),
                                   m, cudaMemcpyDeviceToHost, s));
    EGX_CUDA_TRY(cudaStreamSynchronize(s));
    return EGX_OK;
}
