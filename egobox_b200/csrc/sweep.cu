// Implementation of the shared sweep environment (see sweep.cuh).
#include <cmath>
#include <cstdlib>

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
    ev_panel.assign(max_block_cols, nullptr);
    ev_bulk.assign(max_block_cols, nullptr);
    for (int k = 0; k < max_block_cols; ++k) {
        EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_panel[k], cudaEventDisableTiming));
        EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_bulk[k], cudaEventDisableTiming));
    }
    EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    EGX_CUDA_TRY(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    const char* e = getenv("EGX_LOOKAHEAD");
    lookahead = !(e != nullptr && atoi(e) == 0);
    if ((e = getenv("EGX_OZAKI")) != nullptr) ozaki = atoi(e);
    if ((e = getenv("EGX_OZAKI_MIN_TRI")) != nullptr) ozaki_min_tri = atoi(e) > 1 ? atoi(e) : 1;
    ozaki_min_tri_solve = ozaki_min_tri;
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
    oz_S = nullptr;
    oz_scale = nullptr;
    for (int i = 0; i < 2; ++i) {
        egx_dev_free(oz_rmaxq[i]);
        oz_rmaxq[i] = nullptr;
    }
    if (ozaki) {
        EGX_CUDA_TRY(egx_dev_malloc(&oz_S, ozaki_slice_bytes(rows)));
        EGX_CUDA_TRY(egx_dev_malloc(&oz_scale, static_cast<size_t>(rows) * sizeof(double)));
        for (int i = 0; i < 2; ++i) EGX_CUDA_TRY(egx_dev_malloc(&oz_rmaxq[i], static_cast<size_t>(rows) * 4 * sizeof(double)));
    }
    p_rows = rows;
    ++generation;
    return EGX_OK;
}

void SweepEnv::destroy() {
    if (sb) cudaStreamSynchronize(sb);
    if (sp) cudaStreamSynchronize(sp);
    prof.destroy();
    for (auto e : ev_panel)
        if (e) cudaEventDestroy(e);
    for (auto e : ev_bulk)
        if (e) cudaEventDestroy(e);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    egx_dev_free(P2[0]);
    egx_dev_free(P2[1]);
    egx_dev_free(oz_S);
    egx_dev_free(oz_scale);
    oz_S = nullptr;
    oz_scale = nullptr;
    for (int i = 0; i < 2; ++i) {
        egx_dev_free(oz_rmaxq[i]);
        oz_rmaxq[i] = nullptr;
    }
    if (sp) cudaStreamDestroy(sp);
    if (sb) cudaStreamDestroy(sb);
    sb = sp = nullptr;
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

void blocked_sweep(SweepEnv& env, const FactorRef& f, bool factor, double* rows, long ld_rows, int row_tiles_all,
                   int slabs64_all, bool upper_rows) {
    const int T = f.T, Qt = f.qpad / EGX_NB;
    const long ld = f.ld;
    const long LDP = 2 * EGX_NB;
    const bool la = env.lookahead && factor && T > 4 && static_cast<int>(env.ev_panel.size()) >= (T + 1) / 2;
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
        if (!two) break;
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
    for (int k = f.T - 1; k >= 0; --k) {
        const double* Lkk = f.M + static_cast<long>(k) * EGX_NB * f.ld + static_cast<long>(k) * EGX_NB;
        StageScope sc(env.prof, EGX_STAGE_BACKSOLVE, k > 0 ? 2 : 1, env.sb);
        launch_backsolve_diag(Lkk, f.ld, v + k * EGX_NB, env.sb);
        if (k > 0) launch_backsolve_update(f.M + static_cast<long>(k) * EGX_NB * f.ld, f.ld, v + k * EGX_NB, v, k, env.sb);
    }
}
