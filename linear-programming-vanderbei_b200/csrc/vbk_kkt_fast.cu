// vbk_kkt_fast.cu -- host orchestration of FAST mode (kernels in vbk_sparse_level.cuh, vbk_schur.cuh,
// vbk_dense_panel.cuh, vbk_dense_update.cuh, vbk_window_solve.cuh).
#include "vbk_kkt.h"
#include "vbk_strict_factor.cuh"
#include "vbk_schur.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace vbk {

void Kkt::prepare_fast()
{
    const int N = sym_.N, T = sym_.dense_start, W = N - T;
    fast_ready_ = false;
    sparse_tuned_ = 0;
    if (W <= 0) return;
    Sw_.alloc((size_t)W * W);
    P_.alloc((size_t)2 * W * kOuterPanel);          // two panels of L21*D: the look-ahead keeps two updates in flight
    dvec_.alloc(W); wmag_.alloc(W); wmark_.alloc(W);
    pan_d_.alloc(kPanelW); pan_keep_.alloc(kPanelW); panel_buf_.alloc(kPanelBufDoubles);
#ifndef VBK_EMU
    panel_buf2_.alloc((size_t)2 * kPanelBuf2Doubles);
#endif
    {
        const int npan = (W + kTriPW - 1) / kTriPW;
        tinv_.alloc((size_t)npan * kTriPW * kTriPW);
        tri_racc_.alloc((size_t)npan * kTriPW * kTriSplit);
        tri3_flags_.alloc((size_t)2 * npan + 1);
    }
    {
        // end of the sparse prefix (columns j < T) of every window row's ascending list
        std::vector<int> spend((size_t)W);
        for (int i = T; i < N; ++i) {
            const int* b = sym_.rj_asc.data() + sym_.rowptr[i];
            const int* e = sym_.rj_asc.data() + sym_.rowptr[i + 1];
            spend[(size_t)(i - T)] = sym_.rowptr[i] + (int)(std::lower_bound(b, e, T) - b);
        }
        sp_end_.upload(spend, stream_);
        // mean length of the tails the Schur assembly walks (entries of a sparse contributor below the window row)
        double tails = 0.0, contributors = 0.0;
        for (int i = T; i < N; ++i)
            for (int t = sym_.rowptr[i]; t < spend[(size_t)(i - T)]; ++t) {
                tails += sym_.kL[sym_.rj_asc[t] + 1] - sym_.rk_asc[t] - 1;
                contributors += 1.0;
            }
        schur_mean_tail_ = contributors > 0 ? tails / contributors : 0.0;
        // etree levels of the sparse columns (parent restricted to j < T): every level is one launch of
        // k_sparse_level for its light columns and one of k_sparse_level_heavy for the heavy ones
        // (vbk_sparse_level.cuh).  weight of a column = entries its contributors' tails apply to it
        std::vector<int> lev((size_t)std::max(T, 1), 0);
        std::vector<char> heavy((size_t)std::max(T, 1), 0);
        int nlev = 0, cap_l = 1, cap_h = 1;
        const long long heavy_w = std::getenv("VBK_SPARSE_HEAVY") ? std::atoll(std::getenv("VBK_SPARSE_HEAVY")) : 2048;
        for (int j = 0; j < T; ++j) {
            const int pj = sym_.parent[j];
            if (pj >= 0 && pj < T && lev[pj] < lev[j] + 1) lev[pj] = lev[j] + 1;
            nlev = std::max(nlev, lev[j] + 1);
            long long wgt = 0;
            for (int t = sym_.rowptr[j]; t < sym_.rowptr[j + 1]; ++t)
                wgt += sym_.kL[sym_.rj_asc[t] + 1] - sym_.rk_asc[t] - 1;
            heavy[j] = wgt > heavy_w;
            int& cap = heavy[j] ? cap_h : cap_l;
            cap = std::max(cap, sym_.kL[j + 1] - sym_.kL[j]);
        }
        // per level: light columns first, then heavy ones; sp_lvlptr_ has 2 entries per level + end
        sp_lvlptr_.assign((size_t)2 * nlev + 1, 0);
        for (int j = 0; j < T; ++j) ++sp_lvlptr_[(size_t)2 * lev[j] + (heavy[j] ? 1 : 0) + 1];
        for (size_t l = 0; l + 1 < sp_lvlptr_.size(); ++l) sp_lvlptr_[l + 1] += sp_lvlptr_[l];
        std::vector<int> cols((size_t)std::max(T, 1)), fill(sp_lvlptr_.begin(), sp_lvlptr_.end());
        for (int j = 0; j < T; ++j) cols[(size_t)fill[(size_t)2 * lev[j] + (heavy[j] ? 1 : 0)]++] = j;
        sp_lvlcol_.upload(cols, stream_);
        sp_cap_ = (cap_l + 1) & ~1;
        sp_cap_heavy_ = (cap_h + 1) & ~1;
    }
    VBK_CUDA(cudaFuncSetAttribute(k_sparse_level, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_));
    VBK_CUDA(cudaFuncSetAttribute(k_sparse_level_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_));
    VBK_CUDA(cudaFuncSetAttribute(k_schur_window2<kSchur2ThreadsSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_));
    VBK_CUDA(cudaFuncSetAttribute(k_schur_window2<kSchur2ThreadsLarge>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_));
    VBK_CUDA(cudaFuncSetAttribute(k_window_tinv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTinvSmem));
    VBK_CUDA(cudaFuncSetAttribute(k_window_tri3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTriV3Smem));
    VBK_CUDA(cudaFuncSetAttribute(k_dense_update_k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)(sizeof(double) * 2 * kPanelMax * kUpdTD)));
    VBK_CUDA(cudaFuncSetAttribute(k_panel_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPanelDiagSmem));
    VBK_CUDA(cudaFuncSetAttribute(k_panel_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPanelRowsSmem));
#ifndef VBK_EMU
    VBK_CUDA(cudaFuncSetAttribute(k_dense_update_m<128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UpdMma<128, 64>::kSmem));
    VBK_CUDA(cudaFuncSetAttribute(k_dense_update_m<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UpdMma<64, 64>::kSmem));
    VBK_CUDA(cudaFuncSetAttribute(k_panel_rows_m, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPanelRowsMSmem));
    if (!stream2_) {
        VBK_CUDA(cudaStreamCreateWithFlags(&stream2_, cudaStreamNonBlocking));
        for (int u = 0; u < 2; ++u) {
            VBK_CUDA(cudaEventCreateWithFlags(&ev_rows_[u], cudaEventDisableTiming));
            VBK_CUDA(cudaEventCreateWithFlags(&ev_updb_[u], cudaEventDisableTiming));
        }
    }
#endif
    fast_ready_ = true;
}

// Numeric factorisation in fast mode: sparse columns (bit-exact), Schur assembly, dense window, solve operands.
void Kkt::factor_window_fast()
{
    const int N = sym_.N, T = sym_.dense_start, W = N - T;
    const int sparse_tasks = sym_.tasks_ok ? sym_.col_task0[T] : 0;
    int launches = 2;

    // 1. columns j < T keep the reference's arithmetic.  Two kernels produce the same bits for them -- the strict slice
    //    tasks (vbk_strict_factor.cuh) and one launch per elimination-tree level (vbk_sparse_level.cuh) -- so the choice
    //    is a pure timing question and is made by measurement: the first factorisation of a handle runs the task
    //    kernel, the second the level kernels (CUDA events around this phase, one synchronisation each), every later
    //    one whichever was faster.  Deep, thin trees (pilot87: 99 levels) favour the task kernel's column-level
    //    dataflow, shallow wide ones (dfl001, multicommodity) the level kernels.  $VBK_SPARSE=strict|level pins it.
    const char* esp = std::getenv("VBK_SPARSE");
    const size_t sp_smem = (size_t)kSpWarps * sp_cap_ * (2 * sizeof(double) + sizeof(int));
    const size_t sph_smem = sizeof(double) * ((size_t)2 * sp_cap_heavy_ + 2 * kSpHeavyBatch + kSpHeavyThreads)
                            + sizeof(int) * ((size_t)sp_cap_heavy_ + W + 2 * kSpHeavyBatch + 2);
    const bool levels_ok = sp_smem <= (size_t)smem_optin_ && sph_smem <= (size_t)smem_optin_;
    bool sparse_levels = levels_ok;
    int tune_slot = -1;
    if (!sym_.tasks_ok) {
        if (!levels_ok) { std::fprintf(stderr, "vbkkt: neither the slice tasks nor the level kernels fit this LP\n"); std::exit(1); }
        sparse_levels = true;                 // no slice tasks for this LP (vbk_symbolic.cpp)
    }
    else if (esp && std::string(esp) == "strict") sparse_levels = false;
    else if (esp && std::string(esp) == "level") sparse_levels = levels_ok;
    else if (levels_ok && sparse_tasks > 0 && T > 0) {
        if (sparse_tuned_ < 2) { tune_slot = sparse_tuned_; sparse_levels = tune_slot == 1; }
        else sparse_levels = sparse_ms_[1] <= sparse_ms_[0];
    }
    if (tune_slot >= 0) {
        if (!ev_sp0_) { VBK_CUDA(cudaEventCreate(&ev_sp0_)); VBK_CUDA(cudaEventCreate(&ev_sp1_)); }
        VBK_CUDA(cudaEventRecord(ev_sp0_, stream_));
    }
    if (sparse_tasks > 0 && !sparse_levels) launch_factor_pipe(sparse_tasks, false);
    else VBK_LAUNCH(k_pipe_reset, vec_grid(N), kVecThreads, 0, stream_, N, col_pub_.p, col_done_.p, counters_.p);   // ndep = 0
    if (T > 0 && sparse_levels) {
        SparseLevelArgs sl;
        sl.n_ld = sym_.n; sl.T = T; sl.W = W; sl.kL = kL_.p; sl.iL = iL_.p; sl.L = L_.p; sl.diag = diag_.p; sl.mark = mark_.p;
        sl.perm = perm_.p; sl.rowptr = rowptr_.p; sl.rk = rk_sig_.p; sl.rj = rj_sig_.p; sl.counters = counters_.p;
        sl.scal_bits = bits_.p; sl.epsnum = 0.0;                     // _EPSNUM, ldlt.c:29
        for (size_t l = 0; l + 2 < sp_lvlptr_.size(); l += 2) {
            const int nl = sp_lvlptr_[l + 1] - sp_lvlptr_[l], nh = sp_lvlptr_[l + 2] - sp_lvlptr_[l + 1];
            if (nl > 0) {
                sl.cols = sp_lvlcol_.p + sp_lvlptr_[l]; sl.ncols = nl; sl.cap = sp_cap_;
                const int g = std::max(1, std::min((nl + kSpWarps - 1) / kSpWarps, num_sms_ * 8));
                VBK_LAUNCH(k_sparse_level, g, kSpWarps * 32, sp_smem, stream_, sl);
                ++launches;
            }
            if (nh > 0) {
                sl.cols = sp_lvlcol_.p + sp_lvlptr_[l + 1]; sl.ncols = nh; sl.cap = sp_cap_heavy_;
                VBK_LAUNCH(k_sparse_level_heavy, std::min(nh, num_sms_ * 2), kSpHeavyThreads, sph_smem, stream_, sl);
                ++launches;
            }
        }
    }
    if (tune_slot >= 0) {
        VBK_CUDA(cudaEventRecord(ev_sp1_, stream_));
        VBK_CUDA(cudaEventSynchronize(ev_sp1_));
        float ms = 0.f;
        VBK_CUDA(cudaEventElapsedTime(&ms, ev_sp0_, ev_sp1_));
        sparse_ms_[tune_slot] = ms;
        sparse_tuned_ = tune_slot + 1;
    }
    // 2. Schur complement of the sparse columns on the window, written densely; no dependencies.
    //    Entries outside the fill pattern are never written: start from zero.
    VBK_CUDA(cudaMemsetAsync(Sw_.p, 0, sizeof(double) * (size_t)W * W, stream_));
    {
        // long tails (large multicommodity LPs: ~1100 entries per contributor on average) want many threads per CTA, short
        // ones few (measured: R=50/K=40 961 -> 480 ms per factorisation with 1024 threads, R=32/K=25 and dfl001 ~10 % slower); $VBK_SCHUR_THREADS=small|large pins the choice
        bool large = schur_mean_tail_ >= 768.0;
        if (const char* e = std::getenv("VBK_SCHUR_THREADS")) large = std::string(e) == "large";
        const int nt = large ? kSchur2ThreadsLarge : kSchur2ThreadsSmall;
        size_t sc2_smem = sizeof(double) * ((size_t)W + 2 * nt) + sizeof(int) * 2 * nt;
        if (sc2_smem > (size_t)smem_optin_ && large) {
            large = false;
            sc2_smem = sizeof(double) * ((size_t)W + 2 * kSchur2ThreadsSmall) + sizeof(int) * 2 * kSchur2ThreadsSmall;
        }
        if (sc2_smem > (size_t)smem_optin_) {
            std::fprintf(stderr, "vbkkt: dense window of %d columns exceeds the Schur assembly's shared memory\n", W);
            std::exit(1);
        }
        Schur2Args sc;
        sc.N = N; sc.T = T; sc.ld = W; sc.cap = W; sc.kL = kL_.p; sc.iL = iL_.p; sc.L = L_.p; sc.diag = diag_.p;
        sc.rowptr = rowptr_.p; sc.rk = rk_asc_.p; sc.rj = rj_asc_.p; sc.spend = sp_end_.p; sc.S = Sw_.p; sc.wmag = wmag_.p;
        if (large) VBK_LAUNCH(k_schur_window2<kSchur2ThreadsLarge>, std::min(W, num_sms_ * 6), kSchur2ThreadsLarge, sc2_smem, stream_, sc);
        else       VBK_LAUNCH(k_schur_window2<kSchur2ThreadsSmall>, std::min(W, num_sms_ * 6), kSchur2ThreadsSmall, sc2_smem, stream_, sc);
    }

    // 3. blocked right-looking dense LDL^T of the window: 128-column panels (vbk_dense_panel.cuh) -- diagonal block,
    //    rows below, rank-128 trailing update on the FP64 tensor path (vbk_dense_update.cuh)
    DenseArgs da;
    da.W = W; da.ld = W; da.S = Sw_.p; da.P = P_.p; da.dvec = dvec_.p; da.wmag = wmag_.p; da.wmark = wmark_.p;
    da.pan_d = pan_d_.p; da.pan_keep = pan_keep_.p; da.PB = panel_buf_.p; da.prof = nullptr; da.perm = perm_.p; da.T = T; da.n_ld = sym_.n;
    da.counters = counters_.p;
    // The reference's "exactly zero" pivots come from absorption: a term T swallows the running value
    // (|value| < ulp(T)/2) and is then cancelled exactly.  With re-associated sums the same pivot comes
    // out as the tiny true value instead, so the test is |d| <= 2^-52 * (largest term magnitude).
    da.tol = 2.220446049250313e-16;
    if (const char* e = std::getenv("VBK_PIVOT_TOL_ULPS")) da.tol *= std::max(0.0, std::atof(e));
    // A window pivot that fails the test above is rounding noise of terms of magnitude wmag.  The reference
    // substitutes sgn*1e-8 whatever the scale (ldlt.c:612); with noise-level pivots flagged too, that choice
    // multiplies the column by up to 1e8*wmag and overflowed on the random m=2000 LPs of BASELINE config 4
    // (profiles/r01_summary.md).  Fast mode substitutes sgn*max(1e-8, sqrt(eps)*wmag) instead ("static
    // pivoting"; iterative refinement absorbs the perturbation).  $VBK_PIVOT_STATIC=0 restores the literal rule.
    da.piv_scale = 1.4901161193847656e-08;
    if (const char* e = std::getenv("VBK_PIVOT_STATIC")) da.piv_scale = std::atof(e);
    {
        static const bool prof_on = std::getenv("VBK_PROF") != nullptr;
        if (prof_on) {
            panel_prof_.alloc(16);
            VBK_CUDA(cudaMemsetAsync(panel_prof_.p, 0, 16 * sizeof(unsigned long long), stream_));
            da.prof = panel_prof_.p;
        }
        // Look-ahead of depth one.  Panel k's rank-128 update is split: the columns of panel k+1 ("A part", few
        // tiles) stay on the main stream, everything to the right of them ("B part", nearly all the flops) goes to
        // a second stream and overlaps with the factorisation of panel k+1, which is a latency-bound chain on one or
        // a few SMs.  Hazards: B_k reads P_k and the columns of panel k and writes strictly-lower entries right of
        // panel k+1 only -- nothing panel k+1's kernels touch (they write their own columns, the other P buffer and
        // diagonal entries); A_{k+1} and rows_{k+2} (which reuses P_k's buffer) wait for B_k.
#ifdef VBK_EMU
        cudaStream_t sB = stream_;               // same split of the update, one (emulated) stream
        auto launch_update = [&](int tiles, cudaStream_t st, const DenseArgs& d) {
            VBK_LAUNCH(k_dense_update_k, dim3(tiles, tiles), kUpdThreads, sizeof(double) * 2 * kPanelMax * kUpdTD, st, d);
        };
#else
        cudaStream_t sB = stream2_;
        auto launch_update = [&](int tiles, cudaStream_t st, const DenseArgs& d) {
            VBK_LAUNCH((k_dense_update_m<128, 64>), dim3(2 * tiles, tiles), (UpdMma<128, 64>::kThreads), (UpdMma<128, 64>::kSmem), st, d);
        };
        da.PB2 = panel_buf2_.p;
#endif
        int k = 0, last_b = -1;
        for (int P0 = 0; P0 < W; P0 += kPanelW, ++k) {
            da.p = P0; da.nb = std::min(kPanelW, W - P0); da.pcol0 = 0;
            da.P = P_.p + (size_t)(k & 1) * W * kOuterPanel;
            VBK_LAUNCH(k_panel_diag, 1, kDiagThreads, kPanelDiagSmem, stream_, da);
            ++launches;
            const int below = W - P0 - da.nb;
            if (below <= 0) continue;
#ifndef VBK_EMU
            {
                const int gm = std::min((below + 16 * kRowsMWarps - 1) / (16 * kRowsMWarps), num_sms_ * 2);
                VBK_LAUNCH(k_panel_rows_m, gm, kRowsMWarps * 32, kPanelRowsMSmem, stream_, da);
            }
#else
            {
                const int g = std::min((below + kRowsPerCta - 1) / kRowsPerCta, num_sms_ * 4);
                VBK_LAUNCH(k_panel_rows, g, kRowThreads, kPanelRowsSmem, stream_, da);
            }
#endif
            ++launches;
            const int kend = P0 + da.nb;
            da.kcol0 = P0; da.klen = da.nb;
            const int rest = W - (kend + kPanelW);                  // columns right of panel k+1
            if (rest > 0) {                                         // B part on the second stream
#ifndef VBK_EMU
                VBK_CUDA(cudaEventRecord(ev_rows_[k & 1], stream_));
                VBK_CUDA(cudaStreamWaitEvent(sB, ev_rows_[k & 1], 0));
#endif
                DenseArgs db = da;
                db.rbase = kend + kPanelW; db.cmax = W;
                launch_update((rest + kUpdTD - 1) / kUpdTD, sB, db);
#ifndef VBK_EMU
                VBK_CUDA(cudaEventRecord(ev_updb_[k & 1], sB));
#endif
                ++launches;
            }
            // A part: columns of panel k+1, all rows below panel k.  Those columns were last written by B_{k-1}.
#ifndef VBK_EMU
            if (last_b >= 0) VBK_CUDA(cudaStreamWaitEvent(stream_, ev_updb_[last_b & 1], 0));
#endif
            last_b = rest > 0 ? k : -1;
            da.rbase = kend; da.cmax = std::min(kend + kPanelW, W);
            const int tr = (below + kStripTD - 1) / kStripTD, tc = (da.cmax - kend + kStripTD - 1) / kStripTD;
#ifndef VBK_EMU
            VBK_LAUNCH((k_dense_update_m<64, 64>), dim3(tc, tr), (UpdMma<64, 64>::kThreads), (UpdMma<64, 64>::kSmem), stream_, da);
#else
            VBK_LAUNCH(k_dense_update_strip, dim3(tc, tr), kStripThreads, sizeof(double) * 2 * kPanelMax * kStripTD, stream_, da);
#endif
            ++launches;
        }
#ifndef VBK_EMU
        if (last_b >= 0) VBK_CUDA(cudaStreamWaitEvent(stream_, ev_updb_[last_b & 1], 0));
#endif
    }
    if (da.prof) {
        unsigned long long h[16];
        panel_prof_.download(h, 16, stream_);
        VBK_CUDA(cudaStreamSynchronize(stream_));
        std::fprintf(stderr, "vbkkt panel profile (cycles of thread 0, CTA 0, summed over %d panels): diag load %llu, warp LDL %llu, "
                     "(%llu sub-blocks, %llu repeated with the full pivot rule) block trsm %llu, block update %llu, store %llu | rows: panel fetch %llu, row loads %llu, rank update %llu, "
                     "stages %llu, stores %llu\n", (W + kPanelW - 1) / kPanelW, h[0], h[1], h[14], h[15], h[2], h[3], h[4], h[8], h[9], h[10], h[11], h[12]);
    }
    // 4. mirror L into the upper triangle (the backward sweep then reads rows, like the forward one)
    {
        const int nt32 = (W + 31) / 32;
        VBK_LAUNCH(k_window_mirror, dim3(nt32, nt32), kVecThreads, 32 * 33 * sizeof(double), stream_, W, W, Sw_.p);
        ++launches;
    }
    // 4b. inverses of the 128 x 128 diagonal blocks for the triangular sweeps (vbk_window_solve.cuh)
    VBK_LAUNCH(k_window_tinv, (W + kTriPW - 1) / kTriPW, kTriPW, kTinvSmem, stream_, W, W, Sw_.p, tinv_.p);
    ++launches;
    // 5. back into the packed storage the strict-layout consumers (tests, get_factor) read
    {
        const int gx = std::max(1, std::min((W + kVecThreads - 1) / kVecThreads, 64));
        const int gy = std::max(1, std::min(W, 1024));
        VBK_LAUNCH(k_window_store, dim3(gx, gy), kVecThreads, 0, stream_, W, T, W, Sw_.p, dvec_.p, wmark_.p, kL_.p,
                   iL_.p, L_.p, diag_.p, mark_.p);
        ++launches;
    }
    stats.kernel_launches += launches;
}

void Kkt::rawsolve_window_fast(FlagSolveArgs& fs, SolveArgs& sa, size_t flag_smem)
{
    const int N = sym_.N, T = sym_.dense_start, W = N - T;
    WindowSolveArgs wa;
    wa.N = N; wa.T = T; wa.ld = W; wa.S = Sw_.p; wa.kL = kL_.p; wa.L = L_.p; wa.mark = mark_.p;
    wa.rowptr = rowptr_.p; wa.rk = rk_asc_.p; wa.rj = rj_asc_.p; wa.z = z_.p; wa.spend = sp_end_.p;
    wa.counters = counters_.p; wa.scal_bits = bits_.p; wa.epssol = 1.0e-6;
    fs.nclaim = T;
    fs.fast = 1;
    const int gsolve = std::max(1, std::min(solve_grid_, (T + 3) / 4));
    const int ggather = std::max(1, std::min(num_sms_ * 4, (W * 32 + kSolveThreads - 1) / kSolveThreads));

    VBK_LAUNCH(k_flags_reset, vec_grid(N), kVecThreads, 0, stream_, N, done_.p, counters_.p, 1);
    if (T > 0) VBK_LAUNCH(k_fwd_flags<1>, gsolve, kSolveThreads, flag_smem, stream_, fs);
    VBK_LAUNCH(k_window_gather, ggather, kSolveThreads, 0, stream_, wa);
    // window: 128-row panels with inverted diagonal blocks (vbk_window_solve.cuh)
    Tri3Args t3;
    t3.W = W; t3.ld = W; t3.npan = (W + kTriPW - 1) / kTriPW; t3.S = Sw_.p; t3.Tinv = tinv_.p; t3.z = z_.p + T;
    t3.mark = mark_.p + T; t3.racc = tri_racc_.p; t3.flags = tri3_flags_.p; t3.counters = counters_.p;
    t3.scal_bits = bits_.p; t3.epssol = 1.0e-6;
    const int g3 = std::max(1, std::min(t3.npan * kTriSplit, num_sms_));
    auto sweep3 = [&](int dir) {
        VBK_CUDA(cudaMemsetAsync(tri3_flags_.p, 0, sizeof(int) * (size_t)(2 * t3.npan + 1), stream_));
        VBK_CUDA(cudaMemsetAsync(tri_racc_.p, 0, sizeof(double) * (size_t)t3.npan * kTriPW * kTriSplit, stream_));
        t3.dir = dir;
        VBK_LAUNCH(k_window_tri3, g3, kTriV3Threads, kTriV3Smem, stream_, t3);
    };
    sweep3(0);
    VBK_LAUNCH(k_diag_strict, vec_grid(N), kVecThreads, 0, stream_, sa);
    sweep3(1);
    VBK_LAUNCH(k_flags_reset, vec_grid(N), kVecThreads, 0, stream_, N, done_.p, counters_.p, 0);
    if (T > 0) VBK_LAUNCH(k_bwd_flags, gsolve, kSolveThreads, flag_smem, stream_, fs);
    VBK_CHECK_LAUNCH();
    stats.kernel_launches += 10;
}

}  // namespace vbk
