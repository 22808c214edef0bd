// vbk_fast3.cuh -- third-generation dense-window factorisation (fast mode): 128-column panels.
//
// Measured on B200 with the kernels of vbk_fast2.cuh (dfl001, padded window W=4277): 13.5 of the 16.9 ms
// of a factorisation are the 134 x (k_dense_diag_w, k_dense_trsm_u, strip update) launches of the
// 32-column panels -- a chain of ~400 small dependent kernels, ~100 us per panel, while the rank-128
// trailing updates that hold nearly all the flops take ~1.5 ms.  Here a panel is 128 columns and costs
// three launches in all:
//   k_panel_diag  ONE CTA factorises the 128 x 128 diagonal block in shared memory: four 32-column
//                 sub-blocks, each = warp-level LDL^T in registers (warp 0), row-parallel substitution for
//                 the block rows below it, rank-32 update of the rest of the block (all warps);
//   k_panel_rows  every row below the block is owned by FOUR threads (a quad: 8 of the 32 columns of a sub-block
//                 each) for the whole panel; L11^T, the reciprocal pivots and the keep flags arrive in shared
//                 memory as ONE bulk asynchronous copy (cp.async.bulk + mbarrier) of the packed panel buffer
//                 the diagonal kernel leaves behind; per sub-block a rank-(32b) update from the row's own
//                 earlier results, then four 8-column substitution stages chained by quad shuffles;
//   k_dense_update_k  (vbk_fast2.cuh) the rank-128 update of the trailing matrix.
// The dependent-pivot rule (reference ldlt.c:600-614) keeps its meaning: when a pivot is "zero" the
// maximum of the updated column below it decides between dropping the row and substituting the pivot;
// rows whose panel columns have not been touched yet get them applied on the fly by the whole CTA
// (rare path, O(rows * c^2)).
#pragma once
#include "vbk_fast2.cuh"

namespace vbk {

#ifdef VBK_EMU
constexpr int kPanelW = 64;                   // emulated build: small panels, so that tiny windows span several
#else
constexpr int kPanelW = kOuterPanel;          // 128 columns per panel
#endif
constexpr int kLDD = kPanelW + 1;             // row stride of the diagonal block in k_panel_diag (odd: conflict-free)
constexpr int kLDT = kPanelW + 2;             // column stride of L11 in k_panel_rows (even: 16-byte aligned rows)
#ifdef VBK_EMU
constexpr int kDiagThreads = 64;
constexpr int kRowsPerCta = 8;
#else
constexpr int kDiagThreads = 256;
// 32 rows = 4 warps per CTA: k_panel_rows is bound by shared-memory bandwidth (every DFMA takes a 16-byte operand
// that only a quarter-warp shares), so the rows are spread over as many SMs as possible (measured: 64 rows per
// CTA 45 us, see profiles/)
constexpr int kRowsPerCta = 32;
#endif
constexpr int kRowThreads = 4 * kRowsPerCta;
// packed panel buffer written by k_panel_diag, read by k_panel_rows: L11 transposed (element (row, col) at
// [col * kLDT + row], zero on and above the diagonal and in the padding), reciprocal pivots, keep flags (0/1)
constexpr int kPanelBufDoubles = kPanelW * kLDT + 2 * kPanelW;
constexpr size_t kPanelDiagSmem = sizeof(double) * (kPanelW * kLDD + (kPanelW - 32) * 33 + 3 * kPanelW + kDiagThreads)
                                  + sizeof(int) * (kPanelW + 4);
constexpr size_t kPanelRowsSmem = sizeof(double) * (kPanelBufDoubles + (kPanelW - 32) * kRowsPerCta) + 16;

// correctly rounded reciprocal: one MUFU seed + Newton steps, about half the dependent chain of a full division
__device__ __forceinline__ double vbk_rcp(double d) {
#ifdef VBK_EMU
    return 1.0 / d;
#else
    return __drcp_rn(d);
#endif
}

// $VBK_PROF: thread 0 of CTA 0 adds the cycles since *t to slot and restarts the clock
__device__ __forceinline__ void panel_tick(const DenseArgs& a, int slot, long long* t)
{
    if (a.prof && threadIdx.x == 0 && blockIdx.x == 0) {
        const long long now = vbk_clock();
        atomicAdd(&a.prof[slot], (unsigned long long)(now - *t));
        *t = now;
    }
}

// One row against a unit-lower 32 x 32 block (element (c, c0) at L[c * rs + c0 * cs]): on entry l holds the row of
// the Schur complement, on exit l = row of L and w = l * d (= the substituted value itself).  Right-looking, so
// that the 31 - c updates after column c are independent; pivots enter as reciprocals (one multiply on the
// critical path instead of a division).  Columns whose keep flag is 0 (dropped rows, padding) give l = w = 0.
__device__ __forceinline__ void trsm32(double (&l)[32], double (&w)[32], const double* L, int rs, int cs,
                                       const double* invd, const int* keepp)
{
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const double sc = keepp[c] ? l[c] : 0.0;
        w[c] = sc;
        l[c] = sc * invd[c];
#pragma unroll
        for (int c2 = c + 1; c2 < 32; ++c2) l[c2] = fma(-sc, L[c2 * rs + c * cs], l[c2]);
    }
}

// rare path: max |column cabs of the Schur complement| over the rows below the current sub-block
// [b0, b0+nbb); called by every thread of the CTA, rows dealt round-robin.  Columns of earlier sub-blocks hold
// L in blk; the columns of the current sub-block still hold L*d (they are scaled when the sub-block is done),
// hence the sinv factor.
#ifndef VBK_EMU
__noinline__
#endif
__device__ double panel_colmax(const DenseArgs& a, int b0, int nbb, int cabs, int tid, int nt,
                               const double* blk, const double* sd, const double* sinv, const int* skeep)
{
    const int p = a.p, nb = a.nb;
    double mymax = 0.0;
    const int k = cabs - b0;
    // rows of the diagonal block: final through the earlier sub-blocks, the current one still to apply
    for (int r = b0 + nbb + tid; r < nb; r += nt) {
        double l[32];
        for (int c1 = 0; c1 < k; ++c1) {
            double s = blk[r * kLDD + b0 + c1];
            for (int c0 = 0; c0 < c1; ++c0) s = fma(-l[c0], blk[(b0 + c1) * kLDD + b0 + c0], s);     // l*d*L = l * (L*d)
            l[c1] = skeep[b0 + c1] ? s * sinv[b0 + c1] : 0.0;
        }
        double v = blk[r * kLDD + cabs];
        for (int c0 = 0; c0 < k; ++c0) v = fma(-l[c0], blk[cabs * kLDD + b0 + c0], v);
        if (fabs(v) > mymax) mymax = fabs(v);
    }
    // rows below the block: nothing of this panel has been applied to them yet
    for (int r = nb + tid; r < a.W - p; r += nt) {
        double l[kPanelW];
        for (int c1 = 0; c1 < cabs; ++c1) {
            double s = SW(a, p + r, p + c1);
            for (int c0 = 0; c0 < c1; ++c0) {
                const double ld = (c0 >= b0) ? blk[c1 * kLDD + c0] : blk[c1 * kLDD + c0] * sd[c0];  // L[c1][c0] * d[c0]
                s = fma(-l[c0], ld, s);
            }
            l[c1] = skeep[c1] ? s * sinv[c1] : 0.0;
        }
        double v = SW(a, p + r, p + cabs);
        for (int c0 = 0; c0 < cabs; ++c0) {
            const double ld = (c0 >= b0) ? blk[cabs * kLDD + c0] : blk[cabs * kLDD + c0] * sd[c0];
            v = fma(-l[c0], ld, v);
        }
        if (fabs(v) > mymax) mymax = fabs(v);
    }
    return mymax;
}

static __global__ void __launch_bounds__(kDiagThreads) k_panel_diag(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* blk = reinterpret_cast<double*>(raw);             // [kPanelW][kLDD] row-major diagonal block
    double* wbuf = blk + kPanelW * kLDD;                      // [kPanelW-32][33]  L*D of the rows below a sub-block
    double* sd = wbuf + (kPanelW - 32) * 33;                  // [kPanelW] pivots
    double* wm = sd + kPanelW;                                // [kPanelW] largest term magnitude of each diagonal entry
    double* sinv = wm + kPanelW;                              // [kPanelW] reciprocal pivots (0 for dropped rows)
    double* red = sinv + kPanelW;                             // [kDiagThreads]
    int* skeep = reinterpret_cast<int*>(red + kDiagThreads);  // [kPanelW]
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int p = a.p, nb = a.nb;
    long long tk = vbk_clock();

    if (nb < kPanelW)                                              // partial panel: the padding must read as zero
        for (int e = tid; e < kPanelW * kLDD; e += nt) blk[e] = 0.0;
    for (int e = tid; e < kPanelW; e += nt) { sd[e] = 1.0; sinv[e] = 0.0; skeep[e] = 0; wm[e] = (e < nb) ? a.wmag[p + e] : 0.0; }
    __syncthreads();
#pragma unroll 8
    for (int e = tid; e < kPanelW * kPanelW; e += nt) {            // consecutive threads = consecutive rows: coalesced
        const int r = e % kPanelW, c = e / kPanelW;                // kPanelW is a power of two
        if (r >= c && r < nb) blk[r * kLDD + c] = SW(a, p + r, p + c);
    }
    __syncthreads();
    panel_tick(a, 0, &tk);

    for (int b0 = 0; b0 < nb; b0 += 32) {
        const int nbb = (nb - b0 < 32) ? (nb - b0) : 32;
        // ---- (a) the 32 x 32 sub-block, right-looking, by the whole CTA: per column one barrier, the pivot and
        // its reciprocal (every thread for itself), then <= 4 independent updates per thread.  The columns keep
        // a = l*d until the sub-block is finished (nobody reads a finished column again), so no thread ever
        // waits for a scaled column: the dependent chain of a column is barrier + load + reciprocal + 2 multiplies
        // (a warp-level left-looking version measured 1000+ cycles per column, profiles/).
        for (int c = 0; c < nbb; ++c) {
            __syncthreads();
            double d = blk[(b0 + c) * kLDD + b0 + c];
            const double magc = wm[b0 + c];
            int keep = 1;
            if (fabs(d) <= a.tol * magc) {                                     // uniform over the CTA; ldlt.c:600-614
                double mymax = 0.0;
                for (int r = c + 1 + tid; r < nbb; r += nt) mymax = fmax(mymax, fabs(blk[(b0 + r) * kLDD + b0 + c]));
                const double below = panel_colmax(a, b0, nbb, b0 + c, tid, nt, blk, sd, sinv, skeep);
                red[tid] = fmax(mymax, below);
                __syncthreads();
                double m = 0.0;
                for (int u = 0; u < nt; ++u) m = fmax(m, red[u]);
                __syncthreads();
                if (m < 1.0e+6 * 1.0e-8) keep = 0;
                else {
                    double sub = a.piv_scale * magc;
                    if (!(sub > 1.0e-8)) sub = 1.0e-8;
                    d = (a.perm[a.T + p + b0 + c] < a.n_ld ? -1 : 1) * sub;
                }
                if (tid == 0) atomicAdd(&a.counters[C_NDEP], 1);
            }
            const double inv = keep ? vbk_rcp(d) : 0.0;
            if (tid == 0) { sd[b0 + c] = d; sinv[b0 + c] = inv; skeep[b0 + c] = keep; }
            const int c2 = lane;
            if (c2 > c && c2 < nbb) {
                const double ac2 = blk[(b0 + c2) * kLDD + b0 + c] * inv;        // l_{c2,c}
                // warp w owns rows w, w + nwarps, ...: all loads first, then the arithmetic, then the stores (a
                // load-update-store loop is serialised by the compiler: the store may alias the next load)
                constexpr int kRpw = 32 / (kDiagThreads / 32);
                double ar[kRpw], tv[kRpw];
#pragma unroll
                for (int i = 0; i < kRpw; ++i) {
                    const int r = warp + i * (kDiagThreads / 32);
                    const bool on = r >= c2 && r < nbb;
                    ar[i] = on ? blk[(b0 + r) * kLDD + b0 + c] : 0.0;           // a_{r,c}
                    tv[i] = on ? blk[(b0 + r) * kLDD + b0 + c2] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < kRpw; ++i) {
                    const int r = warp + i * (kDiagThreads / 32);
                    const double upd = ar[i] * ac2;                             // a_{r,c} * l_{c2,c} = l d l
                    tv[i] -= upd;
                    if (r == c2 && fabs(upd) > wm[b0 + r]) wm[b0 + r] = fabs(upd);
                }
#pragma unroll
                for (int i = 0; i < kRpw; ++i) {
                    const int r = warp + i * (kDiagThreads / 32);
                    if (r >= c2 && r < nbb) blk[(b0 + r) * kLDD + b0 + c2] = tv[i];
                }
            }
        }
        __syncthreads();
        // finished: a = l*d  ->  l   (dropped columns: inv = 0 gives l = 0)
        for (int e = tid; e < 32 * 32; e += nt) {
            const int r = e >> 5, c = e & 31;
            if (c < r && r < nbb) blk[(b0 + r) * kLDD + b0 + c] *= sinv[b0 + c];
        }
        __syncthreads();
        panel_tick(a, 1, &tk);
        const int rem = nb - b0 - nbb;
        if (rem <= 0) break;                                                   // uniform
        // ---- (b) block rows below the sub-block: substitution, one thread per row
        for (int t = tid; t < rem; t += nt) {
            const int r = b0 + nbb + t;
            double l[32], w[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) l[c] = (c < nbb) ? blk[r * kLDD + b0 + c] : 0.0;
            trsm32(l, w, blk + b0 * kLDD + b0, kLDD, 1, sinv + b0, skeep + b0);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                if (c < nbb) { blk[r * kLDD + b0 + c] = l[c]; wbuf[t * 33 + c] = w[c]; }
            }
        }
        __syncthreads();
        panel_tick(a, 2, &tk);
        // ---- (c) rank-nbb update of the rest of the block (lower triangle incl. diagonal)
        // 4 x 4 register tiles (16 FMAs per 8 shared-memory loads), rows and columns of a tile INTERLEAVED
        // (tr + i*nt4, tc + j*nt4) so that consecutive threads touch consecutive rows: stride-129 / stride-33
        // accesses are conflict-free, blocked tiles (stride 4*129) were 8-way conflicted
        {
            const int nt4 = (rem + 3) >> 2;
            for (int e = tid; e < nt4 * nt4; e += nt) {
                const int tr = e % nt4, tc = e / nt4;
                double acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
                double dmag[4] = {0.0, 0.0, 0.0, 0.0};
                for (int c = 0; c < nbb; ++c) {
                    double lv[4], wv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        lv[i] = (tr + i * nt4 < rem) ? blk[(b0 + nbb + tr + i * nt4) * kLDD + b0 + c] : 0.0;
                        wv[i] = (tc + i * nt4 < rem) ? wbuf[(tc + i * nt4) * 33 + c] : 0.0;
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fma(lv[i], wv[j], acc[i][j]);
                    if (tr == tc) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) dmag[i] = fmax(dmag[i], fabs(lv[i] * wv[i]));
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int rr = tr + i * nt4, cc = tc + j * nt4;
                        if (rr < rem && cc < rem && rr >= cc) blk[(b0 + nbb + rr) * kLDD + b0 + nbb + cc] -= acc[i][j];
                    }
                if (tr == tc) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = tr + i * nt4;
                        if (rr < rem && dmag[i] > wm[b0 + nbb + rr]) wm[b0 + nbb + rr] = dmag[i];
                    }
                }
            }
        }
        __syncthreads();
        panel_tick(a, 3, &tk);
    }
    __syncthreads();
    for (int c = warp; c < kPanelW; c += (nt >> 5)) {
        for (int r = lane; r < kPanelW; r += 32) {
            const double v = (r > c && r < nb) ? blk[r * kLDD + c] : 0.0;
            if (r > c && r < nb) SW(a, p + r, p + c) = v;
            a.PB[c * kLDT + r] = v;                                 // packed copy for k_panel_rows
        }
        if (lane < kLDT - kPanelW) a.PB[c * kLDT + kPanelW + lane] = 0.0;
    }
    for (int e = tid; e < kPanelW; e += nt) {
        a.PB[kPanelW * kLDT + e] = sinv[e];
        a.PB[kPanelW * kLDT + kPanelW + e] = skeep[e] ? 1.0 : 0.0;
        if (e < nb) {
            a.dvec[p + e] = sd[e];
            a.wmark[p + e] = skeep[e];
            a.wmag[p + e] = wm[e];
            a.pan_d[e] = sd[e];
            a.pan_keep[e] = skeep[e];
        }
    }
    panel_tick(a, 4, &tk);
}

// rows below the diagonal block of the panel: L21 = S21 L11^{-T} D^{-1}, P = L21 D, trailing diagonal
static __global__ void __launch_bounds__(kRowThreads) k_panel_rows(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* pb = reinterpret_cast<double*>(raw);              // packed panel buffer, see kPanelBufDoubles
    const double* l11 = pb;                                   // element (row, col) at l11[col * kLDT + row]
    const double* sinv = pb + kPanelW * kLDT;
    const double* keepd = sinv + kPanelW;
    double* wsh = pb + kPanelBufDoubles;                      // [kPanelW-32][kRowsPerCta] L*D of the earlier sub-blocks
    const int tid = threadIdx.x, lane = tid & 31, nb = a.nb, p = a.p;
    const int rloc = tid >> 2, q = tid & 3;
    long long tk = vbk_clock();
#ifdef VBK_EMU
    for (int e = tid; e < kPanelBufDoubles; e += blockDim.x) pb[e] = a.PB[e];
    __syncthreads();
#else
    {
        // one bulk asynchronous copy (TMA engine) brings the whole packed panel; everyone waits on the mbarrier
        unsigned long long* mbar = reinterpret_cast<unsigned long long*>(wsh + (kPanelW - 32) * kRowsPerCta);
        const unsigned mbar_s = (unsigned)__cvta_generic_to_shared(mbar);
        const unsigned dst_s = (unsigned)__cvta_generic_to_shared(pb);
        constexpr unsigned kBytes = (unsigned)(sizeof(double) * kPanelBufDoubles);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(kBytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst_s), "l"(a.PB), "r"(kBytes), "r"(mbar_s) : "memory");
        }
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(mbar_s), "r"(0u) : "memory");
        }
    }
#endif
    panel_tick(a, 8, &tk);
    const int nsub = (nb + 31) / 32;
    const int nblocks = (a.W - p - nb + kRowsPerCta - 1) / kRowsPerCta;
    for (int blk_i = blockIdx.x; blk_i < nblocks; blk_i += gridDim.x) {
        const int r = p + nb + blk_i * kRowsPerCta + rloc;
        const bool valid = r < a.W;
        double dsum = 0.0, dabs = 0.0;
        for (int b = 0; b < nsub; ++b) {
            const int b0 = 32 * b, cq = b0 + 8 * q;
            double l[8], wv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { l[j] = (valid && cq + j < nb) ? SW(a, r, p + cq + j) : 0.0; wv[j] = 0.0; }
            if (l[0] == 123.456) dsum += 1.0;                  // (keeps the loads ahead of the tick below)
            panel_tick(a, 9, &tk);
            // the earlier sub-blocks of this panel: l[j] -= sum_{c0 < b0} w[c0] * L11[cq + j][c0]
#pragma unroll 4
            for (int c0 = 0; c0 < b0; ++c0) {
                const double nwv = -wsh[c0 * kRowsPerCta + rloc];
                const double* col = l11 + c0 * kLDT + cq;
#pragma unroll
                for (int j = 0; j < 8; ++j) l[j] = fma(nwv, col[j], l[j]);
            }
            panel_tick(a, 10, &tk);
            // four 8-column stages: quad member s substitutes through its own 8 x 8 triangle, then hands its
            // eight w = l*d to the members on its right
#pragma unroll
            for (int s8 = 0; s8 < 4; ++s8) {
                if (q == s8) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double sc = (keepd[cq + j] != 0.0) ? l[j] : 0.0;
                        wv[j] = sc;
                        l[j] = sc * sinv[cq + j];
#pragma unroll
                        for (int j2 = j + 1; j2 < 8; ++j2) l[j2] = fma(-sc, l11[(cq + j) * kLDT + cq + j2], l[j2]);
                    }
                }
                if (s8 < 3) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const double ws = __shfl_sync(0xffffffffu, wv[j], (lane & ~3) | s8);
                        if (q > s8) {
                            const double* col = l11 + (b0 + 8 * s8 + j) * kLDT + cq;
#pragma unroll
                            for (int j2 = 0; j2 < 8; ++j2) l[j2] = fma(-ws, col[j2], l[j2]);
                        }
                    }
                }
            }
            panel_tick(a, 11, &tk);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = cq + j;
                if (valid && col < nb) {
                    SW(a, r, p + col) = l[j];
                    a.P[(size_t)r + (size_t)(a.pcol0 + col) * a.W] = wv[j];
                    const double t = l[j] * wv[j];
                    dsum += t;
                    if (fabs(t) > dabs) dabs = fabs(t);
                }
                if (col < kPanelW - 32) wsh[col * kRowsPerCta + rloc] = wv[j];
            }
            __syncwarp();                                      // the quad reads each other's wsh entries next round
            panel_tick(a, 12, &tk);
        }
#pragma unroll
        for (int sft = 1; sft < 4; sft <<= 1) {
            dsum += __shfl_xor_sync(0xffffffffu, dsum, sft);
            const double o = __shfl_xor_sync(0xffffffffu, dabs, sft);
            if (o > dabs) dabs = o;
        }
        if (valid && q == 0) {
            SW(a, r, r) -= dsum;
            if (dabs > a.wmag[r]) a.wmag[r] = dabs;
        }
        __syncwarp();
    }
}

// Look-ahead "A part": rank-klen update of the column strip [rbase, cmax) (at most kPanelW columns: the NEXT panel)
// for all rows >= rbase.  It sits on the critical path between two panel factorisations, so it is cut into many
// small CTAs -- 64 x 64 tiles, 256 threads, 4 x 4 register micro-tiles (2048 FMAs per thread) -- instead of the
// 128 x 128 / 8 x 8 tiles of k_dense_update_k (8192 FMAs per thread, ~60 us for a grid that is too small to fill
// the GPU anyway).
#ifdef VBK_EMU
constexpr int kStripTG = 4;
#else
constexpr int kStripTG = 16;
#endif
constexpr int kStripTD = 4 * kStripTG;
constexpr int kStripThreads = kStripTG * kStripTG;
static __global__ void __launch_bounds__(kStripThreads) k_dense_update_strip(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* As = reinterpret_cast<double*>(raw);             // [kPanelMax][kStripTD]
    double* Bs = As + kPanelMax * kStripTD;
    const int r0 = a.rbase + blockIdx.y * kStripTD, c0 = a.rbase + blockIdx.x * kStripTD;
    if (r0 + kStripTD <= c0 || c0 >= a.cmax) return;
    const int tid = threadIdx.x, tx = tid % kStripTG, ty = tid / kStripTG;
    double acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
    for (int kc = 0; kc < a.klen; kc += kPanelMax) {
        const int kn = (a.klen - kc < kPanelMax) ? (a.klen - kc) : kPanelMax;
        __syncthreads();
        for (int e = tid; e < kPanelMax * kStripTD; e += kStripThreads) {
            const int x = e % kStripTD, c = e / kStripTD;
            As[c * kStripTD + x] = (c < kn && r0 + x < a.W) ? SW(a, r0 + x, a.kcol0 + kc + c) : 0.0;
            Bs[c * kStripTD + x] = (c < kn && c0 + x < a.W) ? a.P[(size_t)(c0 + x) + (size_t)(a.pcol0 + kc + c) * a.W] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int c = 0; c < kPanelMax; ++c) {
            double av[4], bv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { av[u] = As[c * kStripTD + tx + kStripTG * u]; bv[u] = Bs[c * kStripTD + ty + kStripTG * u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
        }
    }
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int c2 = c0 + ty + kStripTG * v;
        if (c2 >= a.W || c2 >= a.cmax) continue;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = r0 + tx + kStripTG * u;
            if (r < a.W && r > c2) SW(a, r, c2) -= acc[u][v];
        }
    }
}

}  // namespace vbk
