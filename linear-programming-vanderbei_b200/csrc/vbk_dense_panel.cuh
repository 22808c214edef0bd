// vbk_dense_panel.cuh -- dense-window factorisation (fast mode): 128-column panels.
//
// A panel is 128 columns and costs three launches in all (the 32-column panels of the first generations were a chain
// of ~400 small dependent kernels, ~100 us per panel on dfl001, while the rank-128 updates that hold nearly all the
// flops took ~1.5 ms):
//   k_panel_diag  ONE CTA factorises the 128 x 128 diagonal block in shared memory: four 32-column
//                 sub-blocks, each = warp-level LDL^T in registers (warp 0), row-parallel substitution for
//                 the block rows below it, rank-32 update of the rest of the block (all warps);
//   k_panel_rows  every row below the block is owned by FOUR threads (a quad: 8 of the 32 columns of a sub-block
//                 each) for the whole panel; L11^T, the reciprocal pivots and the keep flags arrive in shared
//                 memory as ONE bulk asynchronous copy (cp.async.bulk + mbarrier) of the packed panel buffer
//                 the diagonal kernel leaves behind; per sub-block a rank-(32b) update from the row's own
//                 earlier results, then four 8-column substitution stages chained by quad shuffles;
//   k_dense_update_m  (vbk_dense_update.cuh) the rank-128 update of the trailing matrix on the FP64 tensor path.
// The dependent-pivot rule (reference ldlt.c:600-614) keeps its meaning: when a pivot is "zero" the
// maximum of the updated column below it decides between dropping the row and substituting the pivot;
// rows whose panel columns have not been touched yet get them applied on the fly by the whole CTA
// (rare path, O(rows * c^2)).
#pragma once
#include "vbk_dense_common.cuh"

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
constexpr size_t kPanelDiagSmem = sizeof(double) * (kPanelW * kLDD + (kPanelW - 32) * 33 + 3 * kPanelW + kDiagThreads + 192)
                                  + sizeof(int) * (kPanelW + 4);
constexpr size_t kPanelRowsSmem = sizeof(double) * (kPanelBufDoubles + (kPanelW - 32) * kRowsPerCta) + 16;
// second packed buffer, for the tensor-path rows kernel (k_panel_rows_m): the six off-diagonal 32 x 32 blocks of L11
// as MMA "B" operands (block (b, b'), b' < b, element [i * kPB2Ld + j] = L11(32 b + j, 32 b' + i)), the four inverted
// diagonal blocks ((I + L_bb)^-1 row-major, leading dimension kPB2Ld), reciprocal pivots, keep flags.  kPB2Ld = 36:
// a fragment load B[t][g] then hits banks 4 t + g.
constexpr int kPB2Ld = 36;
constexpr int kPB2Blk = 32 * kPB2Ld;
constexpr int kPB2Off = 0, kPB2Inv = 6 * kPB2Blk, kPB2Sinv = 10 * kPB2Blk, kPB2Keep = kPB2Sinv + 128;
constexpr int kPanelBuf2Doubles = kPB2Keep + 128;

#ifndef VBK_EMU
__device__ __forceinline__ void cp_async8(void* dst, const void* src, bool ok)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = ok ? 8 : 0;                  // src-size 0: destination is zero-filled, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool ok)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// FP64 tensor-path MMA, D(8x8) += A(8x4) B(4x8).  With g = lane / 4, t = lane % 4 a thread holds A[g][t], B[t][g]
// and C[g][2t], C[g][2t+1] (PTX ISA, mma.m8n8k4 .f64 fragments).
__device__ __forceinline__ void dmma884(double& c0, double& c1, double av, double bv)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
}
#endif

// correctly rounded reciprocal: one MUFU seed + Newton steps, about half the dependent chain of a full division
__device__ __forceinline__ double vbk_rcp(double d) {
#ifdef VBK_EMU
    return 1.0 / d;
#else
    return __drcp_rn(d);
#endif
}

// reciprocal to full double precision without the special-case call of __drcp_rn (that call is a scheduling
// barrier for the compiler): MUFU seed (relative error ~2^-20), one cubic and one quadratic Newton step.  Pivots
// outside the normal range do not get here (the dependent-pivot test catches them).
__device__ __forceinline__ double vbk_rcp_fast(double d) {
#ifdef VBK_EMU
    return 1.0 / d;
#else
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    e = fma(e, e, e);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    return fma(x, e, x);
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

// Rare path of k_panel_diag, called by all lanes of warp 0 when the pivot of column c of the current sub-block is
// "zero" (|d| <= tol * largest term, reference rule ldlt.c:600-614): the other warps are released from their
// command loop to help with max |column| over every row below (panel_colmax), the result decides between dropping
// the row (returns 0) and substituting the pivot (returns the substitute, never 0).  One copy, out of line: the
// unrolled column loop stays small.
#ifndef VBK_EMU
__noinline__
#endif
__device__ double panel_rare_pivot(const DenseArgs& a, int b0, int nbb, int c, double mymax, double magc,
                                   const double* blk, const double* sd, const double* sinv, const int* skeep,
                                double* red, volatile int* s_cmd)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    if (lane == 0) *s_cmd = c;
    __syncthreads();                                       // A: the helpers pick the command up
    const double below = panel_colmax(a, b0, nbb, b0 + c, tid, nt, blk, sd, sinv, skeep);
    red[tid] = fmax(mymax, below);
    __syncthreads();                                       // B
    double m = 0.0;
    for (int u = lane; u < nt; u += 32) m = fmax(m, red[u]);
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, sft));
    if (lane == 0) atomicAdd(&a.counters[C_NDEP], 1);
    if (m < 1.0e+6 * 1.0e-8) return 0.0;
    double sub = a.piv_scale * magc;
    if (!(sub > 1.0e-8)) sub = 1.0e-8;
    return (a.perm[a.T + a.p + b0 + c] < a.n_ld ? -1 : 1) * sub;
}

// LDL^T of one 32 x 32 sub-block of k_panel_diag by warp 0, row `lane` in the lane's registers.  kRare = false: no
// dependent-pivot branch, returns whether some pivot failed the test (the caller then repeats the sub-block with
// kRare = true); finished columns are parked as a = l*d at park[lane * park_ld + c].
// (Taking this function out of line -- so that it is compiled like the stand-alone micro-benchmark, profiles/ubench/ubench2.cu,
// 127 cycles per column -- did not help: 440 cycles per column, the shared-memory arrays become generic pointers.)
#ifndef VBK_EMU
// (I + L_bb)^-1 of the finished 32 x 32 sub-block starting at column bs of the panel, for the tensor-path rows kernel:
// one warp, lane j owns column j (x = e_j, right-looking substitution in registers, L_bb as a broadcast), result straight
// to the packed buffer (row-major, coalesced over the lanes); rows past a partial panel: identity.
__device__ __forceinline__ void panel_inverse32(const DenseArgs& a, const double* blk, int bs, int lane)
{
    double x[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) x[u] = (u == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int kk = 0; kk < 31; ++kk) {
        const double xk = x[kk];
#pragma unroll
        for (int u = kk + 1; u < 32; ++u) x[u] = fma(-blk[(bs + u) * kLDD + bs + kk], xk, x[u]);
    }
    double* dst = a.PB2 + kPB2Inv + (bs >> 5) * kPB2Blk;
#pragma unroll
    for (int u = 0; u < 32; ++u) dst[u * kPB2Ld + lane] = (bs + u < a.nb) ? x[u] : (u == lane ? 1.0 : 0.0);
}
#endif

// Ordering point between a lane's store into the column ring and the other lanes' loads: a real warp barrier
// (bar.warp.sync orders the participating lanes' shared-memory accesses; lockstep execution of straight-line code is
// not something CUDA guarantees under independent thread scheduling).  The optimistic pass issues the barrier as inline
// PTX: __syncwarp() there made the compiler emit a divergence check and re-derive the shared-memory window base
// (S2UR SR_CgaCtaId) after every one of them, right on the dependent chain.
template <bool kRare> __device__ __forceinline__ void ldl_fence()
{
#ifdef VBK_EMU
    __syncwarp();
#else
    if (kRare) __syncwarp();
    else asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
#endif
}

template <bool kRare>
__device__ __forceinline__
bool panel_ldl32(const DenseArgs& a, int b0, int nbb, int lane, double* blk, double* sd, double* sinv, int* skeep, double* wm,
                 double* colbuf, double* red, volatile int* s_cmd)
{
    // rows past a partial sub-block are padded with unit pivots (d = 1, nothing below them): the column
    // loop below is then branch-free apart from the rare path, whatever nbb is
    __syncwarp();
    double ar[32];
#pragma unroll
    for (int j = 0; j < 32; ++j)
        ar[j] = (lane < nbb && j <= lane) ? blk[(b0 + lane) * kLDD + b0 + j] : ((lane >= nbb && j == lane) ? 1.0 : 0.0);
    double wmr = (lane < nbb) ? wm[b0 + lane] : 0.0;
    // Column c of the current matrix is broadcast through a small shared-memory buffer (three copies in
    // rotation, 16-byte aligned): one 8-byte store per lane, then every lane reads the pivot, and the
    // a_{j,c} it needs two at a time.  (A shuffle per (c, j) pair made the unrolled loop 60 KB of code --
    // twice the instruction cache -- and ran at 480 cycles per column; shuffles for the pivot alone were no
    // faster than the shared-memory round trip, profiles/r01_summary.md.)
    // The loop is software-pipelined by hand, because a warp issues in order: as soon as column c has
    // updated a_{.,c+1} (ONE fma per lane), column c+1 is published and its pivot's reciprocal started; the
    // other 30 - c updates of column c then fill that latency.  Dependent chain per column: fma, store,
    // load, reciprocal (one MUFU + five fma, no slow-path call), multiply.
    // Bookkeeping is kept off that chain in the optimistic pass (kRare = false; profiles/ubench/ubench2.cu: 126 cycles per
    // column bare, 233 with per-column pivot stores + parked column + broadcast term magnitude): the pivot test is
    // made by the lane that owns the pivot on its own registers and voted on once at the end, pivots and
    // reciprocals stay in the owning lane's registers, and the finished columns -- a_{r,c} is not touched again
    // after column c -- are written from the registers after the loop.
    double d, magc = 0.0, inv, nxt;
    double myd = 1.0, myinv = 0.0;
    bool badl = false;
    {
        double* cb = colbuf;
        cb[lane] = ar[0];
        if (kRare) cb[32 + lane] = wmr;
        ldl_fence<kRare>();
        d = cb[0]; nxt = cb[1];
        if (kRare) magc = cb[32];
        inv = vbk_rcp_fast(d);
        if (!kRare) badl = lane == 0 && fabs(ar[0]) <= a.tol * wmr;
    }
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const double* cb = colbuf + (c % 3) * 64;
        const double arc = ar[c];                                      // a_{r,c} = l_{r,c} d_c   (lanes r > c)
        int keep = 1;
        if (kRare && __builtin_expect(fabs(d) <= a.tol * magc, 0)) {   // uniform over the warp; ldlt.c:600-614
            const double nd = panel_rare_pivot(a, b0, nbb, c, (lane > c && lane < nbb) ? fabs(arc) : 0.0, magc,
                                               blk, sd, sinv, skeep, red, s_cmd);
            if (nd != 0.0) { d = nd; inv = vbk_rcp(d); }               // substituted pivot
            else { keep = 0; inv = 0.0; }                              // dependent row: dropped
        }
        const double lr = arc * inv;                                   // l_{r,c}
        const bool mine = lane == c && c < nbb, below = lane > c && lane < nbb;
        if (kRare) {
            if (mine) { sd[b0 + c] = d; sinv[b0 + c] = inv; skeep[b0 + c] = keep; }
            if (below) blk[(b0 + lane) * kLDD + b0 + c] = arc;         // parked as l*d (the rare path reads it)
        } else if (mine) { myd = d; myinv = inv; }
        const double term = fabs(lr * arc);                            // what this column adds to a_{r,r}
        wmr = (below && term > wmr) ? term : wmr;
        // a_{r,j} -= l_{r,c} d_c l_{j,c} = lr * a_{j,c}   (meaningful for r >= j; the rest is never read)
        if (c + 1 < 32) {
            ar[c + 1] = fma(-lr, nxt, ar[c + 1]);                      // column c+1 is final now: publish it
            if (!kRare) badl = badl || (lane == c + 1 && fabs(ar[c + 1]) <= a.tol * wmr);
            double* cn = colbuf + ((c + 1) % 3) * 64;
            cn[lane] = ar[c + 1];
            if (kRare) cn[32 + lane] = wmr;
            ldl_fence<kRare>();
            if (((c + 1) & 1) == 0) {                                  // pivot and a_{c+2,c+1} in one 16-byte load
                const double2 v = *reinterpret_cast<const double2*>(cn + c + 1);
                d = v.x; nxt = v.y;
            } else { d = cn[c + 1]; nxt = (c + 2 < 32) ? cn[c + 2] : 0.0; }
            if (kRare) magc = cn[32 + c + 1];
            inv = vbk_rcp_fast(d);
            // the rest of column c's update, in the shadow of that reciprocal
            if (c & 1) {                                               // c + 2 odd
                if (c + 2 < 32) ar[c + 2] = fma(-lr, cb[c + 2], ar[c + 2]);
#pragma unroll
                for (int j = c + 3; j < 32; j += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(cb + j);
                    ar[j] = fma(-lr, v.x, ar[j]);
                    ar[j + 1] = fma(-lr, v.y, ar[j + 1]);
                }
            } else {
#pragma unroll
                for (int j = c + 2; j < 32; j += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(cb + j);
                    ar[j] = fma(-lr, v.x, ar[j]);
                    ar[j + 1] = fma(-lr, v.y, ar[j + 1]);
                }
            }
        }
    }
    if (kRare) {
        if (lane < nbb) wm[b0 + lane] = wmr;
        return false;
    }
    const bool bad = __any_sync(0xffffffffu, badl && lane < nbb);
    if (!bad) {
        if (lane < nbb) { sd[b0 + lane] = myd; sinv[b0 + lane] = myinv; skeep[b0 + lane] = 1; wm[b0 + lane] = wmr; }
#pragma unroll
        for (int c = 0; c < 31; ++c)
            if (lane > c && lane < nbb) blk[(b0 + lane) * kLDD + b0 + c] = ar[c];    // as l*d; scaled by the pass below
    }
    return bad;
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
    double* colbuf = red + kDiagThreads;                      // [3][64] column broadcast buffers of the warp LDL^T
    int* skeep = reinterpret_cast<int*>(colbuf + 192);        // [kPanelW]
    volatile int* s_cmd = skeep + kPanelW;                    // rare-path command slot (see below)
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int p = a.p, nb = a.nb;
    long long tk = vbk_clock();

    if (nb < kPanelW)                                              // partial panel: the padding must read as zero
        for (int e = tid; e < kPanelW * kLDD; e += nt) blk[e] = 0.0;
    for (int e = tid; e < kPanelW; e += nt) { sd[e] = 1.0; sinv[e] = 0.0; skeep[e] = 0; wm[e] = (e < nb) ? a.wmag[p + e] : 0.0; }
    __syncthreads();
    // lower triangle of the block, column-major in HBM -> row-major (stride 129) in shared memory.  Asynchronous
    // copies: 64 independent 8-byte transfers per thread in flight at once (plain load/store pairs were serialised
    // by the compiler -- 18 us per panel, profiles/r01_summary.md)
    for (int e = tid; e < kPanelW * kPanelW; e += nt) {            // consecutive threads = consecutive rows: coalesced
        const int r = e % kPanelW, c = e / kPanelW;                // kPanelW is a power of two
        if (r >= c && r < nb) {
#ifdef VBK_EMU
            blk[r * kLDD + c] = SW(a, p + r, p + c);
#else
            cp_async8(&blk[r * kLDD + c], &SW(a, p + r, p + c), true);
#endif
        }
    }
#ifndef VBK_EMU
    cp_async_commit();
    cp_async_wait<0>();
#endif
    __syncthreads();
    panel_tick(a, 0, &tk);

    for (int b0 = 0; b0 < nb; b0 += 32) {
        const int nbb = (nb - b0 < 32) ? (nb - b0) : 32;
        // ---- (a) the 32 x 32 sub-block: right-looking LDL^T by WARP 0 alone, row `lane` of the sub-block in the
        // lane's registers.  Column c: the pivot comes by shuffle from lane c, every lane takes the reciprocal for
        // itself, and the rank-1 update of the remaining columns is one shuffle (a_{j,c} from lane j) and one FMA
        // per column and lane -- no barrier, no shared-memory round trip on the dependent chain (shuffle + reciprocal
        // + multiply + FMA, ~110 cycles per column; the CTA-wide version with a barrier per column took 790,
        // profiles/r01_summary.md).  A finished column is parked in shared memory as a = l*d (what the rare path
        // and the scaling pass below expect).  The other warps wait in a command loop and only help in the rare
        // dependent-pivot path (reference ldlt.c:600-614), which needs max |column| over all rows below.
        if (warp == 0) {
            // optimistic pass without the dependent-pivot branch and without per-column bookkeeping; if a pivot failed
            // the test, the sub-block -- still untouched in blk -- is redone with the full rule.
            const bool bad = panel_ldl32<false>(a, b0, nbb, lane, blk, sd, sinv, skeep, wm, colbuf, red, s_cmd);
            if (a.prof && lane == 0) { atomicAdd(&a.prof[14], 1ull); if (bad) atomicAdd(&a.prof[15], 1ull); }
            if (bad) panel_ldl32<true>(a, b0, nbb, lane, blk, sd, sinv, skeep, wm, colbuf, red, s_cmd);
            if (lane == 0) *s_cmd = -1;
            __syncthreads();                                                   // A: releases the helpers
        } else {
#ifndef VBK_EMU
            // the last warp inverts the PREVIOUS sub-block while warp 0 factorises this one (nobody waits for it)
            if (a.PB2 && warp == (nt >> 5) - 1 && b0 > 0) panel_inverse32(a, blk, b0 - 32, lane);
#endif
            for (;;) {
                __syncthreads();                                               // A
                const int cmd = *s_cmd;
                if (cmd < 0) break;
                red[tid] = panel_colmax(a, b0, nbb, b0 + cmd, tid, nt, blk, sd, sinv, skeep);
                __syncthreads();                                               // B
            }
        }
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
            double dm = 0.0;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                if (c < nbb) { blk[r * kLDD + b0 + c] = l[c]; wbuf[t * 33 + c] = w[c]; dm = fmax(dm, fabs(l[c] * w[c])); }
            }
            if (dm > wm[r]) wm[r] = dm;                                        // largest term of the update of a_{r,r} below
        }
        __syncthreads();
        panel_tick(a, 2, &tk);
        // ---- (c) rank-nbb update of the rest of the block (lower triangle incl. diagonal)
#ifndef VBK_EMU
        // 32 x 32 output blocks, one warp each, on the FP64 tensor path: 16 accumulator tiles, 8 steps of k.  Lane
        // t of a fragment takes k = 16 (s / 4) + 4 t + s % 4 in step s (any assignment works as long as both
        // operands use it): with the row strides 129 and 33 a half-warp then reads banks g + 4 t + const -- all
        // different.  nbb == 32 here (a partial sub-block is the last one and has nothing below it).
        if (a.PB2 && warp == (nt >> 5) - 1) {
            // the finished off-diagonal blocks L11[b][b'] (b' = this sub-block, b below it) go to the packed buffer of
            // k_panel_rows_m now, by a warp that owns no output block (at most 6 pairs for 8 warps)
            const int bp = b0 >> 5;
            for (int b = bp + 1; b < 4; ++b) {
                double* dst = a.PB2 + kPB2Off + (b * (b - 1) / 2 + bp) * kPB2Blk;
#pragma unroll 8
                for (int i = 0; i < 32; ++i) dst[i * kPB2Ld + lane] = blk[(32 * b + lane) * kLDD + b0 + i];   // [i][j] = L(32b + j, 32b' + i)
            }
        }
        {
            const int g = lane >> 2, t4 = lane & 3;
            const int nblk = (rem + 31) >> 5, npairs = nblk * (nblk + 1) / 2;
            for (int pr = warp; pr < npairs; pr += (nt >> 5)) {
                int bi = 0;
                while ((bi + 1) * (bi + 2) / 2 <= pr) ++bi;
                const int bj = pr - bi * (bi + 1) / 2;
                double acc[4][4][2];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
                const double* Ap = blk + (b0 + nbb + 32 * bi + g) * kLDD + b0 + 4 * t4;
                const double* Bp = wbuf + (32 * bj + g) * 33 + 4 * t4;
#pragma unroll
                for (int s8 = 0; s8 < 8; ++s8) {
                    const int kk = 16 * (s8 >> 2) + (s8 & 3);
                    double av[4], bv[4];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) av[mi] = Ap[mi * 8 * kLDD + kk];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) bv[ni] = (32 * bj + 8 * ni + g < rem) ? Bp[ni * 8 * 33 + kk] : 0.0;
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], av[mi], bv[ni]);
                }
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int rr = 32 * bi + 8 * mi + g, cc = 32 * bj + 8 * ni + 2 * t4 + j;
                            if (rr < rem && cc <= rr) blk[(b0 + nbb + rr) * kLDD + b0 + nbb + cc] -= acc[mi][ni][j];
                        }
            }
        }
#else
        // emulated build: 4 x 4 register tiles, rows and columns of a tile interleaved (tr + i*nt4, tc + j*nt4)
        {
            const int nt4 = (rem + 3) >> 2;
            for (int e = tid; e < nt4 * nt4; e += nt) {
                const int tr = e % nt4, tc = e / nt4;
                double acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
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
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int rr = tr + i * nt4, cc = tc + j * nt4;
                        if (rr < rem && cc < rem && rr >= cc) blk[(b0 + nbb + rr) * kLDD + b0 + nbb + cc] -= acc[i][j];
                    }
            }
        }
#endif
        __syncthreads();
        panel_tick(a, 3, &tk);
    }
    __syncthreads();
#ifndef VBK_EMU
    if (a.PB2 && warp == (nt >> 5) - 1) panel_inverse32(a, blk, ((nb - 1) >> 5) << 5, lane);     // the last sub-block's
#endif
    const bool want_pb = a.PB2 == nullptr;                          // packed copy for the DFMA rows kernel (k_panel_rows) only
    for (int c = warp; c < kPanelW; c += (nt >> 5)) {
        for (int r = lane; r < kPanelW; r += 32) {
            const double v = (r > c && r < nb) ? blk[r * kLDD + c] : 0.0;
            if (r > c && r < nb) SW(a, p + r, p + c) = v;
            if (want_pb) a.PB[c * kLDT + r] = v;
        }
        if (want_pb && lane < kLDT - kPanelW) a.PB[c * kLDT + kPanelW + lane] = 0.0;
    }
#ifndef VBK_EMU
    if (a.PB2) {
        for (int e = tid; e < kPanelW; e += nt) { a.PB2[kPB2Sinv + e] = sinv[e]; a.PB2[kPB2Keep + e] = skeep[e] ? 1.0 : 0.0; }
    }
#endif
    for (int e = tid; e < kPanelW; e += nt) {
        if (want_pb) {
            a.PB[kPanelW * kLDT + e] = sinv[e];
            a.PB[kPanelW * kLDT + kPanelW + e] = skeep[e] ? 1.0 : 0.0;
        }
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


#ifndef VBK_EMU
// rows below the diagonal block on the FP64 tensor path.  With w = l * d the row equation w L11^T = s is solved 32
// columns at a time:  w_b = (s_b - sum_{b' < b} w_b' L11[b][b']^T) (I + L_bb)^-T  -- products of a 16-row slab with
// 32 x 32 blocks, all mma.sync.m8n8k4.  A warp owns 16 rows for the whole panel; its w (the next products' "A"
// operand) and the accumulator it has to turn into an operand go through a per-warp shared-memory slab (leading
// dimension 132: fragment loads hit banks 4 g + t).  The blocks arrive as ONE bulk asynchronous copy of the packed
// buffer k_panel_diag leaves behind (kPanelBuf2Doubles).  Same outputs as k_panel_rows: L21 in place, P = L21 D,
// trailing diagonal and its largest term.
constexpr int kRowsMWarps = 4;
constexpr int kRowsMLd = 132;
constexpr size_t kPanelRowsMSmem = sizeof(double) * (kPanelBuf2Doubles + kRowsMWarps * 16 * kRowsMLd) + 16;

static __global__ void __launch_bounds__(kRowsMWarps * 32) k_panel_rows_m(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* pb = reinterpret_cast<double*>(raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    double* wsl = pb + kPanelBuf2Doubles + warp * 16 * kRowsMLd;          // this warp's 16 x 128 slab
    const int nb = a.nb, p = a.p;
    {
        unsigned long long* mbar = reinterpret_cast<unsigned long long*>(pb + kPanelBuf2Doubles + kRowsMWarps * 16 * kRowsMLd);
        const unsigned mbar_s = (unsigned)__cvta_generic_to_shared(mbar);
        const unsigned dst_s = (unsigned)__cvta_generic_to_shared(pb);
        constexpr unsigned kBytes = (unsigned)(sizeof(double) * kPanelBuf2Doubles);
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_s));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_s), "r"(kBytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst_s), "l"(a.PB2), "r"(kBytes), "r"(mbar_s) : "memory");
        }
        unsigned done = 0;
        while (!done) {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(mbar_s), "r"(0u) : "memory");
        }
    }
    const double* sinv = pb + kPB2Sinv;
    const double* keepd = pb + kPB2Keep;
    const int nslabs_all = (a.W - p - nb + 15) / 16;
    const int nslabs = nslabs_all < a.slab_hi ? nslabs_all : a.slab_hi;
    for (int sl = a.slab_lo + blockIdx.x * kRowsMWarps + warp; sl < nslabs; sl += gridDim.x * kRowsMWarps) {
        const int r0 = p + nb + sl * 16;
        // the slab of S21: rows r0 + i, columns p + c  ->  wsl[i][c]   (lane = row pair / column: coalesced over rows)
        for (int e = lane; e < 16 * kPanelW; e += 32) {
            const int i = e & 15, c = e >> 4;
            wsl[i * kRowsMLd + c] = (r0 + i < a.W && c < nb) ? SW(a, r0 + i, p + c) : 0.0;
        }
        __syncwarp();
        double dsum[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, dabs[2][2] = {{0.0, 0.0}, {0.0, 0.0}};    // [m tile][row half]: rows g (both), see below
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            // acc = s_b as C fragments: C[g][2t + j] of m tile mi (rows 8 mi + g), n tile ni (columns 32 b + 8 ni + 2t + j)
            double acc[2][4][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int j = 0; j < 2; ++j) acc[mi][ni][j] = wsl[(8 * mi + g) * kRowsMLd + 32 * b + 8 * ni + 2 * t + j];
            // minus the earlier sub-blocks' w times L11[b][b']^T
#pragma unroll
            for (int bp = 0; bp < b; ++bp) {
                const double* Bk = pb + kPB2Off + (b * (b - 1) / 2 + bp) * kPB2Blk;
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                    double av[2], bv[4];
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) av[mi] = -wsl[(8 * mi + g) * kRowsMLd + 32 * bp + 4 * k4 + t];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) bv[ni] = Bk[(4 * k4 + t) * kPB2Ld + 8 * ni + g];
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], av[mi], bv[ni]);
                }
            }
            // the accumulator becomes an operand: through the slab (its s_b columns are consumed)
            __syncwarp();
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int j = 0; j < 2; ++j) wsl[(8 * mi + g) * kRowsMLd + 32 * b + 8 * ni + 2 * t + j] = acc[mi][ni][j];
            __syncwarp();
            // w_b = acc (I + L_bb)^-T :  C[m][n] = sum_k acc[m][k] Inv[n][k]
            double wv[2][4][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) { wv[mi][ni][0] = 0.0; wv[mi][ni][1] = 0.0; }
            const double* Iv = pb + kPB2Inv + b * kPB2Blk;
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                double av[2], bv[4];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi) av[mi] = wsl[(8 * mi + g) * kRowsMLd + 32 * b + 4 * k4 + t];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) bv[ni] = Iv[(8 * ni + g) * kPB2Ld + 4 * k4 + t];
#pragma unroll
                for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma884(wv[mi][ni][0], wv[mi][ni][1], av[mi], bv[ni]);
            }
            __syncwarp();
            // dropped columns give w = l = 0; outputs; w back into the slab for the later sub-blocks
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int c = 32 * b + 8 * ni + 2 * t + j, r = r0 + 8 * mi + g;
                        const double w = (keepd[c] != 0.0) ? wv[mi][ni][j] : 0.0;
                        const double l = w * sinv[c];
                        wsl[(8 * mi + g) * kRowsMLd + c] = w;
                        if (r < a.W && c < nb) {
                            SW(a, r, p + c) = l;
                            a.P[(size_t)r + (size_t)(a.pcol0 + c) * a.W] = w;
                            const double tt = l * w;
                            dsum[mi][0] += tt;
                            dabs[mi][0] = fmax(dabs[mi][0], fabs(tt));
                        }
                    }
            __syncwarp();
        }
        // row r0 + 8 mi + g: its four lanes (t = 0..3) hold partial sums
#pragma unroll
        for (int mi = 0; mi < 2; ++mi) {
            double ds = dsum[mi][0], da = dabs[mi][0];
#pragma unroll
            for (int sft = 1; sft < 4; sft <<= 1) {
                ds += __shfl_xor_sync(0xffffffffu, ds, sft);
                da = fmax(da, __shfl_xor_sync(0xffffffffu, da, sft));
            }
            const int r = r0 + 8 * mi + g;
            if (t == 0 && r < a.W) {
                SW(a, r, r) -= ds;
                if (da > a.wmag[r]) a.wmag[r] = da;
            }
        }
        __syncwarp();
    }
}
#endif  // !VBK_EMU

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
