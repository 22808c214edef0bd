// vbk_strict_factor.cuh -- STRICT numeric LDL^T (lltnum, reference src/ipo/ldlt.c:565-631), third
// generation: the same bits as the reference, organised around the only thing that bounds strict
// mode -- the dependent chain of rounded additions.
//
// Column i of the reference accumulates, for every row below it,
//     temp[row] = (((0 + p_1) + p_2) + ...),   p_q = (L[i,j_q] * d_jq) * L[row,j_q]      (ldlt.c:572,583,588)
// over its contributing columns j_q in link-list order sigma_i.  The products
// are independently rounded, so they can be formed by anybody at any time; only the additions are a
// chain, one chain per entry of the column, |sigma_i| links long.  sigma_i starts with the column that
// finished LAST (in the dense tail sigma_i = [i-1, i-3, ..., i-4, i-2]), so a column's chains cannot
// start before its predecessor is final: per factorisation the critical path is ~Lnz dependent FP64
// additions (dfl001: 6.96 M of 7.16 M), 8 cycles each on B200.  The round-1 kernel spent ~150 cycles per link: load, scatter, barrier, add phases of a whole CTA
// in sequence, all of it starting only when the last child had finished.  Here a CTA is a pipeline:
//
//   producer warps   form the products of 32 contributors at a time into a ring of shared-memory
//                    tiles tile[stage][q][slot] -- row q holds contributor q's products for every row
//                    of the task, +0.0 where the contributor has no entry (adding +0.0 is exact and the
//                    accumulator is never -0.0).  Which rows a contributor holds is a precomputed bit
//                    mask per (task, contributor) pair (k_pipe_masks), so every row costs the same few
//                    instructions.  They wait per CONTRIBUTOR for its column to be final (col_done[j]),
//                    so everything except the products of the last child is staged before that child
//                    finishes;
//   consumer warp    one lane per row of the task (up to 4 rows per lane when row blocks are wider
//                    than 32): the accumulator lives in a register, a ring stage is 32 back-to-back
//                    dependent additions per lane, in the reference's order, kPipeMulti stages per
//                    flag check;
//   pivot warp       (pivot-owning slice only) the diagonal's own chain diagi -= lij*(lij*dj)
//                    (ldlt.c:573) from the staged products, with the consumer's loop: this chain is as
//                    long as the rows' chains and must not be any slower.
//
// Tasks are (column, 32-row block) slices as before (vbk_symbolic.h), claimed in index order by
// persistent CTAs.  The LAST slice of a column owns the pivot: the other slices publish their
// undivided entries and leave at once (no CTA ever waits for a task claimed after its own, so the
// kernel cannot deadlock whatever the grid or however many handles share the GPU); the owner applies
// the pivot rule (ldlt.c:600-614, max|offdiag| over all slices only when the pivot is exactly 0),
// divides the whole column and raises col_done[i].
#pragma once
#include "vbk_kernels.cuh"

namespace vbk {

constexpr int kPipeQ = 32;            // contributors per ring stage
#ifndef VBK_PIPE_MULTI
#ifdef VBK_EMU
#define VBK_PIPE_MULTI 2      // the emulated ring has three stages
#else
#define VBK_PIPE_MULTI 6
#endif
#endif
constexpr int kPipeMulti = VBK_PIPE_MULTI;      // ring stages a chain takes per flag check (when the ring is deeper than that)
constexpr int kPipeMaxChains = 4;     // rows per consumer lane: tasks of up to 128 rows
#ifdef VBK_EMU
constexpr int kPipeWarpsDefault = 4;  // consumer, pivot, 2 producers (every CUDA thread is an OS thread there)
constexpr int kPipeStagesMax = 3;
#else
constexpr int kPipeWarpsDefault = 16; // consumer, pivot, 14 producers
constexpr int kPipeStagesMax = 32;
#endif

struct PipeArgs {
    int N, n_ld, ntasks;
    int nstages;        // ring depth
    const int* kL; const int* iL; double* L; double* diag; int* mark;
    const int* rowptr; const int* rk; const int* rj;      // row lists in sigma order
    const int* perm;
    const int* task_col; const int* task_blk; const int* task_pos0; const int* task_cnt;
    const int* col_task0; const int* col_ntask;
    const int* winptr; int nblk, rowblk, slice_row0;
    // presence masks (k_pipe_masks, once per analysis): for pair number task_pair0[t] + 32*g + q -- contributor q of group g
    // of task t -- NCH words whose bit s says that the contributor's column holds row s of the task
    const unsigned* masks; const long long* task_pair0;
    int* col_pub;       // [N] non-owner slices of the column that have published their undivided entries
    int* col_done;      // [N] 1 once L[:,j], diag[j] and mark[j] are final
    double* task_max;   // [ntasks] max|undivided entry| of a non-owner slice
    int* counters; const unsigned long long* scal_bits; double epsnum;
    unsigned backoff_ns;    // sleep of a producer between two looks at a full ring
    // optional [16] cycle counters ($VBK_PROF), lane 0 of each role: 0 consumer waits for a stage, 1 consumer adds,
    // 2 producer waits for a free slot, 3 static structure + zero fill, 4 waits for the contributors' columns,
    // 5 fence + lij, dj, 6 products, 7 publish; 8 claim + task setup, 9 epilogue up to col_pub / pivot wait,
    // 10 owner waits for the other slices, 11 pivot rule, 12 divide, 13 fence + col_done, 14 groups staged, 15 tasks
    unsigned long long* prof;
    // optional [N][8] per-column event times in ns (globaltimer), written by the pivot-owning slice ($VBK_PROF):
    // 0 claim, 1 first ring stage consumed, 2 chains finished, 3 all slices published, 4 col_done raised,
    // 5 contributor groups, 6 slices, 7 SM
    long long* trace;
};

// bytes of dynamic shared memory the kernel needs
inline size_t pipe_smem_bytes(int cap, int nstages, int rowblk, int nwarps)
{
    (void)rowblk;
    size_t d = (size_t)nstages * kPipeQ * cap      // product tiles
             + (size_t)nstages * kPipeQ            // lij*(lij*dj) per staged contributor (the pivot chain's operands)
             + (size_t)cap + nwarps + 2;           // undivided column, per-warp max, pivot
    size_t i = (size_t)nstages + 2 + 8;
    return d * sizeof(double) + i * sizeof(int);
}

__device__ __forceinline__ long long vbk_globaltimer() {
#ifdef VBK_EMU
    return 0;
#else
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
#endif
}
static __global__ void k_pipe_reset(int N, int* __restrict__ col_pub, int* __restrict__ col_done, int* __restrict__ counters)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) { col_pub[t] = 0; col_done[t] = 0; }
    if (blockIdx.x == 0 && threadIdx.x == 0) { counters[C_NEXT] = 0; counters[C_NDEP] = 0; }
}

#ifndef VBK_PIPE_FB
#define VBK_PIPE_FB 32
#endif

// Presence masks of every (task, contributor) pair, computed once per analysis: lane = contributor of a group, one
// CTA-stride loop over the tasks.  Bit s of the pair's mask is set when the contributor's column j holds row s of the
// task; its entries in the task's rows are consecutive in column j (they start at kb, see the producers), so the entry of
// slot s is kb + (number of set bits below s).
template <int NCH>
static __global__ void k_pipe_masks(PipeArgs a, unsigned* __restrict__ masks)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int wstride = a.nblk + 1;
    for (int t = blockIdx.x; t < a.ntasks; t += gridDim.x) {
        const int i = a.task_col[t], blk = a.task_blk[t], p0 = a.task_pos0[t], cnt = a.task_cnt[t];
        const int rb = a.rowptr[i], re = a.rowptr[i + 1];
        const int ngroups = (re - rb + kPipeQ - 1) / kPipeQ;
        const int* trow = a.iL + p0;                  // the task's rows, ascending
        for (int g = warp; g < ngroups; g += nwarps) {
            const int tq = rb + g * kPipeQ + lane;
            unsigned m[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) m[c] = 0u;
            if (tq < re) {
                const int k = a.rk[tq], j = a.rj[tq];
                int kb = k + 1, ke = a.kL[j + 1];
                if (blk >= 0) {
                    const int* wp = a.winptr + (size_t)j * wstride;
                    if (wp[blk] > kb) kb = wp[blk];
                    if (wp[blk + 1] < ke) ke = wp[blk + 1];
                }
                int s = 0;                            // both lists ascend: one merge pass
                for (int e = kb; e < ke; ++e) {
                    const int r = a.iL[e];
                    while (s < cnt && trow[s] < r) ++s;
                    if (s < cnt && trow[s] == r) {
#pragma unroll
                        for (int c = 0; c < NCH; ++c) if ((s >> 5) == c) m[c] |= 1u << (s & 31);
                    }
                }
            }
            const size_t pair = (size_t)a.task_pair0[t] + (size_t)g * kPipeQ + lane;
#pragma unroll
            for (int c = 0; c < NCH; ++c) masks[pair * NCH + c] = m[c];
        }
    }
}

// Products of (a subset of) one ring stage.  A producer warp is bound by its own instruction stream and by the round
// trips it exposes, so every row is the same few instructions and there are no look-ups at all:
//  * lane q of the warp holds contributor q's record: w = lij*dj, kb = first entry of column j in the rows of the task,
//    mask = which of the task's rows column j holds (precomputed, k_pipe_masks); rows get it by warp shuffle;
//  * lane s of row q loads entry kb + popcount(mask below s) -- its own entry when bit s is set -- and stores
//    w * value, or +0.0 when the contributor has no entry in that row (x + (+0.0) == x; the accumulator is never -0.0);
//  * the loads are UNCONDITIONAL (a predicated load keeps its predicate alive until the value is used, and with seven
//    predicate registers the compiler then serialises a batch six rows at a time) and always hit the contributor's own
//    entries or the ones right behind them (never a fixed dummy address: thousands of warps reading one line make it a
//    hot spot in the L2); L is padded for the reads past the last column.
// rowsel: one bit per contributor of the stage to STORE in this call (warp-uniform); the others are loaded and dropped.
template <int NCH, int FB>
__device__ __forceinline__ void pipe_load_rows(const PipeArgs& a, int kb, const unsigned (&mask)[NCH], int q0,
                                               double (&val)[FB][NCH], unsigned (&pres)[NCH], int lane)
{
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int c = 0; c < NCH; ++c) pres[c] = 0u;
#pragma unroll
    for (int u = 0; u < FB; ++u) {
        int pos = __shfl_sync(0xffffffffu, kb, q0 + u);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const unsigned m = __shfl_sync(0xffffffffu, mask[c], q0 + u);
            val[u][c] = __ldcg(a.L + pos + __popc(m & lt));
            pres[c] |= ((m >> lane) & 1u) << u;
            if (c + 1 < NCH) pos += __popc(m);
        }
    }
}

template <int NCH, int FB>
__device__ __forceinline__ void pipe_store_rows(double w, int q0, const double (&val)[FB][NCH], const unsigned (&pres)[NCH],
                                                double* __restrict__ tp, unsigned rowsel, int cnt, int lane)
{
    constexpr int cap = 32 * NCH;
    double* tpl = tp + lane;
#pragma unroll
    for (int u = 0; u < FB; ++u) {
        const double wq = __shfl_sync(0xffffffffu, w, q0 + u);
        const bool sel = (rowsel >> (q0 + u)) & 1u;                 // warp-uniform
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            if (sel && lane + 32 * c < cnt) tpl[(q0 + u) * cap + 32 * c] = ((pres[c] >> u) & 1u) ? wq * val[u][c] : 0.0;   // lij_dj*AAt[kk], ldlt.c:583
    }
}

// The same for a FEW rows (the second pass of a stage: the contributors that were not final when the stage was prepared,
// as a rule the youngest child alone -- this sits on the critical path between two columns): only the selected rows are
// loaded, up to four at a time.
template <int NCH>
__device__ __forceinline__ void pipe_products_few(const PipeArgs& a, double w, int kb, const unsigned (&mask)[NCH], double* __restrict__ tp,
                                                  unsigned rowsel, int cnt, int lane)
{
    constexpr int cap = 32 * NCH;
    const unsigned lt = (1u << lane) - 1u;
    double* tpl = tp + lane;
    while (rowsel) {
        int qs[4];
        double val[4][NCH];
        unsigned pres[NCH];
        int nrows = 0, qlast = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) pres[c] = 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (rowsel) { qlast = __ffs(rowsel) - 1; rowsel &= rowsel - 1; ++nrows; }
            qs[u] = qlast;                             // a slot past the end repeats the last row
            int pos = __shfl_sync(0xffffffffu, kb, qlast);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const unsigned m = __shfl_sync(0xffffffffu, mask[c], qlast);
                val[u][c] = __ldcg(a.L + pos + __popc(m & lt));
                pres[c] |= ((m >> lane) & 1u) << u;
                if (c + 1 < NCH) pos += __popc(m);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double wq = __shfl_sync(0xffffffffu, w, qs[u]);
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                if (u < nrows && lane + 32 * c < cnt) tpl[qs[u] * cap + 32 * c] = ((pres[c] >> u) & 1u) ? wq * val[u][c] : 0.0;   // lij_dj*AAt[kk], ldlt.c:583
        }
    }
}

template <int NCH, bool PROF>
static __global__ void __launch_bounds__(kPipeWarpsDefault * 32) k_factor_pipe(PipeArgs a)
{
    VBK_DYN_SMEM(raw);
    constexpr int cap = 32 * NCH;
    const int S = a.nstages;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int P = nwarps - 2;                                   // producer warps
    double* tile = reinterpret_cast<double*>(raw);              // [S][kPipeQ][cap]
    double* s_l = tile + (size_t)S * kPipeQ * cap;              // [S][kPipeQ] lij*(lij*dj) of the staged contributors
    double* temp = s_l + S * kPipeQ;                            // [cap] undivided entries of the task
    double* s_red = temp + cap;                                 // [nwarps]
    double* s_dbl = s_red + nwarps;                             // [0] pivot, [1] diagonal chain result
    int* full = reinterpret_cast<int*>(s_dbl + 2);              // [S] group number + 1 staged in the slot
    int* cons = full + S;                                       // [0] groups consumed by the rows, [1] by the pivot chain
    int* s_ctl = cons + 2;                                      // [0] task, [1] keep

    const double thresh = a.epsnum * bits_to_double(a.scal_bits[S_MAXDIAG]);     // ldlt.c:600
    const int wstride = a.nblk + 1;
    unsigned pacc[PROF ? 16 : 1];           // 32-bit per lane: enough for one factorisation, half the registers
#pragma unroll
    for (int u = 0; u < (PROF ? 16 : 1); ++u) pacc[u] = 0;
    long long tlast = PROF ? vbk_clock() : 0;
    const bool profiling = PROF && a.prof != nullptr && lane == 0;
#define VBK_PTICK(slot) do { if (PROF && profiling) { const long long now_ = vbk_clock(); pacc[PROF ? slot : 0] += (unsigned)(now_ - tlast); tlast = now_; } } while (0)

    for (;;) {
        __syncthreads();
        if (tid == 0) s_ctl[0] = atomicAdd(&a.counters[C_NEXT], 1);
        __syncthreads();
        const int t = s_ctl[0];
        if (t >= a.ntasks) break;
        const int i = a.task_col[t], blk = a.task_blk[t], p0 = a.task_pos0[t], cnt = a.task_cnt[t];
        const int nslices = a.col_ntask[i];
        const bool owner = (t == a.col_task0[i] + nslices - 1);
        const int rb = a.rowptr[i], re = a.rowptr[i + 1];
        const int ngroups = (re - rb + kPipeQ - 1) / kPipeQ;

        if (tid < S) full[tid] = 0;
        if (tid < 2) cons[tid] = 0;
        __syncthreads();
        VBK_PTICK(8);
        if (PROF && profiling && tid == 0) pacc[PROF ? 15 : 0] += 1;
        const bool tracing = PROF && a.trace != nullptr && owner && lane == 0;
        if (tracing && tid == 0) {
            long long* tr = a.trace + (size_t)i * 8;
            tr[0] = vbk_globaltimer(); tr[1] = tr[0]; tr[2] = tr[0]; tr[5] = ngroups; tr[6] = nslices;
#ifndef VBK_EMU
            unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); tr[7] = smid;
#endif
        }

        double acc[NCH];
        double kval[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) { acc[c] = 0.0; kval[c] = 0.0; }
        double diagi = 0.0;

        if (warp == 0) {
            // ---- consumer: one register accumulator per row, the reference's additions in the reference's order
#pragma unroll
            for (int c = 0; c < NCH; ++c) if (lane + 32 * c < cnt) kval[c] = __ldcg(&a.L[p0 + lane + 32 * c]);
            // up to kPipeMulti ring stages per wait: one flag check, one release and one pipeline fill of the shared-memory
            // loads per kPipeMulti*32 links instead of per 32
            for (int g = 0; g < ngroups;) {
                const int left = ngroups - g;
                if (left >= kPipeMulti && S > kPipeMulti) {
#pragma unroll
                    for (int u = 0; u < kPipeMulti; ++u) { while (vbk_lds_acquire(&full[(g + u) % S]) != g + u + 1) vbk_pause(); }
                    VBK_PTICK(0);
                    if (tracing && g == 0) a.trace[(size_t)i * 8 + 1] = vbk_globaltimer();
                    {
#pragma unroll
                        for (int u = 0; u < kPipeMulti; ++u) {
                            const double* tp = tile + (size_t)((g + u) % S) * kPipeQ * cap + lane;
#pragma unroll
                            for (int q = 0; q < kPipeQ; ++q) {
#pragma unroll
                                for (int c = 0; c < NCH; ++c) acc[c] += tp[q * cap + 32 * c];      // temp[row] += lij_dj*AAt[kk]
                            }
                        }
                    }
                    g += kPipeMulti;
                } else {
                    while (vbk_lds_acquire(&full[g % S]) != g + 1) vbk_pause();
                    VBK_PTICK(0);
                    if (tracing && g == 0) a.trace[(size_t)i * 8 + 1] = vbk_globaltimer();
                    const double* tp = tile + (size_t)(g % S) * kPipeQ * cap + lane;
                    {
#pragma unroll
                        for (int q = 0; q < kPipeQ; ++q) {
#pragma unroll
                            for (int c = 0; c < NCH; ++c) acc[c] += tp[q * cap + 32 * c];          // temp[row] += lij_dj*AAt[kk]
                        }
                    }
                    g += 1;
                }
                __syncwarp();
                if (lane == 0) vbk_sts_release(&cons[0], g);
                VBK_PTICK(1);
            }
        } else if (warp == 1) {
            // ---- pivot chain (owner slice only): diagi -= lij*lij_dj in the same order (ldlt.c:573)
            if (owner) {
                diagi = __ldcg(&a.diag[i]);
                // the producers stage the rounded products lij*(lij*dj); two ring stages per wait like the consumer --
                // this chain is as long as the rows' chains and holds the column's last slice back if it is any slower
                for (int g = 0; g < ngroups;) {
                    if (ngroups - g >= kPipeMulti && S > kPipeMulti) {
#pragma unroll
                        for (int u = 0; u < kPipeMulti; ++u) { while (vbk_lds_acquire(&full[(g + u) % S]) != g + u + 1) vbk_pause(); }
#pragma unroll
                        for (int u = 0; u < kPipeMulti; ++u) {
                            const double* pl = s_l + ((g + u) % S) * kPipeQ;
#pragma unroll
                            for (int q = 0; q < kPipeQ; ++q) diagi -= pl[q];                       // diag[i] -= lij*lij_dj
                        }
                        g += kPipeMulti;
                    } else {
                        while (vbk_lds_acquire(&full[g % S]) != g + 1) vbk_pause();
                        const double* pl = s_l + (g % S) * kPipeQ;
#pragma unroll
                        for (int q = 0; q < kPipeQ; ++q) diagi -= pl[q];
                        g += 1;
                    }
                    __syncwarp();
                    if (lane == 0) vbk_sts_release(&cons[1], g);
                }
            }
        } else {
            // ---- producers: stage the products of 32 contributors per ring slot.  Software pipeline over the warp's
            // groups: the static structure of the NEXT group (row-list entries, column end, block range) is fetched
            // while the products of the current one are formed, and so are its readiness flag and lij, dj when the
            // contributing columns are already final -- in steady state only the product loads are exposed.
            const int pw = warp - 2;
            const size_t pair0 = (size_t)a.task_pair0[t];
            int k = 0, j = 0, kb = p0, flag = 1;       // a lane without a contributor reads the task's own entries and stores +0.0
            unsigned mask[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) mask[c] = 0u;
            bool valid = false;
            double lij = 0.0, dj = 0.0;
            if (pw < ngroups) {
                const int tq = rb + pw * kPipeQ + lane;
                valid = tq < re;
                if (valid) {
                    k = a.rk[tq]; j = a.rj[tq];
#pragma unroll
                    for (int c = 0; c < NCH; ++c) mask[c] = a.masks[(pair0 + (size_t)pw * kPipeQ + lane) * NCH + c];
                    kb = k + 1;
                    if (blk >= 0) { const int lo = a.winptr[(size_t)j * wstride + blk]; if (lo > kb) kb = lo; }
                    flag = vbk_ld_acquire(&a.col_done[j]);
                    if (flag) { lij = __ldcg(&a.L[k]); dj = __ldcg(&a.diag[j]); }
                }
            }
            for (int g = pw; g < ngroups; g += P) {
                const int st = g % S;
                double* tp = tile + (size_t)st * kPipeQ * cap;
                // (1) row-list entries and presence masks of the warp's next group
                const int gn = g + P;
                const int tqn = rb + gn * kPipeQ + lane;
                const bool nvalid = gn < ngroups && tqn < re;
                int nk = 0, nj = 0;
                unsigned nmask[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) nmask[c] = 0u;
                if (nvalid) {
                    nk = a.rk[tqn]; nj = a.rj[tqn];
#pragma unroll
                    for (int c = 0; c < NCH; ++c) nmask[c] = a.masks[(pair0 + (size_t)gn * kPipeQ + lane) * NCH + c];
                }
                VBK_PTICK(3);
                // (2) the first batch of rows is loaded BEFORE the ring slot is waited for: once the consumer runs, a freed
                //     slot is refilled after the stores alone, not after a round trip to the L2
                constexpr int FB = VBK_PIPE_FB / NCH;
                const unsigned readym = __ballot_sync(0xffffffffu, !valid || flag != 0);
                double val[FB][NCH];
                unsigned pres[NCH];
                pipe_load_rows<NCH, FB>(a, kb, mask, 0, val, pres, lane);
                if (g >= S) {
                    const int need = g - S + 1;
                    while (vbk_ld_volatile(&cons[0]) < need || (owner && vbk_ld_volatile(&cons[1]) < need)) vbk_backoff(a.backoff_ns);
                    __threadfence_block();
                }
                VBK_PTICK(2);
                // (3) first entry of the next group's contributors in the block of this task
                int nlo = 0;
                if (nvalid && blk >= 0) nlo = a.winptr[(size_t)nj * wstride + blk];
                VBK_PTICK(4);
                // (4) products, in two passes: first every contributor whose column was final when the group was
                //     prepared -- all of them in steady state, all but the youngest child when the task runs ahead of the
                //     critical path -- then, as its column becomes final, what is left.  When the last child finishes only
                //     its own rows remain to be staged.
                double w = lij * dj;                                   // lij_dj, ldlt.c:572
                pipe_store_rows<NCH, FB>(w, 0, val, pres, tp, readym, cnt, lane);
#pragma unroll 1
                for (int q0 = FB; q0 < kPipeQ; q0 += FB) {
                    pipe_load_rows<NCH, FB>(a, kb, mask, q0, val, pres, lane);
                    pipe_store_rows<NCH, FB>(w, q0, val, pres, tp, readym, cnt, lane);
                }
                VBK_PTICK(5);
                if (~readym) {
                    if (valid && !flag) {
                        while (vbk_ld_volatile(&a.col_done[j]) == 0) { __nanosleep(20); vbk_pause(); }
                        (void)vbk_ld_acquire(&a.col_done[j]);          // acquire: what is read below is the column's final state
                        lij = __ldcg(&a.L[k]); dj = __ldcg(&a.diag[j]);
                        w = lij * dj;
                    }
                    __syncwarp();
                    pipe_products_few<NCH>(a, w, kb, mask, tp, ~readym, cnt, lane);
                }
                s_l[st * kPipeQ + lane] = lij * w;                     // the pivot chain's operand lij*lij_dj (ldlt.c:573), rounded like there
                VBK_PTICK(6);
                // (6) publish the stage -- before the next group's readiness is looked at: that acquire load stalls the warp for a
                //     round trip, which must not sit between a finished stage and its consumer
                __syncwarp();
                if (lane == 0) vbk_sts_release(&full[st], g + 1);
                // (7) next group: first entry, readiness, lij and dj if the column is final already
                int nkb = nvalid ? nk + 1 : p0, nflag = 1;
                double nlij = 0.0, ndj = 0.0;
                if (nvalid) {
                    if (nlo > nkb) nkb = nlo;
                    nflag = vbk_ld_acquire(&a.col_done[nj]);
                    if (nflag) { nlij = __ldcg(&a.L[nk]); ndj = __ldcg(&a.diag[nj]); }
                }
                VBK_PTICK(7);
                if (PROF && profiling) pacc[PROF ? 14 : 0] += 1;
                k = nk; j = nj; kb = nkb; valid = nvalid; flag = nflag; lij = nlij; dj = ndj;
#pragma unroll
                for (int c = 0; c < NCH; ++c) mask[c] = nmask[c];
            }
        }

        // ---- column entries minus the accumulated updates (ldlt.c:596-599), max|.| of the slice
        if (warp == 0) {
            if (tracing && ngroups > 0) a.trace[(size_t)i * 8 + 2] = vbk_globaltimer();
            double mymax = 0.0;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int s = lane + 32 * c;
                if (s < cnt) {
                    const double v = kval[c] - acc[c];
                    temp[s] = v;
                    if (!owner) a.L[p0 + s] = v;
                    const double av = fabs(v);
                    if (av > mymax) mymax = av;
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) { const double o = __shfl_xor_sync(0xffffffffu, mymax, d); if (o > mymax) mymax = o; }
            if (lane == 0) {
                s_red[0] = mymax;
                if (!owner) a.task_max[t] = mymax;
            }

        } else if (warp == 1 && owner && lane == 0) {
            s_dbl[1] = diagi;
        }
        __syncthreads();
        if (PROF && profiling) tlast = vbk_clock();         // roles that idled at the barrier do not count it
        if (!owner) {
            // undivided entries + slice maximum were written before the barrier: release, then raise col_pub
            if (tid == 0) { vbk_fence_release(); atomicAdd(&a.col_pub[i], 1); }
            VBK_PTICK(9);
            continue;
        }
        VBK_PTICK(9);
        if (tid == 0) {
            if (nslices > 1) {
                while (vbk_ld_acquire(&a.col_pub[i]) != nslices - 1) vbk_pause();
            }
            VBK_PTICK(10);
            if (tracing) a.trace[(size_t)i * 8 + 3] = vbk_globaltimer();
            double piv = s_dbl[1];
            int keep = 1;
            if (fabs(piv) <= thresh) {                                  // dependent pivot, ldlt.c:600-614
                double colmax = s_red[0];
                const int tb = a.col_task0[i];
                for (int u = 0; u + 1 < nslices; ++u) { const double m = __ldcg(&a.task_max[tb + u]); if (m > colmax) colmax = m; }
                atomicAdd(&a.counters[C_NDEP], 1);
                if (colmax < 1.0e+6 * 1.0e-8) keep = 0;
                else piv = (a.perm[i] < a.n_ld ? -1 : 1) * 1.0e-8;
            }
            a.diag[i] = piv;
            if (!keep) a.mark[i] = 0;
            s_dbl[0] = piv;
            s_ctl[1] = keep;
            VBK_PTICK(11);
        }
        __syncthreads();
        if (PROF && profiling) tlast = vbk_clock();
        {
            const double piv = s_dbl[0];
            const int keep = s_ctl[1];
            for (int s = tid; s < cnt; s += nt) a.L[p0 + s] = keep ? temp[s] / piv : 0.0;         // ldlt.c:621-627
            // the other slices' rows: four independent loads in flight per thread
            for (int p = a.kL[i] + tid; p < p0; p += 4 * nt) {
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = (p + u * nt < p0) ? __ldcg(&a.L[p + u * nt]) : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u) if (p + u * nt < p0) a.L[p + u * nt] = keep ? v[u] / piv : 0.0;
            }
        }
        VBK_PTICK(12);
        __syncthreads();
        if (tid == 0) { vbk_fence_release(); atomicExch(&a.col_done[i], 1); if (tracing) a.trace[(size_t)i * 8 + 4] = vbk_globaltimer(); }
        VBK_PTICK(13);
    }
    if (PROF && profiling) {
        // counters 0-1 come from warp 0, 2-7 and 14 from the producers, 8-13 and 15 from thread 0 (warp 0): no double counting
        if (warp == 0) { for (int u = 0; u < 2; ++u) atomicAdd(&a.prof[u], (unsigned long long)pacc[u]);
                         for (int u = 8; u < 14; ++u) atomicAdd(&a.prof[u], (unsigned long long)pacc[u]);
                         atomicAdd(&a.prof[15], (unsigned long long)pacc[15]); }
        else if (warp >= 2) { for (int u = 2; u < 8; ++u) atomicAdd(&a.prof[u], (unsigned long long)pacc[u]);
                              atomicAdd(&a.prof[14], (unsigned long long)pacc[14]); }
    }
#undef VBK_PTICK
}

}  // namespace vbk
