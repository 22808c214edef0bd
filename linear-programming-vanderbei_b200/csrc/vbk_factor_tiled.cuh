// vbk_factor_tiled.cuh -- second-generation STRICT kernels: same bits as vbk_kernels.cuh's simple
// kernels (and as the reference), restructured around the critical path.
//
// Measured on B200 (profiles/r01_timing.md): the reference's left-looking order makes almost every
// column update a link of one long dependent chain (pilot87: 520 273 of 582 242 contributor steps
// are on the critical path), so what matters is the LATENCY per contributor step.  The simple
// kernel pays one __syncthreads plus two dependent L2 round trips per contributor.  Here:
//
//  * contributors are taken in batches of up to 128: their products w_q*L[kk] are scattered into a
//    zero-initialised shared-memory tile[q][slot] by ALL threads at once (all loads independent),
//    then each slot's owner adds its column of the tile in contributor order -- the same rounded
//    adds in the same order as the reference (adding the +0.0 of an absent entry is exact and the
//    accumulator is never -0.0), two block barriers per BATCH instead of one per contributor;
//  * a column longer than `whole_cap` rows is cut by global row blocks into several tasks that run
//    on different CTAs; each contributing column finds its entries for a block through the
//    winptr table (vbk_symbolic.h), the first slice owns the pivot and publishes it;
//  * forward substitution waits per COLUMN (a flag per finished row) instead of per etree child
//    set, so a row consumes z[j] as soon as it exists and the dense tail pipelines; the dependent
//    subtract chains run through shared memory instead of warp shuffles.
#pragma once
#include "vbk_kernels.cuh"

namespace vbk {

constexpr int kTileMaxBatch = 128;
#ifdef VBK_EMU
constexpr int kTiledThreads = 32;
#else
constexpr int kTiledThreads = 256;
#endif

struct TiledArgs {
    int N, n_ld, ntasks, tile_doubles, temp_cap;
    const int* kL; const int* iL; double* L; double* diag; int* mark;
    const int* rowptr; const int* rk; const int* rj;
    const int* parent; const int* perm;
    const int* task_col; const int* task_blk; const int* task_pos0; const int* task_cnt;
    const int* col_task0; const int* col_ntask;
    const int* winptr; int nblk, rowblk, slice_row0;
    int* pend;          // unfinished etree children per column
    int* col_left;      // unfinished tasks per column
    int* col_ready;     // slices of the column that published their max|v|
    int* piv_flag;      // 1 once the pivot of the column is published
    double* piv_val; int* piv_keep; double* task_max;
    int* counters; const unsigned long long* scal_bits; double epsnum;
    int* slotmap;       // [gridDim.x][N]
    unsigned long long* prof;   // optional [8] cycle counters per phase (thread 0 of every CTA), or null
    // fast mode (vbk_fast.cuh): phase 0 = plain factorisation of tasks [0, ntasks); phase 2 = Schur
    // assembly of the trailing dense window: tasks [task_base, task_base+ntasks) of columns >= T take
    // only contributors j < T, need no dependency waits, and write K - sum into the dense scratch
    // Sw (column-major, leading dimension ldw) instead of pivoting
    int phase, task_base, T, ldw;
    double* Sw; double* wmag;
};

// per-launch reset of the tiled factor's dataflow state
static __global__ void k_tiled_reset(int N, const int* __restrict__ nchild, const int* __restrict__ col_ntask,
                                     int* __restrict__ pend, int* __restrict__ col_left, int* __restrict__ col_ready,
                                     int* __restrict__ piv_flag, int* __restrict__ counters)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) {
        pend[t] = nchild[t];
        col_left[t] = col_ntask[t];
        col_ready[t] = 0;
        piv_flag[t] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { counters[C_NEXT] = 0; counters[C_NDEP] = 0; }
}

static __global__ void __launch_bounds__(kTiledThreads) k_factor_tiled(TiledArgs a)
{
    VBK_DYN_SMEM(raw);
    double* tile = reinterpret_cast<double*>(raw);
    double* temp = tile + a.tile_doubles;
    double* s_w = temp + a.temp_cap;
    double* s_l = s_w + kTileMaxBatch;
    double* s_red = s_l + kTileMaxBatch;              // [kTiledThreads]
    double* s_dbl = s_red + kTiledThreads;            // [0] pivot
    int* s_kb = reinterpret_cast<int*>(s_dbl + 2);
    int* s_off = s_kb + kTileMaxBatch;                // exclusive prefix of the batch's entry counts, [B+1]
    int* blockmap = s_off + kTileMaxBatch + 1;        // [rowblk] row-in-block -> slot
    int* s_ctl = blockmap + a.rowblk;                 // [0] task, [1] dependent, [2] keep

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    int* myslot = a.slotmap + (size_t)blockIdx.x * a.N;
    const double thresh = a.epsnum * bits_to_double(a.scal_bits[S_MAXDIAG]);     // ldlt.c:600

    for (int q = tid; q < a.tile_doubles; q += nt) tile[q] = 0.0;   // consumed entries are re-zeroed below

    // optional phase profile: 0 claim+init, 1 wait for children, 2 stage, 3 scan, 4 scatter,
    // 5 accumulate, 6 pivot (reduce + hand-off), 7 write+release
    long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = vbk_clock();
#define VBK_TICK(slot) do { if (a.prof && tid == 0) { long long now_ = vbk_clock(); pacc[slot] += now_ - tlast; tlast = now_; } } while (0)

    for (;;) {
        __syncthreads();
        if (tid == 0) s_ctl[0] = atomicAdd(&a.counters[C_NEXT], 1);
        __syncthreads();
        if (s_ctl[0] >= a.ntasks) break;
        const int t = s_ctl[0] + a.task_base;
        const int i = a.task_col[t], blk = a.task_blk[t], p0 = a.task_pos0[t], cnt = a.task_cnt[t];
        const int nslices = a.col_ntask[i];
        const bool first = (t == a.col_task0[i]);
        const int bs = a.slice_row0 + (blk < 0 ? 0 : blk) * a.rowblk;            // first row of the block
        const int* wp_base = a.winptr;
        const int wstride = a.nblk + 1;

        for (int s = tid; s < cnt; s += nt) {
            temp[s] = 0.0;
            int row = a.iL[p0 + s];
            if (blk < 0) myslot[row] = s; else blockmap[row - bs] = s;
        }
        int B = cnt > 0 ? a.tile_doubles / cnt : kTileMaxBatch;
        if (B > kTileMaxBatch) B = kTileMaxBatch;

        VBK_TICK(0);
        if (tid == 0 && a.phase != 2) {   // every etree child final => every contributing column final
            while (vbk_ld_volatile(&a.pend[i]) != 0) __nanosleep(64);
            __threadfence();
        }
        __syncthreads();
        VBK_TICK(1);

        double diagi = 0.0, dmag = 0.0;
        if (first && tid == nt - 1) { diagi = __ldcg(&a.diag[i]); dmag = fabs(diagi); }
        const int rb = a.rowptr[i], re = a.rowptr[i + 1];
        for (int t0 = rb; t0 < re; t0 += B) {
            const int nb = (re - t0 < B) ? (re - t0) : B;
            // -- stage the batch: lij, lij*dj and the entry range of each contributor for this task
            for (int q = tid; q < nb; q += nt) {
                const int k = a.rk[t0 + q], j = a.rj[t0 + q];
                double lij = 0.0, dj = 0.0;
                int kb = 0, ke = 0;
                if (a.phase != 2 || j < a.T) {       // Schur assembly: window columns contribute later
                    lij = __ldcg(&a.L[k]);
                    dj = __ldcg(&a.diag[j]);
                    kb = k + 1; ke = a.kL[j + 1];
                    if (blk >= 0) {
                        const int* wp = wp_base + (size_t)j * wstride;
                        int lo = wp[blk], hi = wp[blk + 1];
                        if (lo > kb) kb = lo;
                        if (hi < ke) ke = hi;
                    }
                }
                s_l[q] = lij;
                s_w[q] = lij * dj;                                     // ldlt.c:572
                s_kb[q] = kb;
                s_off[q + 1] = (ke > kb) ? (ke - kb) : 0;
            }
            if (tid == 0) s_off[0] = 0;
            __syncthreads();
            VBK_TICK(2);
            if (tid < 32) {                                            // inclusive scan of <=128 counts
                int v[4], sum = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) { int idx = lane * 4 + u; v[u] = (idx < nb) ? s_off[idx + 1] : 0; sum += v[u]; }
                int incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += o; }
                int run = incl - sum;
#pragma unroll
                for (int u = 0; u < 4; ++u) { int idx = lane * 4 + u; run += v[u]; if (idx < nb) s_off[idx + 1] = run; }
            }
            __syncthreads();
            VBK_TICK(3);
            // -- scatter all products of the batch into the tile; every load is independent
            const int E = s_off[nb];
            constexpr int U = 8;       // elements in flight per thread: the loads of a group overlap
            for (int e0 = tid; e0 < E; e0 += nt * U) {
                int qq[U], kk[U], row[U], sl[U];
                double val[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int e = e0 + u * nt;
                    kk[u] = -1;
                    qq[u] = 0;
                    if (e < E) {
                        int lo = 0, hi = nb;                           // largest q with s_off[q] <= e
                        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_off[mid] <= e) lo = mid; else hi = mid; }
                        qq[u] = lo;
                        kk[u] = s_kb[lo] + (e - s_off[lo]);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) if (kk[u] >= 0) { row[u] = a.iL[kk[u]]; val[u] = __ldcg(&a.L[kk[u]]); }
#pragma unroll
                for (int u = 0; u < U; ++u) if (kk[u] >= 0) sl[u] = (blk < 0) ? myslot[row[u]] : blockmap[row[u] - bs];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (kk[u] >= 0) tile[qq[u] * cnt + sl[u]] = s_w[qq[u]] * val[u];   // lij_dj*AAt[kk], ldlt.c:583
            }
            __syncthreads();
            VBK_TICK(4);
            // -- each slot's owner replays the batch in contributor order; the tile is left zeroed
            for (int s = tid; s < cnt; s += nt) {
                double acc = temp[s];
#pragma unroll 8
                for (int q = 0; q < nb; ++q) { acc += tile[q * cnt + s]; tile[q * cnt + s] = 0.0; }
                temp[s] = acc;
            }
            if (first && tid == nt - 1)
                for (int q = 0; q < nb; ++q) { double p = s_l[q] * s_w[q]; diagi -= p; if (fabs(p) > dmag) dmag = fabs(p); }   // ldlt.c:573
            __syncthreads();
            VBK_TICK(5);
        }

        if (a.phase == 2) {
            // Schur complement of the sparse part, written densely: Sw[r-T, i-T] = K[r,i] - sum; the
            // diagonal entry carries the magnitude of its terms for the fast-mode zero-pivot test
            const size_t ci = (size_t)(i - a.T);
            for (int s = tid; s < cnt; s += nt) {
                const int row = a.iL[p0 + s];
                a.Sw[(size_t)(row - a.T) + ci * a.ldw] = __ldcg(&a.L[p0 + s]) - temp[s];
            }
            if (first && tid == nt - 1) { a.Sw[ci + ci * a.ldw] = diagi; a.wmag[ci] = dmag; }
            continue;
        }

        // L[:,i] -= temp (ldlt.c:596-599); max|.| of this slice for the dependent-pivot rule
        double mymax = 0.0;
        for (int s = tid; s < cnt; s += nt) {
            double v = __ldcg(&a.L[p0 + s]) - temp[s];
            temp[s] = v;
            double av = fabs(v);
            if (av > mymax) mymax = av;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { double o = __shfl_xor_sync(0xffffffffu, mymax, d); if (o > mymax) mymax = o; }
        if (lane == 0) s_red[tid >> 5] = mymax;
        if (first && tid == nt - 1) s_dbl[1] = diagi;
        __syncthreads();
        if (tid == 0) {
            for (int wv = 1; wv < (nt >> 5); ++wv) if (s_red[wv] > s_red[0]) s_red[0] = s_red[wv];
            if (nslices > 1) {
                a.task_max[t] = s_red[0];
                __threadfence();
                atomicAdd(&a.col_ready[i], 1);
            }
            double piv;
            int keep = 1;
            if (first) {
                piv = s_dbl[1];
                if (fabs(piv) <= thresh) {                              // dependent pivot, ldlt.c:600-614
                    double colmax = s_red[0];
                    if (nslices > 1) {
                        while (vbk_ld_volatile(&a.col_ready[i]) != nslices) __nanosleep(64);
                        __threadfence();
                        const int tb = a.col_task0[i];
                        for (int u = 0; u < nslices; ++u) { double m = __ldcg(&a.task_max[tb + u]); if (m > colmax) colmax = m; }
                    }
                    atomicAdd(&a.counters[C_NDEP], 1);
                    if (colmax < 1.0e+6 * 1.0e-8) keep = 0;
                    else piv = (a.perm[i] < a.n_ld ? -1 : 1) * 1.0e-8;
                }
                a.diag[i] = piv;
                if (!keep) a.mark[i] = 0;
                if (nslices > 1) {
                    a.piv_val[i] = piv;
                    a.piv_keep[i] = keep;
                    __threadfence();
                    atomicExch(&a.piv_flag[i], 1);
                }
            } else {
                while (vbk_ld_volatile(&a.piv_flag[i]) == 0) __nanosleep(64);
                __threadfence();
                piv = __ldcg(&a.piv_val[i]);
                keep = __ldcg(&a.piv_keep[i]);
            }
            s_dbl[0] = piv;
            s_ctl[2] = keep;
        }
        __syncthreads();
        VBK_TICK(6);
        const double piv = s_dbl[0];
        const int keep = s_ctl[2];
        for (int s = tid; s < cnt; s += nt) a.L[p0 + s] = keep ? temp[s] / piv : 0.0;   // ldlt.c:621-627

        __threadfence();
        __syncthreads();
        if (tid == 0) {
            if (atomicSub(&a.col_left[i], 1) == 1) {     // last slice of the column: release the parent
                __threadfence();
                int p = a.parent[i];
                if (p >= 0) atomicSub(&a.pend[p], 1);
            }
        }
        VBK_TICK(7);
    }
    if (a.prof && tid == 0)
        for (int u = 0; u < 8; ++u) atomicAdd(&a.prof[u], (unsigned long long)pacc[u]);
#undef VBK_TICK
}

// --------------------------------------------------------------------------------------------
// Forward / backward substitution with per-column completion flags (rawsolve, ldlt.c:433-505).
// done[j] != 0  <=>  z[j] is final.  One warp per row (forward, ascending claims) or per column
// (backward, descending claims); dependencies always point to indices claimed earlier.
// --------------------------------------------------------------------------------------------
struct FlagSolveArgs {
    int N;
    int nclaim;   // rows [0, nclaim) forward / columns nclaim-1..0 backward (N, or the window start in fast mode)
    int fast;     // 1: sums may be re-associated (tree reductions, FMA)
    const int* kL; const int* iL; const double* L; const double* diag; const int* mark;
    const int* rowptr; const int* rk; const int* rj;   // ascending row lists
    const int* parent;
    double* z;
    int* done; int* counters;
    const unsigned long long* scal_bits;
    double epssol;
};

static __global__ void k_flags_reset(int N, int* __restrict__ done, int* __restrict__ counters, int set_consistent)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) done[t] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counters[C_NEXT] = 0;
        if (set_consistent) counters[C_CONSISTENT] = 1;
    }
}

// acc - p[0] - p[1] - ... strictly left to right; a full batch is loaded first (32 independent
// shared-memory reads) so that only the 32 dependent subtractions remain on the critical path
__device__ __forceinline__ double chain_sub(double acc, const double* p, int cnt)
{
    if (cnt == 32) {
        double v[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = p[q];
#pragma unroll
        for (int q = 0; q < 32; ++q) acc = acc - v[q];
    } else {
        for (int q = 0; q < cnt; ++q) acc = acc - p[q];
    }
    return acc;
}

__device__ __forceinline__ double flag_solve_eps(const FlagSolveArgs& a) {
    return a.counters[C_NDEP] ? a.epssol * bits_to_double(a.scal_bits[S_ZMAX]) : 0.0;   // ldlt.c:446
}

static __global__ void __launch_bounds__(kSolveThreads) k_fwd_flags(FlagSolveArgs a)
{
    VBK_DYN_SMEM(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sp = reinterpret_cast<double*>(raw) + warp * 32;                    // products
    const double eps = flag_solve_eps(a);
    for (;;) {
        int r = 0;
        if (lane == 0) r = atomicAdd(&a.counters[C_NEXT], 1);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.nclaim) break;
        double acc = a.z[r];                         // right-hand side entry, written before the launch
        const int rb = a.rowptr[r], re = a.rowptr[r + 1];
        for (int t0 = rb; t0 < re; t0 += 32) {
            const int t = t0 + lane;
            double p = 0.0;        // an unmarked column contributes nothing (ldlt.c:455); x - (+0.0) == x
            if (t < re) {
                const int j = a.rj[t];
                const double l = a.L[a.rk[t]];
                while (vbk_ld_volatile(&a.done[j]) == 0) __nanosleep(20);
                __threadfence();
                if (a.mark[j]) p = l * __ldcg(&a.z[j]);
            }
            sp[lane] = p;
            __syncwarp();
            if (lane == 0) acc = chain_sub(acc, sp, (re - t0 < 32) ? (re - t0) : 32);   // z[row] -= AAt[k]*beta
            __syncwarp();
        }
        if (lane == 0) {
            if (a.mark[r]) a.z[r] = acc;
            else if (fabs(acc) > eps) { a.z[r] = acc; a.counters[C_CONSISTENT] = 0; }
            else a.z[r] = 0.0;
            __threadfence();
            atomicExch(&a.done[r], 1);
        }
        __syncwarp();
    }
}

static __global__ void __launch_bounds__(kSolveThreads) k_bwd_flags(FlagSolveArgs a)
{
    VBK_DYN_SMEM(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sp = reinterpret_cast<double*>(raw) + warp * 128;
    const double eps = flag_solve_eps(a);
    for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(&a.counters[C_NEXT], 1);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= a.nclaim) break;
        const int i = a.nclaim - 1 - c;
        const int par = a.parent[i];
        if (lane == 0 && par >= 0 && par < a.nclaim) {   // columns >= nclaim were solved before this launch
            while (vbk_ld_volatile(&a.done[par]) == 0) __nanosleep(20);
            __threadfence();
        }
        __syncwarp();
        double beta = a.z[i];                        // z[i] after the diagonal sweep (previous launch)
        if (a.mark[i]) {
            int kb = a.kL[i];
            const int ke = a.kL[i + 1];
            if (a.fast) {
                // fast mode: the order of the sum is free -- per-lane partial sums, then a shuffle tree
                // four entries per lane in flight: the column's row index -> z loads are two dependent L2 round trips
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                for (int k = kb + lane; k < ke; k += 128) {
                    const int k1 = k + 32, k2 = k + 64, k3 = k + 96;
                    const int i0 = a.iL[k], i1 = k1 < ke ? a.iL[k1] : -1, i2 = k2 < ke ? a.iL[k2] : -1, i3 = k3 < ke ? a.iL[k3] : -1;
                    s0 = fma(a.L[k], __ldcg(&a.z[i0]), s0);
                    if (i1 >= 0) s1 = fma(a.L[k1], __ldcg(&a.z[i1]), s1);
                    if (i2 >= 0) s2 = fma(a.L[k2], __ldcg(&a.z[i2]), s2);
                    if (i3 >= 0) s3 = fma(a.L[k3], __ldcg(&a.z[i3]), s3);
                }
                double s = (s0 + s1) + (s2 + s3);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                beta -= s;
                kb = ke;          // nothing left for the ordered path below
            }
            // 128 entries per round: the raw operands of the NEXT round are loaded before the dependent
            // subtract chain of the current one and multiplied only afterwards, so the two L2 round
            // trips (iL -> z) hide behind the chain instead of stalling it
            double ln[4], zn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int k = kb + u * 32 + lane;
                ln[u] = 0.0; zn[u] = 0.0;
                if (k < ke) { ln[u] = a.L[k]; zn[u] = __ldcg(&a.z[a.iL[k]]); }
            }
            for (int k0 = kb; k0 < ke; k0 += 128) {
#pragma unroll
                for (int u = 0; u < 4; ++u) sp[u * 32 + lane] = ln[u] * zn[u];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + 128 + u * 32 + lane;
                    ln[u] = 0.0; zn[u] = 0.0;
                    if (k < ke) { ln[u] = a.L[k]; zn[u] = __ldcg(&a.z[a.iL[k]]); }
                }
                __syncwarp();
                if (lane == 0) {
                    const int rem = ke - k0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = rem - u * 32;
                        if (c > 0) beta = chain_sub(beta, sp + u * 32, c < 32 ? c : 32);     // ldlt.c:494
                    }
                }
                __syncwarp();
            }
            if (lane == 0) a.z[i] = beta;
        } else if (lane == 0) {
            if (fabs(beta) > eps) a.counters[C_CONSISTENT] = 0;
            else a.z[i] = 0.0;
        }
        if (lane == 0) {
            __threadfence();
            atomicExch(&a.done[i], 1);
        }
        __syncwarp();
    }
}

}  // namespace vbk
