// vbk_sparse_level.cuh -- fast mode, the sparse columns j < T factorised level by level of their elimination tree.
//
//   k_sparse_level    one launch per etree level of the sparse columns (levels computed on the host from the parent
//                     array restricted to j < T); all columns of a level are independent.  One WARP per column,
//                     left-looking: the column (K values scattered by k_scatter) sits in shared memory with its row
//                     indices, the contributors of the row list come in batches of 32 (one per lane: L_jk, d_k, tail
//                     range), tails are applied four contributors at a time so that their loads overlap, a target
//                     slot is found by binary search of the column's own sorted row indices.
//   k_sparse_level_heavy  the columns of a level whose contributors' tails add up to more than a warp should chew
//                     (the top of the sparse tree: ~1000 window rows per column, ~100 contributors): one CTA each.
// Both apply the contributors of a column in the reference's own list order (sigma) with
// separately rounded products and sums, the diagonal as one serial chain, and the reference's exact pivot rule
// (ldlt.c:600-614): like the strict task kernel (vbk_strict_factor.cuh) they reproduce the sparse columns of the
// reference's factor BIT FOR BIT -- parallelism is across columns and across the entries of a tail, never inside a sum.
#pragma once
#include "vbk_window_solve.cuh"

namespace vbk {

// (compiled for the host thread emulator too -- tests/test_emu.py checks the bit-exactness claim on the CPU -- with
// smaller CTAs: every emulated thread is an OS thread)
constexpr int kSpWarps = 4;
#ifndef VBK_SP_HEAVY_THREADS
#define VBK_SP_HEAVY_THREADS 1024
#endif
#ifdef VBK_EMU
constexpr int kSpHeavyThreads = 64;
constexpr int kSpHeavyBatch = 64;
#else
constexpr int kSpHeavyThreads = VBK_SP_HEAVY_THREADS;
constexpr int kSpHeavyBatch = VBK_SP_HEAVY_THREADS;
#endif

struct SparseLevelArgs {
    const int* cols; int ncols;        // the columns of this launch (one etree level, light or heavy share)
    int n_ld, cap, T, W;               // cap: shared-memory slots per column (>= longest sparse column)
    const int* kL; const int* iL; double* L; double* diag; int* mark; const int* perm;
    const int* rowptr; const int* rk; const int* rj;       // row lists in the reference's sigma order (rk_sig / rj_sig)
    int* counters; const unsigned long long* scal_bits; double epsnum;
};

__device__ __forceinline__ int sp_find(const int* rows, int c, int r)   // position of r in rows[0..c) (present by construction)
{
    int lo = 0, hi = c - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rows[mid] < r) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// pivot rule of lltnum (ldlt.c:600-614), exact form: the sums above are the reference's, bit for bit
__device__ __forceinline__ double sp_pivot(const SparseLevelArgs& a, int j, double d, double colmax, int* keep, bool leader)
{
    const double thresh = a.epsnum * bits_to_double(a.scal_bits[S_MAXDIAG]);
    *keep = 1;
    if (fabs(d) <= thresh) {
        if (leader) atomicAdd(&a.counters[C_NDEP], 1);
        if (colmax < 1.0e+6 * 1.0e-8) *keep = 0;
        else d = (a.perm[j] < a.n_ld ? -1 : 1) * 1.0e-8;
    }
    return d;
}

// light columns: one warp per column
static __global__ void __launch_bounds__(kSpWarps * 32) k_sparse_level(SparseLevelArgs a)
{
    VBK_DYN_SMEM(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* val = reinterpret_cast<double*>(raw) + (size_t)warp * 2 * a.cap;       // K values of the column
    double* acc = val + a.cap;                                                      // the reference's temp[]
    int* rows = reinterpret_cast<int*>(reinterpret_cast<double*>(raw) + (size_t)kSpWarps * 2 * a.cap) + (size_t)warp * a.cap;
    for (int idx = blockIdx.x * kSpWarps + warp; idx < a.ncols; idx += gridDim.x * kSpWarps) {
        const int j = a.cols[idx];
        const int p0 = a.kL[j], c = a.kL[j + 1] - p0;
        for (int s = lane; s < c; s += 32) { rows[s] = a.iL[p0 + s]; val[s] = a.L[p0 + s]; acc[s] = 0.0; }
        double d = a.diag[j];
        __syncwarp();
        const int t1 = a.rowptr[j + 1];
        for (int tb = a.rowptr[j]; tb < t1; tb += 32) {
            const int t = tb + lane;
            double w = 0.0, p = 0.0;
            int e0 = 0, e1 = 0;
            if (t < t1) {
                const int kc = a.rj[t], k = a.rk[t];
                const double lij = a.L[k];
                w = lij * a.diag[kc];                      // lij_dj, ldlt.c:572
                p = lij * w;
                e0 = k + 1; e1 = a.kL[kc + 1];
            }
            const int nb = (t1 - tb < 32) ? (t1 - tb) : 32;
            for (int b = 0; b < nb; ++b) d = d - __shfl_sync(0xffffffffu, p, b);    // diagi -= lij*lij_dj in list order, ldlt.c:573
            for (int b = 0; b < nb; b += 4) {
                // four contributors at a time: their first 32 tail entries are loaded together, applied in order
                int rr[4]; double vv[4], ww[4]; int s0[4], s1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ww[u] = __shfl_sync(0xffffffffu, w, (b + u) & 31);
                    s0[u] = __shfl_sync(0xffffffffu, e0, (b + u) & 31);
                    s1[u] = __shfl_sync(0xffffffffu, e1, (b + u) & 31);
                    if (b + u >= nb) s1[u] = s0[u];
                    const int e = s0[u] + lane;
                    rr[u] = -1; vv[u] = 0.0;
                    if (e < s1[u]) { rr[u] = a.iL[e]; vv[u] = a.L[e]; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (rr[u] >= 0) { const int slot = sp_find(rows, c, rr[u]); acc[slot] = acc[slot] + ww[u] * vv[u]; }   // ldlt.c:583,588
                    __syncwarp();
                    for (int e = s0[u] + 32 + lane; e < s1[u]; e += 32) {         // long tails: the rest, 32 at a time
                        const int slot = sp_find(rows, c, a.iL[e]);
                        acc[slot] = acc[slot] + ww[u] * a.L[e];
                    }
                    __syncwarp();
                }
            }
        }
        double cm = 0.0;
        for (int s = lane; s < c; s += 32) { const double v = val[s] - acc[s]; val[s] = v; cm = fmax(cm, fabs(v)); }   // ldlt.c:598
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, sft));
        int keep;
        d = sp_pivot(a, j, d, cm, &keep, lane == 0);
        if (lane == 0) { a.diag[j] = d; if (!keep) a.mark[j] = 0; }
        for (int s = lane; s < c; s += 32) a.L[p0 + s] = keep ? val[s] / d : 0.0;                                      // ldlt.c:621-627
        __syncwarp();
    }
}

// heavy columns (many contributors with long tails -- the top of the sparse elimination tree, ~1000 window rows per
// column): one CTA per column, contributors still one after the other (the reference's order per target entry), all
// threads on one tail; window rows find their slot through a shared-memory map, the few sparse rows by binary search.
static __global__ void __launch_bounds__(kSpHeavyThreads) k_sparse_level_heavy(SparseLevelArgs a)
{
    VBK_DYN_SMEM(raw);
    double* val = reinterpret_cast<double*>(raw);            // [cap]
    double* acc = val + a.cap;                               // [cap]
    double* sw = acc + a.cap;                                // [batch] lij_dj
    double* sp = sw + kSpHeavyBatch;                         // [batch] lij*lij_dj
    double* red = sp + kSpHeavyBatch;                        // [threads]
    int* rows = reinterpret_cast<int*>(red + kSpHeavyThreads);   // [cap]
    int* winmap = rows + a.cap;                              // [W]
    int* se0 = winmap + a.W;                                 // [batch]
    int* se1 = se0 + kSpHeavyBatch;                          // [batch]
    int* s_ncs = se1 + kSpHeavyBatch;                        // [1]
    const int tid = threadIdx.x;
    for (int idx = blockIdx.x; idx < a.ncols; idx += gridDim.x) {
        const int j = a.cols[idx];
        const int p0 = a.kL[j], c = a.kL[j + 1] - p0;
        __syncthreads();
        if (tid == 0) *s_ncs = c;
        __syncthreads();
        for (int s = tid; s < c; s += kSpHeavyThreads) {
            const int r = a.iL[p0 + s];
            rows[s] = r; val[s] = a.L[p0 + s]; acc[s] = 0.0;
            if (r >= a.T) {
                winmap[r - a.T] = s;
                if (s == 0 || a.iL[p0 + s - 1] < a.T) *s_ncs = s;        // first window row: rows before it are sparse
            }
        }
        double d = a.diag[j];
        __syncthreads();
        const int ncs = *s_ncs;
        const int t1 = a.rowptr[j + 1];
        for (int tb = a.rowptr[j]; tb < t1; tb += kSpHeavyBatch) {
            const int t = tb + tid;
            if (t < t1 && tid < kSpHeavyBatch) {
                const int kc = a.rj[t], k = a.rk[t];
                const double lij = a.L[k];
                const double w = lij * a.diag[kc];
                sw[tid] = w; sp[tid] = lij * w; se0[tid] = k + 1; se1[tid] = a.kL[kc + 1];
            }
            __syncthreads();
            const int nb = (t1 - tb < kSpHeavyBatch) ? (t1 - tb) : kSpHeavyBatch;
            for (int b = 0; b < nb; ++b) d = d - sp[b];                    // every thread keeps its own copy of diagi
            // software pipeline over the contributors, four at a time: the tid-th tail entries of the NEXT four are in
            // flight while these four are applied (one L2 round trip per four contributors instead of per contributor)
            int rq[4][2]; double vq[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    rq[u][h] = -1; vq[u][h] = 0.0;
                    if (u < nb) { const int e = se0[u] + tid + h * kSpHeavyThreads; if (e < se1[u]) { rq[u][h] = a.iL[e]; vq[u][h] = a.L[e]; } }
                }
            for (int b = 0; b < nb; b += 4) {
                int rc[4][2]; double vc[4][2];
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int h = 0; h < 2; ++h) { rc[u][h] = rq[u][h]; vc[u][h] = vq[u][h]; }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        rq[u][h] = -1;
                        if (b + 4 + u < nb) {
                            const int e = se0[b + 4 + u] + tid + h * kSpHeavyThreads;
                            if (e < se1[b + 4 + u]) { rq[u][h] = a.iL[e]; vq[u][h] = a.L[e]; }
                        }
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (b + u >= nb) break;                                   // uniform
                    const double wb = sw[b + u];
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (rc[u][h] >= 0) {
                            const int slot = (rc[u][h] >= a.T) ? winmap[rc[u][h] - a.T] : sp_find(rows, ncs, rc[u][h]);
                            acc[slot] = acc[slot] + wb * vc[u][h];
                        }
                    for (int e = se0[b + u] + 2 * kSpHeavyThreads + tid; e < se1[b + u]; e += kSpHeavyThreads) {
                        const int r = a.iL[e];
                        const int slot = (r >= a.T) ? winmap[r - a.T] : sp_find(rows, ncs, r);
                        acc[slot] = acc[slot] + wb * a.L[e];
                    }
                    __syncthreads();
                }
            }
        }
        double cm = 0.0;
        for (int s = tid; s < c; s += kSpHeavyThreads) { const double v = val[s] - acc[s]; val[s] = v; cm = fmax(cm, fabs(v)); }
        red[tid] = cm;
        __syncthreads();
        for (int sft = kSpHeavyThreads / 2; sft > 0; sft >>= 1) { if (tid < sft) red[tid] = fmax(red[tid], red[tid + sft]); __syncthreads(); }
        int keep;
        d = sp_pivot(a, j, d, red[0], &keep, tid == 0);
        if (tid == 0) { a.diag[j] = d; if (!keep) a.mark[j] = 0; }
        for (int s = tid; s < c; s += kSpHeavyThreads) a.L[p0 + s] = keep ? val[s] / d : 0.0;
    }
}

}  // namespace vbk
