// vbk_strict_solve.cuh -- STRICT backward substitution (rawsolve, reference src/ipo/ldlt.c:490-502) as a
// producer / consumer pipeline per column.
//
// The reference computes z[i] -= L[k]*z[row_k] for the entries k of column i in ascending k, a chain of rounded
// subtractions that STARTS with the row that became final last (the first stored row of a column is its
// elimination-tree parent).  So a column's chain cannot start before its parent is final and is c_i links long:
// one backward sweep is a critical path of ~Lnz dependent FP64 subtractions (dfl001: 7.0 M), 8 cycles each.  The
// round-1 kernel (k_bwd_flags, vbk_flag_solve.cuh) ran loads, products and chain of a 128-entry round in
// one warp, one after the other, and started a column's loads only when the parent had finished (~19 cycles per link).
// Here a CTA is a pipeline over one column:
//   producer warps  form the products L[k]*z[row_k] of 128 entries at a time into a ring of shared-memory stages;
//                   they wait per ENTRY for its row to be final (done[row]), so everything but the parent's product
//                   is staged while the parent is still running;
//   consumer        lane 0 of warp 0 subtracts a stage's products in order, 32 at a time from registers.
// Columns are claimed in descending order by persistent CTAs; a column only waits for rows with larger indices, which
// were claimed earlier: no deadlock for any grid.
#pragma once
#include "vbk_flag_solve.cuh"

namespace vbk {

constexpr int kBwdChunk = 128;        // entries per ring stage (4 per producer lane)
#ifdef VBK_EMU
constexpr int kBwdStages = 3;
constexpr int kBwdWarps = 3;
#else
constexpr int kBwdStages = 12;
constexpr int kBwdWarps = 8;          // consumer + 7 producers
#endif

struct BwdPipeArgs {
    int N;
    int nclaim;        // columns nclaim-1 .. 0
    const int* kL; const int* iL; const double* L; const int* mark;
    double* z;
    int* done;         // [N] 1 once z[i] is final (zeroed before the launch; columns >= nclaim count as final)
    int* counters; const unsigned long long* scal_bits; double epssol;
    double* z2 = nullptr;   // second right-hand side (k_bwd_pipe<2>)
    int rhs = 0;            // scalar slots of z when there is one right-hand side
};

inline size_t bwd_pipe_smem_bytes(int nrhs = 1) { return sizeof(double) * nrhs * kBwdStages * kBwdChunk + sizeof(int) * (kBwdStages + 4); }

// NRHS = 2: the even lanes of the consumer warp run the chain of z, the odd lanes the chain of z2 (same instructions,
// two shared-memory addresses per load); the producers stage both products of an entry behind one wait.
template <int NRHS>
static __global__ void __launch_bounds__(kBwdWarps * 32) k_bwd_pipe(BwdPipeArgs a)
{
    VBK_DYN_SMEM(raw);
    double* ring = reinterpret_cast<double*>(raw);              // [NRHS][kBwdStages][kBwdChunk]
    int* full = reinterpret_cast<int*>(ring + NRHS * kBwdStages * kBwdChunk);   // [kBwdStages] chunk number + 1
    int* cons = full + kBwdStages;                              // [0] chunks consumed
    int* s_ctl = cons + 1;                                      // [0] claim
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int P = kBwdWarps - 1;
    const int myrhs = (NRHS == 2) ? (lane & 1) : a.rhs;
    const double eps = a.counters[C_NDEP] ? a.epssol * bits_to_double(a.scal_bits[S_ZMAX + kRhsSlotStride * myrhs]) : 0.0;   // ldlt.c:446
    double* const zmine = (NRHS == 2 && (lane & 1)) ? a.z2 : a.z;
    const double* const myring = ring + ((NRHS == 2 && (lane & 1)) ? kBwdStages * kBwdChunk : 0);

    for (;;) {
        __syncthreads();
        if (tid == 0) s_ctl[0] = atomicAdd(&a.counters[C_NEXT], 1);
        if (tid < kBwdStages) full[tid] = 0;
        if (tid == 0) cons[0] = 0;
        __syncthreads();
        const int c = s_ctl[0];
        if (c >= a.nclaim) break;
        const int i = a.nclaim - 1 - c;
        const int kb = a.kL[i], ke = a.kL[i + 1];
        const int marked = a.mark[i];
        const int nchunks = marked ? (ke - kb + kBwdChunk - 1) / kBwdChunk : 0;

        if (warp == 0) {
            double beta = zmine[i];                     // z[i] after the diagonal sweep (previous launch)
            if (marked) {
                for (int g = 0; g < nchunks; ++g) {
                    const int st = g % kBwdStages;
                    while (vbk_lds_acquire(&full[st]) != g + 1) vbk_pause();
                    const double* p = myring + st * kBwdChunk;
                    // fully unrolled: the compiler keeps a few loads ahead of the dependent subtractions
#pragma unroll
                    for (int q = 0; q < kBwdChunk; ++q) beta = beta - p[q];         // z[i] -= AAt[k]*z[row], ldlt.c:494
                    __syncwarp();
                    if (lane == 0) vbk_sts_release(&cons[0], g + 1);
                }
                if (lane < NRHS) zmine[i] = beta;
            } else if (lane < NRHS) {
                if (fabs(beta) > eps) a.counters[C_CONSISTENT + myrhs] = 0;
                else zmine[i] = 0.0;
            }
            if (NRHS == 2) __syncwarp();
            if (lane == 0) { vbk_fence_release(); atomicExch(&a.done[i], 1); }
        } else {
            const int pw = warp - 1;
            for (int g = pw; g < nchunks; g += P) {
                const int st = g % kBwdStages;
                const int k0 = kb + g * kBwdChunk + lane;
                double l[4];
                int r[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + 32 * u;
                    l[u] = 0.0; r[u] = -1;
                    if (k < ke) { l[u] = a.L[k]; r[u] = a.iL[k]; }
                }
                bool pending = false;
#pragma unroll
                for (int u = 0; u < 4; ++u) if (r[u] >= 0 && r[u] < a.nclaim && vbk_ld_volatile(&a.done[r[u]]) == 0) pending = true;
                if (pending) {
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (r[u] >= 0 && r[u] < a.nclaim) while (vbk_ld_volatile(&a.done[r[u]]) == 0) { __nanosleep(20); vbk_pause(); }
                }
                // The z loads below are issued only after the polls have returned (control dependency) and bypass L1, the
                // writer released z before raising the flag: the classic volatile-flag hand-off, no reader-side fence.
                double zr[4], zr2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    zr[u] = (r[u] >= 0) ? __ldcg(&a.z[r[u]]) : 0.0;
                    zr2[u] = (NRHS == 2 && r[u] >= 0) ? __ldcg(&a.z2[r[u]]) : 0.0;
                }
                if (g >= kBwdStages) { while (vbk_lds_acquire(&cons[0]) < g - kBwdStages + 1) vbk_pause(); }
                double* p = ring + st * kBwdChunk + lane;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    p[32 * u] = l[u] * zr[u];           // +0.0 beyond the column's end: x - (+0.0) == x
                    if (NRHS == 2) p[kBwdStages * kBwdChunk + 32 * u] = l[u] * zr2[u];
                }
                __syncwarp();
                if (lane == 0) vbk_sts_release(&full[st], g + 1);
            }
        }
    }
}

}  // namespace vbk
