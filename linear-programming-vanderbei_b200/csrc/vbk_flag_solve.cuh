// vbk_flag_solve.cuh -- forward / backward substitution with per-column completion flags (rawsolve, reference
// src/ipo/ldlt.c:433-505).  done[j] != 0  <=>  z[j] is final.  One warp per row (forward, ascending claims) or per
// column (backward, descending claims); dependencies always point to indices claimed earlier, so the kernels cannot
// deadlock for any grid.  Forward substitution waits per COLUMN, so a row consumes z[j] as soon as it exists and the
// dense tail pipelines; the dependent subtract chains run through shared memory.  Strict mode uses k_fwd_flags for the
// forward sweep (its backward sweep is k_bwd_pipe, vbk_strict_solve.cuh); fast mode uses both for the sparse columns
// below the dense window (fast = 1: sums may be re-associated).
#pragma once
#include "vbk_kernels.cuh"

namespace vbk {

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
    double* z2 = nullptr;   // second right-hand side (k_fwd_flags<2>): same structure, same waits, its own chain
    int rhs = 0;            // scalar slots of z when there is one right-hand side
};

static __global__ void k_flags_reset(int N, int* __restrict__ done, int* __restrict__ counters, int set_consistent)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) done[t] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        counters[C_NEXT] = 0;
        if (set_consistent) { counters[C_CONSISTENT] = 1; counters[C_CONSISTENT2] = 1; }
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

__device__ __forceinline__ double flag_solve_eps(const FlagSolveArgs& a, int rhs = 0) {
    return a.counters[C_NDEP] ? a.epssol * bits_to_double(a.scal_bits[S_ZMAX + kRhsSlotStride * rhs]) : 0.0;   // ldlt.c:446
}

// NRHS = 2: two right-hand sides in one sweep (solve2: the two systems of an hsd iteration share the factor).  The
// sweep is bound by its dependency chain, not by arithmetic: lane 0 runs the chain of z, lane 1 the chain of z2, side by
// side in the same instructions, and every wait, row list and L value is shared.
template <int NRHS>
static __global__ void __launch_bounds__(kSolveThreads) k_fwd_flags(FlagSolveArgs a)
{
    VBK_DYN_SMEM(raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sp = reinterpret_cast<double*>(raw) + warp * 32 * NRHS;             // products: [NRHS][32]
    const int myrhs = (NRHS == 2) ? (lane & 1) : a.rhs;                          // slots of the chain this lane may run
    const double eps = flag_solve_eps(a, myrhs);
    double* const zmine = (NRHS == 2 && lane == 1) ? a.z2 : a.z;
    for (;;) {
        int r = 0;
        if (lane == 0) r = atomicAdd(&a.counters[C_NEXT], 1);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= a.nclaim) break;
        double acc = zmine[r];                       // right-hand side entry, written before the launch
        const int rb = a.rowptr[r], re = a.rowptr[r + 1];
        // the row's entries 32 at a time; the static part of the NEXT batch (column, L value) is fetched while this one
        // waits for its columns and runs its chain
        int nj = -1;
        double nl = 0.0;
        if (rb + lane < re) { nj = a.rj[rb + lane]; nl = a.L[a.rk[rb + lane]]; }
        for (int t0 = rb; t0 < re; t0 += 32) {
            const int j = nj;
            const double l = nl;
            nj = -1; nl = 0.0;
            if (t0 + 32 + lane < re) { nj = a.rj[t0 + 32 + lane]; nl = a.L[a.rk[t0 + 32 + lane]]; }
            // Wait for the batch's columns.  Every lane looks at its flag once; then only the lane with the smallest
            // unfinished column spins (a row of the dense tail waits for the next 32 columns of the chain: 32 spinning
            // lanes in each of thousands of warps would saturate the L2 with polls), the others follow one by one.
            unsigned pending = __ballot_sync(0xffffffffu, j >= 0 && vbk_ld_volatile(&a.done[j]) == 0);
            while (pending) {
                const int first = __ffs(pending) - 1;
                if (lane == first) { while (vbk_ld_volatile(&a.done[j]) == 0) __nanosleep(20); }
                __syncwarp();
                pending &= pending - 1;          // the others are not looked at again here: each gets its own turn (its first
                                                 // poll is the check), which keeps a round trip off the chain of the dense tail
            }
            // z[j] is read only after its flag has been seen set (control dependency) and bypasses L1; the writer
            // released z before raising the flag
            const bool live = j >= 0 && a.mark[j];   // an unmarked column contributes nothing (ldlt.c:455); x - (+0.0) == x
            sp[lane] = live ? l * __ldcg(&a.z[j]) : 0.0;
            if (NRHS == 2) sp[32 + lane] = live ? l * __ldcg(&a.z2[j]) : 0.0;
            __syncwarp();
            if (lane < NRHS) acc = chain_sub(acc, sp + 32 * lane, (re - t0 < 32) ? (re - t0) : 32);   // z[row] -= AAt[k]*beta
            __syncwarp();
        }
        if (lane < NRHS) {
            if (a.mark[r]) zmine[r] = acc;
            else if (fabs(acc) > eps) { zmine[r] = acc; a.counters[C_CONSISTENT + myrhs] = 0; }
            else zmine[r] = 0.0;
        }
        if (NRHS == 2) __syncwarp();
        if (lane == 0) {
            vbk_fence_release();
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
            // z of the finished rows is read below with __ldcg, after this poll has returned (control dependency): the
            // volatile-flag hand-off of k_bwd_pipe, no reader-side fence
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
            vbk_fence_release();
            atomicExch(&a.done[i], 1);
        }
        __syncwarp();
    }
}

}  // namespace vbk
