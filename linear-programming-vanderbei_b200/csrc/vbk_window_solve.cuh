// vbk_window_solve.cuh -- dense-window triangular sweeps (fast mode): 128-row panels with inverted
// diagonal blocks.
//
// A sweep that hands over from one 32-row panel to the next through a global flag pays 134 hand-offs on dfl001's window
// (4.8 us each, 640 us per sweep, four sweeps per KKT step).  Here
//   * a panel is 128 rows: 34 hand-offs instead of 134;
//   * the diagonal blocks are inverted ONCE per factorisation (k_window_tinv), so the triangular solve inside a panel
//     becomes a 128 x 128 matrix-vector product from shared memory -- no dependent chain at all;
//   * the row block left of (forward) / right of (backward) the diagonal block is cut into kTriSplit interleaved
//     column slices owned by different CTAs, so that no CTA has to stream more than a quarter of a panel's row; the
//     slices meet in a global accumulator, and only the slice that holds the LAST block waits on the critical path.
// Storage: S keeps L in the lower triangle and L^T mirrored in the upper one (k_window_mirror), so both sweeps read
// rows of the same column-major array; Tinv[p] mirrors (I + L_pp)^-1 the same way (lower: forward, upper: its
// transpose = the inverse of the backward sweep's unit upper block).
// Dependent rows (mark = 0; reference ldlt.c:470-476, 484-489): column r of L is zero, so z_r never feeds a forward
// update and the forward test can run after the product; in the backward sweep row r of L^T is zero, so the test
// runs on the right-hand side BEFORE the product (the zeroed value is what the rows above must see).
#pragma once
#include "vbk_dense_update.cuh"

namespace vbk {

// (also compiled for the host thread emulator: tests/test_emu.py runs the fast-mode solves through these kernels)
constexpr int kTriPW = 128;                    // rows per panel
constexpr int kTriSplit = 4;                   // column slices per panel row
constexpr int kTriV3Threads = 256;
constexpr int kTinvLd = 129;
constexpr size_t kTinvSmem = sizeof(double) * (kTriPW * kTinvLd + kTriPW * (kTriPW - 1) / 2);
constexpr size_t kTriV3Smem = sizeof(double) * (kTriPW * kTriPW + 3 * kTriPW + 2 * kTriPW) + 16;

struct Tri3Args {
    int W, ld, npan, dir;          // dir 0: forward (lower), 1: backward (upper)
    const double* S; const double* Tinv; double* z; const int* mark;     // z, mark already offset to the window
    double* racc;                  // [npan][128] partial sums of the non-owning slices (zeroed before the sweep)
    int* flags;                    // [npan] panel done | [npan] slices arrived | [1] claim counter   (zeroed)
    int* counters; const unsigned long long* scal_bits; double epssol;
};

// Tinv[p] := mirrored inverse of the unit lower triangular diagonal block p of the window.  One CTA per block,
// thread j owns column j of the inverse (right-looking substitution, X in shared memory, L packed in shared memory
// and read as a broadcast).  ~20 us, all blocks in parallel, once per factorisation.
static __global__ void __launch_bounds__(kTriPW) k_window_tinv(int W, int ld, const double* __restrict__ S, double* __restrict__ Tinv)
{
    VBK_DYN_SMEM(raw);
    double* X = reinterpret_cast<double*>(raw);               // [128][129]
    double* Lp = X + kTriPW * kTinvLd;                        // packed strictly-lower: (i, k) at i (i - 1) / 2 + k
    const int p = blockIdx.x, P0 = p * kTriPW, j = threadIdx.x;
    const int nb = (W - P0 < kTriPW) ? (W - P0) : kTriPW;
    for (int k = 0; k < kTriPW; ++k)                          // column k of the block, rows below it: coalesced over i = j
        if (j > k) Lp[j * (j - 1) / 2 + k] = (j < nb) ? S[(size_t)(P0 + j) + (size_t)(P0 + k) * ld] : 0.0;
    for (int i = 0; i < kTriPW; ++i) X[i * kTinvLd + j] = 0.0;
    __syncthreads();
    // Column j of the inverse, 32 rows at a time in registers (rows above the column's own chunk are zero): first the
    // finished rows k of the earlier chunks (x_k from shared memory, L[i][k] as a broadcast), then the chunk's own
    // triangle.  One load and one fma per update, nothing dependent through shared memory (a version that kept x in
    // shared memory took 308 us per factorisation, profiles/).
    const int w = j >> 5;
#pragma unroll
    for (int R = 0; R < kTriPW / 32; ++R) {
        if (R < w) continue;                                                  // uniform per warp
        double x[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) x[u] = (32 * R + u == j) ? 1.0 : 0.0;
        for (int k = 32 * w; k < 32 * R; ++k) {
            const double xk = X[k * kTinvLd + j];
            const double* lc = Lp + k;
#pragma unroll
            for (int u = 0; u < 32; ++u) x[u] = fma(-lc[(32 * R + u) * (32 * R + u - 1) / 2], xk, x[u]);
        }
#pragma unroll
        for (int kk = 0; kk < 31; ++kk) {
            const double xk = x[kk];
            const double* lc = Lp + 32 * R + kk;
#pragma unroll
            for (int u = kk + 1; u < 32; ++u) x[u] = fma(-lc[(32 * R + u) * (32 * R + u - 1) / 2], xk, x[u]);
        }
#pragma unroll
        for (int u = 0; u < 32; ++u) X[(32 * R + u) * kTinvLd + j] = x[u];
    }
    __syncthreads();
    double* M = Tinv + (size_t)p * kTriPW * kTriPW;
    for (int c = 0; c < kTriPW; ++c) {
        // element (row j, col c): lower part = X[j][c]; upper part (j < c) = X[c][j]; both reads conflict-free / broadcast-free
        const double v = (j == c) ? 1.0 : (j > c ? X[j * kTinvLd + c] : X[c * kTinvLd + j]);
        M[j + (size_t)c * kTriPW] = v;
    }
}

static __global__ void __launch_bounds__(kTriV3Threads, 1) k_window_tri3(Tri3Args a)
{
    VBK_DYN_SMEM(raw);
    double* Msh = reinterpret_cast<double*>(raw);             // [128][128] Tinv block of the owned panel (column-major)
    double* zq = Msh + kTriPW * kTriPW;                       // [128] published z of the source panel
    double* rsh = zq + kTriPW;                                // [128] right-hand side of the diagonal product
    double* part = rsh + kTriPW;                              // [2][128] halves of a product   (+ [128] spare)
    volatile int* s_claim = reinterpret_cast<volatile int*>(part + 3 * kTriPW);
    const int tid = threadIdx.x, row = tid & (kTriPW - 1), half = tid >> 7;
    int* done = a.flags;
    int* arrived = a.flags + a.npan;
    int* claim = a.flags + 2 * a.npan;
    const double eps = a.counters[C_NDEP] ? a.epssol * bits_to_double(a.scal_bits[S_ZMAX]) : 0.0;

    for (;;) {
        // items (panel, slice) are claimed in dependency order: whatever an item waits for is already owned by a running
        // CTA (or done), whatever else shares the GPU -- the batch driver runs many sweeps at once
        __syncthreads();
        if (tid == 0) *s_claim = atomicAdd(claim, 1);
        __syncthreads();
        const int item = *s_claim;
        if (item >= a.npan * kTriSplit) break;
        // pp: position in the sweep.  The slice that holds the last block (qq = pp - 1) owns the panel and is claimed
        // LAST of the four: everything it waits for -- the other slices included -- was claimed before it.
        const int pp = item / kTriSplit;
        const int own = (pp == 0) ? 0 : (pp - 1) % kTriSplit;
        const int sl = (own + 1 + item % kTriSplit) % kTriSplit;
        const int p = a.dir ? a.npan - 1 - pp : pp;
        const bool owner = sl == own;
        const int grow = p * kTriPW + row;
        const bool rok = grow < a.W;
        if (owner) {
            // this panel's inverted diagonal block, asynchronously: needed only at the very end
            const double* M = a.Tinv + (size_t)p * kTriPW * kTriPW;
#ifdef VBK_EMU
            for (int e = tid; e < kTriPW * kTriPW; e += kTriV3Threads) Msh[e] = M[e];
#else
            for (int e = tid * 2; e < kTriPW * kTriPW; e += kTriV3Threads * 2) cp_async16(Msh + e, M + e, true);
            cp_async_commit();
#endif
        }
        double acc = 0.0;
        // the owner's own right-hand side entries and the other slices' sum are fetched off the critical path: z[grow]
        // now, the accumulator together with the last block's z (the other slices finished a hand-off ago)
        double rz = 0.0, rothers = 0.0;
        if (owner && tid < kTriPW) rz = rok ? a.z[grow] : 0.0;
        for (int qq = sl; qq < pp; qq += kTriSplit) {
            const int q = a.dir ? a.npan - 1 - qq : qq;
            const bool lastblk = owner && qq + kTriSplit >= pp;
            // this thread's 64 entries of the block S[R_p, C_q]: row `row`, columns half*64 .. +63 (coalesced over rows)
            double b[64];
            const double* src = a.S + (size_t)grow + (size_t)(q * kTriPW + half * 64) * a.ld;
#pragma unroll
            for (int c = 0; c < 64; ++c) b[c] = (rok && q * kTriPW + half * 64 + c < a.W) ? __ldg(src + (size_t)c * a.ld) : 0.0;
            if (tid == 0) {
                if (lastblk) while (vbk_ld_volatile(&arrived[p]) < kTriSplit - 1) __nanosleep(20);
                while (vbk_ld_volatile(&done[q]) == 0) __nanosleep(20);
                __threadfence();
            }
            __syncthreads();
            if (tid < kTriPW) {
                zq[tid] = (q * kTriPW + tid < a.W) ? __ldcg(&a.z[q * kTriPW + tid]) : 0.0;
                if (lastblk) {                   // the other slices' partial sums, added in slice order: the same bits every run
#pragma unroll
                    for (int s2 = 0; s2 < kTriSplit; ++s2)
                        if (s2 != own) rothers += __ldcg(&a.racc[((size_t)p * kTriSplit + s2) * kTriPW + tid]);
                }
            }
            __syncthreads();
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            const double* zz = zq + half * 64;
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
                s0 = fma(b[c], zz[c], s0); s1 = fma(b[c + 1], zz[c + 1], s1);
                s2 = fma(b[c + 2], zz[c + 2], s2); s3 = fma(b[c + 3], zz[c + 3], s3);
            }
            acc += (s0 + s1) + (s2 + s3);
        }
        part[half * kTriPW + row] = acc;
        __syncthreads();
        if (!owner) {
            if (pp > 0) {
                if (tid < kTriPW && sl < pp) a.racc[((size_t)p * kTriSplit + sl) * kTriPW + tid] = part[tid] + part[kTriPW + tid];
                __threadfence();
                __syncthreads();
                if (tid == 0) atomicAdd(&arrived[p], 1);
            }
            continue;
        }
        // owner: build the right-hand side (pp > 0: the last block's iteration above has waited for the other slices)
        if (tid < kTriPW) {
            double r = rz;
            r -= part[tid] + part[kTriPW + tid];
            r -= rothers;
            if (a.dir && rok && !a.mark[grow]) {                          // backward: the test comes before the product
                if (fabs(r) > eps) a.counters[C_CONSISTENT] = 0; else r = 0.0;
            }
            rsh[tid] = r;
        }
#ifndef VBK_EMU
        cp_async_wait<0>();
#endif
        __syncthreads();
        // v = Tinv_pp r : forward uses the strictly lower part (+ unit diagonal), backward the strictly upper part
        {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            const double* mrow = Msh + row + (size_t)(half * 64) * kTriPW;
            const double* rr = rsh + half * 64;
#pragma unroll
            for (int c = 0; c < 64; c += 4) {
                const int col = half * 64 + c;
                const double m0 = mrow[(size_t)(c) * kTriPW], m1 = mrow[(size_t)(c + 1) * kTriPW];
                const double m2 = mrow[(size_t)(c + 2) * kTriPW], m3 = mrow[(size_t)(c + 3) * kTriPW];
                const bool u0 = a.dir ? (col > row) : (col < row), u1 = a.dir ? (col + 1 > row) : (col + 1 < row);
                const bool u2 = a.dir ? (col + 2 > row) : (col + 2 < row), u3 = a.dir ? (col + 3 > row) : (col + 3 < row);
                s0 = fma(u0 ? m0 : 0.0, rr[c], s0); s1 = fma(u1 ? m1 : 0.0, rr[c + 1], s1);
                s2 = fma(u2 ? m2 : 0.0, rr[c + 2], s2); s3 = fma(u3 ? m3 : 0.0, rr[c + 3], s3);
            }
            part[half * kTriPW + row] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
        if (tid < kTriPW && rok) {
            double v = rsh[tid] + part[tid] + part[kTriPW + tid];
            if (!a.dir && !a.mark[grow]) {                                // forward: the test follows the substitution
                if (fabs(v) > eps) a.counters[C_CONSISTENT] = 0; else v = 0.0;
            }
            a.z[grow] = v;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicExch(&done[p], 1);
    }
}

}  // namespace vbk
