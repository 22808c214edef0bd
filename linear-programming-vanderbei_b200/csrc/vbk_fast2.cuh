// vbk_fast2.cuh -- second-generation kernels of the dense-window factorisation (fast mode).
//
// Measured on B200 with the first-generation kernels of vbk_fast.cuh (dfl001, padded window W=4277,
// 134 panels of 32): k_dense_diag 28 us, k_dense_trsm 45 us and the rank-32 update 65 us per panel --
// launch- and latency-bound, 1.4 TFLOP/s overall.  Here:
//   * the 32x32 diagonal block is factorised by ONE warp with the block held in registers (one row
//     per lane, columns exchanged by shuffles): no barriers;
//   * the row-parallel TRSM is fully unrolled (the recurrence lives in registers, not local memory);
//   * two-level blocking: inside an outer panel of 128 columns only the panel's own strip is updated
//     after every 32 columns; the trailing matrix gets ONE rank-128 update per outer panel, so the
//     trailing matrix is read and written four times less often and the FP64 pipe sees 4x the work
//     per byte (the register-tiled update kernel loops over the 128 columns in chunks of 32).
#pragma once
#include "vbk_fast.cuh"

namespace vbk {

constexpr int kOuterPanel = 4 * kPanelMax;     // 128 columns per trailing update

// rare path of the pivot rule (ldlt.c:600-614), kept out of line so that the unrolled main path of
// k_dense_diag_w stays small enough for the instruction cache: max |updated column c below the pivot|
// over the rows below the block (the block's own rows are handled by the caller from registers)
#ifndef VBK_EMU
__noinline__
#endif
__device__ double dense_colmax_below(const DenseArgs& a, int c, int lane, const double* blk,
                                                  const double* sd, const int* skeep)
{
    const int LDB = kPanelMax + 1, p = a.p, nb = a.nb;
    double mymax = 0.0;
    for (int r = nb + lane; r < a.W - p; r += 32) {
        double l[kPanelMax];
        for (int c1 = 0; c1 < c; ++c1) {
            double s = SW(a, p + r, p + c1);
            for (int c0 = 0; c0 < c1; ++c0) s = fma(-l[c0] * sd[c0], blk[c1 * LDB + c0], s);
            l[c1] = skeep[c1] ? s / sd[c1] : 0.0;
        }
        double v = SW(a, p + r, p + c);
        for (int c0 = 0; c0 < c; ++c0) v = fma(-l[c0] * sd[c0], blk[c * LDB + c0], v);
        if (fabs(v) > mymax) mymax = fabs(v);
    }
    return mymax;
}

// one warp: LDL^T of the nb x nb diagonal block at (p,p); lane r owns row r
static __global__ void __launch_bounds__(32) k_dense_diag_w(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* blk = reinterpret_cast<double*>(raw);            // finished L11 columns, for the rare path
    double* sd = blk + kPanelMax * (kPanelMax + 1);
    int* skeep = reinterpret_cast<int*>(sd + kPanelMax);
    const int lane = threadIdx.x, nb = a.nb, p = a.p;
    const int LDB = kPanelMax + 1;
    double row[kPanelMax];
#pragma unroll
    for (int c = 0; c < kPanelMax; ++c) row[c] = (lane < nb && c <= lane && c < nb) ? SW(a, p + lane, p + c) : 0.0;
    double mymag = (lane < nb) ? a.wmag[p + lane] : 0.0;
    double myd = 0.0;
    int mykeep = 1;
#pragma unroll
    for (int c = 0; c < kPanelMax; ++c) {
        if (c < nb) {                                           // uniform branch
            double d = __shfl_sync(0xffffffffu, row[c], c);
            const double magc = __shfl_sync(0xffffffffu, mymag, c);
            int keep = 1;
            if (fabs(d) <= a.tol * magc) {
                // rare path (ldlt.c:600-614): max |updated column below the pivot|; rows of the block are
                // in registers, rows below get the panel's earlier columns applied on the fly
                double mymax = (lane > c && lane < nb) ? fabs(row[c]) : 0.0;
                const double below = dense_colmax_below(a, c, lane, blk, sd, skeep);
                if (below > mymax) mymax = below;
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) { double o = __shfl_xor_sync(0xffffffffu, mymax, s); if (o > mymax) mymax = o; }
                if (mymax < 1.0e+6 * 1.0e-8) keep = 0;
                else {
                    double sub = a.piv_scale * magc;
                    if (!(sub > 1.0e-8)) sub = 1.0e-8;
                    d = (a.perm[a.T + p + c] < a.n_ld ? -1 : 1) * sub;
                }
                if (lane == 0) atomicAdd(&a.counters[C_NDEP], 1);
            }
            if (lane == c) { myd = d; mykeep = keep; row[c] = d; sd[c] = d; skeep[c] = keep; }
            double l = 0.0;
            if (lane > c && lane < nb) { l = keep ? row[c] / d : 0.0; row[c] = l; }
            blk[lane * LDB + c] = l;                            // column c of L11 (0 on and above the diagonal)
            __syncwarp();
            if (keep) {
#pragma unroll
                for (int c2 = c + 1; c2 < kPanelMax; ++c2) {
                    const double lc2 = __shfl_sync(0xffffffffu, l, c2);
                    if (c2 < nb && lane >= c2) {
                        const double upd = l * d * lc2;
                        row[c2] -= upd;
                        if (lane == c2 && fabs(upd) > mymag) mymag = fabs(upd);
                    }
                }
            }
        }
    }
    if (lane < nb) {
#pragma unroll
        for (int c = 0; c < kPanelMax; ++c) if (c < lane) SW(a, p + lane, p + c) = row[c];
        a.dvec[p + lane] = myd;
        a.wmark[p + lane] = mykeep;
        a.wmag[p + lane] = mymag;
        a.pan_d[lane] = myd;
        a.pan_keep[lane] = mykeep;
    }
}

// rows below the block: L21 = S21 * L11^{-T} * D11^{-1}; P[:, pcol0..] = L21 * D11; trailing diagonal
static __global__ void __launch_bounds__(kDenseThreads) k_dense_trsm_u(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* l11 = reinterpret_cast<double*>(raw);            // [kPanelMax][kPanelMax+1]
    double* sd = l11 + kPanelMax * (kPanelMax + 1);
    int* skeep = reinterpret_cast<int*>(sd + kPanelMax);
    const int tid = threadIdx.x, nt = blockDim.x, nb = a.nb, p = a.p;
    const int LDB = kPanelMax + 1;
    for (int e = tid; e < kPanelMax * kPanelMax; e += nt) {
        int r = e % kPanelMax, c = e / kPanelMax;
        l11[r * LDB + c] = (r > c && r < nb) ? SW(a, p + r, p + c) : 0.0;
    }
    if (tid < kPanelMax) { sd[tid] = (tid < nb) ? a.pan_d[tid] : 1.0; skeep[tid] = (tid < nb) ? a.pan_keep[tid] : 0; }
    __syncthreads();
    for (int r = p + nb + blockIdx.x * nt + tid; r < a.W; r += gridDim.x * nt) {
        double l[kPanelMax];
#pragma unroll
        for (int c = 0; c < kPanelMax; ++c) l[c] = (c < nb) ? SW(a, r, p + c) : 0.0;
#pragma unroll
        for (int c = 0; c < kPanelMax; ++c) {
            double s = l[c];
#pragma unroll
            for (int c0 = 0; c0 < c; ++c0) s = fma(-l[c0] * sd[c0], l11[c * LDB + c0], s);
            l[c] = skeep[c] ? s / sd[c] : 0.0;
        }
        double dsum = 0.0, dabs = 0.0;
#pragma unroll
        for (int c = 0; c < kPanelMax; ++c) {
            if (c < nb) {
                const double w = l[c] * sd[c];
                SW(a, r, p + c) = l[c];
                a.P[(size_t)r + (size_t)(a.pcol0 + c) * a.W] = w;
                const double t = l[c] * w;
                dsum += t;
                if (fabs(t) > dabs) dabs = fabs(t);
            }
        }
        SW(a, r, r) -= dsum;
        if (dabs > a.wmag[r]) a.wmag[r] = dabs;
    }
}

// rank-klen update of the strictly-lower part of S[rbase.., rbase..cmax) in kUpdTD x kUpdTD tiles
static __global__ void __launch_bounds__(kUpdThreads) k_dense_update_k(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* As = reinterpret_cast<double*>(raw);             // [kPanelMax][kUpdTD]
    double* Bs = As + kPanelMax * kUpdTD;
    const int tr = blockIdx.y, tc = blockIdx.x;
    const int r0 = a.rbase + tr * kUpdTD, c0 = a.rbase + tc * kUpdTD;
    if (r0 + kUpdTD <= c0 || c0 >= a.cmax) return;            // tile entirely above the diagonal / outside
    const int tid = threadIdx.x, tx = tid % kUpdTG, ty = tid / kUpdTG;
    double acc[8][8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = 0.0;
    for (int kc = 0; kc < a.klen; kc += kPanelMax) {
        const int kn = (a.klen - kc < kPanelMax) ? (a.klen - kc) : kPanelMax;
        __syncthreads();
        for (int e = tid; e < kPanelMax * kUpdTD; e += kUpdThreads) {
            const int x = e % kUpdTD, c = e / kUpdTD;
            As[c * kUpdTD + x] = (c < kn && r0 + x < a.W) ? SW(a, r0 + x, a.kcol0 + kc + c) : 0.0;
            Bs[c * kUpdTD + x] = (c < kn && c0 + x < a.W) ? a.P[(size_t)(c0 + x) + (size_t)(a.pcol0 + kc + c) * a.W] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < kPanelMax; ++c) {
            double av[8], bv[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                av[2 * u] = As[c * kUpdTD + 2 * tx + 2 * kUpdTG * u];
                av[2 * u + 1] = As[c * kUpdTD + 2 * tx + 2 * kUpdTG * u + 1];
                bv[2 * u] = Bs[c * kUpdTD + 2 * ty + 2 * kUpdTG * u];
                bv[2 * u + 1] = Bs[c * kUpdTD + 2 * ty + 2 * kUpdTG * u + 1];
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < 8; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
        }
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const int c2 = c0 + 2 * ty + 2 * kUpdTG * (v >> 1) + (v & 1);
        if (c2 >= a.W || c2 >= a.cmax) continue;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r0 + 2 * tx + 2 * kUpdTG * (u >> 1) + (u & 1);
            if (r < a.W && r > c2) SW(a, r, c2) -= acc[u][v];
        }
    }
}

// --------------------------------------------------------------------------------------------
// Dense-window triangular sweeps on many CTAs.  After the factorisation the strictly-lower L of the
// window is mirrored into the upper triangle of the scratch matrix (k_window_mirror), so that the
// backward sweep (L^T) reads rows of the same column-major array as the forward sweep (L): both become
//     z[R_p] = T_pp^{-1} ( z[R_p] - sum_q  S[R_p, C_q] z[C_q] )      q < p forward, q > p backward
// over 32-row panels.  Panels are dealt round-robin to the CTAs; a panel's CTA multiplies the
// 32x32 blocks of its panel row as soon as the corresponding z[C_q] is published (one flag per
// panel), its warps sharing the blocks; the block of the panel that finishes last is already in
// registers when its flag flips, so the critical path per panel is one 32x32 mat-vec, a shared-memory
// reduction, the 32-step diagonal solve and one flag hand-off.
// --------------------------------------------------------------------------------------------
static __global__ void k_window_mirror(int W, int ld, double* __restrict__ S)
{
    VBK_DYN_SMEM(raw);
    double* t = reinterpret_cast<double*>(raw);      // [32][33]
    const int bx = blockIdx.x, by = blockIdx.y;      // tile (rows by, cols bx) of the lower triangle
    if (by < bx) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, nty = blockDim.x >> 5;
    for (int c = ty; c < 32; c += nty) {
        const int r = by * 32 + tx, cc = bx * 32 + c;
        t[c * 33 + tx] = (r < W && cc < W && r > cc) ? S[(size_t)r + (size_t)cc * ld] : 0.0;
    }
    __syncthreads();
    for (int c = ty; c < 32; c += nty) {
        // element (row = bx*32+tx, col = by*32+c) of the upper triangle = L[by*32+c, bx*32+tx]
        const int r = bx * 32 + tx, cc = by * 32 + c;
        if (r < W && cc < W && r < cc) S[(size_t)r + (size_t)cc * ld] = t[tx * 33 + c];
    }
}

#ifdef VBK_EMU
constexpr int kTriThreads = 64;
#else
constexpr int kTriThreads = 256;
#endif

struct TriArgs {
    int W, ld, npanels, dir;       // dir 0: forward (lower), 1: backward (upper)
    const double* S; double* z; const int* mark;     // z, mark already offset to the window
    int* flags; int* counters; const unsigned long long* scal_bits; double epssol;
};

static __global__ void __launch_bounds__(kTriThreads) k_window_tri(TriArgs a)
{
    VBK_DYN_SMEM(raw);
    double* part = reinterpret_cast<double*>(raw);            // [nwarps][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const double eps = a.counters[C_NDEP] ? a.epssol * bits_to_double(a.scal_bits[S_ZMAX]) : 0.0;
    volatile int* s_claimp = reinterpret_cast<volatile int*>(part + nwarps * 32);   // after part[nwarps][32]
    for (;;) {
        // panels are CLAIMED in dependency order (flags[npanels] is the claim counter): a panel some CTA
        // waits for is always owned by a CTA that is already running, whatever else shares the GPU
        // (the batch driver runs many of these sweeps concurrently on different streams)
        if (tid == 0) *s_claimp = atomicAdd(&a.flags[a.npanels], 1);
        __syncthreads();
        const int pp = *s_claimp;
        if (pp >= a.npanels) break;
        const int p = a.dir ? a.npanels - 1 - pp : pp;
        const int row = 32 * p + lane;
        const bool rok = row < a.W;
        double acc = 0.0;
        const int nblk = a.dir ? (a.npanels - 1 - p) : p;
        for (int bi = warp; bi < nblk; bi += nwarps) {
            const int q = a.dir ? (a.npanels - 1 - bi) : bi;
            double blk[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int col = 32 * q + c;
                blk[c] = (rok && col < a.W) ? a.S[(size_t)row + (size_t)col * a.ld] : 0.0;
            }
            if (lane == 0) {
                while (vbk_ld_volatile(&a.flags[q]) == 0) __nanosleep(20);
                __threadfence();
            }
            __syncwarp();
            const double zq = (32 * q + lane < a.W) ? __ldcg(&a.z[32 * q + lane]) : 0.0;
#pragma unroll
            for (int c = 0; c < 32; ++c) acc = fma(blk[c], __shfl_sync(0xffffffffu, zq, c), acc);
        }
        part[warp * 32 + lane] = acc;
        __syncthreads();
        if (warp == 0) {
            double v = rok ? a.z[row] : 0.0;
            for (int w = 0; w < nwarps; ++w) v -= part[w * 32 + lane];
            const int mk = rok ? a.mark[row] : 1;
            double tri[32];                                     // this lane's row of the diagonal block
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int col = 32 * p + c;
                const bool use = a.dir ? (c > lane) : (c < lane);
                tri[c] = (use && rok && col < a.W) ? a.S[(size_t)row + (size_t)col * a.ld] : 0.0;
            }
            if (!a.dir) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if (lane == c && !mk) { if (fabs(v) > eps) a.counters[C_CONSISTENT] = 0; else v = 0.0; }
                    const double zc = __shfl_sync(0xffffffffu, v, c);
                    if (tri[c] != 0.0) v = fma(-tri[c], zc, v);
                }
            } else {
#pragma unroll
                for (int c = 31; c >= 0; --c) {
                    if (lane == c && !mk) { if (fabs(v) > eps) a.counters[C_CONSISTENT] = 0; else v = 0.0; }
                    const double zc = __shfl_sync(0xffffffffu, v, c);
                    if (tri[c] != 0.0) v = fma(-tri[c], zc, v);
                }
            }
            if (rok) a.z[row] = v;
            __threadfence();
            __syncwarp();
            if (lane == 0) atomicExch(&a.flags[p], 1);
        }
        __syncthreads();
    }
}

}  // namespace vbk
