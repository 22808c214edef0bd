// vbk_fast.cuh -- FAST arithmetic mode: the trailing dense window of L as a dense problem.
//
// In the reference's ordering the last W columns of L are completely full (the "dense window",
// ldlt.c:1027; W = 239 / 695 / 2766 for 25fv47 / pilot87 / dfl001) and hold half or more of the
// factorisation's flops, but the reference's left-looking order makes them one dependent chain of W
// columns.  Fast mode gives up that order (so results agree with the reference to rounding, not bit
// for bit) and treats the window as what it is:
//   1. the sparse columns j < T are factorised by the task kernel of vbk_factor_tiled.cuh;
//   2. the same kernel in "phase 2" assembles the Schur complement of those columns into a dense
//      W x W scratch matrix (no dependencies: every task can run at once);
//   3. a blocked right-looking dense LDL^T (panel width <= 32: diagonal block, row-parallel TRSM,
//      tiled rank-nb update with explicit FP64 FMAs) factorises the scratch matrix;
//   4. triangular solves use dense panel sweeps on the window and the flag kernels below it.
// The exact-zero pivot rule of the reference (ldlt.c:600-614) is kept bit-exact in the sparse part;
// inside the window, where sums are re-associated, "exactly zero" becomes |d| <= tol * (sum of the
// magnitudes of the terms that formed d) (option ii).
#pragma once
#include "vbk_factor_tiled.cuh"

namespace vbk {

constexpr int kPanelMax = 32;
#ifdef VBK_EMU
constexpr int kDenseThreads = 32;
constexpr int kTileDim = 8;       // trailing-update tile edge in the emulated build
#else
constexpr int kDenseThreads = 256;
constexpr int kTileDim = 64;
#endif

struct DenseArgs {
    int W, ld, p, nb;          // window size, leading dimension, panel start, panel width
    double* S;                 // dense scratch, column-major; lower triangle + diagonal
    double* P;                 // panel scratch W x kPanelMax (ld = W): L21 * D of the current panel
    double* dvec; double* wmag; int* wmark;
    double* pan_d; int* pan_keep;      // [kPanelMax] pivots / marks of the current panel
    const int* perm; int T, n_ld;
    int* counters;
    double tol;
    // scale of the substitute for a dependent pivot: reference rule sgn*1e-8 (ldlt.c:612) when 0, otherwise
    // sgn*max(1e-8, piv_scale * largest term magnitude) ("static pivoting")
    double piv_scale;
    unsigned long long* prof;  // $VBK_PROF: cycle counters of the panel kernels (16 slots), else nullptr
    double* PB;                // packed panel buffer (vbk_fast3.cuh): L11^T, reciprocal pivots, keep flags
    double* PB2 = nullptr;     // second packed buffer for k_panel_rows_m: off-diagonal blocks, inverted diagonal blocks
    // two-level blocking (vbk_fast2.cuh): the rank-k update takes its k columns S[:, kcol0..kcol0+klen)
    // and P[:, pcol0..pcol0+klen) and touches target rows/columns [rbase, W) x [rbase, cmax)
    int kcol0, klen, pcol0, rbase, cmax;
    // split look-ahead (vbk_kkt_fast.cu): rows skipped below rbase by the tensor-path update, slab range of k_panel_rows_m
    int rskip = 0, slab_lo = 0, slab_hi = 0x7fffffff;
};

__device__ __forceinline__ double& SW(const DenseArgs& a, int r, int c) { return a.S[(size_t)r + (size_t)c * a.ld]; }

// A1: factor the nb x nb diagonal block (one CTA).  Unblocked LDL^T in shared memory.
static __global__ void __launch_bounds__(kDenseThreads) k_dense_diag(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* blk = reinterpret_cast<double*>(raw);            // [nb][kPanelMax+1]
    double* sd = blk + kPanelMax * (kPanelMax + 1);          // [kPanelMax] pivots
    double* red = sd + kPanelMax;                            // [kDenseThreads]
    int* skeep = reinterpret_cast<int*>(red + kDenseThreads); // [kPanelMax]
    const int tid = threadIdx.x, nt = blockDim.x, nb = a.nb, p = a.p;
    const int LDB = kPanelMax + 1;
    for (int e = tid; e < nb * nb; e += nt) {
        int r = e % nb, c = e / nb;
        blk[r * LDB + c] = (r >= c) ? SW(a, p + r, p + c) : 0.0;
    }
    __syncthreads();
    for (int c = 0; c < nb; ++c) {
        double d = blk[c * LDB + c];
        int keep = 1;
        const bool dep = fabs(d) <= a.tol * a.wmag[p + c];
        if (dep) {
            // rare path (ldlt.c:600-614): max |updated column below the pivot|.  Rows inside the block
            // are up to date; rows below the block get the panel's earlier columns applied on the fly.
            double mymax = 0.0;
            for (int r = c + 1 + tid; r < a.W - p; r += nt) {
                double v;
                if (r < nb) v = blk[r * LDB + c];
                else {
                    double l[kPanelMax];
                    for (int c1 = 0; c1 < c; ++c1) {
                        double s = SW(a, p + r, p + c1);
                        for (int c0 = 0; c0 < c1; ++c0) s = fma(-l[c0] * sd[c0], blk[c1 * LDB + c0], s);
                        l[c1] = skeep[c1] ? s / sd[c1] : 0.0;
                    }
                    v = SW(a, p + r, p + c);
                    for (int c0 = 0; c0 < c; ++c0) v = fma(-l[c0] * sd[c0], blk[c * LDB + c0], v);
                }
                if (fabs(v) > mymax) mymax = fabs(v);
            }
            red[tid] = mymax;
            __syncthreads();
            if (tid == 0) {
                double m = 0.0;
                for (int u = 0; u < nt; ++u) if (red[u] > m) m = red[u];
                red[0] = m;
            }
            __syncthreads();
            if (red[0] < 1.0e+6 * 1.0e-8) keep = 0;
            else d = (a.perm[a.T + p + c] < a.n_ld ? -1 : 1) * 1.0e-8;
            if (tid == 0) atomicAdd(&a.counters[C_NDEP], 1);
            __syncthreads();
        }
        if (tid == 0) { sd[c] = d; skeep[c] = keep; blk[c * LDB + c] = d; }
        // column c of L inside the block, then the trailing part of the block
        for (int r = c + 1 + tid; r < nb; r += nt) blk[r * LDB + c] = keep ? blk[r * LDB + c] / d : 0.0;
        __syncthreads();
        if (keep) {
            const int rem = nb - c - 1;
            for (int e = tid; e < rem * rem; e += nt) {
                int r = c + 1 + e % rem, c2 = c + 1 + e / rem;
                if (r >= c2) {
                    double upd = blk[r * LDB + c] * d * blk[c2 * LDB + c];
                    blk[r * LDB + c2] -= upd;
                    if (r == c2 && fabs(upd) > a.wmag[p + r]) a.wmag[p + r] = fabs(upd);
                }
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < nb * nb; e += nt) {
        int r = e % nb, c = e / nb;
        if (r > c) SW(a, p + r, p + c) = blk[r * LDB + c];
    }
    if (tid < nb) {
        a.dvec[p + tid] = sd[tid];
        a.wmark[p + tid] = skeep[tid];
        a.pan_d[tid] = sd[tid];
        a.pan_keep[tid] = skeep[tid];
    }
}

// A2: rows below the block: L21 = S21 * L11^{-T} * D11^{-1}, one thread per row; also P = L21*D11 and
// the diagonal of the trailing matrix (with its term magnitudes).
static __global__ void __launch_bounds__(kDenseThreads) k_dense_trsm(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* l11 = reinterpret_cast<double*>(raw);            // [nb][kPanelMax+1]
    double* sd = l11 + kPanelMax * (kPanelMax + 1);
    int* skeep = reinterpret_cast<int*>(sd + kPanelMax);
    const int tid = threadIdx.x, nt = blockDim.x, nb = a.nb, p = a.p;
    const int LDB = kPanelMax + 1;
    for (int e = tid; e < nb * nb; e += nt) {
        int r = e % nb, c = e / nb;
        l11[r * LDB + c] = (r > c) ? SW(a, p + r, p + c) : 0.0;
    }
    if (tid < nb) { sd[tid] = a.pan_d[tid]; skeep[tid] = a.pan_keep[tid]; }
    __syncthreads();
    for (int r = p + nb + blockIdx.x * nt + tid; r < a.W; r += gridDim.x * nt) {
        double l[kPanelMax];
        double dsum = 0.0, dabs = 0.0;
        for (int c = 0; c < nb; ++c) {
            double s = SW(a, r, p + c);
            for (int c0 = 0; c0 < c; ++c0) s = fma(-l[c0] * sd[c0], l11[c * LDB + c0], s);
            l[c] = skeep[c] ? s / sd[c] : 0.0;
        }
        for (int c = 0; c < nb; ++c) {
            const double w = l[c] * sd[c];
            SW(a, r, p + c) = l[c];
            a.P[(size_t)r + (size_t)c * a.W] = w;
            const double t = l[c] * w;
            dsum += t;
            if (fabs(t) > dabs) dabs = fabs(t);
        }
        SW(a, r, r) -= dsum;
        if (dabs > a.wmag[r]) a.wmag[r] = dabs;
    }
}

// A3: strictly-lower trailing update  S[r,c2] -= sum_c L21[r,c] * P[c2,c]  in square tiles.
static __global__ void __launch_bounds__(kDenseThreads) k_dense_update(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* As = reinterpret_cast<double*>(raw);             // [nb][kTileDim]   L21 rows of this tile
    double* Bs = As + kPanelMax * kTileDim;                  // [nb][kTileDim]   P rows of this tile's columns
    const int tr = blockIdx.y, tc = blockIdx.x;
    if (tr < tc) return;
    const int base = a.p + a.nb, nb = a.nb;
    const int r0 = base + tr * kTileDim, c0 = base + tc * kTileDim;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < nb * kTileDim; e += nt) {
        int x = e % kTileDim, c = e / kTileDim;
        As[c * kTileDim + x] = (r0 + x < a.W) ? SW(a, r0 + x, a.p + c) : 0.0;
        Bs[c * kTileDim + x] = (c0 + x < a.W) ? a.P[(size_t)(c0 + x) + (size_t)c * a.W] : 0.0;
    }
    __syncthreads();
    // each thread: a strip of the tile, 4 rows x (kTileDim*kTileDim/(4*nt)) columns
    // each work item: RPT rows (interleaved by `rgroups`, so a warp touches consecutive rows of one
    // column = coalesced) of one tile column
    constexpr int RPT = 4;
    const int rgroups = kTileDim / RPT;
    for (int item = tid; item < rgroups * kTileDim; item += nt) {
        const int rg = item % rgroups, cc = item / rgroups;
        double acc[RPT] = {0.0, 0.0, 0.0, 0.0};
        for (int c = 0; c < nb; ++c) {
            const double b = Bs[c * kTileDim + cc];
#pragma unroll
            for (int u = 0; u < RPT; ++u) acc[u] = fma(As[c * kTileDim + u * rgroups + rg], b, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < RPT; ++u) {
            const int r = r0 + u * rgroups + rg, c2 = c0 + cc;
            if (r < a.W && c2 < a.W && r > c2) SW(a, r, c2) -= acc[u];
        }
    }
}

// A3, register-tiled: TD x TD tiles, TG x TG threads, an 8 x 8 micro-tile per thread held in registers
// (rows 2*tx+{0,1}+2*TG*u, columns 2*ty+{0,1}+2*TG*v: consecutive threads touch consecutive rows of a
// column-major tile => coalesced C traffic and conflict-free 16-byte shared-memory reads).  64 FMAs per
// 8 shared-memory loads of 16 bytes: the FP64 pipe, not the LSU, is the limiter.
#ifdef VBK_EMU
constexpr int kUpdTG = 4;
#else
constexpr int kUpdTG = 16;
#endif
constexpr int kUpdTD = 8 * kUpdTG;
constexpr int kUpdThreads = kUpdTG * kUpdTG;

static __global__ void __launch_bounds__(kUpdThreads) k_dense_update_rt(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* As = reinterpret_cast<double*>(raw);             // [nb][kUpdTD]
    double* Bs = As + kPanelMax * kUpdTD;                    // [nb][kUpdTD]
    const int tr = blockIdx.y, tc = blockIdx.x;
    if (tr < tc) return;
    const int base = a.p + a.nb, nb = a.nb;
    const int r0 = base + tr * kUpdTD, c0 = base + tc * kUpdTD;
    const int tid = threadIdx.x, tx = tid % kUpdTG, ty = tid / kUpdTG;
    for (int e = tid; e < nb * kUpdTD; e += kUpdThreads) {
        const int x = e % kUpdTD, c = e / kUpdTD;
        As[c * kUpdTD + x] = (r0 + x < a.W) ? SW(a, r0 + x, a.p + c) : 0.0;
        Bs[c * kUpdTD + x] = (c0 + x < a.W) ? a.P[(size_t)(c0 + x) + (size_t)c * a.W] : 0.0;
    }
    __syncthreads();
    double acc[8][8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = 0.0;
    for (int c = 0; c < nb; ++c) {
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
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const int c2 = c0 + 2 * ty + 2 * kUpdTG * (v >> 1) + (v & 1);
        if (c2 >= a.W) continue;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r0 + 2 * tx + 2 * kUpdTG * (u >> 1) + (u & 1);
            if (r < a.W && r > c2) SW(a, r, c2) -= acc[u][v];
        }
    }
}

// copy the factored window back into the packed storage of L, diag and mark
static __global__ void k_window_store(int W, int T, int ld, const double* __restrict__ S, const double* __restrict__ dvec,
                                      const int* __restrict__ wmark, const int* __restrict__ kL, const int* __restrict__ iL,
                                      double* __restrict__ L, double* __restrict__ diag, int* __restrict__ mark)
{
    for (int c = blockIdx.y; c < W; c += gridDim.y) {
        const int kb = kL[T + c], ke = kL[T + c + 1];       // only the entries of the fill pattern exist in L
        for (int k = kb + blockIdx.x * blockDim.x + threadIdx.x; k < ke; k += gridDim.x * blockDim.x)
            L[k] = S[(size_t)(iL[k] - T) + (size_t)c * ld];
        if (blockIdx.x == 0 && threadIdx.x == 0) { diag[T + c] = dvec[c]; if (!wmark[c]) mark[T + c] = 0; }
    }
}

// Schur assembly, light version: one CTA per window column, contributors applied one after the
// other (used when the window rows see few sparse columns -- the usual case once the window is
// padded, because nearly all flops then live inside the window; the task kernel's phase 2 handles the
// heavy case).  S must be zero on entry.
struct SchurArgs {
    int N, T, ld;
    const int* kL; const int* iL; const double* L; const double* diag;
    const int* rowptr; const int* rk; const int* rj;       // ASCENDING row lists: sparse columns come first
    double* S; double* wmag;
};
static __global__ void __launch_bounds__(kDenseThreads) k_schur_window(SchurArgs a)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = a.T + blockIdx.x; i < a.N; i += gridDim.x) {
        double* col = a.S + (size_t)(i - a.T) * a.ld;
        for (int k = a.kL[i] + tid; k < a.kL[i + 1]; k += nt) col[a.iL[k] - a.T] = a.L[k];   // K[:, i]
        double d = a.diag[i], mag = fabs(d);
        __syncthreads();
        for (int t = a.rowptr[i]; t < a.rowptr[i + 1]; ++t) {
            const int j = a.rj[t];
            if (j >= a.T) break;
            const int k = a.rk[t];
            const double lij = a.L[k], w = lij * a.diag[j];
            const double p = lij * w;
            d -= p;
            if (fabs(p) > mag) mag = fabs(p);
            for (int kk = k + 1 + tid; kk < a.kL[j + 1]; kk += nt) {
                double* dst = &col[a.iL[kk] - a.T];
                *dst = fma(-w, a.L[kk], *dst);
            }
            __syncthreads();      // two contributors may touch the same row from different threads
        }
        if (tid == 0) { col[i - a.T] = d; a.wmag[i - a.T] = mag; }
    }
}

static __global__ void k_zero_counter(int* counters, int slot) { if (threadIdx.x == 0 && blockIdx.x == 0) counters[slot] = 0; }

// --------------------------------------------------------------------------------------------
// Fast-mode triangular solves on the window (rows/columns T..N-1).  The unit-lower factor of the
// window is read from the dense scratch S (column-major, leading dimension ld; zero outside the
// fill pattern), which stays valid until the next factorisation.
// --------------------------------------------------------------------------------------------
struct WindowSolveArgs {
    int N, T, ld;
    const double* S;
    const int* kL; const double* L; const int* mark;
    const int* rowptr; const int* rk; const int* rj;     // ascending row lists (for the coupling rows)
    const int* spend = nullptr;                          // [W] end of the sparse prefix of every window row's list, or null
    double* z;
    int* counters; const unsigned long long* scal_bits; double epssol;
};
__device__ __forceinline__ double win_eps(const WindowSolveArgs& a) {
    return a.counters[C_NDEP] ? a.epssol * bits_to_double(a.scal_bits[S_ZMAX]) : 0.0;
}
__device__ __forceinline__ double WL(const WindowSolveArgs& a, int r, int c) {   // L[T+r, T+c], r > c
    return a.S[(size_t)r + (size_t)c * a.ld];
}

// z[r] -= sum_{j<T} L[r,j] z[j] for window rows r (the sparse columns' contribution), warp per row
static __global__ void __launch_bounds__(kSolveThreads) k_window_gather(WindowSolveArgs a)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = a.T + warp; r < a.N; r += nwarps) {
        // the sparse columns are the leading part of the ascending row list: stop where the window columns begin
        // (a window row's list holds up to W of those), four entries per lane in flight
        const int tend = a.spend ? a.spend[r - a.T] : a.rowptr[r + 1];
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int t = a.rowptr[r] + lane; t < tend; t += 128) {
            const int t1 = t + 32, t2 = t + 64, t3 = t + 96;
            const int j0 = a.rj[t], j1 = t1 < tend ? a.rj[t1] : a.T, j2 = t2 < tend ? a.rj[t2] : a.T, j3 = t3 < tend ? a.rj[t3] : a.T;
            if (j0 < a.T && a.mark[j0]) s0 = fma(a.L[a.rk[t]], a.z[j0], s0);
            if (j1 < a.T && a.mark[j1]) s1 = fma(a.L[a.rk[t1]], a.z[j1], s1);
            if (j2 < a.T && a.mark[j2]) s2 = fma(a.L[a.rk[t2]], a.z[j2], s2);
            if (j3 < a.T && a.mark[j3]) s3 = fma(a.L[a.rk[t3]], a.z[j3], s3);
        }
        double s = (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        if (lane == 0) a.z[r] -= s;
    }
}

// forward sweep on the window: one CTA, 32-column panels
static __global__ void __launch_bounds__(kDenseThreads) k_window_fwd(WindowSolveArgs a)
{
    VBK_DYN_SMEM(raw);
    double* zp = reinterpret_cast<double*>(raw);      // [32] finished panel entries
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, W = a.N - a.T;
    const double eps = win_eps(a);
    double* z = a.z + a.T;
    for (int p = 0; p < W; p += 32) {
        const int nb = (W - p < 32) ? (W - p) : 32;
        if (tid < 32) {
            double v = (lane < nb) ? z[p + lane] : 0.0;
            double lrow[32];                                    // this lane's row of the diagonal block
#pragma unroll
            for (int c = 0; c < 32; ++c)
                lrow[c] = (c < lane && lane < nb && a.mark[a.T + p + c]) ? WL(a, p + lane, p + c) : 0.0;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                double zc = __shfl_sync(0xffffffffu, v, c);
                if (c < nb && lane == c && !a.mark[a.T + p + c]) {   // rawsolve's unmarked-row rule
                    if (fabs(v) > eps) a.counters[C_CONSISTENT] = 0; else v = 0.0;
                }
                if (lrow[c] != 0.0) v = fma(-lrow[c], zc, v);   // lrow[c] == 0 where nothing applies
            }
            if (lane < nb) { z[p + lane] = v; zp[lane] = a.mark[a.T + p + lane] ? v : 0.0; }
        }
        __syncthreads();
        for (int r = p + nb + tid; r < W; r += nt) {
            double acc = z[r];
            for (int c = 0; c < nb; ++c) acc = fma(-WL(a, r, p + c), zp[c], acc);
            z[r] = acc;
        }
        __syncthreads();
    }
}

// backward sweep on the window (L^T): one CTA; each warp reduces one panel column's tail product
static __global__ void __launch_bounds__(kDenseThreads) k_window_bwd(WindowSolveArgs a)
{
    VBK_DYN_SMEM(raw);
    double* part = reinterpret_cast<double*>(raw);    // [32]
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int W = a.N - a.T;
    const double eps = win_eps(a);
    double* z = a.z + a.T;
    const int npanels = (W + 31) / 32;
    for (int pi = npanels - 1; pi >= 0; --pi) {
        const int p = pi * 32;
        const int nb = (W - p < 32) ? (W - p) : 32;
        for (int c = warp; c < nb; c += nwarps) {
            double s = 0.0;
            if (a.mark[a.T + p + c])
                for (int r = p + nb + lane; r < W; r += 32) s = fma(WL(a, r, p + c), z[r], s);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            if (lane == 0) part[c] = s;
        }
        __syncthreads();
        if (tid < 32) {
            double v = (lane < nb) ? z[p + lane] - part[lane] : 0.0;
            double lcol[32];                                    // this lane's column of the diagonal block
#pragma unroll
            for (int c = 0; c < 32; ++c)
                lcol[c] = (c > lane && c < nb && a.mark[a.T + p + lane]) ? WL(a, p + c, p + lane) : 0.0;
#pragma unroll
            for (int c = 31; c >= 0; --c) {
                if (c < nb && lane == c && !a.mark[a.T + p + c]) {
                    if (fabs(v) > eps) a.counters[C_CONSISTENT] = 0; else v = 0.0;
                }
                double zc = __shfl_sync(0xffffffffu, v, c);
                if (lcol[c] != 0.0) v = fma(-lcol[c], zc, v);
            }
            if (lane < nb) z[p + lane] = v;
        }
        __syncthreads();
    }
}

}  // namespace vbk
