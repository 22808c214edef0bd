// vbk_dense_common.cuh -- fast mode, shared definitions of the dense-window kernels: the argument block, the scalar
// (DFMA) rank-k update that stands in for the tensor-path update in the host-emulated test build, and the kernels that
// move the window between its dense scratch and the packed storage of L (mirror, store, gather of the sparse part).
#pragma once
#include "vbk_flag_solve.cuh"

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
    double* PB;                // packed panel buffer (vbk_dense_panel.cuh): L11^T, reciprocal pivots, keep flags
    double* PB2 = nullptr;     // second packed buffer for k_panel_rows_m: off-diagonal blocks, inverted diagonal blocks
    // the rank-k update takes its k columns S[:, kcol0..kcol0+klen)
    // and P[:, pcol0..pcol0+klen) and touches target rows/columns [rbase, W) x [rbase, cmax)
    int kcol0, klen, pcol0, rbase, cmax;
    // split look-ahead (vbk_kkt_fast.cu): rows skipped below rbase by the tensor-path update, slab range of k_panel_rows_m
    int rskip = 0, slab_lo = 0, slab_hi = 0x7fffffff;
};

__device__ __forceinline__ double& SW(const DenseArgs& a, int r, int c) { return a.S[(size_t)r + (size_t)c * a.ld]; }

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

constexpr int kOuterPanel = 4 * kPanelMax;     // 128 columns per trailing update


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

}  // namespace vbk
