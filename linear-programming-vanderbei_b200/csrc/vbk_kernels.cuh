// vbk_kernels.cuh -- device kernels of the KKT step (sm_100a), STRICT arithmetic mode.
//
// Strict mode replays the reference's floating-point operations in the reference's order with
// the reference's roundings (no FMA contraction: this translation unit is compiled with
// -fmad=false), so that every sum is bit-identical to the x86-64 SSE2 build of the reference and
// the exact-zero pivot rule (reference src/ipo/ldlt.c:600-614) fires on exactly the same pivots
// (SURVEY.md section 7, H1/H2).  Parallelism is only ever ACROSS independent sums: rows of a
// column, columns of independent elimination-tree subtrees, rows of an SpMV -- never inside a sum.
//
// Scheduling is dataflow inside one persistent kernel per phase: CTAs (factor) or warps (solves)
// claim columns in index order from a global counter and wait on a per-column counter of
// unfinished elimination-tree children.  Claims ascend, dependencies always point to smaller
// indices, so the lowest unfinished column can always run: no deadlock for any grid size.
#pragma once
#include "vbk_device.h"

#include <cmath>
#include <cstdint>

namespace vbk {

#ifdef VBK_EMU
// the host emulation spawns one OS thread per CUDA thread: keep CTAs small there
constexpr int kSolveThreads = 32;
constexpr int kVecThreads = 32;
constexpr int kScanThreads = 32;
#else
constexpr int kSolveThreads = 128;    // 4 warps per CTA in the dataflow solve kernels
constexpr int kVecThreads = 256;
constexpr int kScanThreads = 1024;
#endif

// Device-resident scalars of one factor object (one allocation, index = enum below).
enum ScalarSlot {
    S_EPSDIAG = 0,   // current diagonal perturbation (ldlt.c:215,302), persists across calls
    S_MAXDIAG,       // bits of max|diag| before lltnum (ldlt.c:554-556)
    S_MINDIAG,       // bits of min|diag| after lltnum (ldlt.c:294-299)
    S_ZMAX,          // bits of maxv(z, m) for rawsolve's eps (ldlt.c:446)
    S_MAXBC_B, S_MAXBC_C,   // bits of maxv(b), maxv(c) (ldlt.c:367)
    S_MAXR, S_MAXS,  // bits of maxv(r), maxv(s) (ldlt.c:401)
    S_ZMAX2, S_MAXBC_B2, S_MAXBC_C2, S_MAXR2, S_MAXS2,   // the same five for the second right-hand side of solve2
    S_COUNT
};
constexpr int kRhsSlotStride = 5;     // S_ZMAX + kRhsSlotStride * rhs etc.: scalar slots of right-hand side `rhs` (0 or 1)
enum CounterSlot {
    C_NEXT = 0,      // column claim counter of the running dataflow kernel
    C_NDEP,          // dependent pivots of the last factorisation (ldlt.c:558,604)
    C_CONSISTENT,    // rawsolve's consistency flag (ldlt.c:439)
    C_CONSISTENT2,   // ... of the second right-hand side of solve2
    C_COUNT
};

__device__ __forceinline__ double vbk_abs(double v) { return v > 0 ? v : -v; }   // macros.h:3
__device__ __forceinline__ double bits_to_double(unsigned long long b) {
    double d;
#ifdef VBK_EMU
    std::memcpy(&d, &b, 8);
#else
    d = __longlong_as_double((long long)b);
#endif
    return d;
}
__device__ __forceinline__ unsigned long long double_to_bits(double d) {
#ifdef VBK_EMU
    unsigned long long b; std::memcpy(&b, &d, 8); return b;
#else
    return (unsigned long long)__double_as_longlong(d);
#endif
}
// max over non-negative doubles through their (monotone) bit patterns; NaNs are skipped.
// fabs, not the reference macro ABS: ABS(+0.0) is -0.0, whose bit pattern would win every max
__device__ __forceinline__ void atomic_absmax(unsigned long long* slot, double v) {
    double a = fabs(v);
    if (a == a) atomicMax(slot, double_to_bits(a));
}

// --------------------------------------------------------------------------------------------
// K1  assemble: diag <- (-max(dn,eps), +max(dm,eps)) permuted; mark <- TRUE; then the strictly
// lower triangle of the permuted K is scattered into L's value array through index maps that
// were computed once (inv_num, ldlt.c:235-269,280).
// --------------------------------------------------------------------------------------------
static __global__ void k_set_diag(int n, int m, const int* __restrict__ iperm, const double* __restrict__ dn,
                           const double* __restrict__ dm, const double* __restrict__ scal,
                           double* __restrict__ diag, int* __restrict__ mark,
                           unsigned long long* __restrict__ scal_bits)
{
    const double eps = scal[S_EPSDIAG];
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n + m; t += gridDim.x * blockDim.x) {
        double v;
        if (t < n) { double d = dn[t]; v = -(d > eps ? d : eps); }
        else       { double d = dm[t - n]; v = (d > eps ? d : eps); }
        diag[iperm[t]] = v;
        mark[t] = 1;
        atomic_absmax(&scal_bits[S_MAXDIAG], v);
    }
}

static __global__ void k_scatter(int nz, const int* __restrict__ map, const double* __restrict__ val,
                          double* __restrict__ L)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nz; k += gridDim.x * blockDim.x) {
        int p = map[k];
        if (p >= 0) L[p] = val[k];
    }
}

// mindiag < 1e-14  =>  epsdiag *= 10 (ldlt.c:293-306).  Two tiny kernels: reduction, then update.
static __global__ void k_min_absdiag(int N, const double* __restrict__ diag, unsigned long long* __restrict__ scal_bits)
{
    // min over |diag| via bit patterns of non-negative doubles (monotone); NaN never lowers the
    // minimum in the reference either (NaN < x is false)
    unsigned long long best = ~0ull;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < N; t += gridDim.x * blockDim.x) {
        double a = fabs(diag[t]);
        if (a == a) { unsigned long long b = double_to_bits(a); if (b < best) best = b; }
    }
    if (best != ~0ull) {
        // atomicMin on 64-bit: emulate through atomicMax of the complement
        atomicMax(&scal_bits[S_MINDIAG], ~best);
    }
}
static __global__ void k_update_epsdiag(double* scal, unsigned long long* scal_bits)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long c = scal_bits[S_MINDIAG];
        double mindiag = (c == 0ull) ? HUGE_VAL : bits_to_double(~c);
        if (mindiag < 1.0e-14) scal[S_EPSDIAG] *= 10;
    }
}

// --------------------------------------------------------------------------------------------
// K5  triangular solves, strict (rawsolve, ldlt.c:433-505).
// Forward: the reference scatters column by column, so z[r] receives its updates in ascending
// column order; the gather below walks row r's ascending list and subtracts one rounded product
// at a time -> same bits.  One warp per row: lanes fetch 32 products, then every lane replays the
// same 32 dependent subtractions (shuffle broadcast), so no divergence and no shared memory.
// --------------------------------------------------------------------------------------------
struct SolveArgs {
    int N, m_ld;
    const int* kL; const int* iL; const double* L; const double* diag; const int* mark;
    const int* rowptr; const int* rk; const int* rj;   // ascending lists
    const int* parent;
    double* z;
    int* counters;
    const unsigned long long* scal_bits;
    double epssol;
    int rhs = 0;       // which right-hand side's scalar slots (solve2)
};

__device__ __forceinline__ double solve_eps(const SolveArgs& a) {
    // ldlt.c:446: if (ndep) eps = epssol * maxv(z,m)
    return a.counters[C_NDEP] ? a.epssol * bits_to_double(a.scal_bits[S_ZMAX + kRhsSlotStride * a.rhs]) : 0.0;
}

static __global__ void k_diag_strict(SolveArgs a)
{
    const double eps = solve_eps(a);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.N; i += gridDim.x * blockDim.x) {
        double v = a.z[i];
        if (a.mark[i]) a.z[i] = v / a.diag[i];                  // ldlt.c:476
        else if (fabs(v) > eps) a.counters[C_CONSISTENT + a.rhs] = 0;
        else a.z[i] = 0.0;
    }
}

// --------------------------------------------------------------------------------------------
// K6/K7  refinement pieces of solve() (ldlt.c:367-401) and the SpMV gather.
// y = M x with M given row-wise (ptr, idx, val) where each row lists its entries in ascending
// source-column order -- the order in which the reference's CSC scatter smx (linalg.c:62-70) adds
// into y[row].  One thread per row, one rounded multiply and one rounded add per entry.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ double row_dot_strict(const int* __restrict__ ptr, const int* __restrict__ idx,
                                                 const double* __restrict__ val, const double* __restrict__ x, int r)
{
    double acc = 0.0;
    for (int k = ptr[r]; k < ptr[r + 1]; ++k) acc += val[k] * x[idx[k]];
    return acc;
}

static __global__ void k_spmv_rows(int nrows, const int* __restrict__ ptr, const int* __restrict__ idx,
                            const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x)
        y[r] = row_dot_strict(ptr, idx, val, x, r);
}

// s[j] = c[j] - (A'^T y - Dn*x)   and   r[i] = b[i] - (A' x + Dm*y), with the two max-norms
// (ldlt.c:389-401).  side 0: s-side (minus), side 1: r-side (plus).
static __global__ void k_residual(int nrows, int side, const int* __restrict__ ptr, const int* __restrict__ idx,
                           const double* __restrict__ val, const double* __restrict__ src,
                           const double* __restrict__ dvec, const double* __restrict__ own,
                           const double* __restrict__ rhs, double* __restrict__ out,
                           unsigned long long* __restrict__ maxslot)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
        double t = row_dot_strict(ptr, idx, val, src, r);
        double v;
        if (side == 0) v = rhs[r] - (t - dvec[r] * own[r]);      // ldlt.c:394 (Q empty)
        else           v = rhs[r] - (t + dvec[r] * own[r]);      // ldlt.c:397
        out[r] = v;
        atomic_absmax(maxslot, v);
    }
}

static __global__ void k_absmax(int n, const double* __restrict__ x, unsigned long long* __restrict__ slot)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        atomic_absmax(slot, x[t]);
}

static __global__ void k_zero_bits(unsigned long long* slots, int first, int count)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) slots[first + t] = 0ull;
}

// z[iperm[j]] = c[j], z[iperm[n+i]] = b[i]   (ldlt.c:371-377)
static __global__ void k_permute_in(int n, int m, const int* __restrict__ iperm, const double* __restrict__ c,
                             const double* __restrict__ b, double* __restrict__ z)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n + m; t += gridDim.x * blockDim.x)
        z[iperm[t]] = (t < n) ? c[t] : b[t - n];
}
// x_k (+)= z[iperm[j]], y_k (+)= z[iperm[n+i]]   (ldlt.c:381-387); sign -1 undoes (ldlt.c:413-416)
static __global__ void k_permute_out(int n, int m, int mode, const int* __restrict__ iperm,
                              const double* __restrict__ z, double* __restrict__ xk, double* __restrict__ yk)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n + m; t += gridDim.x * blockDim.x) {
        double v = z[iperm[t]];
        double* dst = (t < n) ? &xk[t] : &yk[t - n];
        if (mode == 0) *dst = v;
        else if (mode == 1) *dst = *dst + v;
        else *dst = *dst - v;
    }
}

// --------------------------------------------------------------------------------------------
// K9  reductions.  Strict dot product: the reference adds left to right (linalg.c:22), a single
// dependent chain.  Products are formed in parallel by the whole CTA into shared memory, thread 0
// walks them in order.  One CTA per dot product; several independent dot products per launch.
// --------------------------------------------------------------------------------------------
struct DotJob { const double* x; const double* y; int n; };
struct DotBatch { DotJob job[8]; int count; };
constexpr int kDotChunk = 2048;

static __global__ void __launch_bounds__(kVecThreads) k_dot_strict(DotBatch b, double* __restrict__ out)
{
    VBK_DYN_SMEM(raw);
    double* buf = reinterpret_cast<double*>(raw);          // [2][kDotChunk]
    const DotJob jb = b.job[blockIdx.x];
    const int tid = threadIdx.x, nt = blockDim.x;
    double acc = 0.0;
    int stage = 0;
    for (int q = tid; q < kDotChunk && q < jb.n; q += nt) buf[q] = jb.x[q] * jb.y[q];
    __syncthreads();
    for (int base = 0; base < jb.n; base += kDotChunk) {
        const int len = (jb.n - base < kDotChunk) ? (jb.n - base) : kDotChunk;
        double* cur = buf + stage * kDotChunk;
        double* nxt = buf + (stage ^ 1) * kDotChunk;
        if (tid == 0) {
            for (int q = 0; q < len; ++q) acc += cur[q];
        } else {
            const int nb = base + kDotChunk;
            for (int q = tid - 1; q < kDotChunk && nb + q < jb.n; q += nt - 1)
                nxt[q] = jb.x[nb + q] * jb.y[nb + q];
        }
        __syncthreads();
        stage ^= 1;
    }
    if (tid == 0) out[blockIdx.x] = acc;
}

// fast-mode dot product: fixed-shape tree (deterministic run to run, NOT the reference's order)
static __global__ void __launch_bounds__(kVecThreads) k_dot_tree(DotBatch b, double* __restrict__ out)
{
    VBK_DYN_SMEM(raw);
    double* red = reinterpret_cast<double*>(raw);
    const DotJob jb = b.job[blockIdx.x];
    double acc = 0.0;
    for (int q = threadIdx.x; q < jb.n; q += blockDim.x) acc += jb.x[q] * jb.y[q];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

// --------------------------------------------------------------------------------------------
// K8  CSC transpose on the device (atnum, linalg.c:75-103): histogram, exclusive scan (single
// CTA, chunked), unordered atomic scatter of source positions, then a per-row rank sort of those
// positions so that every output column lists its entries in ascending input column -- the
// reference's order.
// --------------------------------------------------------------------------------------------
static __global__ void k_hist(int nz, const int* __restrict__ ia, int* __restrict__ cnt)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nz; k += gridDim.x * blockDim.x)
        atomicAdd(&cnt[ia[k]], 1);
}
// exclusive scan of cnt[0..m) into ptr[0..m], single CTA
static __global__ void k_scan_single(int m, const int* __restrict__ cnt, int* __restrict__ ptr)
{
    VBK_DYN_SMEM(raw);
    int* s = reinterpret_cast<int*>(raw);      // [blockDim.x]
    const int tid = threadIdx.x, nt = blockDim.x;
    const int per = (m + nt - 1) / nt;
    const int lo = tid * per, hi = (lo + per < m) ? lo + per : m;
    int sum = 0;
    for (int q = lo; q < hi; ++q) sum += cnt[q];
    s[tid] = sum;
    __syncthreads();
    if (tid == 0) { int run = 0; for (int q = 0; q < nt; ++q) { int v = s[q]; s[q] = run; run += v; } ptr[m] = run; }
    __syncthreads();
    int run = s[tid];
    for (int q = lo; q < hi; ++q) { ptr[q] = run; run += cnt[q]; }
}
static __global__ void k_scatter_pos(int nz, const int* __restrict__ ia, const int* __restrict__ ptr,
                              int* __restrict__ fill, int* __restrict__ pos)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nz; k += gridDim.x * blockDim.x) {
        int r = ia[k];
        pos[ptr[r] + atomicAdd(&fill[r], 1)] = k;
    }
}
// one warp per output column: rank-sort the source positions (ascending position == ascending
// input column, stable), then emit (input column, value)
static __global__ void k_transpose_emit(int m, int n, const int* __restrict__ ka, const double* __restrict__ a,
                                 const int* __restrict__ ptr, const int* __restrict__ pos,
                                 int* __restrict__ iat, double* __restrict__ at)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < m; r += nwarps) {
        const int b = ptr[r], e = ptr[r + 1];
        for (int t = b + lane; t < e; t += 32) {
            const int k = pos[t];
            int rank = 0;
            for (int u = b; u < e; ++u) rank += (pos[u] < k) ? 1 : 0;
            // input column of position k: largest j with ka[j] <= k
            int lo = 0, hi = n;
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (ka[mid] <= k) lo = mid; else hi = mid; }
            iat[b + rank] = lo;
            at[b + rank] = a[k];
        }
    }
}

// --------------------------------------------------------------------------------------------
// K10  vector kernels of the METHOD loops.  Expression shapes are those of the reference source
// lines quoted beside them: with FMA contraction off each line rounds exactly like the C build.
// --------------------------------------------------------------------------------------------
static __global__ void k_fill(int n, double v, double* __restrict__ x)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) x[t] = v;
}
static __global__ void k_neg_copy(int n, const double* __restrict__ src, double* __restrict__ dst)
{   // fx[j] = -sigma[j]; gx[j] = -c[j]  (hsd.c:220,225)
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[t] = -src[t];
}
static __global__ void k_copy(int n, const double* __restrict__ src, double* __restrict__ dst)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) dst[t] = src[t];
}
static __global__ void k_ratio(int n, const double* __restrict__ num, const double* __restrict__ den, double* __restrict__ out)
{   // D[j] = z[j]/x[j]; E[i] = w[i]/y[i]  (hsd.c:215-216)
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) out[t] = num[t] / den[t];
}
// hsd.c:183-185 / 192-194.  rho = rho - b*phi + w ;  sigma = -sigma + c*phi + z
static __global__ void k_hsd_infeas(int n, int dual, double phi, const double* __restrict__ bc,
                             const double* __restrict__ wz, double* __restrict__ v)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        if (dual) v[t] = -v[t] + bc[t] * phi + wz[t];
        else      v[t] = v[t] - bc[t] * phi + wz[t];
    }
}
// hsd.c:187-189 / 196-198.  v = -(1-delta)*v + wz - delta*mu/yx
static __global__ void k_hsd_rhs(int n, double delta, double mu, const double* __restrict__ wz,
                          const double* __restrict__ yx, double* __restrict__ v)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        v[t] = -(1 - delta) * v[t] + wz[t] - delta * mu / yx[t];
}
// hsd.c:233-234.  d = f - g*dphi
static __global__ void k_hsd_dir(int n, double dphi, const double* __restrict__ f, const double* __restrict__ g, double* __restrict__ d)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) d[t] = f[t] - g[t] * dphi;
}
// hsd.c:236-237 / intpt.c:204-205.  dz = dmu/x - z - D*dx   (dmu = delta*mu, resp. mu)
static __global__ void k_comp_dir(int n, double dmu, const double* __restrict__ x, const double* __restrict__ z,
                           const double* __restrict__ D, const double* __restrict__ dx, double* __restrict__ dz)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
        dz[t] = dmu / x[t] - z[t] - D[t] * dx[t];
}
// ratio test (hsd.c:248-255, intpt.c:211-218): max over -dx/x and -dz/z; values below 0 never win
// because theta starts at 0; NaN never wins because "theta < NaN" is false
static __global__ void k_ratio_test(int n, const double* __restrict__ dx, const double* __restrict__ x,
                             const double* __restrict__ dz, const double* __restrict__ z,
                             unsigned long long* __restrict__ slot)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        double a = -dx[t] / x[t], b = -dz[t] / z[t];
        if (a > 0.0) atomicMax(slot, double_to_bits(a));
        if (b > 0.0) atomicMax(slot, double_to_bits(b));
    }
}
// Long-step variant (reference src/ipo/hsdls.c): per-component step length from the quadratic
//   (x + t dx)(z + t dz) >= (1 - beta) mu(t)      (linesearch, hsdls.c:296-336, same expression shapes => same roundings)
__device__ __forceinline__ double vbk_linesearch(double xj, double zj, double dxj, double dzj, double beta, double delta, double mu)
{
    const double a = dxj * dzj;
    const double b = zj * dxj + xj * dzj + (1 - beta) * (1 - delta) * mu;
    const double c = xj * zj - (1 - beta) * mu;
    const double d = b * b - 4 * a * c;
    if (a == 0.0) return -c / b;
    if (a > 0) {
        if (b < 0) return d >= 0 ? 2 * c / (-b + sqrt(d)) : HUGE_VAL;
        return HUGE_VAL;
    }
    if (b < 0) return 2 * c / (-b + sqrt(d));
    return (-b - sqrt(d)) / (2 * a);
}
// doubles -> unsigned keys with the same order (negative values included), for atomicMin
__device__ __forceinline__ unsigned long long vbk_order_key(double v) {
    const unsigned long long b = double_to_bits(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double vbk_order_unkey(unsigned long long k) {
    return bits_to_double((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k);
}
// hsdls.c:222-236 folds theta = MIN(theta, linesearch(...)) over the components with the macro MIN(a,b) = a<b ? a : b:
// a NaN value REPLACES the running minimum and is itself replaced by the next value.  The fold therefore equals the
// minimum over the entries after the last NaN.  Pass 1 stores the values and finds the last NaN, pass 2 takes the minimum
// behind it (both order-free), the host finishes the fold with the initial 1.0 and the (phi, psi) term.
static __global__ void k_linesearch(int n, int base, const double* __restrict__ x, const double* __restrict__ z,
                                    const double* __restrict__ dx, const double* __restrict__ dz, double beta, double delta,
                                    double mu, double* __restrict__ vals, int* __restrict__ lastnan)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const double v = vbk_linesearch(x[t], z[t], dx[t], dz[t], beta, delta, mu);
        vals[base + t] = v;
        if (v != v) atomicMax(lastnan, base + t);
    }
}
static __global__ void k_min_after(int total, const double* __restrict__ vals, const int* __restrict__ lastnan,
                                   unsigned long long* __restrict__ slot)
{
    const int first = *lastnan + 1;
    for (int t = first + blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x)
        atomicMin(slot, vbk_order_key(vals[t]));
}
// x += theta*dx; z += theta*dz  (hsd.c:265-268)
static __global__ void k_step2(int n, double theta, const double* __restrict__ dx, const double* __restrict__ dz,
                        double* __restrict__ x, double* __restrict__ z)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        x[t] = x[t] + theta * dx[t];
        z[t] = z[t] + theta * dz[t];
    }
}
// x /= phi; z /= phi (hsd.c:277-284)
static __global__ void k_scale2(int n, double phi, double* __restrict__ x, double* __restrict__ z)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        x[t] /= phi;
        z[t] /= phi;
    }
}
// intpt.c:140 / 146.  rho = b - rho - w ;  sigma = c - sigma + z
static __global__ void k_pf_infeas(int n, int dual, const double* __restrict__ bc, const double* __restrict__ wz, double* __restrict__ v)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        if (dual) v[t] = bc[t] - v[t] + wz[t];
        else      v[t] = bc[t] - v[t] - wz[t];
    }
}
// intpt.c:199-200.  dx = sigma - z + mu/x ;  dy = rho + w - mu/y
static __global__ void k_pf_rhs(int n, int dual, double mu, const double* __restrict__ v, const double* __restrict__ wz,
                         const double* __restrict__ yx, double* __restrict__ d)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        if (dual) d[t] = v[t] - wz[t] + mu / yx[t];
        else      d[t] = v[t] + wz[t] - mu / yx[t];
    }
}

}  // namespace vbk
