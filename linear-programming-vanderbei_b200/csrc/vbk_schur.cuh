// vbk_schur.cuh -- fast mode, Schur complement of the sparse columns on the dense window.
//   k_schur_window2   one CTA per window column: the column's dense image (N - i doubles) lives in shared memory,
//                     contributor records are staged 256 at a time and applied one after the other by all threads
//                     (ascending order, fma: the window is the tolerance part of fast mode), one coalesced write.
#pragma once
#include "vbk_sparse_level.cuh"

namespace vbk {

// Schur assembly of one window column per CTA.  S need not be zeroed for the rows >= i of column i (all are written).
// NT threads per CTA = contributor records per staged batch.  Every thread holds two entries of each of the next four
// tails in flight, so tails of up to 2*NT entries never take an unprefetched trip to the L2: 256 threads for the
// small windows of the netlib LPs, 1024 for the long tails (~1000 entries) of the large multicommodity LPs.
#ifdef VBK_EMU
constexpr int kSchur2ThreadsSmall = 64;
constexpr int kSchur2ThreadsLarge = 128;
#else
constexpr int kSchur2ThreadsSmall = 256;
constexpr int kSchur2ThreadsLarge = 1024;
#endif
struct Schur2Args {
    int N, T, ld, cap;                 // cap: doubles of shared memory for the column image
    const int* kL; const int* iL; const double* L; const double* diag;
    const int* rowptr; const int* rk; const int* rj;       // ascending row lists: sparse columns come first
    const int* spend;                  // [W] end of the sparse prefix (columns < T) of every window row's list
    double* S; double* wmag;
};

template <int NT>
static __global__ void __launch_bounds__(NT) k_schur_window2(Schur2Args a)
{
    constexpr int kSchur2Threads = NT, kSchur2Batch = NT;
    VBK_DYN_SMEM(raw);
    double* acc = reinterpret_cast<double*>(raw);                         // [cap]
    double* sw = acc + a.cap;                                             // [batch] L_ij d_j
    double* red = sw + kSchur2Batch;                                      // [threads] reductions
    int* se0 = reinterpret_cast<int*>(red + kSchur2Threads);              // [batch] tail start
    int* se1 = se0 + kSchur2Batch;                                        // [batch] tail end
    const int tid = threadIdx.x;
    for (int i = a.T + blockIdx.x; i < a.N; i += gridDim.x) {
        const int len = a.N - i;
        for (int s = tid; s < len; s += kSchur2Threads) acc[s] = 0.0;
        __syncthreads();
        for (int k = a.kL[i] + tid; k < a.kL[i + 1]; k += kSchur2Threads) acc[a.iL[k] - i] = a.L[k];   // K[i+1.., i]
        double dsum = 0.0, mag = 0.0;
        // the sparse contributors are the leading part of the ascending row list
        const int t0 = a.rowptr[i], tend = a.spend[i - a.T];
        for (int tb = t0; tb < tend; tb += kSchur2Batch) {
            __syncthreads();
            const int t = tb + tid;
            int e0 = 0, e1 = 0; double w = 0.0;
            if (t < tend) {
                const int j = a.rj[t], k = a.rk[t];
                const double lij = a.L[k];
                w = lij * a.diag[j];
                const double p = lij * w;
                dsum += p;
                mag = fmax(mag, fabs(p));
                e0 = k + 1; e1 = a.kL[j + 1];
            }
            sw[tid] = w; se0[tid] = e0; se1[tid] = e1;
            __syncthreads();
            const int nb = (tend - tb < kSchur2Batch) ? (tend - tb) : kSchur2Batch;
            // contributors one after the other (two of them may meet in a row; a fixed order also keeps the result
            // the same from run to run), every thread on one tail; the next four tails' entries are in flight meanwhile
            int rq[4][2]; double vq[4][2];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    rq[u][h] = -1; vq[u][h] = 0.0;
                    if (u < nb) { const int e = se0[u] + tid + h * kSchur2Threads; if (e < se1[u]) { rq[u][h] = a.iL[e]; vq[u][h] = a.L[e]; } }
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
                            const int e = se0[b + 4 + u] + tid + h * kSchur2Threads;
                            if (e < se1[b + 4 + u]) { rq[u][h] = a.iL[e]; vq[u][h] = a.L[e]; }
                        }
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (b + u >= nb) break;                                   // uniform
                    const double wb = -sw[b + u];
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (rc[u][h] >= 0) acc[rc[u][h] - i] = fma(wb, vc[u][h], acc[rc[u][h] - i]);
                    for (int e = se0[b + u] + 2 * kSchur2Threads + tid; e < se1[b + u]; e += kSchur2Threads)
                        acc[a.iL[e] - i] = fma(wb, a.L[e], acc[a.iL[e] - i]);
                    __syncthreads();
                }
            }
        }
        // diagonal and its largest term
        red[tid] = dsum;
        __syncthreads();
        for (int sft = kSchur2Threads / 2; sft > 0; sft >>= 1) { if (tid < sft) red[tid] += red[tid + sft]; __syncthreads(); }
        const double dtot = red[0];
        __syncthreads();
        red[tid] = mag;
        __syncthreads();
        for (int sft = kSchur2Threads / 2; sft > 0; sft >>= 1) { if (tid < sft) red[tid] = fmax(red[tid], red[tid + sft]); __syncthreads(); }
        const double d0 = a.diag[i];
        if (tid == 0) { acc[0] = d0 - dtot; a.wmag[i - a.T] = fmax(fabs(d0), red[0]); }
        __syncthreads();
        double* col = a.S + (size_t)(i - a.T) * a.ld + (i - a.T);
        for (int s = tid; s < len; s += kSchur2Threads) col[s] = acc[s];
        __syncthreads();
    }
}

}  // namespace vbk
