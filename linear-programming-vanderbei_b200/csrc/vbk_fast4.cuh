// vbk_fast4.cuh -- software-pipelined rank-k update of the dense window (fast mode).
//
// k_dense_update_k (vbk_fast2.cuh) stages a 32-deep slab of both operands with plain loads, waits, computes,
// and repeats: with one 256-thread CTA per SM (64 accumulators per thread) every slab exposes a full L2/HBM
// round trip, and the kernel stops at ~9 TFLOP/s (25 % of the measured 37 TFLOP/s DFMA rate).  Here the same
// 128 x 128 tile / 8 x 8 register micro-tile is fed by a three-stage cp.async ring of 16-deep slabs: the copies
// of slab s+2 are in flight while slab s is multiplied, one barrier per slab.  The arithmetic is unchanged
// (acc += L21[r,k] * (L21 D)[c,k], k ascending, explicit fma), so results are bit-identical to k_dense_update_k.
//
// Operand traffic per tile: 2 * klen * 128 doubles from L2 (256 KB at klen = 128) for 2 * 128 * 128 * klen flops;
// shared-memory reads: 8 LDS.128 per 64 DFMA and thread, 12 wavefronts per warp and k step against 128 cycles of
// the FP64 pipe -- the kernel is bound by the DFMA pipe (64 DFMA / clk / SM), which is the roofline it is held to.
#pragma once
#include "vbk_fast3.cuh"

namespace vbk {

#ifndef VBK_EMU
constexpr int kUpKC = 16;                      // slab depth
constexpr int kUpStages = 3;
constexpr int kUpT = 128;                      // tile edge
constexpr size_t kUpdPipeSmem = sizeof(double) * kUpStages * 2 * kUpKC * kUpT;   // 96 KB

// rank-klen update of the strictly-lower part of S[rbase.., rbase..cmax): S[r,c] -= sum_k S[r,kcol0+k] * P[c,pcol0+k]
static __global__ void __launch_bounds__(256, 1) k_dense_update_p(DenseArgs a)
{
    VBK_DYN_SMEM(raw);
    double* sm = reinterpret_cast<double*>(raw);            // [stage][A|B][kUpKC][kUpT]
    const int tr = blockIdx.y, tc = blockIdx.x;
    const int r0 = a.rbase + tr * kUpT, c0 = a.rbase + tc * kUpT;
    if (r0 + kUpT <= c0 || c0 >= a.cmax) return;             // tile entirely above the diagonal / outside
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const double* Ag = a.S + (size_t)a.kcol0 * a.ld;
    const double* Bg = a.P + (size_t)a.pcol0 * a.W;
    // 16-byte copies need every column of both operands 16-byte aligned: even leading dimensions and even
    // tile origins (the bases come from cudaMalloc)
    const bool al16 = ((a.ld | a.W | r0 | c0) & 1) == 0 && ((((size_t)a.S) | ((size_t)a.P)) & 15) == 0;
    const int nslab = (a.klen + kUpKC - 1) / kUpKC;

    auto issue = [&](int s) {
        if (s < nslab) {
            double* As = sm + (size_t)(s % kUpStages) * 2 * kUpKC * kUpT;
            double* Bs = As + kUpKC * kUpT;
            const int kc = s * kUpKC;
            if (al16) {
#pragma unroll
                for (int i = 0; i < kUpKC * kUpT / 2 / 256; ++i) {
                    const int e = tid + i * 256, x = (e & 63) * 2, c = e >> 6;
                    const bool kok = kc + c < a.klen;
                    const bool aok = kok && r0 + x < a.W, bok = kok && c0 + x < a.W;
                    cp_async16(As + c * kUpT + x, aok ? Ag + (size_t)(r0 + x) + (size_t)(kc + c) * a.ld : Ag, aok);
                    cp_async16(Bs + c * kUpT + x, bok ? Bg + (size_t)(c0 + x) + (size_t)(kc + c) * a.W : Bg, bok);
                }
            } else {
#pragma unroll
                for (int i = 0; i < kUpKC * kUpT / 256; ++i) {
                    const int e = tid + i * 256, x = e & 127, c = e >> 7;
                    const bool kok = kc + c < a.klen;
                    const bool aok = kok && r0 + x < a.W, bok = kok && c0 + x < a.W;
                    cp_async8(As + c * kUpT + x, aok ? Ag + (size_t)(r0 + x) + (size_t)(kc + c) * a.ld : Ag, aok);
                    cp_async8(Bs + c * kUpT + x, bok ? Bg + (size_t)(c0 + x) + (size_t)(kc + c) * a.W : Bg, bok);
                }
            }
        }
        cp_async_commit();                                   // one group per slab, empty past the end
    };

    double acc[8][8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int v = 0; v < 8; ++v) acc[u][v] = 0.0;

    issue(0);
    issue(1);
    for (int s = 0; s < nslab; ++s) {
        cp_async_wait<kUpStages - 2>();                      // slab s has landed (this thread's copies)
        __syncthreads();                                     // ... everyone's; and slab s-1 is no longer read
        issue(s + 2);                                        // refills the buffer slab s-1 used
        const double* As = sm + (size_t)(s % kUpStages) * 2 * kUpKC * kUpT;
        const double* Bs = As + kUpKC * kUpT;
#pragma unroll 4
        for (int c = 0; c < kUpKC; ++c) {
            double av[8], bv[8];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double2 x = *reinterpret_cast<const double2*>(As + c * kUpT + 2 * tx + 32 * u);
                const double2 y = *reinterpret_cast<const double2*>(Bs + c * kUpT + 2 * ty + 32 * u);
                av[2 * u] = x.x; av[2 * u + 1] = x.y; bv[2 * u] = y.x; bv[2 * u + 1] = y.y;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < 8; ++v) acc[u][v] = fma(av[u], bv[v], acc[u][v]);
        }
    }
    cp_async_wait<0>();

    // Epilogue.  A load-subtract-store sequence per element is serialised by the compiler (a store may alias
    // the next load): 64 dependent L2 round trips, longer than the whole main loop (ncu source page, round 1).
    // Two columns at a time: 16 independent loads, then the arithmetic, then the stores.
    const bool interior = r0 >= c0 + kUpT && r0 + kUpT <= a.W && c0 + kUpT <= a.cmax && ((a.ld | r0) & 1) == 0
                          && (((size_t)a.S) & 15) == 0;
    if (interior) {
#pragma unroll
        for (int vb = 0; vb < 8; vb += 4) {
            double2 t[4][4];
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    t[v][u] = *reinterpret_cast<const double2*>(&SW(a, r0 + 2 * tx + 32 * u, c0 + 2 * ty + 32 * ((vb + v) >> 1) + ((vb + v) & 1)));
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int u = 0; u < 4; ++u) { t[v][u].x -= acc[2 * u][vb + v]; t[v][u].y -= acc[2 * u + 1][vb + v]; }
#pragma unroll
            for (int v = 0; v < 4; ++v)
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    *reinterpret_cast<double2*>(&SW(a, r0 + 2 * tx + 32 * u, c0 + 2 * ty + 32 * ((vb + v) >> 1) + ((vb + v) & 1))) = t[v][u];
        }
        return;
    }
#pragma unroll
    for (int vb = 0; vb < 8; vb += 2) {
        double t[2][8];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int c2 = c0 + 2 * ty + 32 * ((vb + v) >> 1) + ((vb + v) & 1);
            const bool cok = c2 < a.W && c2 < a.cmax;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + 2 * tx + 32 * (u >> 1) + (u & 1);
                t[v][u] = (cok && r < a.W && r > c2) ? SW(a, r, c2) : 0.0;
            }
        }
#pragma unroll
        for (int v = 0; v < 2; ++v)
#pragma unroll
            for (int u = 0; u < 8; ++u) t[v][u] -= acc[u][vb + v];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int c2 = c0 + 2 * ty + 32 * ((vb + v) >> 1) + ((vb + v) & 1);
            const bool cok = c2 < a.W && c2 < a.cmax;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = r0 + 2 * tx + 32 * (u >> 1) + (u & 1);
                if (cok && r < a.W && r > c2) SW(a, r, c2) = t[v][u];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// The same update on the FP64 tensor path: mma.sync.m8n8k4 (DMMA).  Measured on B200 (scratch/ubench.cu,
// profiles/): DMMA peaks at the same 37 TFLOP/s as DFMA (it is the same pipe -- interleaving both adds nothing),
// but it reaches that peak with 4 warps per SM and 2 accumulators per warp, where an 8 x 8 DFMA outer product
// needs 8 warps to reach 33 TFLOP/s and spends every second issue slot on it.  More important here: operands.
// The DFMA tile reads 8 LDS.128 per k and warp (24 shared-memory cycles against 32 cycles of FP64 pipe: 76 % of
// the shared-memory pipe at full FP64 rate); a 64 x 32 warp tile of DMMA fragments reads 12 LDS.64 per FOUR k
// (6 cycles per k).  Fragment layout (PTX ISA, m8n8k4 .f64): with g = lane / 4, t = lane % 4 a thread holds
// A[g][t], B[t][g] and C[g][2t], C[g][2t+1].  Rows of a slab are padded to 132 doubles: a half-warp's 16 lanes
// (g < 4, t < 4) then hit banks 4 t + g -- all different.
// Not bit-identical to the DFMA kernels (the four products of one instruction are summed inside the tensor
// path); fast mode is the tolerance mode.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMmLd = 132;                     // padded slab row of the 128-row operand (banks 4 t + g)

// TN = columns per tile (128 or 64).  128 x 64 tiles run as 128-thread CTAs, two per SM: one CTA's epilogue and
// pipeline fill (about 40 % of a 128-deep tile's life, measured) overlap the other's main loop.
template <int TN> struct UpdMma {
    static constexpr int kThreads = 2 * (TN / 32) * 32;
    static constexpr int kLdB = TN + 4;
    static constexpr int kStage = kUpKC * (kMmLd + kLdB);                     // doubles per stage
    static constexpr size_t kSmem = sizeof(double) * kUpStages * kStage;
};

template <int TN>
static __global__ void __launch_bounds__(UpdMma<TN>::kThreads, 256 / UpdMma<TN>::kThreads) k_dense_update_m(DenseArgs a)
{
    using U = UpdMma<TN>;
    VBK_DYN_SMEM(raw);
    double* sm = reinterpret_cast<double*>(raw);            // [stage][A: kUpKC x kMmLd | B: kUpKC x kLdB]
    const int tr = blockIdx.y, tc = blockIdx.x;
    const int r0 = a.rbase + tr * kUpT, c0 = a.rbase + tc * TN;
    if (r0 + kUpT <= c0 || c0 >= a.cmax) return;             // tile entirely above the diagonal / outside
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 64, wn = (warp >> 1) * 32;   // this warp's 64 x 32 corner of the tile
    const double* Ag = a.S + (size_t)a.kcol0 * a.ld;
    const double* Bg = a.P + (size_t)a.pcol0 * a.W;
    const bool al16 = ((a.ld | a.W | r0 | c0) & 1) == 0 && ((((size_t)a.S) | ((size_t)a.P)) & 15) == 0;
    const int nslab = (a.klen + kUpKC - 1) / kUpKC;
    // a warp tile that lies entirely on or above the diagonal has nothing to compute
    const bool dead = r0 + wm + 63 <= c0 + wn;

    auto issue = [&](int s) {
        if (s < nslab) {
            double* As = sm + (size_t)(s % kUpStages) * U::kStage;
            double* Bs = As + kUpKC * kMmLd;
            const int kc = s * kUpKC;
            if (al16) {
#pragma unroll
                for (int i = 0; i < kUpKC * kUpT / 2 / U::kThreads; ++i) {
                    const int e = tid + i * U::kThreads, x = (e % (kUpT / 2)) * 2, c = e / (kUpT / 2);
                    const bool ok = kc + c < a.klen && r0 + x < a.W;
                    cp_async16(As + c * kMmLd + x, ok ? Ag + (size_t)(r0 + x) + (size_t)(kc + c) * a.ld : Ag, ok);
                }
#pragma unroll
                for (int i = 0; i < kUpKC * TN / 2 / U::kThreads; ++i) {
                    const int e = tid + i * U::kThreads, x = (e % (TN / 2)) * 2, c = e / (TN / 2);
                    const bool ok = kc + c < a.klen && c0 + x < a.W;
                    cp_async16(Bs + c * U::kLdB + x, ok ? Bg + (size_t)(c0 + x) + (size_t)(kc + c) * a.W : Bg, ok);
                }
            } else {
#pragma unroll
                for (int i = 0; i < kUpKC * kUpT / U::kThreads; ++i) {
                    const int e = tid + i * U::kThreads, x = e % kUpT, c = e / kUpT;
                    const bool ok = kc + c < a.klen && r0 + x < a.W;
                    cp_async8(As + c * kMmLd + x, ok ? Ag + (size_t)(r0 + x) + (size_t)(kc + c) * a.ld : Ag, ok);
                }
#pragma unroll
                for (int i = 0; i < kUpKC * TN / U::kThreads; ++i) {
                    const int e = tid + i * U::kThreads, x = e % TN, c = e / TN;
                    const bool ok = kc + c < a.klen && c0 + x < a.W;
                    cp_async8(Bs + c * U::kLdB + x, ok ? Bg + (size_t)(c0 + x) + (size_t)(kc + c) * a.W : Bg, ok);
                }
            }
        }
        cp_async_commit();
    };

    double acc[8][4][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }

    issue(0);
    issue(1);
    for (int s = 0; s < nslab; ++s) {
        cp_async_wait<kUpStages - 2>();
        __syncthreads();
        issue(s + 2);
        if (dead) continue;
        const double* As = sm + (size_t)(s % kUpStages) * U::kStage + wm + g;
        const double* Bs = sm + (size_t)(s % kUpStages) * U::kStage + kUpKC * kMmLd + wn + g;
#pragma unroll
        for (int k4 = 0; k4 < kUpKC / 4; ++k4) {
            double av[8], bv[4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) av[mi] = As[(k4 * 4 + t) * kMmLd + mi * 8];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) bv[ni] = Bs[(k4 * 4 + t) * U::kLdB + ni * 8];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], av[mi], bv[ni]);
        }
    }
    cp_async_wait<0>();
    if (dead) return;

    // epilogue: loads of one 8-column block (16 values per thread) first, then subtract, then store
    const int rb = r0 + wm + g, cb = c0 + wn + 2 * t;
    const bool interior = r0 >= c0 + TN && r0 + kUpT <= a.W && c0 + TN <= a.cmax;
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
        double tv[8][2];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = rb + mi * 8, c = cb + ni * 8 + j;
                const bool ok = interior || (r < a.W && c < a.W && c < a.cmax && r > c);
                tv[mi][j] = ok ? SW(a, r, c) : 0.0;
            }
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int j = 0; j < 2; ++j) tv[mi][j] -= acc[mi][ni][j];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = rb + mi * 8, c = cb + ni * 8 + j;
                const bool ok = interior || (r < a.W && c < a.W && c < a.cmax && r > c);
                if (ok) SW(a, r, c) = tv[mi][j];
            }
    }
}
#endif  // !VBK_EMU

}  // namespace vbk
