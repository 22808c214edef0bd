// vbk_dense_update.cuh -- rank-k update of the dense window on the FP64 tensor path (fast mode): DMMA
// (mma.sync.m8n8k4.f64) tiles fed by a three-stage cp.async ring of 16-deep slabs (see the comment block below).
// tcgen05 has no FP64 kind; DMMA reaches the same 37 TFLOP/s as DFMA with a quarter of the issue slots and of the
// shared-memory operand traffic (profiles/r01_ubench.txt).
#pragma once
#include "vbk_dense_panel.cuh"

namespace vbk {

#ifndef VBK_EMU
constexpr int kUpKC = 16;                      // slab depth
constexpr int kUpStages = 3;
constexpr int kUpT = 128;                      // tile edge
constexpr size_t kUpdPipeSmem = sizeof(double) * kUpStages * 2 * kUpKC * kUpT;   // 96 KB

// ---------------------------------------------------------------------------------------------------------
// The same update on the FP64 tensor path: mma.sync.m8n8k4 (DMMA).  Measured on B200 (profiles/ubench/ubench.cu,
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

// TM x TN = rows x columns per tile.  128 x 64 tiles (the trailing update) run as 128-thread CTAs, two per SM: one
// CTA's epilogue and pipeline fill (about 40 % of a 128-deep tile's life, measured) overlap the other's main loop.
// 64 x 64 tiles serve the look-ahead strip (columns of the NEXT panel only): it sits on the critical path between
// two panel factorisations and must be spread over all SMs.  A warp owns kWR rows x 32 columns.
template <int TM, int TN> struct UpdMma {
    static constexpr int kWR = (TM == 128) ? 64 : 32;
    static constexpr int kWarps = (TM / kWR) * (TN / 32);
    static constexpr int kThreads = 32 * kWarps;
    static constexpr int kLdA = TM + 4, kLdB = TN + 4;                        // padded slab rows (banks 4 t + g)
    static constexpr int kStage = kUpKC * (kLdA + kLdB);                      // doubles per stage
    static constexpr size_t kSmem = sizeof(double) * kUpStages * kStage;
};

template <int TM, int TN>
static __global__ void __launch_bounds__(UpdMma<TM, TN>::kThreads, 256 / UpdMma<TM, TN>::kThreads) k_dense_update_m(DenseArgs a)
{
    using U = UpdMma<TM, TN>;
    constexpr int NI = U::kWR / 8;
    VBK_DYN_SMEM(raw);
    double* sm = reinterpret_cast<double*>(raw);            // [stage][A: kUpKC x kLdA | B: kUpKC x kLdB]
    const int tr = blockIdx.y, tc = blockIdx.x;
    const int r0 = a.rbase + a.rskip + tr * TM, c0 = a.rbase + tc * TN;
    if (r0 + TM <= c0 || c0 >= a.cmax || r0 >= a.W) return;             // tile entirely above the diagonal / outside
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp % (TM / U::kWR)) * U::kWR, wn = (warp / (TM / U::kWR)) * 32;   // this warp's corner of the tile
    const double* Ag = a.S + (size_t)a.kcol0 * a.ld;
    const double* Bg = a.P + (size_t)a.pcol0 * a.W;
    const bool al16 = ((a.ld | a.W | r0 | c0) & 1) == 0 && ((((size_t)a.S) | ((size_t)a.P)) & 15) == 0;
    const int nslab = (a.klen + kUpKC - 1) / kUpKC;
    // a warp tile that lies entirely on or above the diagonal has nothing to compute
    const bool dead = r0 + wm + U::kWR - 1 <= c0 + wn;

    // Copy addressing, hoisted (the per-slab address arithmetic was 16 % of the kernel's issue slots, ncu source
    // page of round 1): with 16-byte copies a thread always moves the same pair of rows, slab row c = ca + i * cstep.
    constexpr int kRowsA = 2 * U::kThreads / TM, kRowsB = 2 * U::kThreads / TN;   // slab rows covered per pass
    const int xa = (tid % (TM / 2)) * 2, ca = tid / (TM / 2);
    const int xb = (tid % (TN / 2)) * 2, cbr = tid / (TN / 2);
    const bool aok = r0 + xa < a.W, bok = c0 + xb < a.W;
    const double* pa = aok ? Ag + (size_t)(r0 + xa) + (size_t)ca * a.ld : Ag;
    const double* pb = bok ? Bg + (size_t)(c0 + xb) + (size_t)cbr * a.W : Bg;
    const unsigned sa0 = (unsigned)__cvta_generic_to_shared(sm + ca * U::kLdA + xa);
    const unsigned sb0 = (unsigned)__cvta_generic_to_shared(sm + kUpKC * U::kLdA + cbr * U::kLdB + xb);
    const size_t astep = (size_t)kRowsA * a.ld, bstep = (size_t)kRowsB * a.W;

    auto issue = [&](int s) {
        if (s < nslab) {
            const int kc = s * kUpKC;
            if (al16) {
                const unsigned so = (unsigned)((s % kUpStages) * U::kStage * sizeof(double));
                const double* qa = pa + (size_t)kc * a.ld;
                const double* qb = pb + (size_t)kc * a.W;
#pragma unroll
                for (int i = 0; i < kUpKC / kRowsA; ++i) {
                    const int n = (aok && kc + ca + i * kRowsA < a.klen) ? 16 : 0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n"
                                 ::"r"(sa0 + so + (unsigned)(i * kRowsA * U::kLdA * sizeof(double))), "l"(n ? qa + i * astep : Ag), "r"(n) : "memory");
                }
#pragma unroll
                for (int i = 0; i < kUpKC / kRowsB; ++i) {
                    const int n = (bok && kc + cbr + i * kRowsB < a.klen) ? 16 : 0;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n"
                                 ::"r"(sb0 + so + (unsigned)(i * kRowsB * U::kLdB * sizeof(double))), "l"(n ? qb + i * bstep : Bg), "r"(n) : "memory");
                }
            } else {
                double* As = sm + (size_t)(s % kUpStages) * U::kStage;
                double* Bs = As + kUpKC * U::kLdA;
#pragma unroll
                for (int i = 0; i < kUpKC * TM / U::kThreads; ++i) {
                    const int e = tid + i * U::kThreads, x = e % TM, c = e / TM;
                    const bool ok = kc + c < a.klen && r0 + x < a.W;
                    cp_async8(As + c * U::kLdA + x, ok ? Ag + (size_t)(r0 + x) + (size_t)(kc + c) * a.ld : Ag, ok);
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

    // The MMA's M index runs over COLUMNS of S (operand "A" = rows of P), its N index over ROWS of S (operand "B" =
    // rows of L21): a thread then holds C[g][2t], C[g][2t+1] = two consecutive rows of one column of S, and the
    // read-modify-write of the tile is 16 bytes wide.
    double acc[4][NI][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < NI; ++ni) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }

    issue(0);
    issue(1);
    for (int s = 0; s < nslab; ++s) {
        cp_async_wait<kUpStages - 2>();
        __syncthreads();
        issue(s + 2);
        if (dead) continue;
        const double* As = sm + (size_t)(s % kUpStages) * U::kStage + wm + g;                       // rows of S
        const double* Bs = sm + (size_t)(s % kUpStages) * U::kStage + kUpKC * U::kLdA + wn + g;     // columns of S
#pragma unroll
        for (int k4 = 0; k4 < kUpKC / 4; ++k4) {
            double cv[4], rv[NI];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) cv[mi] = Bs[(k4 * 4 + t) * U::kLdB + mi * 8];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) rv[ni] = As[(k4 * 4 + t) * U::kLdA + ni * 8];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], cv[mi], rv[ni]);
        }
    }
    cp_async_wait<0>();
    if (dead) return;

    // epilogue, one 8-column block at a time: 8 independent loads, subtract, store (a load-subtract-store per element
    // is serialised by the compiler -- a store may alias the next load)
    const int rb = r0 + wm + 2 * t, cb = c0 + wn + g;
    const bool interior = r0 >= c0 + TN && r0 + TM <= a.W && c0 + TN <= a.cmax;
    if (interior && al16) {
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            double2 tv[NI];
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) tv[ni] = *reinterpret_cast<const double2*>(&SW(a, rb + ni * 8, cb + mi * 8));
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) { tv[ni].x -= acc[mi][ni][0]; tv[ni].y -= acc[mi][ni][1]; }
#pragma unroll
            for (int ni = 0; ni < NI; ++ni) *reinterpret_cast<double2*>(&SW(a, rb + ni * 8, cb + mi * 8)) = tv[ni];
        }
        return;
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
        double tv[NI][2];
        const int c = cb + mi * 8;
        const bool cok = c < a.W && c < a.cmax;
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = rb + ni * 8 + j;
                tv[ni][j] = (cok && r < a.W && r > c) ? SW(a, r, c) : 0.0;
            }
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int j = 0; j < 2; ++j) tv[ni][j] -= acc[mi][ni][j];
#pragma unroll
        for (int ni = 0; ni < NI; ++ni)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = rb + ni * 8 + j;
                if (cok && r < a.W && r > c) SW(a, r, c) = tv[ni][j];
            }
    }
}
#endif  // !VBK_EMU

}  // namespace vbk
