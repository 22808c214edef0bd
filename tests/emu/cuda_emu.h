// tests/emu/cuda_emu.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny host emulation of the CUDA execution model, used by the CPU-only tests to run the very
// same kernel sources (csrc/*.cu compiled as C++ with -DVBK_EMU) on tiny inputs: every CUDA thread
// becomes an OS thread, __syncthreads()/__syncwarp() are std::barriers, warp shuffles go through a
// per-warp exchange buffer, atomics/fences map to GCC builtins.  It exists so that kernel logic
// (index maps, dataflow waits, accumulation order) is checked in the GPU-less container before
// GPU minutes are spent.  It is NOT a fallback: the product library libvbkkt.so is built by nvcc
// only, never contains this header, and exits loudly when no CUDA device is present.
#pragma once
#ifndef VBK_EMU
#error "cuda_emu.h is only for the -DVBK_EMU test build"
#endif

#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <sched.h>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __noinline__ __attribute__((noinline))
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

struct double2 { double x, y; };

namespace emu {
struct Cta {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<std::barrier<>>> warpbar;
    std::vector<unsigned char> smem;
    std::vector<unsigned long long> xchg;  // per-thread shuffle exchange slot
};
struct Grid {
    std::unique_ptr<std::barrier<>> bar;
};
inline thread_local Cta* cta = nullptr;
inline thread_local Grid* grid = nullptr;
}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

inline void __syncthreads() { emu::cta->bar->arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::cta->warpbar[threadIdx.x / 32]->arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __nanosleep(unsigned) { sched_yield(); }
inline void emu_grid_sync() { emu::grid->bar->arrive_and_wait(); }

template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }

inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicSub(int* p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline int atomicMin(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline double atomicAdd(double* p, double v) {
    unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = __atomic_load_n(q, __ATOMIC_SEQ_CST), nw;
    double o;
    do {
        std::memcpy(&o, &old, 8);
        double s = o + v;
        std::memcpy(&nw, &s, 8);
    } while (!__atomic_compare_exchange_n(q, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return o;
}

// warp shuffles: every lane of the (full) warp must call
template <class T> inline T emu_shfl_from(T v, int srclane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    emu::Cta* c = emu::cta;
    unsigned tid = threadIdx.x, warp = tid / 32, lane = tid % 32;
    unsigned nlanes = std::min(32u, blockDim.x - warp * 32);
    unsigned long long raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    c->xchg[tid] = raw;
    c->warpbar[warp]->arrive_and_wait();
    unsigned src = (srclane >= 0 && (unsigned)srclane < nlanes) ? (unsigned)srclane : lane;
    unsigned long long got = c->xchg[warp * 32 + src];
    c->warpbar[warp]->arrive_and_wait();
    T out;
    std::memcpy(&out, &got, sizeof(T));
    return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_shfl_from(v, src & 31); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d) { return emu_shfl_from(v, (int)(threadIdx.x % 32 + d)); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    int lane = (int)(threadIdx.x % 32);
    return emu_shfl_from(v, lane >= (int)d ? lane - (int)d : lane);
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl_from(v, (int)((threadIdx.x % 32) ^ m)); }

inline int __any_sync(unsigned, int pred) {
    int v = pred ? 1 : 0;
    for (int m = 16; m > 0; m >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned v = pred ? (1u << (threadIdx.x % 32)) : 0u;
    for (int m = 16; m > 0; m >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }

// ---- runtime API subset ----
typedef int cudaError_t;
typedef int cudaStream_t;
struct EmuEvent { std::chrono::steady_clock::time_point t; };
typedef EmuEvent* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
struct cudaDeviceProp { int multiProcessorCount; size_t sharedMemPerBlockOptin; char name[64]; int major, minor; size_t totalGlobalMem; };

inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    std::memset(p, 0, sizeof(*p));
    p->multiProcessorCount = 1; p->sharedMemPerBlockOptin = 227 * 1024; std::strcpy(p->name, "emu");
    p->major = 10; p->minor = 0; p->totalGlobalMem = (size_t)8 << 30;
    return 0;
}
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = 0; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now(); return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return 0;
}
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return 0; }

namespace emu {
// Run `body` once per emulated CUDA thread.  All CTAs run concurrently (needed by the dataflow
// kernels, whose CTAs wait on one another).
inline void launch(dim3 g, dim3 b, size_t smem_bytes, const std::function<void()>& body) {
    unsigned nthreads = b.x, nctas = g.x * g.y;
    Grid grid_state;
    grid_state.bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)nthreads * nctas);
    std::vector<Cta> ctas(nctas);
    for (auto& c : ctas) {
        c.bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)nthreads);
        for (unsigned w = 0; w * 32 < nthreads; ++w)
            c.warpbar.push_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)std::min(32u, nthreads - w * 32)));
        c.smem.assign(smem_bytes + 16, 0);
        c.xchg.assign(nthreads, 0);
    }
    std::vector<std::thread> pool;
    pool.reserve((size_t)nthreads * nctas);
    for (unsigned cb = 0; cb < nctas; ++cb)
        for (unsigned t = 0; t < nthreads; ++t)
            pool.emplace_back([&, cb, t] {
                cta = &ctas[cb];
                grid = &grid_state;
                threadIdx = dim3(t); blockIdx = dim3(cb % g.x, cb / g.x); blockDim = b; gridDim = g;
                body();
            });
    for (auto& th : pool) th.join();
}
inline unsigned char* dyn_smem() { return cta->smem.data(); }
}  // namespace emu
