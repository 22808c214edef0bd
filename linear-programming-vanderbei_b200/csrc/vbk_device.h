// vbk_device.h -- thin layer between the kernels and the CUDA toolchain.
//
// Product build: nvcc, sm_100a, real CUDA runtime.  The -DVBK_EMU branch exists only so that the
// CPU-only unit tests (tests/emu) can run the same kernel sources on tiny inputs; it is never part
// of libvbkkt.so.
#pragma once

#ifdef VBK_EMU
#include "cuda_emu.h"
#define VBK_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (smem), [&] { kernel(__VA_ARGS__); })
#define VBK_DYN_SMEM(ptrname) unsigned char* ptrname = emu::dyn_smem()
#define VBK_GRID_SYNC() emu_grid_sync()
#else
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define VBK_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define VBK_DYN_SMEM(ptrname) extern __shared__ __align__(16) unsigned char ptrname[]
#endif

// The reference's error convention is "print and exit(1)" (src/common/myalloc.h:17-24); a CUDA
// failure in the drop-in does the same so that no silent CPU path can ever be taken.
#define VBK_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            std::fprintf(stderr, "vbkkt: CUDA error %s at %s:%d: %s\n", #call, __FILE__,   \
                         __LINE__, cudaGetErrorString(e_));                                 \
            std::exit(1);                                                                   \
        }                                                                                   \
    } while (0)

#define VBK_CHECK_LAUNCH() VBK_CUDA(cudaGetLastError())

__device__ __forceinline__ long long vbk_clock() {
#ifdef VBK_EMU
    return 0;
#else
    return clock64();
#endif
}

// volatile (L1-bypassing, non-cached) load used by the dataflow waits
__device__ __forceinline__ int vbk_ld_volatile(const int* p) {
#ifdef VBK_EMU
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
    return *reinterpret_cast<const volatile int*>(p);
#endif
}

// gpu-scope acquire load / release publication used by the dataflow hand-offs (cheaper than the
// sequentially-consistent __threadfence(): no MEMBAR.SC, the acquire itself orders the loads after it)
__device__ __forceinline__ int vbk_ld_acquire(const int* p) {
#ifdef VBK_EMU
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
// orders everything that happened before (own writes and, through a preceding CTA barrier, the other
// threads' writes) ahead of a following flag write
__device__ __forceinline__ void vbk_fence_release() {
#ifdef VBK_EMU
    std::atomic_thread_fence(std::memory_order_seq_cst);
#else
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
#endif
}

__device__ __forceinline__ void vbk_pause() {      // spin-loop body: yields the OS thread in the host emulation
#ifdef VBK_EMU
    sched_yield();
#endif
}
// spin-loop body of a warp that is NOT on the critical path: sleeps, so that its polling does not take issue slots
// from the warp of the same scheduler that runs a dependent chain
__device__ __forceinline__ void vbk_backoff(unsigned ns) {
#ifdef VBK_EMU
    (void)ns; sched_yield();
#else
    __nanosleep(ns);
#endif
}
__device__ __forceinline__ void vbk_st_volatile(int* p, int v) {
#ifdef VBK_EMU
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
    *reinterpret_cast<volatile int*>(p) = v;
#endif
}
// CTA-scope acquire / release on a shared-memory flag (no separate MEMBAR on the waiting side)
__device__ __forceinline__ int vbk_lds_acquire(const int* p) {
#ifdef VBK_EMU
    return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
    return v;
#endif
}
__device__ __forceinline__ void vbk_sts_release(int* p, int v) {
#ifdef VBK_EMU
    __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
    asm volatile("st.release.cta.shared.s32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
#endif
}
