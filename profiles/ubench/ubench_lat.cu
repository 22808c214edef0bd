// Micro-benchmark: dependent-load latency out of L2 under the conditions of the strict factor kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_lat ubench_lat.cu && ./ubench_lat
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ int ld_acq(const int* p) { int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
// mode: 0 plain chase (ld.ca), 1 ld.cg chase, 2 ld.cg chase + other warps spin on smem, 3 + other warps poll a global flag (volatile),
// 4 + other warps poll with ld.acquire, 5: ld.cg chase with 16 independent chains per lane (MLP)
__global__ void k(const int* __restrict__ next, int n, int iters, int mode, int* flag, long long* out, int* sink)
{
    __shared__ int sflag;
    if (threadIdx.x == 0) sflag = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        int p = (blockIdx.x * 977 + lane * 131) % n;
        long long t0 = clock64();
        if (mode == 5) {
            int q[16];
            for (int u = 0; u < 16; ++u) q[u] = (p + u * 4099) % n;
            for (int i = 0; i < iters; ++i)
                for (int u = 0; u < 16; ++u) q[u] = __ldcg(&next[q[u]]);
            for (int u = 0; u < 16; ++u) p ^= q[u];
        } else {
            for (int i = 0; i < iters; ++i) p = (mode == 0) ? next[p] : __ldcg(&next[p]);
        }
        long long t1 = clock64();
        if (lane == 0) out[blockIdx.x] = t1 - t0;
        sink[blockIdx.x * 32 + lane] = p;
        __syncwarp();
        if (lane == 0) { *(volatile int*)&sflag = 1; }
    } else {
        if (mode == 2) { while (*(volatile int*)&sflag == 0) {} }
        else if (mode == 3) { while (*(volatile int*)&sflag == 0) { if (*(volatile int*)flag == 12345) break; } }
        else if (mode == 4) { while (*(volatile int*)&sflag == 0) { if (ld_acq(flag) == 12345) break; } }
    }
}
int main()
{
    const int n = 8 << 20;   // 32 MB of ints: L2 resident
    int* h = new int[n];
    unsigned long long s = 88172645463325252ull;
    for (int i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % n); }
    int *d, *flag, *sink; long long* out;
    cudaMalloc(&d, n * 4); cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&flag, 4); cudaMemset(flag, 0, 4);
    cudaMalloc(&sink, 148 * 32 * 4); cudaMalloc(&out, 148 * 8);
    const int iters = 2000;
    const char* names[] = {"ld.ca chase", "ld.cg chase", "ld.cg chase, 15 warps spin on smem", "ld.cg chase, 15 warps poll global (volatile)",
                           "ld.cg chase, 15 warps poll global (ld.acquire)", "ld.cg, 16 independent chains per lane"};
    for (int grid : {1, 148}) for (int mode = 0; mode < 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) k<<<grid, 512>>>(d, n, iters, mode, flag, out, sink);
        cudaDeviceSynchronize();
        long long ho[148]; cudaMemcpy(ho, out, grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < grid; ++i) avg += ho[i]; avg /= grid;
        printf("grid %3d  %-48s %8.1f cycles per step%s\n", grid, names[mode], avg / iters, mode == 5 ? " (16 loads)" : "");
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
