// L2-resident gather of 256-byte segments, the access pattern of the strict factor's producers (csrc/vbk_strict_factor.cuh):
// every warp loads ROWS segments of 32 consecutive doubles at pseudo-random 8-byte-aligned offsets of a buffer that fits
// in L2, BATCH loads in flight per warp, and writes the products to shared memory.  Reports GB/s per SM and chip-wide for
// several (warps per CTA, batch) points, with one CTA per SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_gather ubench_gather.cu && ./ubench_gather
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int BATCH>
__global__ void k_gather(const double* __restrict__ buf, size_t nelem, int iters, double* sink, int active_ctas)
{
    extern __shared__ double tile[];
    if ((int)blockIdx.x >= active_ctas) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long s = 0x9E3779B97F4A7C15ull * (blockIdx.x * 64 + warp + 1);
    const unsigned mask = (unsigned)(nelem / 2 - 1);
    double acc = 0.0;
    double* my = tile + warp * 32 * BATCH;
    for (int it = 0; it < iters; ++it) {
        double v[BATCH];
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
            s = s * 6364136223846793005ull + 1442695040888963407ull;
            const size_t off = (size_t)((unsigned)(s >> 33) & mask);        // nelem is a power of two here
            v[u] = __ldcg(buf + off + lane);
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u) my[u * 32 + lane] = v[u] * 1.0000001;
        __syncwarp();
        acc += my[lane];
    }
    if (acc == 123.456) sink[0] = acc;
}

template <int BATCH>
void run(const double* buf, size_t nelem, double* sink, int sms, int warps, int active)
{
    const int iters = 4000 / BATCH;
    const size_t smem = (size_t)warps * 32 * BATCH * sizeof(double);
    cudaFuncSetAttribute(k_gather<BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather<BATCH><<<sms, warps * 32, smem>>>(buf, nelem, iters, sink, active);
    cudaEventRecord(e0);
    k_gather<BATCH><<<sms, warps * 32, smem>>>(buf, nelem, iters, sink, active);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = (double)active * warps * iters * BATCH * 256.0;
    printf("active CTAs %3d  warps/CTA %2d  loads in flight/warp %2d : %8.1f GB/s chip, %6.2f GB/s per SM, %6.1f ns per 8 KB group per SM  (%s)\n",
           active, warps, BATCH, bytes / ms / 1e6, bytes / ms / 1e6 / active, 8192.0 / (bytes / ms / 1e6 / active), cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t nelem = (size_t)1 << 24;             // 128 MiB of doubles, offsets drawn from the first 64 MiB (dfl001's L: 57 MiB)
    double* buf; double* sink;
    cudaMalloc(&buf, nelem * 8); cudaMalloc(&sink, 8);
    cudaMemset(buf, 0, nelem * 8);
    for (int active : {sms, 87, 16, 1}) {
        run<4>(buf, nelem, sink, sms, 14, active);
        run<16>(buf, nelem, sink, sms, 14, active);
        run<32>(buf, nelem, sink, sms, 14, active);
        run<16>(buf, nelem, sink, sms, 30, active);
    }
    return 0;
}
