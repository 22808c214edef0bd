// vbk_linalg.cu -- device implementations behind the reference's linalg.h entry points
// (smx, atnum, dotprod, maxv; reference src/common/linalg.c) for HOST buffers: the B1 seam.
// Each call moves its operands to the GPU, runs the kernels, and moves the result back.
#include "vbk_linalg.h"
#include "vbk_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace vbk {

LinalgContext::LinalgContext(int device, int mode, bool shared_stream, cudaStream_t s)
    : device_(device), mode_(mode)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= device_ || device_ < 0) {
        std::fprintf(stderr, "vbkkt: no CUDA device %d available (%s); this library has no CPU path\n",
                     device_, e != cudaSuccess ? cudaGetErrorString(e) : "device count too small");
        std::exit(1);
    }
    VBK_CUDA(cudaSetDevice(device_));
    cudaDeviceProp prop;
    VBK_CUDA(cudaGetDeviceProperties(&prop, device_));
    num_sms_ = prop.multiProcessorCount;
    if (shared_stream) { stream_ = s; own_stream_ = false; }
    else VBK_CUDA(cudaStreamCreate(&stream_));
    VBK_CUDA(cudaMallocHost((void**)&pin_, sizeof(double) * 16));
    out_.alloc(16);
    bits_.alloc(2);
}

LinalgContext::~LinalgContext()
{
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
    if (pin_) cudaFreeHost(pin_);
#ifndef VBK_EMU
    if (own_stream_) cudaStreamDestroy(stream_);
#endif
}

int LinalgContext::grid(long long n) const
{
    long long g = (n + kVecThreads - 1) / kVecThreads;
    if (g < 1) g = 1;
    long long cap = (long long)num_sms_ * 8;
    return (int)(g < cap ? g : cap);
}

// several dot products in one launch; device pointers
void LinalgContext::dots_dev(const DotJob* jobs, int count, double* host_out)
{
    VBK_CUDA(cudaSetDevice(device_));
    DotBatch b;
    b.count = count;
    for (int q = 0; q < count; ++q) b.job[q] = jobs[q];
    if (mode_ == kStrictMode)
        VBK_LAUNCH(k_dot_strict, count, kVecThreads, sizeof(double) * 2 * kDotChunk, stream_, b, out_.p);
    else
        VBK_LAUNCH(k_dot_tree, count, kVecThreads, sizeof(double) * kVecThreads, stream_, b, out_.p);
    VBK_CHECK_LAUNCH();
    launches++;
    out_.download(pin_, count, stream_);
    VBK_CUDA(cudaStreamSynchronize(stream_));
    for (int q = 0; q < count; ++q) host_out[q] = pin_[q];
}

double LinalgContext::absmax_dev(const double* d_x, int n)
{
    VBK_CUDA(cudaSetDevice(device_));
    VBK_CUDA(cudaMemsetAsync(bits_.p, 0, sizeof(unsigned long long), stream_));
    VBK_LAUNCH(k_absmax, grid(n), kVecThreads, 0, stream_, n, d_x, bits_.p);
    VBK_CHECK_LAUNCH();
    launches++;
    VBK_CUDA(cudaMemcpyAsync(pin_, bits_.p, 8, cudaMemcpyDeviceToHost, stream_));
    VBK_CUDA(cudaStreamSynchronize(stream_));
    return pin_[0];
}

double LinalgContext::dotprod_host(const double* x, const double* y, int n)
{
    vx_.upload(x, n, stream_);
    vy_.upload(y, n, stream_);
    DotJob j{vx_.p, vy_.p, n};
    double r = 0.0;
    dots_dev(&j, 1, &r);
    return r;
}

double LinalgContext::maxv_host(const double* x, int n)
{
    vx_.upload(x, n, stream_);
    return absmax_dev(vx_.p, n);
}

// device transpose of the CSC matrix currently held in (ka_, ia_, a_) -> (kat_, iat_, at_)
void LinalgContext::transpose_dev(int m, int n, int nz)
{
    kat_.alloc((size_t)m + 1); iat_.alloc(nz); at_.alloc(nz);
    cnt_.alloc(m); fill_.alloc(m); pos_.alloc(nz);
    VBK_CUDA(cudaMemsetAsync(cnt_.p, 0, sizeof(int) * (size_t)m, stream_));
    VBK_CUDA(cudaMemsetAsync(fill_.p, 0, sizeof(int) * (size_t)m, stream_));
    VBK_LAUNCH(k_hist, grid(nz), kVecThreads, 0, stream_, nz, ia_.p, cnt_.p);
    VBK_LAUNCH(k_scan_single, 1, kScanThreads, sizeof(int) * kScanThreads, stream_, m, cnt_.p, kat_.p);
    VBK_LAUNCH(k_scatter_pos, grid(nz), kVecThreads, 0, stream_, nz, ia_.p, kat_.p, fill_.p, pos_.p);
    VBK_LAUNCH(k_transpose_emit, grid((long long)m * 32), kVecThreads, 0, stream_, m, n, ka_.p, a_.p,
               kat_.p, pos_.p, iat_.p, at_.p);
    VBK_CHECK_LAUNCH();
    launches += 4;
}

void LinalgContext::atnum_host(int m, int n, const int* ka, const int* ia, const double* a,
                               int* kat, int* iat, double* at)
{
    VBK_CUDA(cudaSetDevice(device_));
    const int nz = ka[n];
    ka_.upload(ka, (size_t)n + 1, stream_);
    ia_.upload(ia, nz, stream_);
    a_.upload(a, nz, stream_);
    transpose_dev(m, n, nz);
    kat_.download(kat, (size_t)m + 1, stream_);
    iat_.download(iat, nz, stream_);
    at_.download(at, nz, stream_);
    VBK_CUDA(cudaStreamSynchronize(stream_));
}

void LinalgContext::smx_host(int m, int n, const double* a, const int* ka, const int* ia,
                             const double* x, double* y)
{
    VBK_CUDA(cudaSetDevice(device_));
    const int nz = ka[n];
    ka_.upload(ka, (size_t)n + 1, stream_);
    ia_.upload(ia, nz, stream_);
    a_.upload(a, nz, stream_);
    vx_.upload(x, n, stream_);
    vy_.alloc(m);
    // race-free gather over the transposed matrix; rows list their entries in ascending column
    // order, which is the order the reference's scatter loop adds them (linalg.c:67-69)
    transpose_dev(m, n, nz);
    VBK_LAUNCH(k_spmv_rows, grid(m), kVecThreads, 0, stream_, m, kat_.p, iat_.p, at_.p, vx_.p, vy_.p);
    VBK_CHECK_LAUNCH();
    launches++;
    vy_.download(y, m, stream_);
    VBK_CUDA(cudaStreamSynchronize(stream_));
}


// ------------------------------------------------------------------------------------------------
// Roofline yardsticks.  MEASURED_PEAKS.json carries the HBM copy bandwidth and the bf16 GEMM rate;
// the FP64 pipe rate that bounds the wide-front factor updates is measured here: eight independent
// DFMA chains per thread (explicit fma(), so -fmad=false does not matter), every SM full.
// ------------------------------------------------------------------------------------------------
static __global__ void k_fp64_yardstick(int iters, double a, double b, double* __restrict__ out)
{
    double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
    for (int i = 0; i < iters; ++i) {
        r0 = fma(r0, a, b); r1 = fma(r1, a, b); r2 = fma(r2, a, b); r3 = fma(r3, a, b);
        r4 = fma(r4, a, b); r5 = fma(r5, a, b); r6 = fma(r6, a, b); r7 = fma(r7, a, b);
    }
    double s = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7;
    if (s == 123.456) out[0] = s;   // keep the chains alive
}
static __global__ void k_copy_yardstick(size_t n, const double2* __restrict__ src, double2* __restrict__ dst)
{
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
        dst[t] = src[t];
}

double measure_fp64_tflops(int device)
{
    VBK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VBK_CUDA(cudaGetDeviceProperties(&prop, device));
    DevArray<double> out; out.alloc(1);
    cudaEvent_t e0, e1;
    VBK_CUDA(cudaEventCreate(&e0)); VBK_CUDA(cudaEventCreate(&e1));
    const int grid = prop.multiProcessorCount * 8, block = 256, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        VBK_CUDA(cudaEventRecord(e0, 0));
        VBK_LAUNCH(k_fp64_yardstick, grid, block, 0, 0, iters, 1.0000001, 1e-9, out.p);
        VBK_CUDA(cudaEventRecord(e1, 0));
        VBK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        VBK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * 8.0 * (double)iters * (double)grid * (double)block;
        if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

double measure_hbm_gbs(int device)
{
    VBK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    VBK_CUDA(cudaGetDeviceProperties(&prop, device));
    const size_t n = (size_t)1 << 26;                      // 2^26 double2 = 1 GiB
    DevArray<double> a, b; a.alloc(2 * n); b.alloc(2 * n);
    VBK_CUDA(cudaMemset(a.p, 0, 16 * n));
    cudaEvent_t e0, e1;
    VBK_CUDA(cudaEventCreate(&e0)); VBK_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        VBK_CUDA(cudaEventRecord(e0, 0));
        VBK_LAUNCH(k_copy_yardstick, prop.multiProcessorCount * 16, 512, 0, 0, n, (const double2*)a.p, (double2*)b.p);
        VBK_CUDA(cudaEventRecord(e1, 0));
        VBK_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        VBK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms > 0) best = std::max(best, 2.0 * 16.0 * (double)n / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

}  // namespace vbk
