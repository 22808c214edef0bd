// vbk_rowblock.cu -- the per-rank pieces of the row-block partitioned smx / dotprod / maxv
// (BASELINE.json config 5).
//
// One LP is spread over G GPUs: y-side vectors (length m) are block-partitioned by the rows of A, x-side
// vectors (length n) by its columns.  A rank computes  (A x)[rows]  from its row block of A (a slice of
// the transpose arrays atnum builds, src/common/linalg.c:75-103) after an all-gather of x, and
// (A^T y)[cols] from its column block after an all-gather of y.  Each output entry is one row sum taken
// in ascending column order, i.e. exactly the order in which the reference's scatter-form smx
// (linalg.c:62-70) adds into it: the distributed product is BIT-IDENTICAL to the reference's.
// dotprod / maxv become local partial reductions followed by one all-reduce of a few doubles
// (sum / max); the sum is a fixed-shape tree, so it is deterministic but not the reference's
// left-to-right order (linalg.c:17-25) -- tolerance parity, as in fast mode.
//
// The collectives themselves are torch.distributed / NCCL calls made by the host side
// (linear-programming-vanderbei_b200/rowblock.py); everything here takes device pointers and a
// stream and is asynchronous.
#include "../../include/vbkkt.h"
#include "vbk_kernels.cuh"

namespace vbk {

constexpr int kRedMaxJobs = 8;
constexpr int kRedMaxBlocks = 1024;            // partial sums per job
struct RedBatch { const double* x[kRedMaxJobs]; const double* y[kRedMaxJobs]; long long n[kRedMaxJobs]; int count; };

// stage 1: blockIdx.y = job, blockIdx.x = slice; grid-stride, two independent accumulators per thread
static __global__ void __launch_bounds__(kVecThreads) k_dot_partial(RedBatch b, double* __restrict__ partial)
{
    VBK_DYN_SMEM(raw);
    double* red = reinterpret_cast<double*>(raw);
    const int job = blockIdx.y;
    const double* __restrict__ x = b.x[job];
    const double* __restrict__ y = b.y[job];
    const long long n = b.n[job], stride = (long long)gridDim.x * blockDim.x;
    double a0 = 0.0, a1 = 0.0;
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; q + stride < n; q += 2 * stride) {
        a0 = fma(x[q], y[q], a0);
        a1 = fma(x[q + stride], y[q + stride], a1);
    }
    if (q < n) a0 = fma(x[q], y[q], a0);
    red[threadIdx.x] = a0 + a1;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[(size_t)job * gridDim.x + blockIdx.x] = red[0];
}
// stage 2: one CTA per job folds the partials in a fixed tree
static __global__ void __launch_bounds__(kVecThreads) k_dot_fold(int nblocks, const double* __restrict__ partial,
                                                                 double* __restrict__ out)
{
    VBK_DYN_SMEM(raw);
    double* red = reinterpret_cast<double*>(raw);
    double a = 0.0;
    for (int q = threadIdx.x; q < nblocks; q += blockDim.x) a += partial[(size_t)blockIdx.x * nblocks + q];
    red[threadIdx.x] = a;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = red[0];
}

// max |x| of up to 8 vectors: per-CTA maximum, one atomicMax on the (monotone) bit pattern per CTA
static __global__ void __launch_bounds__(kVecThreads) k_absmax_partial(RedBatch b, unsigned long long* __restrict__ out)
{
    VBK_DYN_SMEM(raw);
    double* red = reinterpret_cast<double*>(raw);
    const int job = blockIdx.y;
    const double* __restrict__ x = b.x[job];
    const long long n = b.n[job], stride = (long long)gridDim.x * blockDim.x;
    double m = 0.0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += stride) {
        const double a = vbk_abs(x[q]);
        if (a > m) m = a;
    }
    red[threadIdx.x] = m;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s && red[threadIdx.x + s] > red[threadIdx.x]) red[threadIdx.x] = red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicMax(&out[job], double_to_bits(red[0]));
}

static int red_grid(long long nmax)
{
    long long g = (nmax + (long long)kVecThreads * 8 - 1) / ((long long)kVecThreads * 8);
    if (g < 1) g = 1;
    if (g > kRedMaxBlocks) g = kRedMaxBlocks;
    return (int)g;
}

}  // namespace vbk

using namespace vbk;

extern "C" {

void vbk_spmv_rows_dev(int nrows, const int* ptr_dev, const int* idx_dev, const double* val_dev,
                       const double* x_dev, double* y_dev, void* stream)
{
    if (nrows <= 0) return;
    cudaStream_t st = (cudaStream_t)(size_t)stream;
    long long g = ((long long)nrows + kVecThreads - 1) / kVecThreads;
    if (g > 148 * 16) g = 148 * 16;
    VBK_LAUNCH(k_spmv_rows, (int)g, kVecThreads, 0, st, nrows, ptr_dev, idx_dev, val_dev, x_dev, y_dev);
    VBK_CHECK_LAUNCH();
}

int vbk_reduce_scratch_doubles(void) { return kRedMaxJobs * kRedMaxBlocks; }

void vbk_dots_partial_dev(int count, const double* const* x_dev, const double* const* y_dev, const long long* n,
                          double* out_dev, double* scratch_dev, void* stream)
{
    if (count <= 0) return;
    if (count > kRedMaxJobs) { std::fprintf(stderr, "vbkkt: at most %d dot products per launch\n", kRedMaxJobs); std::exit(1); }
    cudaStream_t st = (cudaStream_t)(size_t)stream;
    RedBatch b;
    b.count = count;
    long long nmax = 0;
    for (int q = 0; q < count; ++q) { b.x[q] = x_dev[q]; b.y[q] = y_dev[q]; b.n[q] = n[q]; if (n[q] > nmax) nmax = n[q]; }
    const int g = red_grid(nmax);
    VBK_LAUNCH(k_dot_partial, dim3(g, count), kVecThreads, sizeof(double) * kVecThreads, st, b, scratch_dev);
    VBK_LAUNCH(k_dot_fold, count, kVecThreads, sizeof(double) * kVecThreads, st, g, scratch_dev, out_dev);
    VBK_CHECK_LAUNCH();
}

void vbk_absmax_partial_dev(int count, const double* const* x_dev, const long long* n, double* out_dev, void* stream)
{
    if (count <= 0) return;
    if (count > kRedMaxJobs) { std::fprintf(stderr, "vbkkt: at most %d max-norms per launch\n", kRedMaxJobs); std::exit(1); }
    cudaStream_t st = (cudaStream_t)(size_t)stream;
    RedBatch b;
    b.count = count;
    long long nmax = 0;
    for (int q = 0; q < count; ++q) { b.x[q] = x_dev[q]; b.y[q] = nullptr; b.n[q] = n[q]; if (n[q] > nmax) nmax = n[q]; }
    VBK_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(double) * (size_t)count, st));
    // a non-negative double and its bit pattern order the same way: the u64 maximum IS the double maximum
    VBK_LAUNCH(k_absmax_partial, dim3(red_grid(nmax), count), kVecThreads, sizeof(double) * kVecThreads, st, b,
               reinterpret_cast<unsigned long long*>(out_dev));
    VBK_CHECK_LAUNCH();
}

}  // extern "C"
