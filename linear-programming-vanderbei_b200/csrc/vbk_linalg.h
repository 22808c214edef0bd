// vbk_linalg.h -- device linear-algebra context for the reference's linalg.h entry points.
#pragma once
#include "vbk_kkt.h"

namespace vbk {

struct DotJob;
constexpr int kStrictMode = 0;

class LinalgContext {
public:
    // shared_stream: run on an existing stream (the factor object's) instead of creating one
    LinalgContext(int device, int mode, bool shared_stream = false, cudaStream_t s = 0);
    ~LinalgContext();

    // host-buffer entry points (B1 seam); argument meaning as in reference linalg.h:1-8
    double dotprod_host(const double* x, const double* y, int n);
    double maxv_host(const double* x, int n);
    void smx_host(int m, int n, const double* a, const int* ka, const int* ia, const double* x, double* y);
    void atnum_host(int m, int n, const int* ka, const int* ia, const double* a, int* kat, int* iat, double* at);

    // device-buffer entry points (used by the device-resident METHOD loops)
    void dots_dev(const DotJob* jobs, int count, double* host_out);   // up to 8 dot products, one launch
    double absmax_dev(const double* d_x, int n);

    cudaStream_t stream() const { return stream_; }
    int grid(long long n) const;
    int mode() const { return mode_; }
    long long launches = 0;

private:
    void transpose_dev(int m, int n, int nz);
    int device_, mode_, num_sms_ = 1;
    cudaStream_t stream_ = 0;
    bool own_stream_ = true;
    double* pin_ = nullptr;
    DevArray<double> out_, vx_, vy_, a_, at_;
    DevArray<unsigned long long> bits_;
    DevArray<int> ka_, ia_, kat_, iat_, cnt_, fill_, pos_;
};

// yardsticks for the roofline denominators that MEASURED_PEAKS.json does not carry
double measure_fp64_tflops(int device);   // dependent-free DFMA streams on every SM
double measure_hbm_gbs(int device);       // device-to-device copy of 1 GiB

}  // namespace vbk
