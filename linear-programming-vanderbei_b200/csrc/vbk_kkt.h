// vbk_kkt.h -- the device-resident factor object ("K = [-E A; A^T D]" of one LP).
//
// Replaces the file-scope statics of the reference's LU plugin (src/ipo/ldlt.c:108-120) by a handle:
// symbolic analysis once on the host (vbk_symbolic), everything numeric on the GPU.
#pragma once
#include "vbk_device.h"
#include "vbk_symbolic.h"

#include <cstddef>
#include <vector>

namespace vbk {

enum Mode { kStrict = 0, kFast = 1 };
struct PipeArgs;
struct FlagSolveArgs;
struct SolveArgs;

// owning device array
template <class T>
struct DevArray {
    T* p = nullptr;
    size_t n = 0;
    DevArray() = default;
    DevArray(const DevArray&) = delete;
    DevArray& operator=(const DevArray&) = delete;
    ~DevArray() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    void alloc(size_t count) {
        if (count <= n && p) return;
        release();
        VBK_CUDA(cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T)));
        n = count;
    }
    void upload(const T* h, size_t count, cudaStream_t s) {
        alloc(count);
        if (count) VBK_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    void upload(const std::vector<T>& v, cudaStream_t s) { upload(v.data(), v.size(), s); }
    void download(T* h, size_t count, cudaStream_t s) const {
        if (count) VBK_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
};

struct KktStats {
    long long factor_calls = 0, solve_calls = 0, rawsolve_calls = 0;
    long long kernel_launches = 0;
    int last_passes = 0, last_consistent = 1, last_ndep = 0;
    int last_passes2[2] = {0, 0};     // refinement passes of the two right-hand sides of the last solve2
};

class Kkt {
public:
    Kkt(int device, int mode);
    ~Kkt();

    // ldlt-space arguments, as in ldltfac (ldlt.h:1-13): A is m x n in CSC, At its transpose.
    // Host pointers; the matrix is copied to the device once (the reference keeps the pointers
    // and assumes the matrix never changes, ldlt.c:140-160).  device < 0: host analysis only.
    void analyze(int m, int n, const int* kA, const int* iA, const double* A,
                 const int* kAt, const int* iAt, const double* At);
    bool analyzed() const { return analyzed_; }

    // numeric factorisation of K(dn, dm); device pointers of length n and m (inv_num, ldlt.c:164-309)
    void factor_dev(const double* d_dn, const double* d_dm);
    // K^{-1} rhs with iterative refinement; c (length n) and b (length m) are overwritten
    // (solve, ldlt.c:327-425).  Returns the `consistent` flag.
    int solve_dev(const double* d_Dn, const double* d_Dm, double* d_c, double* d_b);
    // two independent right-hand sides on the same factor in one pair of sweeps per refinement pass (hsd.c:223,228)
    void solve2_dev(const double* d_Dn, const double* d_Dm, double* d_c0, double* d_b0, double* d_c1, double* d_b1, int consistent[2]);
    int solve2_host(const double* Dn, const double* Dm, double* c0, double* b0, double* c1, double* b1);
    // one forward/diagonal/backward sweep on the permuted vector in zbuf() (rawsolve, ldlt.c:433-505)
    void rawsolve_dev();
    void rawsolve_rhs(int mask);
    void solve_rhs(int mask, const double* d_Dn, const double* d_Dm, double* const d_c[2], double* const d_b[2], int consistent[2]);

    // host-pointer wrappers (the B1 seam): H2D, device work, D2H
    void factor_host(const double* dn, const double* dm);
    int solve_host(const double* Dn, const double* Dm, double* c, double* b);

    // y[m] = A x[n]  /  y[n] = At x[m]  with the reference's summation order (smx, linalg.c:62-70)
    void spmv_A(const double* d_x, double* d_y);
    void spmv_At(const double* d_x, double* d_y);

    // introspection for tests / bench
    const Symbolic& sym() const { return sym_; }
    void download_factor(double* L, double* diag, int* mark);
    double epsdiag();
    int ndep();
    double* zbuf() { return z_.p; }
    cudaStream_t stream() const { return stream_; }
    int device() const { return device_; }
    int mode() const { return mode_; }
    KktStats stats;
    // device time of the last numeric-factor kernel (CUDA events on the handle's stream), ms
    float last_factor_kernel_ms();
    // $VBK_PROF=1: cycles spent per phase of the strict factor kernel since the last call (16 counters, see
    // vbk_strict_factor.cuh); zeros when profiling is off
    void read_phase_profile(unsigned long long out[16]);
    // $VBK_PROF: per-column event times of the last strict factorisation, [N][8] (vbk_strict_factor.cuh)
    void read_trace(long long* out);

    int num_sms() const { return num_sms_; }
    int vec_grid(long long n) const;

private:
    void require_device(const char* what) const;
    void read_scalars();   // D2H of scalar/bit/counter blocks + stream sync

    int device_, mode_;
    bool debug_ = false;   // $VBK_DEBUG: trace refinement passes on stderr
    int num_sms_ = 1;
    int smem_optin_ = 48 << 10;   // largest dynamic shared memory a CTA may opt in to on this device
    cudaStream_t stream_ = 0;
    bool analyzed_ = false;
    Symbolic sym_;

    // matrix, row-wise gather form (built on the host by stable counting sort)
    DevArray<int> gA_ptr_, gA_idx_, gAt_ptr_, gAt_idx_;
    DevArray<double> gA_val_, gAt_val_;
    // raw CSC values + scatter maps for the assemble kernel
    DevArray<double> A_val_, At_val_;
    DevArray<int> mapA_, mapAt_;
    // symbolic
    DevArray<int> iperm_, perm_, kL_, iL_, parent_, nchild_, rowptr_, rk_sig_, rj_sig_, rk_asc_, rj_asc_;
    // numeric
    DevArray<double> L_, diag_;
    DevArray<int> mark_, counters_;
    DevArray<double> scal_;
    DevArray<unsigned long long> bits_;
    // solve work vectors
    DevArray<double> z_, xk_, yk_, r_, s_;
    DevArray<double> z2_, xk2_, yk2_, r2_, s2_;       // second right-hand side (solve2), allocated on first use
    DevArray<double> h_c2_, h_b2_;
    // host-seam staging
    DevArray<double> h_dn_, h_dm_, h_c_, h_b_;
    // pinned readback
    unsigned long long* pin_bits_ = nullptr;
    double* pin_scal_ = nullptr;
    int* pin_cnt_ = nullptr;

    // slice tasks of the numeric factorisation (vbk_symbolic.h) and the strict factor kernel's hand-off state
    // (vbk_strict_factor.cuh): producer/consumer pipeline per slice task
    DevArray<int> task_col_, task_blk_, task_pos0_, task_cnt_, col_task0_, col_ntask_, winptr_;
    double schur_mean_tail_ = 0.0;             // fast mode: mean tail length of the Schur assembly's contributors
    DevArray<unsigned> pipe_masks_;            // presence masks of the (task, contributor) pairs (k_pipe_masks)
    DevArray<long long> task_pair0_;           // [ntasks] first pair of a task
    void fill_pipe_args(struct PipeArgs& pa, int ntasks);
    DevArray<int> col_pub_, col_done_, done_;
    DevArray<double> task_max_;
    DevArray<long long> trace_;            // $VBK_PROF: per-column event times of the strict factor kernel
    DevArray<unsigned long long> prof_;    // $VBK_PROF: per-phase cycle counters of the strict factor kernel
    int pipe_grid_ = 1, pipe_warps_ = 8, pipe_stages_ = 8, pipe_cap_ = 32;
    size_t pipe_smem_ = 0;
    void launch_factor_pipe(int ntasks, bool timed);

    // fast mode (vbk_kkt_fast.cu): dense scratch for the trailing window
    bool fast_ready_ = false;
    DevArray<double> Sw_, P_, dvec_, wmag_, pan_d_, panel_buf_, panel_buf2_;
    DevArray<int> wmark_, pan_keep_, tri3_flags_;
    DevArray<int> sp_end_, sp_lvlcol_;          // fast sparse columns (vbk_sparse_level.cuh): sparse prefix ends, columns by level
    std::vector<int> sp_lvlptr_;
    int sp_cap_ = 2, sp_cap_heavy_ = 2;
    int sparse_tuned_ = 0;                      // 0, 1: measuring the two sparse-column paths; 2: decided
    float sparse_ms_[2] = {0.f, 0.f};           // [0] task kernel, [1] level kernels
    cudaEvent_t ev_sp0_ = nullptr, ev_sp1_ = nullptr;
    DevArray<double> tinv_, tri_racc_;          // inverted diagonal blocks + slice accumulators of the 128-row sweeps (vbk_window_solve.cuh)
    DevArray<unsigned long long> panel_prof_;
    // look-ahead: the bulk of a panel's trailing update runs on a second stream while the next panel is factorised
    cudaStream_t stream2_ = 0;
    cudaEvent_t ev_rows_[2] = {nullptr, nullptr}, ev_updb_[2] = {nullptr, nullptr};
    void prepare_fast();
    void factor_window_fast();
    void rawsolve_window_fast(FlagSolveArgs& fs, SolveArgs& sa, size_t flag_smem);
    int solve_grid_ = 1;
    cudaEvent_t ev_f0_ = nullptr, ev_f1_ = nullptr;
};

}  // namespace vbk
