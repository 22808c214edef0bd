// vbk_kkt.cu -- host orchestration of the device-resident factor object (see vbk_kkt.h).
#include "vbk_kkt.h"
#include "vbk_kernels.cuh"
#include "vbk_strict_factor.cuh"
#include "vbk_strict_solve.cuh"

#include <stdexcept>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace vbk {

namespace {

// row-wise view of a CSC matrix, entries of each row in ascending column order: the order in which
// the reference's scatter-form smx (linalg.c:62-70) accumulates into y[row]
void csc_to_rows(int nrows, int ncols, const int* kA, const int* iA, const double* A,
                 std::vector<int>& ptr, std::vector<int>& idx, std::vector<double>& val)
{
    const int nz = kA[ncols];
    ptr.assign(nrows + 1, 0);
    for (int k = 0; k < nz; ++k) ptr[iA[k] + 1]++;
    for (int r = 0; r < nrows; ++r) ptr[r + 1] += ptr[r];
    idx.assign(nz, 0);
    val.assign(nz, 0.0);
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (int j = 0; j < ncols; ++j)
        for (int k = kA[j]; k < kA[j + 1]; ++k) {
            int dst = fill[iA[k]]++;
            idx[dst] = j;
            val[dst] = A[k];
        }
}

}  // namespace

Kkt::Kkt(int device, int mode) : device_(device), mode_(mode)
{
    debug_ = std::getenv("VBK_DEBUG") != nullptr;
    if (device_ < 0) return;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= device_) {
        std::fprintf(stderr, "vbkkt: no CUDA device %d available (%s); this library has no CPU path\n",
                     device_, e != cudaSuccess ? cudaGetErrorString(e) : "device count too small");
        std::exit(1);
    }
    VBK_CUDA(cudaSetDevice(device_));
    cudaDeviceProp prop;
    VBK_CUDA(cudaGetDeviceProperties(&prop, device_));
    num_sms_ = prop.multiProcessorCount;
    smem_optin_ = (int)prop.sharedMemPerBlockOptin;
    VBK_CUDA(cudaStreamCreate(&stream_));
    VBK_CUDA(cudaMallocHost((void**)&pin_bits_, sizeof(unsigned long long) * S_COUNT));
    VBK_CUDA(cudaMallocHost((void**)&pin_scal_, sizeof(double) * S_COUNT));
    VBK_CUDA(cudaMallocHost((void**)&pin_cnt_, sizeof(int) * C_COUNT));
    VBK_CUDA(cudaEventCreate(&ev_f0_));
    VBK_CUDA(cudaEventCreate(&ev_f1_));
}

Kkt::~Kkt()
{
    if (device_ < 0) return;
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    if (pin_bits_) cudaFreeHost(pin_bits_);
    if (pin_scal_) cudaFreeHost(pin_scal_);
    if (pin_cnt_) cudaFreeHost(pin_cnt_);
    if (ev_f0_) cudaEventDestroy(ev_f0_);
    if (ev_f1_) cudaEventDestroy(ev_f1_);
    if (ev_sp0_) cudaEventDestroy(ev_sp0_);
    if (ev_sp1_) cudaEventDestroy(ev_sp1_);
    for (int u = 0; u < 2; ++u) {
        if (ev_rows_[u]) cudaEventDestroy(ev_rows_[u]);
        if (ev_updb_[u]) cudaEventDestroy(ev_updb_[u]);
    }
#ifndef VBK_EMU
    if (stream2_) cudaStreamDestroy(stream2_);
    if (stream_) cudaStreamDestroy(stream_);
#endif
}

void Kkt::require_device(const char* what) const
{
    if (device_ < 0) {
        std::fprintf(stderr, "vbkkt: %s needs a CUDA device; this handle was created for host analysis only\n", what);
        std::exit(1);
    }
    if (!analyzed_) {
        std::fprintf(stderr, "vbkkt: %s called before analyze()\n", what);
        std::exit(1);
    }
}

int Kkt::vec_grid(long long n) const
{
    long long g = (n + kVecThreads - 1) / kVecThreads;
    long long cap = (long long)num_sms_ * 8;
    if (g < 1) g = 1;
    return (int)std::min(g, cap);
}

void Kkt::analyze(int m, int n, const int* kA, const int* iA, const double* A,
                  const int* kAt, const int* iAt, const double* At)
{
    sym_.analyze(m, n, kA, iA, kAt, iAt);
    analyzed_ = true;
    if (device_ < 0) return;
    VBK_CUDA(cudaSetDevice(device_));
    const int N = sym_.N, nz = sym_.nzA, lnz = sym_.lnz();

    // matrix: gather forms built from each CSC operand independently
    std::vector<int> ptr, idx;
    std::vector<double> val;
    csc_to_rows(m, n, kA, iA, A, ptr, idx, val);       // rows of A  (y[m] = A x[n])
    gA_ptr_.upload(ptr, stream_); gA_idx_.upload(idx, stream_); gA_val_.upload(val, stream_);
    VBK_CUDA(cudaStreamSynchronize(stream_));
    csc_to_rows(n, m, kAt, iAt, At, ptr, idx, val);    // rows of At (y[n] = At x[m])
    gAt_ptr_.upload(ptr, stream_); gAt_idx_.upload(idx, stream_); gAt_val_.upload(val, stream_);
    A_val_.upload(A, nz, stream_);
    At_val_.upload(At, nz, stream_);
    mapA_.upload(sym_.mapA, stream_);
    mapAt_.upload(sym_.mapAt, stream_);

    iperm_.upload(sym_.iperm, stream_); perm_.upload(sym_.perm, stream_);
    kL_.upload(sym_.kL, stream_); iL_.upload(sym_.iL, stream_);
    parent_.upload(sym_.parent, stream_); nchild_.upload(sym_.nchild, stream_);
    rowptr_.upload(sym_.rowptr, stream_);
    rk_sig_.upload(sym_.rk_sig, stream_); rj_sig_.upload(sym_.rj_sig, stream_);
    rk_asc_.upload(sym_.rk_asc, stream_); rj_asc_.upload(sym_.rj_asc, stream_);

    L_.alloc((size_t)lnz + 160);        // padding: the strict factor kernel's unconditional row loads may run past the last column
    diag_.alloc(N); mark_.alloc(N);
    counters_.alloc(C_COUNT); scal_.alloc(S_COUNT); bits_.alloc(S_COUNT);
    z_.alloc(N); xk_.alloc(n); yk_.alloc(m); r_.alloc(m); s_.alloc(n);
    h_dn_.alloc(n); h_dm_.alloc(m); h_c_.alloc(n); h_b_.alloc(m);

    double scal0[S_COUNT] = {0};
    scal0[S_EPSDIAG] = 1.0e-14;                          // _EPSDIAG, ldlt.c:31,215
    VBK_CUDA(cudaMemcpyAsync(scal_.p, scal0, sizeof(scal0), cudaMemcpyHostToDevice, stream_));
    VBK_CUDA(cudaMemsetAsync(bits_.p, 0, sizeof(unsigned long long) * S_COUNT, stream_));
    VBK_CUDA(cudaMemsetAsync(counters_.p, 0, sizeof(int) * C_COUNT, stream_));

    solve_grid_ = (int)std::max<long long>(1, std::min<long long>((long long)num_sms_ * 8, ((long long)N + 3) / 4));
#ifdef VBK_EMU
    solve_grid_ = std::min(solve_grid_, 3);
#endif

    // slice tasks of the numeric factorisation (vbk_symbolic.h)
    {
        const int ntasks = sym_.ntasks();
        task_col_.upload(sym_.task_col, stream_); task_blk_.upload(sym_.task_blk, stream_);
        task_pos0_.upload(sym_.task_pos0, stream_); task_cnt_.upload(sym_.task_cnt, stream_);
        col_task0_.upload(sym_.col_task0, stream_); col_ntask_.upload(sym_.col_ntask, stream_);
        if (!sym_.winptr.empty()) winptr_.upload(sym_.winptr, stream_); else winptr_.alloc(1);
        done_.alloc(N); task_max_.alloc(std::max(ntasks, 1));
    }
    // strict factor kernel (vbk_strict_factor.cuh): launch geometry
    if (!sym_.tasks_ok && mode_ != kFast) {
        std::fprintf(stderr, "vbkkt: this LP (N = %d) is too large for the strict factor kernel's per-column block table; "
                             "use fast mode (VBK_MODE=fast)\n", N);
        std::exit(1);
    }
    {
        const int ntasks = sym_.ntasks();
        int max_cnt = 1;
        for (int t = 0; t < ntasks; ++t) max_cnt = std::max(max_cnt, sym_.task_cnt[t]);
        pipe_cap_ = max_cnt <= 32 ? 32 : (max_cnt <= 64 ? 64 : 128);
        if (max_cnt > 32 * kPipeMaxChains) {
            std::fprintf(stderr, "vbkkt: a slice task of %d rows exceeds the factor kernel's %d; raise the row-block limit\n",
                         max_cnt, 32 * kPipeMaxChains);
            std::exit(1);
        }
        pipe_warps_ = kPipeWarpsDefault;
#ifndef VBK_EMU
        if (const char* e = std::getenv("VBK_PIPE_WARPS")) pipe_warps_ = std::max(3, std::min(kPipeWarpsDefault, std::atoi(e)));
#endif
        pipe_stages_ = kPipeStagesMax;        // as many as fit in shared memory (25 for 32-row tasks)
        if (const char* e = std::getenv("VBK_PIPE_STAGES")) pipe_stages_ = std::max(2, std::min(kPipeStagesMax, std::atoi(e)));
        while (pipe_stages_ > 2 && pipe_smem_bytes(pipe_cap_, pipe_stages_, sym_.rowblk, pipe_warps_) > (size_t)smem_optin_) --pipe_stages_;
        pipe_smem_ = pipe_smem_bytes(pipe_cap_, pipe_stages_, sym_.rowblk, pipe_warps_);
        if (pipe_smem_ > (size_t)smem_optin_) {
            std::fprintf(stderr, "vbkkt: the strict factor kernel needs %zu bytes of shared memory, the device offers %d\n", pipe_smem_, smem_optin_);
            std::exit(1);
        }
        {
            col_pub_.alloc(N); col_done_.alloc(N);
            int occ3 = 1;
#ifndef VBK_EMU
            VBK_CUDA(cudaFuncSetAttribute(k_factor_pipe<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_));
            VBK_CUDA(cudaFuncSetAttribute(k_factor_pipe<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_));
            VBK_CUDA(cudaFuncSetAttribute(k_factor_pipe<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_));
            VBK_CUDA(cudaFuncSetAttribute(k_factor_pipe<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_));
            VBK_CUDA(cudaFuncSetAttribute(k_factor_pipe<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_));
            VBK_CUDA(cudaFuncSetAttribute(k_factor_pipe<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_));
#endif
            if (pipe_cap_ == 32) VBK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, (k_factor_pipe<1, false>), pipe_warps_ * 32, pipe_smem_));
            else if (pipe_cap_ == 64) VBK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, (k_factor_pipe<2, false>), pipe_warps_ * 32, pipe_smem_));
            else VBK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, (k_factor_pipe<4, false>), pipe_warps_ * 32, pipe_smem_));
            occ3 = std::max(1, occ3);
            if (const char* e = std::getenv("VBK_PIPE_OCC")) occ3 = std::max(1, std::min(occ3, std::atoi(e)));
            pipe_grid_ = (int)std::max<long long>(1, std::min<long long>((long long)num_sms_ * occ3, ntasks));
#ifdef VBK_EMU
            pipe_grid_ = std::max(1, std::min(3, ntasks));
#endif
        }
    }
    // presence masks of the (task, contributor) pairs (vbk_strict_factor.cuh: k_pipe_masks), once per analysis
    if (sym_.tasks_ok && sym_.ntasks() > 0) {
        const int ntasks = sym_.ntasks();
        std::vector<long long> pair0((size_t)ntasks);
        long long pairs = 0;
        for (int t = 0; t < ntasks; ++t) {
            const int i = sym_.task_col[t];
            pair0[t] = pairs;
            pairs += (long long)((sym_.rowptr[i + 1] - sym_.rowptr[i] + kPipeQ - 1) / kPipeQ) * kPipeQ;
        }
        const int nch = pipe_cap_ / 32;
        task_pair0_.upload(pair0, stream_);
        pipe_masks_.alloc((size_t)std::max<long long>(pairs, 1) * nch);
        PipeArgs pa;
        fill_pipe_args(pa, ntasks);
        int g = std::max(1, std::min(ntasks, num_sms_ * 8));
#ifdef VBK_EMU
        g = std::min(g, 3);
        const int mthreads = 64;
#else
        const int mthreads = 256;
#endif
        if (nch == 1) VBK_LAUNCH(k_pipe_masks<1>, g, mthreads, 0, stream_, pa, pipe_masks_.p);
        else if (nch == 2) VBK_LAUNCH(k_pipe_masks<2>, g, mthreads, 0, stream_, pa, pipe_masks_.p);
        else VBK_LAUNCH(k_pipe_masks<4>, g, mthreads, 0, stream_, pa, pipe_masks_.p);
        if (debug_) std::fprintf(stderr, "vbk factor: %lld (task, contributor) pairs, %.1f MB of presence masks\n", pairs, pairs * nch * 4 / 1e6);
    } else {
        task_pair0_.alloc(1); pipe_masks_.alloc(1);
    }
    if (mode_ == kFast) prepare_fast();
    VBK_CUDA(cudaStreamSynchronize(stream_));
}

void Kkt::fill_pipe_args(PipeArgs& pa, int ntasks)
{
    const int N = sym_.N;
    pa.N = N; pa.n_ld = sym_.n; pa.ntasks = ntasks; pa.nstages = pipe_stages_;
    pa.kL = kL_.p; pa.iL = iL_.p; pa.L = L_.p; pa.diag = diag_.p; pa.mark = mark_.p;
    pa.rowptr = rowptr_.p; pa.rk = rk_sig_.p; pa.rj = rj_sig_.p; pa.perm = perm_.p;
    pa.task_col = task_col_.p; pa.task_blk = task_blk_.p; pa.task_pos0 = task_pos0_.p; pa.task_cnt = task_cnt_.p;
    pa.col_task0 = col_task0_.p; pa.col_ntask = col_ntask_.p;
    pa.winptr = winptr_.p; pa.nblk = sym_.nblk; pa.rowblk = sym_.rowblk; pa.slice_row0 = sym_.slice_row0;
    pa.masks = pipe_masks_.p; pa.task_pair0 = task_pair0_.p;
    pa.col_pub = col_pub_.p; pa.col_done = col_done_.p; pa.task_max = task_max_.p;
    pa.counters = counters_.p; pa.scal_bits = bits_.p; pa.epsnum = 0.0;        // _EPSNUM, ldlt.c:29
    { const char* e = std::getenv("VBK_PIPE_BACKOFF"); pa.backoff_ns = e ? (unsigned)std::atoi(e) : 256u; }
    pa.prof = nullptr; pa.trace = nullptr;
}

void Kkt::launch_factor_pipe(int ntasks, bool timed)
{
    const int N = sym_.N;
    PipeArgs pa;
    fill_pipe_args(pa, ntasks);
    if (std::getenv("VBK_PROF") && !prof_.p) { prof_.alloc(16); VBK_CUDA(cudaMemsetAsync(prof_.p, 0, 128, stream_)); }
    pa.prof = prof_.p;
    if (pa.prof && !trace_.p) trace_.alloc((size_t)N * 8);
    pa.trace = trace_.p;
    if (debug_) std::fprintf(stderr, "vbk factor: pipe kernel, grid %d, %d warps, %d stages, cap %d, smem %zu, %d tasks\n",
                             pipe_grid_, pipe_warps_, pipe_stages_, pipe_cap_, pipe_smem_, pa.ntasks);
    VBK_LAUNCH(k_pipe_reset, vec_grid(N), kVecThreads, 0, stream_, N, col_pub_.p, col_done_.p, counters_.p);
    const int grid = std::max(1, std::min(pipe_grid_, ntasks));
    if (timed) VBK_CUDA(cudaEventRecord(ev_f0_, stream_));
    if (pa.prof) {
        if (pipe_cap_ == 32) VBK_LAUNCH((k_factor_pipe<1, true>), grid, pipe_warps_ * 32, pipe_smem_, stream_, pa);
        else if (pipe_cap_ == 64) VBK_LAUNCH((k_factor_pipe<2, true>), grid, pipe_warps_ * 32, pipe_smem_, stream_, pa);
        else VBK_LAUNCH((k_factor_pipe<4, true>), grid, pipe_warps_ * 32, pipe_smem_, stream_, pa);
    } else {
        if (pipe_cap_ == 32) VBK_LAUNCH((k_factor_pipe<1, false>), grid, pipe_warps_ * 32, pipe_smem_, stream_, pa);
        else if (pipe_cap_ == 64) VBK_LAUNCH((k_factor_pipe<2, false>), grid, pipe_warps_ * 32, pipe_smem_, stream_, pa);
        else VBK_LAUNCH((k_factor_pipe<4, false>), grid, pipe_warps_ * 32, pipe_smem_, stream_, pa);
    }
    if (timed) VBK_CUDA(cudaEventRecord(ev_f1_, stream_));
}

void Kkt::read_scalars()
{
    bits_.download(pin_bits_, S_COUNT, stream_);
    scal_.download(pin_scal_, S_COUNT, stream_);
    counters_.download(pin_cnt_, C_COUNT, stream_);
    VBK_CUDA(cudaStreamSynchronize(stream_));
}

void Kkt::read_trace(long long* out)
{
    if (!trace_.p) return;
    trace_.download(out, (size_t)sym_.N * 8, stream_);
    VBK_CUDA(cudaStreamSynchronize(stream_));
}

float Kkt::last_factor_kernel_ms()
{
    require_device("last_factor_kernel_ms");
    float ms = 0.f;
    VBK_CUDA(cudaEventSynchronize(ev_f1_));
    VBK_CUDA(cudaEventElapsedTime(&ms, ev_f0_, ev_f1_));
    return ms;
}

void Kkt::read_phase_profile(unsigned long long out[16])
{
    for (int u = 0; u < 16; ++u) out[u] = 0;
    if (!prof_.p) return;
    prof_.download(out, 16, stream_);
    VBK_CUDA(cudaStreamSynchronize(stream_));
    VBK_CUDA(cudaMemsetAsync(prof_.p, 0, 128, stream_));
}

double Kkt::epsdiag() { require_device("epsdiag"); read_scalars(); return pin_scal_[S_EPSDIAG]; }
int Kkt::ndep() { require_device("ndep"); read_scalars(); return pin_cnt_[C_NDEP]; }

void Kkt::download_factor(double* L, double* diag, int* mark)
{
    require_device("download_factor");
    if (L) L_.download(L, sym_.lnz(), stream_);
    if (diag) diag_.download(diag, sym_.N, stream_);
    if (mark) mark_.download(mark, sym_.N, stream_);
    VBK_CUDA(cudaStreamSynchronize(stream_));
}

// ------------------------------------------------------------------------------------------------
void Kkt::factor_dev(const double* d_dn, const double* d_dm)
{
    require_device("factor");
    VBK_CUDA(cudaSetDevice(device_));
    const int N = sym_.N, n = sym_.n, m = sym_.m, nz = sym_.nzA, lnz = sym_.lnz();
    stats.factor_calls++;

    // K1 assemble (inv_num, ldlt.c:235-269,280)
    VBK_LAUNCH(k_zero_bits, 1, 32, 0, stream_, bits_.p, (int)S_MAXDIAG, 2);      // MAXDIAG, MINDIAG
    VBK_LAUNCH(k_set_diag, vec_grid(N), kVecThreads, 0, stream_, n, m, iperm_.p, d_dn, d_dm,
               scal_.p, diag_.p, mark_.p, bits_.p);
    VBK_CUDA(cudaMemsetAsync(L_.p, 0, sizeof(double) * (size_t)lnz, stream_));
    VBK_LAUNCH(k_scatter, vec_grid(nz), kVecThreads, 0, stream_, nz, mapA_.p, A_val_.p, L_.p);
    VBK_LAUNCH(k_scatter, vec_grid(nz), kVecThreads, 0, stream_, nz, mapAt_.p, At_val_.p, L_.p);
    // K2/K4 numeric LDL^T (lltnum, ldlt.c:565-631): pipelined slice tasks (vbk_strict_factor.cuh); fast mode keeps
    // them for the sparse columns only and factorises the trailing window densely (vbk_kkt_fast.cu)
    if (mode_ == kFast && sym_.dense_start < N && fast_ready_) {
        VBK_CUDA(cudaEventRecord(ev_f0_, stream_));
        factor_window_fast();
        VBK_CUDA(cudaEventRecord(ev_f1_, stream_));
    } else {
        launch_factor_pipe(sym_.ntasks(), true);
    }
    // epsdiag escalation (ldlt.c:293-306)
    VBK_LAUNCH(k_min_absdiag, vec_grid(N), kVecThreads, 0, stream_, N, diag_.p, bits_.p);
    VBK_LAUNCH(k_update_epsdiag, 1, 32, 0, stream_, scal_.p, bits_.p);
    VBK_CHECK_LAUNCH();
    stats.kernel_launches += 9;
}

void Kkt::rawsolve_dev() { rawsolve_rhs(1); }

// mask 1: z_ (right-hand side 0); 2: z2_ (right-hand side 1 on its own); 3: both in one pair of sweeps
void Kkt::rawsolve_rhs(int mask)
{
    require_device("rawsolve");
    const int N = sym_.N;
    stats.rawsolve_calls += (mask == 3) ? 2 : 1;
    if (mask != 1 && !z2_.p) throw std::runtime_error("vbkkt: rawsolve on the second right-hand side before solve2");
    double* const zsingle = (mask == 2) ? z2_.p : z_.p;
    const int rhs_single = (mask == 2) ? 1 : 0;
    SolveArgs sa;
    sa.N = N; sa.m_ld = sym_.m;
    sa.kL = kL_.p; sa.iL = iL_.p; sa.L = L_.p; sa.diag = diag_.p; sa.mark = mark_.p;
    sa.rowptr = rowptr_.p; sa.rk = rk_asc_.p; sa.rj = rj_asc_.p;
    sa.parent = parent_.p; sa.z = zsingle; sa.counters = counters_.p;
    sa.scal_bits = bits_.p; sa.epssol = 1.0e-6;            // _EPSSOL, ldlt.c:28
    sa.rhs = rhs_single;

    // eps = epssol*maxv(z,m) is only used when the factorisation met dependent pivots (ldlt.c:446)
    for (int r = 0; r < 2; ++r) {
        if (!(mask & (1 << r))) continue;
        VBK_LAUNCH(k_zero_bits, 1, 32, 0, stream_, bits_.p, (int)S_ZMAX + kRhsSlotStride * r, 1);
        VBK_LAUNCH(k_absmax, vec_grid(sym_.m), kVecThreads, 0, stream_, sym_.m, r ? z2_.p : z_.p, bits_.p + S_ZMAX + kRhsSlotStride * r);
    }
    FlagSolveArgs fs;
    fs.N = N; fs.kL = kL_.p; fs.iL = iL_.p; fs.L = L_.p; fs.diag = diag_.p; fs.mark = mark_.p;
    fs.rowptr = rowptr_.p; fs.rk = rk_asc_.p; fs.rj = rj_asc_.p; fs.parent = parent_.p;
    fs.z = zsingle; fs.done = done_.p; fs.counters = counters_.p; fs.scal_bits = bits_.p; fs.epssol = 1.0e-6;
    fs.nclaim = N;
    fs.fast = 0;
    fs.rhs = rhs_single;
    if (mask == 3) { fs.z = z_.p; fs.z2 = z2_.p; }
    const size_t sm = (size_t)(kSolveThreads / 32) * 128 * sizeof(double);
    if (mode_ == kFast && sym_.dense_start < N && fast_ready_) {
        if (mask != 1) throw std::runtime_error("vbkkt: the fast-mode window sweeps take one right-hand side");
        rawsolve_window_fast(fs, sa, sm);     // flag kernels below the window, dense sweeps on it
        return;
    }
    // forward: per-column completion flags (vbk_flag_solve.cuh); backward: producer/consumer pipeline per column
    // (resets both consistency flags: the one of a right-hand side that is not in this call was read by the host already)
    VBK_LAUNCH(k_flags_reset, vec_grid(N), kVecThreads, 0, stream_, N, done_.p, counters_.p, 1);
    if (mask == 3) VBK_LAUNCH(k_fwd_flags<2>, solve_grid_, kSolveThreads, sm, stream_, fs);
    else           VBK_LAUNCH(k_fwd_flags<1>, solve_grid_, kSolveThreads, sm, stream_, fs);
    if (mask == 3) {
        SolveArgs sb = sa;
        sa.z = z_.p; sa.rhs = 0;
        sb.z = z2_.p; sb.rhs = 1;
        VBK_LAUNCH(k_diag_strict, vec_grid(N), kVecThreads, 0, stream_, sa);
        VBK_LAUNCH(k_diag_strict, vec_grid(N), kVecThreads, 0, stream_, sb);
    } else {
        VBK_LAUNCH(k_diag_strict, vec_grid(N), kVecThreads, 0, stream_, sa);
    }
    VBK_LAUNCH(k_flags_reset, vec_grid(N), kVecThreads, 0, stream_, N, done_.p, counters_.p, 0);
    {
        BwdPipeArgs ba;
        ba.N = N; ba.nclaim = N; ba.kL = kL_.p; ba.iL = iL_.p; ba.L = L_.p; ba.mark = mark_.p; ba.z = fs.z;
        ba.z2 = fs.z2; ba.rhs = rhs_single;
        ba.done = done_.p; ba.counters = counters_.p; ba.scal_bits = bits_.p; ba.epssol = 1.0e-6;
        int g = (int)std::max<long long>(1, std::min<long long>((long long)num_sms_ * 6, N));
#ifdef VBK_EMU
        g = std::min(g, 3);
#endif
        if (mask == 3) VBK_LAUNCH(k_bwd_pipe<2>, g, kBwdWarps * 32, bwd_pipe_smem_bytes(2), stream_, ba);
        else           VBK_LAUNCH(k_bwd_pipe<1>, g, kBwdWarps * 32, bwd_pipe_smem_bytes(1), stream_, ba);
    }
    VBK_CHECK_LAUNCH();
    stats.kernel_launches += (mask == 3) ? 10 : 7;
}

void Kkt::spmv_A(const double* d_x, double* d_y)
{
    require_device("spmv_A");
    VBK_LAUNCH(k_spmv_rows, vec_grid(sym_.m), kVecThreads, 0, stream_, sym_.m, gA_ptr_.p, gA_idx_.p, gA_val_.p, d_x, d_y);
    stats.kernel_launches++;
}
void Kkt::spmv_At(const double* d_x, double* d_y)
{
    require_device("spmv_At");
    VBK_LAUNCH(k_spmv_rows, vec_grid(sym_.n), kVecThreads, 0, stream_, sym_.n, gAt_ptr_.p, gAt_idx_.p, gAt_val_.p, d_x, d_y);
    stats.kernel_launches++;
}

int Kkt::solve_dev(const double* d_Dn, const double* d_Dm, double* d_c, double* d_b)
{
    double* c[2] = {d_c, nullptr};
    double* b[2] = {d_b, nullptr};
    int cons[2] = {1, 1};
    solve_rhs(1, d_Dn, d_Dm, c, b, cons);
    return cons[0];
}

// The two systems of one hsd iteration (hsd.c:223 and :228) share the factor and do not depend on each other: both
// right-hand sides go through ONE pair of sweeps per refinement pass (the sweeps are bound by their dependency chain,
// so the second right-hand side is almost free).  Each right-hand side keeps the reference's own refinement state
// (ldlt.c:367-416) and arithmetic; when only one of them needs another pass it runs alone.
void Kkt::solve2_dev(const double* d_Dn, const double* d_Dm, double* d_c0, double* d_b0, double* d_c1, double* d_b1,
                     int consistent[2])
{
    double* c[2] = {d_c0, d_c1};
    double* b[2] = {d_b0, d_b1};
    if (mode_ == kFast && sym_.dense_start < sym_.N && fast_ready_) {     // fast-mode window sweeps: one at a time
        consistent[0] = solve_dev(d_Dn, d_Dm, d_c0, d_b0);
        const int p0 = stats.last_passes;
        consistent[1] = solve_dev(d_Dn, d_Dm, d_c1, d_b1);
        stats.last_passes2[0] = p0; stats.last_passes2[1] = stats.last_passes;
        return;
    }
    solve_rhs(3, d_Dn, d_Dm, c, b, consistent);
}

void Kkt::solve_rhs(int mask, const double* d_Dn, const double* d_Dm, double* const d_c[2], double* const d_b[2], int consistent[2])
{
    require_device("solve");
    VBK_CUDA(cudaSetDevice(device_));
    const int n = sym_.n, m = sym_.m, N = sym_.N;
    if (mask & 2) {
        if (!z2_.p) { z2_.alloc(N); xk2_.alloc(n); yk2_.alloc(m); r2_.alloc(m); s2_.alloc(n); }
    }
    double* const zz[2] = {z_.p, z2_.p};
    double* const xk[2] = {xk_.p, xk2_.p};
    double* const yk[2] = {yk_.p, yk2_.p};
    double* const rr[2] = {r_.p, r2_.p};
    double* const ss[2] = {s_.p, s2_.p};
    int pass[2] = {0, 0};
    double maxrs[2] = {HUGE_VAL, HUGE_VAL}, oldmaxrs[2] = {HUGE_VAL, HUGE_VAL}, maxbc[2] = {1.0, 1.0};
    int active = mask;

    for (int q = 0; q < 2; ++q) {
        if (!(mask & (1 << q))) continue;
        stats.solve_calls++;
        const int o = kRhsSlotStride * q;
        VBK_LAUNCH(k_zero_bits, 1, 32, 0, stream_, bits_.p, (int)S_MAXBC_B + o, 2);
        VBK_LAUNCH(k_absmax, vec_grid(m), kVecThreads, 0, stream_, m, d_b[q], bits_.p + S_MAXBC_B + o);
        VBK_LAUNCH(k_absmax, vec_grid(n), kVecThreads, 0, stream_, n, d_c[q], bits_.p + S_MAXBC_C + o);
        stats.kernel_launches += 3;
    }
    while (active) {
        for (int q = 0; q < 2; ++q) {
            if (!(active & (1 << q))) continue;
            if (pass[q] == 0) VBK_LAUNCH(k_permute_in, vec_grid(N), kVecThreads, 0, stream_, n, m, iperm_.p, d_c[q], d_b[q], zz[q]);
            else              VBK_LAUNCH(k_permute_in, vec_grid(N), kVecThreads, 0, stream_, n, m, iperm_.p, ss[q], rr[q], zz[q]);
        }
        rawsolve_rhs(active);
        for (int q = 0; q < 2; ++q) {
            if (!(active & (1 << q))) continue;
            const int o = kRhsSlotStride * q;
            VBK_LAUNCH(k_permute_out, vec_grid(N), kVecThreads, 0, stream_, n, m, pass[q] == 0 ? 0 : 1, iperm_.p, zz[q], xk[q], yk[q]);
            VBK_LAUNCH(k_zero_bits, 1, 32, 0, stream_, bits_.p, (int)S_MAXR + o, 2);
            // r = b - (A x_k + Dm y_k) ; s = c - (At y_k - Dn x_k)      (ldlt.c:389-398)
            VBK_LAUNCH(k_residual, vec_grid(m), kVecThreads, 0, stream_, m, 1, gA_ptr_.p, gA_idx_.p, gA_val_.p,
                       xk[q], d_Dm, yk[q], d_b[q], rr[q], bits_.p + S_MAXR + o);
            VBK_LAUNCH(k_residual, vec_grid(n), kVecThreads, 0, stream_, n, 0, gAt_ptr_.p, gAt_idx_.p, gAt_val_.p,
                       yk[q], d_Dn, xk[q], d_c[q], ss[q], bits_.p + S_MAXS + o);
            stats.kernel_launches += 5;
        }
        read_scalars();
        int next = 0;
        for (int q = 0; q < 2; ++q) {
            if (!(active & (1 << q))) continue;
            const int o = kRhsSlotStride * q;
            consistent[q] = pin_cnt_[C_CONSISTENT + q];
            if (pass[q] == 0) {
                double mb, mc;
                std::memcpy(&mb, &pin_bits_[S_MAXBC_B + o], 8);
                std::memcpy(&mc, &pin_bits_[S_MAXBC_C + o], 8);
                maxbc[q] = (mb > mc ? mb : mc) + 1;                                // ldlt.c:367
            }
            double mr, ms;
            std::memcpy(&mr, &pin_bits_[S_MAXR + o], 8);
            std::memcpy(&ms, &pin_bits_[S_MAXS + o], 8);
            oldmaxrs[q] = maxrs[q];
            maxrs[q] = (mr > ms ? mr : ms);                                         // ldlt.c:401
            pass[q]++;
            if (debug_) std::fprintf(stderr, "vbk solve: rhs %d pass %d maxr %.17g maxs %.17g maxbc %.17g consistent %d\n",
                                     q, pass[q], mr, ms, maxbc[q], consistent[q]);
            if (maxrs[q] > 1.0e-10 * maxbc[q] && maxrs[q] < oldmaxrs[q] / 2) next |= 1 << q;   // ldlt.c:411
        }
        active = next;
    }
    for (int q = 0; q < 2; ++q) {
        if (!(mask & (1 << q))) continue;
        if (maxrs[q] > oldmaxrs[q] && pass[q] > 1) {                                // ldlt.c:413-416
            VBK_LAUNCH(k_permute_out, vec_grid(N), kVecThreads, 0, stream_, n, m, 2, iperm_.p, zz[q], xk[q], yk[q]);
            stats.kernel_launches++;
        }
        VBK_CUDA(cudaMemcpyAsync(d_c[q], xk[q], sizeof(double) * (size_t)n, cudaMemcpyDeviceToDevice, stream_));
        VBK_CUDA(cudaMemcpyAsync(d_b[q], yk[q], sizeof(double) * (size_t)m, cudaMemcpyDeviceToDevice, stream_));
        stats.last_passes2[q] = pass[q];
    }
    VBK_CHECK_LAUNCH();
    stats.last_passes = (mask & 1) ? pass[0] : pass[1];
    stats.last_consistent = (mask & 1) ? consistent[0] : consistent[1];
    stats.last_ndep = pin_cnt_[C_NDEP];
}

void Kkt::factor_host(const double* dn, const double* dm)
{
    require_device("factor");
    VBK_CUDA(cudaSetDevice(device_));
    h_dn_.upload(dn, sym_.n, stream_);
    h_dm_.upload(dm, sym_.m, stream_);
    factor_dev(h_dn_.p, h_dm_.p);
    VBK_CUDA(cudaStreamSynchronize(stream_));
}

int Kkt::solve_host(const double* Dn, const double* Dm, double* c, double* b)
{
    require_device("solve");
    VBK_CUDA(cudaSetDevice(device_));
    h_dn_.upload(Dn, sym_.n, stream_);
    h_dm_.upload(Dm, sym_.m, stream_);
    h_c_.upload(c, sym_.n, stream_);
    h_b_.upload(b, sym_.m, stream_);
    int consistent = solve_dev(h_dn_.p, h_dm_.p, h_c_.p, h_b_.p);
    h_c_.download(c, sym_.n, stream_);
    h_b_.download(b, sym_.m, stream_);
    VBK_CUDA(cudaStreamSynchronize(stream_));
    return consistent;
}

int Kkt::solve2_host(const double* Dn, const double* Dm, double* c0, double* b0, double* c1, double* b1)
{
    require_device("solve2");
    VBK_CUDA(cudaSetDevice(device_));
    if (!h_c2_.p) { h_c2_.alloc(sym_.n); h_b2_.alloc(sym_.m); }
    h_dn_.upload(Dn, sym_.n, stream_);
    h_dm_.upload(Dm, sym_.m, stream_);
    h_c_.upload(c0, sym_.n, stream_);
    h_b_.upload(b0, sym_.m, stream_);
    h_c2_.upload(c1, sym_.n, stream_);
    h_b2_.upload(b1, sym_.m, stream_);
    int cons[2] = {1, 1};
    solve2_dev(h_dn_.p, h_dm_.p, h_c_.p, h_b_.p, h_c2_.p, h_b2_.p, cons);
    h_c_.download(c0, sym_.n, stream_);
    h_b_.download(b0, sym_.m, stream_);
    h_c2_.download(c1, sym_.n, stream_);
    h_b2_.download(b1, sym_.m, stream_);
    VBK_CUDA(cudaStreamSynchronize(stream_));
    return cons[0] | (cons[1] << 1);
}

}  // namespace vbk
