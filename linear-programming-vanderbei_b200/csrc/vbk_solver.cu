// vbk_solver.cu -- device-resident METHOD plugins (the B2 seam).
//
// Same signature, stdout and return codes as the reference's `solver` in src/ipo/hsd.c:27-311
// (homogeneous self-dual predictor/corrector) and src/ipo/intpt.c:33-261 (path following).  All
// vectors live on the GPU for the whole solve; per iteration only a handful of scalars (dot
// products, max-norms, the ratio-test maximum) cross PCIe, for the log line and the step logic.
// The scalar recurrences (mu, gamma, dphi, theta, phi, psi) run on the host in exactly the
// reference's expression order.
#include "vbk_solver.h"
#include "vbk_kernels.cuh"
#include "vbk_linalg.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace vbk {

namespace {

#define VMIN(x, y) ((x) > (y) ? (y) : (x))   // macros.h:2

struct Capture { int iter = -1; double *E, *D, *ry, *rx, *sy, *sx; } g_cap;
void capture_in(int iter, cudaStream_t st, int m, int n, const double* E, const double* D, const double* ry, const double* rx)
{
    if (iter != g_cap.iter) return;
    VBK_CUDA(cudaMemcpyAsync(g_cap.E, E, 8 * (size_t)m, cudaMemcpyDeviceToHost, st));
    VBK_CUDA(cudaMemcpyAsync(g_cap.D, D, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    VBK_CUDA(cudaMemcpyAsync(g_cap.ry, ry, 8 * (size_t)m, cudaMemcpyDeviceToHost, st));
    VBK_CUDA(cudaMemcpyAsync(g_cap.rx, rx, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    VBK_CUDA(cudaStreamSynchronize(st));
}
void capture_out(int iter, cudaStream_t st, int m, int n, const double* sy, const double* sx)
{
    if (iter != g_cap.iter) return;
    VBK_CUDA(cudaMemcpyAsync(g_cap.sy, sy, 8 * (size_t)m, cudaMemcpyDeviceToHost, st));
    VBK_CUDA(cudaMemcpyAsync(g_cap.sx, sx, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
    VBK_CUDA(cudaStreamSynchronize(st));
}

struct Timer {
    bool on;
    cudaStream_t s;
    double* acc;
    std::chrono::steady_clock::time_point t0;
    Timer(bool on_, cudaStream_t s_, double* acc_) : on(on_), s(s_), acc(acc_) {
        if (on) { cudaStreamSynchronize(s); t0 = std::chrono::steady_clock::now(); }
    }
    ~Timer() {
        if (on) {
            cudaStreamSynchronize(s);
            *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
    }
};

// everything both METHODs share: the LP on the device, the factor object, launch helpers
static thread_local long long t_last_launches = 0;   // kernels launched by this thread's last solver_* call

struct Workspace {
    int m, n, nz;
    Kkt kkt;
    LinalgContext la;
    cudaStream_t st;
    DevArray<double> x, z, c, sigma, D, dx, dz;      // length n
    DevArray<double> y, w, b, rho, E, dy, dw;        // length m
    DevArray<unsigned long long> slot;
    unsigned long long* pin_slot = nullptr;
    std::vector<int> kAt, iAt;
    std::vector<double> At;
    SolveProfile* prof;

    Workspace(int device, int mode, int m_, int n_, int nz_, const int* iA, const int* kA, const double* A,
              const double* hb, const double* hc, SolveProfile* prof_)
        : m(m_), n(n_), nz(nz_), kkt(device, mode), la(device, mode, true, kkt.stream()), st(kkt.stream()),
          prof(prof_)
    {
        // hsd.c:111 / intpt.c:108: build A^T once -- on the device, same entry order as atnum
        kAt.resize((size_t)m + 1); iAt.resize((size_t)nz > 0 ? nz : 1); At.resize((size_t)nz > 0 ? nz : 1);
        la.atnum_host(m, n, kA, iA, A, kAt.data(), iAt.data(), At.data());
        // The reference analyses K at the first ldltfac call (hsd.c:218: arguments swapped, the
        // ldlt-space "A" is the solver's A^T).  The symbolic result does not depend on the iterate,
        // so it is done up front and the SpMVs of iteration 0 already use the resident matrix.
        kkt.analyze(n, m, kAt.data(), iAt.data(), At.data(), kA, iA, A);
        x.alloc(n); z.alloc(n); c.alloc(n); sigma.alloc(n); D.alloc(n); dx.alloc(n); dz.alloc(n);
        y.alloc(m); w.alloc(m); b.alloc(m); rho.alloc(m); E.alloc(m); dy.alloc(m); dw.alloc(m);
        c.upload(hc, n, st);
        b.upload(hb, m, st);
        slot.alloc(1);
        VBK_CUDA(cudaMallocHost((void**)&pin_slot, 8));
    }
    ~Workspace() { if (pin_slot) cudaFreeHost(pin_slot); }

    int g(long long len) const { return kkt.vec_grid(len); }
    void launches(int k) { kkt.stats.kernel_launches += k; }

    // rho[m] = A x ;  sigma[n] = A^T y   (solver-space; see vbk_kkt.h for the ldlt-space naming)
    void mul_A(const double* dxv, double* out) { kkt.spmv_At(dxv, out); }
    void mul_At(const double* dyv, double* out) { kkt.spmv_A(dyv, out); }

    void fill(DevArray<double>& v, int len, double val) { VBK_LAUNCH(k_fill, g(len), kVecThreads, 0, st, len, val, v.p); launches(1); }

    double ratio_max() {   // max over (-dx/x, -dz/z, -dy/y, -dw/w) clipped at 0
        VBK_CUDA(cudaMemsetAsync(slot.p, 0, 8, st));
        VBK_LAUNCH(k_ratio_test, g(n), kVecThreads, 0, st, n, dx.p, x.p, dz.p, z.p, slot.p);
        VBK_LAUNCH(k_ratio_test, g(m), kVecThreads, 0, st, m, dy.p, y.p, dw.p, w.p, slot.p);
        launches(2);
        VBK_CUDA(cudaMemcpyAsync(pin_slot, slot.p, 8, cudaMemcpyDeviceToHost, st));
        VBK_CUDA(cudaStreamSynchronize(st));
        double v;
        std::memcpy(&v, pin_slot, 8);
        return v;
    }
    void finish(double* hx, double* hy) {
        x.download(hx, n, st);
        y.download(hy, m, st);
        VBK_CUDA(cudaStreamSynchronize(st));
        t_last_launches = kkt.stats.kernel_launches + la.launches;
        if (prof) {
            prof->factor_calls = kkt.stats.factor_calls;
            prof->solve_calls = kkt.stats.solve_calls;
            prof->rawsolve_calls = kkt.stats.rawsolve_calls;
            prof->kernel_launches = kkt.stats.kernel_launches + la.launches;
            prof->lnz = kkt.sym().lnz();
            prof->narth = kkt.sym().narth;
            prof->N = kkt.sym().N;
        }
    }
};

void show_small_problem(int m, int n, const int* kA, const int* iA, const double* A, const double* b, const double* c)
{   // hsd.c:70-95: tiny problems are echoed
    double AA[20][20];
    for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) AA[i][j] = 0;
    for (int j = 0; j < n; j++) for (int k = kA[j]; k < kA[j + 1]; k++) AA[iA[k]][j] = A[k];
    std::printf("A <= b: \n");
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < n; j++) std::printf(" %5.1f", AA[i][j]);
        std::printf("<= %5.1f \n", b[i]);
    }
    std::printf("\n");
    std::printf("c: \n");
    for (int j = 0; j < n; j++) std::printf(" %5.1f", c[j]);
    std::printf("\n");
}

}  // namespace

// hsdls.c:296-336 on the host (the (phi, psi) component) and the inverse of the device's order key
static double ls_host(double xj, double zj, double dxj, double dzj, double beta, double delta, double mu)
{
    const double a = dxj * dzj;
    const double b = zj * dxj + xj * dzj + (1 - beta) * (1 - delta) * mu;
    const double c = xj * zj - (1 - beta) * mu;
    const double d = b * b - 4 * a * c;
    if (a == 0.0) return -c / b;
    if (a > 0) {
        if (b < 0) return d >= 0 ? 2 * c / (-b + std::sqrt(d)) : HUGE_VAL;
        return HUGE_VAL;
    }
    if (b < 0) return 2 * c / (-b + std::sqrt(d));
    return (-b - std::sqrt(d)) / (2 * a);
}
static double ls_unkey(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
    std::memcpy(&v, &b, 8);
    return v;
}

static int g_itnlim = 200;
void set_iteration_limit(int itnlim) { g_itnlim = itnlim > 0 ? itnlim : 200; }

// The batch driver (vbk_batch.cu) runs many solves concurrently, one host thread + one stream each:
// the iteration log of such a solve is suppressed (per thread), everything else is unchanged.
static thread_local bool t_quiet = false;
static thread_local int t_last_iterations = 0;
void set_thread_quiet(bool quiet) { t_quiet = quiet; }
int last_thread_iterations() { return t_last_iterations; }
long long last_thread_launches() { return t_last_launches; }
#define VBK_LOG(...) do { if (!t_quiet) std::printf(__VA_ARGS__); } while (0)
#define VBK_LOG_FLUSH() do { if (!t_quiet) std::fflush(stdout); } while (0)

void set_capture(int iter, double* E, double* D, double* rhs_y, double* rhs_x, double* sol_y, double* sol_x)
{
    g_cap.iter = iter; g_cap.E = E; g_cap.D = D; g_cap.ry = rhs_y; g_cap.rx = rhs_x; g_cap.sy = sol_y; g_cap.sx = sol_x;
}

// METHOD = hsd (reference src/ipo/hsd.c) and, with longstep, METHOD = hsdls (src/ipo/hsdls.c): the same homogeneous
// self-dual iteration; hsdls keeps delta = 2(1 - beta) constant instead of alternating predictor / corrector, takes the
// step length from a per-component quadratic line search, runs up to 600 iterations and has its own status rules.
static int solver_hsd_impl(bool longstep, int device, int mode, int m, int n, int nz, const int* iA, const int* kA, const double* A,
                           const double* b, const double* c, double f, double* x, double* y, SolveProfile* prof)
{
    const bool timed = prof != nullptr;
    auto t_begin = std::chrono::steady_clock::now();
    Workspace W(device, mode, m, n, nz, iA, kA, A, b, c, prof);
    DevArray<double> fx, gx, fy, gy;
    fx.alloc(n); gx.alloc(n); fy.alloc(m); gy.alloc(m);
    cudaStream_t st = W.st;
    int status = 5;

    if (!longstep && m < 20 && n < 20 && !t_quiet) show_small_problem(m, n, kA, iA, A, b, c);   // hsd.c:70-95 only

    W.fill(W.x, n, 1.0); W.fill(W.z, n, 1.0); W.fill(W.w, m, 1.0); W.fill(W.y, m, 1.0);
    double phi = 1.0, psi = 1.0;

    VBK_LOG("m = %d,n = %d,nz = %d\n", m, n, nz);
    VBK_LOG(
"--------------------------------------------------------------------------\n"
"         |           Primal          |            Dual           |       |\n"
"  Iter   |  Obj Value       Infeas   |  Obj Value       Infeas   |  mu   |\n"
"- - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - \n");
    VBK_LOG_FLUSH();
    if (timed) { cudaStreamSynchronize(st); prof->setup_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count(); }

    const double beta = 0.80, delta_ls = 2 * (1 - beta);                          // hsdls.c:112-113
    const int itnlim = longstep ? (g_itnlim == 200 ? 600 : g_itnlim) : g_itnlim; // MAX_ITER: hsd.c:25, hsdls.c:25
    DevArray<double> lsvals;
    DevArray<int> lsnan;
    if (longstep) { lsvals.alloc((size_t)n + m); lsnan.alloc(1); }
    int iter;
    for (iter = 0; iter < itnlim; iter++) {
        double d[4];
        {
            DotJob jobs[4] = {{W.z.p, W.x.p, n}, {W.w.p, W.y.p, m}, {W.c.p, W.x.p, n}, {W.b.p, W.y.p, m}};
            W.la.dots_dev(jobs, 4, d);
        }
        const double mu = (d[0] + d[1] + phi * psi) / (n + m + 1);              // hsd.c:136
        const double delta = longstep ? delta_ls : ((iter % 2 == 0) ? 0.0 : 1.0);
        const double primal_obj = d[2], dual_obj = d[3];

        if (mu < 1.0e-12) {                                                     // hsd.c:155-176, hsdls.c:134-153
            if (longstep ? (phi > 1.0e-12) : (phi > psi)) { status = 0; break; }
            else if (dual_obj < 0.0) { status = 2; break; }
            else if (primal_obj > 0.0) { status = 4; break; }
            else if (longstep) { status = 7; break; }
            else { VBK_LOG("Trouble in river city \n"); status = 4; break; }
        }

        // infeasibilities (hsd.c:182-198)
        double nn[2];
        W.mul_A(W.x.p, W.rho.p);
        VBK_LAUNCH(k_hsd_infeas, W.g(m), kVecThreads, 0, st, m, 0, phi, W.b.p, W.w.p, W.rho.p);
        W.mul_At(W.y.p, W.sigma.p);
        VBK_LAUNCH(k_hsd_infeas, W.g(n), kVecThreads, 0, st, n, 1, phi, W.c.p, W.z.p, W.sigma.p);
        {
            DotJob jobs[2] = {{W.rho.p, W.rho.p, m}, {W.sigma.p, W.sigma.p, n}};
            W.la.dots_dev(jobs, 2, nn);
        }
        const double normr = std::sqrt(nn[0]) / phi, norms = std::sqrt(nn[1]) / phi;
        VBK_LAUNCH(k_hsd_rhs, W.g(m), kVecThreads, 0, st, m, delta, mu, W.w.p, W.y.p, W.rho.p);
        VBK_LAUNCH(k_hsd_rhs, W.g(n), kVecThreads, 0, st, n, delta, mu, W.z.p, W.x.p, W.sigma.p);
        W.launches(4);

        const double gamma = -(1 - delta) * (dual_obj - primal_obj + psi) + psi - delta * mu / phi;

        VBK_LOG("%8d   %14.7e  %8.1e    %14.7e  %8.1e  %8.1e \n",
                    iter, primal_obj / phi + f, normr, dual_obj / phi + f, norms, mu);
        VBK_LOG_FLUSH();

        // step directions (hsd.c:215-238)
        VBK_LAUNCH(k_ratio, W.g(n), kVecThreads, 0, st, n, W.z.p, W.x.p, W.D.p);
        VBK_LAUNCH(k_ratio, W.g(m), kVecThreads, 0, st, m, W.w.p, W.y.p, W.E.p);
        W.launches(2);
        {
            Timer t(timed, st, timed ? &prof->factor_s : nullptr);
            W.kkt.factor_dev(W.E.p, W.D.p);                                     // hsd.c:218
        }
        VBK_LAUNCH(k_neg_copy, W.g(n), kVecThreads, 0, st, n, W.sigma.p, fx.p);
        VBK_LAUNCH(k_copy, W.g(m), kVecThreads, 0, st, m, W.rho.p, fy.p);
        VBK_LAUNCH(k_neg_copy, W.g(n), kVecThreads, 0, st, n, W.c.p, gx.p);
        VBK_LAUNCH(k_neg_copy, W.g(m), kVecThreads, 0, st, m, W.b.p, gy.p);
        W.launches(4);
        capture_in(iter, st, m, n, W.E.p, W.D.p, fy.p, fx.p);
        {
            // the two systems of the iteration (hsd.c:223 and :228) have independent right-hand sides: one pair of
            // sweeps per refinement pass serves both
            Timer t(timed, st, timed ? &prof->solve_s : nullptr);
            int cons2[2];
            W.kkt.solve2_dev(W.E.p, W.D.p, fy.p, fx.p, gy.p, gx.p, cons2);
        }
        capture_out(iter, st, m, n, fy.p, fx.p);
        if (timed) prof->refine_passes += W.kkt.stats.last_passes2[0] + W.kkt.stats.last_passes2[1];

        {
            DotJob jobs[4] = {{W.c.p, fx.p, n}, {W.b.p, fy.p, m}, {W.c.p, gx.p, n}, {W.b.p, gy.p, m}};
            W.la.dots_dev(jobs, 4, d);
        }
        const double dphi = (d[0] - d[1] + gamma) / (d[2] - d[3] - psi / phi);  // hsd.c:230-231

        VBK_LAUNCH(k_hsd_dir, W.g(n), kVecThreads, 0, st, n, dphi, fx.p, gx.p, W.dx.p);
        VBK_LAUNCH(k_hsd_dir, W.g(m), kVecThreads, 0, st, m, dphi, fy.p, gy.p, W.dy.p);
        VBK_LAUNCH(k_comp_dir, W.g(n), kVecThreads, 0, st, n, delta * mu, W.x.p, W.z.p, W.D.p, W.dx.p, W.dz.p);
        VBK_LAUNCH(k_comp_dir, W.g(m), kVecThreads, 0, st, m, delta * mu, W.y.p, W.w.p, W.E.p, W.dy.p, W.dw.p);
        W.launches(4);
        const double dpsi = delta * mu / phi - psi - (psi / phi) * dphi;

        double theta;
        if (!longstep) {
            // step length (hsd.c:248-259)
            theta = W.ratio_max();
            if (theta < -dphi / phi) theta = -dphi / phi;
            if (theta < -dpsi / psi) theta = -dpsi / psi;
            theta = VMIN(0.95 / theta, 1.0);
        } else {
            // step length (hsdls.c:222-241): MIN-fold of the per-component line searches, see k_linesearch
            const int minus1 = -1;
            const unsigned long long top = ~0ull;
            VBK_CUDA(cudaMemcpyAsync(lsnan.p, &minus1, sizeof(int), cudaMemcpyHostToDevice, st));
            VBK_CUDA(cudaMemcpyAsync(W.slot.p, &top, 8, cudaMemcpyHostToDevice, st));
            VBK_LAUNCH(k_linesearch, W.g(n), kVecThreads, 0, st, n, 0, W.x.p, W.z.p, W.dx.p, W.dz.p, beta, delta, mu, lsvals.p, lsnan.p);
            VBK_LAUNCH(k_linesearch, W.g(m), kVecThreads, 0, st, m, n, W.y.p, W.w.p, W.dy.p, W.dw.p, beta, delta, mu, lsvals.p, lsnan.p);
            VBK_LAUNCH(k_min_after, W.g(n + m), kVecThreads, 0, st, n + m, lsvals.p, lsnan.p, W.slot.p);
            W.launches(3);
            int lastnan = -1;
            unsigned long long key = top;
            VBK_CUDA(cudaMemcpyAsync(W.pin_slot, W.slot.p, 8, cudaMemcpyDeviceToHost, st));
            VBK_CUDA(cudaMemcpyAsync(&lastnan, lsnan.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            VBK_CUDA(cudaStreamSynchronize(st));
            std::memcpy(&key, W.pin_slot, 8);
            theta = 1.0;
            if (lastnan == n + m - 1) theta = std::nan("");
            else if (key != top) {
                const double vmin = ls_unkey(key);
                theta = lastnan >= 0 ? vmin : (theta < vmin ? theta : vmin);
            }
            {   // the (phi, psi) component, hsdls.c:233, on the host in the reference's own expression shapes
                const double v = ls_host(phi, psi, dphi, dpsi, beta, delta, mu);
                theta = theta < v ? theta : v;
            }
            if (theta < 1.0) theta *= 0.9999;
        }

        VBK_LAUNCH(k_step2, W.g(n), kVecThreads, 0, st, n, theta, W.dx.p, W.dz.p, W.x.p, W.z.p);
        VBK_LAUNCH(k_step2, W.g(m), kVecThreads, 0, st, m, theta, W.dy.p, W.dw.p, W.y.p, W.w.p);
        W.launches(2);
        phi = phi + theta * dphi;
        psi = psi + theta * dpsi;
    }

    VBK_LAUNCH(k_scale2, W.g(n), kVecThreads, 0, st, n, phi, W.x.p, W.z.p);     // hsd.c:277-284
    VBK_LAUNCH(k_scale2, W.g(m), kVecThreads, 0, st, m, phi, W.y.p, W.w.p);
    W.launches(2);
    VBK_CHECK_LAUNCH();
    W.finish(x, y);
    t_last_iterations = iter;
    if (timed) {
        prof->iterations = iter;
        prof->total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    }
    return status;
}

int solver_hsd(int device, int mode, int m, int n, int nz, const int* iA, const int* kA, const double* A,
               const double* b, const double* c, double f, double* x, double* y, SolveProfile* prof)
{
    return solver_hsd_impl(false, device, mode, m, n, nz, iA, kA, A, b, c, f, x, y, prof);
}
int solver_hsdls(int device, int mode, int m, int n, int nz, const int* iA, const int* kA, const double* A,
                 const double* b, const double* c, double f, double* x, double* y, SolveProfile* prof)
{
    return solver_hsd_impl(true, device, mode, m, n, nz, iA, kA, A, b, c, f, x, y, prof);
}

int solver_intpt(int device, int mode, int m, int n, int nz, const int* iA, const int* kA, const double* A,
                 const double* b, const double* c, double f, double* x, double* y, SolveProfile* prof)
{
    const bool timed = prof != nullptr;
    auto t_begin = std::chrono::steady_clock::now();
    Workspace W(device, mode, m, n, nz, iA, kA, A, b, c, prof);
    cudaStream_t st = W.st;
    int status = 5;

    if (m < 20 && n < 20 && !t_quiet) show_small_problem(m, n, kA, iA, A, b, c);

    W.fill(W.x, n, 1000.0); W.fill(W.z, n, 1000.0); W.fill(W.w, m, 1000.0); W.fill(W.y, m, 1000.0);
    const double delta = 0.02, r = 0.9;
    double normr0 = HUGE_VAL, norms0 = HUGE_VAL;

    VBK_LOG("m = %d,n = %d,nz = %d\n", m, n, nz);
    VBK_LOG(
"------------------------------------------------------------------\n"
"         |           Primal          |            Dual           |\n"
"  Iter   |  Obj Value       Infeas   |  Obj Value       Infeas   |\n"
"- - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - \n");
    VBK_LOG_FLUSH();
    if (timed) { cudaStreamSynchronize(st); prof->setup_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count(); }

    int iter;
    for (iter = 0; iter < g_itnlim; iter++) {
        // intpt.c:139-149
        W.mul_A(W.x.p, W.rho.p);
        VBK_LAUNCH(k_pf_infeas, W.g(m), kVecThreads, 0, st, m, 0, W.b.p, W.w.p, W.rho.p);
        W.mul_At(W.y.p, W.sigma.p);
        VBK_LAUNCH(k_pf_infeas, W.g(n), kVecThreads, 0, st, n, 1, W.c.p, W.z.p, W.sigma.p);
        W.launches(2);
        double d[6];
        {
            DotJob jobs[6] = {{W.rho.p, W.rho.p, m}, {W.sigma.p, W.sigma.p, n}, {W.z.p, W.x.p, n},
                              {W.y.p, W.w.p, m}, {W.c.p, W.x.p, n}, {W.b.p, W.y.p, m}};
            W.la.dots_dev(jobs, 6, d);
        }
        const float normr = std::sqrt(d[0]);                 // float on purpose, intpt.c:47
        const float norms = std::sqrt(d[1]);
        const double gamma = d[2] + d[3];                    // intpt.c:155
        const float primal_obj = d[4] + f;
        const float dual_obj = d[5] + f;
        VBK_LOG("%8d   %14.7e  %8.1e    %14.7e  %8.1e \n", iter, primal_obj, normr, dual_obj, norms);
        VBK_LOG_FLUSH();

        if (normr < 1.0e-6 && norms < 1.0e-6 && gamma < 1.0e-6) { status = 0; break; }   // intpt.c:171-182
        if (normr > 10 * normr0) { status = 2; break; }
        if (norms > 10 * norms0) { status = 4; break; }

        const double mu = delta * gamma / (n + m);           // intpt.c:188

        VBK_LAUNCH(k_ratio, W.g(n), kVecThreads, 0, st, n, W.z.p, W.x.p, W.D.p);
        VBK_LAUNCH(k_ratio, W.g(m), kVecThreads, 0, st, m, W.w.p, W.y.p, W.E.p);
        W.launches(2);
        {
            Timer t(timed, st, timed ? &prof->factor_s : nullptr);
            W.kkt.factor_dev(W.E.p, W.D.p);                  // intpt.c:197
        }
        VBK_LAUNCH(k_pf_rhs, W.g(n), kVecThreads, 0, st, n, 1, mu, W.sigma.p, W.z.p, W.x.p, W.dx.p);
        VBK_LAUNCH(k_pf_rhs, W.g(m), kVecThreads, 0, st, m, 0, mu, W.rho.p, W.w.p, W.y.p, W.dy.p);
        W.launches(2);
        capture_in(iter, st, m, n, W.E.p, W.D.p, W.dy.p, W.dx.p);
        {
            Timer t(timed, st, timed ? &prof->solve_s : nullptr);
            W.kkt.solve_dev(W.E.p, W.D.p, W.dy.p, W.dx.p);   // intpt.c:202
        }
        capture_out(iter, st, m, n, W.dy.p, W.dx.p);
        if (timed) prof->refine_passes += W.kkt.stats.last_passes;
        VBK_LAUNCH(k_comp_dir, W.g(n), kVecThreads, 0, st, n, mu, W.x.p, W.z.p, W.D.p, W.dx.p, W.dz.p);
        VBK_LAUNCH(k_comp_dir, W.g(m), kVecThreads, 0, st, m, mu, W.y.p, W.w.p, W.E.p, W.dy.p, W.dw.p);
        W.launches(2);

        double theta = W.ratio_max();                        // intpt.c:211-220
        theta = VMIN(r / theta, 1.0);

        VBK_LAUNCH(k_step2, W.g(n), kVecThreads, 0, st, n, theta, W.dx.p, W.dz.p, W.x.p, W.z.p);
        VBK_LAUNCH(k_step2, W.g(m), kVecThreads, 0, st, m, theta, W.dy.p, W.dw.p, W.y.p, W.w.p);
        W.launches(2);
        normr0 = normr;
        norms0 = norms;
    }
    VBK_CHECK_LAUNCH();
    W.finish(x, y);
    t_last_iterations = iter;
    if (timed) {
        prof->iterations = iter;
        prof->total_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    }
    return status;
}

}  // namespace vbk
