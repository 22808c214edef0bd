// vbk_capi.cu -- extern "C" surface of libvbkkt.so (declared in include/vbkkt.h).
#include "../../include/vbkkt.h"
#include "vbk_kkt.h"
#include "vbk_linalg.h"
#include "vbk_solver.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

using namespace vbk;

struct vbk_kkt {
    Kkt impl;
    vbk_kkt(int device, int mode) : impl(device, mode) {}
};

namespace {

int g_mode = -1, g_device = -1;
void read_env_once()
{
    if (g_mode < 0) {
        const char* e = std::getenv("VBK_MODE");
        g_mode = (e && std::strcmp(e, "fast") == 0) ? VBK_MODE_FAST : VBK_MODE_STRICT;
    }
    if (g_device < 0) {
        const char* e = std::getenv("VBK_DEVICE");
        g_device = e ? std::atoi(e) : 0;
    }
}

// the reference's one-factor-object-per-process state (ldlt.c:108-120), bound to handle 0
std::unique_ptr<Kkt> g_kkt;
std::unique_ptr<LinalgContext> g_la;
const int* g_kA = nullptr; const int* g_iA = nullptr; const double* g_A = nullptr;
const int* g_kAt = nullptr; const int* g_iAt = nullptr; const double* g_At = nullptr;
DevArray<double> g_vx, g_vy;

LinalgContext& linalg()
{
    read_env_once();
    if (!g_la) g_la.reset(new LinalgContext(g_device, g_mode));
    return *g_la;
}

}  // namespace

extern "C" {

void vbk_set_mode(int mode) { g_mode = mode; }
int vbk_get_mode(void) { read_env_once(); return g_mode; }
void vbk_set_device(int device) { g_device = device; }
int vbk_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
const char* vbk_version(void)
{
#ifdef VBK_EMU
    return "vbkkt 0.1 (host emulation TEST build - not the product)";
#else
    return "vbkkt 0.1 (sm_100a)";
#endif
}

// ---------------------------------------------------------------------------------------- B1
// The reference's hsd.c calls forwardbackward twice per factorisation (hsd.c:223 and :228) and the SECOND right-hand
// side is the same every iteration (-b, -c).  Behind this one-right-hand-side-per-call interface the library therefore
// speculates: the right-hand side of a factorisation's second call is remembered, and the FIRST call after the next
// ldltfac solves it along with its own in one pair of sweeps (Kkt::solve2_host -- a sweep is bound by its dependency
// chain, the second right-hand side costs ~2 %).  When the second call then arrives with bit-identical Dn, Dm and
// right-hand side, it is answered from that result -- the same bits the call would have computed; any difference in the
// inputs and it is solved as usual.  METHODs with one solve per factorisation (intpt) never trigger it.
// $VBK_SPECULATE=0 switches it off.
namespace {
struct Speculation {
    bool enabled = true, env_read = false;
    long calls = 0;                               // forwardbackward calls since the last ldltfac
    bool have_pred = false, have_result = false;
    std::vector<double> pred_c, pred_b;           // right-hand side of the previous factorisation's second call
    std::vector<double> res_c, res_b, res_dn, res_dm;
    int res_consistent = 1, res_passes = 0;
} g_spec;

bool same_bits(const double* a, const std::vector<double>& b) { return std::memcmp(a, b.data(), b.size() * sizeof(double)) == 0; }

int b1_solve(double* Dn, double* Dm, double* c, double* b)
{
    if (!g_spec.env_read) {
        const char* e = std::getenv("VBK_SPECULATE");
        g_spec.enabled = !(e && std::atoi(e) == 0);
        g_spec.env_read = true;
    }
    const size_t n = (size_t)g_kkt->sym().n, m = (size_t)g_kkt->sym().m;
    const long call = ++g_spec.calls;
    if (!g_spec.enabled) return g_kkt->solve_host(Dn, Dm, c, b);
    if (call >= 2 && g_spec.have_result && same_bits(c, g_spec.pred_c) && same_bits(b, g_spec.pred_b) &&
        same_bits(Dn, g_spec.res_dn) && same_bits(Dm, g_spec.res_dm)) {
        std::memcpy(c, g_spec.res_c.data(), n * sizeof(double));
        std::memcpy(b, g_spec.res_b.data(), m * sizeof(double));
        g_spec.have_result = false;
        g_kkt->stats.last_passes = g_spec.res_passes;
        g_kkt->stats.last_consistent = g_spec.res_consistent;
        return g_spec.res_consistent;
    }
    if (call == 2) {                              // what the next factorisation's second call will most likely ask
        g_spec.pred_c.assign(c, c + n);
        g_spec.pred_b.assign(b, b + m);
        g_spec.have_pred = true;
    }
    if (call == 1 && g_spec.have_pred) {
        g_spec.res_c = g_spec.pred_c;
        g_spec.res_b = g_spec.pred_b;
        const int both = g_kkt->solve2_host(Dn, Dm, c, b, g_spec.res_c.data(), g_spec.res_b.data());
        g_spec.res_dn.assign(Dn, Dn + n);
        g_spec.res_dm.assign(Dm, Dm + m);
        g_spec.res_consistent = (both >> 1) & 1;
        g_spec.res_passes = g_kkt->stats.last_passes2[1];
        g_spec.have_result = true;
        g_kkt->stats.last_passes = g_kkt->stats.last_passes2[0];
        g_kkt->stats.last_consistent = both & 1;
        return both & 1;
    }
    return g_kkt->solve_host(Dn, Dm, c, b);
}
}  // namespace

void ldltfac(int m, int n, int* kA, int* iA, double* A, double* dn, double* dm,
             int* kAt, int* iAt, double* At, int verbose)
{
    (void)verbose;
    read_env_once();
    if (!g_kkt) {
        g_kkt.reset(new Kkt(g_device, g_mode));
        g_kkt->analyze(m, n, kA, iA, A, kAt, iAt, At);
        g_kA = kA; g_iA = iA; g_A = A; g_kAt = kAt; g_iAt = iAt; g_At = At;
    }
    g_kkt->factor_host(dn, dm);
    g_spec.calls = 0;                  // a new factorisation: earlier results no longer apply
    g_spec.have_result = false;
}

void forwardbackward(double* Dn, double* Dm, double* dx, double* dy)
{
    if (!g_kkt) { std::fprintf(stderr, "vbkkt: forwardbackward() before ldltfac()\n"); std::exit(1); }
    b1_solve(Dn, Dm, dx, dy);
}

// lp.h:213-227: the LP-struct forms (ldlt.c:164, :327).  The factor object is bound to the first LP seen, like the
// reference's statics.
void inv_num(void* lp_, double* dn, double* dm)
{
    const vbk_lp_head* lp = static_cast<const vbk_lp_head*>(lp_);
    if (lp->qnz != 0) { std::fprintf(stderr, "vbkkt: inv_num: quadratic terms (qnz = %d) are not supported\n", lp->qnz); std::exit(1); }
    ldltfac(lp->m, lp->n, lp->kA, lp->iA, lp->A, dn, dm, lp->kAt, lp->iAt, lp->At, 0);
}
int solve(void* lp_, double* Dn, double* Dm, double* c, double* b)
{
    (void)lp_;
    if (!g_kkt) { std::fprintf(stderr, "vbkkt: solve() before inv_num()\n"); std::exit(1); }
    return b1_solve(Dn, Dm, c, b);
}

void inv_clo(void)
{
    g_kkt.reset();
    g_spec = Speculation();
    g_kA = g_iA = g_kAt = g_iAt = nullptr;
    g_A = g_At = nullptr;
}

double dotprod(double* x, double* y, int n) { return linalg().dotprod_host(x, y, n); }
double maxv(double* x, int n) { return linalg().maxv_host(x, n); }
void atnum(int m, int n, int* ka, int* ia, double* a, int* kat, int* iat, double* at)
{
    linalg().atnum_host(m, n, ka, ia, a, kat, iat, at);
}
void smx(int m, int n, double* a, int* ka, int* ia, double* x, double* y)
{
    // the matrix ldltfac captured is already resident in gather form: only the vectors move
    if (g_kkt && a == g_A && ka == g_kA && ia == g_iA && m == g_kkt->sym().m && n == g_kkt->sym().n) {
        cudaStream_t s = g_kkt->stream();
        g_vx.upload(x, n, s); g_vy.alloc(m);
        g_kkt->spmv_A(g_vx.p, g_vy.p);
        g_vy.download(y, m, s);
        VBK_CUDA(cudaStreamSynchronize(s));
        return;
    }
    if (g_kkt && a == g_At && ka == g_kAt && ia == g_iAt && m == g_kkt->sym().n && n == g_kkt->sym().m) {
        cudaStream_t s = g_kkt->stream();
        g_vx.upload(x, n, s); g_vy.alloc(m);
        g_kkt->spmv_At(g_vx.p, g_vy.p);
        g_vy.download(y, m, s);
        VBK_CUDA(cudaStreamSynchronize(s));
        return;
    }
    linalg().smx_host(m, n, a, ka, ia, x, y);
}

// ---------------------------------------------------------------------------------------- B2
int vbk_solver_hsd(int m, int n, int nz, int* iA, int* kA, double* A, double* b, double* c, double f,
                   double* x, double* y, double* w, double* z)
{
    read_env_once();
    int st = solver_hsd(g_device, g_mode, m, n, nz, iA, kA, A, b, c, f, x, y, nullptr);
    std::free(w); std::free(z);       // the reference's plugins free these (hsd.c:290-291)
    return st;
}
int vbk_solver_hsdls(int m, int n, int nz, int* iA, int* kA, double* A, double* b, double* c, double f,
                     double* x, double* y, double* w, double* z)
{
    read_env_once();
    int st = solver_hsdls(g_device, g_mode, m, n, nz, iA, kA, A, b, c, f, x, y, nullptr);
    std::free(w); std::free(z);       // hsdls.c:273-274
    return st;
}
int vbk_solver_intpt(int m, int n, int nz, int* iA, int* kA, double* A, double* b, double* c, double f,
                     double* x, double* y, double* w, double* z)
{
    read_env_once();
    int st = solver_intpt(g_device, g_mode, m, n, nz, iA, kA, A, b, c, f, x, y, nullptr);
    std::free(w); std::free(z);       // intpt.c:244-245
    return st;
}

int vbk_solve_lp(int method, int device, int mode, int m, int n, int nz, const int* iA, const int* kA,
                 const double* A, const double* b, const double* c, double f,
                 double* x, double* y, vbk_profile* prof)
{
    SolveProfile p;
    int st = method == 0 ? solver_hsd(device, mode, m, n, nz, iA, kA, A, b, c, f, x, y, prof ? &p : nullptr)
           : method == 2 ? solver_hsdls(device, mode, m, n, nz, iA, kA, A, b, c, f, x, y, prof ? &p : nullptr)
                         : solver_intpt(device, mode, m, n, nz, iA, kA, A, b, c, f, x, y, prof ? &p : nullptr);
    if (prof) {
        prof->total_s = p.total_s; prof->setup_s = p.setup_s; prof->factor_s = p.factor_s; prof->solve_s = p.solve_s;
        prof->factor_calls = p.factor_calls; prof->solve_calls = p.solve_calls; prof->rawsolve_calls = p.rawsolve_calls;
        prof->kernel_launches = p.kernel_launches; prof->refine_passes = p.refine_passes;
        prof->iterations = p.iterations; prof->N = p.N; prof->lnz = p.lnz; prof->narth = p.narth;
    }
    return st;
}

void vbk_set_iteration_limit(int itnlim) { set_iteration_limit(itnlim); }
float vbk_kkt_last_factor_kernel_ms(vbk_kkt* h) { return h->impl.last_factor_kernel_ms(); }
double vbk_measure_fp64_tflops(int device) { return measure_fp64_tflops(device); }
double vbk_measure_hbm_gbs(int device) { return measure_hbm_gbs(device); }

void vbk_kkt_phase_profile(vbk_kkt* h, unsigned long long* out16) { h->impl.read_phase_profile(out16); }
void vbk_kkt_trace(vbk_kkt* h, long long* out) { h->impl.read_trace(out); }

void vbk_capture(int iter, double* E, double* D, double* rhs_y, double* rhs_x, double* sol_y, double* sol_x)
{
    set_capture(iter, E, D, rhs_y, rhs_x, sol_y, sol_x);
}

// ---------------------------------------------------------------------------------------- H
vbk_kkt* vbk_kkt_create(int device, int mode) { return new vbk_kkt(device, mode); }
void vbk_kkt_destroy(vbk_kkt* h) { delete h; }
void vbk_kkt_analyze(vbk_kkt* h, int m, int n, const int* kA, const int* iA, const double* A,
                     const int* kAt, const int* iAt, const double* At)
{
    h->impl.analyze(m, n, kA, iA, A, kAt, iAt, At);
}
void vbk_kkt_factor(vbk_kkt* h, const double* dn, const double* dm) { h->impl.factor_host(dn, dm); }
int vbk_kkt_solve(vbk_kkt* h, const double* Dn, const double* Dm, double* dx, double* dy)
{
    return h->impl.solve_host(Dn, Dm, dx, dy);
}
int vbk_kkt_solve2(vbk_kkt* h, const double* Dn, const double* Dm, double* dx0, double* dy0, double* dx1, double* dy1)
{
    return h->impl.solve2_host(Dn, Dm, dx0, dy0, dx1, dy1);
}
int vbk_kkt_solve2_dev(vbk_kkt* h, const double* Dn, const double* Dm, double* dx0, double* dy0, double* dx1, double* dy1)
{
    int cons[2] = {1, 1};
    h->impl.solve2_dev(Dn, Dm, dx0, dy0, dx1, dy1, cons);
    return cons[0] | (cons[1] << 1);
}
void vbk_kkt_factor_dev(vbk_kkt* h, const double* dn, const double* dm) { h->impl.factor_dev(dn, dm); }
int vbk_kkt_solve_dev(vbk_kkt* h, const double* Dn, const double* Dm, double* dx, double* dy)
{
    return h->impl.solve_dev(Dn, Dm, dx, dy);
}
int vbk_kkt_rawsolve(vbk_kkt* h, double* zperm)
{
    Kkt& k = h->impl;
    const size_t N = (size_t)k.sym().N;
    VBK_CUDA(cudaMemcpyAsync(k.zbuf(), zperm, 8 * N, cudaMemcpyHostToDevice, k.stream()));
    k.rawsolve_dev();
    VBK_CUDA(cudaMemcpyAsync(zperm, k.zbuf(), 8 * N, cudaMemcpyDeviceToHost, k.stream()));
    VBK_CUDA(cudaStreamSynchronize(k.stream()));
    return 1;
}
void vbk_kkt_sync(vbk_kkt* h) { VBK_CUDA(cudaStreamSynchronize(h->impl.stream())); }
void* vbk_kkt_stream(vbk_kkt* h) { return (void*)(size_t)h->impl.stream(); }

int vbk_kkt_dim(const vbk_kkt* h) { return h->impl.sym().N; }
long long vbk_kkt_lnz(const vbk_kkt* h) { return h->impl.sym().lnz(); }
int vbk_kkt_denwin(const vbk_kkt* h) { return h->impl.sym().denwin; }
int vbk_kkt_pdf(const vbk_kkt* h) { return h->impl.sym().pdf; }
int vbk_kkt_window(const vbk_kkt* h) { return h->impl.sym().N - h->impl.sym().dense_start; }
double vbk_kkt_narth(const vbk_kkt* h) { return h->impl.sym().narth; }
int vbk_kkt_nlevels(const vbk_kkt* h) { return h->impl.sym().nlevels; }
int vbk_kkt_nsupernodes(const vbk_kkt* h) { return (int)h->impl.sym().sn_ptr.size() - 1; }
const int* vbk_kkt_perm(const vbk_kkt* h) { return h->impl.sym().perm.data(); }
const int* vbk_kkt_iperm(const vbk_kkt* h) { return h->impl.sym().iperm.data(); }
const int* vbk_kkt_kAAt(const vbk_kkt* h) { return h->impl.sym().kL.data(); }
const int* vbk_kkt_iAAt(const vbk_kkt* h) { return h->impl.sym().iL.data(); }
void vbk_kkt_get_factor(vbk_kkt* h, double* L, double* diag, int* mark) { h->impl.download_factor(L, diag, mark); }
double vbk_kkt_epsdiag(vbk_kkt* h) { return h->impl.epsdiag(); }
int vbk_kkt_ndep(vbk_kkt* h) { return h->impl.ndep(); }
int vbk_kkt_last_passes(const vbk_kkt* h) { return h->impl.stats.last_passes; }
int vbk_kkt_last_passes2(const vbk_kkt* h, int rhs) { return h->impl.stats.last_passes2[rhs ? 1 : 0]; }
long long vbk_kkt_launches(const vbk_kkt* h) { return h->impl.stats.kernel_launches; }

}  // extern "C"
