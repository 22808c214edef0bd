// vbk_batch.cu -- batches of independent LPs on one GPU (BASELINE.json config 4).
//
// The reference solves one LP per process (file-scope factor object, src/ipo/ldlt.c:108-120; METHOD
// entry point `solver`, src/common/solve.c:237).  Here every LP of a batch gets its own factor object,
// its own CUDA stream and a host thread that runs the device-resident METHOD loop (vbk_solver.cu) for
// it; `nstreams` such solves are in flight at once, so that while one LP waits for its per-iteration
// scalars (or sits in the dependent chain of its dense window) the SMs work on the others.  Nothing is
// shared between LPs and there is no collective: across GPUs the batch is dealt round-robin by the
// caller (one process per GPU, linear-programming-vanderbei_b200/batch.py).
#include "../../include/vbkkt.h"
#include "vbk_device.h"
#include "vbk_solver.h"

#include <atomic>
#include <chrono>
#include <csignal>
#include <cstdlib>
#include <execinfo.h>
#include <unistd.h>
#include <thread>
#include <vector>

using namespace vbk;

namespace {
// $VBK_SEGV_TRACE (debugging aid): raw backtrace of the faulting thread on stderr, then the default action.
// Runs on an alternate stack so that a stack overflow is reported too.
void segv_trace(int sig)
{
    void* frames[64];
    const int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    std::signal(sig, SIG_DFL);
    raise(sig);
}
void install_segv_trace()
{
    if (!std::getenv("VBK_SEGV_TRACE")) return;
    void* warm[4];
    backtrace(warm, 4);                      // loads libgcc now, not inside the handler
    stack_t ss;
    ss.ss_sp = std::malloc(1 << 16); ss.ss_size = 1 << 16; ss.ss_flags = 0;
    sigaltstack(&ss, nullptr);
    struct sigaction sa;
    sa.sa_handler = segv_trace; sigemptyset(&sa.sa_mask); sa.sa_flags = SA_ONSTACK;
    sigaction(SIGSEGV, &sa, nullptr);
    sigaction(SIGBUS, &sa, nullptr);
    sigaction(SIGABRT, &sa, nullptr);
}
}  // namespace

static std::atomic<long long> g_batch_launches(0);
extern "C" long long vbk_batch_launches(void) { return g_batch_launches.load(); }

extern "C" int vbk_solve_batch(int method, int device, int mode, int nlp, vbk_lp_desc* lps, int nstreams)
{
    if (nlp <= 0) return 0;
    if (nstreams < 1) nstreams = 1;
    if (nstreams > nlp) nstreams = nlp;
    std::atomic<int> next(0), failed(0);
    g_batch_launches.store(0);
    auto worker = [&]() {
#ifndef VBK_EMU
        VBK_CUDA(cudaSetDevice(device));
#endif
        install_segv_trace();
        set_thread_quiet(true);
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= nlp) break;
            vbk_lp_desc& d = lps[i];
            const auto t0 = std::chrono::steady_clock::now();
            d.status = method == 0
                ? solver_hsd(device, mode, d.m, d.n, d.nz, d.iA, d.kA, d.A, d.b, d.c, d.f, d.x, d.y, nullptr)
                : method == 2
                ? solver_hsdls(device, mode, d.m, d.n, d.nz, d.iA, d.kA, d.A, d.b, d.c, d.f, d.x, d.y, nullptr)
                : solver_intpt(device, mode, d.m, d.n, d.nz, d.iA, d.kA, d.A, d.b, d.c, d.f, d.x, d.y, nullptr);
            d.iterations = last_thread_iterations();
            g_batch_launches.fetch_add(last_thread_launches());
            d.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            double po = d.f, du = d.f;              // solve.c:254-255 (objective of the form the solver saw)
            for (int j = 0; j < d.n; ++j) po += d.c[j] * d.x[j];
            for (int r = 0; r < d.m; ++r) du += d.b[r] * d.y[r];
            d.primal_obj = po; d.dual_obj = du;
            if (d.status != 0) failed.fetch_add(1);
        }
        set_thread_quiet(false);
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nstreams; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    return failed.load();
}
