// vbk_solver.h -- device-resident METHOD plugins (hsd / hsdls / intpt), see vbk_solver.cu.
#pragma once

namespace vbk {

struct SolveProfile {
    double total_s = 0, setup_s = 0, factor_s = 0, solve_s = 0;
    long long factor_calls = 0, solve_calls = 0, rawsolve_calls = 0, kernel_launches = 0;
    long long refine_passes = 0;
    int iterations = 0, N = 0;
    long long lnz = 0;
    double narth = 0;
};

// status: 0 optimal, 2 primal infeasible, 4 dual infeasible, 5 iteration limit, 7 numerical problem (hsdls) (main.c:21-30)
int solver_hsd(int device, int mode, int m, int n, int nz, const int* iA, const int* kA, const double* A,
               const double* b, const double* c, double f, double* x, double* y, SolveProfile* prof);
// METHOD = hsdls (reference src/ipo/hsdls.c:37): homogeneous self-dual, long steps with a per-component line search
int solver_hsdls(int device, int mode, int m, int n, int nz, const int* iA, const int* kA, const double* A,
                 const double* b, const double* c, double f, double* x, double* y, SolveProfile* prof);
int solver_intpt(int device, int mode, int m, int n, int nz, const int* iA, const int* kA, const double* A,
                 const double* b, const double* c, double f, double* x, double* y, SolveProfile* prof);

// Test hook: copy the KKT-step inputs (E[m], D[n], rhs_y[m], rhs_x[n]) and outputs (sol_y, sol_x) of
// iteration `iter` of the next solver_* call into host buffers; iter < 0 disables.
// MAX_ITER of hsd.c:25 / intpt.c:31 is a compile-time 200 in the reference; tests and the bench may
// lower it (e.g. to stop right after a captured iteration).  <= 0 restores 200.
void set_iteration_limit(int itnlim);

// per-thread: suppress the iteration log of solver_* calls made by this thread (batch driver);
// iteration count of the last solver_* call made by this thread
void set_thread_quiet(bool quiet);
int last_thread_iterations();
long long last_thread_launches();   // kernels launched by the last solver_* call made by this thread

void set_capture(int iter, double* E, double* D, double* rhs_y, double* rhs_x, double* sol_y, double* sol_x);

}  // namespace vbk
