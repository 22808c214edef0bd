/*
 * vbkkt.h -- C ABI of libvbkkt.so, the B200 (sm_100a) implementation of the per-iteration KKT step of
 * Vanderbei's interior-point LP codes.  Plain C types only; every pointer is HOST memory unless the
 * name says `_dev`.  There is no CPU path: every numeric entry point prints a message and exit(1)s
 * when no CUDA device is usable (the reference's own error convention, src/common/myalloc.h:17-24).
 *
 * Three groups:
 *   B1  the reference's plugin symbols, exact signatures, so that the reference's C driver links
 *       libvbkkt.so in place of ldlt.o and the linalg.o member of common.a (src/ipo/makefile:49-64);
 *   B2  the METHOD plugin `solver` (same signature/stdout/status as src/ipo/hsd.c:27, intpt.c:33),
 *       device-resident; exported as vbk_solver_hsd / vbk_solver_hsdls / vbk_solver_intpt here and as plain
 *       `solver` by the one-function shims libvbkkt_hsd.so / libvbkkt_hsdls.so / libvbkkt_intpt.so;
 *   H   a handle-based API (no reference equivalent: the reference has one process-global factor
 *       object, src/ipo/ldlt.c:108-120) for tests, the batch driver and the bench.
 */
#ifndef VBKKT_H
#define VBKKT_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * B1 -- replaces src/ipo/ldlt.c (declared in src/ipo/ldlt.h:1-20 and src/common/lp.h:213-239)
 * ---------------------------------------------------------------------------------------------- */
/* ldlt.h:1-13.  First call: symbolic analysis of K = [-diag(dn) A^T; A diag(dm)] and upload of the
 * matrix (pointers are retained like the reference does, ldlt.c:140-160); every call: numeric LDL^T. */
void ldltfac(int m, int n, int *kA, int *iA, double *A, double *dn, double *dm,
             int *kAt, int *iAt, double *At, int verbose);
/* ldlt.h:15-20.  Solves K [dx;dy] = [dx;dy] in place with iterative refinement (ldlt.c:327-425). */
void forwardbackward(double *Dn, double *Dm, double *dx, double *dy);
/* lp.h:236-239 / ldlt.c:507-513: releases the process-global factor object. */
void inv_clo(void);
/* lp.h:213-227 -- the LP-struct forms of the same two calls, for drivers that hold the reference's `LP` (amplio and the
 * like; ipo itself goes through ldltfac / forwardbackward).  Only the leading members of `LP` are read -- m, n, the CSC
 * triples (A, iA, kA) and (At, iAt, kAt) -- through the mirror below, which repeats lp.h:34-62 member for member so that
 * the offsets agree under the platform's C ABI.  Q must be empty (qnz == 0), as it is on every ipo path (ldlt.c:150-153). */
struct vbk_lp_head {
    int m, n, nz;
    double *A; int *iA; int *kA;
    double *b; double *c; double f; double *r; double *l; double *u;
    int *varsgn; char **rowlab; char **collab;
    int qnz; double *Q; int *iQ; int *kQ;
    double *At; int *iAt; int *kAt;
};
void inv_num(void *lp, double *dn, double *dm);                        /* ldlt.c:164-309 */
int  solve(void *lp, double *Dn, double *Dm, double *c, double *b);    /* ldlt.c:327-425; returns `consistent` */

/* replaces the ipo-relevant part of src/common/linalg.c (declared in src/common/linalg.h:1-8) */
double dotprod(double *x, double *y, int n);                                   /* linalg.c:17-25  */
void   smx(int m, int n, double *a, int *ka, int *ia, double *x, double *y);   /* linalg.c:62-70  */
void   atnum(int m, int n, int *ka, int *ia, double *a,
             int *kat, int *iat, double *at);                                  /* linalg.c:75-103 */
double maxv(double *x, int n);                                                 /* linalg.c:108-116 */

/* ------------------------------------------------------------------------------------------------
 * B2 -- replaces the METHOD object (src/ipo/hsd.c:27 / src/ipo/intpt.c:33; caller: solve.c:237)
 * ---------------------------------------------------------------------------------------------- */
int vbk_solver_hsd(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c, double f,
                   double *x, double *y, double *w, double *z);
int vbk_solver_intpt(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c, double f,
                     double *x, double *y, double *w, double *z);
/* METHOD = hsdls (src/ipo/hsdls.c:37): the long-step variant of hsd, per-component line search (hsdls.c:296-336),
 * up to 600 iterations, status 7 = numerical problem.  Plain `solver` in libvbkkt_hsdls.so. */
int vbk_solver_hsdls(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c, double f,
                     double *x, double *y, double *w, double *z);

/* ------------------------------------------------------------------------------------------------
 * H -- handle API
 * ---------------------------------------------------------------------------------------------- */
#define VBK_MODE_STRICT 0   /* order-faithful, unfused: reproduces the reference bit for bit */
#define VBK_MODE_FAST   1   /* tree reductions / supernodal panels: tolerance parity          */

/* process-wide defaults used by the B1/B2 symbols (also read once from $VBK_MODE=strict|fast and
 * $VBK_DEVICE=<ordinal>) */
void vbk_set_mode(int mode);
int  vbk_get_mode(void);
void vbk_set_device(int device);
int  vbk_device_count(void);          /* 0 when no usable CUDA device (never exits) */
const char *vbk_version(void);

typedef struct vbk_kkt vbk_kkt;
/* device < 0: host-side symbolic analysis only (numeric calls on such a handle exit(1)) */
vbk_kkt *vbk_kkt_create(int device, int mode);
void     vbk_kkt_destroy(vbk_kkt *h);
/* ldlt-space arguments exactly as ldltfac: A is m x n CSC, At its transpose */
void vbk_kkt_analyze(vbk_kkt *h, int m, int n, const int *kA, const int *iA, const double *A,
                     const int *kAt, const int *iAt, const double *At);
void vbk_kkt_factor(vbk_kkt *h, const double *dn, const double *dm);
int  vbk_kkt_solve(vbk_kkt *h, const double *Dn, const double *Dm, double *dx, double *dy);
/* The two systems of one hsd iteration (reference src/ipo/hsd.c:223 and :228: two forwardbackward calls on the same
 * factor with independent right-hand sides) in one pair of sweeps per refinement pass.  Each right-hand side gets the
 * reference's own arithmetic and refinement rule (ldlt.c:367-416): the results equal two vbk_kkt_solve calls bit for bit.
 * Returns consistent0 | consistent1 << 1; vbk_kkt_last_passes2 gives the refinement passes of each. */
int  vbk_kkt_solve2(vbk_kkt *h, const double *Dn, const double *Dm, double *dx0, double *dy0, double *dx1, double *dy1);
/* device-pointer variants (no host traffic); all on the handle's stream, asynchronous */
void vbk_kkt_factor_dev(vbk_kkt *h, const double *dn_dev, const double *dm_dev);
int  vbk_kkt_solve_dev(vbk_kkt *h, const double *Dn_dev, const double *Dm_dev, double *dx_dev, double *dy_dev);
int  vbk_kkt_solve2_dev(vbk_kkt *h, const double *Dn_dev, const double *Dm_dev, double *dx0_dev, double *dy0_dev,
                        double *dx1_dev, double *dy1_dev);
/* one raw forward/diagonal/backward sweep (ldlt.c:433-505) on a permuted host vector of length m+n */
int  vbk_kkt_rawsolve(vbk_kkt *h, double *zperm);
void vbk_kkt_sync(vbk_kkt *h);
void *vbk_kkt_stream(vbk_kkt *h);     /* cudaStream_t */

/* symbolic results (valid after analyze; host memory owned by the handle) */
int        vbk_kkt_dim(const vbk_kkt *h);
long long  vbk_kkt_lnz(const vbk_kkt *h);
int        vbk_kkt_denwin(const vbk_kkt *h);
int        vbk_kkt_pdf(const vbk_kkt *h);
/* width of the trailing window fast mode factorises densely (>= dim - denwin: the window is padded, see
 * DESIGN.md); 0 when there is none */
int        vbk_kkt_window(const vbk_kkt *h);
double     vbk_kkt_narth(const vbk_kkt *h);
int        vbk_kkt_nlevels(const vbk_kkt *h);
int        vbk_kkt_nsupernodes(const vbk_kkt *h);
const int *vbk_kkt_perm(const vbk_kkt *h);
const int *vbk_kkt_iperm(const vbk_kkt *h);
const int *vbk_kkt_kAAt(const vbk_kkt *h);
const int *vbk_kkt_iAAt(const vbk_kkt *h);
/* numeric results: copies L values [lnz], diag [dim], mark [dim] to host (any may be NULL) */
void   vbk_kkt_get_factor(vbk_kkt *h, double *L, double *diag, int *mark);
double vbk_kkt_epsdiag(vbk_kkt *h);
int    vbk_kkt_ndep(vbk_kkt *h);
int    vbk_kkt_last_passes(const vbk_kkt *h);
int    vbk_kkt_last_passes2(const vbk_kkt *h, int rhs);    /* after vbk_kkt_solve2: rhs = 0 or 1 */
long long vbk_kkt_launches(const vbk_kkt *h);

/* profile of the last vbk_solver_* / vbk_solve_lp call (seconds; GPU work bracketed by stream syncs) */
typedef struct vbk_profile {
    double total_s, setup_s, factor_s, solve_s;
    long long factor_calls, solve_calls, rawsolve_calls, kernel_launches, refine_passes;
    int iterations, N;
    long long lnz;
    double narth;
} vbk_profile;
/* method: 0 = hsd, 1 = intpt, 2 = hsdls.  Like vbk_solver_* but on an explicit device/mode, const inputs, no
 * ownership quirks (w,z are not touched), optional profile. */
int vbk_solve_lp(int method, int device, int mode, int m, int n, int nz, const int *iA, const int *kA,
                 const double *A, const double *b, const double *c, double f,
                 double *x, double *y, vbk_profile *prof);

/* ------------------------------------------------------------------------------------------------
 * Batches of independent LPs (BASELINE.json config 4).  No reference equivalent: the reference
 * solves one LP per process through `solver` (src/common/solve.c:237); a batch is `nlp` such calls,
 * `nstreams` of them in flight at once on `device`, each with its own factor object and CUDA stream.
 * Across GPUs the caller deals the batch round-robin, one process per GPU, no collective.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vbk_lp_desc {
    int m, n, nz;                     /* solver-space dimensions, as passed to `solver` */
    const int *iA, *kA;               /* CSC row indices [nz], column pointers [n+1]    */
    const double *A, *b, *c;          /* values [nz], right-hand side [m], objective [n] */
    double f;                         /* objective offset                                 */
    double *x, *y;                    /* out: primal [n], dual [m] (caller-allocated)     */
    int status, iterations;           /* out: solver() return code, iteration count       */
    double primal_obj, dual_obj;      /* out: c.x + f, b.y + f (solve.c:254-255)          */
    double seconds;                   /* out: wall time of this LP's solve                */
} vbk_lp_desc;
/* method: 0 = hsd, 1 = intpt, 2 = hsdls.  Returns the number of LPs whose status is not 0. */
int vbk_solve_batch(int method, int device, int mode, int nlp, vbk_lp_desc *lps, int nstreams);
/* kernels launched by all LPs of the last vbk_solve_batch call of this process */
long long vbk_batch_launches(void);

/* ------------------------------------------------------------------------------------------------
 * Row-block partitioned smx / dotprod / maxv (BASELINE.json config 5): the per-rank pieces.  Device
 * pointers, asynchronous on `stream` (a cudaStream_t); the all-gather / all-reduce between them is the
 * caller's (NCCL through torch.distributed, see rowblock.py).
 * ---------------------------------------------------------------------------------------------- */
/* y[r] = sum_k val[k] * x[idx[k]], k in [ptr[r], ptr[r+1]), ascending k: with the rows of a matrix in
 * ascending column order this is the summation order of the reference's smx (linalg.c:62-70). */
void vbk_spmv_rows_dev(int nrows, const int *ptr_dev, const int *idx_dev, const double *val_dev,
                       const double *x_dev, double *y_dev, void *stream);
/* out_dev[q] = x_q . y_q over n[q] local entries, q < count <= 8 (fixed-shape tree: deterministic, not the
 * reference's left-to-right order of linalg.c:17-25).  scratch_dev: vbk_reduce_scratch_doubles() doubles. */
void vbk_dots_partial_dev(int count, const double *const *x_dev, const double *const *y_dev, const long long *n,
                          double *out_dev, double *scratch_dev, void *stream);
/* out_dev[q] = max_i |x_q[i]| (linalg.c:108-116), q < count <= 8 */
void vbk_absmax_partial_dev(int count, const double *const *x_dev, const long long *n, double *out_dev, void *stream);
int  vbk_reduce_scratch_doubles(void);

/* MAX_ITER is a compile-time 200 in the reference (hsd.c:25, intpt.c:31); <=0 restores it */
void vbk_set_iteration_limit(int itnlim);
/* device time (ms, CUDA events on the handle's stream) of the last numeric-factor kernel */
float vbk_kkt_last_factor_kernel_ms(vbk_kkt *h);
/* with $VBK_PROF set: SM cycles per phase of the tiled factor kernel since the last call (8 counters:
 * claim+init, wait, stage, scan, scatter, accumulate, pivot, write) */
void vbk_kkt_phase_profile(vbk_kkt *h, unsigned long long *out16);
/* with $VBK_PROF set: per-column event times (ns) of the last strict factorisation, N x 8 values (csrc/vbk_strict_factor.cuh) */
void vbk_kkt_trace(vbk_kkt *h, long long *out);
/* roofline yardsticks measured on the spot: FP64 DFMA TFLOP/s and device copy GB/s */
double vbk_measure_fp64_tflops(int device);
double vbk_measure_hbm_gbs(int device);

/* Test hook (mirrors the oracle's kko_capture): copy the KKT-step inputs E[m], D[n], rhs_y[m], rhs_x[n]
 * and outputs sol_y[m], sol_x[n] of iteration `iter` of the next solve into host buffers; iter<0 = off. */
void vbk_capture(int iter, double *E, double *D, double *rhs_y, double *rhs_x, double *sol_y, double *sol_x);

#ifdef __cplusplus
}
#endif
#endif /* VBKKT_H */
