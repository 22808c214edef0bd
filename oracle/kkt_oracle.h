/*
 * oracle/kkt_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-thread restatement of the reference's KKT-step hot path
 * (romz-pl/linear-programming-Vanderbei: src/ipo/ldlt.c, src/common/linalg.c,
 * src/ipo/hsd.c, src/ipo/intpt.c).  It exists to CHECK the CUDA product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product (libvbkkt.so) never links, loads or calls it.
 *
 * Parity pinning: tests/test_oracle.py runs this restatement against
 *   (a) the reference's own golden logs evaluate/v1-cf4d5ba/netlib/ipo/<name>.mps.sol
 *       (committed as tests/golden/netlib/<name>.npz, byte-for-byte log equality), and
 *   (b) the compiled reference oracle/_ref/libref_{hsd,intpt}.so where that is present
 *       (bit-equal perm/iperm/kAAt/iAAt/diag/AAt and x,y).
 *
 * Unlike the reference (file-scope statics, ldlt.c:108-120) the factor object is a handle.
 */
#ifndef KKT_ORACLE_H
#define KKT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kko_factor kko_factor;

/* ---- linalg.c ---- */
double kko_dotprod(const double *x, const double *y, int n);              /* linalg.c:17-25  */
void   kko_smx(int m, int n, const double *a, const int *ka, const int *ia,
               const double *x, double *y);                               /* linalg.c:62-70  */
void   kko_atnum(int m, int n, const int *ka, const int *ia, const double *a,
                 int *kat, int *iat, double *at);                         /* linalg.c:75-103 */
double kko_maxv(const double *x, int n);                                  /* linalg.c:108-116 */

/* ---- ldlt.c (argument meaning identical to ldlt.h:1-20; m,n are ldlt-space) ---- */
kko_factor *kko_create(void);
void        kko_destroy(kko_factor *F);
void kko_ldltfac(kko_factor *F, int m, int n, const int *kA, const int *iA, const double *A,
                 const double *dn, const double *dm,
                 const int *kAt, const int *iAt, const double *At);       /* ldlt.c:124-309 */
int  kko_forwardbackward(kko_factor *F, const double *Dn, const double *Dm,
                         double *dx, double *dy);                         /* ldlt.c:311-425 */
int  kko_rawsolve(kko_factor *F, double *zperm);                          /* ldlt.c:433-505 */
int  kko_last_passes(const kko_factor *F);   /* refinement passes of the last forwardbackward */

/* accessors (lengths: N=m+n, Lnz=kAAt[N]) */
int           kko_dim(const kko_factor *F);
int           kko_denwin(const kko_factor *F);
int           kko_pdf(const kko_factor *F);
int           kko_ndep(const kko_factor *F);
double        kko_epsdiag(const kko_factor *F);
const int    *kko_perm(const kko_factor *F);
const int    *kko_iperm(const kko_factor *F);
const int    *kko_kAAt(const kko_factor *F);
const int    *kko_iAAt(const kko_factor *F);
const double *kko_AAt(const kko_factor *F);
const double *kko_diag(const kko_factor *F);
const int    *kko_mark(const kko_factor *F);

/* ---- METHOD plugins; same signature and stdout as hsd.c:27 / intpt.c:33 (they free w and z) ---- */
int kko_solver_hsd(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c,
                   double f, double *x, double *y, double *w, double *z);
int kko_solver_intpt(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c,
                     double f, double *x, double *y, double *w, double *z);

/* Ask the next kko_solver_* call to copy the KKT-step inputs/outputs of iteration `iter`
 * (E[m], D[n], rhs_y[m], rhs_x[n] before the first forwardbackward, and the solution after)
 * into caller buffers; pass iter<0 to disable.  Used to make per-call parity vectors. */
void kko_capture(int iter, double *E, double *D, double *rhs_y, double *rhs_x,
                 double *sol_y, double *sol_x);
/* The reference's MAX_ITER is a compile-time 200; fixtures of mid-solve iterates stop earlier. */
void kko_set_itnlim(int itnlim);
/* wall-clock seconds spent inside ldltfac / forwardbackward during the last kko_solver_* call */
void kko_last_timing(double *t_factor, double *t_solve, int *n_factor, int *n_solve, int *n_rawsolve);

#ifdef __cplusplus
}
#endif
#endif
