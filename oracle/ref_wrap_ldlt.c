/*
 * oracle/ref_wrap_ldlt.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Wrapper translation unit around the UNMODIFIED reference LU plugin.  The
 * reference file is #included where it lies under /root/reference (path given by
 * -DREF_LDLT_C on the command line, see oracle/Makefile); nothing is copied.
 * Including it gives this TU access to the reference's file-scope statics
 * (src/ipo/ldlt.c:108-120) so tests can read the symbolic arrays and the numeric
 * factor for bit-exact comparison, and can reset the process-global state so
 * that more than one LP can be run through the reference in one process.
 */
#include REF_LDLT_C

/* ---- read-only accessors for the reference's static factor object ---- */
int     ref_ldlt_dim(void)      { return lp ? lp->m + lp->n : 0; }
int     ref_ldlt_n(void)        { return lp ? lp->n : 0; }   /* ldlt-space n (=solver m) */
int     ref_ldlt_m(void)        { return lp ? lp->m : 0; }   /* ldlt-space m (=solver n) */
int    *ref_ldlt_perm(void)     { return perm; }
int    *ref_ldlt_iperm(void)    { return iperm; }
int    *ref_ldlt_kAAt(void)     { return kAAt; }
int    *ref_ldlt_iAAt(void)     { return iAAt; }
double *ref_ldlt_AAt(void)      { return AAt; }
double *ref_ldlt_diag(void)     { return diag; }
int    *ref_ldlt_mark(void)     { return mark; }
int     ref_ldlt_denwin(void)   { return denwin; }
int     ref_ldlt_pdf(void)      { return pdf; }
int     ref_ldlt_dense(void)    { return dense; }
int     ref_ldlt_ndep(void)     { return ndep; }
double  ref_ldlt_epsdiag(void)  { return epsdiag; }

/* one raw forward/diagonal/backward sweep on a permuted vector of length m+n
 * (src/ipo/ldlt.c:433-505); exposed for per-call parity of the solve kernels */
int ref_ldlt_rawsolve(double *zperm)
{
    return rawsolve(lp->m, lp->n, zperm);
}

/* Forget the current LP so that the next ldltfac() call redoes the symbolic
 * phase on new matrix pointers (the reference itself can never do this:
 * src/ipo/ldlt.c:140 tests lp==NULL only once per process). */
void ref_ldlt_reset(void)
{
    inv_clo();
    if (lp != NULL) {
        FREE(lp->Q); FREE(lp->iQ); FREE(lp->kQ);
        FREE(lp->bndmark); FREE(lp->rngmark);
        free(lp);
        lp = NULL;
    }
}
