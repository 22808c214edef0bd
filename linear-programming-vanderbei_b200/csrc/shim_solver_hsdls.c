/* libvbkkt_hsdls.so: exports the METHOD plugin symbol `solver` (reference src/ipo/hsdls.c:37) and
 * forwards to the device-resident implementation in libvbkkt.so.  Link this in place of hsdls.o. */
int vbk_solver_hsdls(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c, double f,
                     double *x, double *y, double *w, double *z);
int solver(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c, double f,
           double *x, double *y, double *w, double *z)
{
    return vbk_solver_hsdls(m, n, nz, iA, kA, A, b, c, f, x, y, w, z);
}
