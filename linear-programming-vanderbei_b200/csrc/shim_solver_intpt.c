/* libvbkkt_intpt.so: exports the METHOD plugin symbol `solver` (reference src/ipo/intpt.c:33) and
 * forwards to the device-resident implementation in libvbkkt.so.  Link this in place of intpt.o. */
int vbk_solver_intpt(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c, double f,
                     double *x, double *y, double *w, double *z);
int solver(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c, double f,
           double *x, double *y, double *w, double *z)
{
    return vbk_solver_intpt(m, n, nz, iA, kA, A, b, c, f, x, y, w, z);
}
