/*
 * oracle/ref_dump_solver.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A METHOD plugin (same signature as src/ipo/hsd.c:27 / src/ipo/intpt.c:33) that
 * does not solve anything: it writes the arrays solvelp() hands to solver()
 * (src/common/solve.c:237) to the file named by $VBK_DUMP and returns.  Linked
 * with the reference's unmodified main.c/iolp.c/solve.c it turns any MPS file
 * into the exact solver-space LP (after bound shifts, row splitting, upper-bound
 * rows and the two atnum transposes) that the hot path sees.  Used by
 * tests/golden/make_golden.py to generate the committed fixtures.
 *
 * File layout (little endian): int32 m,n,nz; double f; int32 kA[n+1]; int32 iA[nz];
 * double A[nz]; double b[m]; double c[n].
 */
#include <stdio.h>
#include <stdlib.h>

int solver(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c,
           double f, double *x, double *y, double *w, double *z)
{
    const char *path = getenv("VBK_DUMP");
    FILE *fp;
    (void)x; (void)y;
    if (path == NULL || (fp = fopen(path, "wb")) == NULL) {
        fprintf(stderr, "ref_dump_solver: set VBK_DUMP to a writable path\n");
        exit(1);
    }
    fwrite(&m, sizeof(int), 1, fp);
    fwrite(&n, sizeof(int), 1, fp);
    fwrite(&nz, sizeof(int), 1, fp);
    fwrite(&f, sizeof(double), 1, fp);
    fwrite(kA, sizeof(int), (size_t)n + 1, fp);
    fwrite(iA, sizeof(int), (size_t)nz, fp);
    fwrite(A, sizeof(double), (size_t)nz, fp);
    fwrite(b, sizeof(double), (size_t)m, fp);
    fwrite(c, sizeof(double), (size_t)n, fp);
    fclose(fp);
    /* the real METHOD plugins free w and z (src/ipo/hsd.c:290-291); do the same */
    free(w); free(z);
    return 0;
}
