/*
 * oracle/kkt_oracle.c -- TEST INFRASTRUCTURE ONLY (see kkt_oracle.h).
 *
 * Single-thread plain-C restatement of the reference hot path.  Every function names the
 * reference lines it follows.  Arithmetic is written with the same expression shapes as the
 * reference so that, compiled WITHOUT FMA contraction (-ffp-contract=off, no -march=native),
 * it reproduces the reference bit for bit; that is what the golden-log test checks.
 */
#include "kkt_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define KMAX(a, b) ((a) > (b) ? (a) : (b))
#define KMIN(a, b) ((a) > (b) ? (b) : (a))
#define KABS(a) ((a) > 0 ? (a) : -(a))

static void *xmalloc(size_t nbytes)
{
    void *p = malloc(nbytes ? nbytes : 1);
    if (!p) { fprintf(stderr, "kkt_oracle: out of memory\n"); exit(1); }
    return p;
}
static void *xcalloc(size_t cnt, size_t sz)
{
    void *p = calloc(cnt ? cnt : 1, sz);
    if (!p) { fprintf(stderr, "kkt_oracle: out of memory\n"); exit(1); }
    return p;
}
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ======================================================================================
 * linalg.c
 * ==================================================================================== */

/* linalg.c:17-25 -- strict left-to-right sum */
double kko_dotprod(const double *x, const double *y, int n)
{
    double acc = 0.0;
    int i;
    for (i = 0; i < n; i++) acc += x[i] * y[i];
    return acc;
}

/* linalg.c:62-70 -- CSC scatter; y[r] receives its terms in ascending column order */
void kko_smx(int m, int n, const double *a, const int *ka, const int *ia,
             const double *x, double *y)
{
    int r, j, k;
    for (r = 0; r < m; r++) y[r] = 0.0;
    for (j = 0; j < n; j++)
        for (k = ka[j]; k < ka[j + 1]; k++) y[ia[k]] += a[k] * x[j];
}

/* linalg.c:75-103 -- counting-sort transpose, stable in the input column order */
void kko_atnum(int m, int n, const int *ka, const int *ia, const double *a,
               int *kat, int *iat, double *at)
{
    int *fill = (int *)xcalloc((size_t)m, sizeof(int));
    int r, j, k;
    for (k = 0; k < ka[n]; k++) fill[ia[k]]++;
    kat[0] = 0;
    for (r = 0; r < m; r++) { kat[r + 1] = kat[r] + fill[r]; fill[r] = 0; }
    for (j = 0; j < n; j++)
        for (k = ka[j]; k < ka[j + 1]; k++) {
            int dst = kat[ia[k]] + fill[ia[k]]++;
            iat[dst] = j;
            at[dst] = a[k];
        }
    free(fill);
}

/* linalg.c:108-116 -- note the macro semantics MAX(maxv, ABS(x)) with NaN: a NaN entry
 * replaces the running maximum, the next ordinary entry replaces the NaN again */
double kko_maxv(const double *x, int n)
{
    double best = 0.0;
    int i;
    for (i = 0; i < n; i++) best = KMAX(best, KABS(x[i]));
    return best;
}

/* ======================================================================================
 * ldlt.c : the factor object (statics of ldlt.c:108-120 gathered in a struct)
 * ==================================================================================== */

struct kko_factor {
    int m, n;                       /* ldlt-space: n "column" nodes, then m "row" nodes */
    const int *kA, *iA, *kAt, *iAt; /* retained caller pointers (ldlt.c:151-158) */
    const double *A, *At;
    int *perm, *iperm, *kAAt, *iAAt, *mark;
    double *AAt, *diag;
    double epssol, epsnum, epsdiag;
    int ndep, denwin, pdf, dense;
    int have_symbolic;
    int last_passes;
    /* solve() work vectors, ldlt.c:117-118 */
    double *y_k, *x_k, *r, *s, *z;
};

kko_factor *kko_create(void)
{
    kko_factor *F = (kko_factor *)xcalloc(1, sizeof(*F));
    return F;
}

void kko_destroy(kko_factor *F)
{
    if (!F) return;
    free(F->perm); free(F->iperm); free(F->kAAt); free(F->iAAt); free(F->mark);
    free(F->AAt); free(F->diag);
    free(F->y_k); free(F->x_k); free(F->r); free(F->s); free(F->z);
    free(F);
}

int kko_dim(const kko_factor *F) { return F->m + F->n; }
int kko_denwin(const kko_factor *F) { return F->denwin; }
int kko_pdf(const kko_factor *F) { return F->pdf; }
int kko_ndep(const kko_factor *F) { return F->ndep; }
double kko_epsdiag(const kko_factor *F) { return F->epsdiag; }
const int *kko_perm(const kko_factor *F) { return F->perm; }
const int *kko_iperm(const kko_factor *F) { return F->iperm; }
const int *kko_kAAt(const kko_factor *F) { return F->kAAt; }
const int *kko_iAAt(const kko_factor *F) { return F->iAAt; }
const double *kko_AAt(const kko_factor *F) { return F->AAt; }
const double *kko_diag(const kko_factor *F) { return F->diag; }
const int *kko_mark(const kko_factor *F) { return F->mark; }
int kko_last_passes(const kko_factor *F) { return F->last_passes; }

/* --------------------------------------------------------------------------------------
 * binary min-heap on keys, 1-based positions; ldlt.c:1305-1349 (the static swap-based pair,
 * NOT common/heap.c).  Ties: right child only if strictly smaller, move only if strictly out
 * of order.
 * ------------------------------------------------------------------------------------ */
typedef struct { int *key, *pos, *slot; int count; } kheap;  /* slot[1..count] -> node, pos[node] -> slot */

static void heap_exchange(kheap *h, int a, int b)
{
    int na = h->slot[a], nb = h->slot[b];
    h->slot[a] = nb; h->slot[b] = na;
    h->pos[nb] = a;  h->pos[na] = b;
}
static void heap_sink(kheap *h, int count, int cur)           /* hfall, ldlt.c:1305-1328 */
{
    int child = 2 * cur;
    while (child <= count) {
        if (child < count && h->key[h->slot[child + 1]] < h->key[h->slot[child]]) child++;
        if (h->key[h->slot[cur]] > h->key[h->slot[child]]) {
            heap_exchange(h, cur, child);
            cur = child;
            child = 2 * cur;
        } else break;
    }
}
static void heap_float(kheap *h, int cur)                     /* hrise, ldlt.c:1330-1349 */
{
    int parent = cur / 2;
    while (parent > 0) {
        if (h->key[h->slot[parent]] > h->key[h->slot[cur]]) {
            heap_exchange(h, cur, parent);
            cur = parent;
            parent = cur / 2;
        } else break;
    }
}

static int cmp_int(const void *a, const void *b)
{
    int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

/* --------------------------------------------------------------------------------------
 * Tiered minimum-degree ordering with mass elimination on the explicit-fill elimination
 * graph; ldlt.c:860-1262 (method=_MD only, which is all the ipo path uses).
 * adj[v] / deg[v] / cap[v] are the neighbour lists (ownership taken, freed here).
 * ------------------------------------------------------------------------------------ */
static void order_and_fill(kko_factor *F, int N, int *deg, int **adj, int *tier)
{
    int *perm = (int *)xmalloc(sizeof(int) * N), *iperm = (int *)xmalloc(sizeof(int) * N);
    int *rest = (int *)xmalloc(sizeof(int) * N);      /* dst[]: distinguishable neighbours */
    int *cap = (int *)xmalloc(sizeof(int) * N);       /* spc[] */
    int *stamp = (int *)xmalloc(sizeof(int) * N);     /* iwork[] tag marks */
    kheap h;
    int penalty, v, i, tag = 0, fill = 0, room;
    int *kL, *iL;

    penalty = (int)(1.0 * N);                          /* stablty*m, ldlt.c:889 */
    h.key = (int *)xmalloc(sizeof(int) * N);
    h.pos = (int *)xmalloc(sizeof(int) * N);
    h.slot = (int *)xmalloc(sizeof(int) * (N + 1));
    room = 0;
    for (v = 0; v < N; v++) room += deg[v];
    room /= 2;                                          /* aatnz, ldlt.c:917-920 */
    kL = (int *)xmalloc(sizeof(int) * (N + 1));
    iL = (int *)xmalloc(sizeof(int) * (size_t)KMAX(room, 1));
    for (v = 0; v < N; v++) { cap[v] = deg[v]; perm[v] = -1; iperm[v] = -1; stamp[v] = 0; }

    /* keys: degree, plus a penalty for tier 1; tier-0 nodes denser than `dense` are demoted
     * (ldlt.c:985-999) */
    for (v = 0; v < N; v++) h.key[v] = deg[v];
    for (v = 0; v < N; v++) {
        if (deg[v] > F->dense && tier[v] == 0) tier[v] = 1;
        h.key[v] += tier[v] * penalty;
    }
    /* bottom-up heapify in the reference's visiting order (ldlt.c:1004-1010) */
    h.count = N;
    for (v = N - 1; v >= 0; v--) {
        h.pos[v] = v + 1;
        h.slot[v + 1] = v;
        heap_sink(&h, h.count, v + 1);
    }

    i = 0; kL[0] = 0; F->denwin = N;
    while (i < N) {
        int piv = h.slot[1], d = deg[piv], *pl = adj[piv];
        int nrest = 0, i2 = i + 1, k, kk, ii, grp, need, dd;

        if (d >= N - 1 - i) F->denwin = i;             /* ldlt.c:1027 */
        perm[i] = piv; iperm[piv] = i;

        /* mass elimination: neighbours indistinguishable from the pivot (ldlt.c:1037-1054) */
        for (k = 0; k < d; k++) iperm[pl[k]] = i;
        for (k = 0; k < d; k++) {
            int u = pl[k];
            if (deg[u] == d && tier[u] == tier[piv]) {
                int *ul = adj[u];
                for (kk = 0; kk < d; kk++) if (iperm[ul[kk]] < i) break;
                if (kk == d) { perm[i2] = u; iperm[u] = i2; i2++; }
                else rest[nrest++] = u;
            } else rest[nrest++] = u;
        }
        grp = i2 - i;

        need = fill + (d * (d + 1) - (d - grp) * (d - grp + 1)) / 2;   /* ldlt.c:1060 */
        if (need > room) {
            room = KMAX(need, 2 * room);
            iL = (int *)realloc(iL, sizeof(int) * (size_t)room);
            if (!iL) { fprintf(stderr, "kkt_oracle: out of memory\n"); exit(1); }
        }

        /* emit the column structures of the group members (ldlt.c:1068-1088) */
        dd = d;
        for (ii = i; ii < i2; ii++) {
            int u = perm[ii], *ul = adj[u];
            kL[ii + 1] = kL[ii] + dd;
            for (k = 0; k < deg[u]; k++) {
                int w = ul[k], row = iperm[w];
                if (row > ii) iL[fill++] = w;
                else if (row == i && w != perm[i]) iL[fill++] = w;
            }
            dd--;
        }

        /* remove the pivot, then the other group members, from the remaining neighbours' lists,
         * preserving order (ldlt.c:1094-1120) */
        for (k = 0; k < nrest; k++) {
            int u = rest[k], *ul = adj[u], du;
            deg[u]--;
            du = deg[u];
            for (kk = 0; ul[kk] != piv; kk++) ;
            for (; kk < du; kk++) ul[kk] = ul[kk + 1];
        }
        if (i2 > i + 1) {
            for (k = 0; k < nrest; k++) {
                int u = rest[k], *ul = adj[u], du = deg[u], gone = 0;
                for (kk = 0; kk < du; kk++) {
                    if (iperm[ul[kk]] > i) gone++;
                    else ul[kk - gone] = ul[kk];
                }
                deg[u] -= gone;
            }
        }

        /* delete the group from the heap (ldlt.c:1122-1134) */
        for (ii = i; ii < i2; ii++) {
            int u = perm[ii], cur = h.pos[u], oldkey = h.key[h.slot[cur]];
            h.slot[cur] = h.slot[h.count];
            h.pos[h.slot[cur]] = cur;
            h.count--;
            if (oldkey < h.key[h.slot[cur]]) heap_sink(&h, h.count, cur);
            else heap_float(&h, cur);
        }

        /* make the remaining neighbours a clique: explicit fill edges (ldlt.c:1144-1201) */
        for (k = 0; k < nrest; k++) {
            int u = rest[k], du = deg[u], *ul = adj[u];
            tag++;
            for (kk = 0; kk < du; kk++) stamp[ul[kk]] = tag;
            for (kk = k + 1; kk < nrest; kk++) {
                int w = rest[kk];
                if (stamp[w] != tag) {
                    if (deg[u] >= cap[u]) {
                        cap[u] *= 2;
                        adj[u] = (int *)realloc(adj[u], sizeof(int) * (size_t)cap[u]);
                    }
                    adj[u][deg[u]++] = w;
                    if (deg[w] >= cap[w]) {
                        cap[w] *= 2;
                        adj[w] = (int *)realloc(adj[w], sizeof(int) * (size_t)cap[w]);
                    }
                    adj[w][deg[w]++] = u;
                }
            }
        }

        /* re-key the remaining neighbours: float first, then sink (ldlt.c:1206-1220) */
        for (k = 0; k < nrest; k++) {
            int u = rest[k];
            h.key[u] = deg[u];
            if (tier[u] != 0) h.key[u] += tier[u] * penalty;
            heap_float(&h, h.pos[u]);
            heap_sink(&h, h.count, h.pos[u]);
        }

        for (ii = i; ii < i2; ii++) { free(adj[perm[ii]]); adj[perm[ii]] = NULL; }
        i = i2;
    }

    /* map to new indices, rows ascending inside each column (ldlt.c:1236-1238; the reference's
     * K&R quicksort is a total order on distinct ints, any sort gives the same array) */
    for (i = 0; i < kL[N]; i++) iL[i] = iperm[iL[i]];
    for (i = 0; i < N; i++) qsort(iL + kL[i], (size_t)(kL[i + 1] - kL[i]), sizeof(int), cmp_int);

    F->perm = perm; F->iperm = iperm; F->kAAt = kL;
    F->iAAt = (int *)realloc(iL, sizeof(int) * (size_t)KMAX(kL[N], 1));
    F->AAt = (double *)xmalloc(sizeof(double) * (size_t)KMAX(kL[N], 1));
    F->diag = (double *)xmalloc(sizeof(double) * N);
    free(rest); free(cap); free(stamp); free(h.key); free(h.pos); free(h.slot);
}

/* --------------------------------------------------------------------------------------
 * inv_sym, ldlt.c:638-858, specialised to what ldltfac() feeds it: Q empty, every bndmark
 * BDD_BELOW, every rngmark INFINITE, lp->tier NULL, dense=_DENSE(-1), pdf=_UNSET.
 * ------------------------------------------------------------------------------------ */
static void symbolic(kko_factor *F)
{
    int m = F->m, n = F->n, N = m + n, i, j, k;
    int *deg = (int *)xmalloc(sizeof(int) * N);
    int **adj = (int **)xmalloc(sizeof(int *) * N);
    int *tier = (int *)xmalloc(sizeof(int) * N);
    double dens, fraction, pfillin, dfillin;

    /* fill-in estimates decide which side is eliminated first (ldlt.c:687-717) */
    fraction = 1.0e0;
    for (j = 0; j < n; j++) {
        dens = (double)(F->kA[j + 1] - F->kA[j]) / (m + 1);
        fraction = fraction * (1.0e0 - dens * dens);
    }
    pfillin = 0.5 * m * m * (1.0e0 - fraction);
    fraction = 1.0e0;
    for (i = 0; i < m; i++) {
        dens = (double)(F->kAt[i + 1] - F->kAt[i]) / (n + 1);
        fraction = fraction * (1.0e0 - dens * dens);
    }
    dfillin = 0.5 * n * n * (1.0e0 - fraction);
    F->pdf = (3 * pfillin <= dfillin) ? 1 /* _PRIMAL */ : 2 /* _DUAL */;

    /* adjacency of K: column nodes 0..n-1 see n+iA[k]; row nodes n..n+m-1 see iAt[k]
     * (ldlt.c:727-759) */
    for (j = 0; j < n; j++) {
        int ne = F->kA[j + 1] - F->kA[j];
        adj[j] = (int *)xmalloc(sizeof(int) * (size_t)ne);
        for (k = 0; k < ne; k++) adj[j][k] = n + F->iA[F->kA[j] + k];
        deg[j] = ne;
    }
    for (i = 0; i < m; i++) {
        int ne = F->kAt[i + 1] - F->kAt[i];
        adj[n + i] = (int *)xmalloc(sizeof(int) * (size_t)ne);
        for (k = 0; k < ne; k++) adj[n + i][k] = F->iAt[F->kAt[i] + k];
        deg[n + i] = ne;
    }

    /* tiers (ldlt.c:766-809): the favoured side is tier 0, the other tier 1 */
    for (j = 0; j < n; j++) tier[j] = (F->pdf == 1) ? 0 : 1;
    for (i = 0; i < m; i++) tier[n + i] = (F->pdf == 1) ? 1 : 0;

    /* dense-column threshold (ldlt.c:814-846): with n1==0 the histogram scan stops at degree 0,
     * so dense = (int)(3*1) = 3 whatever the matrix */
    F->dense = 3;

    F->mark = (int *)xmalloc(sizeof(int) * N);
    order_and_fill(F, N, deg, adj, tier);
    free(deg); free(adj); free(tier);
}

/* --------------------------------------------------------------------------------------
 * lltnum, ldlt.c:517-636 -- left-looking column LDL^T with link lists
 * ------------------------------------------------------------------------------------ */
static void numeric(kko_factor *F)
{
    int N = F->m + F->n, n = F->n, i, j, nextj, k, kk, kb, ke, row;
    const int *kL = F->kAAt, *iL = F->iAAt;
    double *L = F->AAt, *d = F->diag;
    double *acc = (double *)xcalloc((size_t)N, sizeof(double));
    int *cursor = (int *)xmalloc(sizeof(int) * N), *chain = (int *)xmalloc(sizeof(int) * N);
    double maxdiag = 0.0;

    for (i = 0; i < N; i++) chain[i] = -1;
    for (i = 0; i < N; i++) if (KABS(d[i]) > maxdiag) maxdiag = KABS(d[i]);
    F->ndep = 0;

    for (i = 0; i < N; i++) {
        double di = d[i];
        int sgn = F->perm[i] < n ? -1 : 1;
        for (j = chain[i]; j != -1; j = nextj) {
            double lij, lij_dj;
            nextj = chain[j];
            k = cursor[j];
            lij = L[k];
            lij_dj = lij * d[j];
            di -= lij * lij_dj;
            kb = k + 1; ke = kL[j + 1];
            if (kb < ke) {
                cursor[j] = kb;
                row = iL[kb];
                chain[j] = chain[row];
                chain[row] = j;
                if (j < F->denwin) {
                    for (kk = kb; kk < ke; kk++) acc[iL[kk]] += lij_dj * L[kk];
                } else {                                  /* contiguous rows (ldlt.c:584-591) */
                    double *p = &acc[row];
                    for (kk = kb; kk < ke; kk++) { *p += lij_dj * L[kk]; p++; }
                }
            }
        }
        kb = kL[i]; ke = kL[i + 1];
        for (kk = kb; kk < ke; kk++) L[kk] -= acc[iL[kk]];
        if (fabs(di) <= F->epsnum * maxdiag || F->mark[i] == 0) {   /* ldlt.c:600-614 */
            double maxoff = 0.0;
            F->ndep++;
            for (kk = kb; kk < ke; kk++) maxoff = KMAX(maxoff, KABS(L[kk]));
            if (maxoff < 1.0e+6 * 1.0e-8) F->mark[i] = 0;
            else di = sgn * 1.0e-8;
        }
        d[i] = di;
        if (kb < ke) {
            cursor[i] = kb;
            row = iL[kb];
            chain[i] = chain[row];
            chain[row] = i;
            for (kk = kb; kk < ke; kk++) {
                if (F->mark[i]) L[kk] /= di; else L[kk] = 0.0;
                acc[iL[kk]] = 0.0;
            }
        }
    }
    free(chain); free(cursor); free(acc);
}

/* ldltfac -> inv_num, ldlt.c:124-309 */
void kko_ldltfac(kko_factor *F, int m, int n, const int *kA, const int *iA, const double *A,
                 const double *dn, const double *dm,
                 const int *kAt, const int *iAt, const double *At)
{
    int i, j, k, N, *where;
    if (!F->have_symbolic) {
        F->m = m; F->n = n;
        F->kA = kA; F->iA = iA; F->A = A; F->kAt = kAt; F->iAt = iAt; F->At = At;
        F->epssol = 1.0e-6; F->epsnum = 0.0; F->epsdiag = 1.0e-14;      /* ldlt.c:212-219 */
        symbolic(F);
        F->have_symbolic = 1;
    }
    m = F->m; n = F->n; N = m + n;
    where = (int *)xmalloc(sizeof(int) * N);

    for (j = 0; j < n; j++) F->diag[F->iperm[j]] = -KMAX(dn[j], F->epsdiag);      /* :235 */
    for (i = 0; i < m; i++) F->diag[F->iperm[n + i]] = KMAX(dm[i], F->epsdiag);   /* :236 */

    /* strictly lower triangle of the permuted K into L's storage (ldlt.c:243-269) */
    for (j = 0; j < n; j++) {
        int col = F->iperm[j];
        for (k = F->kAAt[col]; k < F->kAAt[col + 1]; k++) { where[F->iAAt[k]] = k; F->AAt[k] = 0.0; }
        for (k = F->kA[j]; k < F->kA[j + 1]; k++) {
            int row = F->iperm[n + F->iA[k]];
            if (row > col) F->AAt[where[row]] = F->A[k];
        }
    }
    for (i = 0; i < m; i++) {
        int col = F->iperm[n + i];
        for (k = F->kAAt[col]; k < F->kAAt[col + 1]; k++) { where[F->iAAt[k]] = k; F->AAt[k] = 0.0; }
        for (k = F->kAt[i]; k < F->kAt[i + 1]; k++) {
            int row = F->iperm[F->iAt[k]];
            if (row > col) F->AAt[where[row]] = F->At[k];
        }
    }
    free(where);

    for (i = 0; i < N; i++) F->mark[i] = 1;
    numeric(F);

    {   /* diagonal-perturbation escalation, persists across calls (ldlt.c:293-306) */
        double mindiag = HUGE_VAL;
        for (i = 0; i < N; i++) if (KABS(F->diag[i]) < mindiag) mindiag = KABS(F->diag[i]);
        if (mindiag < 1.0e-14) F->epsdiag *= 10;
    }
}

/* rawsolve, ldlt.c:433-505 */
int kko_rawsolve(kko_factor *F, double *z)
{
    int N = F->m + F->n, i, k, consistent = 1;
    double eps = 0.0, beta;
    const int *kL = F->kAAt, *iL = F->iAAt, *mark = F->mark;
    const double *L = F->AAt, *d = F->diag;

    if (F->ndep) eps = F->epssol * kko_maxv(z, F->m);          /* sic: first m entries, :446 */

    for (i = 0; i < N; i++) {
        if (mark[i]) {
            beta = z[i];
            for (k = kL[i]; k < kL[i + 1]; k++) z[iL[k]] -= L[k] * beta;
        } else if (fabs(z[i]) > eps) consistent = 0;
        else z[i] = 0.0;
    }
    for (i = N - 1; i >= 0; i--) {
        if (mark[i]) z[i] = z[i] / d[i];
        else if (fabs(z[i]) > eps) consistent = 0;
        else z[i] = 0.0;
    }
    for (i = N - 1; i >= 0; i--) {
        if (mark[i]) {
            beta = z[i];
            for (k = kL[i]; k < kL[i + 1]; k++) beta -= L[k] * z[iL[k]];
            z[i] = beta;
        } else if (fabs(z[i]) > eps) consistent = 0;
        else z[i] = 0.0;
    }
    return consistent;
}

static int g_n_rawsolve;

/* forwardbackward -> solve, ldlt.c:311-425 (Q empty, so Qx == 0 and max*Qx[j] == 0) */
int kko_forwardbackward(kko_factor *F, const double *Dn, const double *Dm, double *c, double *b)
{
    int m = F->m, n = F->n, i, j, pass = 0, consistent = 1;
    double maxrs, oldmaxrs, maxbc;
    double *y_k, *x_k, *r, *s, *z;
    const int *ip = F->iperm;

    F->y_k = y_k = (double *)realloc(F->y_k, sizeof(double) * (size_t)KMAX(m, 1));
    F->x_k = x_k = (double *)realloc(F->x_k, sizeof(double) * (size_t)KMAX(n, 1));
    F->r = r = (double *)realloc(F->r, sizeof(double) * (size_t)KMAX(m, 1));
    F->s = s = (double *)realloc(F->s, sizeof(double) * (size_t)KMAX(n, 1));
    F->z = z = (double *)realloc(F->z, sizeof(double) * (size_t)KMAX(m + n, 1));

    maxbc = KMAX(kko_maxv(b, m), kko_maxv(c, n)) + 1;
    maxrs = HUGE_VAL;
    do {
        if (pass == 0) {
            for (j = 0; j < n; j++) z[ip[j]] = c[j];
            for (i = 0; i < m; i++) z[ip[n + i]] = b[i];
        } else {
            for (j = 0; j < n; j++) z[ip[j]] = s[j];
            for (i = 0; i < m; i++) z[ip[n + i]] = r[i];
        }
        consistent = kko_rawsolve(F, z);
        g_n_rawsolve++;
        if (pass == 0) {
            for (j = 0; j < n; j++) x_k[j] = z[ip[j]];
            for (i = 0; i < m; i++) y_k[i] = z[ip[n + i]];
        } else {
            for (j = 0; j < n; j++) x_k[j] = x_k[j] + z[ip[j]];
            for (i = 0; i < m; i++) y_k[i] = y_k[i] + z[ip[n + i]];
        }
        kko_smx(m, n, F->A, F->kA, F->iA, x_k, r);
        kko_smx(n, m, F->At, F->kAt, F->iAt, y_k, s);
        /* ldlt.c:394: s = c - (s - Dn*x - max*Qx) with max=1, Qx=0: the "- 1*0.0" is exact */
        for (j = 0; j < n; j++) s[j] = c[j] - (s[j] - Dn[j] * x_k[j] - 1 * 0.0);
        for (i = 0; i < m; i++) r[i] = b[i] - (r[i] + Dm[i] * y_k[i]);
        oldmaxrs = maxrs;
        maxrs = KMAX(kko_maxv(r, m), kko_maxv(s, n));
        pass++;
    } while (maxrs > 1.0e-10 * maxbc && maxrs < oldmaxrs / 2);

    if (maxrs > oldmaxrs && pass > 1) {
        for (j = 0; j < n; j++) x_k[j] = x_k[j] - z[ip[j]];
        for (i = 0; i < m; i++) y_k[i] = y_k[i] - z[ip[n + i]];
    }
    for (j = 0; j < n; j++) c[j] = x_k[j];
    for (i = 0; i < m; i++) b[i] = y_k[i];
    F->last_passes = pass;
    return consistent;
}

/* ======================================================================================
 * METHOD plugins
 * ==================================================================================== */

static struct {
    int iter;
    double *E, *D, *rhs_y, *rhs_x, *sol_y, *sol_x;
} g_cap = { -1, 0, 0, 0, 0, 0, 0 };
static double g_t_factor, g_t_solve;
static int g_n_factor, g_n_solve;

void kko_capture(int iter, double *E, double *D, double *rhs_y, double *rhs_x,
                 double *sol_y, double *sol_x)
{
    g_cap.iter = iter; g_cap.E = E; g_cap.D = D; g_cap.rhs_y = rhs_y; g_cap.rhs_x = rhs_x;
    g_cap.sol_y = sol_y; g_cap.sol_x = sol_x;
}
void kko_last_timing(double *t_factor, double *t_solve, int *n_factor, int *n_solve, int *n_raw)
{
    if (t_factor) *t_factor = g_t_factor;
    if (t_solve) *t_solve = g_t_solve;
    if (n_factor) *n_factor = g_n_factor;
    if (n_solve) *n_solve = g_n_solve;
    if (n_raw) *n_raw = g_n_rawsolve;
}
static int g_itnlim = 200;   /* MAX_ITER, hsd.c:25 / intpt.c:31 */
void kko_set_itnlim(int itnlim) { g_itnlim = itnlim > 0 ? itnlim : 200; }
static void timing_reset(void) { g_t_factor = g_t_solve = 0.0; g_n_factor = g_n_solve = g_n_rawsolve = 0; }

static void show_small_problem(int m, int n, const int *kA, const int *iA, const double *A,
                               const double *b, const double *c)
{   /* hsd.c:70-95 / intpt.c: "Verify input" dump for m<20 && n<20 */
    int i, j, k;
    double AA[20][20];
    for (j = 0; j < n; j++) for (i = 0; i < m; i++) AA[i][j] = 0;
    for (j = 0; j < n; j++) for (k = kA[j]; k < kA[j + 1]; k++) AA[iA[k]][j] = A[k];
    printf("A <= b: \n");
    for (i = 0; i < m; i++) {
        for (j = 0; j < n; j++) printf(" %5.1f", AA[i][j]);
        printf("<= %5.1f \n", b[i]);
    }
    printf("\n");
    printf("c: \n");
    for (j = 0; j < n; j++) printf(" %5.1f", c[j]);
    printf("\n");
}

/* hsd.c:27-311 -- homogeneous self-dual predictor/corrector */
int kko_solver_hsd(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c,
                   double f, double *x, double *y, double *w, double *z)
{
    double *dx, *dw, *dy, *dz, *fx, *fy, *gx, *gy, *rho, *sigma, *D, *E, *At;
    double phi, psi, dphi, dpsi, normr, norms, gamma, delta, mu, theta;
    double primal_obj, dual_obj, t0;
    int *iAt, *kAt, i, j, iter, status = 5;
    kko_factor *F = kko_create();

    timing_reset();
    dx = xmalloc(8 * (size_t)n); dz = xmalloc(8 * (size_t)n); sigma = xmalloc(8 * (size_t)n);
    D = xmalloc(8 * (size_t)n); fx = xmalloc(8 * (size_t)n); gx = xmalloc(8 * (size_t)n);
    dw = xmalloc(8 * (size_t)m); dy = xmalloc(8 * (size_t)m); rho = xmalloc(8 * (size_t)m);
    E = xmalloc(8 * (size_t)m); fy = xmalloc(8 * (size_t)m); gy = xmalloc(8 * (size_t)m);
    At = xmalloc(8 * (size_t)nz); iAt = xmalloc(4 * (size_t)nz); kAt = xmalloc(4 * ((size_t)m + 1));

    if (m < 20 && n < 20) show_small_problem(m, n, kA, iA, A, b, c);

    for (j = 0; j < n; j++) { x[j] = 1.0; z[j] = 1.0; }
    for (i = 0; i < m; i++) { w[i] = 1.0; y[i] = 1.0; }
    phi = 1.0; psi = 1.0;
    kko_atnum(m, n, kA, iA, A, kAt, iAt, At);

    printf("m = %d,n = %d,nz = %d\n", m, n, nz);
    printf(
"--------------------------------------------------------------------------\n"
"         |           Primal          |            Dual           |       |\n"
"  Iter   |  Obj Value       Infeas   |  Obj Value       Infeas   |  mu   |\n"
"- - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - \n");
    fflush(stdout);

    for (iter = 0; iter < g_itnlim; iter++) {
        mu = (kko_dotprod(z, x, n) + kko_dotprod(w, y, m) + phi * psi) / (n + m + 1);
        delta = (iter % 2 == 0) ? 0.0 : 1.0;
        primal_obj = kko_dotprod(c, x, n);
        dual_obj = kko_dotprod(b, y, m);

        if (mu < 1.0e-12) {                                            /* hsd.c:155-176 */
            if (phi > psi) { status = 0; break; }
            else if (dual_obj < 0.0) { status = 2; break; }
            else if (primal_obj > 0.0) { status = 4; break; }
            else { printf("Trouble in river city \n"); status = 4; break; }
        }

        kko_smx(m, n, A, kA, iA, x, rho);
        for (i = 0; i < m; i++) rho[i] = rho[i] - b[i] * phi + w[i];
        normr = sqrt(kko_dotprod(rho, rho, m)) / phi;
        for (i = 0; i < m; i++) rho[i] = -(1 - delta) * rho[i] + w[i] - delta * mu / y[i];

        kko_smx(n, m, At, kAt, iAt, y, sigma);
        for (j = 0; j < n; j++) sigma[j] = -sigma[j] + c[j] * phi + z[j];
        norms = sqrt(kko_dotprod(sigma, sigma, n)) / phi;
        for (j = 0; j < n; j++) sigma[j] = -(1 - delta) * sigma[j] + z[j] - delta * mu / x[j];

        gamma = -(1 - delta) * (dual_obj - primal_obj + psi) + psi - delta * mu / phi;

        printf("%8d   %14.7e  %8.1e    %14.7e  %8.1e  %8.1e \n",
               iter, primal_obj / phi + f, normr, dual_obj / phi + f, norms, mu);
        fflush(stdout);

        for (j = 0; j < n; j++) D[j] = z[j] / x[j];
        for (i = 0; i < m; i++) E[i] = w[i] / y[i];

        t0 = now_s();
        kko_ldltfac(F, n, m, kAt, iAt, At, E, D, kA, iA, A);          /* hsd.c:218 (swapped) */
        g_t_factor += now_s() - t0; g_n_factor++;

        for (j = 0; j < n; j++) fx[j] = -sigma[j];
        for (i = 0; i < m; i++) fy[i] = rho[i];
        if (iter == g_cap.iter) {
            memcpy(g_cap.E, E, 8 * (size_t)m); memcpy(g_cap.D, D, 8 * (size_t)n);
            memcpy(g_cap.rhs_y, fy, 8 * (size_t)m); memcpy(g_cap.rhs_x, fx, 8 * (size_t)n);
        }
        t0 = now_s();
        kko_forwardbackward(F, E, D, fy, fx);
        g_t_solve += now_s() - t0; g_n_solve++;
        if (iter == g_cap.iter) {
            memcpy(g_cap.sol_y, fy, 8 * (size_t)m); memcpy(g_cap.sol_x, fx, 8 * (size_t)n);
        }

        for (j = 0; j < n; j++) gx[j] = -c[j];
        for (i = 0; i < m; i++) gy[i] = -b[i];
        t0 = now_s();
        kko_forwardbackward(F, E, D, gy, gx);
        g_t_solve += now_s() - t0; g_n_solve++;

        dphi = (kko_dotprod(c, fx, n) - kko_dotprod(b, fy, m) + gamma) /
               (kko_dotprod(c, gx, n) - kko_dotprod(b, gy, m) - psi / phi);

        for (j = 0; j < n; j++) dx[j] = fx[j] - gx[j] * dphi;
        for (i = 0; i < m; i++) dy[i] = fy[i] - gy[i] * dphi;
        for (j = 0; j < n; j++) dz[j] = delta * mu / x[j] - z[j] - D[j] * dx[j];
        for (i = 0; i < m; i++) dw[i] = delta * mu / y[i] - w[i] - E[i] * dy[i];
        dpsi = delta * mu / phi - psi - (psi / phi) * dphi;

        theta = 0.0;
        for (j = 0; j < n; j++) {
            if (theta < -dx[j] / x[j]) theta = -dx[j] / x[j];
            if (theta < -dz[j] / z[j]) theta = -dz[j] / z[j];
        }
        for (i = 0; i < m; i++) {
            if (theta < -dy[i] / y[i]) theta = -dy[i] / y[i];
            if (theta < -dw[i] / w[i]) theta = -dw[i] / w[i];
        }
        if (theta < -dphi / phi) theta = -dphi / phi;
        if (theta < -dpsi / psi) theta = -dpsi / psi;
        theta = KMIN(0.95 / theta, 1.0);

        for (j = 0; j < n; j++) { x[j] = x[j] + theta * dx[j]; z[j] = z[j] + theta * dz[j]; }
        for (i = 0; i < m; i++) { y[i] = y[i] + theta * dy[i]; w[i] = w[i] + theta * dw[i]; }
        phi = phi + theta * dphi;
        psi = psi + theta * dpsi;
    }

    for (j = 0; j < n; j++) { x[j] /= phi; z[j] /= phi; }
    for (i = 0; i < m; i++) { y[i] /= phi; w[i] /= phi; }

    free(w); free(z);                                                   /* hsd.c:290-291 */
    free(dx); free(dw); free(dy); free(dz); free(rho); free(sigma); free(D); free(E);
    free(fx); free(fy); free(gx); free(gy); free(At); free(iAt); free(kAt);
    kko_destroy(F);
    return status;
}

/* intpt.c:33-261 -- primal-dual path following (float norms/objectives, intpt.c:47) */
int kko_solver_intpt(int m, int n, int nz, int *iA, int *kA, double *A, double *b, double *c,
                     double f, double *x, double *y, double *w, double *z)
{
    double *dx, *dw, *dy, *dz, *rho, *sigma, *D, *E, *At;
    double normr0, norms0, gamma, delta, mu, theta, r, t0;
    float primal_obj, dual_obj, normr, norms;
    int *iAt, *kAt, i, j, iter, status = 5;
    kko_factor *F = kko_create();

    timing_reset();
    dx = xmalloc(8 * (size_t)n); dz = xmalloc(8 * (size_t)n); sigma = xmalloc(8 * (size_t)n);
    D = xmalloc(8 * (size_t)n);
    dw = xmalloc(8 * (size_t)m); dy = xmalloc(8 * (size_t)m); rho = xmalloc(8 * (size_t)m);
    E = xmalloc(8 * (size_t)m);
    At = xmalloc(8 * (size_t)nz); iAt = xmalloc(4 * (size_t)nz); kAt = xmalloc(4 * ((size_t)m + 1));

    if (m < 20 && n < 20) show_small_problem(m, n, kA, iA, A, b, c);

    for (j = 0; j < n; j++) { x[j] = 1000.0; z[j] = 1000.0; }
    for (i = 0; i < m; i++) { w[i] = 1000.0; y[i] = 1000.0; }
    kko_atnum(m, n, kA, iA, A, kAt, iAt, At);
    delta = 0.02; r = 0.9;
    normr0 = HUGE_VAL; norms0 = HUGE_VAL;

    printf("m = %d,n = %d,nz = %d\n", m, n, nz);
    printf(
"------------------------------------------------------------------\n"
"         |           Primal          |            Dual           |\n"
"  Iter   |  Obj Value       Infeas   |  Obj Value       Infeas   |\n"
"- - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - - \n");
    fflush(stdout);

    for (iter = 0; iter < g_itnlim; iter++) {
        kko_smx(m, n, A, kA, iA, x, rho);
        for (i = 0; i < m; i++) rho[i] = b[i] - rho[i] - w[i];
        normr = sqrt(kko_dotprod(rho, rho, m));
        kko_smx(n, m, At, kAt, iAt, y, sigma);
        for (j = 0; j < n; j++) sigma[j] = c[j] - sigma[j] + z[j];
        norms = sqrt(kko_dotprod(sigma, sigma, n));

        gamma = kko_dotprod(z, x, n) + kko_dotprod(y, w, m);

        primal_obj = kko_dotprod(c, x, n) + f;
        dual_obj = kko_dotprod(b, y, m) + f;
        printf("%8d   %14.7e  %8.1e    %14.7e  %8.1e \n", iter, primal_obj, normr, dual_obj, norms);
        fflush(stdout);

        if (normr < 1.0e-6 && norms < 1.0e-6 && gamma < 1.0e-6) { status = 0; break; }
        if (normr > 10 * normr0) { status = 2; break; }
        if (norms > 10 * norms0) { status = 4; break; }

        mu = delta * gamma / (n + m);

        for (j = 0; j < n; j++) D[j] = z[j] / x[j];
        for (i = 0; i < m; i++) E[i] = w[i] / y[i];
        t0 = now_s();
        kko_ldltfac(F, n, m, kAt, iAt, At, E, D, kA, iA, A);          /* intpt.c:197 */
        g_t_factor += now_s() - t0; g_n_factor++;

        for (j = 0; j < n; j++) dx[j] = sigma[j] - z[j] + mu / x[j];
        for (i = 0; i < m; i++) dy[i] = rho[i] + w[i] - mu / y[i];
        if (iter == g_cap.iter) {
            memcpy(g_cap.E, E, 8 * (size_t)m); memcpy(g_cap.D, D, 8 * (size_t)n);
            memcpy(g_cap.rhs_y, dy, 8 * (size_t)m); memcpy(g_cap.rhs_x, dx, 8 * (size_t)n);
        }
        t0 = now_s();
        kko_forwardbackward(F, E, D, dy, dx);
        g_t_solve += now_s() - t0; g_n_solve++;
        if (iter == g_cap.iter) {
            memcpy(g_cap.sol_y, dy, 8 * (size_t)m); memcpy(g_cap.sol_x, dx, 8 * (size_t)n);
        }

        for (j = 0; j < n; j++) dz[j] = mu / x[j] - z[j] - D[j] * dx[j];
        for (i = 0; i < m; i++) dw[i] = mu / y[i] - w[i] - E[i] * dy[i];

        theta = 0.0;
        for (j = 0; j < n; j++) {
            if (theta < -dx[j] / x[j]) theta = -dx[j] / x[j];
            if (theta < -dz[j] / z[j]) theta = -dz[j] / z[j];
        }
        for (i = 0; i < m; i++) {
            if (theta < -dy[i] / y[i]) theta = -dy[i] / y[i];
            if (theta < -dw[i] / w[i]) theta = -dw[i] / w[i];
        }
        theta = KMIN(r / theta, 1.0);

        for (j = 0; j < n; j++) { x[j] = x[j] + theta * dx[j]; z[j] = z[j] + theta * dz[j]; }
        for (i = 0; i < m; i++) { y[i] = y[i] + theta * dy[i]; w[i] = w[i] + theta * dw[i]; }
        normr0 = normr;
        norms0 = norms;
    }

    free(w); free(z);                                                   /* intpt.c:244-245 */
    free(dx); free(dw); free(dy); free(dz); free(rho); free(sigma); free(D); free(E);
    free(At); free(iAt); free(kAt);
    kko_destroy(F);
    return status;
}
