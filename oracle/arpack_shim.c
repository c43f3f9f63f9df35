/* oracle/arpack_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * The reference links the un-vendored, un-pinned Fortran ARPACK for one thing
 * only: lambda_min(C - A*(lambda)) in dual_infeasible()
 * (reference src_semi/data/lorads_sdp_conic.c:1286-1349, entry points declared at
 * src_semi/data/lorads_sdp_conic.h:31-38).  ARPACK is not in this image, so the
 * reference cannot link without a stand-in.  This file provides the two
 * reverse-communication entry points with exactly the calling sequence the
 * reference uses (bmat='I', which="SA", nev=1, mode 1) on top of an explicitly
 * restarted Lanczos iteration with full re-orthogonalisation.
 *
 * Consequence (stated in DESIGN.md): the "Dual Infeasibility" DIMACS lines of
 * the reference binary built here are NOT pinned against real ARPACK.
 *
 * Compile with the same integer-width define as the reference (-DINT32 or
 * -DMAC_INT64): `n` and `ldv` are lorads_int in the reference's declaration.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(MAC_INT64) || defined(UNIX_INT64)
typedef int64_t shim_int;
#else
typedef int shim_int;
#endif

typedef struct {
    shim_int n;
    int k;          /* Krylov dimension per restart cycle            */
    int j;          /* current Lanczos step inside the cycle         */
    int cycles;     /* completed restart cycles                      */
    int maxcycles;
    double tol;
    double *V;      /* n x (k+1) basis, own storage                  */
    double *alpha;  /* k                                             */
    double *beta;   /* k   (beta[j] couples v_j and v_{j+1})          */
    double theta;   /* current smallest Ritz value                   */
    int active;
} lanczos_state;

static lanczos_state S = {0};

/* Cyclic Jacobi for a small dense symmetric matrix (row-major k x k).
 * On exit a's diagonal holds the eigenvalues and q the eigenvectors (columns). */
static void jacobi_eig(int k, double *a, double *q)
{
    for (int i = 0; i < k; ++i)
        for (int j = 0; j < k; ++j)
            q[i * k + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < k; ++i)
            for (int j = i + 1; j < k; ++j)
                off += a[i * k + j] * a[i * k + j];
        if (off < 1e-300) break;
        for (int p = 0; p < k; ++p) {
            for (int r = p + 1; r < k; ++r) {
                double apr = a[p * k + r];
                if (fabs(apr) < 1e-300) continue;
                double app = a[p * k + p], arr = a[r * k + r];
                double tau = (arr - app) / (2.0 * apr);
                double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                for (int i = 0; i < k; ++i) {
                    double aip = a[i * k + p], air = a[i * k + r];
                    a[i * k + p] = c * aip - s * air;
                    a[i * k + r] = s * aip + c * air;
                }
                for (int i = 0; i < k; ++i) {
                    double api = a[p * k + i], ari = a[r * k + i];
                    a[p * k + i] = c * api - s * ari;
                    a[r * k + i] = s * api + c * ari;
                }
                for (int i = 0; i < k; ++i) {
                    double qip = q[i * k + p], qir = q[i * k + r];
                    q[i * k + p] = c * qip - s * qir;
                    q[i * k + r] = s * qip + c * qir;
                }
            }
        }
    }
}

static double vdot(shim_int n, const double *x, const double *y)
{
    double s = 0.0;
    for (shim_int i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

static void state_free(void)
{
    free(S.V); free(S.alpha); free(S.beta);
    memset(&S, 0, sizeof(S));
}

/* Smallest Ritz pair of the m x m tridiagonal built so far.
 * Returns theta, writes the Ritz vector coefficients into y (length m) and the
 * last component (for the residual estimate) into *last. */
static double ritz_smallest(int m, double *y, double *last)
{
    double *T = (double *)calloc((size_t)m * m, sizeof(double));
    double *Q = (double *)calloc((size_t)m * m, sizeof(double));
    for (int i = 0; i < m; ++i) {
        T[i * m + i] = S.alpha[i];
        if (i + 1 < m) { T[i * m + i + 1] = S.beta[i]; T[(i + 1) * m + i] = S.beta[i]; }
    }
    jacobi_eig(m, T, Q);
    int best = 0;
    for (int i = 1; i < m; ++i) if (T[i * m + i] < T[best * m + best]) best = i;
    double theta = T[best * m + best];
    for (int i = 0; i < m; ++i) y[i] = Q[i * m + best];
    *last = Q[(m - 1) * m + best];
    free(T); free(Q);
    return theta;
}

void dsaupd_(int *ido, char *bmat, shim_int *n, char *which, int *nev, double *tol, double *resid,
             int *ncv, double *v, shim_int *ldv, int *iparam, int *ipntr, double *workd,
             double *workl, int *lworkl, int *info)
{
    (void)bmat; (void)which; (void)nev; (void)v; (void)ldv; (void)workl; (void)lworkl;
    const shim_int N = *n;
    if (*ido == 0) {
        if (S.active) state_free();
        S.n = N;
        S.k = (*ncv < (int)N) ? *ncv : (int)N;
        if (S.k < 1) S.k = 1;
        S.maxcycles = iparam[2] > 0 ? iparam[2] : 300;
        S.tol = (*tol > 0) ? *tol : 2.2e-16;
        S.V = (double *)malloc(sizeof(double) * (size_t)N * (size_t)(S.k + 1));
        S.alpha = (double *)calloc((size_t)S.k, sizeof(double));
        S.beta = (double *)calloc((size_t)S.k, sizeof(double));
        S.active = 1;
        S.j = 0; S.cycles = 0;
        if (*info == 0) {   /* deterministic start vector, own generator (libc rand() untouched) */
            uint64_t st = 0x9E3779B97F4A7C15ull;
            for (shim_int i = 0; i < N; ++i) {
                st = st * 6364136223846793005ull + 1442695040888963407ull;
                resid[i] = ((double)(st >> 11) / 9007199254740992.0) - 0.5;
            }
        }
        double nr = sqrt(vdot(N, resid, resid));
        if (nr == 0.0) { resid[0] = 1.0; nr = 1.0; }
        for (shim_int i = 0; i < N; ++i) S.V[i] = resid[i] / nr;
        memcpy(workd, S.V, sizeof(double) * (size_t)N);
        ipntr[0] = 1; ipntr[1] = (int)N + 1;
        *ido = 1; *info = 0;
        return;
    }
    /* ido == 1 / -1 : workd[N..2N) now holds A * v_j */
    double *w = workd + N;
    double *vj = S.V + (size_t)S.j * N;
    double a = vdot(N, vj, w);
    S.alpha[S.j] = a;
    for (shim_int i = 0; i < N; ++i) w[i] -= a * vj[i];
    if (S.j > 0) {
        double b = S.beta[S.j - 1];
        const double *vp = S.V + (size_t)(S.j - 1) * N;
        for (shim_int i = 0; i < N; ++i) w[i] -= b * vp[i];
    }
    for (int pass = 0; pass < 2; ++pass)          /* full re-orthogonalisation */
        for (int l = 0; l <= S.j; ++l) {
            const double *vl = S.V + (size_t)l * N;
            double c = vdot(N, vl, w);
            for (shim_int i = 0; i < N; ++i) w[i] -= c * vl[i];
        }
    double b = sqrt(vdot(N, w, w));
    S.beta[S.j] = b;
    int m = S.j + 1;
    int breakdown = (b < 1e-14 * (fabs(a) + 1.0));
    if (m == S.k || breakdown) {
        double *y = (double *)malloc(sizeof(double) * (size_t)m);
        double last;
        S.theta = ritz_smallest(m, y, &last);
        double est = fabs(b * last);
        double scale = fabs(S.theta) > 2.2e-16 ? fabs(S.theta) : 2.2e-16;
        S.cycles += 1;
        if (breakdown || est <= S.tol * scale || S.cycles >= S.maxcycles || m >= (int)N) {
            free(y);
            iparam[4] = 1;      /* one converged Ritz value */
            *ido = 99; *info = 0;
            return;
        }
        /* explicit restart from the Ritz vector */
        double *x = resid;
        memset(x, 0, sizeof(double) * (size_t)N);
        for (int l = 0; l < m; ++l) {
            const double *vl = S.V + (size_t)l * N;
            for (shim_int i = 0; i < N; ++i) x[i] += y[l] * vl[i];
        }
        free(y);
        double nr = sqrt(vdot(N, x, x));
        for (shim_int i = 0; i < N; ++i) S.V[i] = x[i] / nr;
        S.j = 0;
        memcpy(workd, S.V, sizeof(double) * (size_t)N);
        ipntr[0] = 1; ipntr[1] = (int)N + 1;
        *ido = 1;
        return;
    }
    double *vn = S.V + (size_t)(S.j + 1) * N;
    for (shim_int i = 0; i < N; ++i) vn[i] = w[i] / b;
    S.j += 1;
    memcpy(workd, vn, sizeof(double) * (size_t)N);
    ipntr[0] = 1; ipntr[1] = (int)N + 1;
    *ido = 1;
}

void dseupd_(int *rvec, char *HowMny, int *select, double *d, double *z, shim_int *ldz, double *sigma,
             char *bmat, shim_int *n, char *which, int *nev, double *tol, double *resid,
             int *ncv, double *v, shim_int *ldv, int *iparam, int *ipntr, double *workd,
             double *workl, int *lworkl, int *info)
{
    (void)rvec; (void)HowMny; (void)select; (void)z; (void)ldz; (void)sigma; (void)bmat; (void)n;
    (void)which; (void)nev; (void)tol; (void)resid; (void)ncv; (void)v; (void)ldv; (void)iparam;
    (void)ipntr; (void)workd; (void)workl; (void)lworkl;
    d[0] = S.theta;
    *info = 0;
    if (S.active) state_free();
}
