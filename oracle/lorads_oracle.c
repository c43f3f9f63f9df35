/* oracle/lorads_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C, single-threaded CPU restatement of the LoRADS per-iteration hot path, written from the
 * reference's algorithm (each function cites the reference file:line it follows).  It is the checker the
 * GPU parity tests, __graft_entry__.smoke() and bench.py's cpu_baseline leg use when the compiled reference
 * (oracle/_ref) is not at hand; the product path (lorads_b200/) never links, loads or calls it.
 *
 * PINNING: the reference ships no tests or golden vectors.  This restatement is pinned
 * against the untouched reference compiled in oracle/_ref: tests/test_oracle_vs_reference.py compares every
 * function below with the reference's own functions on the same inputs, and tests/golden/ holds vectors
 * generated from the reference by tests/golden/make_golden.py.
 *
 * Conventions are the reference's: factors are column-major n x r (matElem[row + k*n]); cone data are the
 * SDPA reader's CSC arrays over the packed lower-triangular index with column 0 = C; the union pattern P is
 * sorted by (col,row); off-diagonal entries count twice in inner products.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int64_t n, r;
    /* coefficients as triplets, constraint c = 0 is C, c = i+1 is A_i (reader convention) */
    int64_t *beg;            /* m + 2 */
    int64_t *row, *col, *pos; /* per entry: coordinates and position in the scratch matrix */
    double *val;
    int dense;               /* dense scratch (lorads_sdp_conic.c:884-989) */
    int64_t np;              /* size of the scratch value arrays: |P| or n(n+1)/2 */
    int64_t *prow, *pcol;    /* pattern (sparse scratch only) */
    double *w_sum, *obj_sum; /* sdp_coeff_w_sum / sdp_obj_sum value arrays */
    double *full;            /* n*n scratch for the dense branch */
    int64_t nnz_coeff;       /* number of non-zero constraint matrices */
    int64_t rank_max;
    double *R, *U, *V, *G, *M2, *bl;
    double *cg_r, *cg_p, *cg_q, *cg_qn, *cg_Q;
    int64_t cg_iter;
    double *cv;              /* constrVal of the cone, full length m */
    double cNrm1, cNrm2sq, cNrmInf;
} orc_cone;

typedef struct {
    int64_t m, ncones;
    double *b, *lam, *s, *q1, *q2, *M1;
    orc_cone *K;
    /* L-BFGS ring, lorads_solver.c:470-497 */
    int64_t L, head, N;
    double **hs, **hy, *hbeta, *halpha, *Dtemp;
    double cObjNrm1, cObjNrm2, cObjNrmInf, bNrm1, bNrm2, bNrmInf, rho0;
    double pinf;
} orc_ctx;

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

static void unpack(int64_t n, int64_t p, int64_t *row, int64_t *col) {
    /* tsp_decompress, lorads_sparse_opts.c:37-52 */
    int64_t j = 0, thresh = n;
    while (p >= thresh) { j += 1; thresh += n - j; }
    *row = p - thresh + n;
    *col = j;
}

orc_ctx *orc_create(int64_t m, int64_t ncones, const int64_t *dims, const double *b) {
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->m = m; c->ncones = ncones;
    c->b = (double *)calloc(m, sizeof(double)); memcpy(c->b, b, sizeof(double) * m);
    c->lam = (double *)calloc(m, sizeof(double)); c->s = (double *)calloc(m, sizeof(double));
    c->q1 = (double *)calloc(m, sizeof(double)); c->q2 = (double *)calloc(m, sizeof(double));
    c->M1 = (double *)calloc(m, sizeof(double));
    c->K = (orc_cone *)calloc(ncones, sizeof(orc_cone));
    for (int64_t i = 0; i < ncones; ++i) c->K[i].n = dims[i];
    /* cal_sdp_const, lorads_solver.c:1059-1066: with -DUNDER_BLAS the reference indexes rowRHS by the 1-based
     * result of Fortran idamax_, i.e. it reads the element after the first largest |b_i| (clamped here). */
    int64_t arg = 0;
    double best = -1.0;
    for (int64_t i = 0; i < m; ++i) {
        c->bNrm1 += fabs(b[i]); c->bNrm2 += b[i] * b[i];
        if (fabs(b[i]) > best) { best = fabs(b[i]); arg = i; }
    }
    c->bNrmInf = fabs(b[arg + 1 < m ? arg + 1 : m - 1]);
    c->bNrm2 = sqrt(c->bNrm2);
    return c;
}

/* sdpDataMatSetData lorads_sdp_data.c:811-828 + AConePresolveData lorads_sdp_conic.c:868-1076 */
void orc_set_cone(orc_ctx *c, int64_t ic, const int64_t *beg, const int64_t *idx, const double *elem) {
    orc_cone *K = &c->K[ic];
    const int64_t m = c->m, n = K->n, nnz = beg[m + 1];
    const double packed = (double)(n * (n + 1) / 2);
    K->beg = (int64_t *)malloc(sizeof(int64_t) * (m + 2));
    memcpy(K->beg, beg, sizeof(int64_t) * (m + 2));
    K->row = (int64_t *)malloc(sizeof(int64_t) * (nnz + 1)); K->col = (int64_t *)malloc(sizeof(int64_t) * (nnz + 1));
    K->pos = (int64_t *)malloc(sizeof(int64_t) * (nnz + 1)); K->val = (double *)malloc(sizeof(double) * (nnz + 1));
    int dense = n < 20;
    K->nnz_coeff = 0;
    for (int64_t col = 0; col <= m; ++col) {
        int64_t a = beg[col], e = beg[col + 1];
        if (col > 0 && e > a) K->nnz_coeff++;
        if ((double)(e - a) > 0.1 * packed) dense = 1;           /* a dense coefficient forces the dense scratch */
        /* sort the column by packed index (dataMatCreateSparseImpl, lorads_sdp_data.c:106-108) */
        int64_t cnt = e - a;
        int64_t *ord = (int64_t *)malloc(sizeof(int64_t) * (cnt + 1));
        for (int64_t k = 0; k < cnt; ++k) ord[k] = k;
        for (int64_t k = 1; k < cnt; ++k) {                       /* insertion sort: columns are short or sorted */
            int64_t o = ord[k], j = k;
            while (j > 0 && idx[a + ord[j - 1]] > idx[a + o]) { ord[j] = ord[j - 1]; --j; }
            ord[j] = o;
        }
        for (int64_t k = 0; k < cnt; ++k) {
            int64_t p = idx[a + ord[k]];
            unpack(n, p, &K->row[a + k], &K->col[a + k]);
            K->val[a + k] = elem[a + ord[k]];
            K->pos[a + k] = p;                                    /* dense scratch: nnzIdx2ResIdx = packed index */
        }
        free(ord);
    }
    /* objective norms: dataMatSparseNrm1 / Nrm2Square / NrmInf, lorads_sdp_data.c:148-183 */
    for (int64_t k = beg[0]; k < beg[1]; ++k) {
        double v = K->val[k], a = fabs(v);
        int d = K->row[k] == K->col[k];
        K->cNrm1 += d ? a : 2 * a; K->cNrm2sq += d ? v * v : 2 * v * v;
        if (a > K->cNrmInf) K->cNrmInf = a;
    }
    if (!dense) {
        /* union pattern: unique (row,col) pairs sorted by (col,row), get_unique_tuples lorads_sdp_conic.c:795-824 */
        int64_t *keys = (int64_t *)malloc(sizeof(int64_t) * (nnz + 1));
        for (int64_t k = 0; k < nnz; ++k) keys[k] = K->col[k] * n + K->row[k];
        qsort(keys, nnz, sizeof(int64_t), cmp_i64);
        int64_t np = 0;
        for (int64_t k = 0; k < nnz; ++k)
            if (k == 0 || keys[k] != keys[k - 1]) keys[np++] = keys[k];
        if ((double)np / packed >= 0.1) dense = 1;
        else {
            K->np = np;
            K->prow = (int64_t *)malloc(sizeof(int64_t) * np); K->pcol = (int64_t *)malloc(sizeof(int64_t) * np);
            for (int64_t p = 0; p < np; ++p) { K->pcol[p] = keys[p] / n; K->prow[p] = keys[p] % n; }
            for (int64_t k = 0; k < nnz; ++k) {                   /* nnzIdx2ResIdx (hash lookup in the reference) */
                int64_t key = K->col[k] * n + K->row[k];
                int64_t *f = (int64_t *)bsearch(&key, keys, np, sizeof(int64_t), cmp_i64);
                K->pos[k] = f - keys;
            }
        }
        free(keys);
    }
    K->dense = dense;
    if (dense) { K->np = n * (n + 1) / 2; K->full = (double *)calloc(n * n, sizeof(double)); }
    K->w_sum = (double *)calloc(K->np, sizeof(double)); K->obj_sum = (double *)calloc(K->np, sizeof(double));
    K->cv = (double *)calloc(m, sizeof(double));
}

/* LORADSDetermineRank, lorads_solver.c:290-319 */
void orc_determine_rank(orc_ctx *c, double timesRank) {
    for (int64_t i = 0; i < c->ncones; ++i) {
        orc_cone *K = &c->K[i];
        int64_t cap = (int64_t)sqrt((double)(2 * K->nnz_coeff)) + 1;
        if (cap > K->n) cap = K->n;
        int64_t r;
        if (timesRank <= 1e-6) r = cap;
        else if (K->nnz_coeff / K->n >= 20 && K->n <= 400 && c->ncones <= 3) r = cap;
        else { double t = ceil(timesRank * log((double)K->n)); r = t < (double)cap ? (int64_t)t : cap; }
        K->r = r < 1 ? 1 : r;
        K->rank_max = cap;
    }
}

static void draw(double *x, int64_t n) {
    /* LORADS_RANDOM_rk_MAT, lorads_solver.c:361-370 */
    for (int64_t i = 0; i < n; ++i) { x[i] = (double)rand() / RAND_MAX; x[i] -= (double)rand() / RAND_MAX; }
}

/* LORADSInitALMVars + LORADSInitADMMVars + initial_solver_state, lorads_solver.c:406-498,580-675,1148-1170 */
void orc_init_vars(orc_ctx *c, int64_t lbfgsLen) {
    srand(925);
    c->N = 0;
    int64_t sumdim = 0;
    double n2 = 0;
    for (int64_t i = 0; i < c->ncones; ++i) {
        orc_cone *K = &c->K[i];
        int64_t sz = K->n * K->r;
        K->R = (double *)calloc(sz, sizeof(double)); draw(K->R, sz);
        K->G = (double *)calloc(sz, sizeof(double));
        c->N += sz; sumdim += K->n;
        c->cObjNrm1 += K->cNrm1; n2 += K->cNrm2sq;
        if (K->cNrmInf > c->cObjNrmInf) c->cObjNrmInf = K->cNrmInf;
    }
    c->cObjNrm2 = sqrt(n2);
    for (int64_t i = 0; i < c->ncones; ++i) {
        orc_cone *K = &c->K[i];
        int64_t sz = K->n * K->r;
        K->U = (double *)calloc(sz, sizeof(double)); draw(K->U, sz);
        K->V = (double *)calloc(sz, sizeof(double)); draw(K->V, sz);
        K->M2 = (double *)calloc(sz, sizeof(double)); K->bl = (double *)calloc(sz, sizeof(double));
        K->cg_r = (double *)calloc(sz, sizeof(double)); K->cg_p = (double *)calloc(sz, sizeof(double));
        K->cg_q = (double *)calloc(sz, sizeof(double)); K->cg_qn = (double *)calloc(sz, sizeof(double));
        K->cg_Q = (double *)calloc(sz, sizeof(double));
    }
    c->L = lbfgsLen; c->head = 0;
    c->hs = (double **)calloc(lbfgsLen, sizeof(double *)); c->hy = (double **)calloc(lbfgsLen, sizeof(double *));
    c->hbeta = (double *)calloc(lbfgsLen, sizeof(double)); c->halpha = (double *)calloc(lbfgsLen, sizeof(double));
    for (int64_t k = 0; k < lbfgsLen; ++k) { c->hs[k] = (double *)calloc(c->N, sizeof(double)); c->hy[k] = (double *)calloc(c->N, sizeof(double)); }
    c->Dtemp = (double *)calloc(c->N, sizeof(double));
    c->rho0 = 1.0 / sqrt((double)sumdim);
}

int64_t orc_info(orc_ctx *c, int what, int64_t ic) {
    orc_cone *K = &c->K[ic];
    switch (what) {
    case 0: return c->m; case 1: return c->ncones; case 2: return K->n; case 3: return K->r;
    case 4: return K->dense ? 0 : K->np; case 6: return K->dense; case 10: return K->rank_max;
    }
    return -1;
}
double orc_dinfo(orc_ctx *c, int what) {
    switch (what) {
    case 0: return c->cObjNrm1; case 1: return c->cObjNrm2; case 2: return c->cObjNrmInf;
    case 3: return c->bNrm1; case 4: return c->bNrm2; case 5: return c->bNrmInf; case 6: return c->rho0;
    }
    return NAN;
}
void orc_pattern(orc_ctx *c, int64_t ic, int64_t *rows, int64_t *cols) {
    orc_cone *K = &c->K[ic];
    if (K->dense) return;
    for (int64_t p = 0; p < K->np; ++p) { rows[p] = K->prow[p]; cols[p] = K->pcol[p]; }
}
static double *fac(orc_cone *K, char w) {
    switch (w) { case 'R': return K->R; case 'U': return K->U; case 'V': return K->V; case 'G': return K->G; case 'M': return K->M2; case 'B': return K->bl; }
    return NULL;
}
double *orc_factor_ptr(orc_ctx *c, char w, int64_t ic) { return fac(&c->K[ic], w); }
double *orc_vec_ptr(orc_ctx *c, char w) {
    switch (w) { case 'l': return c->lam; case 's': return c->s; case 'b': return c->b; case 'm': return c->M1; case 'q': return c->q1; case 'Q': return c->q2; }
    return NULL;
}

/* LORADSUVt, lorads_alg_common.c:21-68: z = (U V^T + V U^T)/2 on the scratch matrix */
static void uvt(orc_cone *K, double *z, const double *U, const double *V) {
    const int64_t n = K->n, r = K->r;
    if (!K->dense) {
        for (int64_t p = 0; p < K->np; ++p) {
            int64_t row = K->prow[p], col = K->pcol[p];
            double a = 0.0, b = 0.0;
            for (int64_t k = 0; k < r; ++k) a += U[row + k * n] * V[col + k * n];
            if (row != col) {
                for (int64_t k = 0; k < r; ++k) b += U[col + k * n] * V[row + k * n];
                z[p] = 0.5 * a;
                z[p] += 0.5 * b;
            } else z[p] = a;
        }
    } else {
        /* dsyr2k('L','N', alpha = 0.5) into the full matrix, then repack the lower triangle column by column */
        for (int64_t j = 0; j < n; ++j)
            for (int64_t i = j; i < n; ++i) {
                double a = 0.0;
                for (int64_t k = 0; k < r; ++k) a += U[i + k * n] * V[j + k * n] + V[i + k * n] * U[j + k * n];
                K->full[j * n + i] = 0.5 * a;
            }
        int64_t q = 0;
        for (int64_t j = 0; j < n; ++j)
            for (int64_t i = j; i < n; ++i) z[q++] = K->full[j * n + i];
    }
}

/* sparseAUV / denseAUV, lorads_sdp_data.c:524-567 (one coefficient against the scratch values) */
static double coeff_inner(orc_cone *K, int64_t col, const double *z) {
    double res = 0.0;
    for (int64_t k = K->beg[col]; k < K->beg[col + 1]; ++k) {
        double t = 2 * K->val[k] * z[K->pos[k]];
        res += t;
        if (K->row[k] == K->col[k]) res -= 0.5 * t;
    }
    return res;
}

/* coneAUV, lorads_sdp_conic.c:285-292 / 498-505 (result expanded to length m) */
static void cone_auv(orc_ctx *c, orc_cone *K, const double *z, double *out) {
    for (int64_t i = 0; i < c->m; ++i) out[i] = coeff_inner(K, i + 1, z);
}

void orc_auv(orc_ctx *c, int64_t ic, char u, char v, double *out) {
    orc_cone *K = &c->K[ic];
    uvt(K, K->w_sum, fac(K, u), fac(K, v));     /* LORADSInitConstrVal, lorads_alg_common.c:71-76 */
    cone_auv(c, K, K->w_sum, out);
}
double orc_obj_auv(orc_ctx *c, int64_t ic, char u, char v) {
    orc_cone *K = &c->K[ic];
    uvt(K, K->obj_sum, fac(K, u), fac(K, v));   /* LORADSInitConstrValObjVal, lorads_alg_common.c:97-100 */
    return coeff_inner(K, 0, K->obj_sum);
}

/* zeros + addObjCoeff + sdpDataWSum: lorads_sdp_conic.c:327,437,539,633; lorads_sdp_data.c:589-634 */
static void wsum(orc_ctx *c, orc_cone *K, double *S, const double *w, int addC) {
    memset(S, 0, sizeof(double) * K->np);
    if (addC) for (int64_t k = K->beg[0]; k < K->beg[1]; ++k) S[K->pos[k]] += 1.0 * K->val[k];
    for (int64_t i = 0; i < c->m; ++i)
        for (int64_t k = K->beg[i + 1]; k < K->beg[i + 2]; ++k) S[K->pos[k]] += w[i] * K->val[k];
}

/* mul_rk: dataMatSparseMultiRkMat lorads_sdp_data.c:491-504 / dataMatDenseMultiRkMat :646-671 */
static void mul_rk(orc_cone *K, const double *S, const double *X, double *Y) {
    const int64_t n = K->n, r = K->r;
    memset(Y, 0, sizeof(double) * n * r);
    if (!K->dense) {
        for (int64_t p = 0; p < K->np; ++p) {
            int64_t row = K->prow[p], col = K->pcol[p];
            for (int64_t k = 0; k < r; ++k) Y[row + k * n] += S[p] * X[col + k * n];
            if (row != col) for (int64_t k = 0; k < r; ++k) Y[col + k * n] += S[p] * X[row + k * n];
        }
    } else {
        int64_t q = 0;
        for (int64_t j = 0; j < n; ++j)
            for (int64_t i = j; i < n; ++i) { K->full[j * n + i] = S[q]; K->full[i * n + j] = S[q]; ++q; }
        for (int64_t k = 0; k < r; ++k)
            for (int64_t j = 0; j < n; ++j) {
                double x = X[j + k * n];
                for (int64_t i = 0; i < n; ++i) Y[i + k * n] += K->full[j * n + i] * x;
            }
    }
}

void orc_wsum_mulrk(orc_ctx *c, int64_t ic, const double *w, int addC, char x, double *out) {
    orc_cone *K = &c->K[ic];
    wsum(c, K, K->obj_sum, w, addC);
    mul_rk(K, K->obj_sum, fac(K, x), out);
}

/* ALMSetGrad / ALMCalGrad, lorads_alm.c:9-54 */
double orc_alm_cal_grad(orc_ctx *c, double rho) {
    double lag = 0.0;
    for (int64_t i = 0; i < c->m; ++i) c->M1[i] = (-c->lam[i] - rho * c->b[i]) + rho * c->s[i];
    for (int64_t ic = 0; ic < c->ncones; ++ic) {
        orc_cone *K = &c->K[ic];
        wsum(c, K, K->obj_sum, c->M1, 1);
        mul_rk(K, K->obj_sum, K->R, K->G);
        double nn = 0.0;
        for (int64_t q = 0; q < K->n * K->r; ++q) { K->G[q] *= 2.0; nn += K->G[q] * K->G[q]; }
        lag += nn;
    }
    return lag;
}

/* LORADSInitConstrValAll + LORADSInitConstrValSum, lorads_alg_common.c:78-84,134-142 */
static void constr_val_all(orc_ctx *c, char u, char v) {
    memset(c->s, 0, sizeof(double) * c->m);
    for (int64_t ic = 0; ic < c->ncones; ++ic) {
        orc_cone *K = &c->K[ic];
        uvt(K, K->w_sum, fac(K, u), fac(K, v));
        cone_auv(c, K, K->w_sum, K->cv);
        for (int64_t i = 0; i < c->m; ++i) c->s[i] += K->cv[i];
    }
}

double orc_alm_prepare(orc_ctx *c, double rho) {   /* ALG_START, lorads_alm.c:1004-1014 */
    constr_val_all(c, 'R', 'R');
    return orc_alm_cal_grad(c, rho);
}

/* linSysProduct / ADMMUpdateUVMvec, lorads_admm.c:356-391,421-426 */
static void cg_mvec(orc_ctx *c, orc_cone *K, const double *Vn, const double *x, double *res) {
    uvt(K, K->w_sum, x, Vn);
    cone_auv(c, K, K->w_sum, c->M1);            /* weight vector lives in M1temp */
    wsum(c, K, K->w_sum, c->M1, 0);
    mul_rk(K, K->w_sum, Vn, res);
    for (int64_t q = 0; q < K->n * K->r; ++q) res[q] += x[q];
}
void orc_cg_matvec(orc_ctx *c, int64_t ic, char noupd, const double *x, double *res) {
    cg_mvec(c, &c->K[ic], fac(&c->K[ic], noupd), x, res);
}

/* CGSolve, lorads_cgs.c:81-240 (restart whenever k % 20 == 0, k = 0 included; stop on ||r||_2 / ||b||_1) */
static void cg_solve(orc_ctx *c, orc_cone *K, const double *Vn, double *x, const double *b, double tol, int64_t maxit) {
    const int64_t nr = K->n * K->r;
    double *r = K->cg_r, *p = K->cg_p, *q = K->cg_q, *qn = K->cg_qn, *Q = K->cg_Q;
    double bn = 0.0, rn = 0.0, qTr, qTrNew, pTQ, alpha, beta;
    for (int64_t i = 0; i < nr; ++i) bn += fabs(b[i]);
    cg_mvec(c, K, Vn, x, r);
    for (int64_t i = 0; i < nr; ++i) { r[i] = -(r[i] - b[i]); rn += r[i] * r[i]; }
    if (sqrt(rn) / bn < tol) return;
    memcpy(p, r, sizeof(double) * nr); memcpy(q, r, sizeof(double) * nr);
    K->cg_iter = 0;
    for (int64_t k = 0; k < maxit; ++k) {
        K->cg_iter += 1;
        cg_mvec(c, K, Vn, p, Q);
        qTr = 0.0; pTQ = 0.0;
        for (int64_t i = 0; i < nr; ++i) { qTr += q[i] * r[i]; pTQ += p[i] * Q[i]; }
        alpha = qTr / pTQ;
        rn = 0.0;
        for (int64_t i = 0; i < nr; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Q[i]; rn += r[i] * r[i]; }
        if (sqrt(rn) / bn < tol) return;
        if (k % 20 == 0) {
            cg_mvec(c, K, Vn, x, r);
            qTr = 0.0;
            for (int64_t i = 0; i < nr; ++i) { r[i] = -(r[i] - b[i]); p[i] = r[i]; q[i] = r[i]; qTr += r[i] * r[i]; }
        }
        qTrNew = 0.0;
        for (int64_t i = 0; i < nr; ++i) { qn[i] = r[i]; qTrNew += r[i] * r[i]; }
        beta = qTrNew / qTr;
        for (int64_t i = 0; i < nr; ++i) { p[i] = beta * p[i] + r[i]; q[i] = qn[i]; }
    }
}

/* LORADSUpdateSDPVarOne, lorads_admm.c:428-480 */
int64_t orc_update_sdp_var_one(orc_ctx *c, int64_t ic, char upd, char noupd, double rho, double tol, int64_t maxit) {
    orc_cone *K = &c->K[ic];
    const double *Vn = fac(K, noupd);
    double *x = fac(K, upd);
    const int64_t nr = K->n * K->r;
    for (int64_t i = 0; i < c->m; ++i) c->M1[i] = ((-c->b[i] + c->s[i]) - K->cv[i]) * rho - c->lam[i];
    wsum(c, K, K->obj_sum, c->M1, 1);
    mul_rk(K, K->obj_sum, Vn, K->M2);
    for (int64_t q = 0; q < nr; ++q) { K->M2[q] -= rho * Vn[q]; K->bl[q] = (-1.0 / rho) * K->M2[q]; }
    cg_solve(c, K, Vn, x, K->bl, tol, maxit);
    return K->cg_iter;
}

/* LORADScubic_equation + ALMLineSearch, lorads_alm.c:102-228 */
static double root3(double x) { return x > 0 ? pow(x, 1.0 / 3) : -pow(-x, 1.0 / 3); }
static int64_t cubic(double a, double b, double c, double d, double *res) {
    double A = b * b - 3 * a * c, B = b * c - 9 * a * d, C = c * c - 3 * b * d, delta = B * B - 4 * A * C;
    res[0] = res[1] = res[2] = 0.0;
    if (A == 0 && B == 0) { if (-c / b > res[0]) res[0] = -c / b; return 1; }
    if (delta > 0) {
        double Y1 = A * b + 1.5 * a * (-B + sqrt(delta)), Y2 = A * b + 1.5 * a * (-B - sqrt(delta));
        double t = (-b - root3(Y1) - root3(Y2)) / 3 / a;
        if (t > res[0]) res[0] = t;
        return 1;
    }
    if (delta == 0 && A != 0 && B != 0) { double Kc = B / A; res[0] = -b / a + Kc; res[1] = -Kc / 2; return 2; }
    if (delta < 0) {
        double sqA = sqrt(A), T = (A * b - 1.5 * a * B) / (A * sqA), th = acos(T);
        double cs = cos(th / 3), sn = sqrt(3) * sin(th / 3);
        res[0] = (-b - 2 * sqA * cs) / 3 / a; res[1] = (-b + sqA * (cs + sn)) / 3 / a; res[2] = (-b + sqA * (cs - sn)) / 3 / a;
        return 3;
    }
    return 0;
}
static double quartic(double a, double b, double c, double d, double x) { return a * pow(x, 4) + b * pow(x, 3) + c * pow(x, 2) + d * x; }

static int64_t line_search(orc_ctx *c, double rho, double p1, double p2, double *q0, double *tau) {
    const int64_t m = c->m;
    double n2 = 0, d12 = 0, n1 = 0, d02 = 0, d01 = 0;
    for (int64_t i = 0; i < m; ++i) { n2 += c->q2[i] * c->q2[i]; d12 += c->q1[i] * c->q2[i]; }
    for (int64_t i = 0; i < m; ++i) q0[i] += (1 / rho) * c->lam[i];
    for (int64_t i = 0; i < m; ++i) { n1 += c->q1[i] * c->q1[i]; d02 += q0[i] * c->q2[i]; d01 += q0[i] * c->q1[i]; }
    double a = rho * n2 / 2, b = rho * d12, cc = p2 - rho * d02 + rho * n1 / 2, d = p1 - rho * d01;
    double roots[3];
    int64_t nroot = cubic(4 * a, 3 * b, 2 * cc, d, roots);
    double f0 = 0.0, f1 = quartic(a, b, cc, d, 1.0), fr[3] = {1e+30, 1e+30, 1e+30};
    if (nroot >= 1 && roots[0] > 1e-20 && roots[0] <= 1.0) fr[0] = quartic(a, b, cc, d, roots[0]);
    if (nroot >= 2 && roots[1] > 1e-20 && roots[1] <= 1.0) fr[1] = quartic(a, b, cc, d, roots[1]);
    if (nroot == 3 && roots[2] > 1e-20 && roots[2] <= 1.0) fr[2] = quartic(a, b, cc, d, roots[2]);
    double mn = f0 < f1 ? f0 : f1;
    for (int k = 0; k < 3; ++k) if (fr[k] < mn) mn = fr[k];
    if (fabs(mn - f0) < 1e-10) *tau = 0.0;
    if (fabs(mn - f1) < 1e-10) *tau = 1.0;
    for (int k = 0; k < 3; ++k) if (fabs(mn - fr[k]) < 1e-10) *tau = roots[k];
    return nroot;
}

/* LBFGSDirection + LBFGSDirectionUseGrad, lorads_alm.c:230-391,469-489 (direction stored in U) */
static void lbfgs_direction(orc_ctx *c, int64_t counter) {
    const int64_t N = c->N, L = c->L;
    double *D = c->Dtemp;
    int64_t off = 0;
    for (int64_t ic = 0; ic < c->ncones; ++ic) { int64_t sz = c->K[ic].n * c->K[ic].r; memcpy(D + off, c->K[ic].G, sizeof(double) * sz); off += sz; }
    if (counter > 0) {
        int64_t cnt = counter <= L - 1 ? counter : L;
        int64_t node = ((c->head - 1) % L + L) % L;
        for (int64_t k = 0; k < cnt; ++k) {
            double t = 0.0;
            for (int64_t i = 0; i < N; ++i) t += c->hs[node][i] * D[i];
            c->halpha[node] = c->hbeta[node] * t;
            for (int64_t i = 0; i < N; ++i) D[i] -= c->halpha[node] * c->hy[node][i];
            node = ((node - 1) % L + L) % L;
        }
        node = (node + 1) % L;
        for (int64_t k = 0; k < cnt; ++k) {
            double t = 0.0;
            for (int64_t i = 0; i < N; ++i) t += c->hy[node][i] * D[i];
            double w = c->halpha[node] - c->hbeta[node] * t;
            for (int64_t i = 0; i < N; ++i) D[i] += w * c->hs[node][i];
            node = (node + 1) % L;
        }
    }
    off = 0;
    double dg = 0.0;
    for (int64_t ic = 0; ic < c->ncones; ++ic) {
        orc_cone *K = &c->K[ic];
        int64_t sz = K->n * K->r;
        for (int64_t i = 0; i < sz; ++i) { K->U[i] = -D[off + i]; dg += K->U[i] * K->G[i]; }
        off += sz;
    }
    if (dg >= 0) for (int64_t ic = 0; ic < c->ncones; ++ic) { orc_cone *K = &c->K[ic]; for (int64_t i = 0; i < K->n * K->r; ++i) K->U[i] = -K->G[i]; }
}

/* one ALM inner iteration, lorads_alm.c:1073-1146. out: tau, lagNormSquare, pinf(1), p1, p2 */
int64_t orc_alm_inner_iter(orc_ctx *c, double rho, int64_t counter, double *out) {
    const int64_t m = c->m, L = c->L;
    double tau = 0.0, p1 = 0.0, p2 = 0.0;
    lbfgs_direction(c, counter);
    double *q0 = c->M1;
    for (int64_t i = 0; i < m; ++i) q0[i] = c->b[i] - c->s[i];
    /* ALMCalq12p12, lorads_alm.c:540-560 */
    memset(c->q1, 0, sizeof(double) * m); memset(c->q2, 0, sizeof(double) * m);
    double *tmp = (double *)malloc(sizeof(double) * m);
    for (int64_t ic = 0; ic < c->ncones; ++ic) {
        orc_cone *K = &c->K[ic];
        uvt(K, K->obj_sum, K->R, K->U);
        p1 += coeff_inner(K, 0, K->obj_sum);
        cone_auv(c, K, K->obj_sum, tmp);
        for (int64_t i = 0; i < m; ++i) c->q1[i] += tmp[i];
    }
    for (int64_t i = 0; i < m; ++i) c->q1[i] *= 2.0;
    p1 *= 2;
    for (int64_t ic = 0; ic < c->ncones; ++ic) {
        orc_cone *K = &c->K[ic];
        uvt(K, K->obj_sum, K->U, K->U);
        p2 += coeff_inner(K, 0, K->obj_sum);
        cone_auv(c, K, K->obj_sum, tmp);
        for (int64_t i = 0; i < m; ++i) c->q2[i] += tmp[i];
    }
    free(tmp);
    int64_t nroot = line_search(c, rho, p1, p2, q0, &tau);
    out[0] = tau; out[3] = p1; out[4] = p2;
    if (nroot == 0) return 0;
    /* SetyAsNegGrad :583, ALMupdateVar :619, constrValSum update :1123-1124, ALMCalGrad, setlbfgsHisTwo :657 */
    int64_t off = 0, h = c->head;
    for (int64_t ic = 0; ic < c->ncones; ++ic) {
        orc_cone *K = &c->K[ic];
        int64_t sz = K->n * K->r;
        for (int64_t i = 0; i < sz; ++i) { c->hy[h][off + i] = -K->G[i]; K->R[i] += tau * K->U[i]; }
        off += sz;
    }
    for (int64_t i = 0; i < m; ++i) { c->s[i] += tau * c->q1[i]; c->s[i] += tau * tau * c->q2[i]; }
    double lag = orc_alm_cal_grad(c, rho);
    off = 0;
    double ys = 0.0;
    for (int64_t ic = 0; ic < c->ncones; ++ic) {
        orc_cone *K = &c->K[ic];
        int64_t sz = K->n * K->r;
        for (int64_t i = 0; i < sz; ++i) { c->hs[h][off + i] = tau * K->U[i]; c->hy[h][off + i] += K->G[i]; ys += c->hy[h][off + i] * c->hs[h][off + i]; }
        off += sz;
    }
    c->hbeta[h] = 1.0 / ys;
    c->head = (h + 1) % L;
    /* primalInfeasibility, lorads_alg_common.c:250-258 */
    constr_val_all(c, 'R', 'R');
    double rs = 0.0;
    for (int64_t i = 0; i < m; ++i) { double d = c->b[i] - c->s[i]; rs += d * d; }
    c->pinf = sqrt(rs) / (1 + c->bNrm1);
    out[1] = lag; out[2] = c->pinf;
    return nroot;
}
