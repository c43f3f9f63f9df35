/* integration/lorads_dropin.c -- the reference-side binding of the B200 kernel layer.
 *
 * LoRADS has no plugin / FFI layer: the seam is link time.  This translation unit is
 * what a LoRADS maintainer compiles INSTEAD OF every .c file of src_semi/data, src_semi/linalg and
 * src_semi/lorads_alg: it defines, with the reference's own signatures (it includes the reference's own
 * headers, which stay where they are), every function the UNMODIFIED driver src_semi/main.c calls from those
 * three subsystems, and forwards each one to the C ABI of liblorads_b200.so (include/lorads_b200.h).
 *
 *   gcc -DINT32 -I$REF -I$REF/data -I$REF/lorads_alg -I$REF/linalg -I$REF/io -Iinclude \
 *       $REF/main.c $REF/lorads_utils.c $REF/io/{lorads_cs,lorads_file_io,lorads_user_data}.c integration/lorads_dropin.c \
 *       -Llorads_b200 -llorads_b200 -lm -o lorads_gpu          (oracle/Makefile target `dropin`)
 *
 * Ownership follows the reference: the reader's arrays are borrowed (main.c frees them through
 * LORADSClearUsrData), everything allocated here is released by the LORADSDestroy* calls.  The fields of
 * lorads_solver / lorads_variable that main.c dereferences (main.c:279-298,329,404-410,439-447,510-511) are
 * kept valid; factor matrices live on the device, so the lorads_sdp_dense shells carry sizes but no matElem,
 * and averageUV / copyRtoV act on the device state.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lorads_utils.h"
#include "def_lorads_solver.h"
#include "lorads_solver.h"
#include "lorads_alm.h"
#include "lorads_admm.h"
#include "lorads_alg_common.h"

#include "lorads_b200.h"

int MAX_ALM_SUB_ITER;               /* the reference's global (lorads.h:66); the library keeps its own copy */

static lb2_solver *g_h = NULL;      /* the reference is single-instance and not re-entrant */

static void die(const char *what)
{
    fprintf(stderr, "lorads_b200 drop-in: %s failed: %s\n", what, lb2_last_error());
    exit(2);
}

static lb2_params to_lb2(const lorads_params *p)
{
    lb2_params q;
    lb2_default_params(&q);
    q.initRho = p->initRho; q.rhoMax = p->rhoMax; q.rhoCellingALM = p->rhoCellingALM; q.rhoCellingADMM = p->rhoCellingADMM;
    q.maxALMIter = p->maxALMIter; q.maxADMMIter = p->maxADMMIter; q.timesLogRank = p->timesLogRank; q.rhoFreq = p->rhoFreq;
    q.rhoFactor = p->rhoFactor; q.ALMRhoFactor = p->ALMRhoFactor; q.phase1Tol = p->phase1Tol; q.phase2Tol = p->phase2Tol;
    q.timeSecLimit = p->timeSecLimit; q.heuristicFactor = p->heuristicFactor; q.lbfgsListLength = p->lbfgsListLength;
    q.endTauTol = p->endTauTol; q.endALMSubTol = p->endALMSubTol; q.l2Rescaling = p->l2Rescaling; q.reoptLevel = p->reoptLevel;
    q.dyrankLevel = p->dyrankLevel; q.highAccMode = p->highAccMode; q.verbose = 1;
    return q;
}

static void push_state(const lorads_alm_state *a, const lorads_admm_state *d)
{
    double av[13], dv[13];
    if (a) {
        double t[13] = {(double)a->outerIter, (double)a->innerIter, a->rho, a->l_inf_primal_infeasibility, a->l_1_primal_infeasibility,
                        a->l_2_primal_infeasibility, a->primal_dual_gap, a->primal_objective_value, a->dual_objective_value,
                        a->l_inf_dual_infeasibility, a->l_1_dual_infeasibility, a->l_2_dual_infeasibility, a->tau};
        memcpy(av, t, sizeof(t));
    }
    if (d) {
        double t[13] = {(double)d->iter, (double)d->nBlks, (double)d->cg_iter, d->rho, d->l_1_dual_infeasibility, d->l_inf_dual_infeasibility,
                        d->l_1_primal_infeasibility, d->l_inf_primal_infeasibility, d->l_2_primal_infeasibility,
                        d->l_2_dual_infeasibility, d->primal_objective_value, d->dual_objective_value, d->primal_dual_gap};
        memcpy(dv, t, sizeof(t));
    }
    lb2_set_state(g_h, a ? av : NULL, d ? dv : NULL);
}

static void pull_state(lorads_solver *S, lorads_alm_state *a, lorads_admm_state *d)
{
    double av[13], dv[13];
    lb2_get_state(g_h, av, dv);
    if (a) {
        a->outerIter = (lorads_int)av[0]; a->innerIter = (lorads_int)av[1]; a->rho = av[2]; a->l_inf_primal_infeasibility = av[3];
        a->l_1_primal_infeasibility = av[4]; a->l_2_primal_infeasibility = av[5]; a->primal_dual_gap = av[6];
        a->primal_objective_value = av[7]; a->dual_objective_value = av[8]; a->l_inf_dual_infeasibility = av[9];
        a->l_1_dual_infeasibility = av[10]; a->l_2_dual_infeasibility = av[11]; a->tau = av[12];
    }
    if (d) {
        d->iter = (lorads_int)dv[0]; d->nBlks = (lorads_int)dv[1]; d->cg_iter = (lorads_int)dv[2]; d->rho = dv[3];
        d->l_1_dual_infeasibility = dv[4]; d->l_inf_dual_infeasibility = dv[5]; d->l_1_primal_infeasibility = dv[6];
        d->l_inf_primal_infeasibility = dv[7]; d->l_2_primal_infeasibility = dv[8]; d->l_2_dual_infeasibility = dv[9];
        d->primal_objective_value = dv[10]; d->dual_objective_value = dv[11]; d->primal_dual_gap = dv[12];
    }
    S->pObjVal = lb2_dinfo(g_h, 7); S->dObjVal = lb2_dinfo(g_h, 8);
    S->dimacError[LORADS_DIMAC_ERROR_CONSTRVIO_L1] = lb2_dinfo(g_h, 9);
    S->dimacError[LORADS_DIMAC_ERROR_PDGAP] = lb2_dinfo(g_h, 10);
    S->dimacError[LORADS_DIMAC_ERROR_DUALFEASIBLE_L1] = lb2_dinfo(g_h, 11);
    S->scaleObjHis = lb2_dinfo(g_h, 12);
    for (lorads_int i = 0; i < S->nCones; ++i) {
        lorads_int r = (lorads_int)lb2_info(g_h, 3, i);
        S->var->rankElem[i] = r;
        S->var->U[i]->rank = S->var->V[i]->rank = S->var->R[i]->rank = r;
    }
}

/* the one helper of src_semi/linalg the kept io/ code needs (lorads_sparse_opts.c:13-21): non-empty columns */
extern lorads_int csp_nnz_cols(lorads_int n, lorads_int *Ap)
{
    lorads_int nz = 0;
    for (lorads_int i = 0; i < n; ++i) nz += (Ap[i + 1] - Ap[i] > 0);
    return nz;
}

/* ---- lorads_solver.h ------------------------------------------------------------------------- */
extern void LORADSInitSolver(lorads_solver *S, lorads_int nRows, lorads_int nCones, lorads_int *blkDims, lorads_int nLpCols)
{
    S->nLpCols = nLpCols; S->nRows = nRows; S->nCones = nCones;
    LORADS_INIT(S->rowRHS, double, nRows);
    LORADS_INIT(S->dimacError, double, 5);
    LORADS_INIT(S->var->rankElem, lorads_int, nCones);
    (void)blkDims;
}

extern void LORADSSetDualObjective(lorads_solver *S, double *dObj) { LORADS_MEMCPY(S->rowRHS, dObj, double, S->nRows); }

extern void LORADSInitConeData(lorads_solver *S, user_data **SDPDatas, double **coneMatElem, lorads_int **coneMatBeg,
                               lorads_int **coneMatIdx, lorads_int *BlkDims, lorads_int nConstrs, lorads_int nBlks,
                               lorads_int nLpCols, lorads_int *LpMatBeg, lorads_int *LpMatIdx, double *LpMatElem)
{
    (void)SDPDatas;
    lb2_int *dims = (lb2_int *)malloc(sizeof(lb2_int) * nBlks);
    for (lorads_int i = 0; i < nBlks; ++i) dims[i] = BlkDims[i];
    const char *dev = getenv("LORADS_B200_DEVICE");
    if (lb2_create(&g_h, nConstrs, nBlks, dims, S->rowRHS, dev ? atoi(dev) : 0) != LB2_OK) die("lb2_create");
    free(dims);
    for (lorads_int k = 0; k < nBlks; ++k) {
        const lorads_int nnz = coneMatBeg[k][nConstrs + 1];
        lb2_int *beg = (lb2_int *)malloc(sizeof(lb2_int) * (nConstrs + 2));
        lb2_int *idx = (lb2_int *)malloc(sizeof(lb2_int) * (nnz + 1));
        for (lorads_int i = 0; i < nConstrs + 2; ++i) beg[i] = coneMatBeg[k][i];
        for (lorads_int i = 0; i < nnz; ++i) idx[i] = coneMatIdx[k][i];
        if (lb2_set_cone_data(g_h, k, beg, idx, coneMatElem[k]) != LB2_OK) die("lb2_set_cone_data");
        free(beg); free(idx);
    }
    if (nLpCols > 0) {
        /* LORADSSetLpCone (lorads_solver.c:135-137): the reader's LP arrays go to the device layer as they are */
        const lorads_int nnz = LpMatBeg[nConstrs + 1];
        lb2_int *beg = (lb2_int *)malloc(sizeof(lb2_int) * (nConstrs + 2));
        lb2_int *idx = (lb2_int *)malloc(sizeof(lb2_int) * (nnz + 1));
        for (lorads_int i = 0; i < nConstrs + 2; ++i) beg[i] = LpMatBeg[i];
        for (lorads_int i = 0; i < nnz; ++i) idx[i] = LpMatIdx[i];
        if (lb2_set_lp_data(g_h, nLpCols, beg, idx, LpMatElem) != LB2_OK) die("lb2_set_lp_data");
        free(beg); free(idx);
    }
}

extern void LORADSPreprocess(lorads_solver *S, lorads_int *BlkDims)
{
    (void)BlkDims;
    S->dTimeBegin = LUtilGetTimeStamp();
    printf("Pre-solver starts \n  Processing the cones \n");
    if (lb2_preprocess(g_h) != LB2_OK) die("lb2_preprocess");
    printf("  End preprocess \n");
}

extern void LORADSDetermineRank(lorads_solver *S, lorads_int *blkDims, double timesRank)
{
    (void)blkDims;
    if (lb2_determine_rank(g_h, timesRank) != LB2_OK) die("lb2_determine_rank");
    LORADS_INIT(S->rank_max, lorads_int, S->nCones);
    for (lorads_int i = 0; i < S->nCones; ++i) {
        S->var->rankElem[i] = (lorads_int)lb2_info(g_h, 3, i);
        S->rank_max[i] = (lorads_int)lb2_info(g_h, 9, i);
    }
}

static lorads_sdp_dense *shell(lorads_int n, lorads_int r)
{
    lorads_sdp_dense *x;
    LORADS_INIT(x, lorads_sdp_dense, 1);
    x->nRows = n; x->rank = r; x->matElem = NULL;     /* the factor itself is device resident */
    return x;
}

extern void LORADSInitALMVars(lorads_solver *S, lorads_int *rankElem, lorads_int *BlkDims, lorads_int nBlks, lorads_int nLpCols,
                              lorads_int lbfgsHis)
{
    LORADS_INIT(S->var->R, lorads_sdp_dense *, nBlks);
    LORADS_INIT(S->var->Grad, lorads_sdp_dense *, nBlks);
    LORADS_INIT(S->var->rLp, lorads_lp_dense, 1);
    LORADS_INIT(S->var->gradLp, lorads_lp_dense, 1);
    S->var->rLp->nCols = nLpCols; S->var->gradLp->nCols = nLpCols;     /* values are device resident (lb2_get_lp_vec) */
    for (lorads_int i = 0; i < nBlks; ++i) { S->var->R[i] = shell(BlkDims[i], rankElem[i]); S->var->Grad[i] = shell(BlkDims[i], rankElem[i]); }
    S->hisRecT = lbfgsHis;
}

extern void LORADSInitADMMVars(lorads_solver *S, lorads_int *rankElem, lorads_int *BlkDims, lorads_int nBlks, lorads_int nLpCols)
{
    LORADS_INIT(S->var->U, lorads_sdp_dense *, nBlks);
    LORADS_INIT(S->var->V, lorads_sdp_dense *, nBlks);
    LORADS_INIT(S->var->uLp, lorads_lp_dense, 1);
    LORADS_INIT(S->var->vLp, lorads_lp_dense, 1);
    S->var->uLp->nCols = nLpCols; S->var->vLp->nCols = nLpCols;
    for (lorads_int i = 0; i < nBlks; ++i) { S->var->U[i] = shell(BlkDims[i], rankElem[i]); S->var->V[i] = shell(BlkDims[i], rankElem[i]); }
}

extern void initial_solver_state(lorads_params *params, lorads_solver *S, lorads_alm_state *alm, lorads_admm_state *admm, SDPConst *c)
{
    /* the random start needs the L-BFGS length and initRho, both known only here */
    if (lb2_init_vars(g_h, params->lbfgsListLength, params->initRho) != LB2_OK) die("lb2_init_vars");
    S->cObjNrm1 = lb2_dinfo(g_h, 0); S->cObjNrm2 = lb2_dinfo(g_h, 1); S->cObjNrmInf = lb2_dinfo(g_h, 2);
    S->bRHSNrm1 = lb2_dinfo(g_h, 3); S->bRHSNrm2 = lb2_dinfo(g_h, 4); S->bRHSNrmInf = lb2_dinfo(g_h, 5);
    c->l_1_norm_c = S->cObjNrm1; c->l_2_norm_c = S->cObjNrm2; c->l_inf_norm_c = S->cObjNrmInf;
    c->l_1_norm_b = S->bRHSNrm1; c->l_2_norm_b = S->bRHSNrm2; c->l_inf_norm_b = S->bRHSNrmInf;
    memset(alm, 0, sizeof(*alm)); memset(admm, 0, sizeof(*admm));
    pull_state(S, alm, admm);
}

extern void LORADS_ALMtoADMM(lorads_solver *S, lorads_params *params, lorads_alm_state *alm, lorads_admm_state *admm)
{
    lb2_params q = to_lb2(params);
    push_state(alm, admm);
    if (lb2_alm_to_admm(g_h, &q) != LB2_OK) die("lb2_alm_to_admm");
    params->rhoMax = q.rhoMax;
    pull_state(S, alm, admm);
}

extern double reopt(lorads_params *params, lorads_solver *S, lorads_alm_state *alm, lorads_admm_state *admm, double *reopt_param,
                    lorads_int *reopt_alm_iter, lorads_int *reopt_admm_iter, double timeSolveStart, int *admm_bad_iter_flag,
                    int reopt_level)
{
    lb2_params q = to_lb2(params);
    lb2_int a = *reopt_alm_iter, b = *reopt_admm_iter;
    double sec = 0.0;
    push_state(alm, admm);
    if (lb2_reopt(g_h, &q, reopt_param, &a, &b, timeSolveStart, admm_bad_iter_flag, reopt_level, &sec) != LB2_OK) die("lb2_reopt");
    pull_state(S, alm, admm);
    return sec;
}

extern void calculate_dual_infeasibility_solver(lorads_solver *S)
{
    if (lb2_dual_infeasibility(g_h) != LB2_OK) die("lb2_dual_infeasibility");
    pull_state(S, NULL, NULL);
}

extern void printRes(double pObj, double dObj, double constrVio, double dualInfe, double pdgap, double constrVioInf, double dualInfeInf)
{
    printf("-----------------------------------------------------------------------\n");
    printf("Objective function Value are:\n");
    printf("\t 1.Primal Objective:            : %10.6e\n", pObj);
    printf("\t 2.Dual Objective:              : %10.6e\n", dObj);
    printf("Dimacs Error are:\n");
    printf("\t 1.Constraint Violation(1)      : %10.6e\n", constrVio);
    printf("\t 2.Dual Infeasibility(1)        : %10.6e\n", dualInfe);
    printf("\t 3.Primal Dual Gap              : %10.6e\n", pdgap);
    printf("\t 4.Primal Variable Semidefinite : %10.6e\n", 0.0);
    printf("\t 5.Constraint Violation(Inf)    : %10.6e\n", constrVioInf);
    printf("\t 6.Dual Infeasibility(Inf)      : %10.6e\n", dualInfeInf);
    printf("-----------------------------------------------------------------------\n");
}

extern void LORADSEndProgram(lorads_solver *S)
{
    static const char *why[] = {
        "End Program but the status is unknown, please notify the authors\n",
        "End Program due to reaching `Official terminate criteria`:\n",
        "End Program due to reaching `final terminate criteria`:\n",
        "End Program due to reaching `the maximum number of iterations`:\n",
        "End Program since time limit.\n"};
    printf("final rank: \n");
    for (lorads_int i = 0; i < S->nCones; ++i) printf("cone %lld rank: %lld", (long long)i, (long long)lb2_info(g_h, 3, i));
    printf("\n-----------------------------------------------------------------------\n");
    printf("%s", why[(int)S->AStatus]);
    printRes(S->pObjVal, S->dObjVal, S->dimacError[LORADS_DIMAC_ERROR_CONSTRVIO_L1], S->dimacError[LORADS_DIMAC_ERROR_DUALFEASIBLE_L1],
             S->dimacError[LORADS_DIMAC_ERROR_PDGAP],
             S->dimacError[LORADS_DIMAC_ERROR_CONSTRVIO_L1] * (1 + S->bRHSNrm1) / (1 + S->bRHSNrmInf),
             S->dimacError[LORADS_DIMAC_ERROR_DUALFEASIBLE_L1] * (1 + S->cObjNrm1) / (1 + S->cObjNrmInf));
    printf("GPU kernel launches: %lld\n", (long long)lb2_info(g_h, 11, 0));
}

extern void LORADSDestroyADMMVars(lorads_solver *S)
{
    for (lorads_int i = 0; i < S->nCones; ++i) { LORADS_FREE(S->var->U[i]); LORADS_FREE(S->var->V[i]); }
    LORADS_FREE(S->var->U); LORADS_FREE(S->var->V); LORADS_FREE(S->var->uLp); LORADS_FREE(S->var->vLp);
}

extern void LORADSDestroyALMVars(lorads_solver *S)
{
    for (lorads_int i = 0; i < S->nCones; ++i) { LORADS_FREE(S->var->R[i]); LORADS_FREE(S->var->Grad[i]); }
    LORADS_FREE(S->var->R); LORADS_FREE(S->var->Grad); LORADS_FREE(S->var->rLp); LORADS_FREE(S->var->gradLp);
}

extern void destroyPreprocess(lorads_solver *S) { LORADS_FREE(S->rank_max); }
extern void LORADSDestroyConeData(lorads_solver *S) { (void)S; }
extern void LORADSDestroySolver(lorads_solver *S)
{
    lb2_destroy(g_h);
    g_h = NULL;
    LORADS_FREE(S->rowRHS);
    LORADS_FREE(S->dimacError);
}

/* ---- lorads_alm.h / lorads_admm.h / lorads_alg_common.h ---------------------------------------- */
extern lorads_int LORADS_ALMOptimize(lorads_params *params, lorads_solver *S, lorads_alm_state *alm, double rho_update_factor,
                                     double timeSolveStart)
{
    (void)rho_update_factor;                      /* overwritten inside the reference as well, lorads_alm.c:1020 */
    lb2_params q = to_lb2(params);
    push_state(alm, NULL);
    int rc = lb2_alm_optimize(g_h, &q, timeSolveStart);
    if (rc < 0) die("lb2_alm_optimize");
    pull_state(S, alm, NULL);
    return RET_CODE_OK;
}

extern lorads_int LORADSADMMOptimize(lorads_params *params, lorads_solver *S, lorads_admm_state *admm, lorads_int iter_celling,
                                     double timeSolveStart)
{
    lb2_params q = to_lb2(params);
    push_state(NULL, admm);
    int rc = lb2_admm_optimize(g_h, &q, iter_celling, timeSolveStart);
    if (rc < 0) die("lb2_admm_optimize");
    pull_state(S, NULL, admm);
    return rc;
}

extern void averageUV(lorads_sdp_dense *U, lorads_sdp_dense *V, lorads_sdp_dense *UVavg)
{
    (void)U; (void)V; (void)UVavg;               /* device state: R = (U + V) / 2 for every cone (idempotent per call) */
    if (lb2_average_uv(g_h) != LB2_OK) die("lb2_average_uv");
}
/* lb2_average_uv / lb2_copy_r_to_v act on the whole [cones | LP] vector, so the LP variants need no work of their
 * own: main.c always calls them next to the cone versions (main.c:438-447) */
extern void averageUVLP(lorads_lp_dense *u, lorads_lp_dense *v, lorads_lp_dense *avg) { (void)u; (void)v; (void)avg; }
extern void copyRtoV(lorads_lp_dense *r, lorads_lp_dense *v, lorads_sdp_dense **R, lorads_sdp_dense **V, lorads_int nCones)
{
    (void)r; (void)v; (void)R; (void)V; (void)nCones;
    if (lb2_copy_r_to_v(g_h) != LB2_OK) die("lb2_copy_r_to_v");
}
extern void copyRtoVLP(lorads_lp_dense *r, lorads_lp_dense *v, lorads_sdp_dense **R, lorads_sdp_dense **V, lorads_int nCones)
{
    copyRtoV(r, v, R, V, nCones);
}
