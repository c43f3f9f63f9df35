/* oracle/ref_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin C glue compiled TOGETHER WITH the untouched reference sources (from
 * /root/reference/src_semi, never copied into this repo) into
 * oracle/_ref/libloradsref{32,64}.so.  It exposes the reference's own extern
 * functions to Python/ctypes with plain pointers and 64-bit sizes so that tests
 * can (a) validate the C restatement in oracle/lorads_oracle.c, (b) generate
 * the golden vectors under tests/golden/, and (c) time the reference's CPU
 * implementation of the hot path on the GPU box's host cores
 * (bench.py --impl reference, cpu_baseline.kind = "reference").
 *
 * Every function below only *calls* reference functions; the call order of
 * refh_open() follows the reference driver src_semi/main.c:249-304 and
 * refh_alm_inner_iter() follows src_semi/lorads_alg/lorads_alm.c:1073-1146.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "lorads_file_io.h"
#include "def_lorads_user_data.h"
#include "lorads_user_data.h"
#include "lorads_utils.h"
#include "def_lorads_solver.h"
#include "lorads_solver.h"
#include "lorads_alm.h"
#include "lorads_admm.h"
#include "lorads_alg_common.h"
#include "lorads_vec_opts.h"
#include "lorads_cgs.h"
#include "lorads_sdp_conic.h"

extern void LORADSInitConstrVal(lorads_sdp_cone *ACone, lorads_sdp_dense *U, lorads_sdp_dense *V, double *constrVal);
extern void ADMMUpdateUVMvec(void *pData, double *x, double *res);

typedef struct {
    lorads_params params;
    lorads_solver *S;
    lorads_int nConstrs, nBlks, nLpCols, nCols, nElem;
    lorads_int *BlkDims;
    double *rowRHS;
    lorads_int **coneMatBeg, **coneMatIdx;
    double **coneMatElem;
    lorads_int *LpMatBeg, *LpMatIdx;
    double *LpMatElem;
    user_data **SDPDatas;
    lorads_alm_state alm;
    lorads_admm_state admm;
    SDPConst sdpConst;
    lorads_func *f;            /* the reference's own function set: SDP only, or SDP + LP (lorads_solver.c:717-756) */
} refh_ctx;

static void default_params(lorads_params *p)
{   /* values of src_semi/main.c:19-43,236 */
    memset(p, 0, sizeof(*p));
    p->fname = "NULL";
    p->initRho = 0.0; p->rhoMax = 5000.0; p->rhoCellingALM = 1e+8;
    p->maxALMIter = 200; p->maxADMMIter = 10000; p->timesLogRank = 2.0;
    p->rhoFreq = 5; p->rhoFactor = 1.2; p->ALMRhoFactor = 2.0;
    p->phase1Tol = 1e-3; p->phase2Tol = 1e-5; p->timeSecLimit = 3600.0;
    p->heuristicFactor = 1.0; p->lbfgsListLength = 2; p->endTauTol = 1e-16;
    p->endALMSubTol = 1e-10; p->l2Rescaling = false; p->reoptLevel = 2;
    p->dyrankLevel = 2; p->highAccMode = false;
    p->rhoCellingADMM = p->rhoMax * 200;
}

refh_ctx *refh_open(const char *fname, double timesLogRank)
{
    refh_ctx *c = (refh_ctx *)calloc(1, sizeof(refh_ctx));
    default_params(&c->params);
    c->params.fname = strdup(fname);
    if (timesLogRank > 0) c->params.timesLogRank = timesLogRank;
    lorads_retcode rc = LReadSDPA(c->params.fname, &c->nConstrs, &c->nBlks, &c->BlkDims, &c->rowRHS,
                                  &c->coneMatBeg, &c->coneMatIdx, &c->coneMatElem, &c->nCols, &c->nLpCols,
                                  &c->LpMatBeg, &c->LpMatIdx, &c->LpMatElem, &c->nElem);
    if (rc != LORADS_RETCODE_OK) { free(c); return NULL; }
    lorads_solver *S;
    LORADS_INIT(S, lorads_solver, 1);
    LORADS_INIT(S->var, lorads_variable, 1);
    c->S = S;
    LORADSInitSolver(S, c->nConstrs, c->nBlks, c->BlkDims, c->nLpCols);
    LORADS_INIT(c->SDPDatas, user_data *, c->nBlks);
    LORADSSetDualObjective(S, c->rowRHS);
    LORADSInitConeData(S, c->SDPDatas, c->coneMatElem, c->coneMatBeg, c->coneMatIdx, c->BlkDims,
                       c->nConstrs, c->nBlks, c->nLpCols, c->LpMatBeg, c->LpMatIdx, c->LpMatElem);
    LORADSPreprocess(S, c->BlkDims);
    LORADSDetermineRank(S, c->BlkDims, c->params.timesLogRank);
    LORADSInitALMVars(S, S->var->rankElem, c->BlkDims, c->nBlks, c->nLpCols, c->params.lbfgsListLength);
    S->hisRecT = c->params.lbfgsListLength;
    LORADSInitADMMVars(S, S->var->rankElem, c->BlkDims, c->nBlks, c->nLpCols);
    initial_solver_state(&c->params, S, &c->alm, &c->admm, &c->sdpConst);
    LORADSInitFuncSet(&c->f, c->nLpCols);
    return c;
}

/* what: 0 nRows, 1 nCones, 2 blkDim, 3 rank, 4 |P| (0 if dense scratch), 5 cone type
 * (4 = dense-constraint cone, 5 = sparse-constraint cone), 6 scratch is dense (1/0),
 * 7 reader nnz of the cone, 8 nRowElem (sparse cone) or nRows, 9 nLpCols, 10 rank_max */
int64_t refh_info(refh_ctx *c, int what, int iCone)
{
    lorads_solver *S = c->S;
    switch (what) {
    case 0: return S->nRows;
    case 1: return S->nCones;
    case 2: return c->BlkDims[iCone];
    case 3: return S->var->R[iCone]->rank;
    case 4: {
        sdp_coeff *w = S->SDPCones[iCone]->sdp_coeff_w_sum;
        if (w->dataType == SDP_COEFF_SPARSE) return ((sdp_coeff_sparse *)w->dataMat)->nTriMatElem;
        return 0;
    }
    case 5: return (int64_t)S->SDPCones[iCone]->type;
    case 6: return S->SDPCones[iCone]->sdp_coeff_w_sum->dataType == SDP_COEFF_DENSE;
    case 7: return c->coneMatBeg[iCone][S->nRows + 1];
    case 8:
        if (S->SDPCones[iCone]->type == LORADS_CONETYPE_SPARSE_SDP)
            return ((lorads_cone_sdp_sparse *)S->SDPCones[iCone]->coneData)->nRowElem;
        return S->nRows;
    case 9: return S->nLpCols;
    case 10: return S->rank_max[iCone];
    }
    return -1;
}

double refh_dinfo(refh_ctx *c, int what)
{   /* 0 cObjNrm1, 1 cObjNrm2, 2 cObjNrmInf, 3 bNrm1, 4 bNrm2, 5 bNrmInf, 6 rho0, 7 pObj, 8 dObj,
       9 dimac[constrvio], 10 dimac[pdgap], 11 dimac[dualinf], 12 scaleObjHis */
    lorads_solver *S = c->S;
    switch (what) {
    case 0: return S->cObjNrm1; case 1: return S->cObjNrm2; case 2: return S->cObjNrmInf;
    case 3: return S->bRHSNrm1; case 4: return S->bRHSNrm2; case 5: return S->bRHSNrmInf;
    case 6: return c->alm.rho;  case 7: return S->pObjVal;  case 8: return S->dObjVal;
    case 9: return S->dimacError[LORADS_DIMAC_ERROR_CONSTRVIO_L1];
    case 10: return S->dimacError[LORADS_DIMAC_ERROR_PDGAP];
    case 11: return S->dimacError[LORADS_DIMAC_ERROR_DUALFEASIBLE_L1];
    case 12: return S->scaleObjHis;
    }
    return NAN;
}

/* Reader output of one cone as 64-bit arrays (CSC over packed lower-tri index). */
void refh_reader_csc(refh_ctx *c, int iCone, int64_t *beg, int64_t *idx, double *elem)
{
    lorads_int m = c->S->nRows;
    for (lorads_int i = 0; i < m + 2; ++i) beg[i] = c->coneMatBeg[iCone][i];
    lorads_int nnz = c->coneMatBeg[iCone][m + 1];
    for (lorads_int k = 0; k < nnz; ++k) { idx[k] = c->coneMatIdx[iCone][k]; elem[k] = c->coneMatElem[iCone][k]; }
}

void refh_pattern(refh_ctx *c, int iCone, int64_t *rows, int64_t *cols)
{
    sdp_coeff *w = c->S->SDPCones[iCone]->sdp_coeff_w_sum;
    if (w->dataType != SDP_COEFF_SPARSE) return;
    sdp_coeff_sparse *sp = (sdp_coeff_sparse *)w->dataMat;
    for (lorads_int k = 0; k < sp->nTriMatElem; ++k) { rows[k] = sp->triMatRow[k]; cols[k] = sp->triMatCol[k]; }
}

static lorads_sdp_dense *factor(refh_ctx *c, char which, int iCone)
{
    lorads_variable *v = c->S->var;
    switch (which) {
    case 'R': return v->R[iCone];
    case 'U': return v->U[iCone];
    case 'V': return v->V[iCone];
    case 'G': return v->Grad[iCone];
    case 'M': return v->M2temp[iCone];
    }
    return NULL;
}

double *refh_factor_ptr(refh_ctx *c, char which, int iCone) { return factor(c, which, iCone)->matElem; }

double *refh_vec_ptr(refh_ctx *c, char which)
{
    lorads_solver *S = c->S;
    switch (which) {
    case 'l': return S->var->dualVar;
    case 's': return S->var->constrValSum;
    case 'b': return S->rowRHS;
    case 'm': return S->var->M1temp;
    case 'q': return S->var->ARDSum;
    case 'Q': return S->var->ADDSum;
    case 'v': return S->constrVio;
    }
    return NULL;
}

double *refh_blinsys_ptr(refh_ctx *c, int iCone) { return c->S->var->bLinSys[iCone]; }

/* LP vectors of lorads_variable: 'R' rLp, 'U' uLp, 'V' vLp, 'G' gradLp (NULL without an LP block) */
double *refh_lp_ptr(refh_ctx *c, char which)
{
    lorads_variable *v = c->S->var;
    if (c->nLpCols <= 0) return NULL;
    switch (which) {
    case 'R': return v->rLp->matElem;
    case 'U': return v->uLp->matElem;
    case 'V': return v->vLp->matElem;
    case 'G': return v->gradLp->matElem;
    }
    return NULL;
}

/* One ADMM sweep over all blocks, LP columns included: LORADSUpdateSDPVar / LORADSUpdateSDPLPVar
 * (lorads_alg_common.c:187-249) after constrVal / constrValSum were initialised from (U, V). */
void refh_admm_init_constr(refh_ctx *c)
{
    lorads_solver *S = c->S;
    c->f->InitConstrValAll(S, S->var->uLp, S->var->vLp, S->var->U, S->var->V);
    c->f->InitConstrValSum(S);
}
void refh_admm_update_var(refh_ctx *c, double rho, double tol, int64_t maxit)
{
    c->f->admmUpdateVar(c->S, rho, tol, (lorads_int)maxit);
}

/* out (length nRows) = A(sym(U V^T)) of one cone  -- LORADSInitConstrVal, lorads_alg_common.c:71 */
void refh_auv(refh_ctx *c, int iCone, char u, char v, double *out)
{
    lorads_solver *S = c->S;
    double *val;
    lorads_vec *cv = S->var->constrVal[iCone];
    if (cv->type == LORADS_DENSE_VEC) val = ((dense_vec *)cv->data)->val; else val = ((sparse_vec *)cv->data)->val;
    LORADSInitConstrVal(S->SDPCones[iCone], factor(c, u, iCone), factor(c, v, iCone), val);
    memset(out, 0, sizeof(double) * S->nRows);
    double one = 1.0;
    cv->add(&one, cv->data, out);
}

/* <C, sym(U V^T)> of one cone -- LORADSUVt on sdp_obj_sum then objAUV, lorads_alg_common.c:97-100 */
double refh_obj_auv(refh_ctx *c, int iCone, char u, char v)
{
    lorads_sdp_cone *K = c->S->SDPCones[iCone];
    double res = 0.0;
    LORADSUVt(K->sdp_obj_sum, factor(c, u, iCone), factor(c, v, iCone));
    K->objAUV(K->coneData, factor(c, u, iCone), factor(c, v, iCone), &res, K->sdp_obj_sum);
    return res;
}

/* out (n x r col-major) = ([C] + sum_i w_i A_i) * X -- zeros/addObjCoeff/sdpDataWSum/mul_rk,
 * lorads_alm.c:29-34 */
void refh_wsum_mulrk(refh_ctx *c, int iCone, double *w, int addC, char x, double *out)
{
    lorads_sdp_cone *K = c->S->SDPCones[iCone];
    K->sdp_obj_sum->zeros(K->sdp_obj_sum->dataMat);
    if (addC) K->addObjCoeff(K->coneData, K->sdp_obj_sum);
    K->sdpDataWSum(K->coneData, w, K->sdp_obj_sum);
    K->sdp_obj_sum->mul_rk(K->sdp_obj_sum->dataMat, factor(c, x, iCone), out);
}

/* ALMCalGrad (lorads_alm.c:41) with the solver's current dualVar / constrValSum; returns sum ||G||^2 */
double refh_alm_cal_grad(refh_ctx *c, double rho)
{
    double lag = 0.0;
    c->f->ALMCalGrad(c->S, c->S->var->rLp, c->S->var->gradLp, c->S->var->R, c->S->var->Grad, &lag, rho);
    return lag;
}

/* res = x + (sum_i A(sym(x V^T))_i A_i) V -- ADMMUpdateUVMvec, lorads_admm.c:421. Clobbers M1temp. */
void refh_cg_matvec(refh_ctx *c, int iCone, char noUpdate, double *x, double *res)
{
    admmCG M;
    lorads_sdp_dense shell;
    memset(&shell, 0, sizeof(shell));
    M.ACone = c->S->SDPCones[iCone];
    M.noUpdateVar = factor(c, noUpdate, iCone);
    M.UpdateVarShell = &shell;
    M.weight = c->S->var->M1temp;
    ADMMUpdateUVMvec(&M, x, res);
}

/* LORADSUpdateSDPVarOne (lorads_admm.c:428): one ADMM block solve; returns CG iterations */
int64_t refh_update_sdp_var_one(refh_ctx *c, int iCone, char upd, char noupd, double rho, double tol, int64_t maxit)
{
    LORADSUpdateSDPVarOne(c->S, factor(c, upd, iCone), factor(c, noupd, iCone), iCone, rho, tol, (lorads_int)maxit);
    return c->S->CGLinsys[iCone]->iter;
}

/* state at label ALG_START of LORADS_ALMOptimize (lorads_alm.c:1004-1014) */
double refh_alm_prepare(refh_ctx *c, double rho)
{
    lorads_solver *S = c->S;
    c->f->InitConstrValAll(S, S->var->rLp, S->var->rLp, S->var->R, S->var->R);
    c->f->InitConstrValSum(S);
    double lag = 0.0;
    c->f->ALMCalGrad(S, S->var->rLp, S->var->gradLp, S->var->R, S->var->Grad, &lag, rho);
    return lag;
}

/* One ALM inner iteration, the body of the while loop at lorads_alm.c:1073-1146, built from the
 * reference's own functions.  out[0]=tau out[1]=lagNormSquare out[2]=pinf(1) out[3]=p1 out[4]=p2
 * Returns rootNum (0 = numerical failure). */
int64_t refh_alm_inner_iter(refh_ctx *c, double rho, int64_t lbfgsCounter, double *out)
{
    lorads_solver *S = c->S;
    lorads_int incx = 1;
    double minusOne = -1.0, tau = 0.0;
    c->f->LBFGSDirection(&c->params, S, S->lbfgsHis, S->var->gradLp, S->var->uLp, S->var->Grad, S->var->U, (lorads_int)lbfgsCounter);
    c->f->LBFGSDirUseGrad(S, S->var->uLp, S->var->gradLp, S->var->U, S->var->Grad);
    double *q0 = S->var->M1temp;
    LORADS_MEMCPY(q0, S->rowRHS, double, S->nRows);
    axpy(&(S->nRows), &minusOne, S->var->constrValSum, &incx, q0, &incx);
    double p12[2];
    c->f->ALMCalq12p12(S, S->var->rLp, S->var->uLp, S->var->R, S->var->U, S->var->ARDSum, S->var->ADDSum, p12);
    lorads_int rootNum = ALMLineSearch(rho, S->nRows, S->var->dualVar, p12[0], p12[1], q0, S->var->ARDSum, S->var->ADDSum, &tau);
    out[0] = tau; out[3] = p12[0]; out[4] = p12[1];
    if (rootNum == 0) return 0;
    c->f->setAsNegGrad(S, S->var->gradLp, S->var->Grad);
    c->f->ALMupdateVar(S, S->var->rLp, S->var->uLp, S->var->R, S->var->U, tau);
    double lag = 0.0, tauSquare = tau * tau;
    axpy(&(S->nRows), &tau, S->var->ARDSum, &incx, S->var->constrValSum, &incx);
    axpy(&(S->nRows), &tauSquare, S->var->ADDSum, &incx, S->var->constrValSum, &incx);
    c->f->ALMCalGrad(S, S->var->rLp, S->var->gradLp, S->var->R, S->var->Grad, &lag, rho);
    c->f->setlbfgsHisTwo(S, S->var->gradLp, S->var->uLp, S->var->Grad, S->var->U, tau);
    c->f->updateDimacsALM(S, S->var->R, S->var->R, S->var->rLp, S->var->rLp);
    out[1] = lag;
    out[2] = S->dimacError[LORADS_DIMAC_ERROR_CONSTRVIO_L1];
    return rootNum;
}

/* Runs `iters` inner iterations back to back and returns the wall time in seconds. */
double refh_time_alm_inner_iters(refh_ctx *c, double rho, int64_t iters, double *out)
{
    refh_alm_prepare(c, rho);
    double t0 = LUtilGetTimeStamp();
    for (int64_t k = 0; k < iters; ++k)
        if (refh_alm_inner_iter(c, rho, k, out) == 0) break;
    return LUtilGetTimeStamp() - t0;
}

double refh_now(void) { return LUtilGetTimeStamp(); }

/* Whole solve with the reference's own phase functions, in the order of main.c:321-487
 * (reoptLevel handling included).  out[0]=pObj out[1]=dObj out[2]=pinf(1) out[3]=gap
 * out[4]=dinf(1) out[5]=alm inner its out[6]=admm its out[7]=cg its out[8]=seconds out[9]=status */
void refh_solve(refh_ctx *c, double timeSecLimit, int reoptLevel, double *out)
{
    lorads_solver *S = c->S;
    lorads_params *P = &c->params;
    if (timeSecLimit > 0) P->timeSecLimit = timeSecLimit;
    P->reoptLevel = reoptLevel;
    double t0 = LUtilGetTimeStamp();
    double reopt_param = 5;
    lorads_int reopt_alm_iter = 3, reopt_admm_iter = 50, alm_reopt_min_iter = 3, admm_reopt_min_iter = 50;
    int admm_bad_iter_flag = 0;
    S->AStatus = LORADS_UNKNOWN;
    LORADS_ALMOptimize(P, S, &c->alm, P->maxALMIter, t0);
    if (LUtilGetTimeStamp() - t0 > P->timeSecLimit) { S->AStatus = LORADS_TIME_LIMIT; goto done; }
    LORADS_ALMtoADMM(S, P, &c->alm, &c->admm);
    if (LORADSADMMOptimize(P, S, &c->admm, P->maxADMMIter, t0) == RET_CODE_BAD_ITER) admm_bad_iter_flag = 1;
    int cnt = 0;
    if (P->reoptLevel >= 1) {
        while ((c->alm.primal_dual_gap > P->phase2Tol || c->alm.l_1_primal_infeasibility > P->phase2Tol) &&
               (c->admm.primal_dual_gap > P->phase2Tol || c->admm.l_1_primal_infeasibility > P->phase2Tol)) {
            if (cnt >= 1) break;
            reopt(P, S, &c->alm, &c->admm, &reopt_param, &alm_reopt_min_iter, &admm_reopt_min_iter, t0, &admm_bad_iter_flag, 1);
            cnt += 1;
            if (LUtilGetTimeStamp() - t0 > P->timeSecLimit) { S->AStatus = LORADS_TIME_LIMIT; goto done; }
        }
    }
    calculate_dual_infeasibility_solver(S);
    c->admm.l_1_dual_infeasibility = S->dimacError[LORADS_DIMAC_ERROR_DUALFEASIBLE_L1];
    c->admm.primal_dual_gap = S->dimacError[LORADS_DIMAC_ERROR_PDGAP];
    c->admm.l_1_primal_infeasibility = S->dimacError[LORADS_DIMAC_ERROR_CONSTRVIO_L1];
    if (P->reoptLevel >= 2) {
        int dual_cnt = 0;
        while (c->admm.l_1_dual_infeasibility > P->phase2Tol || c->admm.primal_dual_gap > P->phase2Tol ||
               c->admm.l_1_primal_infeasibility > P->phase2Tol) {
            if (dual_cnt >= 2) break;
            if (!P->highAccMode && c->admm.l_1_dual_infeasibility <= 5 * P->phase2Tol &&
                c->admm.primal_dual_gap <= 5 * P->phase2Tol && c->admm.l_1_primal_infeasibility <= P->phase2Tol) break;
            reopt(P, S, &c->alm, &c->admm, &reopt_param, &reopt_alm_iter, &reopt_admm_iter, t0, &admm_bad_iter_flag, 2);
            if (S->nLpCols > 0) averageUVLP(S->var->uLp, S->var->vLp, S->var->rLp);           /* main.c:438-447 */
            for (lorads_int i = 0; i < S->nCones; ++i) averageUV(S->var->U[i], S->var->V[i], S->var->R[i]);
            c->f->copyRtoV(S->var->rLp, S->var->vLp, S->var->R, S->var->V, S->nCones);
            calculate_dual_infeasibility_solver(S);
            c->admm.l_1_dual_infeasibility = S->dimacError[LORADS_DIMAC_ERROR_DUALFEASIBLE_L1];
            c->admm.primal_dual_gap = S->dimacError[LORADS_DIMAC_ERROR_PDGAP];
            c->admm.l_1_primal_infeasibility = S->dimacError[LORADS_DIMAC_ERROR_CONSTRVIO_L1];
            dual_cnt += 1;
            if (LUtilGetTimeStamp() - t0 > P->timeSecLimit) { S->AStatus = LORADS_TIME_LIMIT; goto done; }
        }
    }
    if (c->admm.l_1_dual_infeasibility <= 5 * P->phase2Tol && c->admm.primal_dual_gap <= 5 * P->phase2Tol &&
        c->admm.l_1_primal_infeasibility <= P->phase2Tol) S->AStatus = LORADS_PRIMAL_DUAL_OPTIMAL;
    else if (c->admm.primal_dual_gap <= 5 * P->phase2Tol && c->admm.l_1_primal_infeasibility <= P->phase2Tol)
        S->AStatus = LORADS_PRIMAL_OPTIMAL;
    else S->AStatus = LORADS_MAXITER;
done:
    out[0] = S->pObjVal; out[1] = S->dObjVal;
    out[2] = S->dimacError[LORADS_DIMAC_ERROR_CONSTRVIO_L1];
    out[3] = S->dimacError[LORADS_DIMAC_ERROR_PDGAP];
    out[4] = S->dimacError[LORADS_DIMAC_ERROR_DUALFEASIBLE_L1];
    out[5] = (double)c->alm.innerIter; out[6] = (double)c->admm.iter; out[7] = (double)S->cgIter;
    out[8] = LUtilGetTimeStamp() - t0; out[9] = (double)S->AStatus;
    fflush(stdout);
}
