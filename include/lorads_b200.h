/* lorads_b200.h -- C ABI of the B200-native LoRADS low-rank kernel layer.
 *
 * Plain pointers and sizes only; no torch / C++ types.  All host matrices are
 * the reference's own formats:
 *   - a factor matrix (U, V, R, Grad, ...) is column-major n x r, exactly
 *     lorads_sdp_dense.matElem (reference src_semi/data/def_lorads_elements.h:29-33);
 *   - cone data are the SDPA reader's output arrays coneMatBeg / coneMatIdx /
 *     coneMatElem (CSC over the packed lower-triangular index, column 0 = C,
 *     column i = A_i; reference src_semi/io/lorads_file_io.c:325-340,
 *     src_semi/data/lorads_sdp_conic.c:160-168);
 *   - m-vectors (b, lambda, constrValSum, ...) are plain double[m].
 * Indices are 64-bit (lb2_int) regardless of the reference's lorads_int width.
 *
 * Every entry point returns 0 on success and a negative lb2 error code
 * otherwise (lb2_last_error() gives the text).  There is NO CPU fallback: when
 * no CUDA device is usable lb2_create() fails with LB2_ERR_CUDA.
 *
 * Each function names the reference function(s) it replaces.
 */
#ifndef LORADS_B200_H
#define LORADS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t lb2_int;
typedef struct lb2_solver lb2_solver;

#define LB2_OK            0
#define LB2_ERR_ARG      -1
#define LB2_ERR_CUDA     -2
#define LB2_ERR_STATE    -3
#define LB2_ERR_UNSUPPORTED -4
#define LB2_ERR_NUMERIC  -5

/* return codes of the phase functions, reference src_semi/lorads.h:62-65 */
#define LB2_RET_OK        0
#define LB2_RET_TIME_OUT  1
#define LB2_RET_NUM_ERR   4
#define LB2_RET_BAD_ITER  8

/* solver status, reference src_semi/lorads.h:45-51 */
#define LB2_STATUS_UNKNOWN             0
#define LB2_STATUS_PRIMAL_DUAL_OPTIMAL 1
#define LB2_STATUS_PRIMAL_OPTIMAL      2
#define LB2_STATUS_MAXITER             3
#define LB2_STATUS_TIME_LIMIT          4

/* Same fields and defaults as lorads_params (reference src_semi/lorads.h:82-105,
 * defaults src_semi/main.c:19-43,236); fname is not part of the kernel layer. */
typedef struct {
    double  initRho;
    double  rhoMax;
    double  rhoCellingALM;
    double  rhoCellingADMM;
    lb2_int maxALMIter;
    lb2_int maxADMMIter;
    double  timesLogRank;
    lb2_int rhoFreq;
    double  rhoFactor;
    double  ALMRhoFactor;
    double  phase1Tol;
    double  phase2Tol;
    double  timeSecLimit;
    double  heuristicFactor;
    lb2_int lbfgsListLength;
    double  endTauTol;
    double  endALMSubTol;
    int     l2Rescaling;
    lb2_int reoptLevel;
    lb2_int dyrankLevel;
    int     highAccMode;
    int     verbose;          /* 0 silent, 1 reference-style log lines */
} lb2_params;

/* What main.c reads back after a solve (reference src_semi/main.c:404-410,478-504
 * and printRes, src_semi/data/lorads_solver.c:908-922). */
typedef struct {
    double  pObj, dObj;
    double  pInfeasL1, dInfeasL1, pdGap;   /* DIMACS errors 1, 2, 3 */
    double  pInfeasInf, dInfeasInf;        /* DIMACS errors 5, 6   */
    lb2_int almOuterIter, almInnerIter, admmIter, cgIter;
    double  almRho, admmRho;
    double  solveSeconds;                  /* host wall clock of lb2_solve      */
    double  almSeconds, admmSeconds;
    int     status;                        /* LB2_STATUS_*                       */
    lb2_int finalRank0;                    /* rank of cone 0 at exit             */
    lb2_int kernelLaunches;                /* kernels launched by this library   */
} lb2_result;

const char *lb2_last_error(void);
const char *lb2_version(void);
void lb2_default_params(lb2_params *p);                       /* main.c:19-43 initCommandLineArgs */

/* ---- lifecycle ---------------------------------------------------------------------------- */

/* LORADSInitSolver + LORADSSetDualObjective (lorads_solver.c:17,125).  `device` is the CUDA ordinal. */
int lb2_create(lb2_solver **out, lb2_int nRows, lb2_int nCones, const lb2_int *blkDims,
               const double *rowRHS, int device);
/* LORADSInitConeData (lorads_solver.c:130) for one PSD block; arrays are borrowed during the call only. */
int lb2_set_cone_data(lb2_solver *s, lb2_int iCone, const lb2_int *coneMatBeg,
                      const lb2_int *coneMatIdx, const double *coneMatElem);
/* The LP cone of the problem (LORADSSetLpCone, lorads_lp_conic.c:217-237, called from LORADSInitConeData
 * lorads_solver.c:135-137): the reader's LpMatBeg / LpMatIdx / LpMatElem, i.e. a CSC matrix with m+1 columns
 * (column 0 = objective, column i+1 = constraint i) whose row index is the LP column.  Call before lb2_preprocess;
 * nLpCols = 0 clears it.  The LP variables ride behind the PSD factors in every factor-sized vector. */
int lb2_set_lp_data(lb2_solver *s, lb2_int nLpCols, const lb2_int *LpMatBeg, const lb2_int *LpMatIdx,
                    const double *LpMatElem);
/* LP vectors (nLpCols doubles): 'R' rLp, 'U' uLp, 'V' vLp, 'G' gradLp (def_lorads_solver.h lorads_variable),
 * 'x' the product represented in constrValLP, 'c' the (scaled) LP objective (get only). */
int lb2_get_lp_vec(lb2_solver *s, char which, double *out);
int lb2_set_lp_vec(lb2_solver *s, char which, const double *in);
/* LORADSPreprocess (lorads_solver.c:189): AConeProcData + AConePresolveData (lorads_sdp_conic.c:758,868):
 * coefficient classification, union pattern, index maps; then the device layouts are built and uploaded. */
int lb2_preprocess(lb2_solver *s);
/* LORADSDetermineRank (lorads_solver.c:290). */
int lb2_determine_rank(lb2_solver *s, double timesLogRank);
/* LORADSInitALMVars + LORADSInitADMMVars + initial_solver_state (lorads_solver.c:406,580,1148); the libc
 * rand() draw order of the reference (srand(925), R, then U, V per cone) is reproduced. */
int lb2_init_vars(lb2_solver *s, lb2_int lbfgsListLength, double initRho);
/* LORADSDestroy* (lorads_solver.c:81,265,280,500,677). */
void lb2_destroy(lb2_solver *s);

/* Multi-GPU: one process per GPU; `id128` is the 128-byte NCCL id produced on rank 0 (lb2_comm_unique_id) and broadcast
 * by the caller (e.g. torch.distributed over gloo/nccl).  Call after lb2_preprocess / lb2_determine_rank and before
 * lb2_init_vars.  The solve is sharded by factor columns (default) or, with LORADS_B200_SHARD=rows in the environment, by
 * row slabs / cone blocks (the independent cone loops of lorads_alg_common.c:78-84,105-112 and lorads_alm.c:46-53 go to
 * different ranks, a huge block is split by rows); one all-reduce of the m-vector per A() evaluation
 * (LORADSInitConstrValSum, lorads_alg_common.c:134-142) and one scalar all-reduce per dot table in both schemes.
 * An LP block together with sharding returns LB2_ERR_UNSUPPORTED. */
int lb2_comm_unique_id(void *id128);
int lb2_comm_init(lb2_solver *s, const void *id128, int rank, int world);

/* ---- SDPA .dat-s ingest (host only) -------------------------------------------------------------------- */
/* Fast replacement of LReadSDPA (lorads_file_io.c:21-418) with the same output convention: per PSD block a CSC
 * matrix with m+1 columns over the packed lower-triangular index, column 0 = objective (negated), entries
 * |v| < 1e-12 dropped, a trailing negative dimension = LP block (info 4 = nLpCols; block index k = number of PSD
 * blocks addresses it in info 3 / lb2_sdpa_get, output = LpMatBeg/Idx/Elem for lb2_set_lp_data).
 * The arrays can be passed straight to lb2_set_cone_data. */
typedef struct lb2_sdpa lb2_sdpa;
int lb2_read_sdpa(const char *path, lb2_sdpa **out);
/* Binary instance cache: writes the reader's output arrays as they are; lb2_read_sdpa recognises such a file by its
 * magic and loads it without parsing (SURVEY 8f-3 "binary instance format"). */
int lb2_sdpa_save(const lb2_sdpa *s, const char *path);
/* what: 0 nConstrs, 1 number of PSD blocks, 2 dimension of block k, 3 non-zeros of block k, 4 nLpCols, 5 nElems */
lb2_int lb2_sdpa_info(const lb2_sdpa *s, int what, lb2_int k);
/* copies block k (beg: m+2, idx/elem: nnz) and/or the right-hand side (m); pass k = -1 for the rhs only */
int lb2_sdpa_get(const lb2_sdpa *s, lb2_int k, lb2_int *coneMatBeg, lb2_int *coneMatIdx, double *coneMatElem, double *rowRHS);
void lb2_sdpa_free(lb2_sdpa *s);
const char *lb2_sdpa_last_error(void);

/* ---- host-only pieces (no GPU needed; used by the CPU test-suite) -------------------------------- */
/* The pre-solve of one cone on the host (what lb2_preprocess runs before uploading).
 * info[0]=|P| (or n(n+1)/2) info[1]=dense scratch? info[2]=dense-constraint cone? info[3]=active constraints
 * info[4]=nnzA info[5]=nnzC info[6]=non-zero constraint matrices info[7]=#rows split over A(UV^T) tiles
 * info[8]=1 if the objective is stored as c*ee^T + sparse remainder (info has room for 10 values).
 * rows/cols (capacity >= |P|, may be NULL) receive the union pattern of a sparse-scratch cone. */
int lb2_host_presolve(lb2_int n, lb2_int m, const lb2_int *coneMatBeg, const lb2_int *coneMatIdx,
                      const double *coneMatElem, lb2_int *info, lb2_int *rows, lb2_int *cols);
/* ALMLineSearch (lorads_alm.c:161-228) from its five m-vector sums {|q2|^2, q1.q2, q0.q2, |q1|^2, q0.q1};
 * returns rootNum, *tau is in/out exactly as in the reference. */
/* Host-only view of the device layouts the pre-solve builds for one cone (no CUDA involved): lets a test validate item
 * lists, the transposed constraint table and the adjacency against the reference WITHOUT a GPU.
 * lb2_layout_info what: 0 n, 1 pattern size, 2 dense scratch path, 3 active constraints, 4 nnzA, 5 items of A,
 *   6 items of [A;C], 7 adjacency entries, 8 tile size of A, 9 tile size of [A;C], 10 T entries, 11 split rows of [A;C]
 * lb2_layout_get which (dst sized by the caller from lb2_layout_info):
 *   int32: 0 act_idx, 1 P_row, 2 P_col, 3 A.ptr, 4 A.irow, 5 A.icol, 6 AC.ptr, 7 AC.irow, 8 AC.icol, 9 T_ptr, 10 T_con,
 *          11 adj_ptr, 12 adj_col, 13 adj_pos;  double: 20 A.coef, 21 AC.coef, 22 T_val, 23 C_onP, 24 {c_rank1} */
typedef struct lb2_layout lb2_layout;
int lb2_layout_build(lb2_int n, lb2_int m, const lb2_int *coneMatBeg, const lb2_int *coneMatIdx, const double *coneMatElem,
                     lb2_layout **out);
lb2_int lb2_layout_info(const lb2_layout *l, int what);
int lb2_layout_get(const lb2_layout *l, int which, void *dst);
void lb2_layout_free(lb2_layout *l);

lb2_int lb2_host_line_search(double rho, const double *sums, double p1, double p2, double *tau);
/* LORADSDetermineRank rule for one cone (lorads_solver.c:290-319); returns the rank, *rankMax the cap. */
lb2_int lb2_host_rank_rule(lb2_int n, lb2_int nNonzeroCoeff, lb2_int nCones, double timesLogRank, lb2_int *rankMax);

/* ---- queries ------------------------------------------------------------------------------ */
/* what: 0 nRows, 1 nCones, 2 blkDim, 3 rank, 4 |P| (pattern entries; n(n+1)/2 on the dense path),
 *       5 1 if the cone is a "dense-constraint cone" (> 30% of rows touch it, lorads_user_data.c:68-70)
 *       6 1 if the cone uses the dense scratch path, 7 nnz of the cone, 8 active constraints of the cone,
 *       9 rank_max, 10 local rank (column shard) , 11 kernel launches so far, 12 nnzA, 13 nnzC,
 *       14 adjacency entries (2|P| - #diag), 15 items of [A;C], 16 items of A, 17 padded row length ld,
 *       18 length N of the concatenated factor vectors (PSD cones only), 19 nLpCols,
 *       20 1 if the pipelined ALM step uses the three-output gather pass (single cone over all constraints,
 *          sparse scratch, no LP block) */
lb2_int lb2_info(const lb2_solver *s, int what, lb2_int iCone);
/* what: 0 cObjNrm1, 1 cObjNrm2, 2 cObjNrmInf, 3 bNrm1, 4 bNrm2, 5 bNrmInf, 6 rho0, 7 pObj, 8 dObj,
 *       9 pinf(1), 10 gap, 11 dinf(1), 12 scaleObjHis */
double  lb2_dinfo(const lb2_solver *s, int what);
/* union pattern P of a sparse-path cone (row >= col, sorted by (col,row)), AConePresolveData :974-1005 */
int lb2_get_pattern(const lb2_solver *s, lb2_int iCone, lb2_int *rows, lb2_int *cols);

/* ---- state transfer (host column-major <-> device layout) ----------------------------------- */
/* which: 'R' 'U' 'V' 'G'(rad) 'M'(2temp) 'B'(bLinSys) */
int lb2_set_factor(lb2_solver *s, char which, lb2_int iCone, const double *colMajor);
int lb2_get_factor(const lb2_solver *s, char which, lb2_int iCone, double *colMajor);
/* which: 'l' dualVar, 's' constrValSum, 'b' rowRHS, 'm' M1temp, 'q' ARDSum, 'Q' ADDSum, 'v' constrVio */
int lb2_set_vec(lb2_solver *s, char which, const double *v);
int lb2_get_vec(const lb2_solver *s, char which, double *v);

/* ---- hot-path operators (host buffers in/out; the parity tests call these) ------------------- */
/* constrVal = A(sym(U V^T)) of one cone, expanded to length nRows; if obj != NULL also <C, sym(U V^T)>.
 * Replaces LORADSUVt + coneAUV (+ objAUV): lorads_alg_common.c:21-76,97-103; lorads_sdp_conic.c:285,294,498,507;
 * lorads_sdp_data.c:524-587,698-715. */
int lb2_auv(lb2_solver *s, lb2_int iCone, char u, char v, double *constrVal, double *obj);
/* out (n x r col-major) = ([C] + sum_i w_i A_i) X.  Replaces zeros + addObjCoeff + sdpDataWSum + mul_rk:
 * lorads_alm.c:29-34; lorads_sdp_conic.c:327,437,539,633; lorads_sdp_data.c:491-504,589-671. */
int lb2_wsum_mulrk(lb2_solver *s, lb2_int iCone, const double *w, int addC, char x, double *out);
/* ALMCalGrad (lorads_alm.c:41) with the current dualVar / constrValSum; Grad stays on device ('G'). */
int lb2_alm_cal_grad(lb2_solver *s, double rho, double *lagNormSquare);
/* res = x + (sum_i A(sym(x V^T))_i A_i) V : ADMMUpdateUVMvec / linSysProduct (lorads_admm.c:376-426). */
int lb2_cg_matvec(lb2_solver *s, lb2_int iCone, char noUpdate, const double *x, double *res);
/* LORADSUpdateSDPVarOne (lorads_admm.c:428) = RHS build + CGSolve (lorads_cgs.c:81), all on device. */
int lb2_update_sdp_var_one(lb2_solver *s, lb2_int iCone, char upd, char noupd, double rho,
                           double cgTol, lb2_int cgMaxIter, lb2_int *cgIters);
/* Head of LORADSADMMOptimize (lorads_admm.c:47-48): constrVal of every block and constrValSum from (U, V), LP columns
 * included (LORADSInitConstrValAllLP / LORADSInitConstrValSumLP). */
int lb2_admm_init_constr(lb2_solver *s);
/* One Gauss-Seidel sweep over every block: LORADSUpdateSDPVar, and with an LP block LORADSUpdateSDPLPVar
 * (lorads_alg_common.c:187-249). */
int lb2_admm_update_var(lb2_solver *s, double rho, double cgTol, lb2_int cgMaxIter);
/* state at ALG_START of LORADS_ALMOptimize (lorads_alm.c:1004-1014); returns sum ||Grad||^2 */
int lb2_alm_prepare(lb2_solver *s, double rho, double *lagNormSquare);
/* one ALM inner iteration, the loop body lorads_alm.c:1073-1146.
 * out[0]=tau out[1]=lagNormSquare out[2]=pinf(1) out[3]=p1 out[4]=p2 ; *rootNum = 0 on numerical failure */
int lb2_alm_inner_iter(lb2_solver *s, double rho, lb2_int lbfgsCounter, double *out, lb2_int *rootNum);
/* `iters` inner iterations back to back, device resident; seconds = CUDA-event time of the loop. */
int lb2_time_alm_inner_iters(lb2_solver *s, double rho, lb2_int iters, double *out, double *seconds);

/* Host-buffer form of the ALM inner loop, the shape a LORADS_ALMOptimize call has for the reference driver:
 * R (all cones, column-major, concatenated) and lambda come from HOST memory, `iters` inner iterations run at
 * fixed rho starting from ALG_START, and R is copied back.  Used for the end-to-end (e2e) benchmark number.
 * out[0..4] as lb2_alm_inner_iter, out[5] = iterations actually done. */
int lb2_alm_run_host(lb2_solver *s, const double *R_in, const double *lambda_in, double rho, lb2_int iters,
                     double *R_out, double *out);
/* Times `reps` back-to-back launches of one hot kernel with CUDA events on the solver's stream (cone 0).
 * which: 0 A(UV^T) dual pass (R,D)+(D,D) with objective, 1 A(RR^T) constraints only, 2 S = C + A^*(w),
 *        3 Y = 2 S R (+ sum Y.Y), 4 one fused BLAS-1 pass (axpby + dot) over the factor vector,
 *        99 a one-thread kernel (what the timing method itself costs per launch),
 *        10 / 11 / 12 the collectives of a sharded step (all-gather of a factor vector, all-reduce of the three m-vectors of
 *        the fused A() pass, all-reduce of the scalar table); every rank has to make the same call.
 * flush_l2 != 0: a buffer three times the size of the L2 is read before every launch and each launch is timed
 * on its own (cold-cache figure); 0: back-to-back launches.  ms = average milliseconds per launch. */
int lb2_bench_kernel(lb2_solver *s, int which, lb2_int reps, int flush_l2, double *ms);

/* ---- phases (host control flow feeding on device-computed scalars) --------------------------- */
/* LORADS_ALMOptimize (lorads_alm.c:991) */
int lb2_alm_optimize(lb2_solver *s, lb2_params *p, double timeSolveStart);
/* LORADS_ALMtoADMM (lorads_solver.c:968) */
int lb2_alm_to_admm(lb2_solver *s, lb2_params *p);
/* LORADSADMMOptimize (lorads_admm.c:33); returns LB2_RET_* */
int lb2_admm_optimize(lb2_solver *s, lb2_params *p, lb2_int iterCelling, double timeSolveStart);
/* calculate_dual_infeasibility_solver (lorads_solver.c:1007) */
int lb2_dual_infeasibility(lb2_solver *s);
/* The whole driver sequence of main.c:321-487 (ALM, ALM->ADMM, ADMM, reopt rounds, dual infeasibility). */
int lb2_solve(lb2_solver *s, lb2_params *p, lb2_result *res);
/* reopt (lorads_solver.c:1075-1117): rescale C and lambda by *reoptParam, re-enter ALM (reopt variant) and ADMM. */
int lb2_reopt(lb2_solver *s, lb2_params *p, double *reoptParam, lb2_int *reoptAlmIter, lb2_int *reoptAdmmIter,
              double timeSolveStart, int *admmBadIterFlag, int reoptLevel, double *seconds);
/* averageUV for every cone (lorads_admm.c:310) and copyRtoV (lorads_alg_common.c:160), on the device state */
int lb2_average_uv(lb2_solver *s);
int lb2_copy_r_to_v(lb2_solver *s);
/* The iteration state the reference keeps in lorads_alm_state / lorads_admm_state (def_lorads_solver.h:130-161),
 * flattened in declaration order.  alm[13]: outerIter, innerIter, rho, pinf_inf, pinf_1, pinf_2, gap, pobj, dobj,
 * dinf_inf, dinf_1, dinf_2, tau.  admm[13]: iter, nBlks, cg_iter, rho, dinf_1, dinf_inf, pinf_1, pinf_inf, pinf_2,
 * dinf_2, pobj, dobj, gap.  set copies caller-side edits (main.c:404-410) back before the next phase call. */
int lb2_get_state(const lb2_solver *s, double *alm13, double *admm13);
int lb2_set_state(lb2_solver *s, const double *alm13, const double *admm13);
/* copies the current primal factor R (n x r col-major) and the dual vector to the host */
int lb2_get_solution(const lb2_solver *s, lb2_int iCone, double *R, double *dualVar);

#ifdef __cplusplus
}
#endif
#endif /* LORADS_B200_H */
