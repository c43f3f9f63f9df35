// alg.cu -- host control flow of the phases (ALM, ADMM, reopt, rank growth, dual infeasibility, the driver
// sequence).  Only scalars cross the PCIe bus inside the loops; every vector operation is a kernel launch.
//
// The control decisions follow the reference so that both solvers walk the same outer trajectory:
//   LORADS_ALMOptimize / _reopt   src_semi/lorads_alg/lorads_alm.c:745-1255
//   LORADSADMMOptimize / _reopt   src_semi/lorads_alg/lorads_admm.c:33-307
//   LORADS_ALMtoADMM, reopt, AUG_RANK, calculate_dual_infeasibility_solver, objScale_dualvar
//                                 src_semi/data/lorads_solver.c:758-1117
//   driver sequence               src_semi/main.c:321-487
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "solver.hpp"

namespace lb2 {

namespace {

// LUtilUpdateCheckEma, src_semi/lorads_utils.c:404-435: exponential moving average of the sub-problem certificate;
// returns 0 only when, at an evaluation point, the relative change of the EMA left the +-threshold band.
struct EmaStall {
    double cur = 0.0, old = 0.0;
    long long counter = 1;
    int update(double value, double alpha, double thr, long long interval) {
        int ok = 1;
        cur = alpha * value + (1 - alpha) * cur;
        if (counter >= interval) {
            if (old != 0) {
                double change = (cur - old) / old;
                ok = (change >= -thr) && (change <= thr);
            }
            old = cur;
            counter = 1;
        } else {
            counter++;
        }
        return ok;
    }
};

void alm_log(const Solver &S, const AlmState &a, double t, bool verbose) {
    if (!verbose) return;
    printf("ALM OuterIter:%lld InnerIter:%lld pObj:%5.5e dObj:%5.5e pInfea(1):%5.5e pInfea(Inf):%5.5e pdGap:%5.5e rho:%3.2f Time:%3.2f\n",
           a.outerIter, a.innerIter, a.pobj, a.dobj, a.pinf_1, a.pinf_inf, a.gap, a.rho, t);
    (void)S;
}

void admm_log(const AdmmState &a, double t, bool verbose) {
    if (!verbose) return;
    printf("ADMM Iter:%lld pObj:%5.5e dObj:%5.5e pInfea(1):%5.5e pInfea(Inf):%5.5e pdGap:%5.5e rho:%3.2f cgIter:%d Time:%3.2f\n",
           a.iter, a.pobj, a.dobj, a.pinf_1, a.pinf_inf, a.gap, a.rho, (int)((double)a.cg_iter / (double)a.nBlks), t);
}

}  // namespace

bool Solver::check_all_rank_max(double aug) const {
    // CheckAllRankMax, lorads_solver.c:758-774
    long long hit = 0;
    for (long long c = 0; c < nCones; ++c) {
        long long nr = (long long)std::min<double>(std::ceil((double)rank[c] * aug), (double)rank_max[c]);
        if (nr >= rank_max[c]) hit++;
    }
    return hit == nCones;
}

bool Solver::aug_rank(double aug) {
    // AUG_RANK, lorads_solver.c:806-906: append columns to U, V, R, Grad (M2temp keeps its old columns), the new
    // block of columns carries 1/sqrt(min(n, #new)) on its leading diagonal (lpRandomDiag :776-786); CG work
    // vectors and the L-BFGS history are re-allocated (zeroed).  Everything stays on the device: the old buffers are
    // set aside, new zero-filled ones are allocated for the larger leading dimensions and a copy kernel moves the
    // old columns across.  Column sharding: ownership is by global column index (k % world), so the columns a rank
    // already holds keep their local positions and only its share of the new columns is appended.
    if (check_all_rank_max(1.0)) return true;
    struct Old { long long off; int ld, r; };
    std::vector<Old> old((size_t)nCones);
    for (long long c = 0; c < nCones; ++c) old[(size_t)c] = Old{cones[c].off, cones[c].ld, cones[c].r};
    const long long N_old = N;
    std::vector<long long> old_rank = rank;
    DBuf<double> oR = std::move(R), oU = std::move(U), oV = std::move(V), oG = std::move(G), oM = std::move(M2);
    for (long long c = 0; c < nCones; ++c) {
        const long long nr = (long long)std::min<double>(std::ceil((double)rank[c] * aug), (double)rank_max[c]);
        rank[c] = nr;
        my_cols[c].clear();
        for (long long k = 0; k < nr; ++k)
            if (!shard_cols() || (int)(k % world) == myrank) my_cols[c].push_back((int)k);
    }
    alloc_vars();
    struct Pair { DBuf<double> *src, *dst; bool diag; };
    const Pair pairs[5] = {{&oR, &R, true}, {&oU, &U, true}, {&oV, &V, true}, {&oG, &G, true}, {&oM, &M2, false}};
    for (long long c = 0; c < nCones; ++c) {
        ConeDev &K = cones[c];
        const Old &o = old[(size_t)c];
        const long long n = blkDims[c], ro = old_rank[c], add = rank[c] - ro;
        // positions (inside this cone's slice) and value of the new diagonal entries held by this rank
        std::vector<int> pos;
        const long long rr = std::min(n, add);
        for (int lc = o.r; lc < K.r; ++lc) {
            const long long i = my_cols[c][(size_t)lc] - ro;          // row of the diagonal entry of global column ro + i
            if (i >= 0 && i < rr) pos.push_back((int)(i * K.ld + lc));
        }
        if (pos.size() > 1024) throw std::runtime_error("rank augmentation: more than 1024 new columns");
        if (!aug_pos.p) { aug_pos.alloc(1024); aug_val.alloc(1024); }          // allocated once, never freed mid-solve
        DBuf<int> &dpos = aug_pos;
        DBuf<double> &dval = aug_val;
        if (!pos.empty()) {
            const std::vector<double> vals(pos.size(), 1.0 / std::sqrt((double)rr));
            LB2_CUDA(cudaMemcpy(dpos.p, pos.data(), sizeof(int) * pos.size(), cudaMemcpyHostToDevice));
            LB2_CUDA(cudaMemcpy(dval.p, vals.data(), sizeof(double) * pos.size(), cudaMemcpyHostToDevice));
        }
        for (const Pair &p : pairs) {
            launch_relayout(ctx, n, o.r, o.ld, K.ld, p.src->p + o.off, p.dst->p + K.off);
            if (p.diag && !pos.empty())
                launch_scatter_add(ctx, p.dst->p + K.off, dval.p, dpos.p, (long long)pos.size(), 1.0, false, nullptr);
        }
        LB2_CUDA(cudaStreamSynchronize(ctx.stream));      // dpos / dval are rewritten for the next cone
    }
    if (nLp > 0) {
        // the LP vectors are separate arrays in the reference and survive AUG_RANK untouched
        for (const Pair &p : pairs)
            if (p.diag)
                LB2_CUDA(cudaMemcpyAsync(p.dst->p + N, p.src->p + N_old, sizeof(double) * nLp, cudaMemcpyDeviceToDevice, ctx.stream));
    }
    LB2_CUDA(cudaStreamSynchronize(ctx.stream));
    for (DBuf<double> *o : {&oR, &oU, &oV, &oG, &oM}) retire(*o);     // parked, not freed (see Solver::retired)
    return check_all_rank_max(aug);
}

void Solver::obj_scale_dualvar(double f) {
    // objScale_dualvar, lorads_solver.c:1040-1052: C *= f on every cone, lambda *= f, scaleObjHis *= f
    scaleObjHis *= f;
    drop_graphs();       // the rank-one objective coefficient is a by-value kernel argument of captured launches
    for (ConeDev &K : cones) {
        launch_scale(ctx, K.C_onP.p, (long long)K.C_onP.n, f);
        const long long nobj = K.listAC.dev.n_items - K.obj_item_begin;
        if (nobj > 0) launch_scale(ctx, K.listAC.coef.p + K.obj_item_begin, nobj, f);
        if (K.vc_on) launch_scale_tagged(ctx, (long long)K.vc_u_val.n, K.vc_u_tag.p, K.vc_u_val.p, f);
        K.c_rank1 *= f;
    }
    if (nLp > 0) launch_scale(ctx, lp_c.p, nLp, f);      // lp_cone_scalObj, lorads_lp_conic.c:197-202
    launch_scale(ctx, lam.p, m, f);
}

// ---------------------------------------------------------------------------------------------------
// ALM phase.  `reopt_variant` selects LORADS_ALMOptimize_reopt (lorads_alm.c:745) instead of
// LORADS_ALMOptimize (lorads_alm.c:991); the two differ in the outer loop shape, exit tests, difficulty classes.
// ---------------------------------------------------------------------------------------------------
int Solver::alm_optimize(lb2_params *P, double, double timeSolveStart, bool reopt_variant, bool early_stop,
                         double reopt_rho_factor) {
    if (!reopt_variant) MAX_ALM_SUB_ITER = 5000;
    const double ori_start = wall_time();
    const bool verbose = P->verbose != 0;
    bool is_rank_max = check_all_rank_max(1.0);
    int retcode = LB2_RET_OK;
    long long last_outer_start = 1;
    double tau = 0.0;
    const double rho_cert = 0.1;
    double cert_tol, cert_val, lagSq;
    char difficulty;
    long long localIter, clearLBFGS, rank_flag;
    const double rank_update_factor = 1.5;
    double rho_update_factor = reopt_variant ? reopt_rho_factor : P->ALMRhoFactor;
    long long rho_factor_flag;
    double rank_flag_thres = 1e8;
    if (P->dyrankLevel == 1) rank_flag_thres = 150;
    else if (P->dyrankLevel == 2) rank_flag_thres = 15;
    else if (P->dyrankLevel == 3) rank_flag_thres = 5;
    const int max_sub_iter_inc = 10000, max_sub_iter_ceil = 25000;
    int update_max_sub_iter_counter;
    long long k, k0;
    enum { GO_LOOP, GO_END_ALM, GO_PRINT_EXIT } jump;

alg_start:
    cert_tol = rho_cert / alm.rho;
    cert_val = 0;
    lagSq = 0.0;
    init_constr_val_all(R.p, R.p, true);
    constr_val_sum();
    lagSq = cal_grad(alm.rho);
    cert_val = std::sqrt(lagSq) / (1 + cObjNrmInf);
    difficulty = 'h';
    localIter = 0; clearLBFGS = 0; rank_flag = 0;
    if (!reopt_variant) rho_update_factor = P->ALMRhoFactor;
    rho_factor_flag = 0;
    update_max_sub_iter_counter = 0;
    k = alm.outerIter; k0 = alm.outerIter;
    jump = GO_LOOP;

    while (true) {
        if (reopt_variant) {
            if ((k > P->maxALMIter) && (alm.pinf_inf <= P->phase1Tol) &&
                ((alm.gap <= std::max(P->phase1Tol, P->phase2Tol * 5)) || !P->highAccMode)) { jump = GO_END_ALM; break; }
        } else {
            if (k > P->maxALMIter) { jump = GO_END_ALM; break; }
        }
        EmaStall ema;
        long long cur_iter_counter = 1;
        if (update_max_sub_iter_counter >= 2) {
            update_max_sub_iter_counter = 0;
            MAX_ALM_SUB_ITER = std::min(MAX_ALM_SUB_ITER + max_sub_iter_inc, max_sub_iter_ceil);
        }
        bool to_update_rho_direct = false;
        while (difficulty != 'e') {
            localIter = 0;
            int stall_ok = ema.update(cert_val, 0.1, 0.005, 5);
            if (!stall_ok && !P->highAccMode) break;
            if (cur_iter_counter >= MAX_ALM_SUB_ITER) { update_max_sub_iter_counter += 1; break; }
            if ((double)rank_flag >= rank_flag_thres && !is_rank_max && (k - last_outer_start >= 3)) break;
            if (cert_val <= cert_tol) break;
            // S_host may already hold the line-search scalars of a speculatively issued front half (same rho,
            // counter spec_counter); it is only reused inside this sub-problem loop
            bool spec = false;
            long long spec_counter = -1;
            auto reset_rule = [&](long long li) { return reopt_variant ? ((li - 1) % 300 == 0) : (li % 300 == 0); };
            while (cert_val - cert_tol > P->endALMSubTol) {
                if (reset_rule(localIter)) clearLBFGS = 0;
                double p12[2];
                if (!spec || spec_counter != clearLBFGS) { push_scalars(tau, alm.rho); enqueue_front(alm.rho, clearLBFGS); read_slots(); }
                spec = false;
                const long long rootNum = finish_front(alm.rho, &tau, p12);
                if (rootNum == 0) { retcode = LB2_RET_NUM_ERR; jump = GO_END_ALM; goto after_loops; }
                if (std::fabs(tau) < P->endTauTol) {
                    if (verbose) printf("update rho:%5.8e since tau is too small.\n", tau);
                    alm.innerIter++; localIter++; cur_iter_counter++; clearLBFGS++;
                    to_update_rho_direct = true;
                    break;
                }
                double pinf1 = 0;
                spec_counter = reset_rule(localIter + 1) ? 0 : clearLBFGS + 1;
                iter_back_front(alm.rho, tau, spec_counter);   // back half + next iteration's front half, one graph
                spec = true;
                finish_back(&lagSq, &pinf1);
                alm.pinf_1 = pinf1;
                alm.pinf_inf = pinf1 * (1 + bNrm1) / (1 + bNrmInf);
                if (!reopt_variant) {
                    if ((alm.pinf_inf <= P->phase1Tol) && ((alm.gap <= P->phase1Tol) || !P->highAccMode)) {
                        alm.outerIter = k;
                        alm.innerIter += 1; localIter += 1; cur_iter_counter += 1; clearLBFGS += 1;
                        jump = GO_END_ALM;
                        goto after_loops;
                    }
                }
                cert_val = std::sqrt(lagSq) / (1 + cObjNrmInf);
                alm.innerIter += 1; localIter++; cur_iter_counter++; clearLBFGS++;
                if (localIter > 800) break;
            }
            if (to_update_rho_direct) break;
            update_dual_var(alm.rho);
            lagSq = cal_grad(alm.rho);
            cert_val = std::sqrt(lagSq) / (1 + cObjNrmInf);
            if (localIter <= 20) difficulty = 'e';
            else if (localIter <= 100) { difficulty = 'm'; rank_flag += 2; }
            else if (reopt_variant || localIter < 400) { difficulty = 'h'; rank_flag += 3; }
            else { difficulty = 's'; rank_flag += 4; }
            if (difficulty == 'e') rank_flag = 0;
        }
        // UpdateRho
        do {
            alm.rho *= rho_update_factor;
            lagSq = cal_grad(alm.rho);
            cert_val = std::sqrt(lagSq) / (1 + cObjNrmInf);
            cert_tol = rho_cert / alm.rho;
        } while (cert_tol >= cert_val);
        if (alm.rho >= 5e4 && rho_factor_flag < 4) { rho_update_factor = std::sqrt(std::sqrt(rho_update_factor)); rho_factor_flag = 4; }
        else if (alm.rho >= 5e6 && rho_factor_flag < 6) { rho_update_factor = std::sqrt(std::sqrt(rho_update_factor)); rho_factor_flag = 6; }
        else if (alm.rho >= 5e8 && rho_factor_flag < 8) { rho_update_factor = std::sqrt(std::sqrt(rho_update_factor)); rho_factor_flag = 8; }
        difficulty = 'h';
        clearLBFGS = 0;
        if (reopt_variant) k += 1;
        alm.outerIter = k;
        if (!reopt_variant) {
            if ((alm.pinf_inf <= P->phase1Tol) && ((alm.gap <= P->phase1Tol) || !P->highAccMode)) { jump = GO_END_ALM; break; }
        }
        pObj = cal_obj(R.p);
        dObj = cal_dual_obj();
        update_dimacs_alm();
        alm.gap = dimac_gap; alm.pobj = pObj; alm.dobj = dObj;
        alm.pinf_1 = dimac_pinf;
        alm.pinf_inf = alm.pinf_1 * (1 + bNrm1) / (1 + bNrmInf);
        alm.dinf_1 = 99; alm.dinf_inf = 99;
        if (reopt_variant) {
            if (early_stop) {
                if (alm.pinf_1 <= P->phase1Tol && alm.gap <= std::max(P->phase1Tol, P->phase2Tol * 5) && (k - k0) > 1) { jump = GO_PRINT_EXIT; break; }
            } else {
                if (alm.gap <= P->phase2Tol && alm.pinf_1 <= P->phase2Tol && (k - k0) > 1) { jump = GO_PRINT_EXIT; break; }
            }
        } else {
            if (alm.gap <= P->phase1Tol * 1e-3 && alm.pinf_1 <= P->phase1Tol * 1e-3) { jump = GO_PRINT_EXIT; break; }
        }
        alm_log(*this, alm, wall_time() - ori_start, verbose);
        if (wall_time() - timeSolveStart >= P->timeSecLimit) { jump = GO_PRINT_EXIT; break; }
        if ((double)rank_flag >= rank_flag_thres && !is_rank_max && (!reopt_variant || nCones <= 10)) {
            rank_flag = 0;
            if (k - last_outer_start >= 2) {
                if (verbose) printf("increase the rank, factor:%f.\n", rank_update_factor);
                is_rank_max = aug_rank(rank_update_factor);
                alm.outerIter = k;
                last_outer_start = alm.outerIter;
                goto alg_start;
            }
        }
        if (!reopt_variant) k++;
    }
after_loops:
    if (jump == GO_END_ALM) {
        pObj = cal_obj(R.p);
        dObj = cal_dual_obj();
        update_dimacs_alm();
        if (reopt_variant) {
            alm.gap = dimac_gap;
            alm.pinf_1 = alm.pinf_inf * (1 + bNrmInf) / (1 + bNrm1);
        } else {
            alm.pobj = pObj; alm.dobj = dObj; alm.gap = dimac_gap;
            alm.pinf_1 = dimac_pinf;
            alm.pinf_inf = alm.pinf_1 * (1 + bNrm1) / (1 + bNrmInf);
        }
        alm.dinf_1 = 99; alm.dinf_inf = 99;
    }
    if (verbose) {
        printf("-----------------------------------------------------------------------\n");
        printf("Exit ALM:\n");
        alm_log(*this, alm, wall_time() - ori_start, true);
        printf("-----------------------------------------------------------------------\n");
    }
    return retcode;
}

void Solver::alm_to_admm(lb2_params *P) {
    // LORADS_ALMtoADMM, lorads_solver.c:968-1004: U = V = R, state hand-off, rho heuristic
    LB2_CUDA(cudaMemcpyAsync(V.p, R.p, sizeof(double) * Nt, cudaMemcpyDeviceToDevice, ctx.stream));
    LB2_CUDA(cudaMemcpyAsync(U.p, R.p, sizeof(double) * Nt, cudaMemcpyDeviceToDevice, ctx.stream));
    admm.dinf_1 = alm.dinf_1; admm.pinf_1 = alm.pinf_1; admm.dinf_2 = alm.dinf_2; admm.dinf_inf = alm.dinf_inf;
    admm.pinf_inf = alm.pinf_inf; admm.pinf_2 = alm.pinf_2; admm.gap = alm.gap;
    admm.rho = alm.rho * P->heuristicFactor;
    if (alm.rho > P->rhoMax) {
        admm.rho = std::min(std::sqrt(std::max(P->rhoMax, alm.rho) / P->rhoMax) * P->rhoMax, alm.rho);
        P->rhoMax = admm.rho;
    }
}

int Solver::admm_optimize(lb2_params *P, long long iter_celling, double timeSolveStart, bool reopt_variant) {
    if (admm.gap <= P->phase2Tol && admm.pinf_1 <= P->phase2Tol) return LB2_RET_OK;
    const bool verbose = P->verbose != 0;
    double endCG_tol = std::min(admm.pinf_1 * 1e-4, 1e-8);
    const long long CGMaxIter = 800;
    admm.rho = std::min(admm.rho, P->rhoMax);
    cgIter = 0;
    init_constr_val_all(U.p, V.p, false);
    constr_val_sum();
    auto refresh_obj = [&]() {
        // LORADSCalObjUV_ADMM (lorads_admm.c:325-337), LORADSCalDualObj and LORADSUpdateDimacsErrorADMM: R = (U+V)/2,
        // <C,RR^T>, b.lambda, |b - A(RR^T)|; everything is enqueued first and the scalars come back in ONE read
        average_uv();
        LB2_CUDA(cudaMemsetAsync(S.p + SL_OBJ, 0, sizeof(double), ctx.stream));
        for (ConeDev &K : cones) cone_auv(K, true, R.p, R.p, true, 1.0, K.t1.p, S.p + SL_OBJ);
        if (world > 1) allreduce(S.p + SL_OBJ, 1);
        if (nLp > 0) launch_lp_obj(ctx, lp, R.p + N, R.p + N, 1.0, S.p + SL_OBJ, 0.0, nullptr);
        launch_dot(ctx, m, b.p, lam.p, S.p, SL_DOBJ);
        primal_infeasibility(R.p);    // also leaves constrVal / constrValSum = A(RR^T), as updateDimacsADMM does
        read_slots();
        pObj = S_host[SL_OBJ] / scaleObjHis;
        dObj = S_host[SL_DOBJ] / scaleObjHis;
        dimac_pinf = std::sqrt(S_host[SL_PINF]) / (1 + bNrm1);
        dimac_gap = std::fabs(pObj - dObj) / (1 + std::fabs(pObj) + std::fabs(dObj));
    };
    refresh_obj();
    admm.pobj = pObj; admm.dobj = dObj; admm.gap = dimac_gap; admm.pinf_1 = dimac_pinf;
    admm.pinf_inf = dimac_pinf * (1 + bNrm1) / (1 + bNrmInf);
    admm.pinf_2 = dimac_pinf * (1 + bNrm1) / (1 + bNrm2);
    if (reopt_variant && verbose) { printf("enter admm reopt \n"); admm_log(admm, 0, true); }
    cgTime = 0.0;
    double cur_rho_max = P->rhoMax;
    double old_mean = 1e30;
    double buffer[10] = {0};
    int bad_pd = 0;
    const int bad_pd_limit = reopt_variant ? 200 : 800;
    const double origTime = wall_time();
    while (admm.iter <= P->maxADMMIter || admm.gap >= P->phase2Tol || admm.pinf_1 >= P->phase2Tol) {
        if (admm.iter >= iter_celling) {
            if (reopt_variant && verbose) {
                admm_log(admm, 0, true);
                printf("exit admm since maxiter greater than iter_celling:%lld\n", iter_celling);
            }
            break;
        }
        endCG_tol = std::min(admm.pinf_1 * (reopt_variant ? 1e-4 : 1e-2), 1e-8);
        update_sdp_var(admm.rho, endCG_tol, CGMaxIter);
        admm.cg_iter = cgIter;
        refresh_obj();
        admm.pobj = pObj; admm.dobj = dObj;
        admm.gap = dimac_gap;
        if (reopt_variant) {
            admm.pinf_1 = dimac_pinf;
            admm.pinf_2 = dimac_pinf * (1 + bNrm1) / (1 + bNrm2);
        }
        admm.pinf_inf = dimac_pinf * (1 + bNrm1) / (1 + bNrmInf);
        if (admm.pinf_inf >= 1e10 || admm.gap >= 1 - 1e-8) {
            if (reopt_variant) admm_log(admm, wall_time() - origTime, verbose);
            if (verbose) printf("Numerical Error!\n");
            return LB2_RET_NUM_ERR;
        }
        if (admm.gap <= P->phase2Tol * 5) bad_pd = std::max(0, bad_pd - 5);
        else if (admm.gap <= P->phase2Tol) bad_pd = std::max(0, bad_pd - 10);
        if (admm.gap >= P->phase1Tol * 1e2) bad_pd += 2;
        if (bad_pd >= bad_pd_limit) return LB2_RET_OK;
        buffer[0] = admm.pinf_inf;   // the reference indexes with a counter that never advances
        if (!reopt_variant) {
            if (admm.pinf_inf <= P->phase2Tol) {
                update_dimacs_admm();
                admm.pobj = pObj; admm.dobj = dObj; admm.gap = dimac_gap; admm.pinf_1 = dimac_pinf;
                admm_log(admm, wall_time() - origTime, verbose);
                return LB2_RET_OK;
            }
        } else if (admm.pinf_1 <= P->phase2Tol) {
            update_dimacs_admm();
            admm.pobj = pObj; admm.dobj = dObj; admm.gap = dimac_gap; admm.pinf_1 = dimac_pinf;
            if (admm.gap <= P->phase2Tol) {
                admm_log(admm, wall_time() - origTime, verbose);
                return LB2_RET_OK;
            }
        }
        update_dual_var(admm.rho);
        const long long it_eff = reopt_variant ? admm.iter : admm.iter + 1;
        if (it_eff % P->rhoFreq == 0) {
            admm.rho *= P->rhoFactor;
            if (admm.rho >= cur_rho_max) {
                admm.rho = cur_rho_max;
                if (it_eff % (P->rhoFreq * 100) == 0) {
                    double mean = 0;
                    for (double v : buffer) mean += std::fabs(v);
                    mean /= 10.0;
                    if (mean / old_mean >= 0.65) {
                        admm.rho *= std::pow(P->rhoFactor, std::round(std::log((double)(P->rhoFreq * 100)) / std::log((double)P->rhoFreq)));
                        cur_rho_max = admm.rho;
                    }
                    old_mean = mean;
                }
            }
            if (admm.rho >= P->rhoCellingADMM) admm.rho = P->rhoCellingADMM;
        }
        if (admm.iter % 50 == 0) {
            update_dimacs_admm();
            admm.pobj = pObj; admm.dobj = dObj; admm.gap = dimac_gap; admm.pinf_1 = dimac_pinf;
            admm_log(admm, wall_time() - origTime, verbose);
            if (wall_time() - timeSolveStart >= P->timeSecLimit) return LB2_RET_TIME_OUT;
        }
        if (admm.gap <= P->phase2Tol * 1e-3 && admm.pinf_1 <= P->phase2Tol * 1e-3) {
            if (verbose) printf("Early Stop When DIMACS Errors Are Well-Satisfied");
            return LB2_RET_OK;
        }
        admm.iter++;
    }
    if (reopt_variant) admm_log(admm, wall_time() - origTime, verbose);
    return LB2_RET_OK;
}

double Solver::reopt(lb2_params *P, double *reopt_param, long long *alm_iter, long long *admm_iter, double timeSolveStart,
                     int *admm_bad_iter_flag, int reopt_level) {
    // reopt, lorads_solver.c:1075-1117
    const long long old_maxALM = P->maxALMIter, old_maxADMM = P->maxADMMIter;
    const double old_rhoMax = P->rhoMax;
    P->maxALMIter = alm_iter[0] - 1 + alm.outerIter;
    P->maxADMMIter = admm_iter[0];
    obj_scale_dualvar(*reopt_param);
    if (admm.rho <= P->rhoMax) alm.rho = std::max(admm.rho, alm.rho);
    const double t0 = wall_time();
    alm_optimize(P, 0.0, timeSolveStart, true, true, std::sqrt(P->ALMRhoFactor));
    P->rhoMax = std::max(std::sqrt(std::max(admm.rho, alm.rho) / admm.rho) * admm.rho, P->rhoMax);
    alm_to_admm(P);
    if (*admm_bad_iter_flag == 0 || reopt_level < 2) {
        int rc = admm_optimize(P, std::min(admm.iter * 4, admm.iter + old_maxADMM), timeSolveStart, true);
        *admm_bad_iter_flag = (rc == LB2_RET_BAD_ITER) ? 1 : 0;
    }
    const double t1 = wall_time();
    P->maxALMIter = old_maxALM; P->maxADMMIter = old_maxADMM; P->rhoMax = old_rhoMax;
    return t1 - t0;
}

// ---------------------------------------------------------------------------------------------------
// dual infeasibility: lambda_min(C - A^*(lambda)) per cone, explicitly restarted Lanczos (same algorithm as the
// ARPACK stand-in linked into the reference oracle, oracle/arpack_shim.c); the mat-vec runs on the device.
// ---------------------------------------------------------------------------------------------------
static void jacobi_eig(int k, std::vector<double> &a, std::vector<double> &q) {
    q.assign((size_t)k * k, 0.0);
    for (int i = 0; i < k; ++i) q[(size_t)i * k + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < k; ++i)
            for (int j = i + 1; j < k; ++j) off += a[(size_t)i * k + j] * a[(size_t)i * k + j];
        if (off < 1e-300) break;
        for (int p = 0; p < k; ++p)
            for (int r = p + 1; r < k; ++r) {
                double apr = a[(size_t)p * k + r];
                if (std::fabs(apr) < 1e-300) continue;
                double app = a[(size_t)p * k + p], arr = a[(size_t)r * k + r];
                double tau = (arr - app) / (2.0 * apr);
                double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
                double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
                for (int i = 0; i < k; ++i) {
                    double aip = a[(size_t)i * k + p], air = a[(size_t)i * k + r];
                    a[(size_t)i * k + p] = c * aip - s * air; a[(size_t)i * k + r] = s * aip + c * air;
                }
                for (int i = 0; i < k; ++i) {
                    double api = a[(size_t)p * k + i], ari = a[(size_t)r * k + i];
                    a[(size_t)p * k + i] = c * api - s * ari; a[(size_t)r * k + i] = s * api + c * ari;
                }
                for (int i = 0; i < k; ++i) {
                    double qip = q[(size_t)i * k + p], qir = q[(size_t)i * k + r];
                    q[(size_t)i * k + p] = c * qip - s * qir; q[(size_t)i * k + r] = s * qip + c * qir;
                }
            }
    }
}

void Solver::dual_infeasibility() {
    // calculate_dual_infeasibility_solver, lorads_solver.c:1007-1037.  The Lanczos basis, the mat-vec
    // (S = C - A^*(lambda) on the pattern) and the re-orthogonalisation all stay on the device; per Lanczos step
    // the host reads two scalars (alpha, |w|^2) and solves the small tridiagonal eigenproblem per restart cycle.
    static const bool timing = getenv("LORADS_B200_TIMING") != nullptr;
    const double t_begin = wall_time();
    int n_cycles = 0;
    launch_axpby_dot(ctx, m, M1.p, coef_const(-1.0), lam.p, coef_const(0.0), nullptr, nullptr, S.p, SL_T1, false);
    double total = 0.0;
    if (nLp > 0) {
        // LP columns first (lorads_solver.c:1015-1023): sum_j |min(c_j - a_j^T lambda, 0)|
        launch_lp_dinf(ctx, lp, M1.p, S.p, SL_T0);
        read_slots();
        total += S_host[SL_T0];
    }
    for (long long c = 0; c < nCones; ++c) {
        ConeDev &K = cones[c];
        cone_wsum(K, M1.p, false, true, true);
        const long long n = K.n;
        int kdim = (40 > n) ? 2 : 40;       // dual_infeasible: ncv = 40, or 2 when ncv > n (lorads_sdp_conic.c:1288-1294)
        if (kdim > n) kdim = (int)n;
        const long long np2 = (n + 1) & ~1LL;            // even stride keeps every basis vector 16-byte aligned
        auto want = [&](DBuf<double> &buf, size_t count) { if (buf.n < count) { retire(buf); buf.alloc(count); } };
        want(lz_basis, (size_t)np2 * (kdim + 1)); want(lz_w, (size_t)np2); want(lz_x0, (size_t)np2);
        want(lz_h, kLanczosMax); want(lz_scratch, (size_t)512 * kLanczosMax); want(lz_tab, 2 * (size_t)kLanczosMax);
        DBuf<double> &Vb = lz_basis, &w = lz_w, &x0 = lz_x0, &hbuf = lz_h, &hscratch = lz_scratch, &lz = lz_tab;
        std::vector<double> v0((size_t)n);
        uint64_t st = 0x9E3779B97F4A7C15ull;
        double nr0 = 0.0;
        for (long long i = 0; i < n; ++i) {
            st = st * 6364136223846793005ull + 1442695040888963407ull;
            v0[(size_t)i] = ((double)(st >> 11) / 9007199254740992.0) - 0.5;
            nr0 += v0[(size_t)i] * v0[(size_t)i];
        }
        nr0 = std::sqrt(nr0);
        for (double &e : v0) e /= nr0;
        LB2_CUDA(cudaMemcpyAsync(Vb.p, v0.data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx.stream));
        auto vec = [&](int l) { return Vb.p + (size_t)l * np2; };
        auto matvec = [&](const double *in, double *out) {
            if (K.dense_path) launch_dense_symv(ctx, n, K.S.p, in, out);
            else if (K.c_rank1 != 0.0) {
                launch_sum(ctx, n, in, S.p, SL_DG);
                launch_spmv_sym(ctx, n, K.adj_ptr.p, K.adj_col.p, K.adj_pos.p, K.S.p, in, out, K.c_rank1, S.p + SL_DG);
            } else launch_spmv_sym(ctx, n, K.adj_ptr.p, K.adj_col.p, K.adj_pos.p, K.S.p, in, out);
        };
        double theta = 0.0;
        const double tol = 1e-2;
        std::vector<double> alpha, beta;
        // One restart cycle = kdim Lanczos steps issued back to back: alpha_j and |w_j|^2 land in a small device table
        // and are read ONCE per cycle.  A breakdown (beta ~ 0) inside the cycle only produces garbage in the later
        // steps, which the host discards by truncating the tridiagonal matrix at the first tiny beta.
        std::vector<double> lz_host(2 * (size_t)kLanczosMax);
        const double t_alloc = wall_time();
        auto enqueue_cycle = [&]() {
            for (int j = 0; j < kdim; ++j) {
                // w = S v_j, then two classical Gram-Schmidt passes against v_0..v_j, each one dots kernel + one
                // update kernel (full re-orthogonalisation; the first pass also yields alpha_j = v_j . S v_j and
                // removes the beta_{j-1} v_{j-1} term), |w|^2 from the last update, v_{j+1} = w/|w|
                matvec(vec(j), w.p);
                launch_lanczos_dots(ctx, n, Vb.p, np2, j + 1, w.p, hbuf.p, hscratch.p, lz.p, j);
                launch_lanczos_update(ctx, n, Vb.p, np2, j + 1, hbuf.p, w.p, lz.p, -1);
                launch_lanczos_dots(ctx, n, Vb.p, np2, j + 1, w.p, hbuf.p, hscratch.p, lz.p, -1);
                launch_lanczos_update(ctx, n, Vb.p, np2, j + 1, hbuf.p, w.p, lz.p, kLanczosMax + j);
                if (j + 1 < kdim) launch_lanczos_next(ctx, n, w.p, vec(j + 1), lz.p, kLanczosMax + j);
            }
        };
        // the ~6*kdim launches of a cycle never change between restarts: record them once, replay as one graph
        cudaGraphExec_t cycle_exec = nullptr;
        long long cycle_launches = 0;
        for (int cycle = 0; cycle < 600; ++cycle) {
            ++n_cycles;
            alpha.clear(); beta.clear();
            if (!use_graphs) enqueue_cycle();
            else {
                if (!cycle_exec) {
                    const long long l0 = ctx.launches;
                    cudaGraph_t graph = nullptr;
                    LB2_CUDA(cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal));
                    try { enqueue_cycle(); } catch (...) {
                        cudaStreamEndCapture(ctx.stream, &graph);
                        if (graph) cudaGraphDestroy(graph);
                        throw;
                    }
                    LB2_CUDA(cudaStreamEndCapture(ctx.stream, &graph));
                    LB2_CUDA(cudaGraphInstantiate(&cycle_exec, graph, 0));
                    LB2_CUDA(cudaGraphDestroy(graph));
                    cycle_launches = ctx.launches - l0;
                    ctx.launches = l0;
                }
                LB2_CUDA(cudaGraphLaunch(cycle_exec, ctx.stream));
                ctx.launches += cycle_launches;
            }
            LB2_CUDA(cudaMemcpyAsync(lz_host.data(), lz.p, sizeof(double) * lz_host.size(), cudaMemcpyDeviceToHost, ctx.stream));
            LB2_CUDA(cudaStreamSynchronize(ctx.stream));
            int mdim = kdim;
            bool breakdown = false;
            for (int j = 0; j < kdim; ++j) {
                const double a = lz_host[(size_t)j], bb = std::sqrt(lz_host[(size_t)kLanczosMax + j]);
                alpha.push_back(a); beta.push_back(bb);
                if (!(bb >= 1e-14 * (std::fabs(a) + 1.0))) { mdim = j + 1; breakdown = true; break; }   // also catches NaN
            }
            const double bb = beta[(size_t)mdim - 1];
            std::vector<double> T((size_t)mdim * mdim, 0.0), Q;
            for (int i = 0; i < mdim; ++i) {
                T[(size_t)i * mdim + i] = alpha[i];
                if (i + 1 < mdim) { T[(size_t)i * mdim + i + 1] = beta[i]; T[(size_t)(i + 1) * mdim + i] = beta[i]; }
            }
            jacobi_eig(mdim, T, Q);
            int best = 0;
            for (int i = 1; i < mdim; ++i) if (T[(size_t)i * mdim + i] < T[(size_t)best * mdim + best]) best = i;
            theta = T[(size_t)best * mdim + best];
            const double est = std::fabs(bb * Q[(size_t)(mdim - 1) * mdim + best]);
            const double scale = std::max(std::fabs(theta), 2.2e-16);
            if (breakdown || est <= tol * scale || mdim >= n) break;
            // explicit restart from the Ritz vector x0 = sum_l Q[l][best] v_l, normalised on the device
            std::vector<double> hq((size_t)mdim);
            for (int l = 0; l < mdim; ++l) hq[(size_t)l] = -Q[(size_t)l * mdim + best];
            LB2_CUDA(cudaMemcpyAsync(hbuf.p, hq.data(), sizeof(double) * mdim, cudaMemcpyHostToDevice, ctx.stream));
            LB2_CUDA(cudaStreamSynchronize(ctx.stream));      // hq is pageable stack-lifetime memory
            LB2_CUDA(cudaMemsetAsync(x0.p, 0, sizeof(double) * np2, ctx.stream));
            launch_lanczos_update(ctx, n, Vb.p, np2, mdim, hbuf.p, x0.p, S.p, SL_T1);
            launch_lanczos_next(ctx, n, x0.p, vec(0), S.p, SL_T1);
        }
        if (cycle_exec) { LB2_CUDA(cudaStreamSynchronize(ctx.stream)); cudaGraphExecDestroy(cycle_exec); }
        if (timing) fprintf(stderr, "  dual infeasibility cone %lld: setup %.3f s, %d cycles %.3f s\n", c, t_alloc - t_begin, n_cycles, wall_time() - t_alloc);
        total += std::fabs(std::min(theta, 0.0));
    }
    dimac_dinf = total / scaleObjHis / (cObjNrm1 + 1);
    if (timing) { sync(); fprintf(stderr, "  dual infeasibility total %.3f s\n", wall_time() - t_begin); }
}

// ---------------------------------------------------------------------------------------------------
// the driver sequence of src_semi/main.c:321-487
// ---------------------------------------------------------------------------------------------------
int Solver::solve(lb2_params *P, lb2_result *res) {
    const double t0 = wall_time();
    const long long launches0 = ctx.launches;
    double reopt_param = 5;
    long long reopt_alm_iter = 3, reopt_admm_iter = 50, alm_reopt_min_iter = 3, admm_reopt_min_iter = 50;
    if (P->highAccMode) admm_reopt_min_iter = 1000;
    int admm_bad_iter_flag = 0;
    status = LB2_STATUS_UNKNOWN;
    double tA = wall_time();
    alm_optimize(P, 0.0, t0, false, false, 0.0);
    double almSeconds = wall_time() - tA, admmSeconds = 0;
    bool timed_out = false;
    if (wall_time() - t0 > P->timeSecLimit) { status = LB2_STATUS_TIME_LIMIT; timed_out = true; }
    if (!timed_out) {
        alm_to_admm(P);
        tA = wall_time();
        if (admm_optimize(P, P->maxADMMIter, t0, false) == LB2_RET_BAD_ITER) admm_bad_iter_flag = 1;
        admmSeconds = wall_time() - tA;
        int cnt = 0;
        if (P->reoptLevel >= 1) {
            while ((alm.gap > P->phase2Tol || alm.pinf_1 > P->phase2Tol) && (admm.gap > P->phase2Tol || admm.pinf_1 > P->phase2Tol)) {
                if (cnt >= 1) break;
                if (P->verbose) printf("******  reopt parameter:%.3f\n", reopt_param);
                reopt(P, &reopt_param, &alm_reopt_min_iter, &admm_reopt_min_iter, t0, &admm_bad_iter_flag, 1);
                cnt += 1;
                if (wall_time() - t0 > P->timeSecLimit) { status = LB2_STATUS_TIME_LIMIT; timed_out = true; break; }
            }
        }
    }
    if (!timed_out) {
        dual_infeasibility();
        auto refresh = [&]() {
            admm.dinf_1 = dimac_dinf;
            admm.dinf_inf = dimac_dinf * (1 + cObjNrm1) / (1 + cObjNrmInf);
            admm.dinf_2 = dimac_dinf * (1 + cObjNrm1) / (1 + cObjNrm2);
            admm.gap = dimac_gap;
            admm.pinf_1 = dimac_pinf;
            admm.pinf_inf = dimac_pinf * (1 + bNrm1) / (1 + bNrmInf);
            admm.pinf_2 = dimac_pinf * (1 + bNrm1) / (1 + bNrm2);
        };
        refresh();
        if (P->reoptLevel >= 2) {
            int dual_cnt = 0;
            while (admm.dinf_1 > P->phase2Tol || admm.gap > P->phase2Tol || admm.pinf_1 > P->phase2Tol) {
                if (dual_cnt >= 2) break;
                if (!P->highAccMode && admm.dinf_1 <= 5 * P->phase2Tol && admm.gap <= 5 * P->phase2Tol && admm.pinf_1 <= P->phase2Tol) break;
                if (P->verbose) printf("******  reopt parameter:%.3f\n", reopt_param);
                reopt(P, &reopt_param, &reopt_alm_iter, &reopt_admm_iter, t0, &admm_bad_iter_flag, 2);
                average_uv();                                                        // main.c:441-443
                LB2_CUDA(cudaMemcpyAsync(V.p, R.p, sizeof(double) * Nt, cudaMemcpyDeviceToDevice, ctx.stream));   // copyRtoV / copyRtoVLP
                dual_infeasibility();
                refresh();
                dual_cnt += 1;
                if (wall_time() - t0 > P->timeSecLimit) { status = LB2_STATUS_TIME_LIMIT; timed_out = true; break; }
            }
        }
    }
    if (!timed_out) {
        if (admm.dinf_1 <= 5 * P->phase2Tol && admm.gap <= 5 * P->phase2Tol && admm.pinf_1 <= P->phase2Tol) status = LB2_STATUS_PRIMAL_DUAL_OPTIMAL;
        else if (admm.gap <= 5 * P->phase2Tol && admm.pinf_1 <= P->phase2Tol) status = LB2_STATUS_PRIMAL_OPTIMAL;
        else status = LB2_STATUS_MAXITER;
    }
    sync();
    if (res) {
        res->pObj = pObj; res->dObj = dObj;
        res->pInfeasL1 = dimac_pinf; res->dInfeasL1 = dimac_dinf; res->pdGap = dimac_gap;
        res->pInfeasInf = dimac_pinf * (1 + bNrm1) / (1 + bNrmInf);
        res->dInfeasInf = dimac_dinf * (1 + cObjNrm1) / (1 + cObjNrmInf);
        res->almOuterIter = alm.outerIter; res->almInnerIter = alm.innerIter; res->admmIter = admm.iter; res->cgIter = cgIter;
        res->almRho = alm.rho; res->admmRho = admm.rho;
        res->solveSeconds = wall_time() - t0; res->almSeconds = almSeconds; res->admmSeconds = admmSeconds;
        res->status = status; res->finalRank0 = rank.empty() ? 0 : rank[0];
        res->kernelLaunches = ctx.launches - launches0;
    }
    return LB2_OK;
}

}  // namespace lb2
