// solver.cu -- setup, state transfer and the hot-path operators of the device-resident solver.
//
// Host code here replaces src_semi/data/lorads_solver.c (allocation, rank choice, random start) and the glue of
// src_semi/lorads_alg/lorads_alg_common.c; every numerical loop is a kernel from kernels.cu.
#include "solver.hpp"

#include <sys/time.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

namespace lb2 {

double wall_time() {
    struct timeval tv;
    gettimeofday(&tv, nullptr);
    return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}

void ItemListBufs::upload(const ItemList &L) {
    ptr.upload(L.ptr); irow.upload(L.irow); icol.upload(L.icol); coef.upload(L.coef);
    split_row.upload(L.split_row); split_first_slot.upload(L.split_first_slot);
    split_tile_a.upload(L.split_tile_a); split_tile_b.upload(L.split_tile_b);
    dev.n_items = L.n_items(); dev.n_rows = L.n_rows; dev.n_tiles = L.n_tiles(); dev.n_split = (long long)L.split_row.size();
    dev.ptr = ptr.p; dev.irow = irow.p; dev.icol = icol.p; dev.coef = coef.p;
    dev.split_row = split_row.p; dev.split_first_slot = split_first_slot.p;
    dev.split_tile_a = split_tile_a.p; dev.split_tile_b = split_tile_b.p;
    tile_row_lo.upload(L.tile_row_lo); tile_row_hi.upload(L.tile_row_hi);
    dev.tile_row_lo = tile_row_lo.p; dev.tile_row_hi = tile_row_hi.p;
    dev.tile = L.tile;
    dev.obj_row = L.has_obj ? (int)(L.n_rows - 1) : -1;
    dev.has_empty_rows = false;
    for (int64_t r = 0; r < L.n_rows; ++r)
        if (L.ptr[r + 1] == L.ptr[r]) dev.has_empty_rows = true;
}

void Solver::drop_graphs() {
    for (auto &kv : iter_graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    iter_graphs.clear();
}

void Solver::push_scalars(double tau, double rho) {
    push_host[0] = tau; push_host[1] = rho;
    LB2_CUDA(cudaMemcpyAsync(S.p + SL_TAU, push_host, 2 * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
}

Solver::~Solver() {
    drop_graphs();
    if (push_host) cudaFreeHost(push_host);
    if (S_host) cudaFreeHost(S_host);
    if (side_event) cudaEventDestroy(side_event);
    if (side_stream) cudaStreamDestroy(side_stream);
    if (ctx.stream) cudaStreamDestroy(ctx.stream);
}

void Solver::create(long long nRows, long long nC, const lb2_int *dims, const double *rhs, int dev) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        throw CudaError("no usable CUDA device: the LoRADS B200 kernel layer has no CPU fallback");
    if (dev < 0 || dev >= count) throw CudaError("CUDA device ordinal out of range");
    device = dev;
    LB2_CUDA(cudaSetDevice(dev));
    LB2_CUDA(cudaStreamCreate(&ctx.stream));
    cudaDeviceProp prop;
    LB2_CUDA(cudaGetDeviceProperties(&prop, dev));
    ctx.num_sms = prop.multiProcessorCount;
    m = nRows; nCones = nC;
    blkDims.assign(dims, dims + nC);
    b_h.assign(rhs, rhs + nRows);
    inputs.resize(nC);
    red_partials.alloc(8192 * 8);
    red_counter.alloc(4);
    ctx.rs.partials = red_partials.p;
    ctx.rs.counter = red_counter.p;
    ctx.ticket = red_counter.p + 1;
    S.alloc(kNumSlots + 2 * (nC + 1));
    LB2_CUDA(cudaMallocHost((void **)&S_host, sizeof(double) * (kNumSlots + 2 * (nC + 1))));
    std::memset(S_host, 0, sizeof(double) * (kNumSlots + 2 * (nC + 1)));
    S_host[SL_ONE] = 1.0;
    LB2_CUDA(cudaMallocHost((void **)&push_host, sizeof(double) * 4));
    use_graphs = (getenv("LORADS_B200_NO_GRAPH") == nullptr);
    // L-BFGS with history 2: the Gram-table form (every inner product of the two-loop recursion from one pass over the
    // vectors, launch_lbfgs_pair / launch_lbfgs_dir) is the default on one GPU and under sharding alike -- 26 % more steps
    // per second on MaxCut n = 1e5 and one reduction instead of five; LORADS_B200_EXACT_LBFGS=1 keeps the reference's
    // two-loop recursion pass by pass (lorads_alm.c:230-391).  Both agree with the reference to rounding (tests).
    vf_lbfgs = (getenv("LORADS_B200_EXACT_LBFGS") == nullptr);
    LB2_CUDA(cudaMemcpy(S.p, S_host, sizeof(double) * (kNumSlots + 2 * (nC + 1)), cudaMemcpyHostToDevice));
}

void Solver::set_cone(long long i, const lb2_int *beg, const lb2_int *idx, const double *elem) {
    if (i < 0 || i >= nCones) throw std::invalid_argument("cone index out of range");
    // the arrays come straight from the caller (reader format, lorads_file_io.c:325-340): column pointers over
    // columns 0..m ascending from 0, packed lower-triangular positions inside [0, n(n+1)/2)
    if (!beg) throw std::invalid_argument("cone data: null column pointer array");
    if (beg[0] != 0) throw std::invalid_argument("cone data: column pointers must start at 0");
    for (long long c = 0; c <= m; ++c)
        if (beg[c + 1] < beg[c]) throw std::invalid_argument("cone data: column pointers must not decrease");
    if (beg[m + 1] > 0 && (!idx || !elem)) throw std::invalid_argument("cone data: null index / value array with non-zeros present");
    {
        const long long nn = blkDims[i];
        const long long packed = nn * (nn + 1) / 2;
        for (long long k = 0; k < beg[m + 1]; ++k)
            if (idx[k] < 0 || idx[k] >= packed) throw std::invalid_argument("cone data: packed index outside [0, n(n+1)/2)");
    }
    ConeInput &in = inputs[i];
    in.beg.assign(beg, beg + m + 2);
    in.idx.assign(idx, idx + beg[m + 1]);
    in.elem.assign(elem, elem + beg[m + 1]);
    in.set = true;
}

void Solver::set_lp(long long nLpCols, const lb2_int *beg, const lb2_int *idx, const double *elem) {
    if (preprocessed) throw std::logic_error("set the LP block before lb2_preprocess");
    if (nLpCols < 0) throw std::invalid_argument("negative number of LP columns");
    nLp = nLpCols;
    lp_input = ConeInput();
    if (nLp == 0) return;
    lp_input.beg.assign(beg, beg + m + 2);
    lp_input.idx.assign(idx, idx + beg[m + 1]);
    lp_input.elem.assign(elem, elem + beg[m + 1]);
    lp_input.set = true;
}

void Solver::build_lp() {
    // lp_cone_proc + lp_cone_presolve, lorads_lp_conic.c:14-113.  Input: CSC by constraint (column 0 = objective,
    // column i+1 = constraint i), entries (LP column, value).  Output: objective vector, the constraint matrix by
    // constraint (CSR) and by LP column (CSC, rows ascending as the reference builds them), |a_j|^2.
    // Deviation (DESIGN.md "reference quirks"): the reference's dense column type (nnz >= m/4) reads an
    // uninitialised row count and ends up as an all-zero column; here every column keeps its entries.
    if (!lp_input.set) throw std::logic_error("LP data missing");
    const std::vector<int64_t> &beg = lp_input.beg, &idx = lp_input.idx;
    const std::vector<double> &val = lp_input.elem;
    lp_c_h.assign((size_t)nLp, 0.0);
    for (int64_t k = beg[0]; k < beg[1]; ++k) {
        if (idx[k] < 0 || idx[k] >= nLp) throw std::invalid_argument("LP column index out of range");
        lp_c_h[(size_t)idx[k]] = val[k];
    }
    const int64_t nnz = beg[m + 1] - beg[1];
    if (nnz > 0x7fffffffLL) throw std::invalid_argument("LP block too large");
    std::vector<int> rbeg((size_t)m + 1), rcol((size_t)nnz), cbeg((size_t)nLp + 1, 0), crow((size_t)nnz);
    std::vector<double> rval((size_t)nnz), cval((size_t)nnz), n2((size_t)nLp, 0.0);
    for (long long i = 0; i <= m; ++i) rbeg[(size_t)i] = (int)(beg[(size_t)i + 1] - beg[1]);
    for (int64_t k = 0; k < nnz; ++k) {
        const int64_t j = idx[(size_t)(beg[1] + k)];
        if (j < 0 || j >= nLp) throw std::invalid_argument("LP column index out of range");
        rcol[(size_t)k] = (int)j; rval[(size_t)k] = val[(size_t)(beg[1] + k)];
        cbeg[(size_t)j + 1]++;
    }
    for (long long j = 0; j < nLp; ++j) cbeg[(size_t)j + 1] += cbeg[(size_t)j];
    {
        std::vector<int> cur(cbeg.begin(), cbeg.end() - 1);
        for (long long i = 0; i < m; ++i)
            for (int k = rbeg[(size_t)i]; k < rbeg[(size_t)i + 1]; ++k) {
                const int q = cur[(size_t)rcol[(size_t)k]]++;
                crow[(size_t)q] = (int)i; cval[(size_t)q] = rval[(size_t)k];
            }
    }
    for (long long j = 0; j < nLp; ++j) {
        // nrm2 then squared, as lp_cone_presolve does (lorads_lp_conic.c:97-98)
        double ss = 0.0;
        for (int k = cbeg[(size_t)j]; k < cbeg[(size_t)j + 1]; ++k) ss += cval[(size_t)k] * cval[(size_t)k];
        const double nr = std::sqrt(ss);
        n2[(size_t)j] = nr * nr;
        if (cbeg[(size_t)j] == cbeg[(size_t)j + 1])
            printf("File [%30s] Line [%d]\n", "lorads_b200 lp column without entries", (int)j);   // LP_COEFF_ZERO trace
    }
    // level schedule of the Gauss-Seidel sweep: a column waits for every earlier column that shares a row with it
    std::vector<int> row_level((size_t)m, 0), level((size_t)nLp, 0);
    int n_lvl = 0;
    for (long long j = 0; j < nLp; ++j) {
        int lv = 0;
        for (int k = cbeg[(size_t)j]; k < cbeg[(size_t)j + 1]; ++k) lv = std::max(lv, row_level[(size_t)crow[(size_t)k]]);
        level[(size_t)j] = lv;
        for (int k = cbeg[(size_t)j]; k < cbeg[(size_t)j + 1]; ++k) row_level[(size_t)crow[(size_t)k]] = lv + 1;
        n_lvl = std::max(n_lvl, lv + 1);
    }
    std::vector<int> lptr((size_t)n_lvl + 1, 0), lcol((size_t)nLp);
    for (long long j = 0; j < nLp; ++j) lptr[(size_t)level[(size_t)j] + 1]++;
    for (int l = 0; l < n_lvl; ++l) lptr[(size_t)l + 1] += lptr[(size_t)l];
    {
        std::vector<int> cur(lptr.begin(), lptr.end() - 1);
        for (long long j = 0; j < nLp; ++j) lcol[(size_t)cur[(size_t)level[(size_t)j]]++] = (int)j;
    }
    // launches of the sweep: a level of 64 columns or more (over eight per warp of one CTA) gets its own multi-CTA launch,
    // runs of narrower levels share a one-CTA launch
    lp_segs.clear();
    for (int l = 0; l < n_lvl; ++l) {
        const int cols = lptr[(size_t)l + 1] - lptr[(size_t)l];
        if (cols >= 64) lp_segs.push_back(LpSeg{l, l + 1, std::min((cols + 7) / 8, 8 * ctx.num_sms)});
        else if (!lp_segs.empty() && lp_segs.back().blocks == 1) lp_segs.back().lv_hi = l + 1;
        else lp_segs.push_back(LpSeg{l, l + 1, 1});
    }
    lp_c.upload(lp_c_h); lp_rbeg.upload(rbeg); lp_rcol.upload(rcol); lp_rval.upload(rval);
    lp_cbeg.upload(cbeg); lp_crow.upload(crow); lp_cval.upload(cval); lp_nrm2sq.upload(n2);
    lp_lvl_ptr.upload(lptr); lp_lvl_col.upload(lcol);
    lp_x.alloc((size_t)nLp);
    lp.n = nLp; lp.m = m; lp.c = lp_c.p; lp.rbeg = lp_rbeg.p; lp.rcol = lp_rcol.p; lp.rval = lp_rval.p;
    lp.cbeg = lp_cbeg.p; lp.crow = lp_crow.p; lp.cval = lp_cval.p; lp.nrm2sq = lp_nrm2sq.p;
    lp.lvl_ptr = lp_lvl_ptr.p; lp.lvl_col = lp_lvl_col.p; lp.n_lvl = n_lvl;
    lp_input = ConeInput();
}

void Solver::preprocess() {
    LB2_CUDA(cudaSetDevice(device));
    cones.clear();
    cones.resize(nCones);
    cObjNrm1 = 0; cObjNrmInf = 0;
    double n2 = 0;
    for (long long c = 0; c < nCones; ++c) {
        if (!inputs[c].set) throw std::logic_error("cone data missing: call lb2_set_cone_data for every cone");
        ConeLayout L = build_cone_layout(blkDims[c], m, inputs[c].beg.data(), inputs[c].idx.data(), inputs[c].elem.data());
        if (L.c_rank1 != 0.0 && L.dense_path)   // the dense scratch kernels take C as it is
            L = build_cone_layout(blkDims[c], m, inputs[c].beg.data(), inputs[c].idx.data(), inputs[c].elem.data(), false);
        ConeDev &K = cones[c];
        K.c_rank1 = L.c_rank1;
        if (K.c_rank1 != 0.0) { K.csA.alloc(256); K.csB.alloc(256); K.cs_scratch.alloc((size_t)4 * ctx.num_sms * 256); }
        K.n = L.n; K.dense_path = L.dense_path; K.dense_cone = L.dense_cone; K.n_act = L.n_act;
        K.np = L.psize(); K.nnzA = L.nnzA; K.nnzC = L.nnzC; K.n_nonzero_coeff = L.n_nonzero_coeff;
        K.cNrm1 = L.cNrm1; K.cNrm2Sq = L.cNrm2Sq; K.cNrmInf = L.cNrmInf;
        K.act_idx_h = L.act_idx;
        K.identity_act = (L.n_act == m);
        K.act_idx.upload(L.act_idx);
        K.listA.upload(L.listA);
        K.listAC.upload(L.listAC);
        K.obj_item_begin = L.listAC.ptr[L.n_act];
        // the carry buffers serve every item list of the cone (lists of different length use different tile sizes)
        const size_t ntile = (size_t)std::max<long long>(
            1, std::max<long long>(std::max<long long>(K.listAC.dev.n_tiles, K.listA.dev.n_tiles), L.vc.listRes.n_tiles()));
        K.carry1.alloc(2 * ntile); K.carry2.alloc(2 * ntile); K.carry3.alloc(2 * ntile);
        K.C_onP.upload(L.C_onP);
        K.S.alloc((size_t)K.np);
        K.T_ptr.upload(L.T_ptr); K.T_con.upload(L.T_con); K.T_val.upload(L.T_val);
        if (K.dense_path) {
            K.D_pos.upload(L.D_pos);
            K.n_pos = (long long)L.D_pos.size();
            K.Z1.alloc((size_t)K.np); K.Z2.alloc((size_t)K.np);
        } else {
            K.adj_ptr.upload(L.adj_ptr); K.adj_col.upload(L.adj_col); K.adj_pos.upload(L.adj_pos);
            K.P_row_h = L.P_row; K.P_col_h = L.P_col;
            K.vc_on = L.vc.on;
            if (K.vc_on) {
                const VcLayout &V = L.vc;
                K.vc_nnz_res = V.nnz_res;
                K.nnzC_adj = 0;
                for (int32_t t : V.u_tag) K.nnzC_adj += (t == -1);
                K.vc_order.upload(V.order); K.vc_order_l.upload(V.order_l);
                K.vc_order_h = V.order; K.vc_order_l_h = V.order_l;
                K.row_weight_h.resize((size_t)K.n);
                for (long long i = 0; i < K.n; ++i)
                    K.row_weight_h[(size_t)i] = 8 + (V.u_ptr[i + 1] - V.u_ptr[i]) + 2 * (V.l_ptr[i + 1] - V.l_ptr[i]);
                K.vc.n = K.n; K.vc.obj_row = (int)K.n_act;
                K.vc.order = K.vc_order.p; K.vc.order_l = K.vc_order_l.p;
                K.vc_u_ptr.upload(V.u_ptr); K.vc_u_mid.upload(V.u_mid); K.vc_u_col.upload(V.u_col); K.vc_u_tag.upload(V.u_tag);
                K.vc_u_val.upload(V.u_val);
                K.vc.u_ptr = K.vc_u_ptr.p; K.vc.u_mid = K.vc_u_mid.p; K.vc.u_col = K.vc_u_col.p; K.vc.u_tag = K.vc_u_tag.p;
                K.vc.u_val = K.vc_u_val.p;
                if (V.n_single > 0) {
                    K.vc_d_con.upload(V.d_con); K.vc_d_coef.upload(V.d_coef);
                    K.vc.d_con = K.vc_d_con.p; K.vc.d_coef = K.vc_d_coef.p;
                }
                if (!V.l_row.empty()) {
                    K.vc_l_ptr.upload(V.l_ptr); K.vc_l_row.upload(V.l_row); K.vc_l_col.upload(V.l_col); K.vc_l_con.upload(V.l_con);
                    K.vc_l_coef.upload(V.l_coef);
                    K.vc.l_ptr = K.vc_l_ptr.p; K.vc.l_row = K.vc_l_row.p; K.vc.l_col = K.vc_l_col.p; K.vc.l_con = K.vc_l_con.p;
                    K.vc.l_coef = K.vc_l_coef.p;
                    K.vc.l_lo = 0; K.vc.l_hi = (long long)V.l_row.size();
                    K.vc_l_ptr_h = V.l_ptr;
                }
                if (V.nnz_res > 0) {
                    K.vc_Tr_ptr.upload(V.Tr_ptr); K.vc_Tr_con.upload(V.Tr_con); K.vc_Tr_val.upload(V.Tr_val);
                    K.vc_listRes.upload(V.listRes);
                }
            }
        }
        K.row_lo = 0; K.row_hi = K.n; K.lead = true;
        K.cv.alloc((size_t)K.n_act + 1); K.t1.alloc((size_t)K.n_act + 1); K.t2.alloc((size_t)K.n_act + 1);
        cObjNrm1 += K.cNrm1; n2 += K.cNrm2Sq; cObjNrmInf = std::max(cObjNrmInf, K.cNrmInf);
        inputs[c] = ConeInput();   // the reader arrays are no longer needed
    }
    if (nLp > 0) {
        build_lp();
        // LORADSNrm1Obj / Nrm2Obj / NrmInfObj, lorads_solver.c:149-183, lorads_alg_common.c:300-316: the LP block
        // contributes |c|_1 to the 1-norm, |c|_1^2 (sic, lp_cone_obj_nrm2Square) to the squared 2-norm and, in the
        // -DUNDER_BLAS build, the entry after the first largest one (1-based idamax_ used 0-based) to the inf-norm
        double l1 = 0.0, best = -1.0;
        long long arg = 0;
        for (long long j = 0; j < nLp; ++j) {
            l1 += std::fabs(lp_c_h[(size_t)j]);
            if (std::fabs(lp_c_h[(size_t)j]) > best) { best = std::fabs(lp_c_h[(size_t)j]); arg = j; }
        }
        cObjNrm1 += l1; n2 += l1 * l1;
        cObjNrmInf = std::max(cObjNrmInf, std::fabs(lp_c_h[(size_t)std::min<long long>(arg + 1, nLp - 1)]));
    }
    cObjNrm2 = std::sqrt(n2);
    bNrm1 = 0; bNrmInf = 0; double b2 = 0;
    for (double v : b_h) { bNrm1 += std::fabs(v); b2 += v * v; }
    bNrm2 = std::sqrt(b2);
    {
        // cal_sdp_const, lorads_solver.c:1059-1066.  The reference's Linux build (-DUNDER_BLAS) indexes rowRHS
        // with the 1-based result of Fortran idamax_, i.e. it reads the element AFTER the first entry of largest
        // magnitude.  The value only scales the reported / tested "Inf" infeasibility, but it steers the phase
        // exits, so it is reproduced here (clamped to the last element); see DESIGN.md "reference quirks".
        long long arg = 0;
        double best = -1.0;
        for (long long i = 0; i < m; ++i)
            if (std::fabs(b_h[i]) > best) { best = std::fabs(b_h[i]); arg = i; }
        bNrmInf = std::fabs(b_h[std::min<long long>(arg + 1, m - 1)]);
    }
    single_identity = (nCones == 1 && cones[0].identity_act);
    b.alloc((size_t)m + 1); lam.alloc((size_t)m + 1); s.alloc((size_t)m + 1);
    const size_t mpad = ((size_t)m + 2) & ~(size_t)1;     // even stride keeps q2 16-byte aligned
    q12.alloc(3 * mpad); q1.p = q12.p; q2.p = q12.p + mpad; q3.p = q12.p + 2 * mpad;
    M1.alloc((size_t)m + 1); cvfull.alloc((size_t)m + 1);
    LB2_CUDA(cudaMemcpy(b.p, b_h.data(), sizeof(double) * m, cudaMemcpyHostToDevice));
    preprocessed = true;
}

void Solver::determine_rank(double timesRank) {
    // LORADSDetermineRank, lorads_solver.c:290-319
    if (!preprocessed) throw std::logic_error("preprocess first");
    rank.assign(nCones, 1); rank_max.assign(nCones, 1);
    for (long long c = 0; c < nCones; ++c) {
        lb2_int cap = 0;
        rank[c] = lb2_host_rank_rule(blkDims[c], cones[c].n_nonzero_coeff, nCones, timesRank, &cap);
        // the gather kernels hold a factor row in at most 8 passes of 32 columns: the rank (and with it the growth of
        // AUG_RANK) stops at that width; reaching it counts as "rank at its maximum"
        rank_max[c] = std::min<long long>(cap, kMaxRank);
        rank[c] = std::min<long long>(rank[c], rank_max[c]);
    }
}

void Solver::alloc_vars() {
    drop_graphs();      // captured launches hold the old buffer addresses
    N = 0;
    for (long long c = 0; c < nCones; ++c) {
        ConeDev &K = cones[c];
        K.r = (int)my_cols[c].size();
        K.ld = std::max(4, ((K.r + 3) / 4) * 4);
        K.off = N;
        N += K.n * K.ld;
    }
    Nt = N + nLp;
    compute_owned_ranges();
    Nalloc = Nt;
    if (shard_rows() && equal_rows > 0) Nalloc = std::max<long long>(Nt, equal_rows * world * (long long)cones[0].ld);
    for (DBuf<double> *v : {&R, &U, &V, &G, &M2, &Bls, &cg_r, &cg_p, &cg_Q, &Dtemp}) { retire(*v); v->alloc((size_t)Nalloc); }
    {
        // split-K scratch of the dense symmetric product: up to 32 partial n x ldp blocks of the largest dense cone
        size_t need = 0;
        for (const ConeDev &K : cones)
            if (K.dense_path) need = std::max(need, (size_t)32 * (size_t)K.n * (size_t)(((K.ld + 7) / 8) * 8));
        need = std::min(need, (size_t)1 << 28);     // cap at 2 GiB of doubles; the launcher lowers the split count
        if (need > dense_part.n) { retire(dense_part); dense_part.alloc(need, false); }
        ctx.dense_part = dense_part.p;
        ctx.dense_part_cap = need;
    }
    for (DBuf<double> &v : lb_s) retire(v);
    for (DBuf<double> &v : lb_y) retire(v);
    lb_s.clear(); lb_y.clear();
    lb_s.resize(lbfgs_len); lb_y.resize(lbfgs_len);
    for (int k = 0; k < lbfgs_len; ++k) { lb_s[k].alloc((size_t)Nalloc); lb_y[k].alloc((size_t)Nalloc); }
    lb_head = 0;
}

static void assign_columns(std::vector<int> &cols, long long r, int world, int myrank) {
    cols.clear();
    for (long long k = 0; k < r; ++k)
        if ((int)(k % world) == myrank) cols.push_back((int)k);
}

// ---------------------------------------------------------------------------------------------------
// row sharding: partition of the concatenated rows (by cone block, huge blocks by row slabs)
// ---------------------------------------------------------------------------------------------------
void Solver::setup_row_partition() {
    // weight of a row = its share of the gather work; cones that cannot be split (dense scratch, generic item path)
    // count as one lump and go to the rank their centre of mass falls into
    std::vector<double> cone_w((size_t)nCones, 0.0);
    double total = 0.0;
    for (long long c = 0; c < nCones; ++c) {
        const ConeDev &K = cones[c];
        double w = 0.0;
        if (K.vc_on) for (int32_t x : K.row_weight_h) w += x;
        else w = (double)K.np * 4.0 + (double)K.n * 8.0;
        cone_w[(size_t)c] = w; total += w;
    }
    part_rows.assign((size_t)world + 1, 0);
    long long rows_total = 0;
    for (long long c = 0; c < nCones; ++c) rows_total += cones[c].n;
    part_rows[(size_t)world] = rows_total;
    equal_rows = 0;
    if (nCones == 1 && cones[0].vc_on && getenv("LORADS_B200_UNEQUAL_SLABS") == nullptr) {
        // one big block: slabs of equal length (the work per row of these patterns is uniform enough) so that the
        // all-gather of a factor vector is ONE in-place ncclAllGather instead of a group of broadcasts
        equal_rows = (rows_total + world - 1) / world;
        for (int k = 1; k < world; ++k) part_rows[(size_t)k] = std::min<long long>(rows_total, equal_rows * k);
    } else {
        double acc = 0.0;
        long long row0 = 0;
        int next = 1;
        for (long long c = 0; c < nCones && next < world; ++c) {
            const ConeDev &K = cones[c];
            if (K.vc_on) {
                for (long long i = 0; i < K.n && next < world; ++i) {
                    acc += K.row_weight_h[(size_t)i];
                    while (next < world && acc >= total * next / world) part_rows[(size_t)next++] = row0 + i + 1;
                }
                if (next >= world) break;
            } else {
                const double mid = acc + 0.5 * cone_w[(size_t)c];
                acc += cone_w[(size_t)c];
                // every boundary that falls inside the lump moves to the nearer end of the cone
                while (next < world && acc >= total * next / world)
                    part_rows[(size_t)next] = (total * next / world <= mid) ? row0 : row0 + K.n, ++next;
            }
            row0 += K.n;
        }
        for (; next < world; ++next) part_rows[(size_t)next] = rows_total;
        for (int k = 1; k <= world; ++k) part_rows[(size_t)k] = std::max(part_rows[(size_t)k], part_rows[(size_t)k - 1]);
    }
    // this rank's rows of every cone, its filtered row orders, the lead rank of every cone
    long long row0 = 0;
    for (long long c = 0; c < nCones; ++c) {
        ConeDev &K = cones[c];
        const long long lo = part_rows[(size_t)myrank], hi = part_rows[(size_t)myrank + 1];
        K.row_lo = std::min(std::max(lo - row0, 0LL), K.n);
        K.row_hi = std::min(std::max(hi - row0, 0LL), K.n);
        int lead_rank = 0;
        while (lead_rank + 1 < world && part_rows[(size_t)lead_rank + 1] <= row0) ++lead_rank;     // owner of the cone's row 0
        K.lead = (lead_rank == myrank);
        if (!K.vc_on && K.row_hi > K.row_lo && (K.row_lo != 0 || K.row_hi != K.n))
            throw std::logic_error("row partition split a cone that cannot be split");
        if (K.vc_on) {
            std::vector<int32_t> o, ol;
            for (int32_t i : K.vc_order_h) if (i >= K.row_lo && i < K.row_hi) o.push_back(i);
            for (int32_t i : K.vc_order_l_h) if (i >= K.row_lo && i < K.row_hi) ol.push_back(i);
            if (o.empty()) { o.push_back(0); ol.push_back(0); }      // keep the device arrays non-null
            K.vc_order.upload(o); K.vc_order_l.upload(ol);
            K.vc.order = K.vc_order.p; K.vc.order_l = K.vc_order_l.p;
            K.vc.n = K.row_hi - K.row_lo;
            if (!K.vc_l_ptr_h.empty()) {      // lowA is sorted by column: the owned columns are one contiguous range
                K.vc.l_lo = K.vc_l_ptr_h[(size_t)K.row_lo];
                K.vc.l_hi = K.vc_l_ptr_h[(size_t)K.row_hi];
            }
        }
        row0 += K.n;
    }
}

void Solver::compute_owned_ranges() {
    vo = 0; vn = Nt;
    own_off.assign((size_t)world, 0); own_cnt.assign((size_t)world, 0);
    if (!shard_rows()) return;
    if (part_rows.empty()) throw std::logic_error("row partition missing");
    // element offset of a concatenated row index
    auto elem_of = [&](long long row) {
        long long row0 = 0;
        for (long long c = 0; c < nCones; ++c) {
            const ConeDev &K = cones[c];
            if (row < row0 + K.n) return K.off + (row - row0) * (long long)K.ld;
            row0 += K.n;
        }
        return N;
    };
    for (int k = 0; k < world; ++k) {
        own_off[(size_t)k] = elem_of(part_rows[(size_t)k]);
        own_cnt[(size_t)k] = elem_of(part_rows[(size_t)k + 1]) - own_off[(size_t)k];
    }
    vo = own_off[(size_t)myrank]; vn = own_cnt[(size_t)myrank];
}

void Solver::allgather_owned(double *X) { if (shard_rows()) bcast_ranges(X, 0, N); }

void Solver::allgather_cone(const ConeDev &K, double *X) {
    if (shard_rows()) bcast_ranges(X, K.off, K.off + K.n * (long long)K.ld);
}

void Solver::init_vars(long long lbfgsLen, double initRho) {
    if (rank.empty()) throw std::logic_error("determine the rank first");
    if (lbfgsLen < 1 || lbfgsLen > kMaxLbfgs) throw std::invalid_argument("lbfgsListLength must be in 1..16");
    LB2_CUDA(cudaSetDevice(device));
    lbfgs_len = (int)lbfgsLen;
    my_cols.resize(nCones);
    // column sharding deals the factor columns round-robin; row sharding (and one GPU) keeps every column everywhere
    for (long long c = 0; c < nCones; ++c) assign_columns(my_cols[c], rank[c], shard_cols() ? world : 1, shard_cols() ? myrank : 0);
    alloc_vars();
    // the reference's libc draw order: srand(925); R of every cone (lorads_solver.c:415-446), later U then V of
    // every cone (lorads_solver.c:631-658); each element is rand()/RAND_MAX - rand()/RAND_MAX (:361-370)
    // glibc's rand() is random() on a 128-byte TYPE_3 state behind a lock; random_r on a private state of the same
    // size yields the identical stream without the lock (1.7e8 draws at n = 1e6, r = 28)
    struct random_data rd;
    char rstate[128];
    std::memset(&rd, 0, sizeof(rd));
    initstate_r(925, rstate, sizeof(rstate), &rd);
    auto draw = [&rd](std::vector<double> &x) {
        int32_t a, b;
        for (double &v : x) {
            random_r(&rd, &a); random_r(&rd, &b);
            v = (double)a / RAND_MAX;
            v -= (double)b / RAND_MAX;
        }
    };
    std::vector<double> tmp;
    for (long long c = 0; c < nCones; ++c) {
        tmp.resize((size_t)(blkDims[c] * rank[c]));
        draw(tmp);
        upload_factor(R.p, cones[c], tmp.data());
    }
    if (nLp > 0) {
        // rLp right after the cone factors (lorads_solver.c:458-466); uLp, vLp open LORADSInitADMMVars (:593-605)
        tmp.resize((size_t)nLp);
        for (double *dst : {R.p, U.p, V.p}) {
            draw(tmp);
            LB2_CUDA(cudaMemcpy(dst + N, tmp.data(), sizeof(double) * nLp, cudaMemcpyHostToDevice));
        }
    }
    for (long long c = 0; c < nCones; ++c) {
        tmp.resize((size_t)(blkDims[c] * rank[c]));
        draw(tmp);
        upload_factor(U.p, cones[c], tmp.data());
        draw(tmp);
        upload_factor(V.p, cones[c], tmp.data());
    }
    LB2_CUDA(cudaMemset(lam.p, 0, sizeof(double) * (m + 1)));
    // initial_solver_state, lorads_solver.c:1148-1170
    double rho;
    if (initRho == 0) {
        long long sum = 0;
        for (long long c = 0; c < nCones; ++c) sum += blkDims[c];
        rho = 1.0 / std::sqrt((double)sum);
    } else rho = initRho;
    alm = AlmState(); alm.rho = rho;
    admm = AdmmState(); admm.rho = rho; admm.nBlks = nCones;
    scaleObjHis = 1.0;
    cgIter = 0;
    vars_ready = true;
}

// ---------------------------------------------------------------------------------------------------
// transfers
// ---------------------------------------------------------------------------------------------------
double *Solver::factor_ptr(char which) {
    switch (which) {
    case 'R': return R.p; case 'U': return U.p; case 'V': return V.p; case 'G': return G.p;
    case 'M': return M2.p; case 'B': return Bls.p;
    }
    throw std::invalid_argument("unknown factor name");
}

double *Solver::vec_ptr(char which) {
    switch (which) {
    case 'l': return lam.p; case 's': return s.p; case 'b': return b.p; case 'm': return M1.p;
    case 'q': return q1.p; case 'Q': return q2.p; case 'v': return cvfull.p;
    }
    throw std::invalid_argument("unknown vector name");
}

// The caller's column-major buffer goes to the device as it is (one copy when this rank holds every column, one
// per owned column under column sharding) and is transposed there; the copy is asynchronous when the buffer is
// pinned.  The staging buffer is reused across calls.
void Solver::upload_factor(double *dst, const ConeDev &K, const double *colMajor) {
    const long long c = &K - cones.data();
    const std::vector<int> &cols = my_cols[c];
    const size_t need = (size_t)K.n * (size_t)std::max(K.r, 1);
    if (xfer_stage.n < need) { retire(xfer_stage); xfer_stage.alloc(need, false); }
    if (K.r == rank[c]) {
        LB2_CUDA(cudaMemcpyAsync(xfer_stage.p, colMajor, sizeof(double) * need, cudaMemcpyHostToDevice, ctx.stream));
    } else {
        for (int k = 0; k < K.r; ++k)
            LB2_CUDA(cudaMemcpyAsync(xfer_stage.p + (size_t)k * K.n, colMajor + (size_t)cols[k] * K.n, sizeof(double) * K.n,
                                     cudaMemcpyHostToDevice, ctx.stream));
    }
    launch_cm_to_rm(ctx, K.n, K.r, K.ld, xfer_stage.p, dst + K.off);
    LB2_CUDA(cudaStreamSynchronize(ctx.stream));
}

void Solver::download_factor(const double *src, const ConeDev &K, double *colMajor) const {
    Solver &self = const_cast<Solver &>(*this);
    const long long c = &K - cones.data();
    const std::vector<int> &cols = my_cols[c];
    const size_t need = (size_t)K.n * (size_t)std::max(K.r, 1);
    if (self.xfer_stage.n < need) { self.retire(self.xfer_stage); self.xfer_stage.alloc(need, false); }
    launch_rm_to_cm(self.ctx, K.n, K.r, K.ld, src + K.off, self.xfer_stage.p);
    if (K.r == rank[c]) {
        LB2_CUDA(cudaMemcpyAsync(colMajor, xfer_stage.p, sizeof(double) * need, cudaMemcpyDeviceToHost, ctx.stream));
    } else {
        // sharded: the columns held elsewhere read as zero (the caller sums or gathers over ranks)
        std::memset(colMajor, 0, sizeof(double) * (size_t)(K.n * rank[c]));
        for (int k = 0; k < K.r; ++k)
            LB2_CUDA(cudaMemcpyAsync(colMajor + (size_t)cols[k] * K.n, xfer_stage.p + (size_t)k * K.n, sizeof(double) * K.n,
                                     cudaMemcpyDeviceToHost, ctx.stream));
    }
    LB2_CUDA(cudaStreamSynchronize(ctx.stream));
}

void Solver::set_factor(char which, long long c, const double *colMajor) { upload_factor(factor_ptr(which), cones.at(c), colMajor); }
void Solver::get_factor(char which, long long c, double *colMajor) const {
    Solver &self = const_cast<Solver &>(*this);
    // row sharding: R, U, V are replicated; the others hold only this rank's rows -- complete them first (collective:
    // every rank has to ask for the same factor)
    if (shard_rows() && which != 'R' && which != 'U' && which != 'V') self.allgather_cone(cones.at(c), self.factor_ptr(which));
    download_factor(self.factor_ptr(which), cones.at(c), colMajor);
}

void Solver::sync() { LB2_CUDA(cudaStreamSynchronize(ctx.stream)); }

// Host wait on the per-iteration round trips.  cudaStreamSynchronize follows the context's scheduling flag, and the
// default ("auto") turns into a blocking wait with tens of microseconds of wake-up latency when the process sees few
// CPU cores (containers, eight ranks on one node): the step time then varied by +-25 % from box to box with identical
// kernel and collective times.  Polling the stream keeps the latency at one driver call on every box.
// LORADS_B200_BLOCKING_SYNC=1 restores the blocking wait.
static void spin_sync(cudaStream_t s) {
    static const bool blocking = getenv("LORADS_B200_BLOCKING_SYNC") != nullptr;
    if (blocking) { LB2_CUDA(cudaStreamSynchronize(s)); return; }
    cudaError_t e;
    while ((e = cudaStreamQuery(s)) == cudaErrorNotReady) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    LB2_CUDA(e);
}

void Solver::read_slots_side() {
    if (!side_stream) {
        LB2_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
        LB2_CUDA(cudaEventCreateWithFlags(&side_event, cudaEventDisableTiming));
    }
    LB2_CUDA(cudaEventRecord(side_event, ctx.stream));
    LB2_CUDA(cudaStreamWaitEvent(side_stream, side_event, 0));
    LB2_CUDA(cudaMemcpyAsync(S_host, S.p, sizeof(double) * (kNumSlots + 2 * (nCones + 1)), cudaMemcpyDeviceToHost, side_stream));
}

void Solver::read_slots() {
    LB2_CUDA(cudaMemcpyAsync(S_host, S.p, sizeof(double) * (kNumSlots + 2 * (nCones + 1)), cudaMemcpyDeviceToHost, ctx.stream));
    spin_sync(ctx.stream);
}

// ---------------------------------------------------------------------------------------------------
// operators
// ---------------------------------------------------------------------------------------------------
void Solver::cone_auv(ConeDev &K, bool with_obj, const double *Um, const double *Vm, bool same, double scale, double *out,
                      double *obj) {
    ItemListBufs &L = with_obj ? K.listAC : K.listA;
    const bool rows = shard_rows();
    // row sharding: every rank adds its part into zeroed outputs, the all-reduce below completes them; the parts that
    // cannot be split are evaluated by the cone's lead rank only
    if (rows) LB2_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)L.dev.n_rows, ctx.stream));
    if (K.dense_path) {
        // <C,Z> is taken inside the tile kernel from this rank's share of Z (the slot is summed over ranks by the
        // caller); the constraint rows come from the A-only item list over the all-reduced Z
        if (!rows || K.lead) {
            launch_dense_uvt(ctx, K.n, K.r, K.ld, Um + K.off, Vm + K.off, K.Z1.p, same, with_obj ? K.C_onP.p : nullptr, scale,
                             with_obj ? obj : nullptr);
            if (shard_cols()) allreduce(K.Z1.p, K.np);
            launch_auv(ctx, AUV_FROMZ, K.listA.dev, K.Z1.p, nullptr, K.ld, scale, 0.0, out, nullptr, K.carry1.p, nullptr, nullptr,
                       nullptr);
        }
        if (rows) allreduce(out, K.listA.dev.n_rows);
        return;
    }
    // with sharding the objective is summed over ranks by the caller (the slot is all-reduced once)
    if (K.vc_on) {
        // multi-entry constraints through the item kernel (it clears the rows it does not own), then the singleton
        // rows and the objective row from the vertex-centric pass
        if (K.vc_nnz_res > 0 && (!rows || K.lead))
            launch_auv(ctx, same ? AUV_SAME : AUV_PAIR, K.vc_listRes.dev, Um + K.off, Vm + K.off, K.ld, scale, 0.0, out, nullptr,
                       K.carry1.p, nullptr, nullptr, nullptr);
        if (K.vc.n > 0)
            launch_vc_auv(ctx, same ? AUV_SAME : AUV_PAIR, K.vc, K.ld, with_obj, Um + K.off, Vm + K.off, scale, 0.0, out, nullptr,
                          nullptr, obj, nullptr);
    } else if (!rows || K.lead)
        launch_auv(ctx, same ? AUV_SAME : AUV_PAIR, L.dev, Um + K.off, Vm + K.off, K.ld, scale, 0.0, out, nullptr,
                   K.carry1.p, nullptr, obj, nullptr);
    if (with_obj && obj && K.c_rank1 != 0.0 && (!rows || K.lead)) {
        // <c ee^T, sym(U V^T)> = c (e^T U)(V^T e)
        launch_colsum(ctx, K.n, K.ld, Um + K.off, K.csA.p, K.cs_scratch.p);
        if (!same) launch_colsum(ctx, K.n, K.ld, Vm + K.off, K.csB.p, K.cs_scratch.p);
        launch_rank1_obj(ctx, K.ld, K.csA.p, same ? K.csA.p : K.csB.p, scale * K.c_rank1, obj, 0.0, nullptr);
    }
    if (world > 1) allreduce(out, L.dev.n_rows);
}

void Solver::cone_auv_dual(ConeDev &K, const double *Rm, const double *Dm, double *out1, double *out2, double *obj1,
                           double *obj2) {
    ItemListBufs &L = K.listAC;
    const bool rows = shard_rows();
    if (rows) {
        LB2_CUDA(cudaMemsetAsync(out1, 0, sizeof(double) * (size_t)L.dev.n_rows, ctx.stream));
        LB2_CUDA(cudaMemsetAsync(out2, 0, sizeof(double) * (size_t)L.dev.n_rows, ctx.stream));
    }
    if (K.dense_path) {
        if (!rows || K.lead) {
            launch_dense_uvt_dual(ctx, K.n, K.r, K.ld, Rm + K.off, Dm + K.off, K.Z1.p, K.Z2.p, K.C_onP.p, 2.0, 1.0, obj1, obj2);
            if (shard_cols()) { allreduce(K.Z1.p, K.np); allreduce(K.Z2.p, K.np); }
            launch_auv(ctx, AUV_FROMZ, K.listA.dev, K.Z1.p, nullptr, K.ld, 2.0, 0.0, out1, nullptr, K.carry1.p, nullptr, nullptr,
                       nullptr);
            launch_auv(ctx, AUV_FROMZ, K.listA.dev, K.Z2.p, nullptr, K.ld, 1.0, 0.0, out2, nullptr, K.carry1.p, nullptr, nullptr,
                       nullptr);
        }
        if (rows && !(out1 == q1.p && out2 == q2.p)) { allreduce(out1, K.listA.dev.n_rows); allreduce(out2, K.listA.dev.n_rows); }
        return;
    }
    if (K.vc_on) {
        if (K.vc_nnz_res > 0 && (!rows || K.lead))
            launch_auv(ctx, AUV_DUAL, K.vc_listRes.dev, Rm + K.off, Dm + K.off, K.ld, 2.0, 1.0, out1, out2, K.carry1.p, K.carry2.p,
                       nullptr, nullptr);
        if (K.vc.n > 0)
            launch_vc_auv(ctx, AUV_DUAL, K.vc, K.ld, true, Rm + K.off, Dm + K.off, 2.0, 1.0, out1, out2, nullptr, obj1, obj2);
    } else if (!rows || K.lead)
        launch_auv(ctx, AUV_DUAL, L.dev, Rm + K.off, Dm + K.off, K.ld, 2.0, 1.0, out1, out2, K.carry1.p, K.carry2.p, obj1, obj2);
    if (K.c_rank1 != 0.0 && obj1 && obj2 && (!rows || K.lead)) {
        launch_colsum(ctx, K.n, K.ld, Rm + K.off, K.csA.p, K.cs_scratch.p);
        launch_colsum(ctx, K.n, K.ld, Dm + K.off, K.csB.p, K.cs_scratch.p);
        launch_rank1_obj(ctx, K.ld, K.csA.p, K.csB.p, 2.0 * K.c_rank1, obj1, K.c_rank1, obj2);
    }
    if (world > 1 && !(out1 == q1.p && out2 == q2.p)) { allreduce(out1, L.dev.n_rows); allreduce(out2, L.dev.n_rows); }
}

void Solver::cone_wsum(ConeDev &K, const double *w, bool w_compact, bool addC, bool materialize) {
    const int *map = K.identity_act ? nullptr : K.act_idx.p;
    K.S_has_C = addC;
    if (K.vc_on && !materialize) {
        // nothing is written for C and the singleton constraints: cone_mul reads w directly.  Only the weighted sum
        // of the multi-entry constraints is materialised on the pattern.
        K.pw = w; K.pw_compact = w_compact;
        if (K.vc_nnz_res > 0)
            launch_wsum(ctx, K.S.p, K.np, K.C_onP.p, K.vc_Tr_ptr.p, K.vc_Tr_con.p, K.vc_Tr_val.p, w, map, w_compact, false);
        return;
    }
    K.pw = nullptr;
    if (K.dense_path)
        launch_dense_wsum(ctx, K.S.p, K.np, K.C_onP.p, K.D_pos.p, K.n_pos, K.T_ptr.p, K.T_con.p, K.T_val.p, w, map,
                          w_compact, addC);
    else
        launch_wsum(ctx, K.S.p, K.np, K.C_onP.p, K.T_ptr.p, K.T_con.p, K.T_val.p, w, map, w_compact, addC);
}

void Solver::cone_mul(ConeDev &K, const double *X, double a, double bcoef, const double *Z, const double *Z2, double *Y,
                      double *red) {
    if (shard_rows() && (K.row_hi == K.row_lo || (K.vc_on && K.vc.n == 0))) {
        // nothing of this cone lives here: only the reduction slots are cleared so that the all-reduce sums the owners
        if (red) LB2_CUDA(cudaMemsetAsync(red, 0, 2 * sizeof(double), ctx.stream));
        return;
    }
    if (K.dense_path)
        launch_dense_symm(ctx, K.n, K.r, K.ld, K.S.p, X + K.off, a, bcoef, Z ? Z + K.off : nullptr,
                          Z2 ? Z2 + K.off : nullptr, Y + K.off, red);
    else {
        const bool r1 = K.c_rank1 != 0.0 && K.S_has_C;      // (S + c ee^T) X = S X + c e (e^T X)
        if (r1) launch_colsum(ctx, K.n, K.ld, X + K.off, K.csA.p, K.cs_scratch.p);
        if (K.vc_on && K.pw) {
            const int *map = (K.pw_compact || K.identity_act) ? nullptr : K.act_idx.p;
            launch_vc_spmm(ctx, K.vc, K.ld, K.S_has_C, K.pw, map, K.vc_nnz_res > 0 ? K.S.p : nullptr, X + K.off, a, bcoef,
                           Z ? Z + K.off : nullptr, Z2 ? Z2 + K.off : nullptr, Y + K.off, red, r1 ? K.csA.p : nullptr, K.c_rank1);
            return;
        }
        launch_spmm(ctx, K.n, K.ld, K.adj_ptr.p, K.adj_col.p, K.adj_pos.p, K.S.p, X + K.off, a, bcoef,
                    Z ? Z + K.off : nullptr, Z2 ? Z2 + K.off : nullptr, Y + K.off, red, r1 ? K.csA.p : nullptr, K.c_rank1);
    }
}

void Solver::init_constr_val_all(const double *Um, const double *Vm, bool same) {
    for (ConeDev &K : cones) cone_auv(K, false, Um, Vm, same, 1.0, K.cv.p);
    // LORADSInitConstrValAllLP, lorads_alg_common.c:86-95: constrValLP[j] = a_j * (u_j v_j); only the product is kept
    if (nLp > 0) launch_lp_prod(ctx, nLp, Um + N, Vm + N, lp_x.p);
}

void Solver::constr_val_sum() {
    if (single_identity) {
        LB2_CUDA(cudaMemcpyAsync(s.p, cones[0].cv.p, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx.stream));
    } else {
        LB2_CUDA(cudaMemsetAsync(s.p, 0, sizeof(double) * m, ctx.stream));
        for (ConeDev &K : cones)
            launch_scatter_add(ctx, s.p, K.cv.p, K.identity_act ? nullptr : K.act_idx.p, K.n_act, 1.0, true, nullptr);
    }
    // LORADSInitConstrValSumLP, lorads_alg_common.c:144-158
    if (nLp > 0) launch_lp_rows(ctx, lp, lp_x.p, nullptr, 1.0, s.p, 0.0, nullptr);
}

void Solver::update_constr_val(long long c, const double *Um, const double *Vm) {
    cone_auv(cones[c], false, Um, Vm, false, 1.0, cones[c].cv.p);
}

void Solver::expand_cv_from(ConeDev &K, const double *src, double *full) {
    if (K.identity_act) {
        LB2_CUDA(cudaMemcpyAsync(full, src, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx.stream));
        return;
    }
    LB2_CUDA(cudaMemsetAsync(full, 0, sizeof(double) * m, ctx.stream));
    launch_scatter_add(ctx, full, src, K.act_idx.p, K.n_act, 1.0, false, nullptr);
}

void Solver::expand_cv(ConeDev &K, double *full) { expand_cv_from(K, K.cv.p, full); }

// G = 2 (C + A^*(M1)) R per cone with M1 already on the device; per-cone sum G.G lands in S[kNumSlots + 2c]
static void grad_from_M1(Solver &S_) {
    for (long long c = 0; c < S_.nCones; ++c) {
        ConeDev &K = S_.cones[c];
        S_.cone_wsum(K, S_.M1.p, false, true);
        S_.cone_mul(K, S_.R.p, 2.0, 0.0, nullptr, nullptr, S_.G.p, S_.S.p + kNumSlots + 2 * c);
    }
    // ALMSetGradLP, lorads_alm.c:56-78 (same multiplier vector M1 as the cones)
    if (S_.nLp > 0)
        launch_lp_grad(S_.ctx, S_.lp, S_.M1.p, S_.R.p + S_.N, S_.G.p + S_.N, S_.S.p + kNumSlots + 2 * S_.nCones);
}

static double sum_grad_sq(Solver &S_) {
    double t = 0.0;
    for (long long c = 0; c <= S_.nCones; ++c) t += S_.S_host[kNumSlots + 2 * c];    // slot nCones: LP block (0 without one)
    return t;
}

double Solver::cal_grad(double rho) {
    // ALMSetGrad / ALMCalGrad, lorads_alm.c:9-54
    push_scalars(0.0, rho);
    launch_alm_m_update(ctx, m, S.p + SL_TAU, nullptr, nullptr, s.p, lam.p, b.p, S.p + SL_RHO, M1.p);
    grad_from_M1(*this);
    if (world > 1) allreduce(S.p + kNumSlots, 2 * (nCones + 1));
    vf_valid = false;          // the gradient changed without a new (s, y) pair
    read_slots();
    return sum_grad_sq(*this);
}

void Solver::lbfgs_direction(long long counter) {
    // LBFGSDirection, lorads_alm.c:230-391 (history ring of lbfgs_len nodes, lb_head = oldest) followed by
    // LBFGSDirectionUseGrad, lorads_alm.c:469-489.  The search direction lives in U, as in the reference.
    // every vector pass runs on the rows this rank owns ([vo, vo + vn) of the concatenated vectors; everything without
    // row sharding); the finished direction is then gathered, because A(UV^T) and the R update need all of its rows
    const int L = lbfgs_len;
    double *D = U.p + vo, *q = Dtemp.p + vo;
    const double *Gv = G.p + vo;
    auto sv = [&](int k) { return lb_s[k].p + vo; };
    auto yv = [&](int k) { return lb_y[k].p + vo; };
    if (use_vf()) {
        // vector-free form: two passes over the factor vectors, every inner product from the Gram table
        const int depth = (int)std::min<long long>(counter, 2);
        const int a = (lb_head + 1) % 2, bnode = lb_head;       // a = newest pair, bnode = the older one
        if (depth >= 1 && !vf_valid) {
            launch_lbfgs_pair(ctx, vn, false, yv(a), Gv, sv(a), yv(bnode), sv(bnode), S.p, SL_VF_D,
                              SL_BETA0 + a, SL_VF_YY + a, world == 1);
            if (world > 1) { allreduce(S.p + SL_VF_D, 8); launch_lbfgs_pair_finalize(ctx, S.p, SL_VF_D, SL_BETA0 + a, SL_VF_YY + a); }
            vf_valid = true;
        }
        launch_lbfgs_dir(ctx, vn, depth, D, Gv, yv(a), sv(a), yv(bnode), sv(bnode), S.p, SL_VF_D,
                         SL_BETA0 + a, SL_BETA0 + bnode, SL_VF_YY + bnode, S.p + kNumSlots, (int)nCones + 1);
        allgather_owned(U.p);
        return;
    }
    if (counter == 0) {
        // D = -G; the <D,G> >= 0 test of LBFGSDirectionUseGrad can only fire for G = 0, where it is a no-op
        launch_axpby_dot(ctx, vn, D, coef_const(-1.0), Gv, coef_const(0.0), nullptr, nullptr, S.p, SL_DG, false);
        allgather_owned(U.p);
        return;
    }
    const int K = (int)((counter <= L - 1) ? counter : L);
    auto node = [&](int i) { return ((lb_head - i) % L + L) % L; };   // i = 1 newest ... K oldest used
    // <s_newest, G>: the pass that completed the newest pair (enqueue_back) already produced it next to <y,s>
    int t0_slot = SL_T0;
    if (vf_valid) t0_slot = SL_VF_D + 4;
    else {
        launch_dot(ctx, vn, sv(node(1)), Gv, S.p, SL_T0);
        if (world > 1) allreduce(S.p + SL_T0, 1);
    }
    for (int i = 1; i <= K; ++i) {
        const int nd = node(i);
        const double *zvec = (i < K) ? sv(node(i + 1)) : yv(nd);
        // q = q - alpha*y, alpha = beta*<s,q>; -alpha is remembered for the second loop
        launch_axpby_dot(ctx, vn, q, coef_const(1.0), (i == 1) ? Gv : q,
                         coef_prod(SL_BETA0 + nd, (i == 1) ? t0_slot : SL_T0, -1.0, SL_NEGALPHA0 + nd), yv(nd), zvec, S.p, SL_T0, false);
        if (world > 1) allreduce(S.p + SL_T0, 1);
    }
    for (int i = K; i >= 1; --i) {
        const int nd = node(i);
        if (i > 1) {
            // q = q + (alpha - beta*<y,q>) s
            Coef w{-1.0, SL_NEGALPHA0 + nd, -1.0, SL_BETA0 + nd, SL_T0, -1};
            launch_axpby_dot(ctx, vn, q, coef_const(1.0), q, w, sv(nd), yv(node(i - 1)), S.p, SL_T0, false);
            if (world > 1) allreduce(S.p + SL_T0, 1);
        } else {
            // D = -(q + w s)
            Coef negw{1.0, SL_NEGALPHA0 + nd, 1.0, SL_BETA0 + nd, SL_T0, -1};
            launch_axpby_dot(ctx, vn, D, coef_const(-1.0), q, negw, sv(nd), Gv, S.p, SL_DG, false);
            if (world > 1) allreduce(S.p + SL_DG, 1);
        }
    }
    launch_neg_if_nonneg(ctx, vn, D, Gv, S.p, SL_DG);
    allgather_owned(U.p);
}

bool Solver::p12_from_rows() const {
    // single cone covering every constraint, sparse scratch, no rank-one objective part: row m of q1 / q2 is the
    // complete objective term, so the all-reduce of [q1 | q2] already delivers p1 and p2
    return single_identity && !cones[0].dense_path && cones[0].c_rank1 == 0.0 && nLp == 0;
}

bool Solver::tri_ok() const { return single_identity && !cones[0].dense_path && nLp == 0; }

void Solver::q12p12() {
    // ALMCalq12p12, lorads_alm.c:540-560: q1 = 2 A(sym(R D^T)), p1 = 2 <C, sym(R D^T)>, q2 = A(D D^T), p2 = <C, D D^T>
    LB2_CUDA(cudaMemsetAsync(S.p + SL_P1, 0, 2 * sizeof(double), ctx.stream));
    if (single_identity) {
        // row m of the outputs is the objective row; its value is also accumulated into the P1 / P2 slots.
        // Sharded: q1 and q2 are contiguous, one all-reduce completes both (P1 / P2 are reduced by the caller).
        const bool tri = tri_ok();
        if (tri) {
            // third output of the same gather pass: A(RR^T) of the CURRENT point, i.e. the exact constrValSum and the
            // primal infeasibility that updateDimacsALM recomputes every iteration (lorads_alg_common.c:250-258)
            ConeDev &K = cones[0];
            if (shard_rows()) LB2_CUDA(cudaMemsetAsync(q12.p, 0, sizeof(double) * (size_t)((q3.p - q1.p) + m + 1), ctx.stream));
            if (K.vc_on) {
                if (K.vc_nnz_res > 0 && (!shard_rows() || K.lead))
                    launch_auv(ctx, AUV_TRI, K.vc_listRes.dev, R.p + K.off, U.p + K.off, K.ld, 2.0, 1.0, q1.p, q2.p, K.carry1.p,
                               K.carry2.p, nullptr, nullptr, q3.p, K.carry3.p);
                if (K.vc.n > 0)
                    launch_vc_auv(ctx, AUV_TRI, K.vc, K.ld, true, R.p + K.off, U.p + K.off, 2.0, 1.0, q1.p, q2.p, q3.p, S.p + SL_P1,
                                  S.p + SL_P2);
            } else if (!shard_rows() || K.lead)
                launch_auv(ctx, AUV_TRI, K.listAC.dev, R.p + K.off, U.p + K.off, K.ld, 2.0, 1.0, q1.p, q2.p, K.carry1.p, K.carry2.p,
                           S.p + SL_P1, S.p + SL_P2, q3.p, K.carry3.p);
            if (K.c_rank1 != 0.0 && (!shard_rows() || K.lead)) {
                launch_colsum(ctx, K.n, K.ld, R.p + K.off, K.csA.p, K.cs_scratch.p);
                launch_colsum(ctx, K.n, K.ld, U.p + K.off, K.csB.p, K.cs_scratch.p);
                launch_rank1_obj(ctx, K.ld, K.csA.p, K.csB.p, 2.0 * K.c_rank1, S.p + SL_P1, K.c_rank1, S.p + SL_P2);
            }
        } else {
            cone_auv_dual(cones[0], R.p, U.p, q1.p, q2.p, S.p + SL_P1, S.p + SL_P2);
        }
        // column sharding on a dense-scratch cone: cone_auv_dual all-reduced Z1 / Z2 and every rank already holds the
        // complete q1, q2 -- summing them again would scale them by `world`
        const bool complete = shard_cols() && cones[0].dense_path;
        if (world > 1 && !complete) {
            allreduce(q12.p, (long long)((tri ? q3.p : q2.p) - q1.p) + m + 1);
            if (p12_from_rows()) {
                // the reduced objective rows ARE p1 and p2: no separate scalar all-reduce
                LB2_CUDA(cudaMemcpyAsync(S.p + SL_P1, q1.p + m, sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
                LB2_CUDA(cudaMemcpyAsync(S.p + SL_P2, q2.p + m, sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
            }
        }
        tri_pending = tri;        // enqueue_front folds s := q3 and the residual into its line-search pass
        lp_q12p12();
        return;
    }
    LB2_CUDA(cudaMemsetAsync(q1.p, 0, sizeof(double) * m, ctx.stream));
    LB2_CUDA(cudaMemsetAsync(q2.p, 0, sizeof(double) * m, ctx.stream));
    for (ConeDev &K : cones) {
        cone_auv_dual(K, R.p, U.p, K.t1.p, K.t2.p, S.p + SL_P1, S.p + SL_P2);
        const int *map = K.identity_act ? nullptr : K.act_idx.p;
        launch_scatter_add(ctx, q1.p, K.t1.p, map, K.n_act, 1.0, true, nullptr);
        launch_scatter_add(ctx, q2.p, K.t2.p, map, K.n_act, 1.0, true, nullptr);
    }
    lp_q12p12();
}

void Solver::lp_q12p12() {
    // ALMCalq12p12LP, lorads_alm.c:563-581: the LP block adds 2 A(r.d), A(d.d) and 2 c.(r.d), c.(d.d)
    if (nLp == 0) return;
    launch_lp_rows(ctx, lp, R.p + N, U.p + N, 2.0, q1.p, 1.0, q2.p);
    launch_lp_obj(ctx, lp, R.p + N, U.p + N, 2.0, S.p + SL_P1, 1.0, S.p + SL_P2);
}

void Solver::primal_infeasibility(const double *Rm) {
    // primalInfeasibility, lorads_alg_common.c:250-258 (the caller reads S_host[SL_PINF] after read_slots())
    init_constr_val_all(Rm, Rm, true);
    constr_val_sum();
    launch_resid_sq(ctx, m, b.p, s.p, S.p, SL_PINF);
}

double Solver::cal_obj(const double *Rm) {
    // LORADSCalObjRR_ALM, lorads_alm.c:1259-1268
    LB2_CUDA(cudaMemsetAsync(S.p + SL_OBJ, 0, sizeof(double), ctx.stream));
    for (ConeDev &K : cones) cone_auv(K, true, Rm, Rm, true, 1.0, K.t1.p, S.p + SL_OBJ);
    if (world > 1) allreduce(S.p + SL_OBJ, 1);
    if (nLp > 0) launch_lp_obj(ctx, lp, Rm + N, Rm + N, 1.0, S.p + SL_OBJ, 0.0, nullptr);   // LORADSCalObjRR_ALM_LP
    read_slots();
    return S_host[SL_OBJ] / scaleObjHis;
}

double Solver::cal_dual_obj() {
    // LORADSCalDualObj, lorads_alg_common.c:334-340
    launch_dot(ctx, m, b.p, lam.p, S.p, SL_DOBJ);
    read_slots();
    return S_host[SL_DOBJ] / scaleObjHis;
}

void Solver::average_uv() {
    // averageUV, lorads_admm.c:310-315
    launch_axpby_dot(ctx, Nt, R.p, coef_const(0.5), U.p, coef_const(0.5), V.p, nullptr, S.p, SL_T1, false);
}

void Solver::update_dual_var(double rho) { launch_dual_update(ctx, m, rho, b.p, s.p, lam.p); }

void Solver::update_dimacs_alm() {
    // LORADSUpdateDimacsErrorALM, lorads_alg_common.c:270-274
    primal_infeasibility(R.p);
    read_slots();
    dimac_pinf = std::sqrt(S_host[SL_PINF]) / (1 + bNrm1);
    dimac_gap = std::fabs(pObj - dObj) / (1 + std::fabs(pObj) + std::fabs(dObj));
}

void Solver::update_dimacs_admm() {
    // LORADSUpdateDimacsErrorADMM, lorads_alg_common.c:282-290: R = (U+V)/2, then the ALM formula; note that
    // this overwrites constrVal / constrValSum with A(R R^T), which the dual update then uses
    average_uv();
    update_dimacs_alm();
}

// ---------------------------------------------------------------------------------------------------
// ALM inner iteration, lorads_alm.c:1073-1146
// ---------------------------------------------------------------------------------------------------
static double nth_root3(double base) {
    // LORADSnthroot(base, 3), lorads_alm.c:102-112
    return base > 0 ? std::pow(base, 1.0 / 3) : -std::pow(-base, 1.0 / 3);
}

static long long cubic_roots(double a, double b, double c, double d, double *res) {
    // Shengjin's formulas as used by LORADScubic_equation, lorads_alm.c:114-154
    const double A = b * b - 3 * a * c, B = b * c - 9 * a * d, C = c * c - 3 * b * d;
    const double delta = B * B - 4 * A * C;
    res[0] = res[1] = res[2] = 0.0;
    if (A == 0 && B == 0) {
        res[0] = std::max(res[0], -c / b);
        return 1;
    }
    if (delta > 0) {
        const double sq = std::sqrt(delta);
        const double Y1 = A * b + 1.5 * a * (-B + sq), Y2 = A * b + 1.5 * a * (-B - sq);
        res[0] = std::max(res[0], (-b - nth_root3(Y1) - nth_root3(Y2)) / 3 / a);
        return 1;
    }
    if (delta == 0 && A != 0 && B != 0) {
        const double K = B / A;
        res[0] = -b / a + K;
        res[1] = -K / 2;
        return 2;
    }
    if (delta < 0) {
        const double sqA = std::sqrt(A);
        const double T = (A * b - 1.5 * a * B) / (A * sqA);
        const double theta = std::acos(T);
        const double cs = std::cos(theta / 3), sn = std::sqrt(3.0) * std::sin(theta / 3);
        res[0] = (-b - 2 * sqA * cs) / 3 / a;
        res[1] = (-b + sqA * (cs + sn)) / 3 / a;
        res[2] = (-b + sqA * (cs - sn)) / 3 / a;
        return 3;
    }
    return 0;
}

long long line_search(double rho, const double *sums, double p1, double p2, double *tau) {
    // ALMLineSearch, lorads_alm.c:161-228: exact minimiser over [0,1] of a t^4 + b t^3 + c t^2 + d t.
    // sums = {|q2|^2, q1.q2, q0.q2, |q1|^2, q0.q1} with q0 = b - A(RR^T) + lambda/rho.
    const double a = rho * sums[0] / 2;
    const double b = rho * sums[1];
    const double c = p2 - rho * sums[2] + rho * sums[3] / 2;
    const double d = p1 - rho * sums[4];
    double roots[3];
    const long long rootNum = cubic_roots(4 * a, 3 * b, 2 * c, d, roots);
    auto f = [&](double x) { return a * std::pow(x, 4) + b * std::pow(x, 3) + c * std::pow(x, 2) + d * x; };
    const double f0 = 0.0, f1 = f(1.0);
    double fr[3] = {1e+30, 1e+30, 1e+30};
    for (int k = 0; k < 3; ++k)
        if (rootNum >= k + 1 && (k < 2 || rootNum == 3) && roots[k] > 1e-20 && roots[k] <= 1.0) fr[k] = f(roots[k]);
    const double mn = std::min(std::min(std::min(std::min(f0, f1), fr[0]), fr[1]), fr[2]);
    if (std::fabs(mn - f0) < 1e-10) *tau = 0.0;
    if (std::fabs(mn - f1) < 1e-10) *tau = 1.0;
    for (int k = 0; k < 3; ++k)
        if (std::fabs(mn - fr[k]) < 1e-10) *tau = roots[k];
    return rootNum;
}

// The inner iteration is issued in two halves so that the host needs ONE synchronisation per iteration:
//   front(k): L-BFGS direction, q1/q2/p1/p2, line-search sums          (device, no host input but the counter)
//   back(k) : needs tau_k (host line search on the 7 scalars of front(k))
// The loop enqueues back(k) and, speculatively, front(k+1) right behind it, then reads all scalars at once.
// front() only writes the direction U, Dtemp, q1, q2 and scalar slots, so a speculative front that turns out
// not to be needed (loop exit) is simply discarded.
void Solver::enqueue_front(double rho, long long counter) {
    (void)rho;     // read by the kernels from S[SL_RHO] (push_scalars)
    lbfgs_direction(counter);
    tri_pending = false;
    q12p12();
    if (world > 1 && !p12_from_rows()) allreduce(S.p + SL_P1, 2);
    if (tri_pending) launch_linesearch_resid(ctx, m, b.p, q3.p, s.p, lam.p, S.p + SL_RHO, q1.p, q2.p, S.p, SL_LS, SL_PINF);
    else launch_linesearch_dots(ctx, m, b.p, s.p, lam.p, S.p + SL_RHO, q1.p, q2.p, S.p, SL_LS);
    tri_pending = false;
}

long long Solver::finish_front(double rho, double *tau, double *p12) {
    p12[0] = S_host[SL_P1]; p12[1] = S_host[SL_P2];
    return line_search(rho, S_host + SL_LS, p12[0], p12[1], tau);
}

void Solver::enqueue_back(double rho, double tau, bool front_follows) {
    (void)rho; (void)tau;     // both are read from the scalar slots S[SL_RHO], S[SL_TAU] (push_scalars)
    const int head = lb_head;
    // variable step + m-vector update (s += tau q1 + tau^2 q2, M1 = -lambda - rho b + rho s) in one launch
    launch_alm_step_m(ctx, vn, S.p + SL_TAU, G.p + vo, U.p + vo, R.p + vo, lb_y[head].p + vo, lb_s[head].p + vo, m, q1.p, q2.p,
                      s.p, lam.p, b.p, S.p + SL_RHO, M1.p);
    if (shard_rows()) {
        // R is replicated: the rows owned elsewhere take the same fused multiply-add from the gathered direction
        launch_axpy_slot(ctx, vo, S.p + SL_TAU, U.p, R.p);
        launch_axpy_slot(ctx, N - (vo + vn), S.p + SL_TAU, U.p + vo + vn, R.p + vo + vn);
    }
    grad_from_M1(*this);
    // setlbfgsHisTwo, lorads_alm.c:657-678: y += G_new, beta = 1/<y,s>, advance the ring
    if (use_vf()) {
        const int other = (head + 1) % 2;
        launch_lbfgs_pair(ctx, vn, true, lb_y[head].p + vo, G.p + vo, lb_s[head].p + vo, lb_y[other].p + vo, lb_s[other].p + vo,
                          S.p, SL_VF_D, SL_BETA0 + head, SL_VF_YY + head, world == 1);
        if (world > 1) {
            // one all-reduce for the 8 dots (slots 64..71) and the per-cone gradient sums (slots 80..): 72..79 are unused
            allreduce(S.p + SL_VF_D, (kNumSlots - SL_VF_D) + 2 * (nCones + 1));
            launch_lbfgs_pair_finalize(ctx, S.p, SL_VF_D, SL_BETA0 + head, SL_VF_YY + head);
        }
        vf_valid = true;
    } else {
        if ((vn & 1) == 0) {
            // y += G_new, beta = 1/<y,s> and, from the same pass, <s,G_new> for the first step of the next two-loop
            // recursion (the pair kernel works on double2: an odd length -- LP block -- takes the plain path below)
            launch_lbfgs_pair(ctx, vn, true, lb_y[head].p + vo, G.p + vo, lb_s[head].p + vo, nullptr, nullptr, S.p, SL_VF_D,
                              SL_BETA0 + head, SL_VF_YY + head, world == 1);
            if (world > 1) {
                allreduce(S.p + SL_VF_D, (kNumSlots - SL_VF_D) + 2 * (nCones + 1));
                launch_lbfgs_pair_finalize(ctx, S.p, SL_VF_D, SL_BETA0 + head, SL_VF_YY + head);
            }
            vf_valid = true;
        } else {
            if (world > 1) allreduce(S.p + kNumSlots, 2 * (nCones + 1));
            launch_axpby_dot(ctx, vn, lb_y[head].p + vo, coef_const(1.0), lb_y[head].p + vo, coef_const(1.0), G.p + vo,
                             lb_s[head].p + vo, S.p, SL_BETA0 + head, world == 1);
            if (world > 1) {
                allreduce(S.p + SL_BETA0 + head, 1);
                launch_recip(ctx, S.p, SL_BETA0 + head);
            }
            vf_valid = false;
        }
    }
    lb_head = (head + 1) % lbfgs_len;
    // the front half that follows in the same submission recomputes A(RR^T) inside its dual gather pass
    if (!(front_follows && tri_ok())) primal_infeasibility(R.p);
}

void Solver::finish_back(double *lagNormSq, double *pinf1) {
    *lagNormSq = sum_grad_sq(*this);
    dimac_pinf = std::sqrt(S_host[SL_PINF]) / (1 + bNrm1);
    dimac_gap = std::fabs(pObj - dObj) / (1 + std::fabs(pObj) + std::fabs(dObj));
    *pinf1 = dimac_pinf;
}

void Solver::iter_back_front(double rho, double tau, long long next_counter) {
    const size_t slot_bytes = sizeof(double) * (kNumSlots + 2 * (nCones + 1));
    // NCCL all-reduces are captured into the graph too (validated at 2 ranks); LORADS_B200_NO_GRAPH_NCCL=1 opts out
    static const bool graph_nccl = getenv("LORADS_B200_NO_GRAPH_NCCL") == nullptr;
    static const bool mirror_scalars = getenv("LORADS_B200_NO_MIRROR") == nullptr;
    if (!use_graphs || (world > 1 && !graph_nccl)) {
        push_scalars(tau, rho);
        enqueue_back(rho, tau, next_counter >= 0);
        if (next_counter >= 0) enqueue_front(rho, next_counter);
        read_slots();
        return;
    }
    const int depth = next_counter < 0 ? 63 : (int)std::min<long long>(next_counter, lbfgs_len);
    const int key = lb_head * 64 + depth;
    auto it = iter_graphs.find(key);
    if (it == iter_graphs.end()) {
        // first visit of this (ring head, history depth): record the launches once ...
        IterGraph g;
        const long long l0 = ctx.launches;
        const int head0 = lb_head;
        cudaGraph_t graph = nullptr;
        LB2_CUDA(cudaStreamBeginCapture(ctx.stream, cudaStreamCaptureModeThreadLocal));
        try {
            // first node: tau, rho from the pinned pair the host fills before every replay; last kernel: the line-search
            // sums, whose final block mirrors the scalar slots into pinned host memory (no copy node behind it)
            LB2_CUDA(cudaMemcpyAsync(S.p + SL_TAU, push_host, 2 * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
            enqueue_back(rho, tau, next_counter >= 0);
            if (next_counter >= 0 && mirror_scalars) {
                ctx.mirror = S_host; ctx.mirror_n = (int)(slot_bytes / sizeof(double));
            }
            if (next_counter >= 0) enqueue_front(rho, next_counter);
            const bool mirrored = ctx.mirror != nullptr;
            ctx.mirror = nullptr; ctx.mirror_n = 0;
            if (!mirrored) LB2_CUDA(cudaMemcpyAsync(S_host, S.p, slot_bytes, cudaMemcpyDeviceToHost, ctx.stream));
        } catch (...) {
            ctx.mirror = nullptr; ctx.mirror_n = 0;
            cudaStreamEndCapture(ctx.stream, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        LB2_CUDA(cudaStreamEndCapture(ctx.stream, &graph));
        LB2_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
        LB2_CUDA(cudaGraphDestroy(graph));
        g.launches = ctx.launches - l0;
        ctx.launches = l0;
        lb_head = head0;                 // nothing has executed yet: the replay below advances the ring
        it = iter_graphs.emplace(key, g).first;
    }
    // ... and replay them as one graph launch
    push_host[0] = tau; push_host[1] = rho;
    LB2_CUDA(cudaGraphLaunch(it->second.exec, ctx.stream));
    ctx.launches += it->second.launches;
    lb_head = (lb_head + 1) % lbfgs_len;
    spin_sync(ctx.stream);
}

int Solver::alm_inner_front(double rho, long long counter, double *tau, double *p12, long long *rootNum) {
    push_scalars(*tau, rho);
    enqueue_front(rho, counter);
    read_slots();
    *rootNum = finish_front(rho, tau, p12);
    return 0;
}

void Solver::alm_inner_back(double rho, double tau, double *lagNormSq, double *pinf1) {
    push_scalars(tau, rho);
    enqueue_back(rho, tau);
    read_slots();
    finish_back(lagNormSq, pinf1);
}

// `iters` inner iterations at fixed rho with the pipelined schedule (benchmark / host-buffer entry points);
// counter k = 0,1,2,... as in a fresh sub-problem.  Returns the number of iterations done.
long long Solver::run_inner_iters(double rho, long long iters, double *out) {
    double tau = 0.0, p12[2] = {0, 0}, lag = 0, pinf = 0;
    long long done = 0;
    if (iters > 0) { push_scalars(0.0, rho); enqueue_front(rho, 0); read_slots(); }
    for (long long k = 0; k < iters; ++k) {
        const long long rn = finish_front(rho, &tau, p12);
        if (rn == 0) break;
        iter_back_front(rho, tau, (k + 1 < iters) ? k + 1 : -1);
        finish_back(&lag, &pinf);
        done++;
    }
    out[0] = tau; out[1] = lag; out[2] = pinf; out[3] = p12[0]; out[4] = p12[1]; out[5] = (double)done;
    return done;
}

// ---------------------------------------------------------------------------------------------------
// ADMM block update: RHS + CG, lorads_admm.c:376-480 and lorads_cgs.c:81-240
// ---------------------------------------------------------------------------------------------------
void Solver::cg_matvec(ConeDev &K, const double *x, const double *Vn, double *res, const double *Z2, double *red) {
    // linSysProduct: weight = A(sym(x V^T)); W = sum_i weight_i A_i; res = W V + x
    cone_auv(K, false, x, Vn, false, 1.0, K.t1.p);
    cone_wsum(K, K.t1.p, true, false);
    cone_mul(K, Vn, 1.0, 1.0, x, Z2, res, red);
}

void Solver::update_sdp_var_one(long long c, double *upd, const double *noupd, double rho, double tol, long long maxit) {
    ConeDev &K = cones[c];
    // row sharding: every vector pass of the CG runs on this rank's rows of the cone; the vectors the mat-vec gathers
    // from (the iterate at a restart, the search direction every iteration) are completed by an all-gather
    const long long off = K.off + (shard_rows() ? K.row_lo * (long long)K.ld : 0);
    const long long nk = (shard_rows() ? (K.row_hi - K.row_lo) : K.n) * (long long)K.ld;
    // M1 = rho (constrValSum - constrVal[c] - b) - lambda
    const double *cvf = nullptr;
    if (K.identity_act) cvf = K.cv.p; else { expand_cv(K, cvfull.p); cvf = cvfull.p; }
    launch_admm_m1(ctx, m, b.p, s.p, cvf, lam.p, rho, M1.p);
    // bLinSys = -( (C + A^*(M1)) V - rho V ) / rho
    cone_wsum(K, M1.p, false, true);
    cone_mul(K, noupd, -1.0 / rho, 1.0, noupd, nullptr, Bls.p, nullptr);

    double *x = upd + off, *bl = Bls.p + off, *r = cg_r.p + off, *p = cg_p.p + off;
    // shifted views so that the cone_* helpers (which add K.off) see this cone's slice
    double *r0 = cg_r.p, *p0 = cg_p.p, *Q0 = cg_Q.p;
    const double t0 = wall_time();
    launch_asum(ctx, nk, bl, S.p, SL_CG_BN);
    cg_matvec(K, upd, noupd, r0, nullptr, nullptr);
    int rrA = SL_CG_RR_A, rrB = SL_CG_RR_B;
    // r = b - A x ; <r,r>
    launch_axpby_dot(ctx, nk, r, coef_const(-1.0), r, coef_const(1.0), bl, r, S.p, rrA, false);
    if (world > 1) { allreduce(S.p + SL_CG_BN, 1); allreduce(S.p + rrA, 1); }
    read_slots();
    const double bNorm = S_host[SL_CG_BN];
    double rr = S_host[rrA];
    if (std::sqrt(rr) / bNorm < tol) {   // converged at the warm start; cg->iter keeps its previous value
        cgTime += wall_time() - t0;
        cgIter += K.cg_iter_last;
        return;
    }
    if (nk > 0) LB2_CUDA(cudaMemcpyAsync(p, r, sizeof(double) * nk, cudaMemcpyDeviceToDevice, ctx.stream));
    long long iter = 0;
    for (long long k = 0; k < maxit; ++k) {
        iter += 1;
        allgather_cone(K, p0);
        cg_matvec(K, p0, noupd, Q0, p0, S.p + SL_CG_RED);       // Q = A p, S[SL_CG_RED+1] = <p,Q>
        if (world > 1) allreduce(S.p + SL_CG_RED, 2);
        // single GPU, no restart due: beta is formed on the device by the update kernel and p = r + beta p is enqueued
        // BEFORE the scalar read-back, so the round trip overlaps it (same IEEE division and FMA as the host path; when
        // the solve stops here the extra p is never used)
        const bool early_p = (world == 1) && (k % 20 != 0);
        launch_cg_update(ctx, nk, x, r, p, cg_Q.p + off, S.p, rrA, SL_CG_RED + 1, rrB, early_p ? SL_CG_BETA : -1);
        if (world > 1) allreduce(S.p + rrB, 1);
        if (early_p) {
            read_slots_side();       // scalars as of the update kernel, copied on the side stream ...
            launch_axpby_dot(ctx, nk, p, coef_slot(SL_CG_BETA), p, coef_const(1.0), r, nullptr, S.p, SL_T1, false);
            spin_sync(side_stream);      // ... while the direction update runs
        } else read_slots();
        double rrNew = S_host[rrB];
        const double resi = std::sqrt(rrNew);
        if (resi / bNorm < tol) break;
        if (early_p) {
            rr = rrNew;
            std::swap(rrA, rrB);
            if (resi != resi) printf("File [%30s] Line [%d]\n", "lorads_b200 cg", (int)k);
            continue;
        }
        double beta;
        if (k % 20 == 0) {
            // restart (also at k = 0, lorads_cgs.c:195): r = b - A x, p = q = r, then beta = <r,r>/<r,r>
            allgather_cone(K, upd);
            cg_matvec(K, upd, noupd, r0, nullptr, nullptr);
            launch_axpby_dot(ctx, nk, r, coef_const(-1.0), r, coef_const(1.0), bl, r, S.p, rrB, false);
            if (world > 1) allreduce(S.p + rrB, 1);
            if (nk > 0) LB2_CUDA(cudaMemcpyAsync(p, r, sizeof(double) * nk, cudaMemcpyDeviceToDevice, ctx.stream));
            beta = 1.0;
        } else {
            beta = rrNew / rr;
        }
        // p = r + beta p
        launch_axpby_dot(ctx, nk, p, coef_const(beta), p, coef_const(1.0), r, nullptr, S.p, SL_T1, false);
        if (k % 20 == 0) { read_slots(); rrNew = S_host[rrB]; }
        rr = rrNew;
        std::swap(rrA, rrB);
        if (resi != resi) printf("File [%30s] Line [%d]\n", "lorads_b200 cg", (int)k);   // NaN trace, as the reference
    }
    allgather_cone(K, upd);        // the other half of the sweep and A(UV^T) gather from every row of the new iterate
    K.cg_iter_last = iter;
    cgTime += wall_time() - t0;
    cgIter += iter;
}

void Solver::update_sdp_var(double rho, double tol, long long maxit) {
    // LORADSUpdateSDPVar, lorads_alg_common.c:187-215 (Gauss-Seidel over cones)
    for (long long c = 0; c < nCones; ++c) {
        ConeDev &K = cones[c];
        const int *map = K.identity_act ? nullptr : K.act_idx.p;
        for (int half = 0; half < 2; ++half) {
            if (half == 0) update_sdp_var_one(c, U.p, V.p, rho, tol, maxit);
            else update_sdp_var_one(c, V.p, U.p, rho, tol, maxit);
            launch_scatter_add(ctx, s.p, K.cv.p, map, K.n_act, -1.0, true, nullptr);
            update_constr_val(c, U.p, V.p);
            launch_scatter_add(ctx, s.p, K.cv.p, map, K.n_act, 1.0, true, nullptr);
        }
    }
    // LORADSUpdateSDPLPVar, lorads_alg_common.c:225-249: the LP columns follow the cones in the same sweep
    if (nLp > 0) launch_lp_sweep(ctx, lp, lp_segs.data(), (int)lp_segs.size(), rho, b.p, lam.p, s.p, lp_x.p, U.p + N, V.p + N);
}

void Solver::get_lp_vec(char which, double *out) {
    if (nLp == 0) return;
    const double *src = (which == 'x') ? lp_x.p : (which == 'c') ? lp_c.p : factor_ptr(which) + N;
    LB2_CUDA(cudaMemcpyAsync(out, src, sizeof(double) * nLp, cudaMemcpyDeviceToHost, ctx.stream));
    LB2_CUDA(cudaStreamSynchronize(ctx.stream));
}

void Solver::set_lp_vec(char which, const double *in) {
    if (nLp == 0) return;
    double *dst = (which == 'x') ? lp_x.p : factor_ptr(which) + N;
    LB2_CUDA(cudaMemcpyAsync(dst, in, sizeof(double) * nLp, cudaMemcpyHostToDevice, ctx.stream));
    LB2_CUDA(cudaStreamSynchronize(ctx.stream));
}

}  // namespace lb2
