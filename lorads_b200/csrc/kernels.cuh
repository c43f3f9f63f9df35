// kernels.cuh -- launch interface of the hand-written sm_100a kernels (definitions in kernels.cu).
#pragma once
#include "common.cuh"

namespace lb2 {

// Launch context: one stream, one reduction scratch, a launch counter (reported as gpu_launches).
struct Ctx {
    cudaStream_t stream = nullptr;
    ReduceScratch rs{nullptr, nullptr};
    unsigned int *ticket = nullptr;    // zero-initialised counter for "last block finishes" kernels
    double *dense_part = nullptr;      // split-K partials of the dense symmetric product (dense-scratch cones only)
    size_t dense_part_cap = 0;         // capacity of dense_part in doubles
    int num_sms = 148;
    long long launches = 0;
    // When set, the kernel that ends an iteration (the line-search sums) copies the first mirror_n scalar slots into
    // this pinned host buffer from its last block, replacing a device-to-host copy node behind it.
    double *mirror = nullptr;
    int mirror_n = 0;
};

// ------------------------------------------------------------------------------------------------
// A(UV^T): items = non-zeros of [A_1..A_m ; C] sorted by constraint (layout.hpp ItemList)
// ------------------------------------------------------------------------------------------------
struct ItemListDev {
    long long n_items = 0, n_rows = 0, n_tiles = 0, n_split = 0;
    const int *ptr = nullptr, *irow = nullptr, *icol = nullptr;
    const double *coef = nullptr;
    const int *split_row = nullptr, *split_first_slot = nullptr, *split_tile_a = nullptr, *split_tile_b = nullptr;
    const int *tile_row_lo = nullptr, *tile_row_hi = nullptr;
    int tile = 512;          // items per CTA tile (512 or 128)
    int obj_row = -1;        // index of the objective row (last row of an [A;C] list), -1 if none
    bool has_empty_rows = false;
};

enum AuvMode {
    AUV_SAME = 0,   // U == V            : z = U_i . U_j
    AUV_PAIR = 1,   // U != V            : z = (U_i.V_j + U_j.V_i)/2
    AUV_DUAL = 2,   // (R,D) and (D,D)   : z1 = (R_i.D_j + R_j.D_i)/2 , z2 = D_i.D_j  (one gather pass)
    AUV_FROMZ = 3,  // dense path        : z is materialised (packed), item.irow is the packed position
    AUV_TRI = 4     // AUV_DUAL + out3 = A(R R^T) from the rows the dual pass already holds (z3 = R_i.R_j, scale 1)
};

// out1/out2: n_rows values each; carry: 2*n_tiles doubles per output.  U,V row-major n x ld (ld % 4 == 0).
// obj1/obj2 (may be null): the value of the objective row is ADDED to *obj1 / *obj2 as well.
void launch_auv(Ctx &c, AuvMode mode, const ItemListDev &L, const double *U, const double *V, int ld,
                double scale1, double scale2, double *out1, double *out2, double *carry1, double *carry2,
                double *obj1 = nullptr, double *obj2 = nullptr, double *out3 = nullptr, double *carry3 = nullptr);

// dst[idx ? idx[a] : a] (+)= alpha * src[a]  for a < n ; when obj_dst != nullptr: *obj_dst += alpha*src[n]
void launch_scatter_add(Ctx &c, double *dst, const double *src, const int *idx, long long n, double alpha,
                        bool accumulate, double *obj_dst);

// ------------------------------------------------------------------------------------------------
// (C + A^*(w)) X : weighted sum on the pattern, then symmetric SpMM in gather form
// ------------------------------------------------------------------------------------------------
// S[p] = (addC ? C_onP[p] : 0) + sum_k w[map(T_con[k])] * T_val[k],  map = act_idx unless w is compact
void launch_wsum(Ctx &c, double *S, long long np, const double *C_onP, const int *T_ptr, const int *T_con,
                 const double *T_val, const double *w, const int *act_idx, bool w_is_compact, bool addC);

// Y[i,:] = a * sum_{e in adj(i)} S[adj_pos[e]] * X[adj_col[e],:] + b * Z[i,:]   (Z may be null)
// red[0] = sum Y.Y, red[1] = sum Y.Z2 (Z2 may be null) when red != nullptr.
// cs != nullptr adds the rank-one part of the objective: S_eff = S + c1 * e e^T with cs = column sums of X.
void launch_spmm(Ctx &c, long long n, int ld, const int *adj_ptr, const int *adj_col, const int *adj_pos,
                 const double *S, const double *X, double a, double b, const double *Z, const double *Z2,
                 double *Y, double *red, const double *cs = nullptr, double c1 = 0.0);

// ------------------------------------------------------------------------------------------------
// vertex-centric fast path (vc_kernels.cu; layout.hpp VcLayout): class-split adjacency, nothing materialised
// ------------------------------------------------------------------------------------------------
struct VcDev {
    long long n = 0;
    int obj_row = 0;                       // index of the objective row in the A(UV^T) outputs (= n_act)
    const int *order = nullptr;            // rows by decreasing SpMM work
    const int *order_l = nullptr;          // rows by decreasing A(UV^T) work
    // adjU: row i = [u_ptr[i], u_mid[i]) weight-dependent entries (tag >= 0 singleton constraint, tag <= -2 residual
    // pattern position -2-tag), [u_mid[i], u_ptr[i+1]) objective entries (tag -1, value inline)
    const int *u_ptr = nullptr, *u_mid = nullptr, *u_col = nullptr, *u_tag = nullptr;
    const double *u_val = nullptr;
    const int *d_con = nullptr; const double *d_coef = nullptr;                                           // diagonal singletons
    // lowA: the other singleton entries, flat in (col,row) order; [l_lo, l_hi) is the part this rank evaluates
    const int *l_ptr = nullptr, *l_row = nullptr, *l_col = nullptr, *l_con = nullptr; const double *l_coef = nullptr;
    long long l_lo = 0, l_hi = 0;
};
int vc_max_ld();
// Y[i,:] = a * sum_{e in adjU(i)} s_e X[col_e,:] + b Z[i,:],  s_e = value (objective entries, only when useC),
// w[map(tag)] * value (singleton constraints) or Sres[-2-tag] (residual positions).  w == nullptr and Sres == nullptr
// skip the weight-dependent part; wmap (compact -> global constraint index) may be null.
// red / cs / c1 as launch_spmm.
void launch_vc_spmm(Ctx &c, const VcDev &V, int ld, bool useC, const double *w, const int *wmap, const double *Sres,
                    const double *X, double a, double b, const double *Z, const double *Z2, double *Y, double *red,
                    const double *cs = nullptr, double c1 = 0.0);
// A(sym(U W^T)) of the singleton constraints (written straight to their rows) and, with_obj, the objective row
// <C, sym(U W^T)> (also added to *obj1 / *obj2).  mode: AUV_SAME / AUV_PAIR / AUV_DUAL / AUV_TRI.
void launch_vc_auv(Ctx &c, AuvMode mode, const VcDev &V, int ld, bool with_obj, const double *U, const double *W, double s1,
                   double s2, double *out1, double *out2, double *out3, double *obj1, double *obj2);

// rank-one objective C = c1 * e e^T (+ sparse remainder): column sums of a factor, and the objective terms
// out[k] = sum_i X[i,k]  (deterministic two-level sum; scratch >= 4*148*ld doubles)
void launch_colsum(Ctx &c, long long n, int ld, const double *X, double *out, double *scratch);
// *obj1 += k1 * <csA, csB> ; if obj2: *obj2 += k2 * <csB, csB>
void launch_rank1_obj(Ctx &c, int ld, const double *csA, const double *csB, double k1, double *obj1, double k2, double *obj2);

// ------------------------------------------------------------------------------------------------
// dense path (cones whose scratch matrices are dense, lorads_sdp_conic.c:884-963)
// ------------------------------------------------------------------------------------------------
// Zp (packed lower, column-major as the reference) = sym(U V^T);  replaces dsyr2k + repack (lorads_alg_common.c:50-67)
void launch_dense_uvt(Ctx &c, long long n, int r, int ld, const double *U, const double *V, double *Zp, bool same,
                      const double *Cp = nullptr, double scale = 0.0, double *obj = nullptr);
// dual variant: Z1 = sym(R D^T), Z2 = D D^T
void launch_dense_uvt_dual(Ctx &c, long long n, int r, int ld, const double *R, const double *D, double *Z1, double *Z2,
                           const double *Cp = nullptr, double s1 = 0.0, double s2 = 0.0, double *obj1 = nullptr,
                           double *obj2 = nullptr);
// Sp = (addC ? Cp : 0) then Sp[D_pos[u]] += sum_k w[..] T_val[k]
void launch_dense_wsum(Ctx &c, double *Sp, long long psize, const double *Cp, const long long *D_pos, long long n_pos,
                       const int *T_ptr, const int *T_con, const double *T_val, const double *w, const int *act_idx,
                       bool w_is_compact, bool addC);
// Y = a * Sp(sym, packed) * X + b * Z ; reductions as launch_spmm
void launch_dense_symm(Ctx &c, long long n, int r, int ld, const double *Sp, const double *X, double a, double b,
                       const double *Z, const double *Z2, double *Y, double *red);

// y = S x for one n-vector (dual infeasibility Lanczos; reference mv: dataMatSparseMV / dataMatDenseMV,
// lorads_sdp_data.c:506-521,673-696)
// when sum_slot != nullptr: y += c1 * (*sum_slot)   (rank-one part: c1 * e (e^T x), *sum_slot = sum of x)
void launch_spmv_sym(Ctx &c, long long n, const int *adj_ptr, const int *adj_col, const int *adj_pos, const double *S,
                     const double *x, double *y, double c1 = 0.0, const double *sum_slot = nullptr);
void launch_sum(Ctx &c, long long n, const double *x, double *S, int slot);   // S[slot] = sum x
// Block Gram-Schmidt steps of the Lanczos process (basis vectors V[l] = Vb + l*stride, l < nv <= kLanczosMax):
//   dots  : h[l] = V[l] . w for every l in ONE pass over w (deterministic two-level reduction; scratch holds
//           gridDim * nv partials); h[nv-1] is also copied to S[alpha_slot] when alpha_slot >= 0
//   update: w -= sum_l h[l] V[l] in one pass; S[wsq_slot] = sum w^2 afterwards when wsq_slot >= 0
//   next  : out = w / sqrt(S[wsq_slot])
constexpr int kLanczosMax = 40;
void launch_lanczos_dots(Ctx &c, long long n, const double *Vb, long long stride, int nv, const double *w, double *h,
                         double *scratch, double *S, int alpha_slot);
void launch_lanczos_update(Ctx &c, long long n, const double *Vb, long long stride, int nv, const double *h, double *w,
                           double *S, int wsq_slot);
void launch_lanczos_next(Ctx &c, long long n, const double *w, double *out, const double *S, int wsq_slot);
void launch_dense_symv(Ctx &c, long long n, const double *Sp, const double *x, double *y);

// ------------------------------------------------------------------------------------------------
// fused BLAS-1 on the concatenated factor vectors (length N) and on m-vectors
// ------------------------------------------------------------------------------------------------
// A coefficient evaluated on the device from scalar slots: value = k0*S[i0] + k1*S[i1]*S[i2];
// if store >= 0 thread 0 writes it to S[store].  Slot 0 holds 1.0.
struct Coef {
    double k0; int i0; double k1; int i1, i2; int store;
};
inline Coef coef_const(double v) { return Coef{v, 0, 0.0, 0, 0, -1}; }
inline Coef coef_slot(int i, double k = 1.0) { return Coef{k, i, 0.0, 0, 0, -1}; }
inline Coef coef_prod(int i1, int i2, double k = 1.0, int store = -1) { return Coef{0.0, 0, k, i1, i2, store}; }
inline Coef coef_affine(int i0, double k1, int i1, int i2, int store = -1) { return Coef{1.0, i0, k1, i1, i2, store}; }

// out = a*x + b*y ; if z != nullptr: S[dot_slot] = sum out*z ; if recip: S[dot_slot] = 1/sum
// out may alias x or y.  y may be nullptr (treated as 0).
void launch_axpby_dot(Ctx &c, long long n, double *out, Coef a, const double *x, Coef b, const double *y,
                      const double *z, double *S, int dot_slot, bool recip);
// S[slot] = sum x*y
void launch_dot(Ctx &c, long long n, const double *x, const double *y, double *S, int slot);
// S[slot] = sum |x|
void launch_asum(Ctx &c, long long n, const double *x, double *S, int slot);
// if (S[cond_slot] >= 0) D = -G           (LBFGSDirectionUseGrad, lorads_alm.c:469-489)
void launch_neg_if_nonneg(Ctx &c, long long n, double *D, const double *G, const double *S, int cond_slot);
// y = -G ; s = tau*D ; R += tau*D         (SetyAsNegGrad :583, ALMupdateVar :619, first half of setlbfgsHisTwo :657)
// tau is read from *tau_p (a device scalar slot) so that the launch can live in a replayed CUDA graph
void launch_alm_step(Ctx &c, long long n, const double *tau_p, const double *G, const double *D, double *R, double *y, double *s);
// the same plus launch_alm_m_update(m, tau_p, q1, q2, sm, lam, b, rho_p, M1) in ONE launch (extra CTAs; both passes are
// element-wise and independent, same arithmetic as the two separate kernels)
void launch_alm_step_m(Ctx &c, long long n, const double *tau_p, const double *G, const double *D, double *R, double *y, double *s,
                       long long m, const double *q1, const double *q2, double *sm, const double *lam, const double *b,
                       const double *rho_p, double *M1);
// x += alpha*p ; r -= alpha*Q ; S[slot_rr] = sum r*r  with alpha = S[slot_num]/S[slot_den]  (lorads_cgs.c:183-189)
// slot_beta >= 0: S[slot_beta] = (new sum r*r) / S[slot_num], the beta of the next search direction
void launch_cg_update(Ctx &c, long long n, double *x, double *r, const double *p, const double *Q, double *S,
                      int slot_num, int slot_den, int slot_rr, int slot_beta = -1);

// ---- vector-free L-BFGS (history length 2): all inner products of the two-loop recursion come from a small Gram
// table, so the recursion needs two passes over the factor vectors instead of five, and one reduction.
// Pair completion (setlbfgsHisTwo, lorads_alm.c:657-678, fused with the dots the next direction needs):
//   ya += G ; S[d+0]=ya.sa (S[beta_slot]=1/that) S[d+1]=ya.ya S[d+2]=ya.yb S[d+3]=ya.sb
//   S[d+4]=sa.G S[d+5]=sb.G S[d+6]=ya.G S[d+7]=yb.G          (yb, sb may be null: their dots are 0)
// update_y = false leaves ya untouched and only refreshes the dots (G changed without a new pair).
// finalize: also store S[beta_slot] = 1/(ya.sa) and S[yy_slot] = ya.ya (single GPU; sharded runs all-reduce the
// eight dots first and call launch_lbfgs_pair_finalize).
void launch_lbfgs_pair(Ctx &c, long long n, bool update_y, double *ya, const double *G, const double *sa, const double *yb,
                       const double *sb, double *S, int d_slot, int beta_slot, int yy_slot, bool finalize);
void launch_lbfgs_pair_finalize(Ctx &c, double *S, int d_slot, int beta_slot, int yy_slot);
// Direction (LBFGSDirection + LBFGSDirectionUseGrad, lorads_alm.c:230-391,469-489) from the Gram table:
// depth 0: D = -G; depth 1: newest pair a only; depth 2: pairs a (newest) and b.  gg_slot holds G.G.
void launch_lbfgs_dir(Ctx &c, long long n, int depth, double *D, const double *G, const double *ya, const double *sa,
                      const double *yb, const double *sb, const double *S, int d_slot, int beta_a, int beta_b,
                      int yy_b_slot, const double *gg_slots, int n_gg);

// line-search sums over m (ALMLineSearch, lorads_alm.c:161-172): q0 = b - s + lambda/rho
// S[slot+0]=|q2|^2  S[slot+1]=q1.q2  S[slot+2]=q0.q2  S[slot+3]=|q1|^2  S[slot+4]=q0.q1
void launch_linesearch_dots(Ctx &c, long long m, const double *b, const double *s, const double *lam, const double *rho_p,
                            const double *q1, const double *q2, double *S, int slot);
// the same five sums with s := src first and S[pinf_slot] = sum (b - s)^2 (primalInfeasibility) from the same pass
void launch_linesearch_resid(Ctx &c, long long m, const double *b, const double *src, double *s, const double *lam,
                             const double *rho_p, const double *q1, const double *q2, double *S, int slot, int pinf_slot);
// s += tau*q1 + tau^2*q2 (when q1 != nullptr) ; M1 = -lambda - rho*b + rho*s   (lorads_alm.c:1123-1124, 15-27)
void launch_alm_m_update(Ctx &c, long long m, const double *tau_p, const double *q1, const double *q2, double *s,
                         const double *lam, const double *b, const double *rho_p, double *M1);
// S[slot] = sum (b - s)^2  (primalInfeasibility, lorads_alg_common.c:255-257)
void launch_resid_sq(Ctx &c, long long m, const double *b, const double *s, double *S, int slot);
// lambda = (lambda + rho*b) - rho*s  (LORADSUpdateDualVar, lorads_alg_common.c:319-332)
void launch_dual_update(Ctx &c, long long m, double rho, const double *b, const double *s, double *lam);
// M1 = ((s - b) - cv) * rho - lambda  (LORADSUpdateSDPVarOne, lorads_admm.c:432-445); cv is the cone's
// constraint vector expanded to length m (may be null)
void launch_admm_m1(Ctx &c, long long m, const double *b, const double *s, const double *cv, const double *lam,
                    double rho, double *M1);
// ------------------------------------------------------------------------------------------------
// LP cone (diagonal block of the SDPA file): x_j = u_j * v_j, columns a_j of the m x nLp constraint matrix
// (reference: data/lorads_lp_conic.c, lorads_lp_data.c and the *LP function set, lorads_solver.c:717-735)
// ------------------------------------------------------------------------------------------------
struct LpDev {
    long long n = 0, m = 0;                   // LP columns, constraints
    const double *c = nullptr;                // objective coefficients (scaled with the objective)
    const int *rbeg = nullptr, *rcol = nullptr;   // by constraint row (CSR)
    const double *rval = nullptr;
    const int *cbeg = nullptr, *crow = nullptr;   // by LP column (CSC), rows ascending inside a column
    const double *cval = nullptr;
    const double *nrm2sq = nullptr;           // |a_j|^2
    const int *lvl_ptr = nullptr, *lvl_col = nullptr;   // level schedule of the Gauss-Seidel sweep
    int n_lvl = 0;
};
// x[j] = u[j] * v[j]
void launch_lp_prod(Ctx &c, long long n, const double *u, const double *v, double *x);
// out1[i] += s1 * sum_j a_ij u_j v_j ; if out2: out2[i] += s2 * sum_j a_ij v_j v_j   (v == nullptr: factor 1)
// (lp_cone_AUV lorads_lp_conic.c:164-167 for every column + the add into constrValSum / q1 / q2,
//  lorads_alg_common.c:144-158, lorads_alm.c:525-537)
void launch_lp_rows(Ctx &c, const LpDev &L, const double *u, const double *v, double s1, double *out1, double s2, double *out2);
// *obj1 += s1 * sum c_j u_j v_j ; if obj2: *obj2 += s2 * sum c_j v_j v_j   (lp_cone_objAUV lorads_lp_conic.c:186-195)
void launch_lp_obj(Ctx &c, const LpDev &L, const double *u, const double *v, double s1, double *obj1, double s2, double *obj2);
// g_j = 2 (c_j + a_j^T w) r_j ; red[0] = sum g_j^2   (ALMSetGradLP / ALMCalGradLP, lorads_alm.c:56-99)
void launch_lp_grad(Ctx &c, const LpDev &L, const double *w, const double *r, double *g, double *red);
// One Gauss-Seidel sweep over the LP columns, u_j then v_j, with constrValSum kept current
// (LORADSUpdateSDPLPVar + LORADSUpdateLPVarOne, lorads_alg_common.c:225-249, lorads_admm.c:595-628);
// x[j] is the product currently represented in constrValSum for column j (constrValLP[j] = a_j x[j]).
struct LpSeg { int lv_lo, lv_hi, blocks; };     // levels [lv_lo, lv_hi) in one launch of `blocks` CTAs
void launch_lp_sweep(Ctx &c, const LpDev &L, const LpSeg *segs, int n_segs, double rho, const double *b, const double *lam,
                     double *cvs, double *x, double *u, double *v);
// S[slot] = sum_j |min(c_j + a_j^T w, 0)|   (calculate_dual_infeasibility_solver, lorads_solver.c:1015-1023; w = -lambda)
void launch_lp_dinf(Ctx &c, const LpDev &L, const double *w, double *S, int slot);

// Rank augmentation (AUG_RANK, lorads_solver.c:806-906): dst (n x ld_new, zero-filled) receives the first r_old
// columns of src (n x ld_old); rows i < n_diag of column (c_diag + i) get diag_val when diag_val != 0
// (lpRandomDiag, lorads_solver.c:776-786).
void launch_relayout(Ctx &c, long long n, int r_old, int ld_old, int ld_new, const double *src, double *dst);
// C-ABI boundary (xfer_kernels.cu): r_own contiguous columns of length n (the caller's column-major layout) <-> the
// row-major n x ld device layout; padding columns are zero-filled on the way in.
void launch_cm_to_rm(Ctx &c, long long n, int r_own, int ld, const double *src, double *dst);
void launch_rm_to_cm(Ctx &c, long long n, int r_own, int ld, const double *src, double *dst);
// cold-cache timing helper: reads `n` doubles (n * 8 bytes >> L2) so that the L2 holds only clean lines of `buf`
void launch_l2_flush_read(Ctx &c, const double *buf, size_t n, double *sink);
// val[e] *= f for the objective entries (tag == -1) of a vertex-centric adjacency
void launch_scale_tagged(Ctx &c, long long n, const int *tag, double *val, double f);
// R += (*tau_p) * D with the same fma as launch_alm_step (rows not owned under row sharding)
void launch_axpy_slot(Ctx &c, long long n, const double *tau_p, const double *D, double *R);
// ------------------------------------------------------------------------------------------------
// One-shot all-reduce over NVLink peer memory (column-sharded runs).  Every rank owns an exchange buffer of
// 2 sets x world slots x cap doubles and 2 x world arrival flags, mapped into every peer through CUDA IPC.
//   push  : the rank stores its `count` partial values into slot[myrank] of EVERY peer's buffer (its own included)
//           and, once all of its stores are fenced, raises flag[myrank] = epoch on every peer;
//   reduce: waits until the `world` local flags carry the epoch, then adds the slots in rank order (the same order
//           on every rank, so all ranks hold bit-identical sums) and writes the result back over the input.
// Two buffer sets alternate with the epoch parity: a peer can be at most one all-reduce ahead, so a slot is never
// overwritten while its owner still reads it.  The epoch counter lives on the device, which keeps the pair of
// launches replayable inside a CUDA graph.
// ------------------------------------------------------------------------------------------------
struct P2PDev {
    int world = 1, rank = 0;
    size_t cap = 0;                               // doubles per slot
    double *const *peer_x = nullptr;              // [world] exchange buffers (device array of peer pointers)
    unsigned long long *const *peer_f = nullptr;  // [world] flag arrays
    double *x = nullptr;                          // this rank's exchange buffer
    unsigned long long *f = nullptr;              // this rank's flags
    unsigned long long *epoch = nullptr;          // completed all-reduces (device counter)
    unsigned int *ticket = nullptr;               // block tickets of the two kernels (2 counters)
};
void launch_p2p_allreduce(Ctx &c, const P2PDev &P, double *data, long long count);
void launch_recip(Ctx &c, double *S, int slot);   // S[slot] = 1/S[slot]
void launch_scale(Ctx &c, double *x, long long n, double f);   // x *= f

}  // namespace lb2
