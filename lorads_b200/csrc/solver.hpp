// solver.hpp -- device-resident solver state and the operators of the hot path.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/lorads_b200.h"
#include "kernels.cuh"
#include "layout.hpp"

namespace lb2 {

// scalar slots on the device (double S[kNumSlots + 2*nCones])
enum Slot : int {
    SL_ONE = 0, SL_ZERO = 1,
    SL_T0 = 2, SL_T1 = 3,
    SL_TAU = 4, SL_RHO = 5,   // step length / penalty of the current inner iteration (written by the host)
    SL_LS = 8,            // 8..12 line-search sums
    SL_P1 = 13, SL_P2 = 14,
    SL_PINF = 17, SL_DG = 19, SL_OBJ = 20, SL_DOBJ = 21,
    SL_CG_RED = 24,       // 24 = sum Q.Q (unused), 25 = p.Q
    SL_CG_RR_A = 26, SL_CG_RR_B = 27, SL_CG_BN = 28,
    SL_CG_BETA = 31,      // beta = <r,r>_new / <r,r>_old of the CG, computed by the update kernel
    SL_NEGALPHA0 = 32,    // 32..47  -alpha of the L-BFGS two-loop, per history node
    SL_BETA0 = 48,        // 48..63  beta = 1/<y,s> per history node
    SL_VF_D = 64,         // 64..71 Gram dots of the vector-free L-BFGS (see launch_lbfgs_pair)
    SL_VF_YY = 29,        // 29..30 y.y of the two history nodes (kept away from 64..79 so that one all-reduce can
                          // cover the dot table and the per-cone gradient sums that follow kNumSlots)
    kNumSlots = 80        // followed by 2 slots per cone (sum G.G, sum G.Z2 of that cone) and 2 for the LP block
};
constexpr int kMaxLbfgs = 16;
constexpr long long kMaxRank = 256;      // widest factor the gather kernels take (vc_kernels.cu vc_passes, kernels.cu)

struct ItemListBufs {
    DBuf<int> ptr, irow, icol, split_row, split_first_slot, split_tile_a, split_tile_b, tile_row_lo, tile_row_hi;
    DBuf<double> coef;
    ItemListDev dev;
    void upload(const ItemList &L);
};

struct ConeDev {
    long long n = 0;
    int r = 0;           // rank (number of factor columns held on this device)
    int ld = 0;          // padded leading dimension of the row-major factor storage
    long long off = 0;   // offset of this cone inside the concatenated N-vectors
    bool dense_path = false, dense_cone = true, identity_act = false;
    long long n_act = 0, np = 0, nnzA = 0, nnzC = 0, n_nonzero_coeff = 0, n_pos = 0;
    long long obj_item_begin = 0;   // first objective item of listAC
    double cNrm1 = 0, cNrm2Sq = 0, cNrmInf = 0;
    double c_rank1 = 0.0;            // C = c_rank1 * e e^T + sparse remainder (0: no rank-one part)
    bool S_has_C = false;            // the last weighted sum written to S included the objective
    DBuf<double> csA, csB, cs_scratch;   // column sums of factors for the rank-one objective
    std::vector<int32_t> act_idx_h;
    std::vector<int32_t> P_row_h, P_col_h;          // kept for lb2_get_pattern / dual infeasibility
    DBuf<int> act_idx;
    ItemListBufs listA, listAC;
    DBuf<double> carry1, carry2, carry3;
    DBuf<double> C_onP, S, T_val;
    DBuf<int> T_ptr, T_con, adj_ptr, adj_col, adj_pos;
    DBuf<long long> D_pos;
    DBuf<double> Z1, Z2;                            // dense path: packed sym(UV^T)
    // vertex-centric fast path (layout.hpp VcLayout)
    bool vc_on = false;
    long long vc_nnz_res = 0;                       // non-zeros of the residual (multi-entry) constraints
    long long nnzC_adj = 0;                         // objective entries of the vertex-centric adjacency
    VcDev vc;
    DBuf<int> vc_order, vc_order_l, vc_u_ptr, vc_u_mid, vc_u_col, vc_u_tag, vc_d_con, vc_l_ptr, vc_l_row, vc_l_col, vc_l_con, vc_Tr_ptr,
        vc_Tr_con;
    DBuf<double> vc_u_val, vc_d_coef, vc_l_coef, vc_Tr_val;
    ItemListBufs vc_listRes;
    // row sharding (Solver::shard_rows): rows [row_lo, row_hi) of this cone are owned by this rank; `lead` marks the
    // rank that evaluates the unsplittable parts (residual item list, rank-one objective term, dense / generic cones)
    long long row_lo = 0, row_hi = 0;
    bool lead = true;
    std::vector<int32_t> vc_order_h, vc_order_l_h, row_weight_h, vc_l_ptr_h;   // host copies, kept for the partition
    const double *pw = nullptr;                     // weights of the pending (C + A^*(w)) product (set by cone_wsum)
    bool pw_compact = false;
    DBuf<double> cv;                                // constrVal of the cone (compact, n_act + 1)
    DBuf<double> t1, t2;                            // compact AUV outputs (n_act + 1)
    long long cg_iter_last = 0;
};

struct AlmState {   // lorads_alm_state, def_lorads_solver.h:130-145
    long long outerIter = 0, innerIter = 0;
    double rho = 0, pinf_inf = 1e30, pinf_1 = 1e30, pinf_2 = 0, gap = 1e30, pobj = 1e30, dobj = 1e30;
    double dinf_inf = 1e30, dinf_1 = 1e30, dinf_2 = 0, tau = 0;
};
struct AdmmState {  // lorads_admm_state, def_lorads_solver.h:147-161
    long long iter = 0, nBlks = 0, cg_iter = 0;
    double rho = 0, dinf_1 = 1e30, dinf_inf = 1e30, pinf_1 = 1e30, pinf_inf = 1e30, pinf_2 = 1e30, dinf_2 = 1e30;
    double pobj = 1e30, dobj = 1e30, gap = 1e30;
};

struct Solver {
    int device = 0;
    Ctx ctx;
    long long m = 0, nCones = 0;
    std::vector<long long> blkDims;
    std::vector<double> b_h;
    // reader arrays, kept until preprocess
    struct ConeInput { std::vector<int64_t> beg, idx; std::vector<double> elem; bool set = false; };
    std::vector<ConeInput> inputs;
    std::vector<ConeDev> cones;
    // LP cone (diagonal block): nLp columns appended behind the cone factors in every N-vector, so that the
    // L-BFGS / BLAS-1 kernels treat [cones | LP] as one vector exactly like the reference's Dtemp (lorads_alm.c:419-426)
    long long nLp = 0;
    ConeInput lp_input;                 // reader arrays (by constraint, column 0 = objective), kept until preprocess
    std::vector<double> lp_c_h;
    DBuf<double> lp_c, lp_rval, lp_cval, lp_nrm2sq, lp_x;
    DBuf<int> lp_rbeg, lp_rcol, lp_cbeg, lp_crow, lp_lvl_ptr, lp_lvl_col;
    std::vector<LpSeg> lp_segs;     // launches of the Gauss-Seidel sweep (kernels.cuh)
    LpDev lp;
    void set_lp(long long nLpCols, const lb2_int *beg, const lb2_int *idx, const double *elem);
    void build_lp();
    void get_lp_vec(char which, double *out);
    void set_lp_vec(char which, const double *in);
    bool preprocessed = false, vars_ready = false;
    bool single_identity = false;      // one cone whose active set is all constraints: no scatter passes

    // m-vectors (capacity m + 1: slot m receives the objective value of fused evaluations)
    DBuf<double> b, lam, s, q12, M1, cvfull;   // q12 = [q1 (m+1) | q2 (m+1)] contiguous: one all-reduce covers both
    struct VecView { double *p = nullptr; } q1, q2, q3;    // q3: A(RR^T) as third output of the fused pass
    // N-vectors
    long long N = 0;      // cone part: sum of n * ld
    long long Nt = 0;     // N + nLp: length of the concatenated [cones | LP] vectors
    DBuf<double> R, U, V, G, M2, Bls, cg_r, cg_p, cg_Q, Dtemp;
    std::vector<DBuf<double>> lb_s, lb_y;
    // cudaFree synchronises the device and was measured to stall for up to 1.5 s on this platform, so nothing is
    // freed while a solve is running: buffers replaced by rank augmentation are parked here until the solver dies
    std::vector<DBuf<double>> retired;
    size_t retired_bytes = 0;
    void retire(DBuf<double> &b) {
        if (!b.p) return;
        retired_bytes += b.n * sizeof(double);
        retired.emplace_back(std::move(b));
        if (retired_bytes > ((size_t)32 << 30)) { retired.clear(); retired_bytes = 0; }   // cap the parked memory at 32 GiB
    }
    // Lanczos workspace of the dual infeasibility (allocated once, grown on demand)
    DBuf<double> lz_basis, lz_w, lz_x0, lz_h, lz_scratch, lz_tab;
    DBuf<int> aug_pos;            // scatter list of the diagonal entries of freshly appended factor columns
    DBuf<double> aug_val;
    int lbfgs_len = 2, lb_head = 0;
    bool vf_lbfgs = true;      // vector-free (Gram-table) L-BFGS for history length 2
    // the pair kernel works on double2: an odd vector length (LP block) and longer histories use the plain recursion
    bool use_vf() const { return vf_lbfgs && lbfgs_len == 2 && (vn & 1) == 0; }
    bool vf_valid = false;     // the G-dots in the Gram table belong to the current gradient
    // scalars
    DBuf<double> S;
    double *S_host = nullptr;          // pinned mirror
    DBuf<double> red_partials, dense_part;
    DBuf<double> flush_buf;            // lb2_bench_kernel: L2 eviction buffer
    DBuf<double> xfer_stage;           // column-major staging of upload_factor / download_factor (device)
    DBuf<unsigned int> red_counter;
    double *push_host = nullptr;       // pinned staging for {tau, rho}
    // replayable CUDA graphs of the steady-state inner iteration, keyed by (L-BFGS ring head, history depth)
    struct IterGraph { cudaGraphExec_t exec = nullptr; long long launches = 0; };
    std::map<int, IterGraph> iter_graphs;
    bool use_graphs = true;
    void drop_graphs();
    void push_scalars(double tau, double rho);
    // back half of iteration k (+ front half of k+1 when next_counter >= 0), then the scalar read-back and sync
    void iter_back_front(double rho, double tau, long long next_counter);

    // ranks
    std::vector<long long> rank, rank_max;
    // communicator.  Two ways to shard a solve over the GPUs of a node:
    //   rows (LORADS_B200_SHARD=rows, north_star): the concatenated factor rows of all cones are cut into `world` contiguous slabs of
    //     equal work -- small cones go to one rank as a whole (partition by cone block), a huge cone is split by rows.
    //     Every rank keeps full copies of the gathered operands (R, the direction, U, V) and computes only its slab:
    //     its rows of the gradient / CG vectors, the pattern entries of its columns in A(UV^T).  Per A() evaluation
    //     one all-reduce of the m-vectors, per dot product a scalar all-reduce, per new direction one all-gather.
    //   columns (default): every rank holds all rows of r/world factor columns; no factor traffic, but every rank
    //     still walks the whole pattern.  Faster than rows on random sparse graphs (measured).
    int world = 1, myrank = 0;
    void *nccl = nullptr;
    int shard_mode = 0;                        // 1 = rows, 0 = columns
    bool shard_rows() const { return world > 1 && shard_mode == 1; }
    bool shard_cols() const { return world > 1 && shard_mode == 0; }
    long long vo = 0, vn = 0;                  // owned range [vo, vo + vn) of the concatenated factor vectors
    std::vector<long long> part_rows;          // world + 1 boundaries in the concatenated row index space
    std::vector<long long> own_off, own_cnt;   // owned element range of every rank (all-gather schedule)
    long long equal_rows = 0;                  // > 0: one splittable cone cut into slabs of this many rows (the last one may
                                               // be shorter): the all-gather is a single in-place ncclAllGather
    long long Nalloc = 0;                      // allocated length of the factor vectors (>= Nt; padded for equal slabs)
    void setup_row_partition();                // after preprocess + comm init
    void compute_owned_ranges();               // after alloc_vars (leading dimensions known)
    void allgather_owned(double *X);           // every rank broadcasts its slab of a concatenated vector
    void allgather_cone(const ConeDev &K, double *X);   // the same restricted to one cone's slice
    void bcast_ranges(double *X, long long lo, long long hi);
    // one-shot all-reduce over NVLink peer memory (kernels.cuh P2PDev); messages above p2p.cap go through NCCL
    P2PDev p2p;
    bool p2p_on = false;
    DBuf<double> p2p_x;
    DBuf<unsigned long long> p2p_f, p2p_epoch;
    DBuf<unsigned int> p2p_ticket;
    DBuf<double *> p2p_peer_x;
    DBuf<unsigned long long *> p2p_peer_f;
    std::vector<void *> p2p_opened;      // peer mappings to close
    std::vector<std::vector<int>> my_cols;   // per cone: global column ids held here

    // problem constants / convergence state (lorads_solver fields, def_lorads_solver.h:76-105)
    double cObjNrm1 = 0, cObjNrm2 = 0, cObjNrmInf = 0, bNrm1 = 0, bNrm2 = 0, bNrmInf = 0;
    double pObj = 0, dObj = 0, scaleObjHis = 1.0;
    double dimac_pinf = 0, dimac_gap = 0, dimac_dinf = 0;
    long long cgIter = 0;
    double cgTime = 0;
    int status = LB2_STATUS_UNKNOWN;
    AlmState alm;
    AdmmState admm;
    int MAX_ALM_SUB_ITER = 5000;   // the reference's global, lorads_alm.c:7

    ~Solver();
    // setup
    void create(long long nRows, long long nCones, const lb2_int *blkDims, const double *rhs, int dev);
    void set_cone(long long i, const lb2_int *beg, const lb2_int *idx, const double *elem);
    void preprocess();
    void determine_rank(double timesLogRank);
    void init_vars(long long lbfgsLen, double initRho);
    void alloc_vars();
    // transfers
    double *factor_ptr(char which);
    double *vec_ptr(char which);
    void set_factor(char which, long long c, const double *colMajor);
    void get_factor(char which, long long c, double *colMajor) const;
    void upload_factor(double *dst, const ConeDev &K, const double *colMajor);
    void download_factor(const double *src, const ConeDev &K, double *colMajor) const;
    void read_slots();          // S -> S_host (synchronises the stream)
    // S -> S_host as of the work enqueued so far, on a second stream: kernels enqueued afterwards on the main stream keep
    // running while the host waits for the scalars
    void read_slots_side();
    cudaStream_t side_stream = nullptr;
    cudaEvent_t side_event = nullptr;
    void sync();
    void allreduce(double *p, long long count);
    // operators
    // A(sym(U V^T)) of cone c with the A or A+C item list: writes (scale*values) to out1 (n_rows of the list)
    void cone_auv(ConeDev &K, bool with_obj, const double *U, const double *V, bool same, double scale, double *out,
                  double *obj = nullptr);
    void cone_auv_dual(ConeDev &K, const double *Rm, const double *Dm, double *out1, double *out2, double *obj1 = nullptr,
                       double *obj2 = nullptr);
    // materialize: write S = C + A^*(w) on the pattern even for vertex-centric cones (dual infeasibility SpMV)
    void cone_wsum(ConeDev &K, const double *w, bool w_compact, bool addC, bool materialize = false);
    void cone_mul(ConeDev &K, const double *X, double a, double bcoef, const double *Z, const double *Z2, double *Y, double *red);
    // constrVal[c] = A(sym(U V^T)) for all cones, then constrValSum (LORADSInitConstrValAll + InitConstrValSum)
    void init_constr_val_all(const double *Um, const double *Vm, bool same);
    void constr_val_sum();
    void update_constr_val(long long c, const double *Um, const double *Vm);   // LORADSUpdateConstrVal
    void expand_cv(ConeDev &K, double *full);                                    // compact cv -> length m
    void expand_cv_from(ConeDev &K, const double *src, double *full);
    double cal_grad(double rho);                                                 // ALMCalGrad, returns sum ||G||^2
    void lbfgs_direction(long long counter);
    void q12p12();
    void lp_q12p12();
    bool p12_from_rows() const;
    bool tri_pending = false; // q3 holds A(RR^T) that the line-search pass still has to move into constrValSum
    bool tri_ok() const;      // the dual gather pass can also deliver A(RR^T) (single cone over all constraints, sparse scratch)
    void primal_infeasibility(const double *Rm);                                 // fills S_host[SL_PINF] lazily
    double cal_obj(const double *Rm);                                            // <C, R R^T> / scaleObjHis
    double cal_dual_obj();
    void average_uv();                                                           // R = (U+V)/2
    void cg_matvec(ConeDev &K, const double *x, const double *Vnoupd, double *res, const double *Z2, double *red);
    void update_sdp_var_one(long long c, double *upd, const double *noupd, double rho, double tol, long long maxit);
    void update_sdp_var(double rho, double tol, long long maxit);
    void update_dual_var(double rho);
    // phases
    void enqueue_front(double rho, long long counter);
    long long finish_front(double rho, double *tau, double *p12);
    void enqueue_back(double rho, double tau, bool front_follows = false);
    void finish_back(double *lagNormSq, double *pinf1);
    long long run_inner_iters(double rho, long long iters, double *out);
    int alm_inner_front(double rho, long long counter, double *tau, double *p12, long long *rootNum);
    void alm_inner_back(double rho, double tau, double *lagNormSq, double *pinf1);
    void update_dimacs_alm();
    void update_dimacs_admm();
    int alm_optimize(lb2_params *p, double rho_update_factor_unused, double timeSolveStart, bool reopt, bool early_stop,
                     double reopt_rho_factor);
    void alm_to_admm(lb2_params *p);
    int admm_optimize(lb2_params *p, long long iter_celling, double timeSolveStart, bool reopt);
    double reopt(lb2_params *p, double *reopt_param, long long *alm_iter, long long *admm_iter, double timeSolveStart,
                 int *admm_bad_iter_flag, int reopt_level);
    void obj_scale_dualvar(double f);
    void dual_infeasibility();
    bool check_all_rank_max(double aug_factor) const;
    bool aug_rank(double aug_factor);
    int solve(lb2_params *p, lb2_result *res);
};

double wall_time();
long long line_search(double rho, const double *sums, double p1, double p2, double *tau);

}  // namespace lb2
