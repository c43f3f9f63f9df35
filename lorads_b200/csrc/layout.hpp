// layout.hpp -- host-side pre-solve of one PSD cone: from the SDPA reader's CSC arrays to the index
// structures the device kernels consume.
//
// Replaces (reference, all under src_semi/data/): sdpDataMatSetData lorads_sdp_data.c:811-828 (coefficient
// classification), sdp{Dense,Sparse}ConeProcDataImpl lorads_sdp_conic.c:144-227 (constraint rows, compact
// row list), AConePresolveData lorads_sdp_conic.c:868-1076 (union pattern P, dense/sparse scratch decision,
// nnzIdx2ResIdx maps -- here by sort + binary search instead of the chained hash of lorads_sdp_data.c:31-45).
//
// The reference keeps one heap object + vtable per constraint; here every structure is a flat array:
//   P            union pattern (row >= col), sorted by (col, row)
//   item lists   the non-zeros of [A_1..A_m ; C] stacked, sorted by constraint, each item carrying its
//                pattern coordinates and its inner-product weight (2a off the diagonal, a on it)
//   T            the same non-zeros transposed: CSC by pattern position (for S = C + sum_i w_i A_i)
//   adj          the symmetric adjacency of P in CSR by row (for Y = S X without scatter)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace lb2 {

// items per CTA tile of the A(UV^T) kernel: large lists use 512, small ones 128 so that the grid still fills
// the 148 SMs several times over (kernels.cu instantiates both)
constexpr int kAuvTileLarge = 512, kAuvTileSmall = 128;
constexpr long long kAuvSmallListItems = 400000;

struct ItemList {
    // rows 0..n_rows-1; when has_obj the last row is the objective C
    int64_t n_rows = 0;
    bool has_obj = false;
    std::vector<int32_t> ptr;     // n_rows + 1
    std::vector<int32_t> irow;    // pattern row  (sparse path)  | packed position (dense path)
    std::vector<int32_t> icol;    // pattern col  (sparse path)  | unused (dense path)
    std::vector<double> coef;     // 2a off-diagonal, a on the diagonal (lorads_sdp_data.c:545-549)
    // rows that straddle tile boundaries: their tile partials are summed by the fix-up kernel
    std::vector<int32_t> split_row, split_first_slot, split_tile_a, split_tile_b;
    int tile = kAuvTileLarge;                 // items per tile chosen for this list
    std::vector<int32_t> tile_row_lo, tile_row_hi;   // first / last row with an item in each tile
    int64_t n_items() const { return (int64_t)irow.size(); }
    int64_t n_tiles() const { return (n_items() + tile - 1) / tile; }
};

// Vertex-centric layout of a sparse-scratch cone (the fast path of the two gather kernels).  The coefficients are
// split by class so that nothing has to be materialised on the pattern per evaluation.  Row i of the symmetric
// adjacency `adjU` lists (neighbour, tag, value) entries, the weight-dependent ones first:
//   [u_ptr[i], u_mid[i])    tag >= 0 : entry of SINGLETON constraint `tag` (exactly one stored non-zero: MaxCut
//                                      diagonals, the edge constraints of Lovasz theta, the samples of matrix
//                                      completion), S value = w[tag] * value;
//                           tag <= -2: pattern position -2-tag of the remaining ("residual", multi-entry) constraints,
//                                      whose weighted sum is still materialised on the pattern (T restricted to them);
//   [u_mid[i], u_ptr[i+1])  tag == -1: entry of the objective C, value inline.
// For A(UV^T) the singleton entries are needed once (lower triangle): the diagonal ones per column in d_con/d_coef
// (read with the owner row), the others in lowA, a flat list in (col,row) order walked entry-parallel (every entry is
// an output of its own, so long columns -- matrix completion: 100 entries per column -- split over many lane groups);
// residual constraints go through the generic item kernel (listRes).  `order`/`order_l` list the rows by decreasing work so that the lane groups of a warp walk rows
// of equal length.
struct VcLayout {
    bool on = false;
    int64_t n_single = 0, n_res = 0, nnz_res = 0;
    std::vector<int32_t> order, order_l;
    std::vector<int32_t> u_ptr, u_mid, u_col, u_tag; std::vector<double> u_val;
    std::vector<int32_t> d_con; std::vector<double> d_coef;
    std::vector<int32_t> l_ptr, l_row, l_col, l_con; std::vector<double> l_coef;
    std::vector<int32_t> Tr_ptr, Tr_con; std::vector<double> Tr_val;
    ItemList listRes;
};

struct ConeLayout {
    int64_t n = 0;          // block dimension
    int64_t m = 0;          // global number of constraints
    bool dense_path = false;        // dense scratch matrices (lorads_sdp_conic.c:884,969-973,988-989)
    bool dense_cone = true;         // > 30 % of the constraints touch the block (lorads_user_data.c:68-70)
    int64_t n_act = 0;              // constraints with at least one entry in this block
    std::vector<int32_t> act_idx;   // compact row -> global constraint index
    int64_t nnzA = 0, nnzC = 0;
    int64_t n_nonzero_coeff = 0;    // "nnzStat": number of non-zero constraint matrices (lorads_sdp_data.c:190)
    bool c_is_dense_type = false, any_dense_coeff = false;
    double c_rank1 = 0.0;           // C = c_rank1 * e e^T + (sparse remainder stored as the objective column)
    double cNrm1 = 0, cNrm2Sq = 0, cNrmInf = 0;

    // sparse path
    std::vector<int32_t> P_row, P_col;
    int64_t n_diag = 0;
    std::vector<double> C_onP;                      // C expanded on P (dense path: packed C)
    ItemList listA, listAC;
    std::vector<int32_t> T_ptr, T_con; std::vector<double> T_val;   // CSC by pattern position
    std::vector<int32_t> adj_ptr, adj_col, adj_pos;                 // symmetric adjacency CSR
    VcLayout vc;                                                    // class-split vertex-centric layout
    // dense path: positions touched by any constraint, CSR over those positions
    std::vector<long long> D_pos;                   // unique packed positions with constraint entries
    int64_t psize() const { return dense_path ? n * (n + 1) / 2 : (int64_t)P_row.size(); }
};

inline void unpack_lower(int64_t n, int64_t packed, int64_t &row, int64_t &col) {
    // inverse of PACK_IDX (lorads_utils.h:45); column j starts at j*n - j*(j-1)/2
    double nn = (double)n;
    int64_t j = (int64_t)std::floor(((2 * nn + 1) - std::sqrt((2 * nn + 1) * (2 * nn + 1) - 8.0 * (double)packed)) / 2.0);
    if (j < 0) j = 0;
    if (j > n - 1) j = n - 1;
    while (j > 0 && j * n - j * (j - 1) / 2 > packed) --j;
    while (j + 1 < n && (j + 1) * n - (j + 1) * j / 2 <= packed) ++j;
    col = j;
    row = packed - (j * n - j * (j - 1) / 2) + j;
}

inline void finish_item_list(ItemList &L) {
    L.tile = (L.n_items() < kAuvSmallListItems) ? kAuvTileSmall : kAuvTileLarge;
    const int64_t T = L.tile;
    // static table of the rows whose items straddle tile boundaries
    L.split_row.clear(); L.split_first_slot.clear(); L.split_tile_a.clear(); L.split_tile_b.clear();
    const int64_t nt = L.n_tiles();
    L.tile_row_lo.assign(nt, 0); L.tile_row_hi.assign(nt, 0);
    for (int64_t r = 0; r < L.n_rows; ++r) {
        int64_t a = L.ptr[r], b = L.ptr[r + 1];
        if (b <= a) continue;
        int64_t ta = a / T, tb = (b - 1) / T;
        for (int64_t t = ta; t <= tb; ++t) L.tile_row_hi[t] = (int32_t)r;   // rows ascend: the last writer wins
        if (ta == tb) continue;
        // in tile ta the row is the LAST row of the tile: slot 2*ta+1, unless it starts exactly at the tile
        // start (then it is also the first row: slot 2*ta)
        int64_t first_slot = (a == ta * T) ? 2 * ta : 2 * ta + 1;
        L.split_row.push_back((int32_t)r);
        L.split_first_slot.push_back((int32_t)first_slot);
        L.split_tile_a.push_back((int32_t)ta);
        L.split_tile_b.push_back((int32_t)tb);
    }
    // first row of a tile = row containing the tile's first item
    {
        int64_t r = 0;
        for (int64_t t = 0; t < nt; ++t) {
            const int64_t first = t * T;
            while (r + 1 < L.n_rows && L.ptr[r + 1] <= first) ++r;
            L.tile_row_lo[t] = (int32_t)r;
        }
    }
}

// phase timer of the pre-solve (LORADS_B200_PRESOLVE_TIMING=1 prints the split to stderr)
struct PhaseTimer {
    bool on = getenv("LORADS_B200_PRESOLVE_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "  presolve %-28s %8.3f s\n", what, std::chrono::duration<double>(now - t).count());
        t = now;
    }
};

// Runs fn(lo, hi) over [0, total) cut into nth contiguous chunks, one thread each (nth == 1: inline).
template <class F>
inline void parallel_chunks(int64_t total, int nth, F fn) {
    if (nth <= 1 || total < 2 * (int64_t)nth) { fn((int64_t)0, total); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nth; ++t) th.emplace_back(fn, total * t / nth, total * (t + 1) / nth);
    for (std::thread &x : th) x.join();
}

// threads used by the pre-solve on large cones (LORADS_B200_PRESOLVE_THREADS overrides; every phase produces the
// same arrays for any thread count)
inline int presolve_threads(int64_t nnz_all) {
    if (const char *e = getenv("LORADS_B200_PRESOLVE_THREADS")) return std::max(1, atoi(e));
    return nnz_all > 2000000 ? (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())) : 1;
}


// Symmetric CSR (rows sorted by neighbour) from lower-triangular entries (row >= col) given in (col,row) order.
// alloc(total) is called once the slot count is known, then emit(q, k, other) once per adjacency slot q for entry k
// with `other` = the neighbour index.
template <class Alloc, class Emit>
inline void sym_csr_from_lower(int64_t n, int64_t cnt, const int32_t *row, const int32_t *col, std::vector<int32_t> &ptr,
                               Alloc alloc, Emit emit) {
    ptr.assign(n + 1, 0);
    for (int64_t k = 0; k < cnt; ++k) {
        ptr[row[k] + 1]++;
        if (row[k] != col[k]) ptr[col[k] + 1]++;
    }
    for (int64_t i = 0; i < n; ++i) ptr[i + 1] += ptr[i];
    alloc((int64_t)ptr[n]);
    std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
    for (int64_t k = 0; k < cnt; ++k)            // (row, col < row): columns ascend along the list
        if (row[k] != col[k]) emit(cur[row[k]]++, k, col[k]);
    for (int64_t k = 0; k < cnt; ++k)            // diagonal, then (col, row > col): rows ascend inside a column
        emit(cur[col[k]]++, k, row[k]);
}

// Build the layout of one cone from the reader's arrays (column 0 = C, column i = A_i).
inline ConeLayout build_cone_layout(int64_t n, int64_t m, const int64_t *beg_in, const int64_t *idx, const double *elem,
                                    bool allow_rank_one = true) {
    ConeLayout L;
    PhaseTimer pt;
    L.n = n; L.m = m;
    const double packedSize = (double)(n * (n + 1) / 2);
    // raw arrays from the C ABI (lb2_layout_build, lb2_host_presolve, lb2_set_cone_data): refuse what would send the
    // column walk below out of range
    if (!beg_in || beg_in[0] != 0) throw std::invalid_argument("cone data: column pointers must start at 0");
    for (int64_t c = 0; c <= m; ++c)
        if (beg_in[c + 1] < beg_in[c]) throw std::invalid_argument("cone data: column pointers must not decrease");
    if (beg_in[m + 1] > 0 && (!idx || !elem)) throw std::invalid_argument("cone data: null index / value array");
    for (int64_t k = 0; k < beg_in[m + 1]; ++k)
        if (idx[k] < 0 || idx[k] >= n * (n + 1) / 2) throw std::invalid_argument("cone data: packed index outside [0, n(n+1)/2)");
    if (beg_in[m + 1] > (int64_t)2000000000) throw std::runtime_error("cone has more than 2^31 non-zeros");

    // --- per-column sorted copies (dataMatCreateSparseImpl sorts when needed, lorads_sdp_data.c:106-108)
    std::vector<int64_t> sidx(idx, idx + beg_in[m + 1]);
    std::vector<double> sval(elem, elem + beg_in[m + 1]);
    {
        std::vector<int64_t> perm;
        for (int64_t c = 0; c <= m; ++c) {
            int64_t a = beg_in[c], b = beg_in[c + 1];
            if (b - a < 2 || std::is_sorted(sidx.begin() + a, sidx.begin() + b)) continue;
            perm.resize(b - a);
            std::iota(perm.begin(), perm.end(), (int64_t)0);
            std::stable_sort(perm.begin(), perm.end(), [&](int64_t x, int64_t y) { return idx[a + x] < idx[a + y]; });
            for (int64_t k = 0; k < b - a; ++k) { sidx[a + k] = idx[a + perm[k]]; sval[a + k] = elem[a + perm[k]]; }
        }
    }

    pt.lap("copy + per-column order");
    // objective norms on the ORIGINAL C (dataMatSparseNrm1/Nrm2Square/NrmInf lorads_sdp_data.c:148-183; the dense
    // variants :227-272 are the same sums over the packed array)
    {
        // the entries of C are sorted by packed index: an entry is diagonal iff it is the first position of its matrix
        // column, found by walking the column starts (a closed-form unpack only after a long jump)
        int64_t j = 0, start = 0, next = n;
        for (int64_t k = beg_in[0]; k < beg_in[1]; ++k) {
            const int64_t p = sidx[k];
            if (p < start) { j = 0; start = 0; next = n; }
            if (p - next > 8 * n) {
                int64_t r_, c_; unpack_lower(n, p, r_, c_);
                j = c_; start = j * n - j * (j - 1) / 2; next = start + (n - j);
            }
            while (p >= next) { start = next; ++j; next = start + (n - j); }
            const double v = sval[k], a = std::fabs(v);
            const bool diag = (p == start);
            L.cNrm1 += diag ? a : 2 * a;
            L.cNrm2Sq += diag ? v * v : 2 * v * v;
            L.cNrmInf = std::max(L.cNrmInf, a);
        }
    }

    // --- rank-one objective: a fully populated C whose entries are (almost) all one value c is stored as
    // c * e e^T plus a sparse remainder (north_star "rank-one layouts"; e.g. C = -J of the Lovasz theta SDP).
    // <c ee^T, sym(UV^T)> = c (e^T U)(V^T e) and (c ee^T) X = c e (e^T X) are evaluated from column sums, so the
    // cone stays on the sparse scratch path instead of the O(n^2) dense one the reference takes.
    std::vector<int64_t> vbeg(beg_in, beg_in + m + 2);
    if (allow_rank_one && n >= 20 && vbeg[1] - vbeg[0] == n * (n + 1) / 2) {
        const double v0 = sval[vbeg[0]];
        int64_t same = 0;
        for (int64_t k = vbeg[0]; k < vbeg[1]; ++k) same += (sval[k] == v0);
        if ((double)same >= 0.9 * packedSize) {
            L.c_rank1 = v0;
            // compact the arrays: objective column keeps only the entries that differ from v0 (minus v0)
            int64_t w = 0;
            for (int64_t k = vbeg[0]; k < vbeg[1]; ++k)
                if (sval[k] != v0) { sidx[w] = sidx[k]; sval[w] = sval[k] - v0; ++w; }
            const int64_t removed = (vbeg[1] - vbeg[0]) - w;
            for (int64_t k = vbeg[1]; k < vbeg[m + 1]; ++k) { sidx[k - removed] = sidx[k]; sval[k - removed] = sval[k]; }
            for (int64_t c = 1; c <= m + 1; ++c) vbeg[c] -= removed;
            sidx.resize(vbeg[m + 1]); sval.resize(vbeg[m + 1]);
        }
    }
    const int64_t *beg = vbeg.data();
    const int64_t nnz_all = beg[m + 1];
    const int nthreads = presolve_threads(nnz_all);

    // --- classification (lorads_sdp_data.c:818-824) and statistics
    auto is_dense_type = [&](int64_t nnz) { return (double)nnz > 0.1 * packedSize; };
    L.nnzC = beg[1] - beg[0];
    L.c_is_dense_type = L.nnzC > 0 && is_dense_type(L.nnzC);
    L.any_dense_coeff = L.c_is_dense_type;
    L.n_act = 0; L.nnzA = 0; L.n_nonzero_coeff = 0;
    for (int64_t i = 0; i < m; ++i) {
        int64_t k = beg[i + 2] - beg[i + 1];
        if (k > 0) {
            L.act_idx.push_back((int32_t)i);
            L.n_act++; L.nnzA += k; L.n_nonzero_coeff++;
            if (is_dense_type(k)) L.any_dense_coeff = true;
        }
    }
    L.dense_cone = (double)L.n_act > 0.3 * (double)m;   // LUserDataChooseCone, lorads_user_data.c:68-70

    // rows / cols of every entry
    std::vector<int32_t> erow(nnz_all), ecol(nnz_all);
    {
        // entries inside a column are sorted by packed index => columns of the matrix ascend; walk instead
        // of calling the closed form for every entry
        parallel_chunks(nnz_all, nthreads, [&](int64_t lo, int64_t hi) {
            if (lo >= hi) return;
            int64_t c = (std::upper_bound(beg, beg + m + 2, lo) - beg) - 1;      // coefficient column holding entry lo
            int64_t j = 0, start = 0, next = n;   // packed range of matrix column j is [start, next)
            for (int64_t k = lo; k < hi; ++k) {
                while (k >= beg[c + 1]) { ++c; j = 0; start = 0; next = n; }        // next coefficient: restart the walk
                int64_t p = sidx[k];
                if (p < start) { j = 0; start = 0; next = n; }
                if (p - next > 8 * n) {           // far jump: use the closed form
                    int64_t r_, c_; unpack_lower(n, p, r_, c_);
                    j = c_; start = j * n - j * (j - 1) / 2; next = start + (n - j);
                }
                while (p >= next) { start = next; ++j; next = start + (n - j); }
                erow[k] = (int32_t)(p - start + j);
                ecol[k] = (int32_t)j;
            }
        });
    }

    pt.lap("norms, classes, rows/cols");
    // --- scratch type decision (AConePresolveData, lorads_sdp_conic.c:884-989)
    std::vector<int64_t> keys;     // (col, row) keys of the union pattern
    bool dense = (n < 20) || L.any_dense_coeff;
    if (!dense) {
        // The entries of one coefficient are already in (col,row) order, so the objective is one sorted run and the
        // constraints are one more whenever their keys happen to ascend (MaxCut, theta, matrix completion files do):
        // sort only what is not sorted, then merge the two runs.
        keys.resize(nnz_all);
        parallel_chunks(nnz_all, nthreads, [&](int64_t lo, int64_t hi) {
            for (int64_t k = lo; k < hi; ++k) keys[k] = (int64_t)ecol[k] * n + erow[k];
        });
        auto mid = keys.begin() + beg[1];
        if (!std::is_sorted(keys.begin(), mid)) std::sort(keys.begin(), mid);
        if (!std::is_sorted(mid, keys.end())) std::sort(mid, keys.end());
        std::inplace_merge(keys.begin(), mid, keys.end());
        keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
        double spRatio = (double)keys.size() / packedSize;
        if (spRatio >= 0.1) dense = true;
    }
    L.dense_path = dense;
    pt.lap("union pattern (sort)");

    auto fill_items = [&](ItemList &IL, bool with_obj) {
        IL.has_obj = with_obj;
        IL.n_rows = L.n_act + (with_obj ? 1 : 0);
        IL.ptr.assign(IL.n_rows + 1, 0);
        int64_t total = L.nnzA + (with_obj ? L.nnzC : 0);
        IL.irow.resize(total); IL.icol.resize(total); IL.coef.resize(total);
        int64_t w = 0;
        auto push = [&](int64_t k) {
            bool diag = erow[k] == ecol[k];
            if (dense) {
                IL.irow[w] = (int32_t)sidx[k];   // packed position == nnzIdx2ResIdx on the dense path
                IL.icol[w] = diag ? 1 : 0;
            } else {
                IL.irow[w] = erow[k]; IL.icol[w] = ecol[k];
            }
            IL.coef[w] = diag ? sval[k] : 2.0 * sval[k];
            ++w;
        };
        for (int64_t a = 0; a < L.n_act; ++a) {
            int64_t i = L.act_idx[a];
            for (int64_t k = beg[i + 1]; k < beg[i + 2]; ++k) push(k);
            IL.ptr[a + 1] = (int32_t)w;
        }
        if (with_obj) {
            for (int64_t k = beg[0]; k < beg[1]; ++k) push(k);
            IL.ptr[L.n_act + 1] = (int32_t)w;
        }
        finish_item_list(IL);
    };

    if (dense) {
        if (n * (n + 1) / 2 > (int64_t)2000000000) throw std::runtime_error("dense path: packed size exceeds 2^31");
        fill_items(L.listA, false);
        fill_items(L.listAC, true);
        // packed C
        L.C_onP.assign((size_t)(n * (n + 1) / 2), 0.0);
        for (int64_t k = beg[0]; k < beg[1]; ++k) L.C_onP[(size_t)sidx[k]] += sval[k];
        // constraint entries transposed by packed position (only touched positions)
        std::vector<int64_t> order(L.nnzA);
        std::vector<int64_t> pos(L.nnzA); std::vector<int32_t> con(L.nnzA); std::vector<double> val(L.nnzA);
        int64_t w = 0;
        for (int64_t a = 0; a < L.n_act; ++a) {
            int64_t i = L.act_idx[a];
            for (int64_t k = beg[i + 1]; k < beg[i + 2]; ++k) { pos[w] = sidx[k]; con[w] = (int32_t)a; val[w] = sval[k]; ++w; }
        }
        std::iota(order.begin(), order.end(), (int64_t)0);
        std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return pos[x] < pos[y]; });
        L.T_ptr.clear(); L.T_con.resize(L.nnzA); L.T_val.resize(L.nnzA); L.D_pos.clear();
        for (int64_t q = 0; q < L.nnzA; ++q) {
            int64_t o = order[q];
            if (q == 0 || pos[o] != L.D_pos.back()) { L.D_pos.push_back(pos[o]); L.T_ptr.push_back((int32_t)q); }
            L.T_con[q] = con[o]; L.T_val[q] = val[o];
        }
        L.T_ptr.push_back((int32_t)L.nnzA);
        return L;
    }

    // ----------------------------- sparse path -----------------------------
    const int64_t np = (int64_t)keys.size();
    L.P_row.resize(np); L.P_col.resize(np);
    L.n_diag = 0;
    for (int64_t p = 0; p < np; ++p) {
        L.P_col[p] = (int32_t)(keys[p] / n); L.P_row[p] = (int32_t)(keys[p] % n);
        if (L.P_col[p] == L.P_row[p]) L.n_diag++;
    }
    // pattern position of every entry (replaces the hash dictionary)
    // (merge-join: while the keys of consecutive entries ascend, the search continues from the previous hit with a
    //  galloping step; a binary search over the whole pattern is the fallback for a descending step)
    std::vector<int32_t> epos(nnz_all);
    parallel_chunks(nnz_all, nthreads, [&](int64_t klo, int64_t khi) {
        int64_t prev_key = -1, prev_pos = 0;
        for (int64_t k = klo; k < khi; ++k) {
            const int64_t key = (int64_t)ecol[k] * n + erow[k];
            int64_t lo, hi;
            if (key >= prev_key) {
                lo = prev_pos;
                int64_t step = 1;
                hi = lo;
                while (hi < np && keys[hi] < key) { lo = hi + 1; hi += step; step <<= 1; }
                if (hi > np) hi = np;
            } else { lo = 0; hi = prev_pos; }
            prev_pos = std::lower_bound(keys.begin() + lo, keys.begin() + hi, key) - keys.begin();
            prev_key = key;
            epos[k] = (int32_t)prev_pos;
        }
    });
    pt.lap("pattern positions");
    fill_items(L.listA, false);
    fill_items(L.listAC, true);
    pt.lap("item lists");

    L.C_onP.assign(np, 0.0);
    for (int64_t k = beg[0]; k < beg[1]; ++k) L.C_onP[epos[k]] += sval[k];

    // T: CSC by pattern position, constraint (compact index) ascending inside a position
    L.T_ptr.assign(np + 1, 0);
    for (int64_t a = 0; a < L.n_act; ++a) {
        int64_t i = L.act_idx[a];
        for (int64_t k = beg[i + 1]; k < beg[i + 2]; ++k) L.T_ptr[epos[k] + 1]++;
    }
    for (int64_t p = 0; p < np; ++p) L.T_ptr[p + 1] += L.T_ptr[p];
    L.T_con.resize(L.nnzA); L.T_val.resize(L.nnzA);
    {
        std::vector<int32_t> cur(L.T_ptr.begin(), L.T_ptr.end() - 1);
        for (int64_t a = 0; a < L.n_act; ++a) {
            int64_t i = L.act_idx[a];
            for (int64_t k = beg[i + 1]; k < beg[i + 2]; ++k) {
                int32_t q = cur[epos[k]]++;
                L.T_con[q] = (int32_t)a; L.T_val[q] = sval[k];
            }
        }
    }

    pt.lap("C on P, T");
    // symmetric adjacency CSR: row i lists (j, p) for every pattern entry (i,j) or (j,i)
    L.adj_ptr.assign(n + 1, 0);
    for (int64_t p = 0; p < np; ++p) {
        L.adj_ptr[L.P_row[p] + 1]++;
        if (L.P_row[p] != L.P_col[p]) L.adj_ptr[L.P_col[p] + 1]++;
    }
    for (int64_t i = 0; i < n; ++i) L.adj_ptr[i + 1] += L.adj_ptr[i];
    const int64_t nadj = L.adj_ptr[n];
    L.adj_col.resize(nadj); L.adj_pos.resize(nadj);
    {
        // emit in an order that leaves every row sorted by neighbour index: P is sorted by (col,row); the
        // entries (i, j<i) of row i come from pattern columns j ascending, then (i,i), then (j>i, i) from
        // pattern column i with rows ascending.  Two passes keep that order.
        // The first pass scatters by ROW over the whole array (cache hostile at n = 1e6): it is split over threads by
        // row range, every thread scans P once and fills only the rows it owns, in the same order as a single scan.
        std::vector<int32_t> cur(L.adj_ptr.begin(), L.adj_ptr.end() - 1);
        const int nth = nthreads;
        auto lower_pass = [&](int64_t r0, int64_t r1) {
            for (int64_t p = 0; p < np; ++p) {       // lower part seen from the row: (row, col<row)
                const int32_t r = L.P_row[p], c = L.P_col[p];
                if (r != c && r >= r0 && r < r1) { int32_t q = cur[r]++; L.adj_col[q] = c; L.adj_pos[q] = (int32_t)p; }
            }
        };
        if (nth == 1) lower_pass(0, n);
        else {
            std::vector<std::thread> th;
            for (int t = 0; t < nth; ++t) th.emplace_back(lower_pass, n * t / nth, n * (t + 1) / nth);
            for (std::thread &x : th) x.join();
        }
        // diagonal and upper part: row c gets (r >= c).  P is sorted by column, so a range of columns is a contiguous
        // range of p and the threads touch disjoint rows.
        parallel_chunks(n, nth, [&](int64_t c0, int64_t c1) {
            const int64_t p0 = std::lower_bound(L.P_col.begin(), L.P_col.end(), (int32_t)c0) - L.P_col.begin();
            const int64_t p1 = std::lower_bound(L.P_col.begin(), L.P_col.end(), (int32_t)c1) - L.P_col.begin();
            for (int64_t p = p0; p < p1; ++p) {
                int32_t r = L.P_row[p], c = L.P_col[p];
                int32_t q = cur[c]++; L.adj_col[q] = r; L.adj_pos[q] = (int32_t)p;
            }
        });
    }
    pt.lap("adjacency");
    // ----------------------------- vertex-centric class split -----------------------------
    {
        VcLayout &V = L.vc;
        // Two constructions of the class-split adjacency that must give identical arrays: "rows" (default) walks the
        // symmetric adjacency of P row by row, threads over row ranges, and classifies every pattern position through T
        // and an objective look-up; "sort" (LORADS_B200_VC_BUILD=sort) builds entry lists per class and scatters them
        // into a symmetric CSR.  tests/test_host_layout.py compares the two on every instance shape of the suite.
        const char *build_env = getenv("LORADS_B200_VC_BUILD");
        bool by_rows = !(build_env && std::string(build_env) == "sort");
        std::vector<int32_t> obj_k;           // objective entry at a pattern position (-1: none)
        if (by_rows) {
            obj_k.assign(np, -1);
            for (int64_t k = beg[0]; k < beg[1] && by_rows; ++k) {
                if (obj_k[epos[k]] >= 0) by_rows = false;          // repeated objective entry: the list construction keeps both
                else obj_k[epos[k]] = (int32_t)k;
            }
        }
        // lower-triangular entries (row >= col) of the three classes, each list in (col,row) order
        struct Ent { int32_t row, col, tag; double val; };
        std::vector<Ent> dyn, sta;
        // objective: entries of column 0 are already in (col,row) order
        if (!by_rows) {
            sta.reserve(beg[1] - beg[0]);
            for (int64_t k = beg[0]; k < beg[1]; ++k) sta.push_back({erow[k], ecol[k], -1, sval[k]});
        }
        // singleton constraints
        V.d_con.assign(n, -1); V.d_coef.assign(n, 0.0);
        std::vector<Ent> low;
        std::vector<uint8_t> single(L.n_act, 0);
        V.n_single = 0;
        for (int64_t a = 0; a < L.n_act; ++a) {
            const int64_t i = L.act_idx[a];
            if (beg[i + 2] - beg[i + 1] != 1) continue;
            const int64_t k = beg[i + 1];
            V.n_single++;
            single[a] = 1;
            if (!by_rows) dyn.push_back({erow[k], ecol[k], (int32_t)a, sval[k]});
            if (erow[k] == ecol[k] && V.d_con[erow[k]] < 0) { V.d_con[erow[k]] = (int32_t)a; V.d_coef[erow[k]] = sval[k]; }
            else low.push_back({erow[k], ecol[k], (int32_t)a, (erow[k] == ecol[k]) ? sval[k] : 2.0 * sval[k]});
        }
        V.n_res = L.n_act - V.n_single;
        auto by_pos = [&](const Ent &x, const Ent &y) { return x.col != y.col ? x.col < y.col : x.row < y.row; };
        if (!std::is_sorted(low.begin(), low.end(), by_pos)) std::stable_sort(low.begin(), low.end(), by_pos);
        V.l_ptr.assign(n + 1, 0);
        for (const Ent &e : low) V.l_ptr[e.col + 1]++;
        for (int64_t j = 0; j < n; ++j) V.l_ptr[j + 1] += V.l_ptr[j];
        V.l_row.resize(low.size()); V.l_col.resize(low.size()); V.l_con.resize(low.size()); V.l_coef.resize(low.size());
        for (size_t q = 0; q < low.size(); ++q) {
            V.l_row[q] = low[q].row; V.l_col[q] = low[q].col; V.l_con[q] = low[q].tag; V.l_coef[q] = low[q].val;
        }
        // residual constraints: item list (all n_act rows, singleton rows empty), T restricted to them, and one
        // weight-dependent adjacency entry per touched pattern position
        {
            ItemList &IL = V.listRes;
            IL.has_obj = false;
            IL.n_rows = L.n_act;
            IL.ptr.assign(IL.n_rows + 1, 0);
            int64_t total = 0;
            for (int64_t a = 0; a < L.n_act; ++a) {
                const int64_t i = L.act_idx[a], cntk = beg[i + 2] - beg[i + 1];
                if (cntk != 1) total += cntk;
            }
            V.nnz_res = total;
            IL.irow.resize(total); IL.icol.resize(total); IL.coef.resize(total);
            V.Tr_ptr.assign(np + 1, 0);
            int64_t w = 0;
            for (int64_t a = 0; a < L.n_act; ++a) {
                const int64_t i = L.act_idx[a], cntk = beg[i + 2] - beg[i + 1];
                if (cntk != 1)
                    for (int64_t k = beg[i + 1]; k < beg[i + 2]; ++k) {
                        IL.irow[w] = erow[k]; IL.icol[w] = ecol[k];
                        IL.coef[w] = (erow[k] == ecol[k]) ? sval[k] : 2.0 * sval[k];
                        ++w;
                        V.Tr_ptr[epos[k] + 1]++;
                    }
                IL.ptr[a + 1] = (int32_t)w;
            }
            finish_item_list(IL);
            for (int64_t p = 0; p < np; ++p) V.Tr_ptr[p + 1] += V.Tr_ptr[p];
            V.Tr_con.resize(total); V.Tr_val.resize(total);
            std::vector<int32_t> cur(V.Tr_ptr.begin(), V.Tr_ptr.end() - 1);
            for (int64_t a = 0; a < L.n_act; ++a) {
                const int64_t i = L.act_idx[a], cntk = beg[i + 2] - beg[i + 1];
                if (cntk == 1) continue;
                for (int64_t k = beg[i + 1]; k < beg[i + 2]; ++k) {
                    const int32_t q = cur[epos[k]]++;
                    V.Tr_con[q] = (int32_t)a; V.Tr_val[q] = sval[k];
                }
            }
            if (!by_rows)
                for (int64_t p = 0; p < np; ++p)
                    if (V.Tr_ptr[p + 1] > V.Tr_ptr[p]) dyn.push_back({L.P_row[p], L.P_col[p], (int32_t)(-2 - p), 0.0});
        }
        pt.lap("vc: entry classes");
        if (by_rows) {
            // Row i of adj lists its pattern positions by ascending neighbour -- the order the symmetric CSR of the list
            // construction produces.  At one position the singleton constraints come in ascending compact index (the order
            // of T), then the residual entry; the objective entry of the position goes to the second half of the row.
            // One record per pattern position, filled in a sequential sweep, so that the row walk below costs one
            // random access per adjacency slot: number of weight-dependent entries (nd >> 1), objective flag (nd & 1),
            // tag / value of the weight-dependent entry when there is exactly one (the common case).
            struct PosRec { double dval, oval; int32_t tag, nd; };
            std::vector<PosRec> rec(np);
            parallel_chunks(np, nthreads, [&](int64_t p0, int64_t p1) {
                for (int64_t p = p0; p < p1; ++p) {
                    PosRec r{0.0, 0.0, 0, 0};
                    int32_t cnt = 0;
                    for (int32_t q = L.T_ptr[p]; q < L.T_ptr[p + 1]; ++q)
                        if (single[L.T_con[q]]) { if (cnt++ == 0) { r.tag = L.T_con[q]; r.dval = L.T_val[q]; } }
                    if (V.Tr_ptr[p + 1] > V.Tr_ptr[p]) { if (cnt++ == 0) { r.tag = (int32_t)(-2 - p); r.dval = 0.0; } }
                    r.nd = cnt << 1;
                    if (obj_k[p] >= 0) { r.nd |= 1; r.oval = sval[obj_k[p]]; }
                    rec[p] = r;
                }
            });
            std::vector<int32_t> nd(n), ns(n);
            auto walk = [&](int64_t i, auto dyn_entry, auto sta_entry) {
                for (int32_t e = L.adj_ptr[i]; e < L.adj_ptr[i + 1]; ++e) {
                    const int32_t p = L.adj_pos[e], nb = L.adj_col[e];
                    const PosRec &r = rec[p];
                    if ((r.nd >> 1) == 1) dyn_entry(nb, r.tag, r.dval);
                    else if ((r.nd >> 1) > 1) {
                        for (int32_t q = L.T_ptr[p]; q < L.T_ptr[p + 1]; ++q)
                            if (single[L.T_con[q]]) dyn_entry(nb, L.T_con[q], L.T_val[q]);
                        if (V.Tr_ptr[p + 1] > V.Tr_ptr[p]) dyn_entry(nb, (int32_t)(-2 - p), 0.0);
                    }
                    if (r.nd & 1) sta_entry(nb, r.oval);
                }
            };
            parallel_chunks(n, nthreads, [&](int64_t i0, int64_t i1) {
                for (int64_t i = i0; i < i1; ++i) {
                    int32_t d = 0, t = 0;
                    walk(i, [&](int32_t, int32_t, double) { ++d; }, [&](int32_t, double) { ++t; });
                    nd[i] = d; ns[i] = t;
                }
            });
            int64_t total = 0;
            for (int64_t i = 0; i < n; ++i) total += (int64_t)nd[i] + ns[i];
            if (total > (int64_t)2000000000) throw std::runtime_error("adjacency exceeds 2^31 entries");
            V.u_ptr.assign(n + 1, 0); V.u_mid.assign(n, 0);
            for (int64_t i = 0; i < n; ++i) {
                V.u_mid[i] = V.u_ptr[i] + nd[i];
                V.u_ptr[i + 1] = V.u_mid[i] + ns[i];
            }
            V.u_col.resize(total); V.u_tag.resize(total); V.u_val.resize(total);
            parallel_chunks(n, nthreads, [&](int64_t i0, int64_t i1) {
                for (int64_t i = i0; i < i1; ++i) {
                    int64_t wd = V.u_ptr[i], ws = V.u_mid[i];
                    walk(i,
                         [&](int32_t nb, int32_t tag, double val) { V.u_col[wd] = nb; V.u_tag[wd] = tag; V.u_val[wd] = val; ++wd; },
                         [&](int32_t nb, double val) { V.u_col[ws] = nb; V.u_tag[ws] = -1; V.u_val[ws] = val; ++ws; });
                }
            });
            pt.lap("vc: adjacency by rows");
        } else {
        if (!std::is_sorted(dyn.begin(), dyn.end(), by_pos)) std::stable_sort(dyn.begin(), dyn.end(), by_pos);
        // symmetric adjacency: per row the weight-dependent entries, then the objective entries
        std::vector<int32_t> dptr, sptr, dslot, sslot;
        {
            std::vector<int32_t> r(dyn.size()), c(dyn.size());
            for (size_t q = 0; q < dyn.size(); ++q) { r[q] = dyn[q].row; c[q] = dyn[q].col; }
            std::vector<int32_t> okind, oother;
            sym_csr_from_lower(n, (int64_t)dyn.size(), r.data(), c.data(), dptr, [&](int64_t t) { dslot.resize(t); oother.resize(t); },
                               [&](int32_t q, int64_t k, int32_t other) { dslot[q] = (int32_t)k; oother[q] = other; });
            std::vector<int32_t> r2(sta.size()), c2(sta.size()), sother;
            for (size_t q = 0; q < sta.size(); ++q) { r2[q] = sta[q].row; c2[q] = sta[q].col; }
            sym_csr_from_lower(n, (int64_t)sta.size(), r2.data(), c2.data(), sptr, [&](int64_t t) { sslot.resize(t); sother.resize(t); },
                               [&](int32_t q, int64_t k, int32_t other) { sslot[q] = (int32_t)k; sother[q] = other; });
            const int64_t total = (int64_t)dslot.size() + (int64_t)sslot.size();
            if (total > (int64_t)2000000000) throw std::runtime_error("adjacency exceeds 2^31 entries");
            V.u_ptr.assign(n + 1, 0); V.u_mid.assign(n, 0);
            V.u_col.resize(total); V.u_tag.resize(total); V.u_val.resize(total);
            int64_t w = 0;
            for (int64_t i = 0; i < n; ++i) {
                V.u_ptr[i] = (int32_t)w;
                for (int32_t q = dptr[i]; q < dptr[i + 1]; ++q, ++w) {
                    V.u_col[w] = oother[q]; V.u_tag[w] = dyn[dslot[q]].tag; V.u_val[w] = dyn[dslot[q]].val;
                }
                V.u_mid[i] = (int32_t)w;
                for (int32_t q = sptr[i]; q < sptr[i + 1]; ++q, ++w) {
                    V.u_col[w] = sother[q]; V.u_tag[w] = -1; V.u_val[w] = sta[sslot[q]].val;
                }
            }
            V.u_ptr[n] = (int32_t)w;
        }
        pt.lap("vc: adjacency by sorted lists");
        }
        // rows by decreasing work (stable counting sort).  Sorting pays while the factors are L2 resident; on very
        // large blocks the scattered row order costs more DRAM locality than the balance gains (measured, n = 1e6)
        auto by_work = [&](std::vector<int32_t> &out, auto work) {
            out.resize(n);
            if (n > 250000) { std::iota(out.begin(), out.end(), 0); return; }
            std::vector<int32_t> deg(n);
            int32_t dmax = 0;
            for (int64_t i = 0; i < n; ++i) { deg[i] = work(i); dmax = std::max(dmax, deg[i]); }
            std::vector<int64_t> start((size_t)dmax + 2, 0);
            for (int64_t i = 0; i < n; ++i) start[dmax - deg[i] + 1]++;
            for (int32_t d = 0; d <= dmax; ++d) start[d + 1] += start[d];
            for (int64_t i = 0; i < n; ++i) out[start[dmax - deg[i]]++] = (int32_t)i;
        };
        by_work(V.order, [&](int64_t i) { return V.u_ptr[i + 1] - V.u_ptr[i]; });
        by_work(V.order_l, [&](int64_t i) { return V.u_ptr[i + 1] - V.u_mid[i]; });     // objective entries of the row
        pt.lap("vc: row orders");
        // Bounds of everything the gather kernels index with (compute-sanitizer is not available on the target pool, so the
        // layout is proved in range here, once, on the host): neighbours and rows inside the block, constraint tags
        // inside the compact constraint range, residual tags inside the pattern, monotone pointers.
        {
            auto bad = [](const char *what) { throw std::logic_error(std::string("vertex-centric layout out of range: ") + what); };
            const int64_t nu = (int64_t)V.u_col.size();
            if (V.u_ptr[0] != 0 || V.u_ptr[n] != nu) bad("adjacency pointers");
            for (int64_t i = 0; i < n; ++i)
                if (V.u_ptr[i] > V.u_mid[i] || V.u_mid[i] > V.u_ptr[i + 1]) bad("row split");
            for (int64_t e = 0; e < nu; ++e) {
                if (V.u_col[e] < 0 || V.u_col[e] >= n) bad("neighbour index");
                const int32_t t = V.u_tag[e];
                if (t >= L.n_act || (t <= -2 && -2 - (int64_t)t >= np)) bad("entry tag");
            }
            for (size_t q = 0; q < V.l_row.size(); ++q)
                if (V.l_row[q] < V.l_col[q] || V.l_row[q] >= n || V.l_col[q] < 0 || V.l_con[q] < 0 || V.l_con[q] >= L.n_act) bad("lower-triangular entry");
            for (int64_t i = 0; i < n; ++i)
                if (V.d_con[i] >= L.n_act || V.order[i] < 0 || V.order[i] >= n || V.order_l[i] < 0 || V.order_l[i] >= n) bad("row table");
        }
        // the fast path pays off when singleton constraints carry most of the constraint non-zeros
        V.on = L.nnzA == 0 || 2 * V.n_single >= L.n_act;
        if (const char *e = getenv("LORADS_B200_VC")) V.on = atoi(e) != 0;
    }
    pt.lap("vc: bounds proof");
    return L;
}

}  // namespace lb2
