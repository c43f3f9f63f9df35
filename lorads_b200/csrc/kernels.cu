// kernels.cu -- hand-written sm_100a kernels of the LoRADS low-rank hot path.
//
// All kernels are FP64 and HBM/L2-bandwidth bound (0.1-0.5 flop/B), so the design rules are: row-major
// r-padded factor rows (every gather is whole 32 B sectors), 16 B vector loads, a small lane group per
// gathered row so a warp keeps 8 independent row gathers in flight, warp-shuffle segmented reductions,
// and deterministic grid reductions (fixed summation order, no floating-point atomics).
//
// Reference loops replaced (src_semi/...):
//   auv_items_kernel   LORADSUVt lorads_alg_common.c:21-68 fused with coneAUV/objAUV
//                      lorads_sdp_conic.c:285-301,498-513 -> sparseAUV/denseAUV lorads_sdp_data.c:524-587
//   wsum_kernel        zeros + addObjCoeff + sdpDataWSum lorads_sdp_conic.c:327,437,539,633;
//                      dataMatSparseAddSparseSDPCoeff lorads_sdp_data.c:618-634
//   spmm_sym_kernel    dataMatSparseMultiRkMat lorads_sdp_data.c:491-504 (+ the scal/axpy/nrm2 that follow it
//                      in ALMSetGrad lorads_alm.c:34-37,51 and linSysProduct lorads_admm.c:386-390)
//   dense_*            fds_syr2k + repack lorads_alg_common.c:50-67; dataMatDenseMultiRkMat lorads_sdp_data.c:646-671
//   BLAS-1 kernels     lorads_vec_opts.c:116-329 call sites in lorads_alm.c / lorads_cgs.c / lorads_alg_common.c
#include "kernels.cuh"
#include "layout.hpp"

#include <algorithm>
#include <cstdlib>

namespace lb2 {

static inline int grid_for(long long n, int per_thread, const Ctx &c) {
    long long b = (n + (long long)kBlock * per_thread - 1) / ((long long)kBlock * per_thread);
    long long cap = (long long)c.num_sms * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

#define LB2_LAUNCH_CHECK(c)            \
    do {                               \
        (c).launches++;                \
        LB2_CUDA(cudaGetLastError());  \
    } while (0)

// =================================================================================================
// A(UV^T)
// =================================================================================================
// One CTA per tile of TILE consecutive items.  A group of G lanes evaluates one item: it gathers the needed
// factor rows with fully unrolled 16-byte loads (NP passes of 2*G columns, all issued before the FMAs so that a
// thread keeps up to 4*NP independent loads in flight), reduces with shuffles and parks coef*z in shared
// memory.  The tile is then reduced by constraint row: rows inside the tile are written, rows straddling tiles
// leave partials in `carry`, and the last CTA to finish (atomic ticket) adds the partials of every split row
// in a fixed order, so the result is deterministic.
template <int MODE, int G, int NP, int TILE>
__global__ void __launch_bounds__(kBlock) auv_items_kernel(ItemListDev L, const double *__restrict__ U,
                                                           const double *__restrict__ V, int ld, double scale1,
                                                           double scale2, double *__restrict__ out1,
                                                           double *__restrict__ out2, double *obj1, double *obj2,
                                                           double *carry1, double *carry2, unsigned int *ticket,
                                                           double *__restrict__ out3, double *carry3) {
    constexpr bool TRI = (MODE == AUV_TRI);                 // DUAL + third output A(U U^T) (scale 1)
    constexpr bool DUAL = (MODE == AUV_DUAL) || TRI;
    __shared__ double sp1[TILE];
    __shared__ double sp2[DUAL ? TILE : 1];
    __shared__ double sp3[TRI ? TILE : 1];
    __shared__ double sred[8];
    __shared__ bool s_last;

    const int tile = blockIdx.x;
    const int t0 = tile * TILE;
    const int t1 = min(t0 + TILE, (int)L.n_items);
    const int tid = threadIdx.x;

    if constexpr (MODE == AUV_FROMZ) {
        for (int k = t0 + tid; k < t1; k += kBlock) sp1[k - t0] = scale1 * L.coef[k] * U[L.irow[k]];
    } else {
        constexpr int NG = kBlock / G;          // groups per block
        constexpr int IPG = TILE / NG;          // consecutive items per group
        const int g = tid / G, gl = tid % G;
        // A group walks IPG CONSECUTIVE items.  Items of the objective are in column-major pattern order, so the
        // column-side rows (U_j, V_j) usually repeat from one item to the next: they stay in registers and are
        // gathered again only when the column (or, for the row side, the row) changes.  This halves the L1/L2
        // wavefronts of the gather.  The trip count is warp-uniform (every lane reaches every shuffle).
        int pi = -1, pj = -1;
        double2 x[NP], y[NP], p[NP], qv[NP];    // U_i, V_j (U_j when SAME), U_j, V_i
#pragma unroll
        for (int q = 0; q < NP; ++q) x[q] = y[q] = p[q] = qv[q] = make_double2(0.0, 0.0);
#pragma unroll 1
        for (int s_ = 0; s_ < IPG; ++s_) {
            const int k = t0 + g * IPG + s_;
            const bool live = k < t1;
            double a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, cf = 0.0;
            bool diag = true;
            if (live) {
                const int i = L.irow[k], j = L.icol[k];
                cf = L.coef[k];
                diag = (i == j);
                const double *ui = U + (size_t)i * ld + 2 * gl, *uj = U + (size_t)j * ld + 2 * gl;
                if constexpr (MODE == AUV_SAME) {
                    if (i != pi) {
#pragma unroll
                        for (int q = 0; q < NP; ++q)
                            x[q] = ((2 * gl + 2 * G * q) < ld) ? ld2(ui + 2 * G * q) : make_double2(0.0, 0.0);
                    }
                    if (j != pj) {
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            if (diag) y[q] = x[q];
                            else y[q] = ((2 * gl + 2 * G * q) < ld) ? ld2(uj + 2 * G * q) : make_double2(0.0, 0.0);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < NP; ++q) { a1 = fma(x[q].x, y[q].x, a1); a1 = fma(x[q].y, y[q].y, a1); }
                } else {
                    const double *vi = V + (size_t)i * ld + 2 * gl, *vj = V + (size_t)j * ld + 2 * gl;
                    if (i != pi) {
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            const bool in = (2 * gl + 2 * G * q) < ld;
                            x[q] = in ? ld2(ui + 2 * G * q) : make_double2(0.0, 0.0);      // U_i
                            qv[q] = in ? ld2(vi + 2 * G * q) : make_double2(0.0, 0.0);     // V_i
                        }
                    }
                    if (j != pj) {
#pragma unroll
                        for (int q = 0; q < NP; ++q) {
                            const bool in = (2 * gl + 2 * G * q) < ld;
                            if (!diag) {
                                y[q] = in ? ld2(vj + 2 * G * q) : make_double2(0.0, 0.0);  // V_j
                                p[q] = in ? ld2(uj + 2 * G * q) : make_double2(0.0, 0.0);  // U_j
                            } else { y[q] = qv[q]; p[q] = x[q]; }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < NP; ++q) {
                        a1 = fma(x[q].x, y[q].x, a1); a1 = fma(x[q].y, y[q].y, a1);          // U_i . V_j
                        a2 = fma(p[q].x, qv[q].x, a2); a2 = fma(p[q].y, qv[q].y, a2);        // U_j . V_i
                        if constexpr (DUAL) { a3 = fma(qv[q].x, y[q].x, a3); a3 = fma(qv[q].y, y[q].y, a3); }  // V_i . V_j
                        if constexpr (TRI) { a4 = fma(x[q].x, p[q].x, a4); a4 = fma(x[q].y, p[q].y, a4); }      // U_i . U_j
                    }
                }
                pi = i; pj = j;
            }
            a1 = group_sum<G>(a1);
            if constexpr (MODE != AUV_SAME) a2 = group_sum<G>(a2);
            if constexpr (DUAL) a3 = group_sum<G>(a3);
            if constexpr (TRI) a4 = group_sum<G>(a4);
            if (live && gl == 0) {
                if constexpr (MODE == AUV_SAME) {
                    sp1[k - t0] = scale1 * cf * a1;
                } else {
                    sp1[k - t0] = scale1 * cf * (diag ? a1 : (0.5 * a1 + 0.5 * a2));
                    if constexpr (DUAL) sp2[k - t0] = scale2 * cf * a3;
                    if constexpr (TRI) sp3[k - t0] = cf * a4;
                }
            }
        }
    }
    __syncthreads();

    const int row_lo = L.tile_row_lo[tile], row_hi = L.tile_row_hi[tile];
    const int nrows = row_hi - row_lo + 1;
    const int nit = t1 - t0;
    auto emit = [&](int r, bool complete, double v1, double v2, double v3) {
        if (complete) {
            out1[r] = v1;
            if constexpr (DUAL) out2[r] = v2;
            if constexpr (TRI) out3[r] = v3;
            if (r == L.obj_row) {
                if (obj1) *obj1 += v1;
                if constexpr (DUAL) { if (obj2) *obj2 += v2; }
            }
        } else {
            const int slot = (r == row_lo) ? 2 * tile : 2 * tile + 1;
            carry1[slot] = v1;
            if constexpr (DUAL) carry2[slot] = v2;
            if constexpr (TRI) carry3[slot] = v3;
        }
    };

    if (nrows == 1) {
        // the whole tile belongs to one row: block reduction in a fixed order
        double v1 = 0.0, v2 = 0.0, v3 = 0.0;
        for (int k = tid; k < nit; k += kBlock) {
            v1 += sp1[k];
            if constexpr (DUAL) v2 += sp2[k];
            if constexpr (TRI) v3 += sp3[k];
        }
        v1 = block_sum(v1, sred);
        if constexpr (DUAL) v2 = block_sum(v2, sred);
        if constexpr (TRI) v3 = block_sum(v3, sred);
        if (tid == 0) emit(row_lo, (L.ptr[row_lo] >= t0) && (L.ptr[row_lo + 1] <= t1), v1, v2, v3);
    } else if (nrows * 8 >= nit) {
        // short rows: one thread per row, sequential (deterministic) sum
        for (int r = row_lo + tid; r <= row_hi; r += kBlock) {
            const int pa = L.ptr[r], pb = L.ptr[r + 1];
            const int a = max(pa, t0) - t0, b = min(pb, t1) - t0;
            double v1 = 0.0, v2 = 0.0, v3 = 0.0;
            for (int k = a; k < b; ++k) {
                v1 += sp1[k];
                if constexpr (DUAL) v2 += sp2[k];
                if constexpr (TRI) v3 += sp3[k];
            }
            emit(r, pa >= t0 && pb <= t1, v1, v2, v3);
        }
    } else {
        // longer rows: one warp per row
        const int w = tid >> 5, lane = tid & 31;
        for (int r = row_lo + w; r <= row_hi; r += kBlock / 32) {
            const int pa = L.ptr[r], pb = L.ptr[r + 1];
            const int a = max(pa, t0) - t0, b = min(pb, t1) - t0;
            double v1 = 0.0, v2 = 0.0, v3 = 0.0;
            for (int k = a + lane; k < b; k += 32) {
                v1 += sp1[k];
                if constexpr (DUAL) v2 += sp2[k];
                if constexpr (TRI) v3 += sp3[k];
            }
            v1 = warp_sum(v1);
            if constexpr (DUAL) v2 = warp_sum(v2);
            if constexpr (TRI) v3 = warp_sum(v3);
            if (lane == 0) emit(r, pa >= t0 && pb <= t1, v1, v2, v3);
        }
    }

    if (L.n_split == 0) return;
    // ---- split rows: the last CTA adds the tile partials, one warp per split row, fixed order ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        const int w = tid >> 5, lane = tid & 31;
        for (int sidx = w; sidx < (int)L.n_split; sidx += kBlock / 32) {
            const int row = L.split_row[sidx], fs = L.split_first_slot[sidx], ta = L.split_tile_a[sidx];
            const int ne = L.split_tile_b[sidx] - ta + 1;
            double v1 = 0.0, v2 = 0.0, v3 = 0.0;
            for (int e = lane; e < ne; e += 32) {
                const int slot = (e == 0) ? fs : 2 * (ta + e);
                v1 += __ldcg(carry1 + slot);
                if constexpr (DUAL) v2 += __ldcg(carry2 + slot);
                if constexpr (TRI) v3 += __ldcg(carry3 + slot);
            }
            v1 = warp_sum(v1);
            if constexpr (DUAL) v2 = warp_sum(v2);
            if constexpr (TRI) v3 = warp_sum(v3);
            if (lane == 0) {
                out1[row] = v1;
                if constexpr (DUAL) out2[row] = v2;
                if constexpr (TRI) out3[row] = v3;
                if (row == L.obj_row) {
                    if (obj1) *obj1 += v1;
                    if constexpr (DUAL) { if (obj2) *obj2 += v2; }
                }
            }
        }
        if (tid == 0) *ticket = 0u;
    }
}

template <int MODE, int G, int NP>
static void launch_auv_tile(Ctx &c, const ItemListDev &L, const double *U, const double *V, int ld, double s1, double s2,
                            double *o1, double *o2, double *b1, double *b2, double *c1, double *c2, double *o3 = nullptr,
                            double *c3 = nullptr) {
    const int grid = (int)L.n_tiles;
    if (L.tile == kAuvTileSmall)
        auv_items_kernel<MODE, G, NP, kAuvTileSmall><<<grid, kBlock, 0, c.stream>>>(L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, c.ticket, o3, c3);
    else
        auv_items_kernel<MODE, G, NP, kAuvTileLarge><<<grid, kBlock, 0, c.stream>>>(L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, c.ticket, o3, c3);
    LB2_LAUNCH_CHECK(c);
}

template <int MODE, int G>
static void launch_auv_np(Ctx &c, int np, const ItemListDev &L, const double *U, const double *V, int ld, double s1,
                          double s2, double *o1, double *o2, double *b1, double *b2, double *c1, double *c2, double *o3,
                          double *c3) {
    switch (np) {
    case 1: launch_auv_tile<MODE, G, 1>(c, L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, o3, c3); break;
    case 2: launch_auv_tile<MODE, G, 2>(c, L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, o3, c3); break;
    case 3: launch_auv_tile<MODE, G, 3>(c, L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, o3, c3); break;
    default: launch_auv_tile<MODE, G, 4>(c, L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, o3, c3); break;
    }
}

template <int MODE>
static void launch_auv_mode(Ctx &c, const ItemListDev &L, const double *U, const double *V, int ld, double s1,
                            double s2, double *o1, double *o2, double *b1, double *b2, double *c1, double *c2,
                            double *o3 = nullptr, double *c3 = nullptr) {
    if (MODE == AUV_FROMZ) { launch_auv_tile<MODE, 4, 1>(c, L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2); return; }
    if (ld > 256) throw std::runtime_error("rank above 256 is not supported by the A(UV^T) kernel yet");
    // ld <= 4 (column-sharded factors): two lanes cover a row, so a warp works on 16 items at a time
    const int G = ld <= 4 ? 2 : (ld <= 32 ? 4 : (ld <= 64 ? 8 : (ld <= 128 ? 16 : 32)));
    const int np = (ld + 2 * G - 1) / (2 * G);
    switch (G) {
    case 2: launch_auv_np<MODE, 2>(c, np, L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, o3, c3); break;
    case 4: launch_auv_np<MODE, 4>(c, np, L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, o3, c3); break;
    case 8: launch_auv_np<MODE, 8>(c, np, L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, o3, c3); break;
    case 16: launch_auv_np<MODE, 16>(c, np, L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, o3, c3); break;
    default: launch_auv_np<MODE, 32>(c, np, L, U, V, ld, s1, s2, o1, o2, b1, b2, c1, c2, o3, c3); break;
    }
}

void launch_auv(Ctx &c, AuvMode mode, const ItemListDev &L, const double *U, const double *V, int ld, double scale1,
                double scale2, double *out1, double *out2, double *carry1, double *carry2, double *obj1, double *obj2,
                double *out3, double *carry3) {
    if (L.has_empty_rows) {
        LB2_CUDA(cudaMemsetAsync(out1, 0, sizeof(double) * L.n_rows, c.stream));
        if (mode == AUV_DUAL || mode == AUV_TRI) LB2_CUDA(cudaMemsetAsync(out2, 0, sizeof(double) * L.n_rows, c.stream));
        if (mode == AUV_TRI) LB2_CUDA(cudaMemsetAsync(out3, 0, sizeof(double) * L.n_rows, c.stream));
    }
    if (L.n_items == 0) return;
    switch (mode) {
    case AUV_SAME: launch_auv_mode<AUV_SAME>(c, L, U, V, ld, scale1, scale2, out1, out2, obj1, obj2, carry1, carry2); break;
    case AUV_PAIR: launch_auv_mode<AUV_PAIR>(c, L, U, V, ld, scale1, scale2, out1, out2, obj1, obj2, carry1, carry2); break;
    case AUV_DUAL: launch_auv_mode<AUV_DUAL>(c, L, U, V, ld, scale1, scale2, out1, out2, obj1, obj2, carry1, carry2); break;
    case AUV_TRI: launch_auv_mode<AUV_TRI>(c, L, U, V, ld, scale1, scale2, out1, out2, obj1, obj2, carry1, carry2, out3, carry3); break;
    case AUV_FROMZ: launch_auv_mode<AUV_FROMZ>(c, L, U, V, ld, scale1, scale2, out1, out2, obj1, obj2, carry1, carry2); break;
    }
}

__global__ void __launch_bounds__(kBlock) scatter_add_kernel(double *__restrict__ dst, const double *__restrict__ src,
                                                             const int *__restrict__ idx, long long n, double alpha,
                                                             bool accumulate, double *obj_dst) {
    for (long long a = blockIdx.x * (long long)kBlock + threadIdx.x; a < n; a += (long long)gridDim.x * kBlock) {
        const long long d = idx ? idx[a] : a;
        const double v = alpha * src[a];
        dst[d] = accumulate ? dst[d] + v : v;
    }
    if (obj_dst && blockIdx.x == 0 && threadIdx.x == 0) *obj_dst += alpha * src[n];
}

void launch_scatter_add(Ctx &c, double *dst, const double *src, const int *idx, long long n, double alpha,
                        bool accumulate, double *obj_dst) {
    scatter_add_kernel<<<grid_for(n, 4, c), kBlock, 0, c.stream>>>(dst, src, idx, n, alpha, accumulate, obj_dst);
    LB2_LAUNCH_CHECK(c);
}

// =================================================================================================
// S = C + A^*(w) on the pattern ; Y = S X
// =================================================================================================
__global__ void __launch_bounds__(kBlock) wsum_kernel(double *__restrict__ S, long long np,
                                                      const double *__restrict__ C_onP, const int *__restrict__ T_ptr,
                                                      const int *__restrict__ T_con, const double *__restrict__ T_val,
                                                      const double *__restrict__ w, const int *__restrict__ act_idx,
                                                      bool addC) {
    for (long long p = blockIdx.x * (long long)kBlock + threadIdx.x; p < np; p += (long long)gridDim.x * kBlock) {
        double s = addC ? C_onP[p] : 0.0;
        const int a = T_ptr[p], b = T_ptr[p + 1];
        for (int k = a; k < b; ++k) {
            const int cidx = act_idx ? act_idx[T_con[k]] : T_con[k];
            s += w[cidx] * T_val[k];
        }
        S[p] = s;
    }
}

void launch_wsum(Ctx &c, double *S, long long np, const double *C_onP, const int *T_ptr, const int *T_con,
                 const double *T_val, const double *w, const int *act_idx, bool w_is_compact, bool addC) {
    if (np == 0) return;
    wsum_kernel<<<grid_for(np, 2, c), kBlock, 0, c.stream>>>(S, np, C_onP, T_ptr, T_con, T_val, w,
                                                             w_is_compact ? nullptr : act_idx, addC);
    LB2_LAUNCH_CHECK(c);
}

// One group of G lanes per matrix row i: Y[i,:] = a * sum_e S[pos_e] X[col_e,:] + b Z[i,:].  Four neighbours are
// gathered per round with every 16-byte load issued before the first FMA (4*NP loads in flight per thread);
// the kernel is latency bound on the random row gathers, so memory-level parallelism is what matters.
template <int G, int NP>
__global__ void __launch_bounds__(kBlock) spmm_sym_kernel(long long n, int ld, const int *__restrict__ adj_ptr,
                                                          const int *__restrict__ adj_col,
                                                          const int *__restrict__ adj_pos, const double *__restrict__ S,
                                                          const double *__restrict__ X, double a, double b,
                                                          const double *__restrict__ Z, const double *__restrict__ Z2,
                                                          double *__restrict__ Y, ReduceScratch rs, double *red,
                                                          const double *__restrict__ cs, double c1) {
    constexpr int NG = kBlock / G;
    constexpr int UN = 4;
    const int g = threadIdx.x / G, gl = threadIdx.x % G;
    double ryy = 0.0, ryz = 0.0;
    for (long long i = (long long)blockIdx.x * NG + g; i < n; i += (long long)gridDim.x * NG) {
        double2 acc[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int col = 2 * gl + 2 * G * q;
            acc[q] = (cs && col < ld) ? make_double2(c1 * cs[col], c1 * cs[col + 1]) : make_double2(0.0, 0.0);
        }
        const int ea = adj_ptr[i], eb = adj_ptr[i + 1];
        for (int e = ea; e < eb; e += UN) {
            int jn[UN];
            double sv[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const bool ok = (e + u) < eb;
                jn[u] = ok ? adj_col[e + u] : (int)i;
                sv[u] = ok ? S[adj_pos[ok ? e + u : ea]] : 0.0;
            }
            double2 v[UN][NP];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const double *xr = X + (size_t)jn[u] * ld + 2 * gl;
#pragma unroll
                for (int q = 0; q < NP; ++q)
                    v[u][q] = ((2 * gl + 2 * G * q) < ld) ? ld2(xr + 2 * G * q) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    acc[q].x = fma(sv[u], v[u][q].x, acc[q].x);
                    acc[q].y = fma(sv[u], v[u][q].y, acc[q].y);
                }
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            const int col = 2 * gl + 2 * G * q;
            if (col < ld) {
                double2 y = make_double2(a * acc[q].x, a * acc[q].y);
                if (Z) {
                    double2 z = ld2(Z + (size_t)i * ld + col);
                    y.x = fma(b, z.x, y.x); y.y = fma(b, z.y, y.y);
                }
                st2(Y + (size_t)i * ld + col, y);
                ryy = fma(y.x, y.x, ryy); ryy = fma(y.y, y.y, ryy);
                if (Z2) {
                    double2 z2 = ld2(Z2 + (size_t)i * ld + col);
                    ryz = fma(y.x, z2.x, ryz); ryz = fma(y.y, z2.y, ryz);
                }
            }
        }
    }
    if (red) {
        double v[2] = {ryy, ryz};
        if (grid_reduce<2>(v, rs) && threadIdx.x == 0) { red[0] = v[0]; red[1] = v[1]; }
    }
}

template <int G>
static void launch_spmm_g(Ctx &c, int np, long long n, int ld, const int *ap, const int *ac, const int *apos,
                          const double *S, const double *X, double a, double b, const double *Z, const double *Z2,
                          double *Y, double *red, const double *cs, double c1) {
    constexpr int NG = kBlock / G;
    long long blocks = (n + NG - 1) / NG;      // one row per lane group; the hardware balances the CTAs
    if (blocks > 8192) blocks = 8192;           // capacity of the reduction scratch (grid-stride beyond that)
    const int grid = (int)blocks;
    switch (np) {
    case 1: spmm_sym_kernel<G, 1><<<grid, kBlock, 0, c.stream>>>(n, ld, ap, ac, apos, S, X, a, b, Z, Z2, Y, c.rs, red, cs, c1); break;
    case 2: spmm_sym_kernel<G, 2><<<grid, kBlock, 0, c.stream>>>(n, ld, ap, ac, apos, S, X, a, b, Z, Z2, Y, c.rs, red, cs, c1); break;
    case 3: spmm_sym_kernel<G, 3><<<grid, kBlock, 0, c.stream>>>(n, ld, ap, ac, apos, S, X, a, b, Z, Z2, Y, c.rs, red, cs, c1); break;
    default: spmm_sym_kernel<G, 4><<<grid, kBlock, 0, c.stream>>>(n, ld, ap, ac, apos, S, X, a, b, Z, Z2, Y, c.rs, red, cs, c1); break;
    }
    LB2_LAUNCH_CHECK(c);
}

void launch_spmm(Ctx &c, long long n, int ld, const int *adj_ptr, const int *adj_col, const int *adj_pos,
                 const double *S, const double *X, double a, double b, const double *Z, const double *Z2, double *Y,
                 double *red, const double *cs, double c1) {
    if (ld > 256) throw std::runtime_error("rank above 256 is not supported by the SpMM kernel yet");
    int G = ld <= 4 ? 2 : (ld <= 32 ? 4 : (ld <= 64 ? 8 : (ld <= 128 ? 16 : 32)));
    int np = (ld + 2 * G - 1) / (2 * G);
    switch (G) {
    case 2: launch_spmm_g<2>(c, np, n, ld, adj_ptr, adj_col, adj_pos, S, X, a, b, Z, Z2, Y, red, cs, c1); break;
    case 4: launch_spmm_g<4>(c, np, n, ld, adj_ptr, adj_col, adj_pos, S, X, a, b, Z, Z2, Y, red, cs, c1); break;
    case 8: launch_spmm_g<8>(c, np, n, ld, adj_ptr, adj_col, adj_pos, S, X, a, b, Z, Z2, Y, red, cs, c1); break;
    case 16: launch_spmm_g<16>(c, np, n, ld, adj_ptr, adj_col, adj_pos, S, X, a, b, Z, Z2, Y, red, cs, c1); break;
    default: launch_spmm_g<32>(c, np, n, ld, adj_ptr, adj_col, adj_pos, S, X, a, b, Z, Z2, Y, red, cs, c1); break;
    }
}

// =================================================================================================
// dense path (straightforward shared-memory tiled FP64 kernels; see DESIGN.md for the DMMA plan)
// =================================================================================================
constexpr int kDT = 32;   // tile edge

__device__ __forceinline__ long long packed_col_start(long long n, long long j) { return j * n - j * (j - 1) / 2; }

// ---- FP64 tensor-core (DMMA) versions -----------------------------------------------------------------
// mma.sync.aligned.m8n8k4.row.col.f64: A 8x4 (lane holds A[lane/4][lane%4]), B 4x8 (lane holds B[lane%4][lane/4]),
// C/D 8x8 (lane holds C[lane/4][2*(lane%4)] and the element right of it).
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct DenseObj {
    const double *Cp = nullptr;     // packed dense objective (nullptr: no objective)
    double s1 = 0.0, s2 = 0.0;      // scales applied to <C,Z1>, <C,Z2>
    double *obj1 = nullptr, *obj2 = nullptr;
    double *partials = nullptr;
    unsigned int *ticket = nullptr;
};

// One CTA per lower 32x32 tile of Z = sym(U V^T) (packed).  The four factor tiles are staged in shared memory,
// each warp owns two 8x8 sub-tiles and runs ld/4 DMMA k-steps per product; results are staged back through
// shared memory so that the packed columns are written with the row index fastest (coalesced).
template <bool SAME, bool DUAL>
__global__ void __launch_bounds__(kBlock) dense_uvt_dmma_kernel(long long n, int ld, const double *__restrict__ U,
                                                                const double *__restrict__ V, double *__restrict__ Z1,
                                                                double *__restrict__ Z2, DenseObj ob) {
    long long t = blockIdx.x;
    int bi = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) / 2.0);
    while ((long long)(bi + 1) * (bi + 2) / 2 <= t) ++bi;
    while ((long long)bi * (bi + 1) / 2 > t) --bi;
    const int bj = (int)(t - (long long)bi * (bi + 1) / 2);
    extern __shared__ double sm[];
    const int lds = ld + 1;
    double *Ui = sm, *Uj = Ui + kDT * lds, *Vi = Uj + kDT * lds, *Vj = Vi + kDT * lds;
    double *Zt1 = Vj + kDT * lds, *Zt2 = Zt1 + kDT * (kDT + 1);
    const long long i0 = (long long)bi * kDT, j0 = (long long)bj * kDT;
    for (int q = threadIdx.x; q < kDT * ld; q += kBlock) {
        const int rr = q / ld, cc = q % ld;
        const long long gi = i0 + rr, gj = j0 + rr;
        Ui[rr * lds + cc] = gi < n ? U[gi * ld + cc] : 0.0;
        Uj[rr * lds + cc] = gj < n ? U[gj * ld + cc] : 0.0;
        if (!SAME) {
            Vi[rr * lds + cc] = gi < n ? V[gi * ld + cc] : 0.0;
            Vj[rr * lds + cc] = gj < n ? V[gj * ld + cc] : 0.0;
        }
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int sub = 2 * w + h, si = sub >> 2, sj = sub & 3;
        double p1a = 0, p1b = 0, p2a = 0, p2b = 0, p3a = 0, p3b = 0;
        const double *ai = Ui + (8 * si + fr) * lds + fk, *bj_ = (SAME ? Uj : Vj) + (8 * sj + fr) * lds + fk;
        const double *avi = Vi + (8 * si + fr) * lds + fk, *buj = Uj + (8 * sj + fr) * lds + fk;
        for (int k = 0; k < ld; k += 4) {
            dmma884(p1a, p1b, ai[k], bj_[k]);                 // U_I V_J^T  (U_I U_J^T when SAME)
            if (!SAME) {
                dmma884(p2a, p2b, avi[k], buj[k]);            // V_I U_J^T
                if (DUAL) dmma884(p3a, p3b, avi[k], bj_[k]);  // V_I V_J^T
            }
        }
        const int row = 8 * si + fr, col = 8 * sj + 2 * fk;
        const bool d0 = (i0 + row == j0 + col), d1 = (i0 + row == j0 + col + 1);
        Zt1[row * (kDT + 1) + col] = (SAME || d0) ? p1a : (0.5 * p1a + 0.5 * p2a);
        Zt1[row * (kDT + 1) + col + 1] = (SAME || d1) ? p1b : (0.5 * p1b + 0.5 * p2b);
        if (DUAL) { Zt2[row * (kDT + 1) + col] = p3a; Zt2[row * (kDT + 1) + col + 1] = p3b; }
    }
    __syncthreads();
    const int li = threadIdx.x % kDT;
    double o1 = 0.0, o2 = 0.0;
    for (int q = 0; q < kDT * kDT / kBlock; ++q) {
        const int lj = threadIdx.x / kDT + q * (kBlock / kDT);
        const long long gi = i0 + li, gj = j0 + lj;
        if (gi >= n || gj >= n || gi < gj) continue;
        const long long pos = packed_col_start(n, gj) + (gi - gj);
        const double z1 = Zt1[li * (kDT + 1) + lj];
        Z1[pos] = z1;
        if (DUAL) Z2[pos] = Zt2[li * (kDT + 1) + lj];
        if (ob.Cp) {
            const double cw = (gi == gj) ? ob.Cp[pos] : 2.0 * ob.Cp[pos];
            o1 = fma(cw, z1, o1);
            if (DUAL) o2 = fma(cw, Zt2[li * (kDT + 1) + lj], o2);
        }
    }
    if (ob.Cp) {
        // <C, Z> fused here (objAUV for a dense objective): per-tile partials, last CTA adds them in tile order
        __shared__ double rsm[8];
        __shared__ bool last;
        o1 = block_sum(o1, rsm);
        if (DUAL) o2 = block_sum(o2, rsm);
        if (threadIdx.x == 0) {
            ob.partials[2 * (size_t)blockIdx.x] = o1;
            if (DUAL) ob.partials[2 * (size_t)blockIdx.x + 1] = o2;
        }
        __threadfence();
        if (threadIdx.x == 0) last = (atomicAdd(ob.ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (!last) return;
        __threadfence();
        double t1 = 0.0, t2 = 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += kBlock) {
            t1 += __ldcg(ob.partials + 2 * (size_t)b);
            if (DUAL) t2 += __ldcg(ob.partials + 2 * (size_t)b + 1);
        }
        t1 = block_sum(t1, rsm);
        if (DUAL) t2 = block_sum(t2, rsm);
        if (threadIdx.x == 0) {
            if (ob.obj1) *ob.obj1 += ob.s1 * t1;
            if (DUAL && ob.obj2) *ob.obj2 += ob.s2 * t2;
            *ob.ticket = 0u;
        }
    }
}

// Split-K symmetric product: CTA (x = 64-row block, y = split) accumulates its share of column blocks with DMMA
// (warp w owns rows 8w..8w+7 and every 8-column sub-tile of X) and writes a partial n x ldp block; the finish
// kernel adds the partials in split order and applies the epilogue (deterministic).
// Factors wider than 96 columns are covered by column panels of pw <= 96 columns (col0 = first column of the panel),
// one launch per panel; every output element still sums the same products in the same order.
constexpr int kSymRows = 64, kSymMaxSub = 12;   // <= 96 columns per panel
__global__ void __launch_bounds__(kBlock) dense_symm_dmma_kernel(long long n, int ld, int ldp, int col0, int pw,
                                                                 const double *__restrict__ Sp,
                                                                 const double *__restrict__ X, double *__restrict__ part,
                                                                 int jb_per_split) {
    extern __shared__ double sm[];
    double *St = sm;                                   // kSymRows x (kDT + 1)
    double *Xt = sm + kSymRows * (kDT + 1);            // kDT x (pw + 1)
    const int ldx = pw + 1, nsub = pw / 8;
    const long long i0 = (long long)blockIdx.x * kSymRows;
    const int nb = (int)((n + kDT - 1) / kDT);
    const int jb0 = blockIdx.y * jb_per_split, jb1 = min(nb, jb0 + jb_per_split);
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, fr = lane >> 2, fk = lane & 3;
    double acc[kSymMaxSub][2];
#pragma unroll
    for (int q = 0; q < kSymMaxSub; ++q) { acc[q][0] = 0.0; acc[q][1] = 0.0; }
    // The S tile of the NEXT column block is fetched into registers (8 independent loads per thread) while the
    // tensor-core loop works on the current one: one global round trip per tile instead of eight.
    constexpr int kPer = kSymRows * kDT / kBlock;
    double sreg[kPer];
    auto fetch = [&](int jb) {
        const long long j0 = (long long)jb * kDT;
        if (i0 >= j0 + kDT) {            // below the diagonal: S[i][j] = packed(col j)[i - j], row index fastest
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                const int q = threadIdx.x + u * kBlock, x = q % kSymRows, y = q / kSymRows;
                const long long gi = i0 + x, gj = j0 + y;
                sreg[u] = (gi < n && gj < n) ? __ldg(Sp + packed_col_start(n, gj) + (gi - gj)) : 0.0;
            }
        } else if (i0 + kSymRows <= j0) { // above: S[i][j] = packed(col i)[j - i], column index fastest
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                const int q = threadIdx.x + u * kBlock, x = q % kDT, y = q / kDT;
                const long long gi = i0 + y, gj = j0 + x;
                sreg[u] = (gi < n && gj < n) ? __ldg(Sp + packed_col_start(n, gi) + (gj - gi)) : 0.0;
            }
        } else {
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                const int q = threadIdx.x + u * kBlock, x = q % kSymRows, y = q / kSymRows;
                const long long gi = i0 + x, gj = j0 + y;
                double v = 0.0;
                if (gi < n && gj < n)
                    v = (gi >= gj) ? __ldg(Sp + packed_col_start(n, gj) + (gi - gj)) : __ldg(Sp + packed_col_start(n, gi) + (gj - gi));
                sreg[u] = v;
            }
        }
    };
    auto stash = [&](int jb) {
        const long long j0 = (long long)jb * kDT;
        const bool above = (i0 + kSymRows <= j0) && !(i0 >= j0 + kDT);
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int q = threadIdx.x + u * kBlock;
            if (above) St[(q / kDT) * (kDT + 1) + (q % kDT)] = sreg[u];
            else St[(q % kSymRows) * (kDT + 1) + (q / kSymRows)] = sreg[u];
        }
    };
    if (jb0 < jb1) fetch(jb0);
    for (int jb = jb0; jb < jb1; ++jb) {
        const long long j0 = (long long)jb * kDT;
        __syncthreads();
        stash(jb);
        for (int q = threadIdx.x; q < kDT * pw; q += kBlock) {
            const int rr = q / pw, cc = q % pw;
            Xt[rr * ldx + cc] = ((j0 + rr) < n && col0 + cc < ld) ? X[(j0 + rr) * ld + col0 + cc] : 0.0;
        }
        __syncthreads();
        if (jb + 1 < jb1) fetch(jb + 1);
#pragma unroll
        for (int ks = 0; ks < kDT / 4; ++ks) {
            const double a = St[(8 * w + fr) * (kDT + 1) + 4 * ks + fk];
            const double *xb = Xt + (4 * ks + fk) * ldx + fr;
#pragma unroll
            for (int q = 0; q < kSymMaxSub; ++q)
                if (q < nsub) dmma884(acc[q][0], acc[q][1], a, xb[8 * q]);
        }
    }
    const long long gi = i0 + 8 * w + fr;
    if (gi < n) {
        double *dst = part + ((size_t)blockIdx.y * n + gi) * ldp + col0 + 2 * fk;
#pragma unroll
        for (int q = 0; q < kSymMaxSub; ++q)
            if (q < nsub) { dst[8 * q] = acc[q][0]; dst[8 * q + 1] = acc[q][1]; }
    }
}

__global__ void __launch_bounds__(kBlock) dense_symm_finish_kernel(long long n, int r, int ld, int ldp, int nsplit,
                                                                   const double *__restrict__ part, double a, double b,
                                                                   const double *__restrict__ Z, const double *__restrict__ Z2,
                                                                   double *__restrict__ Y, ReduceScratch rs, double *red) {
    double ryy = 0.0, ryz = 0.0;
    const long long total = n * ld;
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < total; q += (long long)gridDim.x * kBlock) {
        const long long i = q / ld;
        const int col = (int)(q % ld);
        double acc = 0.0;
        for (int sidx = 0; sidx < nsplit; ++sidx) acc += part[((size_t)sidx * n + i) * ldp + col];
        double y = a * acc;
        if (Z) y = fma(b, Z[q], y);
        if (col >= r) y = 0.0;
        Y[q] = y;
        ryy = fma(y, y, ryy);
        if (Z2) ryz = fma(y, Z2[q], ryz);
    }
    if (red) {
        double v[2] = {ryy, ryz};
        if (grid_reduce<2>(v, rs) && threadIdx.x == 0) { red[0] = v[0]; red[1] = v[1]; }
    }
}

void launch_dense_uvt(Ctx &c, long long n, int r, int ld, const double *U, const double *V, double *Zp, bool same,
                      const double *Cp, double scale, double *obj) {
    DenseObj ob;
    if (Cp && obj) { ob.Cp = Cp; ob.s1 = scale; ob.obj1 = obj; ob.partials = c.dense_part; ob.ticket = c.ticket; }
    const int nb = (int)((n + kDT - 1) / kDT);
    const long long tiles = (long long)nb * (nb + 1) / 2;
    const size_t smem = sizeof(double) * (4 * kDT * (ld + 1) + 2 * kDT * (kDT + 1));
    if (same) {
        LB2_CUDA(cudaFuncSetAttribute(dense_uvt_dmma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_uvt_dmma_kernel<true, false><<<(unsigned)tiles, kBlock, smem, c.stream>>>(n, ld, U, V, Zp, nullptr, ob);
    } else {
        LB2_CUDA(cudaFuncSetAttribute(dense_uvt_dmma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_uvt_dmma_kernel<false, false><<<(unsigned)tiles, kBlock, smem, c.stream>>>(n, ld, U, V, Zp, nullptr, ob);
    }
    (void)r;
    LB2_LAUNCH_CHECK(c);
}

void launch_dense_uvt_dual(Ctx &c, long long n, int r, int ld, const double *R, const double *D, double *Z1, double *Z2,
                           const double *Cp, double s1, double s2, double *obj1, double *obj2) {
    DenseObj ob;
    if (Cp && (obj1 || obj2)) {
        ob.Cp = Cp; ob.s1 = s1; ob.s2 = s2; ob.obj1 = obj1; ob.obj2 = obj2; ob.partials = c.dense_part; ob.ticket = c.ticket;
    }
    const int nb = (int)((n + kDT - 1) / kDT);
    const long long tiles = (long long)nb * (nb + 1) / 2;
    const size_t smem = sizeof(double) * (4 * kDT * (ld + 1) + 2 * kDT * (kDT + 1));
    LB2_CUDA(cudaFuncSetAttribute(dense_uvt_dmma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_uvt_dmma_kernel<false, true><<<(unsigned)tiles, kBlock, smem, c.stream>>>(n, ld, R, D, Z1, Z2, ob);
    (void)r;
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) dense_wsum_kernel(double *__restrict__ Sp, const long long *__restrict__ D_pos,
                                                            long long n_pos, const int *__restrict__ T_ptr,
                                                            const int *__restrict__ T_con, const double *__restrict__ T_val,
                                                            const double *__restrict__ w, const int *__restrict__ act_idx) {
    for (long long u = blockIdx.x * (long long)kBlock + threadIdx.x; u < n_pos; u += (long long)gridDim.x * kBlock) {
        double s = Sp[D_pos[u]];
        for (int k = T_ptr[u]; k < T_ptr[u + 1]; ++k) {
            const int cidx = act_idx ? act_idx[T_con[k]] : T_con[k];
            s += w[cidx] * T_val[k];
        }
        Sp[D_pos[u]] = s;
    }
}

void launch_dense_wsum(Ctx &c, double *Sp, long long psize, const double *Cp, const long long *D_pos, long long n_pos,
                       const int *T_ptr, const int *T_con, const double *T_val, const double *w, const int *act_idx,
                       bool w_is_compact, bool addC) {
    if (addC) LB2_CUDA(cudaMemcpyAsync(Sp, Cp, sizeof(double) * psize, cudaMemcpyDeviceToDevice, c.stream));
    else LB2_CUDA(cudaMemsetAsync(Sp, 0, sizeof(double) * psize, c.stream));
    if (n_pos == 0) return;
    dense_wsum_kernel<<<grid_for(n_pos, 1, c), kBlock, 0, c.stream>>>(Sp, D_pos, n_pos, T_ptr, T_con, T_val, w,
                                                                      w_is_compact ? nullptr : act_idx);
    LB2_LAUNCH_CHECK(c);
}

// Y[I,:] = a * sum_J S[I,J] X[J,:] + b Z[I,:], S symmetric packed lower column-major. One CTA per row block I.
__global__ void __launch_bounds__(kBlock) dense_symm_kernel(long long n, int r, int ld, const double *__restrict__ Sp,
                                                            const double *__restrict__ X, double a, double b,
                                                            const double *__restrict__ Z, const double *__restrict__ Z2,
                                                            double *__restrict__ Y, ReduceScratch rs, double *red) {
    extern __shared__ double sm[];
    double *St = sm;                       // kDT x (kDT+1): St[li][lj] = S[i0+li][j0+lj]
    double *Xt = sm + kDT * (kDT + 1);     // kDT x ld
    const long long i0 = (long long)blockIdx.x * kDT;
    const int nb = (int)((n + kDT - 1) / kDT);
    // thread -> (row li, column chunk); kBlock/kDT = 8 threads per row, each owns columns cq, cq+8, ...
    const int li = threadIdx.x / 8, cq = threadIdx.x % 8;
    constexpr int MAXC = 32;               // supports ld <= 256
    double acc[MAXC];
#pragma unroll
    for (int q = 0; q < MAXC; ++q) acc[q] = 0.0;
    for (int bj = 0; bj < nb; ++bj) {
        const long long j0 = (long long)bj * kDT;
        __syncthreads();
        for (int q = threadIdx.x; q < kDT * kDT; q += kBlock) {
            int x = q % kDT, y = q / kDT;
            double v = 0.0;
            if (j0 < i0 || (j0 == i0)) {
                // lower (or diagonal) tile: rows i (fast index x), column j = j0 + y
                const long long gi = i0 + x, gj = j0 + y;
                if (gi < n && gj < n) {
                    if (gi >= gj) v = Sp[packed_col_start(n, gj) + (gi - gj)];
                    else v = Sp[packed_col_start(n, gi) + (gj - gi)];
                }
                St[x * (kDT + 1) + y] = v;
            } else {
                // upper tile: S[i][j] = lower(j, i): column i = i0 + y, rows j (fast index x)
                const long long gi = i0 + y, gj = j0 + x;
                if (gi < n && gj < n) v = Sp[packed_col_start(n, gi) + (gj - gi)];
                St[y * (kDT + 1) + x] = v;
            }
        }
        for (int q = threadIdx.x; q < kDT * ld; q += kBlock) {
            const int rr = q / ld, cc = q % ld;
            Xt[rr * ld + cc] = (j0 + rr) < n ? X[(j0 + rr) * ld + cc] : 0.0;
        }
        __syncthreads();
        for (int lj = 0; lj < kDT; ++lj) {
            const double s = St[li * (kDT + 1) + lj];
#pragma unroll
            for (int q = 0; q < MAXC; ++q) {
                const int col = cq + 8 * q;
                if (col < ld) acc[q] = fma(s, Xt[lj * ld + col], acc[q]);
            }
        }
    }
    double ryy = 0.0, ryz = 0.0;
    const long long gi = i0 + li;
    if (gi < n) {
#pragma unroll
        for (int q = 0; q < MAXC; ++q) {
            const int col = cq + 8 * q;
            if (col < ld) {
                double y = a * acc[q];
                if (Z) y = fma(b, Z[gi * ld + col], y);
                if (col >= r) y = 0.0;
                Y[gi * ld + col] = y;
                ryy = fma(y, y, ryy);
                if (Z2) ryz = fma(y, Z2[gi * ld + col], ryz);
            }
        }
    }
    if (red) {
        double v[2] = {ryy, ryz};
        if (grid_reduce<2>(v, rs) && threadIdx.x == 0) { red[0] = v[0]; red[1] = v[1]; }
    }
}

void launch_dense_symm(Ctx &c, long long n, int r, int ld, const double *Sp, const double *X, double a, double b,
                       const double *Z, const double *Z2, double *Y, double *red) {
    if (ld > 256) throw std::runtime_error("rank above 256 is not supported by the dense symm kernel yet");
    if (c.dense_part) {
        // FP64 tensor-core path: split-K grid of (64-row blocks) x (splits), partials added in a fixed order
        const int ldp = ((ld + 7) / 8) * 8;
        const int nb = (int)((n + kDT - 1) / kDT), nrb = (int)((n + kSymRows - 1) / kSymRows);
        int nsplit = (8 * c.num_sms + nrb - 1) / nrb;
        nsplit = std::max(1, std::min(nsplit, std::min(nb, 32)));
        while ((size_t)nsplit * n * ldp > c.dense_part_cap && nsplit > 1) --nsplit;
        if ((size_t)nsplit * n * ldp <= c.dense_part_cap) {
            const int jbps = (nb + nsplit - 1) / nsplit;
            nsplit = (nb + jbps - 1) / jbps;
            // (a register-tiled DFMA version of this kernel measured 136 us / 414 us at n = 3000, ld = 20 / 68 against
            //  90 us / 119 us for DMMA on the B200: the FP64 tensor pipe is the faster FP64 path on this part)
            // column panels of (almost) equal width, each at most 8 * kSymMaxSub columns
            const int npanel = (ldp + 8 * kSymMaxSub - 1) / (8 * kSymMaxSub);
            const int pw_max = ((ldp / 8 + npanel - 1) / npanel) * 8;
            const size_t smem = sizeof(double) * (kSymRows * (kDT + 1) + kDT * (pw_max + 1));
            LB2_CUDA(cudaFuncSetAttribute(dense_symm_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            for (int col0 = 0; col0 < ldp; col0 += pw_max) {
                const int pw = std::min(pw_max, ldp - col0);
                dense_symm_dmma_kernel<<<dim3(nrb, nsplit), kBlock, smem, c.stream>>>(n, ld, ldp, col0, pw, Sp, X, c.dense_part,
                                                                                       jbps);
                LB2_LAUNCH_CHECK(c);
            }
            dense_symm_finish_kernel<<<grid_for(n * ld, 1, c), kBlock, 0, c.stream>>>(n, r, ld, ldp, nsplit, c.dense_part, a, b, Z, Z2,
                                                                                   Y, c.rs, red);
            LB2_LAUNCH_CHECK(c);
            return;
        }
    }
    const int nb = (int)((n + kDT - 1) / kDT);
    const size_t smem = sizeof(double) * (kDT * (kDT + 1) + kDT * ld);
    LB2_CUDA(cudaFuncSetAttribute(dense_symm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_symm_kernel<<<nb, kBlock, smem, c.stream>>>(n, r, ld, Sp, X, a, b, Z, Z2, Y, c.rs, red);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) spmv_sym_kernel(long long n, const int *__restrict__ adj_ptr,
                                                          const int *__restrict__ adj_col, const int *__restrict__ adj_pos,
                                                          const double *__restrict__ S, const double *__restrict__ x,
                                                          double *__restrict__ y, double c1, const double *sum_slot) {
    const double shift = sum_slot ? c1 * (*sum_slot) : 0.0;
    for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        double acc = shift;
        for (int e = adj_ptr[i]; e < adj_ptr[i + 1]; ++e) acc = fma(S[adj_pos[e]], x[adj_col[e]], acc);
        y[i] = acc;
    }
}

void launch_spmv_sym(Ctx &c, long long n, const int *adj_ptr, const int *adj_col, const int *adj_pos, const double *S,
                     const double *x, double *y, double c1, const double *sum_slot) {
    spmv_sym_kernel<<<grid_for(n, 1, c), kBlock, 0, c.stream>>>(n, adj_ptr, adj_col, adj_pos, S, x, y, c1, sum_slot);
    LB2_LAUNCH_CHECK(c);
}

// one warp per row i of the packed symmetric matrix: columns j <= i are strided reads, rows j > i are contiguous
__global__ void __launch_bounds__(kBlock) dense_symv_kernel(long long n, const double *__restrict__ Sp,
                                                            const double *__restrict__ x, double *__restrict__ y) {
    const long long i = ((long long)blockIdx.x * kBlock + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    double acc = 0.0;
    for (long long j = lane; j < i; j += 32) acc = fma(Sp[packed_col_start(n, j) + (i - j)], x[j], acc);
    const long long base = packed_col_start(n, i);
    for (long long j = i + lane; j < n; j += 32) acc = fma(Sp[base + (j - i)], x[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[i] = acc;
}

void launch_dense_symv(Ctx &c, long long n, const double *Sp, const double *x, double *y) {
    const long long blocks = (n * 32 + kBlock - 1) / kBlock;
    dense_symv_kernel<<<(unsigned)blocks, kBlock, 0, c.stream>>>(n, Sp, x, y);
    LB2_LAUNCH_CHECK(c);
}

// =================================================================================================
// fused BLAS-1
// =================================================================================================
__device__ __forceinline__ double eval_coef(const Coef &k, double *S) {
    double v = k.k0 * S[k.i0];
    if (k.k1 != 0.0) v = fma(k.k1 * S[k.i1], S[k.i2], v);
    return v;
}

__global__ void scalar_prologue_kernel(Coef ca, Coef cb, double *S) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        if (ca.store >= 0) S[ca.store] = eval_coef(ca, S);
        if (cb.store >= 0) S[cb.store] = eval_coef(cb, S);
    }
}

__global__ void __launch_bounds__(kBlock) axpby_dot2_kernel(long long n, double *out, Coef ca, const double *x, Coef cb,
                                                            const double *y, const double *z, double *S, int dot_slot,
                                                            bool recip, ReduceScratch rs) {
    // Every block evaluates the coefficients from the scalar slots itself.  S[dot_slot] is rewritten only by the LAST
    // block to finish, i.e. after every block has read its coefficients, and a remembered coefficient goes to a slot
    // nobody reads in this kernel, so no separate one-thread prologue launch is needed.
    const double a = eval_coef(ca, S), b = eval_coef(cb, S);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (ca.store >= 0) S[ca.store] = a;
        if (cb.store >= 0) S[cb.store] = b;
    }
    double acc = 0.0;
    const long long n2 = n >> 1;
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < n2; q += (long long)gridDim.x * kBlock) {
        double2 xv = ld2(x + 2 * q);
        double2 o = make_double2(a * xv.x, a * xv.y);
        if (y) {
            double2 yv = ld2(y + 2 * q);
            o.x = fma(b, yv.x, o.x); o.y = fma(b, yv.y, o.y);
        }
        st2(out + 2 * q, o);
        if (z) {
            double2 zv = (z == out) ? o : ld2(z + 2 * q);
            acc = fma(o.x, zv.x, acc); acc = fma(o.y, zv.y, acc);
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double o = a * x[n - 1];
        if (y) o = fma(b, y[n - 1], o);
        out[n - 1] = o;
        if (z) acc = fma(o, (z == out) ? o : z[n - 1], acc);
    }
    if (z) {
        double v[1] = {acc};
        if (grid_reduce<1>(v, rs) && threadIdx.x == 0) S[dot_slot] = recip ? 1.0 / v[0] : v[0];
    }
}

void launch_axpby_dot(Ctx &c, long long n, double *out, Coef a, const double *x, Coef b, const double *y,
                      const double *z, double *S, int dot_slot, bool recip) {
    if ((a.store >= 0 && (a.store == dot_slot || a.store == a.i0 || a.store == a.i1 || a.store == a.i2 || a.store == b.i0 ||
                          a.store == b.i1 || a.store == b.i2)) ||
        (b.store >= 0 && (b.store == dot_slot || b.store == a.i0 || b.store == a.i1 || b.store == a.i2 || b.store == b.i0 ||
                          b.store == b.i1 || b.store == b.i2))) {
        // a remembered coefficient that aliases a slot read or written by the same launch would race: evaluate it in
        // a one-thread launch first (not used by the current call sites)
        scalar_prologue_kernel<<<1, 32, 0, c.stream>>>(a, b, S);
        LB2_LAUNCH_CHECK(c);
        if (a.store >= 0) a = coef_slot(a.store);
        if (b.store >= 0) b = coef_slot(b.store);
    }
    axpby_dot2_kernel<<<grid_for(n, 8, c), kBlock, 0, c.stream>>>(n, out, a, x, b, y, z, S, dot_slot, recip, c.rs);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) lbfgs_pair_kernel(long long n, bool update_y, double *__restrict__ ya,
                                                            const double *__restrict__ G,
                                                            const double *__restrict__ sa, const double *__restrict__ yb,
                                                            const double *__restrict__ sb, double *S, int d_slot,
                                                            int beta_slot, int yy_slot, bool finalize, ReduceScratch rs) {
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long n2 = n >> 1;
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < n2; q += (long long)gridDim.x * kBlock) {
        double2 y = ld2(ya + 2 * q);
        const double2 g = ld2(G + 2 * q), s = ld2(sa + 2 * q);
        if (update_y) {
            y.x += g.x; y.y += g.y;
            st2(ya + 2 * q, y);
        }
        v[0] = fma(y.x, s.x, v[0]); v[0] = fma(y.y, s.y, v[0]);
        v[1] = fma(y.x, y.x, v[1]); v[1] = fma(y.y, y.y, v[1]);
        v[4] = fma(s.x, g.x, v[4]); v[4] = fma(s.y, g.y, v[4]);
        v[6] = fma(y.x, g.x, v[6]); v[6] = fma(y.y, g.y, v[6]);
        if (yb) {
            const double2 y2 = ld2(yb + 2 * q), s2 = ld2(sb + 2 * q);
            v[2] = fma(y.x, y2.x, v[2]); v[2] = fma(y.y, y2.y, v[2]);
            v[3] = fma(y.x, s2.x, v[3]); v[3] = fma(y.y, s2.y, v[3]);
            v[5] = fma(s2.x, g.x, v[5]); v[5] = fma(s2.y, g.y, v[5]);
            v[7] = fma(y2.x, g.x, v[7]); v[7] = fma(y2.y, g.y, v[7]);
        }
    }
    if (grid_reduce<8>(v, rs) && threadIdx.x == 0) {
        for (int k = 0; k < 8; ++k) S[d_slot + k] = v[k];
        if (finalize) { S[beta_slot] = 1.0 / v[0]; S[yy_slot] = v[1]; }
    }
}

void launch_lbfgs_pair(Ctx &c, long long n, bool update_y, double *ya, const double *G, const double *sa, const double *yb,
                       const double *sb, double *S, int d_slot, int beta_slot, int yy_slot, bool finalize) {
    if (n & 1) throw std::runtime_error("factor vectors have even length by construction");
    lbfgs_pair_kernel<<<grid_for(n, 8, c), kBlock, 0, c.stream>>>(n, update_y, ya, G, sa, yb, sb, S, d_slot, beta_slot,
                                                                  yy_slot, finalize, c.rs);
    LB2_LAUNCH_CHECK(c);
}

__global__ void lbfgs_pair_finalize_kernel(double *S, int d_slot, int beta_slot, int yy_slot) {
    S[beta_slot] = 1.0 / S[d_slot];
    S[yy_slot] = S[d_slot + 1];
}

void launch_lbfgs_pair_finalize(Ctx &c, double *S, int d_slot, int beta_slot, int yy_slot) {
    lbfgs_pair_finalize_kernel<<<1, 1, 0, c.stream>>>(S, d_slot, beta_slot, yy_slot);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) lbfgs_dir_kernel(long long n, int depth, double *__restrict__ D,
                                                           const double *__restrict__ G, const double *__restrict__ ya,
                                                           const double *__restrict__ sa, const double *__restrict__ yb,
                                                           const double *__restrict__ sb, const double *__restrict__ S,
                                                           int d, int beta_a, int beta_b, int yy_b_slot,
                                                           const double *__restrict__ gg_slots, int n_gg) {
    // coefficients of the two-loop recursion in closed form (every thread evaluates the same ~30 flops)
    double ca = 0.0, cb = 0.0, wa = 0.0, wb = 0.0;      // D = -(G - ca*ya - cb*yb + wb*sb + wa*sa)
    bool use_grad = (depth == 0);
    if (depth >= 1) {
        const double ba = S[beta_a];
        const double sag = S[d + 4], yag = S[d + 6], yaya = S[d + 1];
        double gg = 0.0;
        for (int k = 0; k < n_gg; ++k) gg += gg_slots[2 * k];
        ca = ba * sag;                                              // alpha_a
        double dg;
        if (depth >= 2) {
            const double bb = S[beta_b];
            const double sbg = S[d + 5], ybg = S[d + 7], yayb = S[d + 2], yasb = S[d + 3], ybyb = S[yy_b_slot];
            cb = bb * (sbg - ca * yasb);                            // alpha_b = beta_b * sb.(g - ca*ya)
            wb = cb - bb * (ybg - ca * yayb - cb * ybyb);           // alpha_b - beta_b * yb.q2
            wa = ca - ba * (yag - ca * yaya - cb * yayb + wb * yasb);   // alpha_a - beta_a * ya.q3
            dg = -(gg - ca * yag - cb * ybg + wb * sbg + wa * sag);
        } else {
            wa = ca - ba * (yag - ca * yaya);
            dg = -(gg - ca * yag + wa * sag);
        }
        use_grad = (dg >= 0.0);                                     // LBFGSDirectionUseGrad: fall back to -G
    }
    const long long n2 = n >> 1;
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < n2; q += (long long)gridDim.x * kBlock) {
        const double2 g = ld2(G + 2 * q);
        double2 o = g;
        if (!use_grad) {
            const double2 y1 = ld2(ya + 2 * q), s1 = ld2(sa + 2 * q);
            o.x = fma(-ca, y1.x, o.x); o.y = fma(-ca, y1.y, o.y);
            if (depth >= 2) {
                const double2 y2 = ld2(yb + 2 * q), s2 = ld2(sb + 2 * q);
                o.x = fma(-cb, y2.x, o.x); o.y = fma(-cb, y2.y, o.y);
                o.x = fma(wb, s2.x, o.x); o.y = fma(wb, s2.y, o.y);
            }
            o.x = fma(wa, s1.x, o.x); o.y = fma(wa, s1.y, o.y);
        }
        st2(D + 2 * q, make_double2(-o.x, -o.y));
    }
}

void launch_lbfgs_dir(Ctx &c, long long n, int depth, double *D, const double *G, const double *ya, const double *sa,
                      const double *yb, const double *sb, const double *S, int d_slot, int beta_a, int beta_b,
                      int yy_b_slot, const double *gg_slots, int n_gg) {
    lbfgs_dir_kernel<<<grid_for(n, 8, c), kBlock, 0, c.stream>>>(n, depth, D, G, ya, sa, yb, sb, S, d_slot, beta_a, beta_b,
                                                                 yy_b_slot, gg_slots, n_gg);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) dot_kernel(long long n, const double *__restrict__ x,
                                                     const double *__restrict__ y, double *S, int slot, int mode,
                                                     ReduceScratch rs) {
    double acc = 0.0;   // mode 0: sum x*y, 1: sum |x|, 2: sum x
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < n; q += (long long)gridDim.x * kBlock)
        acc += mode == 0 ? x[q] * y[q] : (mode == 1 ? fabs(x[q]) : x[q]);
    double v[1] = {acc};
    if (grid_reduce<1>(v, rs) && threadIdx.x == 0) S[slot] = v[0];
}

void launch_dot(Ctx &c, long long n, const double *x, const double *y, double *S, int slot) {
    dot_kernel<<<grid_for(n, 8, c), kBlock, 0, c.stream>>>(n, x, y, S, slot, 0, c.rs);
    LB2_LAUNCH_CHECK(c);
}

void launch_asum(Ctx &c, long long n, const double *x, double *S, int slot) {
    dot_kernel<<<grid_for(n, 8, c), kBlock, 0, c.stream>>>(n, x, x, S, slot, 1, c.rs);
    LB2_LAUNCH_CHECK(c);
}

void launch_sum(Ctx &c, long long n, const double *x, double *S, int slot) {
    dot_kernel<<<grid_for(n, 8, c), kBlock, 0, c.stream>>>(n, x, x, S, slot, 2, c.rs);
    LB2_LAUNCH_CHECK(c);
}

// column sums of a row-major n x ld factor: each warp walks rows (lane = column, coalesced), block partials are
// combined in a fixed order by the last block
__global__ void __launch_bounds__(kBlock) colsum_kernel(long long n, int ld, const double *__restrict__ X,
                                                        double *__restrict__ out, double *__restrict__ scratch,
                                                        unsigned int *ticket) {
    __shared__ double sm[8][256];
    __shared__ bool s_last;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.0;
    for (long long i = (long long)blockIdx.x * 8 + w; i < n; i += (long long)gridDim.x * 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int col = lane + 32 * q;
            if (col < ld) acc[q] += X[i * ld + col];
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) sm[w][lane + 32 * q] = acc[q];
    __syncthreads();
    for (int col = threadIdx.x; col < ld; col += kBlock) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += sm[k][col];
        scratch[(size_t)blockIdx.x * ld + col] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int col = threadIdx.x; col < ld; col += kBlock) {
        double t = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) t += __ldcg(scratch + (size_t)b * ld + col);
        out[col] = t;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

void launch_colsum(Ctx &c, long long n, int ld, const double *X, double *out, double *scratch) {
    if (ld > 256) throw std::runtime_error("rank above 256 is not supported by the column-sum kernel yet");
    // 64 rows per CTA at least: the last CTA adds one partial per CTA and column serially, which dominated on small
    // blocks (n = 5000: 592 partials per column for 100 KB of data)
    long long blocks = (n + 63) / 64;
    const long long cap = (long long)c.num_sms * 4;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    colsum_kernel<<<(unsigned)blocks, kBlock, 0, c.stream>>>(n, ld, X, out, scratch, c.ticket);
    LB2_LAUNCH_CHECK(c);
}

__global__ void rank1_obj_kernel(int ld, const double *csA, const double *csB, double k1, double *obj1, double k2, double *obj2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double a = 0.0, b = 0.0;
    for (int k = 0; k < ld; ++k) { a = fma(csA[k], csB[k], a); b = fma(csB[k], csB[k], b); }
    if (obj1) *obj1 += k1 * a;
    if (obj2) *obj2 += k2 * b;
}

void launch_rank1_obj(Ctx &c, int ld, const double *csA, const double *csB, double k1, double *obj1, double k2, double *obj2) {
    rank1_obj_kernel<<<1, 32, 0, c.stream>>>(ld, csA, csB, k1, obj1, k2, obj2);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) neg_if_nonneg_kernel(long long n, double *__restrict__ D,
                                                               const double *__restrict__ G, const double *S, int slot) {
    if (!(S[slot] >= 0.0)) return;
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < n; q += (long long)gridDim.x * kBlock)
        D[q] = -G[q];
}

void launch_neg_if_nonneg(Ctx &c, long long n, double *D, const double *G, const double *S, int cond_slot) {
    neg_if_nonneg_kernel<<<grid_for(n, 4, c), kBlock, 0, c.stream>>>(n, D, G, S, cond_slot);
    LB2_LAUNCH_CHECK(c);
}

// m-vector update of the same point of the iteration (see launch_alm_m_update): carried by extra CTAs of the step
// kernel, so that the two element-wise passes cost one launch
struct MUpdate {
    long long m = 0;
    const double *q1 = nullptr, *q2 = nullptr, *lam = nullptr, *b = nullptr, *rho_p = nullptr;
    double *s = nullptr, *M1 = nullptr;
};

__device__ __forceinline__ void m_update_body(const MUpdate &u, const double *tau_p, long long first, long long stride) {
    const double tau = u.q1 ? *tau_p : 0.0, rho = *u.rho_p;
    const double t2 = tau * tau;
    for (long long i = first; i < u.m; i += stride) {
        double sv = u.s[i];
        if (u.q1) {
            sv = fma(tau, u.q1[i], sv);
            sv = fma(t2, u.q2[i], sv);
            u.s[i] = sv;
        }
        u.M1[i] = fma(rho, sv, fma(-rho, u.b[i], -u.lam[i]));
    }
}

__global__ void __launch_bounds__(kBlock) alm_step_kernel(long long n, const double *tau_p, const double *__restrict__ G,
                                                          const double *__restrict__ D, double *__restrict__ R,
                                                          double *__restrict__ y, double *__restrict__ s, int vec_blocks,
                                                          MUpdate mu) {
    if ((int)blockIdx.x >= vec_blocks) {
        m_update_body(mu, tau_p, (long long)(blockIdx.x - vec_blocks) * kBlock + threadIdx.x, (long long)(gridDim.x - vec_blocks) * kBlock);
        return;
    }
    const double tau = *tau_p;
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < n; q += (long long)vec_blocks * kBlock) {
        const double d = D[q];
        y[q] = -G[q];
        s[q] = tau * d;
        R[q] = fma(tau, d, R[q]);
    }
}

void launch_alm_step(Ctx &c, long long n, const double *tau_p, const double *G, const double *D, double *R, double *y, double *s) {
    const int vb = grid_for(n, 4, c);
    alm_step_kernel<<<vb, kBlock, 0, c.stream>>>(n, tau_p, G, D, R, y, s, vb, MUpdate{});
    LB2_LAUNCH_CHECK(c);
}

void launch_alm_step_m(Ctx &c, long long n, const double *tau_p, const double *G, const double *D, double *R, double *y, double *s,
                       long long m, const double *q1, const double *q2, double *sm, const double *lam, const double *b,
                       const double *rho_p, double *M1) {
    const int vb = grid_for(n, 4, c);
    const int mb = grid_for(m, 4, c);
    MUpdate mu;
    mu.m = m; mu.q1 = q1; mu.q2 = q2; mu.lam = lam; mu.b = b; mu.rho_p = rho_p; mu.s = sm; mu.M1 = M1;
    alm_step_kernel<<<vb + mb, kBlock, 0, c.stream>>>(n, tau_p, G, D, R, y, s, vb, mu);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) cg_update_kernel(long long n, double *__restrict__ x, double *__restrict__ r,
                                                           const double *__restrict__ p, const double *__restrict__ Q,
                                                           double *S, int slot_num, int slot_den, int slot_rr,
                                                           int slot_beta, ReduceScratch rs) {
    const double rr_old = S[slot_num];
    const double alpha = rr_old / S[slot_den];
    double acc = 0.0;
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < n; q += (long long)gridDim.x * kBlock) {
        x[q] = fma(alpha, p[q], x[q]);
        const double rv = fma(-alpha, Q[q], r[q]);
        r[q] = rv;
        acc = fma(rv, rv, acc);
    }
    double v[1] = {acc};
    if (grid_reduce<1>(v, rs) && threadIdx.x == 0) {
        S[slot_rr] = v[0];
        if (slot_beta >= 0) S[slot_beta] = v[0] / rr_old;      // beta of the next direction (lorads_cgs.c:226-228)
    }
}

void launch_cg_update(Ctx &c, long long n, double *x, double *r, const double *p, const double *Q, double *S,
                      int slot_num, int slot_den, int slot_rr, int slot_beta) {
    cg_update_kernel<<<grid_for(n, 8, c), kBlock, 0, c.stream>>>(n, x, r, p, Q, S, slot_num, slot_den, slot_rr, slot_beta, c.rs);
    LB2_LAUNCH_CHECK(c);
}

// Last block of the kernel that ends an iteration: publish the scalar slots to pinned host memory (Ctx::mirror), so that
// the host reads them right after the stream drains without a copy node in between.
__device__ __forceinline__ void mirror_slots(const double *S, double *mirror, int mirror_n) {
    if (!mirror) return;
    __syncthreads();
    for (int t = threadIdx.x; t < mirror_n; t += kBlock) mirror[t] = __ldcg(S + t);
}

__global__ void __launch_bounds__(kBlock) linesearch_dots_kernel(long long m, const double *__restrict__ b,
                                                                 const double *__restrict__ s,
                                                                 const double *__restrict__ lam, const double *rho_p,
                                                                 const double *__restrict__ q1,
                                                                 const double *__restrict__ q2, double *S, int slot,
                                                                 ReduceScratch rs, double *mirror, int mirror_n) {
    const double rhoInv = 1.0 / (*rho_p);
    double v[5] = {0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < m; i += (long long)gridDim.x * kBlock) {
        const double q0 = fma(rhoInv, lam[i], b[i] - s[i]);
        const double a = q1[i], c2 = q2[i];
        v[0] = fma(c2, c2, v[0]);
        v[1] = fma(a, c2, v[1]);
        v[2] = fma(q0, c2, v[2]);
        v[3] = fma(a, a, v[3]);
        v[4] = fma(q0, a, v[4]);
    }
    if (!grid_reduce<5>(v, rs)) return;
    if (threadIdx.x == 0)
        for (int k = 0; k < 5; ++k) S[slot + k] = v[k];
    mirror_slots(S, mirror, mirror_n);
}

// The same sums when the exact constraint values arrive in `src` (third output of the gather pass): s := src, the
// primal residual |b - s|^2 goes to S[pinf_slot], all in one pass.
__global__ void __launch_bounds__(kBlock) linesearch_resid_kernel(long long m, const double *__restrict__ b,
                                                                  const double *__restrict__ src, double *__restrict__ s,
                                                                  const double *__restrict__ lam, const double *rho_p,
                                                                  const double *__restrict__ q1,
                                                                  const double *__restrict__ q2, double *S, int slot,
                                                                  int pinf_slot, ReduceScratch rs, double *mirror,
                                                                  int mirror_n) {
    const double rhoInv = 1.0 / (*rho_p);
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < m; i += (long long)gridDim.x * kBlock) {
        const double sv = src[i];
        s[i] = sv;
        const double d = b[i] - sv;
        const double q0 = fma(rhoInv, lam[i], d);
        const double a = q1[i], c2 = q2[i];
        v[0] = fma(c2, c2, v[0]);
        v[1] = fma(a, c2, v[1]);
        v[2] = fma(q0, c2, v[2]);
        v[3] = fma(a, a, v[3]);
        v[4] = fma(q0, a, v[4]);
        v[5] = fma(d, d, v[5]);
    }
    if (!grid_reduce<6>(v, rs)) return;
    if (threadIdx.x == 0) {
        for (int k = 0; k < 5; ++k) S[slot + k] = v[k];
        S[pinf_slot] = v[5];
    }
    mirror_slots(S, mirror, mirror_n);
}

void launch_linesearch_resid(Ctx &c, long long m, const double *b, const double *src, double *s, const double *lam,
                             const double *rho_p, const double *q1, const double *q2, double *S, int slot, int pinf_slot) {
    linesearch_resid_kernel<<<grid_for(m, 4, c), kBlock, 0, c.stream>>>(m, b, src, s, lam, rho_p, q1, q2, S, slot, pinf_slot, c.rs, c.mirror, c.mirror_n);
    LB2_LAUNCH_CHECK(c);
}

void launch_linesearch_dots(Ctx &c, long long m, const double *b, const double *s, const double *lam, const double *rho_p,
                            const double *q1, const double *q2, double *S, int slot) {
    linesearch_dots_kernel<<<grid_for(m, 4, c), kBlock, 0, c.stream>>>(m, b, s, lam, rho_p, q1, q2, S, slot, c.rs, c.mirror, c.mirror_n);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) alm_m_update_kernel(long long m, const double *tau_p, const double *__restrict__ q1,
                                                              const double *__restrict__ q2, double *__restrict__ s,
                                                              const double *__restrict__ lam,
                                                              const double *__restrict__ b, const double *rho_p,
                                                              double *__restrict__ M1) {
    const double tau = q1 ? *tau_p : 0.0, rho = *rho_p;
    const double t2 = tau * tau;
    for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < m; i += (long long)gridDim.x * kBlock) {
        double sv = s[i];
        if (q1) {
            sv = fma(tau, q1[i], sv);
            sv = fma(t2, q2[i], sv);
            s[i] = sv;
        }
        M1[i] = fma(rho, sv, fma(-rho, b[i], -lam[i]));
    }
}

void launch_alm_m_update(Ctx &c, long long m, const double *tau_p, const double *q1, const double *q2, double *s,
                         const double *lam, const double *b, const double *rho_p, double *M1) {
    alm_m_update_kernel<<<grid_for(m, 4, c), kBlock, 0, c.stream>>>(m, tau_p, q1, q2, s, lam, b, rho_p, M1);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) resid_sq_kernel(long long m, const double *__restrict__ b,
                                                          const double *__restrict__ s, double *S, int slot,
                                                          ReduceScratch rs) {
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < m; i += (long long)gridDim.x * kBlock) {
        const double d = b[i] - s[i];
        acc = fma(d, d, acc);
    }
    double v[1] = {acc};
    if (grid_reduce<1>(v, rs) && threadIdx.x == 0) S[slot] = v[0];
}

void launch_resid_sq(Ctx &c, long long m, const double *b, const double *s, double *S, int slot) {
    resid_sq_kernel<<<grid_for(m, 4, c), kBlock, 0, c.stream>>>(m, b, s, S, slot, c.rs);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) dual_update_kernel(long long m, double rho, const double *__restrict__ b,
                                                             const double *__restrict__ s, double *__restrict__ lam) {
    for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < m; i += (long long)gridDim.x * kBlock)
        lam[i] = fma(-rho, s[i], fma(rho, b[i], lam[i]));
}

void launch_dual_update(Ctx &c, long long m, double rho, const double *b, const double *s, double *lam) {
    dual_update_kernel<<<grid_for(m, 4, c), kBlock, 0, c.stream>>>(m, rho, b, s, lam);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) admm_m1_kernel(long long m, const double *__restrict__ b,
                                                         const double *__restrict__ s, const double *__restrict__ cv,
                                                         const double *__restrict__ lam, double rho,
                                                         double *__restrict__ M1) {
    for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < m; i += (long long)gridDim.x * kBlock) {
        double t = s[i] - b[i];
        if (cv) t -= cv[i];
        M1[i] = fma(t, rho, -lam[i]);
    }
}

void launch_admm_m1(Ctx &c, long long m, const double *b, const double *s, const double *cv, const double *lam,
                    double rho, double *M1) {
    admm_m1_kernel<<<grid_for(m, 4, c), kBlock, 0, c.stream>>>(m, b, s, cv, lam, rho, M1);
    LB2_LAUNCH_CHECK(c);
}

__global__ void recip_kernel(double *S, int slot) { S[slot] = 1.0 / S[slot]; }

void launch_recip(Ctx &c, double *S, int slot) {
    recip_kernel<<<1, 1, 0, c.stream>>>(S, slot);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) scale_kernel(double *__restrict__ x, long long n, double f) {
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < n; q += (long long)gridDim.x * kBlock) x[q] *= f;
}

void launch_scale(Ctx &c, double *x, long long n, double f) {
    if (n <= 0) return;
    scale_kernel<<<grid_for(n, 4, c), kBlock, 0, c.stream>>>(x, n, f);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) relayout_kernel(long long n, int r_old, int ld_old, int ld_new,
                                                          const double *__restrict__ src, double *__restrict__ dst) {
    const long long total = n * r_old;
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < total; q += (long long)gridDim.x * kBlock) {
        const long long i = q / r_old;
        const int k = (int)(q % r_old);
        dst[i * ld_new + k] = src[i * ld_old + k];
    }
}

void launch_relayout(Ctx &c, long long n, int r_old, int ld_old, int ld_new, const double *src, double *dst) {
    if (r_old <= 0) return;
    relayout_kernel<<<grid_for(n * r_old, 2, c), kBlock, 0, c.stream>>>(n, r_old, ld_old, ld_new, src, dst);
    LB2_LAUNCH_CHECK(c);
}

// =================================================================================================
// all-reduce over NVLink peer memory
// =================================================================================================
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(kBlock) p2p_push_kernel(P2PDev P, const double *__restrict__ data, long long count) {
    __shared__ bool last;
    const unsigned long long e = *P.epoch + 1;
    const size_t base = ((size_t)(e & 1) * P.world + P.rank) * P.cap;
    // peer-major striping: consecutive blocks serve different peers so that all NVLink ports are busy at once
    const int nb = gridDim.x / P.world;                       // blocks per peer (grid is a multiple of world)
    const int peer = blockIdx.x % P.world, bid = blockIdx.x / P.world;
    double *dst = P.peer_x[peer] + base;
    if ((reinterpret_cast<unsigned long long>(data) & 15ull) == 0) {
        const long long n2 = count >> 1;
        for (long long q = (long long)bid * kBlock + threadIdx.x; q < n2; q += (long long)nb * kBlock)
            st2(dst + 2 * q, ld2(data + 2 * q));
        if ((count & 1) && bid == 0 && threadIdx.x == 0) dst[count - 1] = data[count - 1];
    } else {
        for (long long q = (long long)bid * kBlock + threadIdx.x; q < count; q += (long long)nb * kBlock) dst[q] = data[q];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x < P.world) st_release_sys(P.peer_f[threadIdx.x] + (size_t)(e & 1) * P.world + P.rank, e);
    if (threadIdx.x == 0) *P.ticket = 0u;
}

__global__ void __launch_bounds__(kBlock) p2p_reduce_kernel(P2PDev P, double *__restrict__ data, long long count) {
    __shared__ bool last;
    const unsigned long long e = *P.epoch + 1;
    const unsigned long long *fl = P.f + (size_t)(e & 1) * P.world;
    if (threadIdx.x < P.world) {
        const long long t0 = clock64();
        while (ld_acquire_sys(fl + threadIdx.x) < e) {
            if (clock64() - t0 > 20000000000LL) __trap();      // ~10 s: a peer is gone; fail loudly instead of hanging
        }
    }
    __syncthreads();
    const double *src = P.x + (size_t)(e & 1) * P.world * P.cap;
    for (long long q = blockIdx.x * (long long)kBlock + threadIdx.x; q < count; q += (long long)gridDim.x * kBlock) {
        double t = 0.0;
        for (int r = 0; r < P.world; ++r) t += __ldcg(src + (size_t)r * P.cap + q);
        data[q] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(P.ticket + 1, 1u) == gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) { P.ticket[1] = 0u; *P.epoch = e; }
}

// Small messages (scalars, dot tables): push, arrival wait and the rank-ordered sum in ONE single-CTA kernel, so the
// whole all-reduce costs one launch plus one NVLink round trip.
constexpr int kP2PSmall = 1024;
__global__ void __launch_bounds__(kBlock) p2p_small_kernel(P2PDev P, double *__restrict__ data, int count) {
    const unsigned long long e = *P.epoch + 1;
    const size_t set = (size_t)(e & 1) * P.world;
    for (int q = threadIdx.x; q < count; q += kBlock) {
        const double v = data[q];
        for (int r = 0; r < P.world; ++r) P.peer_x[r][(set + P.rank) * P.cap + q] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < P.world) {
        st_release_sys(P.peer_f[threadIdx.x] + set + P.rank, e);
        const long long t0 = clock64();
        while (ld_acquire_sys(P.f + set + threadIdx.x) < e) {
            if (clock64() - t0 > 20000000000LL) __trap();
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < count; q += kBlock) {
        double t = 0.0;
        for (int r = 0; r < P.world; ++r) t += __ldcg(P.x + (set + r) * P.cap + q);
        data[q] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) *P.epoch = e;
}

void launch_p2p_allreduce(Ctx &c, const P2PDev &P, double *data, long long count) {
    if ((size_t)count > P.cap) throw std::invalid_argument("p2p all-reduce: message larger than the exchange slot");
    if (count <= kP2PSmall) {
        p2p_small_kernel<<<1, kBlock, 0, c.stream>>>(P, data, (int)count);
        LB2_LAUNCH_CHECK(c);
        return;
    }
    int per_peer = (int)std::min<long long>(16, std::max<long long>(1, count / (2 * kBlock * 4)));
    p2p_push_kernel<<<per_peer * P.world, kBlock, 0, c.stream>>>(P, data, count);
    LB2_LAUNCH_CHECK(c);
    const int rgrid = (int)std::min<long long>(2LL * c.num_sms, std::max<long long>(1, (count + kBlock - 1) / kBlock));
    p2p_reduce_kernel<<<rgrid, kBlock, 0, c.stream>>>(P, data, count);
    LB2_LAUNCH_CHECK(c);
}

// =================================================================================================
// Lanczos helpers (dual infeasibility): classical Gram-Schmidt against the whole basis in one pass
// =================================================================================================
__global__ void __launch_bounds__(kBlock) lanczos_dots_kernel(long long n, const double *__restrict__ Vb, long long stride, int nv,
                                                              const double *__restrict__ w, double *h, double *scratch,
                                                              unsigned int *ticket, double *S, int alpha_slot) {
    __shared__ double sm[kBlock / 32][kLanczosMax];
    __shared__ bool last;
    double acc[kLanczosMax];
#pragma unroll
    for (int l = 0; l < kLanczosMax; ++l) acc[l] = 0.0;
    for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        const double wi = w[i];
#pragma unroll
        for (int l = 0; l < kLanczosMax; ++l)
            if (l < nv) acc[l] = fma(Vb[l * stride + i], wi, acc[l]);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int l = 0; l < kLanczosMax; ++l) {
        if (l < nv) {                               // warp-uniform
            const double t = warp_sum(acc[l]);
            if (lane == 0) sm[warp][l] = t;
        }
    }
    __syncthreads();
    if (threadIdx.x < nv) {
        double t = 0.0;
        for (int q = 0; q < kBlock / 32; ++q) t += sm[q][threadIdx.x];
        scratch[(size_t)blockIdx.x * kLanczosMax + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    // one warp per basis vector: lanes stride the per-block partials (independent loads in flight), fixed shuffle tree
    for (int l = warp; l < nv; l += kBlock / 32) {
        double t = 0.0;
        for (unsigned int b = lane; b < gridDim.x; b += 32) t += __ldcg(scratch + (size_t)b * kLanczosMax + l);
        t = warp_sum(t);
        if (lane == 0) {
            h[l] = t;
            if (alpha_slot >= 0 && l == nv - 1) S[alpha_slot] = t;
        }
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

void launch_lanczos_dots(Ctx &c, long long n, const double *Vb, long long stride, int nv, const double *w, double *h,
                         double *scratch, double *S, int alpha_slot) {
    if (nv < 1 || nv > kLanczosMax) throw std::invalid_argument("Lanczos basis size out of range");
    const int grid = std::min(grid_for(n, 2, c), 512);
    lanczos_dots_kernel<<<grid, kBlock, 0, c.stream>>>(n, Vb, stride, nv, w, h, scratch, c.ticket, S, alpha_slot);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) lanczos_update_kernel(long long n, const double *__restrict__ Vb, long long stride, int nv,
                                                                const double *__restrict__ h, double *__restrict__ w, double *S,
                                                                int wsq_slot, ReduceScratch rs) {
    __shared__ double hs[kLanczosMax];
    if (threadIdx.x < nv) hs[threadIdx.x] = h[threadIdx.x];
    __syncthreads();
    double t[1] = {0.0};
    for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) {
        double wi = w[i];
        for (int l = 0; l < nv; ++l) wi = fma(-hs[l], Vb[l * stride + i], wi);
        w[i] = wi;
        t[0] = fma(wi, wi, t[0]);
    }
    if (wsq_slot >= 0) {
        if (grid_reduce<1>(t, rs) && threadIdx.x == 0) S[wsq_slot] = t[0];
    }
}

void launch_lanczos_update(Ctx &c, long long n, const double *Vb, long long stride, int nv, const double *h, double *w,
                           double *S, int wsq_slot) {
    lanczos_update_kernel<<<grid_for(n, 2, c), kBlock, 0, c.stream>>>(n, Vb, stride, nv, h, w, S, wsq_slot, c.rs);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) lanczos_next_kernel(long long n, const double *__restrict__ w, double *__restrict__ out,
                                                              const double *S, int wsq_slot) {
    const double f = 1.0 / sqrt(S[wsq_slot]);
    for (long long i = blockIdx.x * (long long)kBlock + threadIdx.x; i < n; i += (long long)gridDim.x * kBlock) out[i] = f * w[i];
}

void launch_lanczos_next(Ctx &c, long long n, const double *w, double *out, const double *S, int wsq_slot) {
    lanczos_next_kernel<<<grid_for(n, 2, c), kBlock, 0, c.stream>>>(n, w, out, S, wsq_slot);
    LB2_LAUNCH_CHECK(c);
}

// =================================================================================================
// LP cone
// =================================================================================================
__global__ void __launch_bounds__(kBlock) lp_prod_kernel(long long n, const double *__restrict__ u, const double *__restrict__ v,
                                                         double *__restrict__ x) {
    for (long long j = blockIdx.x * (long long)kBlock + threadIdx.x; j < n; j += (long long)gridDim.x * kBlock) x[j] = u[j] * v[j];
}

void launch_lp_prod(Ctx &c, long long n, const double *u, const double *v, double *x) {
    lp_prod_kernel<<<grid_for(n, 1, c), kBlock, 0, c.stream>>>(n, u, v, x);
    LB2_LAUNCH_CHECK(c);
}

// one warp per constraint row; lanes stride the row's LP entries, fixed shuffle tree -> deterministic
__global__ void __launch_bounds__(kBlock) lp_rows_kernel(LpDev L, const double *__restrict__ u, const double *__restrict__ v,
                                                         double s1, double *__restrict__ out1, double s2,
                                                         double *__restrict__ out2) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (blockIdx.x * (long long)kBlock + threadIdx.x) >> 5, nw = ((long long)gridDim.x * kBlock) >> 5;
    for (long long i = w0; i < L.m; i += nw) {
        const int a = L.rbeg[i], b = L.rbeg[i + 1];
        if (a == b) continue;                       // warp-uniform
        double t1 = 0.0, t2 = 0.0;
        for (int k = a + lane; k < b; k += 32) {
            const int j = L.rcol[k];
            const double av = L.rval[k], vj = v ? v[j] : 1.0;
            t1 = fma(av, u[j] * vj, t1);
            if (out2) t2 = fma(av, vj * vj, t2);
        }
        t1 = warp_sum(t1);
        if (out2) t2 = warp_sum(t2);
        if (lane == 0) {
            out1[i] += s1 * t1;
            if (out2) out2[i] += s2 * t2;
        }
    }
}

void launch_lp_rows(Ctx &c, const LpDev &L, const double *u, const double *v, double s1, double *out1, double s2, double *out2) {
    lp_rows_kernel<<<grid_for(L.m * 32, 1, c), kBlock, 0, c.stream>>>(L, u, v, s1, out1, s2, out2);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) lp_obj_kernel(LpDev L, const double *__restrict__ u, const double *__restrict__ v, double s1,
                                                        double *obj1, double s2, double *obj2, ReduceScratch rs) {
    double t[2] = {0.0, 0.0};
    for (long long j = blockIdx.x * (long long)kBlock + threadIdx.x; j < L.n; j += (long long)gridDim.x * kBlock) {
        const double cj = L.c[j], vj = v[j];
        t[0] = fma(cj, u[j] * vj, t[0]);
        t[1] = fma(cj, vj * vj, t[1]);
    }
    if (grid_reduce<2>(t, rs) && threadIdx.x == 0) {
        if (obj1) *obj1 += s1 * t[0];
        if (obj2) *obj2 += s2 * t[1];
    }
}

void launch_lp_obj(Ctx &c, const LpDev &L, const double *u, const double *v, double s1, double *obj1, double s2, double *obj2) {
    lp_obj_kernel<<<grid_for(L.n, 4, c), kBlock, 0, c.stream>>>(L, u, v, s1, obj1, s2, obj2, c.rs);
    LB2_LAUNCH_CHECK(c);
}

__global__ void __launch_bounds__(kBlock) lp_grad_kernel(LpDev L, const double *__restrict__ w, const double *__restrict__ r,
                                                         double *__restrict__ g, double *red, ReduceScratch rs) {
    double t[1] = {0.0};
    for (long long j = blockIdx.x * (long long)kBlock + threadIdx.x; j < L.n; j += (long long)gridDim.x * kBlock) {
        double sum = L.c[j];
        for (int k = L.cbeg[j]; k < L.cbeg[j + 1]; ++k) sum += L.cval[k] * w[L.crow[k]];   // same order as LPdataMatSparseWeightSum
        const double gj = 2 * sum * r[j];
        g[j] = gj;
        t[0] = fma(gj, gj, t[0]);
    }
    if (grid_reduce<1>(t, rs) && threadIdx.x == 0 && red) red[0] = t[0];
}

void launch_lp_grad(Ctx &c, const LpDev &L, const double *w, const double *r, double *g, double *red) {
    lp_grad_kernel<<<grid_for(L.n, 1, c), kBlock, 0, c.stream>>>(L, w, r, g, red, c.rs);
    LB2_LAUNCH_CHECK(c);
}

// The reference walks the columns one after the other and each column reads constrValSum as left by its
// predecessors.  Columns that share no constraint row commute exactly, so the columns are grouped into levels
// (level(j) = 1 + max level of an earlier column sharing a row with j, built on the host); one CTA runs the
// levels in order, one warp per column inside a level.  Every column sees exactly the values it would see in
// the sequential sweep.
__device__ __forceinline__ void lp_update_one(const LpDev &L, int j, int lane, double rho, const double *b, const double *lam,
                                              double *cvs, double *x, double *upd, const double *noupd) {
    const int a = L.cbeg[j], e = L.cbeg[j + 1];
    const double xo = x[j], nu = noupd[j];
    double t = 0.0;
    for (int k = a + lane; k < e; k += 32) {
        const int i = L.crow[k];
        const double av = L.cval[k];
        const double m1 = ((cvs[i] - b[i]) - av * xo) * rho - lam[i];     // lorads_admm.c:603-616
        t = fma(av, m1, t);
    }
    t = warp_sum(t);
    t = __shfl_sync(0xffffffffu, t, 0);
    const double wsum = L.c[j] + t;
    double M2 = wsum * nu;
    M2 = M2 - rho * nu;
    const double blin = -1.0 * M2 / rho;
    const double un = blin / (1 + L.nrm2sq[j] * nu * nu);                 // lorads_admm.c:622-627
    const double xn = (upd == noupd) ? un * un : un * nu;
    for (int k = a + lane; k < e; k += 32) {
        const int i = L.crow[k];
        const double av = L.cval[k];
        cvs[i] = (cvs[i] - av * xo) + av * xn;                            // lorads_alg_common.c:235-237
    }
    __syncwarp();
    if (lane == 0) { upd[j] = un; x[j] = xn; }
    __syncwarp();
}

// One launch covers the levels [lv_lo, lv_hi).  A wide level is a launch of its own over many CTAs (its columns share no
// row, so any assignment of columns to warps gives the same values); runs of narrow levels share a one-CTA launch with a
// block barrier between levels.  The segments are built on the host (Solver::set_lp).
__global__ void __launch_bounds__(kBlock) lp_sweep_kernel(LpDev L, int lv_lo, int lv_hi, double rho,
                                                          const double *__restrict__ b, const double *__restrict__ lam,
                                                          double *cvs, double *x, double *u, double *v) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5), nw = gridDim.x * (kBlock / 32);
    for (int lv = lv_lo; lv < lv_hi; ++lv) {
        for (int q = L.lvl_ptr[lv] + w; q < L.lvl_ptr[lv + 1]; q += nw) {
            const int j = L.lvl_col[q];
            lp_update_one(L, j, lane, rho, b, lam, cvs, x, u, v);
            lp_update_one(L, j, lane, rho, b, lam, cvs, x, v, u);
        }
        if (lv + 1 < lv_hi) __syncthreads();      // more than one level only in one-CTA launches
    }
}

void launch_lp_sweep(Ctx &c, const LpDev &L, const LpSeg *segs, int n_segs, double rho, const double *b, const double *lam,
                     double *cvs, double *x, double *u, double *v) {
    for (int k = 0; k < n_segs; ++k) {
        const LpSeg &g = segs[k];
        if (g.blocks > 1 && g.lv_hi - g.lv_lo != 1) throw std::logic_error("LP sweep: a multi-CTA segment must be one level");
        lp_sweep_kernel<<<g.blocks, kBlock, 0, c.stream>>>(L, g.lv_lo, g.lv_hi, rho, b, lam, cvs, x, u, v);
        LB2_LAUNCH_CHECK(c);
    }
}

__global__ void __launch_bounds__(kBlock) lp_dinf_kernel(LpDev L, const double *__restrict__ w, double *S, int slot, ReduceScratch rs) {
    double t[1] = {0.0};
    for (long long j = blockIdx.x * (long long)kBlock + threadIdx.x; j < L.n; j += (long long)gridDim.x * kBlock) {
        double sum = L.c[j];
        for (int k = L.cbeg[j]; k < L.cbeg[j + 1]; ++k) sum += L.cval[k] * w[L.crow[k]];
        t[0] += fabs(fmin(sum, 0.0));
    }
    if (grid_reduce<1>(t, rs) && threadIdx.x == 0) S[slot] = t[0];
}

void launch_lp_dinf(Ctx &c, const LpDev &L, const double *w, double *S, int slot) {
    lp_dinf_kernel<<<grid_for(L.n, 1, c), kBlock, 0, c.stream>>>(L, w, S, slot, c.rs);
    LB2_LAUNCH_CHECK(c);
}

}  // namespace lb2
