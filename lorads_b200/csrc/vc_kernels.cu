// vc_kernels.cu -- vertex-centric gather kernels of the sparse-scratch path (sm_100a).
//
// Both hot gather operators walk the class-split adjacency of layout.hpp `VcLayout` row by row:
//   vc_spmm_kernel   Y = a (C + A^*(w)) X + b Z          replaces zeros + addObjCoeff + sdpDataWSum
//                    (lorads_sdp_conic.c:327,437,539,633; lorads_sdp_data.c:589-644) followed by mul_rk
//                    (dataMatSparseMultiRkMat, lorads_sdp_data.c:491-504) and the scal/axpy/nrm2 behind it
//                    (lorads_alm.c:34-37,51; lorads_admm.c:386-390).  S = C + sum w_i A_i is never written: the values of
//                    C are inline in the adjacency, singleton constraints contribute w[con]*a from their inline
//                    (constraint, value) pair, only multi-entry constraints go through a materialised remainder.
//   vc_auv_kernel    A(sym(U V^T)) and <C, sym(U V^T)>     replaces LORADSUVt (lorads_alg_common.c:21-68) +
//                    coneAUV/objAUV (lorads_sdp_conic.c:285-301,498-513; sparseAUV lorads_sdp_data.c:524-567).
//                    <C, sym(U W^T)> = sum_j U_j . (C W)_j : the objective needs ONE gathered row per adjacency entry
//                    and no per-entry reduction; singleton constraints take their entry once from the lower triangle.
//
// Mechanics (chosen from measurements, scripts/gather_probe.cu, profiles/r02_gather_probe*.log): a group of
// G = min(8, ceil(ld/4)) lanes owns a row, every lane moves 32 bytes per gathered row with one 256-bit load
// (LDG.E.ENL2.256), floor(32/G) rows per warp, rows taken in order of decreasing length so the groups of a warp
// finish together, the (neighbour, value) pairs are read coalesced one per lane and broadcast by shuffles.  While
// the factors fit in L2 one gather in flight per lane and many resident warps is fastest (UN = 1); beyond L2 four
// gathers in flight per lane hide the DRAM latency (UN = 4).  TMA staging of the rows (cp.async.bulk per row,
// tile::gather4) was measured 2.5-4x slower than these loads at 192-224 byte rows and is not used.
// All reductions have a fixed order (no floating-point atomics): results are bitwise reproducible.
#include "kernels.cuh"

#include <algorithm>
#include <cstdlib>
#include <stdexcept>

namespace lb2 {

namespace {

struct d4 { double a, b, c, d; };

__device__ __forceinline__ d4 ld4(const double *p) {
    d4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}
// plain (coherent) 256-bit load for operands that another kernel of the same graph wrote just before
__device__ __forceinline__ d4 ld4c(const double *p) {
    d4 r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}
__device__ __forceinline__ void st4(double *p, d4 v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory");
}
__device__ __forceinline__ double dot4(const d4 &x, const d4 &y) { return x.a * y.a + x.b * y.b + x.c * y.c + x.d * y.d; }
__device__ __forceinline__ void fma4(d4 &acc, double s, const d4 &v) {
    acc.a = fma(s, v.a, acc.a); acc.b = fma(s, v.b, acc.b); acc.c = fma(s, v.c, acc.c); acc.d = fma(s, v.d, acc.d);
}
constexpr d4 kZero4 = {0.0, 0.0, 0.0, 0.0};

// lane geometry of a warp: G lanes per row, RW = 32/G rows per warp
struct Lanes {
    int G, RW, grp, gl, base;
    bool valid;
    unsigned gmask;
    __device__ __forceinline__ explicit Lanes(int G_) : G(G_) {
        const int lane = threadIdx.x & 31;
        RW = 32 / G;
        grp = lane / G;
        gl = lane - grp * G;
        base = grp * G;
        valid = grp < RW;
        gmask = valid ? ((G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << base) : 0u;
    }
};

// sum over the G lanes of a group; the total is valid in lane gl == 0 (fixed tree order)
__device__ __forceinline__ double group_total(const Lanes &L, double t) {
    for (int o = 1; o < L.G; o <<= 1) {
        const int src = L.base + ((L.gl + o) < L.G ? L.gl + o : L.gl);
        const double u = __shfl_sync(L.gmask, t, src);
        if (L.gl + o < L.G && (L.gl & (2 * o - 1)) == 0) t += u;
    }
    return t;
}

// acc[q] += sum_{e in [ea,eb)} val(e) * X[col[e], 4*gl + 32*q ..]   (one adjacency segment of a row)
// PL = entries every lane fetches per chunk.  With G >= 3 lanes per row one entry per lane (PL = 1) gives chunks of G
// entries; with one or two lanes per row (ld <= 8: low rank, or the column-sharded slices of a factor) the chunk would
// be a single entry or a pair and the loop would be a chain of dependent loads, so each lane fetches PL = 4 / 2 entries
// and all the row gathers of the chunk are issued before the first multiply-add.
template <int NP, int UN, int PL, class ValF>
__device__ __forceinline__ void gather_segment(const Lanes &L, int ea, int eb, const int *__restrict__ col, ValF valf,
                                               const double *__restrict__ X, int ld, d4 (&acc)[NP]) {
    if constexpr (PL > 1) {
        // PL = 4 is launched with G = 1 (ld <= 4), PL = 2 with G = 2 (ld <= 8): four entries per chunk either way
        static_assert(NP == 1, "several entries per lane only for rows of at most 8 columns");
        constexpr int GS = 4 / PL;
        for (int e = ea; e < eb; e += 4) {
            int jc[PL];
            double sc[PL];
#pragma unroll
            for (int k = 0; k < PL; ++k) {
                const int idx = e + k * GS + L.gl;
                const bool ok = idx < eb;
                jc[k] = ok ? col[idx] : 0;
                sc[k] = ok ? valf(idx) : 0.0;
            }
            const int cnt = min(4, eb - e);
            d4 v[PL][GS];
            double sv[PL][GS];
#pragma unroll
            for (int k = 0; k < PL; ++k)
#pragma unroll
                for (int h = 0; h < GS; ++h) {
                    const bool live = (k * GS + h) < cnt;
                    const int j = (GS == 1) ? jc[k] : __shfl_sync(L.gmask, jc[k], L.base + h);
                    const double s = (GS == 1) ? sc[k] : __shfl_sync(L.gmask, sc[k], L.base + h);
                    sv[k][h] = live ? s : 0.0;
                    v[k][h] = (live && 4 * L.gl < ld) ? ld4(X + (size_t)j * ld + 4 * L.gl) : kZero4;
                }
#pragma unroll
            for (int k = 0; k < PL; ++k)
#pragma unroll
                for (int h = 0; h < GS; ++h) fma4(acc[0], sv[k][h], v[k][h]);
        }
        return;
    }
    for (int e = ea; e < eb; e += L.G) {
        const bool ok = e + L.gl < eb;
        const int jc = ok ? col[e + L.gl] : 0;
        const double sc = ok ? valf(e + L.gl) : 0.0;
        const int cnt = min(L.G, eb - e);
#pragma unroll 1
        for (int h = 0; h < cnt; h += UN) {
            d4 v[UN][NP];
            double sv[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int hh = min(h + u, L.G - 1);
                const bool live = (h + u) < cnt;
                const int j = __shfl_sync(L.gmask, jc, L.base + hh);
                const double s = __shfl_sync(L.gmask, sc, L.base + hh);
                sv[u] = live ? s : 0.0;
                const double *xr = X + (size_t)j * ld + 4 * L.gl;
#pragma unroll
                for (int q = 0; q < NP; ++q) v[u][q] = (live && (4 * L.gl + 32 * q) < ld) ? ld4(xr + 32 * q) : kZero4;
            }
#pragma unroll
            for (int u = 0; u < UN; ++u)
#pragma unroll
                for (int q = 0; q < NP; ++q) fma4(acc[q], sv[u], v[u][q]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
//  Y = a * (useC*C + sum_con w[con] A_con) X + b Z
// ---------------------------------------------------------------------------------------------------------------
template <int NP, int UN, int MINB, int PL>
__global__ void __launch_bounds__(kBlock, MINB)
    vc_spmm_kernel(VcDev V, int ld, int G, bool useC, const double *__restrict__ w, const int *__restrict__ wmap,
                   const double *__restrict__ Sres, const double *__restrict__ X, double a, double b,
                   const double *__restrict__ Z, const double *__restrict__ Z2, double *__restrict__ Y, ReduceScratch rs,
                   double *red, const double *__restrict__ cs, double c1) {
    const Lanes L(G);
    const long long wg = (long long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (kBlock / 32);
    double ryy = 0.0, ryz = 0.0;
    if (L.valid)
        for (long long slot = wg * L.RW + L.grp; slot < V.n; slot += nw * L.RW) {
            const int i = V.order[slot];
            d4 acc[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int col = 4 * L.gl + 32 * q;
                acc[q] = (cs && col < ld) ? d4{c1 * cs[col], c1 * cs[col + 1], c1 * cs[col + 2], c1 * cs[col + 3]} : kZero4;
            }
            {
                const bool dynamic = (w != nullptr) || (Sres != nullptr);
                const int ea = dynamic ? V.u_ptr[i] : V.u_mid[i];
                const int eb = useC ? V.u_ptr[i + 1] : V.u_mid[i];
                const int *tg = V.u_tag;
                const double *uv = V.u_val;
                gather_segment<NP, UN, PL>(L, ea, eb, V.u_col,
                                       [tg, uv, w, wmap, Sres](int e) {
                                           const int t = tg[e];
                                           if (t == -1) return uv[e];
                                           if (t >= 0) return w ? w[wmap ? wmap[t] : t] * uv[e] : 0.0;
                                           return Sres ? Sres[-2 - t] : 0.0;
                                       },
                                       X, ld, acc);
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int col = 4 * L.gl + 32 * q;
                if (col < ld) {
                    d4 y = {a * acc[q].a, a * acc[q].b, a * acc[q].c, a * acc[q].d};
                    if (Z) {
                        const d4 z = ld4c(Z + (size_t)i * ld + col);
                        y.a = fma(b, z.a, y.a); y.b = fma(b, z.b, y.b); y.c = fma(b, z.c, y.c); y.d = fma(b, z.d, y.d);
                    }
                    st4(Y + (size_t)i * ld + col, y);
                    ryy = fma(y.a, y.a, ryy); ryy = fma(y.b, y.b, ryy); ryy = fma(y.c, y.c, ryy); ryy = fma(y.d, y.d, ryy);
                    if (Z2) {
                        const d4 z2 = ld4c(Z2 + (size_t)i * ld + col);
                        ryz = fma(y.a, z2.a, ryz); ryz = fma(y.b, z2.b, ryz); ryz = fma(y.c, z2.c, ryz); ryz = fma(y.d, z2.d, ryz);
                    }
                }
            }
        }
    if (red) {
        double v[2] = {ryy, ryz};
        if (grid_reduce<2>(v, rs) && threadIdx.x == 0) { red[0] = v[0]; red[1] = v[1]; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
//  A(sym(U W^T)) of the singleton constraints + <C, sym(U W^T)>
// ---------------------------------------------------------------------------------------------------------------
// MODE (kernels.cuh AuvMode): SAME  out1 = s1 A(UU^T)
//                             PAIR  out1 = s1 A(sym(U W^T))
//                             DUAL  out1 = s1 A(sym(U W^T)), out2 = s2 A(W W^T)
//                             TRI   DUAL + out3 = A(U U^T)
// The objective row (index V.obj_row) of out1 / out2 receives s1 <C, sym(U W^T)> / s2 <C, W W^T>, which are also
// added to *obj1 / *obj2, as the item kernel does.
template <int MODE, int NP, int UN, int MINB, int PL>
__global__ void __launch_bounds__(kBlock, MINB)
    vc_auv_kernel(VcDev V, int ld, int G, bool with_obj, const double *__restrict__ U, const double *__restrict__ W, double s1,
                  double s2, double *__restrict__ out1, double *__restrict__ out2, double *__restrict__ out3, double *obj1,
                  double *obj2, ReduceScratch rs) {
    constexpr bool SAME = (MODE == AUV_SAME);
    constexpr bool TRI = (MODE == AUV_TRI);
    constexpr bool DUAL = (MODE == AUV_DUAL) || TRI;
    const Lanes L(G);
    const long long wg = (long long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (kBlock / 32);
    const double *__restrict__ W2 = SAME ? U : W;       // the gathered side of the objective
    double p1 = 0.0, p2 = 0.0;
    if (L.valid)
        for (long long slot = wg * L.RW + L.grp; slot < V.n; slot += nw * L.RW) {
            const int j = V.order_l[slot];
            const int ca = with_obj ? V.u_mid[j] : 0, cb = with_obj ? V.u_ptr[j + 1] : 0;
            const int dc = V.d_con ? V.d_con[j] : -1;
            if (ca == cb && dc < 0) continue;
            d4 uj[NP], wj[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int col = 4 * L.gl + 32 * q;
                uj[q] = (col < ld) ? ld4(U + (size_t)j * ld + col) : kZero4;
                wj[q] = SAME ? uj[q] : ((col < ld) ? ld4(W + (size_t)j * ld + col) : kZero4);
            }
            if (cb > ca) {
                d4 t[NP];
#pragma unroll
                for (int q = 0; q < NP; ++q) t[q] = kZero4;
                const double *uv = V.u_val;
                gather_segment<NP, UN, PL>(L, ca, cb, V.u_col, [uv](int e) { return uv[e]; }, W2, ld, t);
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    p1 += dot4(uj[q], t[q]);
                    if constexpr (DUAL) p2 += dot4(wj[q], t[q]);
                }
            }
            if (dc >= 0) {
                // diagonal singleton constraint of this column: everything comes from the owner rows
                const double cf = V.d_coef[j];
                double a1 = 0.0, a3 = 0.0, a4 = 0.0;
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    a1 += dot4(uj[q], wj[q]);
                    if constexpr (DUAL) a3 += dot4(wj[q], wj[q]);
                    if constexpr (TRI) a4 += dot4(uj[q], uj[q]);
                }
                a1 = group_total(L, a1);
                if constexpr (DUAL) a3 = group_total(L, a3);
                if constexpr (TRI) a4 = group_total(L, a4);
                if (L.gl == 0) {
                    out1[dc] = s1 * cf * a1;
                    if constexpr (DUAL) out2[dc] = s2 * cf * a3;
                    if constexpr (TRI) out3[dc] = cf * a4;
                }
            }
        }
    if (with_obj) {
        double v[2] = {p1, p2};
        if (grid_reduce<2>(v, rs) && threadIdx.x == 0) {
            const double o1 = s1 * v[0], o2 = s2 * v[1];
            out1[V.obj_row] = o1;
            if (obj1) *obj1 += o1;
            if constexpr (DUAL) {
                out2[V.obj_row] = o2;
                if (obj2) *obj2 += o2;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
//  narrow factors (ld <= 8: low rank, or the column shards of a factor): entry-parallel lanes
// ---------------------------------------------------------------------------------------------------------------
// A whole row is one or two 32-byte loads, so a lane per ROW (the geometry above with G = 1, 2) leaves every lane
// chasing its own index stream with uncoalesced loads.  Here eight lanes share a row instead: lane gl takes entries
// gl, gl + 8, ... (the eight index loads of a trip are one coalesced access), gathers the WHOLE row of its entry and
// keeps a partial accumulator of its own; the eight partials are added once per row (SpMM) or never (the objective of
// A(UV^T) is linear in them).  NQ = ld / 4.
template <int NQ>
__global__ void __launch_bounds__(kBlock, 4)
    vc_spmm_narrow_kernel(VcDev V, int ld, bool useC, const double *__restrict__ w, const int *__restrict__ wmap,
                          const double *__restrict__ Sres, const double *__restrict__ X, double a, double b,
                          const double *__restrict__ Z, const double *__restrict__ Z2, double *__restrict__ Y, ReduceScratch rs,
                          double *red, const double *__restrict__ cs, double c1) {
    const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
    const long long wg = (long long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (kBlock / 32);
    const bool dynamic = (w != nullptr) || (Sres != nullptr);
    double ryy = 0.0, ryz = 0.0;
    for (long long base = wg * 4; base < V.n; base += nw * 4) {        // warp-uniform trip count: full-mask shuffles below
        const long long slot = base + g;
        const int i = slot < V.n ? V.order[slot] : -1;
        d4 acc[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[q] = kZero4;
        if (i >= 0) {
            const int ea = dynamic ? V.u_ptr[i] : V.u_mid[i];
            const int eb = useC ? V.u_ptr[i + 1] : V.u_mid[i];
            for (int e = ea + gl; e < eb; e += 8) {
                const int j = V.u_col[e], t = V.u_tag[e];
                double sv = V.u_val[e];
                if (t >= 0) sv = w ? w[wmap ? wmap[t] : t] * sv : 0.0;
                else if (t <= -2) sv = Sres ? Sres[-2 - t] : 0.0;
#pragma unroll
                for (int q = 0; q < NQ; ++q) fma4(acc[q], sv, ld4(X + (size_t)j * ld + 4 * q));
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1)
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                acc[q].a += __shfl_xor_sync(0xffffffffu, acc[q].a, o, 8);
                acc[q].b += __shfl_xor_sync(0xffffffffu, acc[q].b, o, 8);
                acc[q].c += __shfl_xor_sync(0xffffffffu, acc[q].c, o, 8);
                acc[q].d += __shfl_xor_sync(0xffffffffu, acc[q].d, o, 8);
            }
        if (i >= 0 && gl < NQ) {          // lane q of the group finishes columns 4q .. 4q+3
            const int col = 4 * gl;
            d4 r = acc[0];
#pragma unroll
            for (int q = 1; q < NQ; ++q) if (gl == q) r = acc[q];
            if (cs) { r.a += c1 * cs[col]; r.b += c1 * cs[col + 1]; r.c += c1 * cs[col + 2]; r.d += c1 * cs[col + 3]; }
            d4 y = {a * r.a, a * r.b, a * r.c, a * r.d};
            if (Z) {
                const d4 z = ld4c(Z + (size_t)i * ld + col);
                y.a = fma(b, z.a, y.a); y.b = fma(b, z.b, y.b); y.c = fma(b, z.c, y.c); y.d = fma(b, z.d, y.d);
            }
            st4(Y + (size_t)i * ld + col, y);
            ryy = fma(y.a, y.a, ryy); ryy = fma(y.b, y.b, ryy); ryy = fma(y.c, y.c, ryy); ryy = fma(y.d, y.d, ryy);
            if (Z2) {
                const d4 z2 = ld4c(Z2 + (size_t)i * ld + col);
                ryz = fma(y.a, z2.a, ryz); ryz = fma(y.b, z2.b, ryz); ryz = fma(y.c, z2.c, ryz); ryz = fma(y.d, z2.d, ryz);
            }
        }
    }
    if (red) {
        double v[2] = {ryy, ryz};
        if (grid_reduce<2>(v, rs) && threadIdx.x == 0) { red[0] = v[0]; red[1] = v[1]; }
    }
}

template <int MODE, int NQ>
__global__ void __launch_bounds__(kBlock, 4)
    vc_auv_narrow_kernel(VcDev V, int ld, bool with_obj, const double *__restrict__ U, const double *__restrict__ W, double s1,
                         double s2, double *__restrict__ out1, double *__restrict__ out2, double *__restrict__ out3, double *obj1,
                         double *obj2, ReduceScratch rs) {
    constexpr bool SAME = (MODE == AUV_SAME);
    constexpr bool TRI = (MODE == AUV_TRI);
    constexpr bool DUAL = (MODE == AUV_DUAL) || TRI;
    const int lane = threadIdx.x & 31, g = lane >> 3, gl = lane & 7;
    const long long wg = (long long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (kBlock / 32);
    const double *__restrict__ W2 = SAME ? U : W;
    double p1 = 0.0, p2 = 0.0;
    for (long long slot = wg * 4 + g; slot < V.n; slot += nw * 4) {
        const int j = V.order_l[slot];
        const int ca = with_obj ? V.u_mid[j] : 0, cb = with_obj ? V.u_ptr[j + 1] : 0;
        const int dc = V.d_con ? V.d_con[j] : -1;
        if (ca == cb && dc < 0) continue;
        d4 uj[NQ], wj[NQ], t[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            uj[q] = ld4(U + (size_t)j * ld + 4 * q);
            wj[q] = SAME ? uj[q] : ld4(W + (size_t)j * ld + 4 * q);
            t[q] = kZero4;
        }
        for (int e = ca + gl; e < cb; e += 8) {
            const int i = V.u_col[e];
            const double cv = V.u_val[e];
#pragma unroll
            for (int q = 0; q < NQ; ++q) fma4(t[q], cv, ld4(W2 + (size_t)i * ld + 4 * q));
        }
        // <U_j, sum_lanes t> = sum_lanes <U_j, t_lane>: no reduction across the lanes of the row
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            p1 += dot4(uj[q], t[q]);
            if constexpr (DUAL) p2 += dot4(wj[q], t[q]);
        }
        if (dc >= 0 && gl == 0) {
            double a1 = 0.0, a3 = 0.0, a4 = 0.0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                a1 += dot4(uj[q], wj[q]);
                if constexpr (DUAL) a3 += dot4(wj[q], wj[q]);
                if constexpr (TRI) a4 += dot4(uj[q], uj[q]);
            }
            const double cf = V.d_coef[j];
            out1[dc] = s1 * cf * a1;
            if constexpr (DUAL) out2[dc] = s2 * cf * a3;
            if constexpr (TRI) out3[dc] = cf * a4;
        }
    }
    if (with_obj) {
        double v[2] = {p1, p2};
        if (grid_reduce<2>(v, rs) && threadIdx.x == 0) {
            const double o1 = s1 * v[0], o2 = s2 * v[1];
            out1[V.obj_row] = o1;
            if (obj1) *obj1 += o1;
            if constexpr (DUAL) {
                out2[V.obj_row] = o2;
                if (obj2) *obj2 += o2;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
//  off-diagonal singleton constraints: one output per lower-triangular entry, entry-parallel
// ---------------------------------------------------------------------------------------------------------------
// A lane group walks CH consecutive entries of lowA (sorted by column, so the column-side rows U_j, W_j stay in
// registers while the column repeats), gathers the row-side rows, takes the full dot products and writes the value to
// the entry's constraint row -- every singleton row has exactly one writer, no reduction across entries.
template <int MODE, int NP>
__global__ void __launch_bounds__(kBlock, 4)
    vc_low_kernel(VcDev V, int ld, int G, const double *__restrict__ U, const double *__restrict__ W, double s1, double s2,
                  double *__restrict__ out1, double *__restrict__ out2, double *__restrict__ out3) {
    constexpr bool SAME = (MODE == AUV_SAME);
    constexpr bool TRI = (MODE == AUV_TRI);
    constexpr bool DUAL = (MODE == AUV_DUAL) || TRI;
    constexpr int CH = 8;
    const Lanes L(G);
    if (!L.valid) return;
    const long long groups = (long long)gridDim.x * (kBlock / 32) * L.RW;
    const long long g0 = ((long long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5)) * L.RW + L.grp;
    for (long long base_e = V.l_lo + g0 * CH; base_e < V.l_hi; base_e += groups * CH) {
        const int cnt = (int)min((long long)CH, V.l_hi - base_e);
        int pj = -1;
        d4 uj[NP], wj[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) uj[q] = wj[q] = kZero4;
#pragma unroll 1
        for (int t = 0; t < cnt; ++t) {
            // one entry per trip (two per trip measured slower: 305 vs 278 us on the 2e6 entries of matrix completion)
            const long long e = base_e + t;
            const int i = V.l_row[e], j = V.l_col[e], con = V.l_con[e];
            const double cf = V.l_coef[e];
            d4 ui[NP], wi[NP];
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const int col = 4 * L.gl + 32 * q;
                const bool in = col < ld;
                ui[q] = in ? ld4(U + (size_t)i * ld + col) : kZero4;
                if constexpr (!SAME) wi[q] = in ? ld4(W + (size_t)i * ld + col) : kZero4;
                if (j != pj) {
                    uj[q] = in ? ld4(U + (size_t)j * ld + col) : kZero4;
                    if constexpr (!SAME) wj[q] = in ? ld4(W + (size_t)j * ld + col) : kZero4;
                }
            }
            pj = j;
            double a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                if constexpr (SAME) {
                    a1 += dot4(ui[q], uj[q]);
                } else {
                    a1 += dot4(ui[q], wj[q]);                           // U_i . W_j
                    a2 += dot4(uj[q], wi[q]);                           // U_j . W_i
                    if constexpr (DUAL) a3 += dot4(wi[q], wj[q]);       // W_i . W_j
                    if constexpr (TRI) a4 += dot4(ui[q], uj[q]);        // U_i . U_j
                }
            }
            a1 = group_total(L, a1);
            if constexpr (!SAME) a2 = group_total(L, a2);
            if constexpr (DUAL) a3 = group_total(L, a3);
            if constexpr (TRI) a4 = group_total(L, a4);
            if (L.gl == 0) {
                if constexpr (SAME) {
                    out1[con] = s1 * cf * a1;
                } else {
                    out1[con] = s1 * cf * ((i == j) ? a1 : (0.5 * a1 + 0.5 * a2));
                    if constexpr (DUAL) out2[con] = s2 * cf * a3;
                    if constexpr (TRI) out3[con] = cf * a4;
                }
            }
        }
    }
}

constexpr int kVcMaxGrid = 8192;     // capacity of the reduction scratch (rows beyond it are taken grid-stride)

inline int vc_lanes(int ld) { return ld >= 32 ? 8 : (ld + 3) / 4; }
inline int vc_passes(int ld) {
    const int need = (ld + 31) / 32;
    for (int np : {1, 2, 3, 4, 6, 8})
        if (np >= need) return np;
    return 0;
}
// persistent grid: `per_sm` resident CTAs per SM walk the rows grid-stride, so the grid reduction (one fence and one
// ticket per CTA) is paid once per resident CTA instead of once per 40 rows
inline int vc_grid(const Ctx &c, long long n, int G, int per_sm) {
    const long long rows_per_block = (long long)(32 / G) * (kBlock / 32);
    long long g = (n + rows_per_block - 1) / rows_per_block;
    g = std::min<long long>(g, (long long)c.num_sms * per_sm);
    g = std::min<long long>(g, kVcMaxGrid);
    return (int)std::max<long long>(g, 1);
}

// Resident CTAs per SM asked of the compiler for factors wider than 32 columns (NP >= 2 passes): 0 = the budget of the
// one-pass kernels (48 / 64 registers, accumulators partly in local memory), 3 and 2 = 80 / 128 registers.  Measured at
// n = 1e5 (scripts/wide_minb.py, profiles/r02_wide_factor_register_budget.log; us hot, settings 0 / 3 / 2):
//   ld  64: TRI  99 /  99 /  94   SpMM  81 / 101 /  96
//   ld  96: TRI 174 / 187 / 163   SpMM 139 / 176 / 126
//   ld 128: TRI 263 / 285 / 312   SpMM 219 / 158 / 199
// The defaults below take the fastest of each row; LORADS_B200_VC_WIDE_MINB (read at every launch, so that one process
// can compare the settings) overrides them.  A() values and the SpMM are bit-identical across the settings; the
// objective sum follows the persistent grid and moves in the last bit.
static int vc_wide_minb(int dflt) {
    const char *e = getenv("LORADS_B200_VC_WIDE_MINB");
    return e ? atoi(e) : dflt;
}

template <int NP, int UN, int MINB = (UN == 1 ? 5 : 3), int PL = 1>
void spmm_launch(Ctx &c, const VcDev &V, int ld, bool useC, const double *w, const int *wmap, const double *Sres, const double *X,
                 double a, double b, const double *Z, const double *Z2, double *Y, double *red, const double *cs, double c1) {
    const int G = vc_lanes(ld);
    vc_spmm_kernel<NP, UN, MINB, PL><<<vc_grid(c, V.n, G, MINB), kBlock, 0, c.stream>>>(V, ld, G, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, c.rs, red, cs, c1);
}

template <int MODE, int NP, int UN, int MINB = (UN == 1 ? 4 : 3), int PL = 1>
void auv_launch(Ctx &c, const VcDev &V, int ld, bool with_obj, const double *U, const double *W, double s1, double s2, double *o1,
                double *o2, double *o3, double *obj1, double *obj2) {
    const int G = vc_lanes(ld);
    vc_auv_kernel<MODE, NP, UN, MINB, PL><<<vc_grid(c, V.n, G, MINB), kBlock, 0, c.stream>>>(V, ld, G, with_obj, U, W, s1, s2, o1, o2, o3, obj1, obj2, c.rs);
}

template <int MODE, int NP>
void low_launch(Ctx &c, const VcDev &V, int ld, const double *U, const double *W, double s1, double s2, double *o1, double *o2,
                double *o3) {
    const int G = vc_lanes(ld);
    const long long groups_needed = (V.l_hi - V.l_lo + 7) / 8;
    const long long per_block = (long long)(32 / G) * (kBlock / 32);
    long long g = (groups_needed + per_block - 1) / per_block;
    g = std::min<long long>(g, (long long)c.num_sms * 8);
    vc_low_kernel<MODE, NP><<<(int)std::max<long long>(g, 1), kBlock, 0, c.stream>>>(V, ld, G, U, W, s1, s2, o1, o2, o3);
}

template <int MODE>
void low_dispatch(Ctx &c, int np, const VcDev &V, int ld, const double *U, const double *W, double s1, double s2, double *o1,
                  double *o2, double *o3) {
    switch (np) {
    case 1: low_launch<MODE, 1>(c, V, ld, U, W, s1, s2, o1, o2, o3); break;
    case 2: low_launch<MODE, 2>(c, V, ld, U, W, s1, s2, o1, o2, o3); break;
    case 3: low_launch<MODE, 3>(c, V, ld, U, W, s1, s2, o1, o2, o3); break;
    case 4: low_launch<MODE, 4>(c, V, ld, U, W, s1, s2, o1, o2, o3); break;
    case 6: low_launch<MODE, 6>(c, V, ld, U, W, s1, s2, o1, o2, o3); break;
    default: low_launch<MODE, 8>(c, V, ld, U, W, s1, s2, o1, o2, o3); break;
    }
}

template <int MODE>
void auv_dispatch(Ctx &c, int np, int un, const VcDev &V, int ld, bool with_obj, const double *U, const double *W, double s1,
                  double s2, double *o1, double *o2, double *o3, double *obj1, double *obj2) {
#define LB2_VC_AUV(NP_, UN_) auv_launch<MODE, NP_, UN_>(c, V, ld, with_obj, U, W, s1, s2, o1, o2, o3, obj1, obj2)
    switch (np) {
    case 1: {
        static const int minb = getenv("LORADS_B200_VC_MINB_AUV") ? atoi(getenv("LORADS_B200_VC_MINB_AUV")) : 0;
        if (ld <= 4) auv_launch<MODE, 1, 1, 3, 4>(c, V, ld, with_obj, U, W, s1, s2, o1, o2, o3, obj1, obj2);
        else if (ld <= 8) auv_launch<MODE, 1, 1, 3, 2>(c, V, ld, with_obj, U, W, s1, s2, o1, o2, o3, obj1, obj2);
        else if (un >= 4 && minb == 4) auv_launch<MODE, 1, 4, 4>(c, V, ld, with_obj, U, W, s1, s2, o1, o2, o3, obj1, obj2);
        else if (un >= 4) LB2_VC_AUV(1, 4);
        else if (un == 2) LB2_VC_AUV(1, 2);
        else if (minb == 5) auv_launch<MODE, 1, 1, 5>(c, V, ld, with_obj, U, W, s1, s2, o1, o2, o3, obj1, obj2);
        else if (minb == 6) auv_launch<MODE, 1, 1, 6>(c, V, ld, with_obj, U, W, s1, s2, o1, o2, o3, obj1, obj2);
        else LB2_VC_AUV(1, 1);
        break;
    }
#define LB2_VC_AUV_WIDE(NP_)                                                                                                    \
    if (wide == 2) auv_launch<MODE, NP_, 1, 2>(c, V, ld, with_obj, U, W, s1, s2, o1, o2, o3, obj1, obj2);                         \
    else if (wide == 3) auv_launch<MODE, NP_, 1, 3>(c, V, ld, with_obj, U, W, s1, s2, o1, o2, o3, obj1, obj2);                    \
    else LB2_VC_AUV(NP_, 1);
    case 2: { const int wide = vc_wide_minb(2); LB2_VC_AUV_WIDE(2) break; }
    case 3: { const int wide = vc_wide_minb(2); LB2_VC_AUV_WIDE(3) break; }
    case 4: { const int wide = vc_wide_minb(0); LB2_VC_AUV_WIDE(4) break; }
    case 6: LB2_VC_AUV(6, 1); break;
    default: LB2_VC_AUV(8, 1); break;
    }
#undef LB2_VC_AUV_WIDE
#undef LB2_VC_AUV
}

}  // namespace

int vc_max_ld() { return 256; }

// loads in flight per lane: one gather per lane and as many resident warps as fit measured fastest from n = 1e5
// (L2 resident) to n = 1e6 (SpMM 624 vs 752 us with four in flight: the extra registers cost more warps than the
// unrolling gains); four only when the block is so small that a row group per row leaves most warp slots empty
// (n = 5000: 1000 warps on 148 SMs) and the parallelism has to come from inside the row.  LORADS_B200_VC_UN overrides.
static int vc_unroll(long long n, int ld) {
    (void)ld;
    static const int forced = getenv("LORADS_B200_VC_UN") ? atoi(getenv("LORADS_B200_VC_UN")) : 0;
    if (forced > 0) return forced;
    return n < 30000 ? 4 : 1;
}

void launch_vc_spmm(Ctx &c, const VcDev &V, int ld, bool useC, const double *w, const int *wmap, const double *Sres,
                    const double *X, double a, double b, const double *Z, const double *Z2, double *Y, double *red,
                    const double *cs, double c1) {
    const int np = vc_passes(ld);
    if (np == 0) throw std::runtime_error("rank above 256 is not supported by the SpMM kernel");
    const int un = vc_unroll(V.n, ld);
    static const bool narrow_on = getenv("LORADS_B200_NO_NARROW") == nullptr;
    // measured at n = 1e5 (profiles/r02_kernels_narrow_factors_cfg2.log): ld = 4: 20.5 vs 28.7 us; ld = 8: 28.1 vs 25.7 us
    // (n = 1e6, index arrays streamed from DRAM, rows in natural order: it loses, 170 vs 134 us -- lane-per-row path there)
    if (ld <= 4 && narrow_on && V.n <= 250000) {
        const long long blocks = std::min<long long>((V.n + 31) / 32, (long long)c.num_sms * 4);
        const int grid = (int)std::max<long long>(1, std::min<long long>(blocks, kVcMaxGrid));
        if (ld <= 4) vc_spmm_narrow_kernel<1><<<grid, kBlock, 0, c.stream>>>(V, ld, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, c.rs, red, cs, c1);
        else vc_spmm_narrow_kernel<2><<<grid, kBlock, 0, c.stream>>>(V, ld, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, c.rs, red, cs, c1);
        c.launches++;
        LB2_CUDA(cudaGetLastError());
        return;
    }
#define LB2_VC_SPMM(NP_, UN_) spmm_launch<NP_, UN_>(c, V, ld, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, red, cs, c1)
    switch (np) {
    case 1: {
        static const int minb = getenv("LORADS_B200_VC_MINB") ? atoi(getenv("LORADS_B200_VC_MINB")) : 0;
        if (ld <= 4) spmm_launch<1, 1, 4, 4>(c, V, ld, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, red, cs, c1);
        else if (ld <= 8) spmm_launch<1, 1, 4, 2>(c, V, ld, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, red, cs, c1);
        else if (un >= 4 && minb == 4) spmm_launch<1, 4, 4>(c, V, ld, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, red, cs, c1);
        else if (un >= 4) LB2_VC_SPMM(1, 4);
        else if (un == 2) LB2_VC_SPMM(1, 2);
        else if (minb == 6) spmm_launch<1, 1, 6>(c, V, ld, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, red, cs, c1);
        else if (minb == 8) spmm_launch<1, 1, 8>(c, V, ld, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, red, cs, c1);
        else LB2_VC_SPMM(1, 1);
        break;
    }
#define LB2_VC_SPMM_WIDE(NP_)                                                                                                   \
    if (wide == 2) spmm_launch<NP_, 1, 2>(c, V, ld, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, red, cs, c1);                          \
    else if (wide == 3) spmm_launch<NP_, 1, 3>(c, V, ld, useC, w, wmap, Sres, X, a, b, Z, Z2, Y, red, cs, c1);                     \
    else LB2_VC_SPMM(NP_, 1);
    case 2: { const int wide = vc_wide_minb(0); LB2_VC_SPMM_WIDE(2) break; }
    case 3: { const int wide = vc_wide_minb(2); LB2_VC_SPMM_WIDE(3) break; }
    case 4: { const int wide = vc_wide_minb(3); LB2_VC_SPMM_WIDE(4) break; }
    case 6: LB2_VC_SPMM(6, 1); break;
    default: LB2_VC_SPMM(8, 1); break;
    }
#undef LB2_VC_SPMM_WIDE
#undef LB2_VC_SPMM
    c.launches++;
    LB2_CUDA(cudaGetLastError());
}

void launch_vc_auv(Ctx &c, AuvMode mode, const VcDev &V, int ld, bool with_obj, const double *U, const double *W, double s1,
                   double s2, double *out1, double *out2, double *out3, double *obj1, double *obj2) {
    const int np = vc_passes(ld);
    if (np == 0) throw std::runtime_error("rank above 256 is not supported by the A(UV^T) kernel");
    const int un = vc_unroll(V.n, ld);
    // row-centric part: objective row and the diagonal singleton constraints (skipped when the cone has neither)
    static const bool narrow_on = getenv("LORADS_B200_NO_NARROW") == nullptr;
    // (ld = 4 with the objective row: 18.5 vs 22.6 us; constraints only, or ld = 8: the lane-per-row path stays faster)
    if (V.n > 0 && with_obj && ld <= 4 && narrow_on && V.n <= 250000) {
        const long long blocks = std::min<long long>((V.n + 31) / 32, (long long)c.num_sms * 4);
        const int grid = (int)std::max<long long>(1, std::min<long long>(blocks, kVcMaxGrid));
#define LB2_VC_AUV_NARROW(M_)                                                                                                   \
        if (ld <= 4) vc_auv_narrow_kernel<M_, 1><<<grid, kBlock, 0, c.stream>>>(V, ld, with_obj, U, W, s1, s2, out1, out2, out3, obj1, obj2, c.rs); \
        else vc_auv_narrow_kernel<M_, 2><<<grid, kBlock, 0, c.stream>>>(V, ld, with_obj, U, W, s1, s2, out1, out2, out3, obj1, obj2, c.rs);
        switch (mode) {
        case AUV_SAME: LB2_VC_AUV_NARROW(AUV_SAME) break;
        case AUV_PAIR: LB2_VC_AUV_NARROW(AUV_PAIR) break;
        case AUV_DUAL: LB2_VC_AUV_NARROW(AUV_DUAL) break;
        case AUV_TRI: LB2_VC_AUV_NARROW(AUV_TRI) break;
        default: throw std::invalid_argument("vertex-centric A(UV^T): unsupported mode");
        }
#undef LB2_VC_AUV_NARROW
        c.launches++;
        LB2_CUDA(cudaGetLastError());
    } else if (V.n > 0 && (with_obj || V.d_con)) {
        switch (mode) {
        case AUV_SAME: auv_dispatch<AUV_SAME>(c, np, un, V, ld, with_obj, U, W, s1, s2, out1, out2, out3, obj1, obj2); break;
        case AUV_PAIR: auv_dispatch<AUV_PAIR>(c, np, un, V, ld, with_obj, U, W, s1, s2, out1, out2, out3, obj1, obj2); break;
        case AUV_DUAL: auv_dispatch<AUV_DUAL>(c, np, un, V, ld, with_obj, U, W, s1, s2, out1, out2, out3, obj1, obj2); break;
        case AUV_TRI: auv_dispatch<AUV_TRI>(c, np, un, V, ld, with_obj, U, W, s1, s2, out1, out2, out3, obj1, obj2); break;
        default: throw std::invalid_argument("vertex-centric A(UV^T): unsupported mode");
        }
        c.launches++;
        LB2_CUDA(cudaGetLastError());
    }
    // entry-parallel part: the off-diagonal singleton constraints
    if (V.l_hi > V.l_lo) {
        switch (mode) {
        case AUV_SAME: low_dispatch<AUV_SAME>(c, np, V, ld, U, W, s1, s2, out1, out2, out3); break;
        case AUV_PAIR: low_dispatch<AUV_PAIR>(c, np, V, ld, U, W, s1, s2, out1, out2, out3); break;
        case AUV_DUAL: low_dispatch<AUV_DUAL>(c, np, V, ld, U, W, s1, s2, out1, out2, out3); break;
        case AUV_TRI: low_dispatch<AUV_TRI>(c, np, V, ld, U, W, s1, s2, out1, out2, out3); break;
        default: throw std::invalid_argument("vertex-centric A(UV^T): unsupported mode");
        }
        c.launches++;
        LB2_CUDA(cudaGetLastError());
    }
    return;
}

}  // namespace lb2
