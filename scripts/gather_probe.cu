// gather_probe.cu -- stand-alone probe of row-gather mechanics for the two gather kernels (SpMM  Y = S X  on a
// symmetric pattern and the vertex-centric A(UV^T) pass) on a random MaxCut-like graph.  Not part of the library:
// it exists to choose the load mechanism (128-bit lane groups, 256-bit lane groups, cp.async.bulk row staging)
// from measurements on the B200.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather_probe gather_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s failed: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

constexpr int kBlock = 256;

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }
struct d4 { double a, b, c, d; };
__device__ __forceinline__ d4 ld4(const double *p) {
    d4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}
__device__ __forceinline__ void st4(double *p, d4 v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a), "d"(v.b), "d"(v.c), "d"(v.d) : "memory");
}

// ---------------------------------------------------------------------------------------------- A: round-1 style
template <int G, int NP>
__global__ void __launch_bounds__(kBlock) spmm_a(int n, int ld, const int *__restrict__ ap, const int *__restrict__ ac,
                                                 const int *__restrict__ apos, const double *__restrict__ S,
                                                 const double *__restrict__ X, double *__restrict__ Y) {
    constexpr int NG = kBlock / G;
    constexpr int UN = 4;
    const int g = threadIdx.x / G, gl = threadIdx.x % G;
    for (int i = blockIdx.x * NG + g; i < n; i += gridDim.x * NG) {
        double2 acc[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) acc[q] = make_double2(0, 0);
        const int ea = ap[i], eb = ap[i + 1];
        for (int e = ea; e < eb; e += UN) {
            int jn[UN];
            double sv[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const bool ok = (e + u) < eb;
                jn[u] = ok ? ac[e + u] : i;
                sv[u] = ok ? S[apos[ok ? e + u : ea]] : 0.0;
            }
            double2 v[UN][NP];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const double *xr = X + (size_t)jn[u] * ld + 2 * gl;
#pragma unroll
                for (int q = 0; q < NP; ++q) v[u][q] = ((2 * gl + 2 * G * q) < ld) ? ld2(xr + 2 * G * q) : make_double2(0, 0);
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
            if (col < ld) st2(Y + (size_t)i * ld + col, acc[q]);
        }
    }
}

// ---------------------------------------------------------------------------------------------- B: 256-bit, 8 lanes/row
// Values inline with the adjacency (no position indirection).  Each lane of the group fetches one (col,val) pair
// of the next 8 entries, the pairs are broadcast by shuffles, UN row gathers are in flight per lane.
template <int UN>
__global__ void __launch_bounds__(kBlock) spmm_b(int n, int ld, const int *__restrict__ ap, const int *__restrict__ ac,
                                                 const double *__restrict__ av, const double *__restrict__ X,
                                                 double *__restrict__ Y, const int *__restrict__ perm) {
    constexpr int G = 8, NG = kBlock / G;
    const int g = threadIdx.x / G, gl = threadIdx.x % G;
    const bool act = 4 * gl < ld;
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
    for (int s = blockIdx.x * NG + g; s < n; s += gridDim.x * NG) {
        const int i = perm ? perm[s] : s;
        d4 acc = {0, 0, 0, 0};
        const int ea = ap[i], eb = ap[i + 1];
        for (int e = ea; e < eb; e += 8) {
            const bool ok = e + gl < eb;
            const int jc = ok ? ac[e + gl] : 0;
            const double sc = ok ? av[e + gl] : 0.0;
            const int cnt = min(8, eb - e);
#pragma unroll
            for (int h = 0; h < 8; h += UN) {
                if (h < cnt) {
                    d4 v[UN];
                    double sv[UN];
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        const int j = __shfl_sync(gmask, jc, h + u, 8);
                        sv[u] = __shfl_sync(gmask, sc, h + u, 8);
                        v[u] = (act && (h + u) < cnt) ? ld4(X + (size_t)j * ld + 4 * gl) : d4{0, 0, 0, 0};
                    }
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        acc.a = fma(sv[u], v[u].a, acc.a);
                        acc.b = fma(sv[u], v[u].b, acc.b);
                        acc.c = fma(sv[u], v[u].c, acc.c);
                        acc.d = fma(sv[u], v[u].d, acc.d);
                    }
                }
            }
        }
        if (act) st4(Y + (size_t)i * ld + 4 * gl, acc);
    }
}

// ---------------------------------------------------------------------------------------------- B2: 128-bit, 16 lanes/row
// (one 16-byte load per lane covers rows up to 32 columns; 2 rows per warp instruction, 2 lines per row)
template <int UN>
__global__ void __launch_bounds__(kBlock) spmm_b2(int n, int ld, const int *__restrict__ ap, const int *__restrict__ ac,
                                                  const double *__restrict__ av, const double *__restrict__ X,
                                                  double *__restrict__ Y) {
    constexpr int G = 16, NG = kBlock / G;
    const int g = threadIdx.x / G, gl = threadIdx.x % G;
    const bool act = 2 * gl < ld;
    const unsigned gmask = 0xffffu << ((threadIdx.x & 31) & ~15);
    for (int i = blockIdx.x * NG + g; i < n; i += gridDim.x * NG) {
        double2 acc = make_double2(0, 0);
        const int ea = ap[i], eb = ap[i + 1];
        for (int e = ea; e < eb; e += 16) {
            const bool ok = e + gl < eb;
            const int jc = ok ? ac[e + gl] : 0;
            const double sc = ok ? av[e + gl] : 0.0;
            const int cnt = min(16, eb - e);
#pragma unroll
            for (int h = 0; h < 16; h += UN) {
                if (h < cnt) {
                    double2 v[UN];
                    double sv[UN];
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        const int j = __shfl_sync(gmask, jc, h + u, 16);
                        sv[u] = __shfl_sync(gmask, sc, h + u, 16);
                        v[u] = (act && (h + u) < cnt) ? ld2(X + (size_t)j * ld + 2 * gl) : make_double2(0, 0);
                    }
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        acc.x = fma(sv[u], v[u].x, acc.x);
                        acc.y = fma(sv[u], v[u].y, acc.y);
                    }
                }
            }
        }
        if (act) st2(Y + (size_t)i * ld + 2 * gl, acc);
    }
}

// ---------------------------------------------------------------------------------------------- E: cp.async.bulk row staging
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int cnt) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done)
                     : "r"(smem_u32(b)), "r"(phase)
                     : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}

// Every warp owns a contiguous range of rows (hence a contiguous range of adjacency entries) and a private ring of
// STAGES x 32 staged rows.  Per batch of 32 entries each lane issues one bulk copy of a whole factor row; the batch
// is then consumed from shared memory with lane = column.
template <int STAGES>
__global__ void __launch_bounds__(kBlock) spmm_e(int n, int ld, const int *__restrict__ ap, const int *__restrict__ ac,
                                                 const double *__restrict__ av, const double *__restrict__ X,
                                                 double *__restrict__ Y, int rows_per_warp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowb = ld * 8;
    unsigned char *ring = smem + (size_t)w * STAGES * 32 * rowb;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)(kBlock / 32) * STAGES * 32 * rowb) + w * STAGES;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int gw = blockIdx.x * (kBlock / 32) + w;
    const int r0 = min(n, gw * rows_per_warp), r1 = min(n, r0 + rows_per_warp);
    if (r0 >= r1) return;
    const int e0 = ap[r0], e1 = ap[r1];
    const int nb = (e1 - e0 + 31) / 32;
    double sreg[STAGES];
    auto issue = [&](int b) {
        const int st = b % STAGES;
        const int base = e0 + b * 32;
        const int cnt = min(32, e1 - base);
        if (lane == 0) mbar_expect_tx(bars + st, (uint32_t)(cnt * rowb));
        __syncwarp();
        double sv = 0.0;
        if (lane < cnt) {
            const int j = ac[base + lane];
            sv = av[base + lane];
            bulk_g2s(ring + ((size_t)st * 32 + lane) * rowb, X + (size_t)j * ld, (uint32_t)rowb, bars + st);
        }
        return sv;
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) sreg[s] = (s < nb) ? issue(s) : 0.0;
    int i = r0;
    int eb = ap[i + 1];
    double acc = 0.0;
    for (int b = 0; b < nb; ++b) {
        // keep STAGES-1 batches in flight
        double snew = (b + STAGES - 1 < nb) ? issue(b + STAGES - 1) : 0.0;
        const int st = b % STAGES;
        mbar_wait(bars + st, (uint32_t)((b / STAGES) & 1));
        const double scur = sreg[0];
#pragma unroll
        for (int s = 0; s < STAGES - 2; ++s) sreg[s] = sreg[s + 1];
        sreg[STAGES - 2] = snew;
        const int base = e0 + b * 32;
        const int cnt = min(32, e1 - base);
        const double *rows = reinterpret_cast<const double *>(ring + (size_t)st * 32 * rowb);
        for (int k = 0; k < cnt; ++k) {
            while (base + k >= eb) {
                if (lane < ld) Y[(size_t)i * ld + lane] = acc;
                acc = 0.0;
                ++i;
                eb = ap[i + 1];
            }
            const double s = __shfl_sync(0xffffffffu, scur, k);
            const double x = (lane < ld) ? rows[k * ld + lane] : 0.0;
            acc = fma(s, x, acc);
        }
        __syncwarp();
    }
    for (; i < r1; ++i) {
        if (lane < ld) Y[(size_t)i * ld + lane] = acc;
        acc = 0.0;
    }
}

// ---------------------------------------------------------------------------------------------- T: vertex-centric A(UV^T)
// Column j of the lower-triangular pattern (CSC, diagonal first): owner rows R_j, D_j are loaded once, the rows of
// the lower neighbours are gathered.  Outputs: q1_j = 2 R_j.D_j, q2_j = D_j.D_j, q3_j = R_j.R_j (diagonal
// constraints) and the two objective sums p1 = 2<C, sym(R D^T)>, p2 = <C, D D^T> accumulated per lane.
template <int UN>
__global__ void __launch_bounds__(kBlock) tri_b(int n, int ld, const int *__restrict__ cp, const int *__restrict__ ci,
                                                const double *__restrict__ cv, const double *__restrict__ R,
                                                const double *__restrict__ D, double *__restrict__ q1,
                                                double *__restrict__ q2, double *__restrict__ q3,
                                                double *__restrict__ partial) {
    constexpr int G = 8, NG = kBlock / G;
    const int g = threadIdx.x / G, gl = threadIdx.x % G;
    const bool act = 4 * gl < ld;
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
    double p1 = 0.0, p2 = 0.0;
    for (int j = blockIdx.x * NG + g; j < n; j += gridDim.x * NG) {
        const d4 rj = act ? ld4(R + (size_t)j * ld + 4 * gl) : d4{0, 0, 0, 0};
        const d4 dj = act ? ld4(D + (size_t)j * ld + 4 * gl) : d4{0, 0, 0, 0};
        const int ea = cp[j], eb = cp[j + 1];
        // diagonal entry (first of the column): constraint outputs need the complete dot products
        double t1 = rj.a * dj.a + rj.b * dj.b + rj.c * dj.c + rj.d * dj.d;
        double t2 = dj.a * dj.a + dj.b * dj.b + dj.c * dj.c + dj.d * dj.d;
        double t3 = rj.a * rj.a + rj.b * rj.b + rj.c * rj.c + rj.d * rj.d;
        const double cjj = cv[ea];
        p1 = fma(cjj, t1, p1);
        p2 = fma(cjj, t2, p2);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            t1 += __shfl_xor_sync(gmask, t1, o, 8);
            t2 += __shfl_xor_sync(gmask, t2, o, 8);
            t3 += __shfl_xor_sync(gmask, t3, o, 8);
        }
        if (gl == 0) { q1[j] = 2.0 * t1; q2[j] = t2; q3[j] = t3; }
        for (int e = ea + 1; e < eb; e += 8) {
            const bool ok = e + gl < eb;
            const int ic = ok ? ci[e + gl] : 0;
            const double cc = ok ? cv[e + gl] : 0.0;
            const int cnt = min(8, eb - e);
#pragma unroll
            for (int h = 0; h < 8; h += UN) {
                if (h < cnt) {
                    d4 ri[UN], di[UN];
                    double c[UN];
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        const int i = __shfl_sync(gmask, ic, h + u, 8);
                        c[u] = __shfl_sync(gmask, cc, h + u, 8);
                        const bool on = act && (h + u) < cnt;
                        ri[u] = on ? ld4(R + (size_t)i * ld + 4 * gl) : d4{0, 0, 0, 0};
                        di[u] = on ? ld4(D + (size_t)i * ld + 4 * gl) : d4{0, 0, 0, 0};
                    }
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        // off-diagonal: weight 2c on sym(R D^T)_ij = (R_i.D_j + R_j.D_i)/2  and on (D D^T)_ij
                        double a = ri[u].a * dj.a + ri[u].b * dj.b + ri[u].c * dj.c + ri[u].d * dj.d;
                        a += rj.a * di[u].a + rj.b * di[u].b + rj.c * di[u].c + rj.d * di[u].d;
                        const double b = di[u].a * dj.a + di[u].b * dj.b + di[u].c * dj.c + di[u].d * dj.d;
                        p1 = fma(c[u], a, p1);
                        p2 = fma(2.0 * c[u], b, p2);
                    }
                }
            }
        }
    }
    // block partials (deterministic): warp shuffle + shared
    __shared__ double sm[2][kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        p1 += __shfl_xor_sync(0xffffffffu, p1, o);
        p2 += __shfl_xor_sync(0xffffffffu, p2, o);
    }
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = p1; sm[1][threadIdx.x >> 5] = p2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int k = 0; k < kBlock / 32; ++k) { a += sm[0][k]; b += sm[1][k]; }
        partial[2 * blockIdx.x] = 2.0 * a;
        partial[2 * blockIdx.x + 1] = b;
    }
}


// ---------------------------------------------------------------------------------------------- ceilings
__device__ __forceinline__ d4 ld4_na(const double *p) {
    d4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.a), "=d"(r.b), "=d"(r.c), "=d"(r.d) : "l"(p));
    return r;
}
__global__ void __launch_bounds__(kBlock) stream_read(const double *__restrict__ X, size_t n4, double *out) {
    double s = 0;
    for (size_t i = blockIdx.x * (size_t)kBlock + threadIdx.x; i < n4; i += (size_t)gridDim.x * kBlock) {
        d4 v = ld4(X + 4 * i);
        s += v.a + v.b + v.c + v.d;
    }
    if (s == 123.456) out[0] = s;
}
// pure row gather: index list read coalesced (one entry per lane, broadcast by shuffle), rows of `ld` doubles,
// 8 lanes x 32 B per row, UN rows in flight per lane; nothing but the sum is kept.
template <int UN, bool NA>
__global__ void __launch_bounds__(kBlock) row_gather(const int *__restrict__ idx, long long cnt, int ld,
                                                     const double *__restrict__ X, double *out) {
    constexpr int G = 8;
    const int gl = threadIdx.x % G;
    const bool act = 4 * gl < ld;
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
    const long long ngroups = (long long)gridDim.x * (kBlock / G);
    const long long g = blockIdx.x * (long long)(kBlock / G) + threadIdx.x / G;
    double s = 0;
    for (long long e = g * 8; e < cnt; e += ngroups * 8) {
        const int jc = (e + gl < cnt) ? idx[e + gl] : 0;
#pragma unroll
        for (int h = 0; h < 8; h += UN) {
            d4 v[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int j = __shfl_sync(gmask, jc, h + u, 8);
                v[u] = act ? (NA ? ld4_na(X + (size_t)j * ld + 4 * gl) : ld4(X + (size_t)j * ld + 4 * gl)) : d4{0, 0, 0, 0};
            }
#pragma unroll
            for (int u = 0; u < UN; ++u) s += v[u].a + v[u].b + v[u].c + v[u].d;
        }
    }
    if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------------------------------------- F: TMA tile::gather4 staging
__device__ __forceinline__ void tma_gather4(void *dst, const void *tmap, int c0, int r0, int r1, int r2, int r3, uint64_t *b) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(b))
        : "memory");
}
template <int STAGES>
__global__ void __launch_bounds__(kBlock) spmm_f(int n, int ld, const int *__restrict__ ap, const int *__restrict__ ac,
                                                 const double *__restrict__ av, const __grid_constant__ CUtensorMap tmap,
                                                 double *__restrict__ Y, int rows_per_warp) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowb = ld * 8;
    unsigned char *ring = smem + (size_t)w * STAGES * 32 * rowb;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)(kBlock / 32) * STAGES * 32 * rowb) + w * STAGES;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int gw = blockIdx.x * (kBlock / 32) + w;
    const int r0 = min(n, gw * rows_per_warp), r1 = min(n, r0 + rows_per_warp);
    if (r0 >= r1) return;
    const int e0 = ap[r0], e1 = ap[r1];
    const int nb = (e1 - e0 + 31) / 32;
    double sreg[STAGES];
    auto issue = [&](int b) {
        const int st = b % STAGES;
        const int base = e0 + b * 32;
        const int cnt = min(32, e1 - base);
        const int nq = (cnt + 3) / 4;
        if (lane == 0) mbar_expect_tx(bars + st, (uint32_t)(nq * 4 * rowb));
        __syncwarp();
        int j = 0;
        double sv = 0.0;
        if (lane < cnt) { j = ac[base + lane]; sv = av[base + lane]; }
        const int j0 = __shfl_sync(0xffffffffu, j, (4 * lane) & 31), j1 = __shfl_sync(0xffffffffu, j, (4 * lane + 1) & 31);
        const int j2 = __shfl_sync(0xffffffffu, j, (4 * lane + 2) & 31), j3 = __shfl_sync(0xffffffffu, j, (4 * lane + 3) & 31);
        if (lane < nq) tma_gather4(ring + ((size_t)st * 32 + 4 * lane) * rowb, &tmap, 0, j0, j1, j2, j3, bars + st);
        return sv;
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) sreg[s] = (s < nb) ? issue(s) : 0.0;
    int i = r0;
    int eb = ap[i + 1];
    double acc = 0.0;
    for (int b = 0; b < nb; ++b) {
        double snew = (b + STAGES - 1 < nb) ? issue(b + STAGES - 1) : 0.0;
        const int st = b % STAGES;
        mbar_wait(bars + st, (uint32_t)((b / STAGES) & 1));
        const double scur = sreg[0];
#pragma unroll
        for (int s = 0; s < STAGES - 2; ++s) sreg[s] = sreg[s + 1];
        sreg[STAGES - 2] = snew;
        const int base = e0 + b * 32;
        const int cnt = min(32, e1 - base);
        const double *rows = reinterpret_cast<const double *>(ring + (size_t)st * 32 * rowb);
        for (int k = 0; k < cnt; ++k) {
            while (base + k >= eb) {
                if (lane < ld) Y[(size_t)i * ld + lane] = acc;
                acc = 0.0;
                ++i;
                eb = ap[i + 1];
            }
            const double s = __shfl_sync(0xffffffffu, scur, k);
            const double x = (lane < ld) ? rows[k * ld + lane] : 0.0;
            acc = fma(s, x, acc);
        }
        __syncwarp();
    }
    for (; i < r1; ++i) {
        if (lane < ld) Y[(size_t)i * ld + lane] = acc;
        acc = 0.0;
    }
}

template <int UN, bool NA, int MINB>
__global__ void __launch_bounds__(kBlock, MINB) spmm_b3(int n, int ld, const int *__restrict__ ap, const int *__restrict__ ac,
                                                  const double *__restrict__ av, const double *__restrict__ X,
                                                  double *__restrict__ Y, const int *__restrict__ perm) {
    constexpr int G = 8, NG = kBlock / G;
    const int g = threadIdx.x / G, gl = threadIdx.x % G;
    const bool act = 4 * gl < ld;
    const unsigned gmask = 0xffu << ((threadIdx.x & 31) & ~7);
    for (int s = blockIdx.x * NG + g; s < n; s += gridDim.x * NG) {
        const int i = perm ? perm[s] : s;
        d4 acc = {0, 0, 0, 0};
        const int ea = ap[i], eb = ap[i + 1];
        for (int e = ea; e < eb; e += 8) {
            const bool ok = e + gl < eb;
            const int jc = ok ? ac[e + gl] : 0;
            const double sc = ok ? av[e + gl] : 0.0;
            const int cnt = min(8, eb - e);
#pragma unroll
            for (int h = 0; h < 8; h += UN) {
                if (h < cnt) {
                    d4 v[UN];
                    double sv[UN];
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        const int j = __shfl_sync(gmask, jc, h + u, 8);
                        sv[u] = __shfl_sync(gmask, sc, h + u, 8);
                        const double *p = X + (size_t)j * ld + 4 * gl;
                        v[u] = (act && (h + u) < cnt) ? (NA ? ld4_na(p) : ld4(p)) : d4{0, 0, 0, 0};
                    }
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        acc.a = fma(sv[u], v[u].a, acc.a);
                        acc.b = fma(sv[u], v[u].b, acc.b);
                        acc.c = fma(sv[u], v[u].c, acc.c);
                        acc.d = fma(sv[u], v[u].d, acc.d);
                    }
                }
            }
        }
        if (act) st4(Y + (size_t)i * ld + 4 * gl, acc);
    }
}

// ---------------------------------------------------------------------------------------------- V: vertex-centric skeleton
// G = ceil(ld/4) lanes per row (<= 8), RW = 32/G rows per warp, 256-bit loads, one gather in flight per lane and
// many resident warps.  adjC = (col, val) static part, diagonal handled through (dcon) weights: S_jj += w[j].
template <int MINB>
__global__ void __launch_bounds__(kBlock, MINB) spmm_v(int n, int ld, int G, const int *__restrict__ order,
                                                        const int *__restrict__ ap, const int *__restrict__ ac,
                                                        const double *__restrict__ av, const double *__restrict__ w,
                                                        const double *__restrict__ X, double a, double *__restrict__ Y,
                                                        double *__restrict__ part) {
    const int lane = threadIdx.x & 31;
    const int RW = 32 / G;
    const int grp = lane / G, gl = lane - grp * G;
    const bool valid = grp < RW;
    const unsigned gmask = valid ? (((1u << G) - 1u) << (grp * G)) : 0u;
    const int base = grp * G;
    const long long wg = (long long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (kBlock / 32);
    double ryy = 0.0;
    if (valid)
        for (long long slot = wg * RW + grp; slot < n; slot += nw * RW) {
            const int i = order[slot];
            d4 acc = {0, 0, 0, 0};
            const int ea = ap[i], eb = ap[i + 1];
            for (int e = ea; e < eb; e += G) {
                const bool ok = e + gl < eb;
                const int jc = ok ? ac[e + gl] : 0;
                double sc = ok ? av[e + gl] : 0.0;
                if (ok && jc == i) sc += w[i];            // MaxCut: A_i = e_i e_i^T, S_ii = C_ii + w_i
                const int cnt = min(G, eb - e);
#pragma unroll 1
                for (int h = 0; h < cnt; ++h) {
                    const int j = __shfl_sync(gmask, jc, base + h);
                    const double s = __shfl_sync(gmask, sc, base + h);
                    const d4 v = ld4(X + (size_t)j * ld + 4 * gl);
                    acc.a = fma(s, v.a, acc.a); acc.b = fma(s, v.b, acc.b);
                    acc.c = fma(s, v.c, acc.c); acc.d = fma(s, v.d, acc.d);
                }
            }
            acc.a *= a; acc.b *= a; acc.c *= a; acc.d *= a;
            st4(Y + (size_t)i * ld + 4 * gl, acc);
            ryy = fma(acc.a, acc.a, ryy); ryy = fma(acc.b, acc.b, ryy); ryy = fma(acc.c, acc.c, ryy); ryy = fma(acc.d, acc.d, ryy);
        }
    __shared__ double sm[kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ryy += __shfl_xor_sync(0xffffffffu, ryy, o);
    if (lane == 0) sm[threadIdx.x >> 5] = ryy;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < kBlock / 32; ++k) t += sm[k];
        part[blockIdx.x] = t;
    }
}

// A(UV^T) TRI, MaxCut-shaped cone: objective through t_j = sum_i c_ij D_i (symmetric adjacency of C, one gather per
// entry), p1 = 2 sum_j R_j.t_j, p2 = sum_j D_j.t_j; diagonal constraints from the owner rows.
template <int MINB>
__global__ void __launch_bounds__(kBlock, MINB) tri_v(int n, int ld, int G, const int *__restrict__ order,
                                                       const int *__restrict__ ap, const int *__restrict__ ac,
                                                       const double *__restrict__ av, const double *__restrict__ R,
                                                       const double *__restrict__ D, double *__restrict__ q1,
                                                       double *__restrict__ q2, double *__restrict__ q3,
                                                       double *__restrict__ part) {
    const int lane = threadIdx.x & 31;
    const int RW = 32 / G;
    const int grp = lane / G, gl = lane - grp * G;
    const bool valid = grp < RW;
    const unsigned gmask = valid ? (((1u << G) - 1u) << (grp * G)) : 0u;
    const int base = grp * G;
    const long long wg = (long long)blockIdx.x * (kBlock / 32) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (kBlock / 32);
    double p1 = 0.0, p2 = 0.0;
    if (valid)
        for (long long slot = wg * RW + grp; slot < n; slot += nw * RW) {
            const int jv = order[slot];
            d4 acc = {0, 0, 0, 0};
            const int ea = ap[jv], eb = ap[jv + 1];
            for (int e = ea; e < eb; e += G) {
                const bool ok = e + gl < eb;
                const int jc = ok ? ac[e + gl] : 0;
                const double sc = ok ? av[e + gl] : 0.0;
                const int cnt = min(G, eb - e);
#pragma unroll 1
                for (int h = 0; h < cnt; ++h) {
                    const int j = __shfl_sync(gmask, jc, base + h);
                    const double s = __shfl_sync(gmask, sc, base + h);
                    const d4 v = ld4(D + (size_t)j * ld + 4 * gl);
                    acc.a = fma(s, v.a, acc.a); acc.b = fma(s, v.b, acc.b);
                    acc.c = fma(s, v.c, acc.c); acc.d = fma(s, v.d, acc.d);
                }
            }
            const d4 rj = ld4(R + (size_t)jv * ld + 4 * gl), dj = ld4(D + (size_t)jv * ld + 4 * gl);
            p1 += rj.a * acc.a + rj.b * acc.b + rj.c * acc.c + rj.d * acc.d;
            p2 += dj.a * acc.a + dj.b * acc.b + dj.c * acc.c + dj.d * acc.d;
            double t1 = rj.a * dj.a + rj.b * dj.b + rj.c * dj.c + rj.d * dj.d;
            double t2 = dj.a * dj.a + dj.b * dj.b + dj.c * dj.c + dj.d * dj.d;
            double t3 = rj.a * rj.a + rj.b * rj.b + rj.c * rj.c + rj.d * rj.d;
            for (int o = 1; o < G; o <<= 1) {
                const int src = base + ((gl + o) < G ? gl + o : gl);
                const double u1 = __shfl_sync(gmask, t1, src), u2 = __shfl_sync(gmask, t2, src), u3 = __shfl_sync(gmask, t3, src);
                if (gl + o < G && (gl & (2 * o - 1)) == 0) { t1 += u1; t2 += u2; t3 += u3; }
            }
            if (gl == 0) { q1[jv] = 2.0 * t1; q2[jv] = t2; q3[jv] = t3; }
        }
    __shared__ double sm[2][kBlock / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        p1 += __shfl_xor_sync(0xffffffffu, p1, o);
        p2 += __shfl_xor_sync(0xffffffffu, p2, o);
    }
    if (lane == 0) { sm[0][threadIdx.x >> 5] = p1; sm[1][threadIdx.x >> 5] = p2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0, y = 0;
        for (int k = 0; k < kBlock / 32; ++k) { x += sm[0][k]; y += sm[1][k]; }
        part[2 * blockIdx.x] = 2.0 * x;
        part[2 * blockIdx.x + 1] = y;
    }
}

__global__ void flush_kernel(double *p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = p[i] * 0.5 + 1.0;
}

// ================================================================================================== host
struct Graph {
    int n;
    std::vector<int> ap, ac, apos;       // symmetric adjacency incl. diagonal, sorted by column index
    std::vector<double> av, S;           // value inline / value by pattern position
    std::vector<int> cp, ci;             // lower-triangular CSC (diagonal first)
    std::vector<double> cv;
};

static Graph make_graph(int n, long long ne, unsigned seed) {
    Graph g;
    g.n = n;
    std::mt19937_64 rng(seed);
    std::vector<std::pair<int, int>> ed;
    ed.reserve(ne);
    std::vector<uint64_t> keys;
    keys.reserve(ne * 11 / 10);
    while ((long long)keys.size() < ne * 11 / 10) {
        int a = (int)(rng() % n), b = (int)(rng() % n);
        if (a == b) continue;
        if (a < b) std::swap(a, b);
        keys.push_back((uint64_t)b << 32 | (uint32_t)a);   // (col=min, row=max)
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    std::shuffle(keys.begin(), keys.end(), rng);
    if ((long long)keys.size() > ne) keys.resize(ne);
    std::sort(keys.begin(), keys.end());
    // pattern P: per column j: (j,j) then lower neighbours
    g.cp.assign(n + 1, 0);
    for (auto k : keys) g.cp[(k >> 32) + 1]++;
    for (int j = 0; j < n; ++j) g.cp[j + 1] += g.cp[j] + 1;
    g.ci.resize(g.cp[n]);
    g.cv.resize(g.cp[n]);
    std::vector<int> fill(n);
    std::vector<double> deg(n, 0.0);
    for (auto k : keys) { deg[k >> 32] += 1; deg[(uint32_t)k] += 1; }
    for (int j = 0; j < n; ++j) { g.ci[g.cp[j]] = j; g.cv[g.cp[j]] = -0.25 * deg[j]; fill[j] = g.cp[j] + 1; }
    for (auto k : keys) { int j = (int)(k >> 32), i = (int)(uint32_t)k; g.ci[fill[j]] = i; g.cv[fill[j]] = 0.25; fill[j]++; }
    g.S = g.cv;
    // symmetric adjacency
    std::vector<int> cnt(n, 1);
    for (auto k : keys) { cnt[k >> 32]++; cnt[(uint32_t)k]++; }
    g.ap.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) g.ap[i + 1] = g.ap[i] + cnt[i];
    g.ac.resize(g.ap[n]); g.apos.resize(g.ap[n]); g.av.resize(g.ap[n]);
    std::vector<int> f2(g.ap.begin(), g.ap.end() - 1);
    // entries sorted by neighbour index: walk columns in order, (col j) contributes to row i>j entry j, and to row j entry i
    for (int j = 0; j < n; ++j) {
        for (int p = g.cp[j]; p < g.cp[j + 1]; ++p) {
            const int i = g.ci[p];
            // row i gets neighbour j at position p
            g.ac[f2[i]] = j; g.apos[f2[i]] = p; g.av[f2[i]] = g.S[p]; f2[i]++;
        }
    }
    // second pass: for row j, larger neighbours i are exactly column j's lower entries (already sorted by i)
    for (int j = 0; j < n; ++j)
        for (int p = g.cp[j] + 1; p < g.cp[j + 1]; ++p) {
            g.ac[f2[j]] = g.ci[p]; g.apos[f2[j]] = p; g.av[f2[j]] = g.S[p]; f2[j]++;
        }
    return g;
}

template <typename T>
static T *dev(const std::vector<T> &h) {
    T *d;
    CK(cudaMalloc(&d, h.size() * sizeof(T)));
    CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

static double *g_flush = nullptr;
static size_t g_flush_n = 0;

template <typename F>
static void timeit(const char *name, F launch, double alg_bytes, bool cold, int reps = 20) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    std::vector<float> ts;
    for (int i = 0; i < reps; ++i) {
        if (cold) flush_kernel<<<1184, 256>>>(g_flush, g_flush_n);
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        ts.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(ts.begin(), ts.end());
    const double us = ts[ts.size() / 2] * 1e3;
    printf("  %-34s %s  %9.1f us  (min %8.1f)  alg %7.1f GB/s  frac %.3f\n", name, cold ? "cold" : "hot ", us, ts[0] * 1e3,
           alg_bytes / us * 1e-3, alg_bytes / us * 1e-3 / 6554.6);
    fflush(stdout);
}

static double maxrel(const std::vector<double> &a, const std::vector<double> &b) {
    double m = 0, s = 0;
    for (size_t i = 0; i < a.size(); ++i) { m = std::max(m, std::fabs(a[i] - b[i])); s = std::max(s, std::fabs(b[i])); }
    return m / (s > 0 ? s : 1);
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 100000;
    const int ld = argc > 2 ? atoi(argv[2]) : 24;
    const long long ne = argc > 3 ? atoll(argv[3]) : 5LL * n;
    const int which = argc > 4 ? atoi(argv[4]) : 0xff;
    printf("gather_probe n=%d ld=%d edges=%lld\n", n, ld, ne);
    Graph g = make_graph(n, ne, 3);
    const size_t N = (size_t)n * ld;
    std::vector<double> hX(N), hD(N);
    std::mt19937_64 rng(925);
    for (auto &x : hX) x = (double)(rng() % 2000001) / 1e6 - 1.0;
    for (auto &x : hD) x = (double)(rng() % 2000001) / 1e6 - 1.0;
    // degree-sorted row order
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return g.ap[a + 1] - g.ap[a] > g.ap[b + 1] - g.ap[b]; });

    int *ap = dev(g.ap), *ac = dev(g.ac), *apos = dev(g.apos), *cp = dev(g.cp), *ci = dev(g.ci), *dperm = dev(perm);
    double *av = dev(g.av), *S = dev(g.S), *cv = dev(g.cv), *X = dev(hX), *D = dev(hD), *Y, *Yref;
    CK(cudaMalloc(&Y, N * 8)); CK(cudaMalloc(&Yref, N * 8));
    g_flush_n = (size_t)48 << 20;   // 384 MB
    CK(cudaMalloc(&g_flush, g_flush_n * 8));
    CK(cudaMemset(g_flush, 0, g_flush_n * 8));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const long long nadj = g.ap[n], np = g.cp[n];
    const double spmm_bytes = 8.0 * nadj + 8.0 * np + 4.0 * (n + 1) + 16.0 * N;   // round-1 model (adj (col,pos) + S + ptr + X + Y)
    const double tri_bytes = 16.0 * N + 16.0 * (np + n) + 4.0 * (n + 2) + 3 * 8.0 * (n + 1);

    // reference result (kernel A, verified on the host for a few rows)
    const int gridA = std::min(8192, (n + 63) / 64);
    spmm_a<4, 3><<<gridA, kBlock>>>(n, ld, ap, ac, apos, S, X, Yref);
    if (ld > 24) spmm_a<4, 4><<<gridA, kBlock>>>(n, ld, ap, ac, apos, S, X, Yref);
    CK(cudaDeviceSynchronize());
    std::vector<double> hY(N), hYr(N);
    CK(cudaMemcpy(hYr.data(), Yref, N * 8, cudaMemcpyDeviceToHost));
    {
        double m = 0;
        for (int i = 0; i < n; i += std::max(1, n / 50)) {
            for (int c = 0; c < ld; ++c) {
                double s = 0;
                for (int e = g.ap[i]; e < g.ap[i + 1]; ++e) s += g.av[e] * hX[(size_t)g.ac[e] * ld + c];
                m = std::max(m, std::fabs(s - hYr[(size_t)i * ld + c]));
            }
        }
        printf("host check of reference kernel: max abs diff %.3e\n", m);
    }
    auto check = [&](const char *nm) {
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hY.data(), Y, N * 8, cudaMemcpyDeviceToHost));
        printf("  [%s] max rel diff vs reference %.3e\n", nm, maxrel(hY, hYr));
        CK(cudaMemset(Y, 0, N * 8));
    };

    double *q1, *q2, *q3, *part;
    CK(cudaMalloc(&q1, 8 * (size_t)n)); CK(cudaMalloc(&q2, 8 * (size_t)n)); CK(cudaMalloc(&q3, 8 * (size_t)n));
    CK(cudaMalloc(&part, 8 * 2 * 65536));
    for (int cold = 0; cold < 2; ++cold) {
        printf("--- SpMM (%s)\n", cold ? "L2 flushed before each launch" : "back to back");
        if (which & 1) {
            if (ld <= 24) timeit("A  G=4 NP=3 pos-indirect", [&] { spmm_a<4, 3><<<gridA, kBlock>>>(n, ld, ap, ac, apos, S, X, Y); }, spmm_bytes, cold);
            else timeit("A  G=4 NP=4 pos-indirect", [&] { spmm_a<4, 4><<<gridA, kBlock>>>(n, ld, ap, ac, apos, S, X, Y); }, spmm_bytes, cold);
        }
        if (which & 2) {
            const int NG = kBlock / 8;
            for (int occ : {4, 8}) {
                const int grid = std::min((n + NG - 1) / NG, sms * occ);
                char nm[64];
                snprintf(nm, 64, "B  256-bit UN=4 grid=%dxSM", occ);
                timeit(nm, [&] { spmm_b<4><<<grid, kBlock>>>(n, ld, ap, ac, av, X, Y, nullptr); }, spmm_bytes, cold);
                snprintf(nm, 64, "B  256-bit UN=8 grid=%dxSM", occ);
                timeit(nm, [&] { spmm_b<8><<<grid, kBlock>>>(n, ld, ap, ac, av, X, Y, nullptr); }, spmm_bytes, cold);
            }
            const int gridf = (n + NG - 1) / NG;
            timeit("B  256-bit UN=4 one group/row", [&] { spmm_b<4><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y, nullptr); }, spmm_bytes, cold);
            timeit("B  256-bit UN=8 one group/row", [&] { spmm_b<8><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y, nullptr); }, spmm_bytes, cold);
            if (!cold) check("B");
            timeit("B  256-bit UN=8 degree-sorted", [&] { spmm_b<8><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y, dperm); }, spmm_bytes, cold);
            timeit("B  256-bit UN=4 degree-sorted", [&] { spmm_b<4><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y, dperm); }, spmm_bytes, cold);
            if (!cold) check("B sorted");
        }
        if (which & 4) {
            const int NG = kBlock / 16;
            const int gridf = (n + NG - 1) / NG;
            timeit("B2 128-bit G=16 UN=4", [&] { spmm_b2<4><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y); }, spmm_bytes, cold);
            timeit("B2 128-bit G=16 UN=8", [&] { spmm_b2<8><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y); }, spmm_bytes, cold);
            if (!cold) check("B2");
        }
        if (which & 8) {
            for (int ctas : {1, 2}) {
                const int warps = sms * ctas * (kBlock / 32);
                for (int mult : {1, 4}) {
                    const int rpw = (n + warps * mult - 1) / (warps * mult);
                    const int grid = (n + rpw * 8 - 1) / (rpw * 8);
                    char nm[64];
                    if (ctas == 1) {
                        const size_t sh = (size_t)8 * 3 * 32 * ld * 8 + 8 * 3 * 8;
                        CK(cudaFuncSetAttribute(spmm_e<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
                        snprintf(nm, 64, "E  bulk rows 3 stages rpw=%d", rpw);
                        timeit(nm, [&] { spmm_e<3><<<grid, kBlock, sh>>>(n, ld, ap, ac, av, X, Y, rpw); }, spmm_bytes, cold);
                    } else {
                        const size_t sh = (size_t)8 * 2 * 32 * ld * 8 + 8 * 2 * 8;
                        CK(cudaFuncSetAttribute(spmm_e<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
                        snprintf(nm, 64, "E  bulk rows 2 stages x2 rpw=%d", rpw);
                        timeit(nm, [&] { spmm_e<2><<<grid, kBlock, sh>>>(n, ld, ap, ac, av, X, Y, rpw); }, spmm_bytes, cold);
                    }
                }
            }
            if (!cold) check("E");
        }
    }


    if (which & 32) {
        printf("--- ceilings\n");
        double *out; CK(cudaMalloc(&out, 64));
        for (int cold = 0; cold < 2; ++cold) {
            timeit("stream read of X (256-bit)", [&] { stream_read<<<sms * 8, kBlock>>>(X, N / 4, out); }, 8.0 * N, cold);
            // random index list, nadj entries
            static int *ridx = nullptr;
            if (!ridx) {
                std::vector<int> h(nadj);
                std::mt19937_64 r2(7);
                for (auto &x : h) x = (int)(r2() % n);
                ridx = dev(h);
            }
            const double gb = (double)nadj * ld * 8;
            for (int occ : {2, 4, 8}) {
                char nm[64];
                snprintf(nm, 64, "row gather UN=1 grid=%dxSM", occ);
                timeit(nm, [&] { row_gather<1, false><<<sms * occ, kBlock>>>(ridx, nadj, ld, X, out); }, gb, cold);
                snprintf(nm, 64, "row gather UN=2 grid=%dxSM", occ);
                timeit(nm, [&] { row_gather<2, false><<<sms * occ, kBlock>>>(ridx, nadj, ld, X, out); }, gb, cold);
                snprintf(nm, 64, "row gather UN=4 grid=%dxSM", occ);
                timeit(nm, [&] { row_gather<4, false><<<sms * occ, kBlock>>>(ridx, nadj, ld, X, out); }, gb, cold);
                snprintf(nm, 64, "row gather UN=8 grid=%dxSM", occ);
                timeit(nm, [&] { row_gather<8, false><<<sms * occ, kBlock>>>(ridx, nadj, ld, X, out); }, gb, cold);
                snprintf(nm, 64, "row gather UN=4 no_alloc grid=%dxSM", occ);
                timeit(nm, [&] { row_gather<4, true><<<sms * occ, kBlock>>>(ridx, nadj, ld, X, out); }, gb, cold);
            }
            // the adjacency's own index list (sorted within rows)
            timeit("row gather UN=4 adj order 8xSM", [&] { row_gather<4, false><<<sms * 8, kBlock>>>(ac, nadj, ld, X, out); }, gb, cold);
        }
    }
    if (which & 128) {
        const int NG = kBlock / 8;
        const int gridf = (n + NG - 1) / NG;
        for (int cold = 0; cold < 2; ++cold) {
            printf("--- SpMM B variants (%s)\n", cold ? "cold" : "hot");
            timeit("B3 UN=2 sorted", [&] { spmm_b3<2, false, 1><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y, dperm); }, spmm_bytes, cold);
            timeit("B3 UN=2 sorted minb=8", [&] { spmm_b3<2, false, 8><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y, dperm); }, spmm_bytes, cold);
            timeit("B3 UN=4 sorted minb=6", [&] { spmm_b3<4, false, 6><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y, dperm); }, spmm_bytes, cold);
            timeit("B3 UN=4 sorted no_alloc", [&] { spmm_b3<4, true, 1><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y, dperm); }, spmm_bytes, cold);
            timeit("B3 UN=2 sorted no_alloc", [&] { spmm_b3<2, true, 1><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y, dperm); }, spmm_bytes, cold);
            timeit("B3 UN=1 sorted minb=8", [&] { spmm_b3<1, false, 8><<<gridf, kBlock>>>(n, ld, ap, ac, av, X, Y, dperm); }, spmm_bytes, cold);
            if (!cold) check("B3");
        }
    }
    if (which & 64) {
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        const int boxrows = argc > 5 ? atoi(argv[5]) : 1;
        CUtensorMap tm;
        cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)n};
        cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
        cuuint32_t box[2] = {(cuuint32_t)ld, (cuuint32_t)boxrows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, X, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("--- TMA gather4 (box rows %d): encode rc=%d\n", boxrows, (int)r);
        if (r == CUDA_SUCCESS) {
            for (int cold = 0; cold < 2; ++cold)
                for (int ctas : {1, 2}) {
                    const int warps = sms * ctas * (kBlock / 32);
                    const int rpw = (n + warps * 2 - 1) / (warps * 2);
                    const int grid = (n + rpw * 8 - 1) / (rpw * 8);
                    char nm[64];
                    if (ctas == 1) {
                        const size_t sh = (size_t)8 * 3 * 32 * ld * 8 + 8 * 3 * 8;
                        CK(cudaFuncSetAttribute(spmm_f<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
                        snprintf(nm, 64, "F  gather4 3 stages rpw=%d", rpw);
                        timeit(nm, [&] { spmm_f<3><<<grid, kBlock, sh>>>(n, ld, ap, ac, av, tm, Y, rpw); }, spmm_bytes, cold, 10);
                    } else {
                        const size_t sh = (size_t)8 * 2 * 32 * ld * 8 + 8 * 2 * 8;
                        CK(cudaFuncSetAttribute(spmm_f<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
                        snprintf(nm, 64, "F  gather4 2 stages x2 rpw=%d", rpw);
                        timeit(nm, [&] { spmm_f<2><<<grid, kBlock, sh>>>(n, ld, ap, ac, av, tm, Y, rpw); }, spmm_bytes, cold, 10);
                    }
                    if (!cold) check("F");
                }
        }
    }

    if (which & 256) {
        const int G = (ld + 3) / 4, RW = 32 / G;
        std::vector<int> ident(n);
        for (int i = 0; i < n; ++i) ident[i] = i;
        int *dident = dev(ident);
        std::vector<double> hw(n, 0.0);
        double *dw = dev(hw);
        double *part2; CK(cudaMalloc(&part2, 8 * 2 * 70000));
        const int rows_per_block = RW * (kBlock / 32);
        const int gridf = (n + rows_per_block - 1) / rows_per_block;
        for (int cold = 0; cold < 2; ++cold) {
            printf("--- vertex-centric skeleton G=%d RW=%d (%s)\n", G, RW, cold ? "cold" : "hot");
            timeit("V spmm sorted minb=8 full grid", [&] { spmm_v<8><<<gridf, kBlock>>>(n, ld, G, dperm, ap, ac, av, dw, X, 1.0, Y, part2); }, spmm_bytes, cold);
            timeit("V spmm sorted minb=5 full grid", [&] { spmm_v<5><<<gridf, kBlock>>>(n, ld, G, dperm, ap, ac, av, dw, X, 1.0, Y, part2); }, spmm_bytes, cold);
            timeit("V spmm natural minb=8 full grid", [&] { spmm_v<8><<<gridf, kBlock>>>(n, ld, G, dident, ap, ac, av, dw, X, 1.0, Y, part2); }, spmm_bytes, cold);
            timeit("V spmm sorted minb=8 grid 8xSM", [&] { spmm_v<8><<<std::min(gridf, sms * 8), kBlock>>>(n, ld, G, dperm, ap, ac, av, dw, X, 1.0, Y, part2); }, spmm_bytes, cold);
            if (!cold) check("V spmm");
            timeit("V tri sorted minb=8 full grid", [&] { tri_v<8><<<gridf, kBlock>>>(n, ld, G, dperm, ap, ac, av, X, D, q1, q2, q3, part2); }, tri_bytes, cold);
            timeit("V tri sorted minb=5 full grid", [&] { tri_v<5><<<gridf, kBlock>>>(n, ld, G, dperm, ap, ac, av, X, D, q1, q2, q3, part2); }, tri_bytes, cold);
            timeit("V tri natural minb=8 full grid", [&] { tri_v<8><<<gridf, kBlock>>>(n, ld, G, dident, ap, ac, av, X, D, q1, q2, q3, part2); }, tri_bytes, cold);
        }
        tri_v<8><<<gridf, kBlock>>>(n, ld, G, dperm, ap, ac, av, X, D, q1, q2, q3, part2);
        CK(cudaDeviceSynchronize());
        std::vector<double> hp(2 * gridf), hq(n);
        CK(cudaMemcpy(hp.data(), part2, 8 * 2 * gridf, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hq.data(), q1, 8 * n, cudaMemcpyDeviceToHost));
        double p1 = 0, p2 = 0;
        for (int b = 0; b < gridf; ++b) { p1 += hp[2 * b]; p2 += hp[2 * b + 1]; }
        double q1ref = 0;
        for (int k = 0; k < ld; ++k) q1ref += 2.0 * hX[(size_t)7 * ld + k] * hD[(size_t)7 * ld + k];
        printf("  tri_v p1 %.15e p2 %.15e  q1[7] %.15e (host %.15e)\n", p1, p2, hq[7], q1ref);
    }
    // ---- vertex-centric A(UV^T)
    if (which & 16) {
        for (int cold = 0; cold < 2; ++cold) {
            printf("--- vertex-centric A(UV^T) TRI (%s)\n", cold ? "L2 flushed" : "back to back");
            const int NG = kBlock / 8;
            const int gridf = std::min(65536, (n + NG - 1) / NG);
            timeit("T  256-bit UN=2", [&] { tri_b<2><<<gridf, kBlock>>>(n, ld, cp, ci, cv, X, D, q1, q2, q3, part); }, tri_bytes, cold);
            timeit("T  256-bit UN=4", [&] { tri_b<4><<<gridf, kBlock>>>(n, ld, cp, ci, cv, X, D, q1, q2, q3, part); }, tri_bytes, cold);
            const int grid8 = std::min(gridf, sms * 8);
            timeit("T  256-bit UN=4 grid=8xSM", [&] { tri_b<4><<<grid8, kBlock>>>(n, ld, cp, ci, cv, X, D, q1, q2, q3, part); }, tri_bytes, cold);
        }
        // check p1, p2 against the host
        const int NG = kBlock / 8;
        const int gridf = std::min(65536, (n + NG - 1) / NG);
        tri_b<4><<<gridf, kBlock>>>(n, ld, cp, ci, cv, X, D, q1, q2, q3, part);
        CK(cudaDeviceSynchronize());
        std::vector<double> hp(2 * gridf);
        CK(cudaMemcpy(hp.data(), part, 8 * 2 * gridf, cudaMemcpyDeviceToHost));
        double p1 = 0, p2 = 0;
        for (int b = 0; b < gridf; ++b) { p1 += hp[2 * b]; p2 += hp[2 * b + 1]; }
        double r1 = 0, r2 = 0;
        for (int j = 0; j < n; ++j)
            for (int p = g.cp[j]; p < g.cp[j + 1]; ++p) {
                const int i = g.ci[p];
                double a = 0, b = 0, c = 0;
                for (int k = 0; k < ld; ++k) {
                    a += hX[(size_t)i * ld + k] * hD[(size_t)j * ld + k];
                    b += hX[(size_t)j * ld + k] * hD[(size_t)i * ld + k];
                    c += hD[(size_t)i * ld + k] * hD[(size_t)j * ld + k];
                }
                const double wgt = (i == j) ? g.cv[p] : 2.0 * g.cv[p];
                r1 += 2.0 * wgt * 0.5 * (a + b);
                r2 += wgt * c;
            }
        printf("  p1 %.15e (host %.15e)  p2 %.15e (host %.15e)\n", p1, r1, p2, r2);
    }
    printf("done\n");
    return 0;
}
