// common.cuh -- error handling, device buffers and deterministic reduction helpers (sm_100a).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace lb2 {

struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

#define LB2_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw lb2::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +   \
                                 __FILE__ + ":" + std::to_string(__LINE__));                        \
    } while (0)

// Owning device array.
template <typename T>
struct DBuf {
    T *p = nullptr;
    size_t n = 0;
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    DBuf(DBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DBuf &operator=(DBuf &&o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr; n = 0;
    }
    void alloc(size_t count, bool zero = true) {
        release();
        n = count;
        if (count == 0) return;
        LB2_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
        if (zero) LB2_CUDA(cudaMemset(p, 0, count * sizeof(T)));
    }
    void upload(const std::vector<T> &h) {
        alloc(h.size(), false);
        if (!h.empty()) LB2_CUDA(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    }
};

#ifdef __CUDACC__

constexpr int kBlock = 256;

template <int G>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum(double v) { return group_sum<32>(v); }

// Deterministic block sum (blockDim.x == kBlock).  Result valid in thread 0.
__device__ __forceinline__ double block_sum(double v, double *sm /* >= 8 doubles */) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < kBlock / 32; ++k) t += sm[k];
    }
    return t;
}

// Grid-wide deterministic reduction of NV per-thread values.  Every block writes its partial, the last
// block to arrive (atomic ticket) adds the partials in a fixed order that does not depend on which
// block is last, and stores the NV totals into out[0..NV).  `scratch` holds gridDim.x*NV doubles.
struct ReduceScratch {
    double *partials;
    unsigned int *counter;
};

// Returns true in every thread of the LAST block; there thread 0 holds the NV totals in v[].
// All NV values travel through one shared-memory exchange (two barriers per phase instead of two per value): warp sums,
// then thread k adds the eight warp partials of value k in warp order -- the same additions in the same order as
// block_sum, so results are bit-identical to the one-value-at-a-time form.
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], ReduceScratch rs) {
    static_assert(NV <= 32, "grid_reduce: at most 32 values");
    __shared__ double sm[NV][kBlock / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const double t = warp_sum(v[k]);
        if (lane == 0) sm[k][w] = t;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < kBlock / 32; ++q) t += sm[threadIdx.x][q];
        rs.partials[(size_t)blockIdx.x * NV + threadIdx.x] = t;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int ticket = atomicAdd(rs.counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = 0.0;
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += kBlock) t += __ldcg(rs.partials + (size_t)b * NV + k);
        t = warp_sum(t);
        if (lane == 0) sm[k][w] = t;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < kBlock / 32; ++q) t += sm[threadIdx.x][q];
        sm[threadIdx.x][0] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] = sm[k][0];
        *rs.counter = 0u;
    }
    return true;
}

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

#endif  // __CUDACC__

}  // namespace lb2
