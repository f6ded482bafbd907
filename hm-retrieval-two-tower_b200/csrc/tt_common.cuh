// tt_common.cuh -- shared helpers for libtt.so (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "tt.h"

namespace tt {

void set_error(const char* fmt, ...);
void count_launch();

#define TT_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            tt::set_error(__VA_ARGS__);       \
            return TT_ERR_ARG;                \
        }                                     \
    } while (0)

#define TT_CUDA_OK(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            tt::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return TT_ERR_CUDA;                                                               \
        }                                                                                     \
    } while (0)

// after a kernel launch: surfaces launch-configuration errors without synchronising
#define TT_LAUNCH_OK(name)                                                                    \
    do {                                                                                      \
        cudaError_t _e = cudaPeekAtLastError();                                               \
        if (_e != cudaSuccess) {                                                              \
            tt::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));           \
            (void)cudaGetLastError();                                                         \
            return TT_ERR_CUDA;                                                               \
        }                                                                                     \
        tt::count_launch();                                                                   \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

// cudaFuncAttributeMaxDynamicSharedMemorySize of one kernel, set once per device and raised when a launch needs more.  Lock-free
// and safe from any number of host threads and devices (a lost race only repeats the idempotent driver call).  Declare one
// `static SmemAttr` next to each launch site.
struct SmemAttr {
    static constexpr int kMaxDevices = 32;
    std::atomic<int> cur[kMaxDevices];
    SmemAttr() { for (auto& c : cur) c.store(0, std::memory_order_relaxed); }
    template <typename F>
    cudaError_t ensure(F* kernel, int bytes) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        const bool tracked = dev >= 0 && dev < kMaxDevices;
        if (tracked && cur[dev].load(std::memory_order_acquire) >= bytes) return cudaSuccess;
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess && tracked) {
            int old = cur[dev].load(std::memory_order_relaxed);
            while (old < bytes && !cur[dev].compare_exchange_weak(old, bytes, std::memory_order_release)) {}
        }
        return e;
    }
};

// carve a caller-provided workspace
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(reinterpret_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return r;
    }
};

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// first element of embedding row `id` of a feature's table.  Row-sharded tables (shards = G > 1): `table` is a device array of G
// shard base pointers; row i lives in shard i % G at local row i / G, possibly in another GPU's HBM (peer-mapped, read over NVLink).
__device__ __forceinline__ const float* feature_row(const tt_feature& ft, int id) {
    if (ft.shards > 1) {
        const unsigned long long* tabs = reinterpret_cast<const unsigned long long*>(ft.table);
        const unsigned g = (unsigned)ft.shards, q = (unsigned)id / g;
        return reinterpret_cast<const float*>(__ldg(tabs + ((unsigned)id - q * g))) + (int64_t)q * ft.e;
    }
    return ft.table + (int64_t)id * ft.e;
}

// strict ordering of the tf.math.top_k contract: higher score first, then lower index
__device__ __forceinline__ bool ranks_before(float sa, int32_t ia, float sb, int32_t ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace tt
