// tt_api.cu -- library plumbing and the small element-wise entry points of tt.h.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "tt_common.cuh"

namespace tt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return 148;
    }
    return cached;
}

// ---- kernels ----------------------------------------------------------------------------------
__global__ void fill_kernel(float* __restrict__ p, float v, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

__global__ void log_kernel(const float* __restrict__ p, float* __restrict__ out, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = logf(p[i]);
}

__global__ void logq_apply_kernel(const float* __restrict__ L, int ldl, const float* __restrict__ bias, int Bq, int Bc,
                                  float* __restrict__ Z, int ldz) {
    int64_t total = (int64_t)Bq * Bc;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int r = (int)(i / Bc), c = (int)(i % Bc);
        float b = bias ? bias[c] : 0.0f;
        Z[(int64_t)r * ldz + c] = __fsub_rn(L[(int64_t)r * ldl + c], b);
    }
}

__global__ void round_tf32_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int64_t rows,
                                  int cols) {
    int64_t total = rows * cols;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < total; i += stride) {
        int64_t r = i / cols;
        int c = (int)(i % cols);
        dst[r * ldd + c] = tf32_rn(src[r * lds + c]);
    }
}

// ResourceApplyAdagradV2: acc += g*g; w -= (g*lr)/(sqrt(acc)+eps).  Every op individually rounded
// (no FMA contraction) so the result is bit-identical to the numpy oracle.
__global__ void dense_adagrad_kernel(float* __restrict__ w, float* __restrict__ acc, const float* __restrict__ g, int64_t n,
                                     float lr, float eps) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float gi = g[i];
        float a = __fadd_rn(acc[i], __fmul_rn(gi, gi));
        acc[i] = a;
        w[i] = __fsub_rn(w[i], __fdiv_rn(__fmul_rn(gi, lr), __fadd_rn(__fsqrt_rn(a), eps)));
    }
}

// ResourceApplyAdam: m += (g-m)(1-b1); v += (g*g-v)(1-b2); w -= (m*lr_t)/(sqrt(v)+eps)
__global__ void dense_adam_kernel(float* __restrict__ w, float* __restrict__ m, float* __restrict__ v,
                                  const float* __restrict__ g, int64_t n, float lr_t, float omb1, float omb2, float eps) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float gi = g[i];
        float mi = __fadd_rn(m[i], __fmul_rn(__fsub_rn(gi, m[i]), omb1));
        float vi = __fadd_rn(v[i], __fmul_rn(__fsub_rn(__fmul_rn(gi, gi), v[i]), omb2));
        m[i] = mi;
        v[i] = vi;
        w[i] = __fsub_rn(w[i], __fdiv_rn(__fmul_rn(mi, lr_t), __fadd_rn(__fsqrt_rn(vi), eps)));
    }
}

// hits[t] += #{(b, j<ks[t]) : cand[b][j] == true[b]} -- integer, exact; one block-level reduction and
// a single integer atomicAdd per (block, k) (integer adds commute, so the result is deterministic).
struct KsArg {
    int32_t ks[TT_MAX_KS];
};

__global__ void recall_hits_kernel(const int32_t* __restrict__ cand, int k_stride, const int32_t* __restrict__ truth, int nq,
                                   KsArg ks, int nk, int32_t* __restrict__ hits) {
    __shared__ int32_t s_hits[TT_MAX_KS];
    if (threadIdx.x < TT_MAX_KS) s_hits[threadIdx.x] = 0;
    __syncthreads();
    int32_t local[TT_MAX_KS];
#pragma unroll
    for (int t = 0; t < TT_MAX_KS; ++t) local[t] = 0;
    // one warp per query row; lanes stride over the k_stride candidates
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int b = warp; b < nq; b += nwarps) {
        int32_t t_id = truth[b];
        for (int j = lane; j < k_stride; j += 32) {
            if (cand[(int64_t)b * k_stride + j] == t_id) {
#pragma unroll
                for (int t = 0; t < TT_MAX_KS; ++t)
                    if (t < nk && j < ks.ks[t]) local[t]++;
            }
        }
    }
#pragma unroll
    for (int t = 0; t < TT_MAX_KS; ++t) {
        if (t < nk) {
            int32_t v = local[t];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0 && v) atomicAdd(&s_hits[t], v);
        }
    }
    __syncthreads();
    if (threadIdx.x < nk && s_hits[threadIdx.x]) atomicAdd(&hits[threadIdx.x], s_hits[threadIdx.x]);
}

__global__ void take_i32_kernel(const int32_t* __restrict__ table, const int32_t* __restrict__ idx, int64_t n, int32_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int32_t j = idx[i];
        out[i] = __ldg(table + (j > 0 ? j : 0));
    }
}

struct StageArgs {
    tt_stage_col col[TT_MAX_STAGE_COLS];
};

// blockIdx.y = column; 16-byte accesses when both ends allow it (device buffers and pinned host memory alike)
__global__ void __launch_bounds__(256) stage_columns_kernel(const __grid_constant__ StageArgs a, int64_t rows) {
    const tt_stage_col c = a.col[blockIdx.y];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c.kind == 0) {
        const bool vec = ((reinterpret_cast<uintptr_t>(c.src) | reinterpret_cast<uintptr_t>(c.dst)) & 15) == 0;
        const int64_t n4 = vec ? rows / 4 : 0;
        const uint4* s4 = reinterpret_cast<const uint4*>(c.src);
        uint4* d4 = reinterpret_cast<uint4*>(c.dst);
        for (int64_t j = i; j < n4; j += stride) d4[j] = s4[j];
        const uint32_t* s1 = reinterpret_cast<const uint32_t*>(c.src);
        uint32_t* d1 = reinterpret_cast<uint32_t*>(c.dst);
        for (int64_t j = n4 * 4 + i; j < rows; j += stride) d1[j] = s1[j];
    } else {
        const long long* s8 = reinterpret_cast<const long long*>(c.src);
        int32_t* d1 = reinterpret_cast<int32_t*>(c.dst);
        for (int64_t j = i; j < rows; j += stride) d1[j] = (int32_t)s8[j];
    }
}

// value (global row r, column c) of the table initialiser: a hash of (seed, r * e + c) only, so a shard that holds rows
// {row0 + i * row_stride} gets exactly the values the whole table would hold there
__device__ __forceinline__ float unit_hash(uint64_t seed, uint64_t ctr) {
    uint64_t z = ctr * 0x9E3779B97F4A7C15ull + seed;      // splitmix64 finaliser over a Weyl sequence
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);   // 24 random bits -> [0, 1)
}

__global__ void fill_uniform_kernel(float* __restrict__ out, int64_t rows_local, int e, int64_t row0, int64_t row_stride, uint64_t seed,
                                    float lo, float span) {
    const int64_t n = rows_local * e;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t r = i / e;
        const int c = (int)(i - r * e);
        out[i] = fmaf(unit_hash(seed, (uint64_t)((row0 + r * row_stride) * e + c)), span, lo);
    }
}

static inline int grid_for(int64_t n, int block) {
    int64_t g = ceil_div(n, block);
    int64_t cap = (int64_t)sm_count() * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace tt

using namespace tt;

extern "C" {

int tt_version(void) { return 100; }

const char* tt_last_error(void) { return tt::g_err; }

int64_t tt_launch_count(void) { return tt::g_launches.load(); }

// ---- peer-shareable memory (CUDA IPC; see tt.h) ---------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == TT_PEER_HANDLE_BYTES, "CUDA IPC handle size");

int tt_peer_alloc(size_t bytes, void** ptr, void* handle) {
    TT_REQUIRE(ptr != nullptr && handle != nullptr && bytes > 0, "tt_peer_alloc: bad arguments");
    void* p = nullptr;
    TT_CUDA_OK(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        tt::set_error("tt_peer_alloc: %s", cudaGetErrorString(e));
        return TT_ERR_CUDA;
    }
    memcpy(handle, &h, sizeof(h));
    *ptr = p;
    return TT_OK;
}

int tt_peer_open(const void* handle, void** ptr) {
    TT_REQUIRE(ptr != nullptr && handle != nullptr, "tt_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    TT_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TT_OK;
}

int tt_peer_close(void* ptr) {
    if (ptr) TT_CUDA_OK(cudaIpcCloseMemHandle(ptr));
    return TT_OK;
}

int tt_peer_free(void* ptr) {
    if (ptr) TT_CUDA_OK(cudaFree(ptr));
    return TT_OK;
}

int tt_device_supports_tc(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

int tt_fill_f32(float* p, float value, int64_t n, void* stream) {
    TT_REQUIRE(p != nullptr || n == 0, "tt_fill_f32: null pointer");
    if (n == 0) return TT_OK;
    fill_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(p, value, n);
    TT_LAUNCH_OK("fill_kernel");
    return TT_OK;
}

int tt_log_f32(const float* p, float* out, int64_t n, void* stream) {
    TT_REQUIRE((p && out) || n == 0, "tt_log_f32: null pointer");
    if (n == 0) return TT_OK;
    log_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(p, out, n);
    TT_LAUNCH_OK("log_kernel");
    return TT_OK;
}

int tt_logq_apply(const float* logits, int ldl, const float* col_bias, int Bq, int Bc, float* Z, int ldz, void* stream) {
    TT_REQUIRE(logits && Z, "tt_logq_apply: null pointer");
    TT_REQUIRE(Bq >= 0 && Bc >= 0 && ldl >= Bc && ldz >= Bc, "tt_logq_apply: bad shape");
    if ((int64_t)Bq * Bc == 0) return TT_OK;
    logq_apply_kernel<<<grid_for((int64_t)Bq * Bc, 256), 256, 0, as_stream(stream)>>>(logits, ldl, col_bias, Bq, Bc, Z, ldz);
    TT_LAUNCH_OK("logq_apply_kernel");
    return TT_OK;
}

int tt_round_tf32(const float* src, int lds, float* dst, int ldd, int64_t rows, int cols, void* stream) {
    TT_REQUIRE(src && dst, "tt_round_tf32: null pointer");
    TT_REQUIRE(rows >= 0 && cols >= 0 && lds >= cols && ldd >= cols, "tt_round_tf32: bad shape");
    if (rows * cols == 0) return TT_OK;
    round_tf32_kernel<<<grid_for(rows * cols, 256), 256, 0, as_stream(stream)>>>(src, lds, dst, ldd, rows, cols);
    TT_LAUNCH_OK("round_tf32_kernel");
    return TT_OK;
}

int tt_dense_adagrad(float* w, float* acc, const float* g, int64_t n, float lr, float eps, void* stream) {
    TT_REQUIRE((w && acc && g) || n == 0, "tt_dense_adagrad: null pointer");
    if (n == 0) return TT_OK;
    dense_adagrad_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(w, acc, g, n, lr, eps);
    TT_LAUNCH_OK("dense_adagrad_kernel");
    return TT_OK;
}

int tt_dense_adam(float* w, float* m, float* v, const float* g, int64_t n, float lr_t, float beta1, float beta2, float eps,
                  void* stream) {
    TT_REQUIRE((w && m && v && g) || n == 0, "tt_dense_adam: null pointer");
    if (n == 0) return TT_OK;
    dense_adam_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(w, m, v, g, n, lr_t, 1.0f - beta1, 1.0f - beta2, eps);
    TT_LAUNCH_OK("dense_adam_kernel");
    return TT_OK;
}

int tt_take_i32(const int32_t* table, const int32_t* idx, int64_t n, int32_t* out, void* stream) {
    TT_REQUIRE((table && idx && out) || n == 0, "tt_take_i32: null pointer");
    if (n == 0) return TT_OK;
    take_i32_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(table, idx, n, out);
    TT_LAUNCH_OK("take_i32_kernel");
    return TT_OK;
}

int tt_stage_columns(const tt_stage_col* cols, int n_cols, int64_t rows, void* stream) {
    TT_REQUIRE(n_cols >= 0 && n_cols <= TT_MAX_STAGE_COLS && rows >= 0, "tt_stage_columns: at most %d columns", TT_MAX_STAGE_COLS);
    if (n_cols == 0 || rows == 0) return TT_OK;
    TT_REQUIRE(cols != nullptr, "tt_stage_columns: null pointer");
    StageArgs a;
    for (int i = 0; i < n_cols; ++i) {
        TT_REQUIRE(cols[i].src && cols[i].dst && (cols[i].kind == 0 || cols[i].kind == 1), "tt_stage_columns: column %d: null pointer or bad kind", i);
        a.col[i] = cols[i];
    }
    int64_t gx = ceil_div(rows, 4 * 256);
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    stage_columns_kernel<<<dim3((unsigned)gx, (unsigned)n_cols), 256, 0, as_stream(stream)>>>(a, rows);
    TT_LAUNCH_OK("stage_columns_kernel");
    return TT_OK;
}

int tt_fill_uniform(float* out, int64_t rows_local, int e, int64_t row0, int64_t row_stride, uint64_t seed, float lo, float hi, void* stream) {
    TT_REQUIRE(out || rows_local == 0, "tt_fill_uniform: null pointer");
    TT_REQUIRE(rows_local >= 0 && e >= 1 && row0 >= 0 && row_stride >= 1 && hi >= lo, "tt_fill_uniform: bad arguments");
    if (rows_local == 0) return TT_OK;
    fill_uniform_kernel<<<grid_for(rows_local * e, 256), 256, 0, as_stream(stream)>>>(out, rows_local, e, row0, row_stride, seed, lo, hi - lo);
    TT_LAUNCH_OK("fill_uniform_kernel");
    return TT_OK;
}

int tt_recall_hits(const int32_t* cand, int k_stride, const int32_t* true_idx, int nq, const int32_t* ks, int nk,
                   int32_t* hits, void* stream) {
    TT_REQUIRE(cand && true_idx && ks && hits, "tt_recall_hits: null pointer");
    TT_REQUIRE(nk >= 1 && nk <= TT_MAX_KS, "tt_recall_hits: nk must be in [1,%d]", TT_MAX_KS);
    TT_REQUIRE(nq >= 0 && k_stride >= 1, "tt_recall_hits: bad shape");
    if (nq == 0) return TT_OK;
    KsArg a;
    for (int t = 0; t < TT_MAX_KS; ++t) a.ks[t] = t < nk ? ks[t] : 0;
    int block = 256;
    int grid = (int)ceil_div((int64_t)nq * 32, block);
    int cap = sm_count() * 4;
    if (grid > cap) grid = cap;
    recall_hits_kernel<<<grid, block, 0, as_stream(stream)>>>(cand, k_stride, true_idx, nq, a, nk, hits);
    TT_LAUNCH_OK("recall_hits_kernel");
    return TT_OK;
}

}  // extern "C"
