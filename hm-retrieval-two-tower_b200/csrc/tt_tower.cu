// tt_tower.cu -- InputLayer gather/concat and the Dense(relu) tower, forward and backward.
//
// Replaces (reference file:line): input_layer.py:37-41,61-68 (gather + concat),
// tower.py:41-49,72-75 (Dense/relu stack) and their autodiff (two_tower_model.py:110-124).
//
// The tower contractions are tiny next to the B x B logits, so they run as exact fp32
// FMA on the CUDA cores in the canonical k-ascending order: tower outputs are bit-identical to
// oracle/tt_oracle.c:tto_dense_fmaf and feed both the exact and the tensor-core logits paths.
#include "tt_common.cuh"
#include "tt_simt_gemm.cuh"

namespace tt {

struct FeatArr {
    tt_feature f[TT_MAX_FEATURES];
    int n;
};

constexpr int kMaxGatherCols = 1024;

// ------------------------------------------------------------------------------------------------
// G1: stand-alone gather + concat (pure HBM traffic).  Vector path: every block is float4-aligned.
// One thread per 16-byte chunk of the output row; consecutive threads write consecutive chunks, so
// stores are fully coalesced and each table row is read with 128-bit loads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gather_concat_vec_kernel(const __grid_constant__ FeatArr fa, int B, int D4, int ld4,
                                                                float4* __restrict__ X) {
    __shared__ uint8_t s_feat[kMaxGatherCols / 4];
    __shared__ uint16_t s_off[kMaxGatherCols / 4];
    for (int c = threadIdx.x; c < ld4; c += blockDim.x) {
        uint8_t ff = 255;
        uint16_t oo = 0;
        for (int f = 0; f < fa.n; ++f) {
            int c0 = fa.f[f].col >> 2, c1 = (fa.f[f].col + fa.f[f].e) >> 2;
            if (c >= c0 && c < c1) { ff = (uint8_t)f; oo = (uint16_t)(c - c0); }
        }
        s_feat[c] = ff;
        s_off[c] = oo;
    }
    __syncthreads();
    const int64_t total = (int64_t)B * ld4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
        int b = (int)(t / ld4), c = (int)(t % ld4);
        int f = s_feat[c];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f != 255) {
            const tt_feature& ft = fa.f[f];
            int id = __ldg(reinterpret_cast<const int32_t*>(ft.src) + b);
            if ((unsigned)id >= (unsigned)ft.rows) id = 0;
            v = __ldg(reinterpret_cast<const float4*>(feature_row(ft, id)) + s_off[c]);
        }
        X[t] = v;
    }
}

// scalar path (numeric features or blocks that are not 16-byte aligned)
__global__ void __launch_bounds__(256) gather_concat_scalar_kernel(const __grid_constant__ FeatArr fa, int B, int D, int ldx,
                                                                   float* __restrict__ X) {
    __shared__ uint8_t s_feat[kMaxGatherCols];
    __shared__ uint16_t s_off[kMaxGatherCols];
    for (int c = threadIdx.x; c < ldx; c += blockDim.x) {
        uint8_t ff = 255;
        uint16_t oo = 0;
        for (int f = 0; f < fa.n; ++f)
            if (c >= fa.f[f].col && c < fa.f[f].col + fa.f[f].e) { ff = (uint8_t)f; oo = (uint16_t)(c - fa.f[f].col); }
        s_feat[c] = ff;
        s_off[c] = oo;
    }
    __syncthreads();
    const int64_t total = (int64_t)B * ldx;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
        int b = (int)(t / ldx), c = (int)(t % ldx);
        int f = s_feat[c];
        float v = 0.f;
        if (f != 255) {
            const tt_feature& ft = fa.f[f];
            if (ft.table == nullptr) {
                v = __ldg(reinterpret_cast<const float*>(ft.src) + b);
            } else {
                int id = __ldg(reinterpret_cast<const int32_t*>(ft.src) + b);
                if ((unsigned)id >= (unsigned)ft.rows) id = 0;
                v = __ldg(feature_row(ft, id) + s_off[c]);
            }
        }
        X[t] = v;
    }
}

// ---- forward -----------------------------------------------------------------------------------
struct DenseA {
    const float* X;
    int ldx, B, K;
    __device__ __forceinline__ float operator()(int m, int k) const { return (m < B && k < K) ? __ldg(X + (int64_t)m * ldx + k) : 0.f; }
};
struct DenseW {
    const float* W;
    int K, N;
    __device__ __forceinline__ float operator()(int k, int n) const { return (k < K && n < N) ? __ldg(W + (int64_t)k * N + n) : 0.f; }
};

template <bool kGather>
__global__ void __launch_bounds__(256) dense_fwd_kernel(const __grid_constant__ FeatArr fa, const float* __restrict__ X, int ldx,
                                                        const float* __restrict__ W, const float* __restrict__ bias,
                                                        float* __restrict__ Xout, float* __restrict__ Y, int ldy,
                                                        float* __restrict__ Ytf32, int B, int K, int N, int relu) {
    __shared__ TileSmem sm;
    __shared__ int32_t s_ids[kGather ? BM * TT_MAX_FEATURES : 1];
    __shared__ uint8_t s_feat[kGather ? kMaxGatherCols : 1];
    __shared__ uint16_t s_off[kGather ? kMaxGatherCols : 1];
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    float acc[TM][TN] = {};
    if constexpr (kGather) {
        for (int c = threadIdx.x; c < K; c += blockDim.x) {
            uint8_t ff = 255;
            uint16_t oo = 0;
            for (int f = 0; f < fa.n; ++f)
                if (c >= fa.f[f].col && c < fa.f[f].col + fa.f[f].e) { ff = (uint8_t)f; oo = (uint16_t)(c - fa.f[f].col); }
            s_feat[c] = ff;
            s_off[c] = oo;
        }
        for (int i = threadIdx.x; i < BM * fa.n; i += blockDim.x) {
            int r = i / fa.n, f = i % fa.n;
            int id = 0;
            if (m0 + r < B && fa.f[f].table != nullptr) {
                id = __ldg(reinterpret_cast<const int32_t*>(fa.f[f].src) + m0 + r);
                if ((unsigned)id >= (unsigned)fa.f[f].rows) id = 0;
            }
            s_ids[r * TT_MAX_FEATURES + f] = id;
        }
        __syncthreads();
        const bool write_x = (Xout != nullptr) && blockIdx.y == 0;
        auto la = [&](int m, int k) -> float {
            if (m >= B || k >= K) return 0.f;
            int f = s_feat[k];
            float v = 0.f;
            if (f != 255) {
                const tt_feature& ft = fa.f[f];
                if (ft.table == nullptr) v = __ldg(reinterpret_cast<const float*>(ft.src) + m);
                else v = __ldg(feature_row(ft, s_ids[(m - m0) * TT_MAX_FEATURES + f]) + s_off[k]);
            }
            if (write_x) Xout[(int64_t)m * ldx + k] = v;
            return v;
        };
        DenseW lb{W, K, N};
        tile_gemm<true, true>(acc, la, lb, m0, n0, 0, K, sm);
        if (write_x) {  // zero the padding columns [K, ldx)
            for (int i = threadIdx.x; i < BM * (ldx - K); i += blockDim.x) {
                int r = i / (ldx - K), c = K + i % (ldx - K);
                if (m0 + r < B) Xout[(int64_t)(m0 + r) * ldx + c] = 0.f;
            }
        }
    } else {
        DenseA la{X, ldx, B, K};
        DenseW lb{W, K, N};
        tile_gemm<true, true>(acc, la, lb, m0, n0, 0, K, sm);
    }
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= B) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n >= N) continue;
            float y = acc[i][j];
            if (bias) y = __fadd_rn(y, __ldg(bias + n));
            if (relu && !(y > 0.f)) y = 0.f;
            Y[(int64_t)m * ldy + n] = y;
            if (Ytf32) Ytf32[(int64_t)m * ldy + n] = tf32_rn(y);
        }
    }
}

// ---- backward ----------------------------------------------------------------------------------
// dpre(b, n) = dY[b][n] * (relu ? Y[b][n] > 0 : 1)
struct DPre {
    const float* dY;
    const float* Y;
    int lddy, ldy, B, N, relu;
    __device__ __forceinline__ float at(int b, int n) const {
        if (b >= B || n >= N) return 0.f;
        float g = __ldg(dY + (int64_t)b * lddy + n);
        if (relu && !(__ldg(Y + (int64_t)b * ldy + n) > 0.f)) g = 0.f;
        return g;
    }
};

// dX[b][k] = sum_n dpre[b][n] * W[k][n]
__global__ void __launch_bounds__(256) dense_bwd_dx_kernel(DPre dp, const float* __restrict__ W, float* __restrict__ dX, int lddx,
                                                           int B, int K, int N) {
    __shared__ TileSmem sm;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;  // output tile: rows b, cols k
    float acc[TM][TN] = {};
    auto la = [&](int m, int kk) -> float { return dp.at(m, kk); };                                            // (b, n)
    auto lb = [&](int kk, int n) -> float { return (kk < N && n < K) ? __ldg(W + (int64_t)n * N + kk) : 0.f; };  // W^T
    tile_gemm<true, false>(acc, la, lb, m0, n0, 0, N, sm);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= B) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n < K) dX[(int64_t)m * lddx + n] = acc[i][j];
        }
    }
}

// partial[z][k][n] = sum_{b in chunk z} X[b][k] * dpre[b][n];   row k == K holds the bias gradient
__global__ void __launch_bounds__(256) dense_bwd_dw_kernel(const float* __restrict__ X, int ldx, DPre dp, float* __restrict__ partial,
                                                           int B, int K, int N, int chunk) {
    __shared__ TileSmem sm;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;  // output tile: rows k (K+1 of them), cols n
    const int b0 = blockIdx.z * chunk;
    const int b1 = min(B, b0 + chunk);
    float acc[TM][TN] = {};
    auto la = [&](int m, int bb) -> float {
        if (bb >= B) return 0.f;
        if (m < K) return __ldg(X + (int64_t)bb * ldx + m);
        return m == K ? 1.0f : 0.f;
    };
    auto lb = [&](int bb, int n) -> float { return dp.at(bb, n); };
    tile_gemm<false, true>(acc, la, lb, m0, n0, b0, b1, sm);
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    float* out = partial + (int64_t)blockIdx.z * (K + 1) * N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m > K) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n < N) out[(int64_t)m * N + n] = acc[i][j];
        }
    }
}

// dW / db = sum over chunks in ascending chunk order (fixed order => deterministic)
__global__ void dense_bwd_reduce_kernel(const float* __restrict__ partial, int nchunk, int K, int N, float* __restrict__ dW,
                                        float* __restrict__ db) {
    int64_t total = (int64_t)(K + 1) * N;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= total) return;
    float s = 0.f;
#pragma unroll 16
    for (int z = 0; z < nchunk; ++z) s = __fadd_rn(s, __ldg(partial + (int64_t)z * total + i));   // loads are independent: keep 16 in flight
    int k = (int)(i / N), n = (int)(i % N);
    if (k < K) dW[(int64_t)k * N + n] = s;
    else if (db) db[n] = s;
}

static int dw_chunks(int B, int K, int N, int* chunk_rows) {
    int64_t tiles = ceil_div(K + 1, BM) * ceil_div(N, BN);
    int64_t want = ceil_div(2 * (int64_t)sm_count(), tiles);
    int64_t maxc = ceil_div(B, 128);
    int64_t nchunk = want < maxc ? want : maxc;
    if (nchunk < 1) nchunk = 1;
    int64_t rows = ceil_div(ceil_div(B, nchunk), BK) * BK;
    if (rows < BK) rows = BK;
    nchunk = ceil_div(B, rows);
    if (nchunk < 1) nchunk = 1;
    *chunk_rows = (int)rows;
    return (int)nchunk;
}

static int check_feats(const tt_feature* feats, int nfeat, int D, const char* who) {
    TT_REQUIRE(feats != nullptr && nfeat >= 1 && nfeat <= TT_MAX_FEATURES, "%s: nfeat must be in [1,%d]", who, TT_MAX_FEATURES);
    TT_REQUIRE(D >= 1 && D <= kMaxGatherCols, "%s: D must be in [1,%d]", who, kMaxGatherCols);
    for (int f = 0; f < nfeat; ++f) {
        TT_REQUIRE(feats[f].src != nullptr, "%s: feature %d has no source", who, f);
        TT_REQUIRE(feats[f].e >= 1 && feats[f].col >= 0 && feats[f].col + feats[f].e <= D, "%s: feature %d block out of range", who, f);
        TT_REQUIRE(feats[f].table == nullptr ? feats[f].e == 1 : feats[f].rows >= 1, "%s: feature %d malformed", who, f);
    }
    return TT_OK;
}

// tt_tower_panel.cu: shared-memory resident variants for K <= 256 / N <= 256
bool panel_fwd_ok(int K, int N);
bool panel_bwd_ok(int K, int N);
int panel_dense_fwd(const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, float* Y32, int B, int K, int N, int relu,
                    cudaStream_t st);
int panel_input_dense_fwd(const tt_feature* feats, int nfeat, int D, const float* W, const float* b, float* Xout, int ldx, float* Y, int ldy,
                          float* Y32, int B, int N, int relu, cudaStream_t st);
size_t panel_bwd_workspace(int B, int K, int N);
int panel_dense_bwd(const float* X, int ldx, const float* W, const float* Y, int ldy, const float* dY, int lddy, float* dX, int lddx, float* dW,
                    float* db, float* partial, int B, int K, int N, int relu, cudaStream_t st);

}  // namespace tt

using namespace tt;

extern "C" {

int tt_gather_concat(const tt_feature* feats, int nfeat, int B, int D, float* X, int ldx, void* stream) {
    if (B == 0) return TT_OK;
    int rc = check_feats(feats, nfeat, D, "tt_gather_concat");
    if (rc) return rc;
    TT_REQUIRE(X != nullptr && B >= 0 && ldx >= D && ldx <= kMaxGatherCols, "tt_gather_concat: bad output shape");
    if (B == 0) return TT_OK;
    FeatArr fa;
    memset(&fa, 0, sizeof(fa));
    fa.n = nfeat;
    bool vec = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
    for (int f = 0; f < nfeat; ++f) {
        fa.f[f] = feats[f];
        if (feats[f].table == nullptr || feats[f].e % 4 || feats[f].col % 4 || (reinterpret_cast<uintptr_t>(feats[f].table) & 15)) vec = false;
    }
    cudaStream_t st = as_stream(stream);
    if (vec) {
        int ld4 = ldx / 4;
        int64_t total = (int64_t)B * ld4;
        int64_t grid = ceil_div(total, 256 * 4);
        int64_t cap = (int64_t)sm_count() * 8;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        gather_concat_vec_kernel<<<(int)grid, 256, 0, st>>>(fa, B, D / 4, ld4, reinterpret_cast<float4*>(X));
        TT_LAUNCH_OK("gather_concat_vec_kernel");
    } else {
        int64_t total = (int64_t)B * ldx;
        int64_t grid = ceil_div(total, 256 * 4);
        int64_t cap = (int64_t)sm_count() * 8;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        gather_concat_scalar_kernel<<<(int)grid, 256, 0, st>>>(fa, B, D, ldx, X);
        TT_LAUNCH_OK("gather_concat_scalar_kernel");
    }
    return TT_OK;
}

int tt_dense_fwd(const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, float* Y_tf32, int B, int K, int N,
                 int relu, void* stream) {
    TT_REQUIRE(X && W && Y, "tt_dense_fwd: null pointer");
    TT_REQUIRE(B >= 0 && K >= 1 && N >= 1 && ldx >= K && ldy >= N, "tt_dense_fwd: bad shape");
    if (B == 0) return TT_OK;
    if (panel_fwd_ok(K, N)) return panel_dense_fwd(X, ldx, W, b, Y, ldy, Y_tf32, B, K, N, relu, as_stream(stream));
    FeatArr fa;
    memset(&fa, 0, sizeof(fa));
    dim3 grid((unsigned)ceil_div(B, BM), (unsigned)ceil_div(N, BN));
    dense_fwd_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(fa, X, ldx, W, b, nullptr, Y, ldy, Y_tf32, B, K, N, relu);
    TT_LAUNCH_OK("dense_fwd_kernel");
    return TT_OK;
}

int tt_input_dense_fwd(const tt_feature* feats, int nfeat, int D, const float* W, const float* b, float* X_out, int ldx, float* Y,
                       int ldy, float* Y_tf32, int B, int N, int relu, void* stream) {
    if (B == 0) return TT_OK;
    int rc = check_feats(feats, nfeat, D, "tt_input_dense_fwd");
    if (rc) return rc;
    TT_REQUIRE(W && Y, "tt_input_dense_fwd: null pointer");
    TT_REQUIRE(B >= 0 && N >= 1 && ldy >= N && (X_out == nullptr || ldx >= D), "tt_input_dense_fwd: bad shape");
    if (B == 0) return TT_OK;
    if (panel_fwd_ok(D, N)) return panel_input_dense_fwd(feats, nfeat, D, W, b, X_out, ldx, Y, ldy, Y_tf32, B, N, relu, as_stream(stream));
    FeatArr fa;
    memset(&fa, 0, sizeof(fa));
    fa.n = nfeat;
    for (int f = 0; f < nfeat; ++f) fa.f[f] = feats[f];
    dim3 grid((unsigned)ceil_div(B, BM), (unsigned)ceil_div(N, BN));
    dense_fwd_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(fa, nullptr, ldx, W, b, X_out, Y, ldy, Y_tf32, B, D, N, relu);
    TT_LAUNCH_OK("dense_fwd_kernel<gather>");
    return TT_OK;
}

size_t tt_dense_bwd_workspace_bytes(int B, int K, int N) {
    if (B <= 0 || K <= 0 || N <= 0) return 256;
    int rows = 0;
    int nchunk = dw_chunks(B, K, N, &rows);
    size_t generic = align_up((size_t)nchunk * (size_t)(K + 1) * (size_t)N * sizeof(float), 256) + 256;
    size_t panel = panel_bwd_ok(K, N) ? panel_bwd_workspace(B, K, N) : 0;
    return generic > panel ? generic : panel;
}

int tt_dense_bwd(const float* X, int ldx, const float* W, const float* Y, int ldy, const float* dY, int lddy, float* dX, int lddx,
                 float* dW, float* db, int B, int K, int N, int relu, void* ws, size_t ws_bytes, void* stream) {
    TT_REQUIRE(X && W && Y && dY && dW, "tt_dense_bwd: null pointer");
    TT_REQUIRE(B >= 0 && K >= 1 && N >= 1 && ldx >= K && ldy >= N && lddy >= N && (dX == nullptr || lddx >= K), "tt_dense_bwd: bad shape");
    TT_REQUIRE(ws != nullptr && ws_bytes >= tt_dense_bwd_workspace_bytes(B, K, N), "tt_dense_bwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    if (B == 0) {
        TT_CUDA_OK(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)K * N, st));
        if (db) TT_CUDA_OK(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, st));
        return TT_OK;
    }
    if (panel_bwd_ok(K, N))
        return panel_dense_bwd(X, ldx, W, Y, ldy, dY, lddy, dX, lddx, dW, db, reinterpret_cast<float*>(ws), B, K, N, relu, st);
    DPre dp{dY, Y, lddy, ldy, B, N, relu};
    if (dX) {
        dim3 grid((unsigned)ceil_div(B, BM), (unsigned)ceil_div(K, BN));
        dense_bwd_dx_kernel<<<grid, 256, 0, st>>>(dp, W, dX, lddx, B, K, N);
        TT_LAUNCH_OK("dense_bwd_dx_kernel");
    }
    int rows = 0;
    int nchunk = dw_chunks(B, K, N, &rows);
    float* partial = reinterpret_cast<float*>(ws);
    dim3 grid((unsigned)ceil_div(K + 1, BM), (unsigned)ceil_div(N, BN), (unsigned)nchunk);
    dense_bwd_dw_kernel<<<grid, 256, 0, st>>>(X, ldx, dp, partial, B, K, N, rows);
    TT_LAUNCH_OK("dense_bwd_dw_kernel");
    int64_t total = (int64_t)(K + 1) * N;
    dense_bwd_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(partial, nchunk, K, N, dW, db);
    TT_LAUNCH_OK("dense_bwd_reduce_kernel");
    return TT_OK;
}

}  // extern "C"
