// tt_index.cu -- brute-force index: exact fp32 scoring fused with top-K selection, K-way merge.
//
// Replaces (reference file:line): brute_force.py:75-78 (scores = Q.C^T), :81 (tf.math.top_k: sorted
// descending, lower index first on equal scores) -- the (nq x n) score matrix never reaches HBM.
//
// Exact path (TT_IMPL_SIMT): scores are the canonical k-ascending fmaf sums, bit-identical to
// oracle/tt_oracle.c:tto_index_topk.  Each CTA scans a slice of the corpus for 64 (or 8) queries and
// keeps, per query, a candidate list in shared memory guarded by a running threshold (the K-th best so
// far); lists are compacted by a warp-level bitonic sort only when they are about to overflow, so the
// expected number of insertions per query is ~K.ln(n/K) rather than n.
#include <math_constants.h>

#include "tt_common.cuh"
#include "tt_simt_gemm.cuh"

namespace tt {

constexpr int32_t kIdxPad = 0x7fffffff;

struct QRows {
    const float* P;
    int ld, row_end, cols;
    __device__ __forceinline__ float operator()(int m, int k) const { return (m < row_end && k < cols) ? __ldg(P + (int64_t)m * ld + k) : 0.f; }
};
struct CTrans {
    const float* P;
    int ld;
    int64_t row_end;
    int cols;
    __device__ __forceinline__ float operator()(int k, int n) const { return (n < row_end && k < cols) ? __ldg(P + (int64_t)n * ld + k) : 0.f; }
};

// warp-cooperative bitonic sort (best first) of CAP (score, idx) pairs living in shared memory
template <int CAP>
__device__ __forceinline__ void warp_bitonic_sort(float* s, int32_t* id, int lane) {
#pragma unroll 1
    for (int k = 2; k <= CAP; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < CAP / 2; t += 32) {
                int i = 2 * t - (t & (j - 1));
                int p = i + j;
                bool up = ((i & k) == 0);  // ascending in rank order: best first
                float si = s[i], sp = s[p];
                int32_t ii = id[i], ip = id[p];
                bool swap = up ? ranks_before(sp, ip, si, ii) : ranks_before(si, ii, sp, ip);
                if (swap) { s[i] = sp; s[p] = si; id[i] = ip; id[p] = ii; }
            }
            __syncwarp();
        }
    }
}

template <int BQ, int CAP>
struct IndexSmem {
    TileSmem tile;
    float bs[BQ][CAP];
    int32_t bi[BQ][CAP];
    int cnt[BQ];
    float thr[BQ];
};

// Work items (query tile, corpus split), item w = split * n_qtiles + tile; a CTA strides over them (the fallback launches a small
// grid: normally no tile is flagged and a CTA only reads a few flag words).  Partial results (split, nq, K), best first, padded
// (-inf, kIdxPad).
template <int BQ, int CAP>
__global__ void __launch_bounds__(256) index_exact_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ C, int ldc, int nq,
                                                          int64_t n, int E, int K, int64_t per_split, float* __restrict__ ps,
                                                          int32_t* __restrict__ pi, const int32_t* __restrict__ flags, int n_qtiles, int n_items) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IndexSmem<BQ, CAP>& sm = *reinterpret_cast<IndexSmem<BQ, CAP>*>(smem_raw);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int bx = item % n_qtiles, by = item / n_qtiles;
    const int q0 = bx * BQ;
    __syncthreads();   // the previous item's lists are done with
    if (flags) {   // fallback mode: only query tiles holding a flagged query do any work
        int mine = 0;
        for (int q = threadIdx.x; q < BQ; q += 256) mine |= (q0 + q < nq) ? flags[q0 + q] : 0;
        if (!__syncthreads_or(mine)) continue;
    }
    const int64_t c_begin = (int64_t)by * per_split;
    const int64_t c_end = min(n, c_begin + per_split);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ty = tid >> 4, tx = tid & 15;
    for (int q = tid; q < BQ; q += 256) { sm.cnt[q] = 0; sm.thr[q] = -CUDART_INF_F; }
    __syncthreads();
    QRows la{Q, ldq, min(nq, q0 + BQ), E};
    CTrans lb{C, ldc, c_end, E};
    for (int n0 = (int)c_begin; n0 < (int)c_end; n0 += BN) {
        float acc[TM][TN] = {};
        tile_gemm<true, false>(acc, la, lb, q0, n0, 0, E, sm.tile);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int ql = ty * TM + i;
            if (ql >= BQ || q0 + ql >= nq) continue;
            float th = sm.thr[ql];
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                int cn = n0 + tx * TN + j;
                if (cn < (int)c_end && acc[i][j] > th) {
                    int pos = atomicAdd(&sm.cnt[ql], 1);
                    sm.bs[ql][pos] = acc[i][j];
                    sm.bi[ql][pos] = cn;
                }
            }
        }
        __syncthreads();
        // compaction of lists that could overflow during the next tile (at most BN pushes per query)
        for (int ql = warp; ql < BQ; ql += 8) {
            int m = sm.cnt[ql];
            if (m > CAP - BN) {
                for (int t = m + lane; t < CAP; t += 32) { sm.bs[ql][t] = -CUDART_INF_F; sm.bi[ql][t] = kIdxPad; }
                __syncwarp();
                warp_bitonic_sort<CAP>(sm.bs[ql], sm.bi[ql], lane);
                if (lane == 0) {
                    sm.cnt[ql] = min(m, K);
                    sm.thr[ql] = (m >= K) ? sm.bs[ql][K - 1] : -CUDART_INF_F;
                }
            }
        }
        __syncthreads();
    }
    // final: sort every list and emit the best K
    for (int ql = warp; ql < BQ; ql += 8) {
        if (q0 + ql >= nq) continue;
        int m = sm.cnt[ql];
        for (int t = m + lane; t < CAP; t += 32) { sm.bs[ql][t] = -CUDART_INF_F; sm.bi[ql][t] = kIdxPad; }
        __syncwarp();
        warp_bitonic_sort<CAP>(sm.bs[ql], sm.bi[ql], lane);
        int64_t o = ((int64_t)by * nq + (q0 + ql)) * K;
        for (int t = lane; t < K; t += 32) { ps[o + t] = sm.bs[ql][t]; pi[o + t] = sm.bi[ql][t]; }
    }
    }
}

// block-wide bitonic sort of P pairs in shared memory, then write the best K (idx padded -> -1)
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ s_in, const int32_t* __restrict__ i_in, int G, int nq, int K,
                                                         int P, int64_t idx_add, float* __restrict__ s_out, int32_t* __restrict__ i_out,
                                                         const int32_t* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s = reinterpret_cast<float*>(smem_raw);
    int32_t* id = reinterpret_cast<int32_t*>(s + P);
    const int total = G * K;
    // fallback mode (flags): a small grid strides over the queries and leaves the rows the filter path already produced alone --
    // normally none is flagged, and 2 CTAs per SM check all flags in a few microseconds
    for (int q = blockIdx.x; q < nq; q += gridDim.x) {
    if (flags && !flags[q]) continue;
    for (int t = threadIdx.x; t < P; t += blockDim.x) {
        if (t < total) {
            int g = t / K, j = t - g * K;
            int64_t o = ((int64_t)g * nq + q) * K + j;
            float sv = s_in[o];
            int32_t iv = i_in[o];
            if (iv < 0) { iv = kIdxPad; sv = -CUDART_INF_F; }
            s[t] = sv;
            id[t] = iv;
        } else { s[t] = -CUDART_INF_F; id[t] = kIdxPad; }
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
                int i = 2 * t - (t & (j - 1));
                int p = i + j;
                bool up = ((i & k) == 0);
                float si = s[i], sp = s[p];
                int32_t ii = id[i], ip = id[p];
                bool swap = up ? ranks_before(sp, ip, si, ii) : ranks_before(si, ii, sp, ip);
                if (swap) { s[i] = sp; s[p] = si; id[i] = ip; id[p] = ii; }
            }
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < K; t += blockDim.x) {
        int32_t iv = id[t];
        bool pad = (iv == kIdxPad);
        s_out[(int64_t)q * K + t] = pad ? -CUDART_INF_F : s[t];
        i_out[(int64_t)q * K + t] = pad ? -1 : (int32_t)(iv + idx_add);
    }
    __syncthreads();
    }
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

struct ExactPlan {
    int bq, cap, nsplit;
    int64_t per_split;
    size_t smem;
};

// `fallback`: the kernel redoes the few queries the filter path flagged, so only a few query tiles do any work; the corpus is then cut
// into as many splits as the merge allows, so that even ONE flagged tile spreads over the chip (10 splits made a flagged tile cost
// 3.9 ms at 105 k rows: ten CTAs of work on 148 SMs)
static int plan_exact(int nq, int64_t n, int K, ExactPlan* pl, bool fallback = false) {
    if (K <= 128) { pl->bq = 64; pl->cap = 256; pl->smem = sizeof(IndexSmem<64, 256>); }
    else if (K <= 1536) { pl->bq = 8; pl->cap = 2048; pl->smem = sizeof(IndexSmem<8, 2048>); }
    else { set_error("tt_index_topk (exact): K=%d > 1536 unsupported", K); return TT_ERR_UNSUPPORTED; }
    int64_t qtiles = ceil_div(nq, pl->bq);
    int64_t want = ceil_div(2 * (int64_t)sm_count(), qtiles);
    int64_t by_work = ceil_div(n, 2048);             // at least 2048 candidates per split
    int64_t by_merge = 16384 / next_pow2(K);         // merge sorts nsplit*K entries in shared memory
    int64_t ns = fallback ? 64 : want;
    if (ns > by_work) ns = by_work;
    if (ns > by_merge) ns = by_merge;
    if (ns > 64) ns = 64;
    if (ns < 1) ns = 1;
    int64_t per = ceil_div(ceil_div(n, ns), BN) * BN;
    if (per < BN) per = BN;
    ns = ceil_div(n, per);
    if (ns < 1) ns = 1;
    pl->nsplit = (int)ns;
    pl->per_split = per;
    return TT_OK;
}

int merge_launch(const float* s_in, const int32_t* i_in, int G, int nq, int K, int64_t idx_add, float* s_out, int32_t* i_out,
                 cudaStream_t st, const int32_t* flags = nullptr) {
    int P = next_pow2(G * K);
    if (P < 2) P = 2;
    if (P > 16384) { set_error("tt_topk_merge: G*K=%d exceeds 16384", G * K); return TT_ERR_UNSUPPORTED; }
    size_t smem = (size_t)P * 8;
    { static SmemAttr smem_attr; TT_CUDA_OK(smem_attr.ensure(topk_merge_kernel, 16384 * 8)); }
    const int grid = (flags && nq > 2 * sm_count()) ? 2 * sm_count() : nq;
    topk_merge_kernel<<<(unsigned)grid, 256, smem, st>>>(s_in, i_in, G, nq, K, P, idx_add, s_out, i_out, flags);
    TT_LAUNCH_OK("topk_merge_kernel");
    return TT_OK;
}

size_t index_exact_workspace(int nq, int64_t n, int K) {
    ExactPlan pl, pf;
    if (plan_exact(nq, n, K, &pl) || plan_exact(nq, n, K, &pf, true)) return 0;
    const size_t ns = (size_t)(pl.nsplit > pf.nsplit ? pl.nsplit : pf.nsplit);      // sized for either use
    return align_up(ns * nq * K * sizeof(float), 256) + align_up(ns * nq * K * sizeof(int32_t), 256) + 256;
}

int index_exact(const float* Q, int ldq, const float* C, int ldc, int nq, int64_t n, int E, int K, int64_t idx_base, float* out_s,
                int32_t* out_i, void* ws, size_t ws_bytes, cudaStream_t st, const int32_t* flags) {
    ExactPlan pl;
    int rc = plan_exact(nq, n, K, &pl, flags != nullptr);
    if (rc) return rc;
    TT_REQUIRE(ws && ws_bytes >= index_exact_workspace(nq, n, K), "tt_index_topk: workspace too small");
    Carver cv(ws);
    float* ps = cv.take<float>((size_t)pl.nsplit * nq * K);
    int32_t* pi = cv.take<int32_t>((size_t)pl.nsplit * nq * K);
    const int n_qtiles = (int)ceil_div(nq, pl.bq);
    const int64_t items = (int64_t)n_qtiles * pl.nsplit;
    TT_REQUIRE(items < (1ll << 31), "tt_index_topk (exact): too many work items");
    const int64_t cap = flags ? 2 * (int64_t)sm_count() : items;      // fallback: a small grid strides over the (mostly unflagged) items
    const unsigned grid = (unsigned)(items < cap ? items : cap);
    if (pl.bq == 64) {
        { static SmemAttr smem_attr; TT_CUDA_OK(smem_attr.ensure(index_exact_kernel<64, 256>, (int)sizeof(IndexSmem<64, 256>))); }
        index_exact_kernel<64, 256><<<grid, 256, pl.smem, st>>>(Q, ldq, C, ldc, nq, n, E, K, pl.per_split, ps, pi, flags, n_qtiles, (int)items);
    } else {
        { static SmemAttr smem_attr; TT_CUDA_OK(smem_attr.ensure(index_exact_kernel<8, 2048>, (int)sizeof(IndexSmem<8, 2048>))); }
        index_exact_kernel<8, 2048><<<grid, 256, pl.smem, st>>>(Q, ldq, C, ldc, nq, n, E, K, pl.per_split, ps, pi, flags, n_qtiles, (int)items);
    }
    TT_LAUNCH_OK("index_exact_kernel");
    return merge_launch(ps, pi, pl.nsplit, nq, K, idx_base, out_s, out_i, st, flags);
}

// tt_index_tc.cu
bool index_tc_supported(int ldq, int ldc, int E, int K, int64_t n, const void* Q, const void* C);
size_t index_tc_workspace(int nq, int64_t n, int E, int K, bool need_corpus_copy);
int index_tc(const float* Q, int ldq, const float* C, int ldc, const float* C32, const float* norms, int nq, int64_t n, int E, int K,
             int64_t idx_base, float* out_s, int32_t* out_i, void* ws, size_t ws_bytes, cudaStream_t st);
namespace tc { int launch_prepare(const float* C, int ldc, int64_t n, int E, float* C32p, float* norms, int64_t n_pad, cudaStream_t st); }

}  // namespace tt

using namespace tt;

extern "C" {

size_t tt_index_workspace_bytes(int nq, int64_t n, int E, int K, int impl, int have_corpus_prepared) {
    if (nq <= 0 || n <= 0 || K <= 0) return 256;
    size_t a = index_exact_workspace(nq, n, K);
    size_t b = (impl == TT_IMPL_SIMT) ? 0 : index_tc_workspace(nq, n, E, K, !have_corpus_prepared);
    return a + b + 256;   // the filter path keeps the exact path's scratch for its (rare) fallback
}

size_t tt_index_prepared_bytes(int64_t n, int E) {
    if (n <= 0 || E <= 0) return 16;
    return (size_t)n * E * (E >= 64 ? 2 : 4);      // fp16 tiles from E = 64 (tt_index_tc.cu idx_half_operands), TF32-rounded fp32 below
}

int tt_index_prepare(const float* corpus, int ldc, int64_t n, int E, void* corpus_prepared, float* corpus_norms, void* stream) {
    TT_REQUIRE(corpus && corpus_prepared && corpus_norms, "tt_index_prepare: null pointer");
    TT_REQUIRE(n >= 0 && E >= 1 && ldc >= E, "tt_index_prepare: bad shape");
    TT_REQUIRE((reinterpret_cast<uintptr_t>(corpus_prepared) & 15) == 0, "tt_index_prepare: corpus_prepared must be 16-byte aligned");
    return tc::launch_prepare(corpus, ldc, n, E, reinterpret_cast<float*>(corpus_prepared), corpus_norms, (int64_t)TT_INDEX_ROWS_PAD(n), as_stream(stream));
}

int tt_index_topk(const float* Q, int ldq, const float* corpus, int ldc, const void* corpus_prepared, const float* corpus_norms, int nq,
                  int64_t n, int E, int K, int64_t idx_base, float* out_scores, int32_t* out_idx, void* ws, size_t ws_bytes, int impl,
                  void* stream) {
    TT_REQUIRE(Q && corpus && out_scores && out_idx, "tt_index_topk: null pointer");
    TT_REQUIRE(nq >= 0 && n >= 1 && E >= 1 && K >= 1 && ldq >= E && ldc >= E, "tt_index_topk: bad shape");
    TT_REQUIRE(idx_base >= 0 && idx_base + n < 2147483647ll - 64, "tt_index_topk: indices must fit int32");
    if (nq == 0) return TT_OK;
    cudaStream_t st = as_stream(stream);
    bool tc_ok = index_tc_supported(ldq, ldc, E, K, n, Q, corpus);
    if (impl == TT_IMPL_TC && !tc_ok) {
        set_error("tt_index_topk: TT_IMPL_TC unsupported for this shape/device");
        return TT_ERR_UNSUPPORTED;
    }
    if (impl == TT_IMPL_TC || (impl == TT_IMPL_AUTO && tc_ok))
        return index_tc(Q, ldq, corpus, ldc, reinterpret_cast<const float*>(corpus_prepared), corpus_norms, nq, n, E, K, idx_base, out_scores, out_idx, ws, ws_bytes, st);
    return index_exact(Q, ldq, corpus, ldc, nq, n, E, K, idx_base, out_scores, out_idx, ws, ws_bytes, st, nullptr);
}

int tt_topk_merge(const float* scores, const int32_t* idx, int G, int nq, int K, float* out_scores, int32_t* out_idx, void* stream) {
    TT_REQUIRE(scores && idx && out_scores && out_idx, "tt_topk_merge: null pointer");
    TT_REQUIRE(G >= 1 && nq >= 0 && K >= 1, "tt_topk_merge: bad shape");
    if (nq == 0) return TT_OK;
    return merge_launch(scores, idx, G, nq, K, 0, out_scores, out_idx, as_stream(stream));
}

}  // extern "C"
