// tt_index_tc.cu -- brute-force index on the tensor cores with EXACT results.
//
// Replaces (reference file:line): brute_force.py:75-78 (scores = Q.C^T) and :81 (top_k, sorted, lower index
// first on ties).  The (nq x n) score matrix never reaches HBM, and the returned (score, index) pairs are the
// canonical fp32 values -- bit-identical to the exact path and to oracle/tt_oracle.c -- although the
// heavy lifting runs in TF32 on tcgen05:
//
//   1. rowpanel_kernel<kIndex>   TF32 scores a_ij with the error bound eps_ij = kappa_i*||c_j||, kappa_i = 2^-9*||q_i||;
//                                only max_j (a_ij - eps_ij) over every group of 32 consecutive corpus rows is
//                                kept (n/32 floats per query).
//   2. select_threshold_kernel   lambda_i = K-th largest group value of the query (radix select).
//   3. rowpanel_kernel<kCollect> same TF32 contraction; every chunk of 32 columns whose maximum can reach lambda_i
//                                (~1.25 K chunks per query) is dumped whole into the query's hit queue.
//   4. collect_rescore_kernel    per column: a_ij + eps_ij >= lambda_i -> exact canonical fp32 score from the
//                                authoritative corpus -> drop s < lambda_i -> exact top-K by (score desc, index asc).
//   5. a query whose queue or list overflowed is redone by the exact CUDA-core kernel (tt_index.cu), on the device.
//
// Why it is exact: |a_ij - s_ij| <= eps_ij (TF32 operand rounding 2^-11 each and Cauchy-Schwarz, plus the fp32
// accumulation error of both evaluations, with a 2x margin).  K groups hold a row with s >= a - eps >= lambda,
// so the exact K-th best score s_K >= lambda; every true top-K row has s >= s_K, hence a + eps >= lambda, and is
// collected; step 4 orders the collected rows by the exact (score desc, index asc) rule.
#include "tt_tc_rowpanel.cuh"

namespace tt {

// tt_index.cu
size_t index_exact_workspace(int nq, int64_t n, int K);
int index_exact(const float* Q, int ldq, const float* C, int ldc, int nq, int64_t n, int E, int K, int64_t idx_base, float* out_s,
                int32_t* out_i, void* ws, size_t ws_bytes, cudaStream_t st, const int32_t* flags);
// tt_softmax_tc.cu
bool softmax_tc_supported(int ldq, int ldc, int E, const void* Q, const void* C);

namespace tc {

constexpr int kGroup = 32;
static inline int idx_bn(int E) { return E <= 64 ? 256 : 128; }   // shared-memory budget of the T ring
constexpr float kEpsCoef = 1.0f / 512.0f;   // 2^-9
static int g_cap_override = 0;              // tests: force tiny candidate lists to exercise the fallback

static inline int cand_cap(int K) { return g_cap_override > 0 ? g_cap_override : 4 * K + 512; }
static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

__device__ __forceinline__ uint32_t ordered_key(float f) {   // monotone float -> uint
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(b);
}

// The filter's group statistics assume the best rows are spread over many groups of 32.  Vocabularies are
// ordered by frequency (features.py:119-127), so popular -- and, once trained, high-scoring -- candidates sit
// next to each other; the prepared copy therefore stores the corpus under the fixed affine permutation
// pos = (orig * A) mod n, orig = (pos * Ainv) mod n, with A ~ 0.618 n coprime to n (a low-discrepancy spread of
// any run of consecutive rows).  Candidate lists hold permuted positions; rescoring maps them back.
struct Perm {
    unsigned long long n, a, a_inv;
};
static unsigned long long gcd_u64(unsigned long long x, unsigned long long y) { while (y) { unsigned long long t = x % y; x = y; y = t; } return x; }
static Perm make_perm(int64_t n) {
    Perm pm;
    pm.n = (unsigned long long)n;
    if (n <= 2) { pm.a = 1; pm.a_inv = 1; return pm; }
    unsigned long long a = (unsigned long long)((double)n * 0.6180339887498949) | 1ull;
    while (a >= pm.n || gcd_u64(a, pm.n) != 1) { a += 2; if (a >= pm.n) a = 1; }
    // modular inverse by the extended Euclid algorithm (signed 128-bit free: values stay below n < 2^31)
    long long t = 0, nt = 1, r = (long long)pm.n, nr = (long long)a;
    while (nr != 0) {
        long long qq = r / nr;
        long long tmp = t - qq * nt; t = nt; nt = tmp;
        tmp = r - qq * nr; r = nr; nr = tmp;
    }
    if (t < 0) t += (long long)pm.n;
    pm.a = a;
    pm.a_inv = (unsigned long long)t;
    return pm;
}
__device__ __forceinline__ int64_t perm_orig(const Perm& pm, int64_t pos) { return (int64_t)(((unsigned long long)pos * pm.a_inv) % pm.n); }

// Q32 = tf32_rn(Q); kappa[q] = kEpsCoef * ||q||.  One warp per query row.
__global__ void __launch_bounds__(256) prep_queries_kernel(const float* __restrict__ Q, int ldq, int nq, int E, float* __restrict__ Q32,
                                                           float* __restrict__ eps) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    float s = 0.f;
    for (int k = lane; k < E; k += 32) {
        float v = Q[(int64_t)q * ldq + k];
        Q32[(int64_t)q * E + k] = tf32_rn(v);
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) eps[q] = kEpsCoef * sqrtf(s) * 1.0001f + 1e-30f;
}

// One CTA per query: lambda = a lower bound, tight to 16 significant bits, of the K-th largest of gmax[q][0..ngroups).
// Any lambda <= the K-th largest group value keeps the filter exact (a smaller lambda only admits more candidates), so
// the radix select skips the bits every key shares (group maxima of one query differ only from about the 9th bit on:
// without the skip nearly all keys land in one histogram bin and the shared-memory atomics serialise) and stops after
// two 8-bit digits, returning the lower edge of the bin that holds the K-th largest key.
__global__ void __launch_bounds__(256) select_threshold_kernel(const float* __restrict__ gmax, int ld, int ngroups, int K,
                                                               float* __restrict__ thr, int32_t* __restrict__ flags) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_remaining, s_red[16];
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* row = gmax + (int64_t)q * ld;
    if (threadIdx.x == 0) flags[q] = 0;
    if (ngroups < K) {   // fewer groups than K: everything is a candidate (the list will overflow unless n is tiny)
        if (threadIdx.x == 0) thr[q] = -CUDART_INF_F;
        return;
    }
    constexpr int kCache = 16;                         // values cached in registers when the row is short
    uint32_t cache[kCache];
    const bool cached = ngroups <= kCache * 256;
    uint32_t kmax = 0u, kmin = 0xFFFFFFFFu;
    if (cached) {
#pragma unroll
        for (int i = 0; i < kCache; ++i) {
            int g = threadIdx.x + i * 256;
            cache[i] = g < ngroups ? ordered_key(row[g]) : 0u;
            if (g < ngroups) { kmax = max(kmax, cache[i]); kmin = min(kmin, cache[i]); }
        }
    } else {
        for (int g = threadIdx.x; g < ngroups; g += 256) {
            uint32_t k = ordered_key(row[g]);
            kmax = max(kmax, k); kmin = min(kmin, k);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    }
    if (lane == 0) { s_red[warp] = kmax; s_red[8 + warp] = kmin; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) { kmax = max(kmax, s_red[w]); kmin = min(kmin, s_red[8 + w]); }
    const uint32_t diff = kmax ^ kmin;
    if (diff == 0u) {   // every group has the same value
        if (threadIdx.x == 0) thr[q] = key_to_float(kmax);
        return;
    }
    const int common = __clz(diff);                    // leading bits shared by every key
    uint32_t mask = common ? (0xFFFFFFFFu << (32 - common)) : 0u;
    uint32_t prefix = kmax & mask, remaining = (uint32_t)K;
    int shift = 32 - common;
    for (int pass = 0; pass < 2 && shift > 0; ++pass) {
        const int bits = shift < 8 ? shift : 8;
        shift -= bits;
        const uint32_t dmask = (1u << bits) - 1u;
        hist[threadIdx.x] = 0;
        __syncthreads();
        if (cached) {
#pragma unroll
            for (int i = 0; i < kCache; ++i) {
                int g = threadIdx.x + i * 256;
                if (g < ngroups && (cache[i] & mask) == prefix) atomicAdd(&hist[(cache[i] >> shift) & dmask], 1u);
            }
        } else {
            for (int g = threadIdx.x; g < ngroups; g += 256) {
                uint32_t k = ordered_key(row[g]);
                if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & dmask], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {   // warp 0: find the bin holding the `remaining`-th largest key (suffix scan over 256 bins)
            uint32_t c[8], tot = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; tot += c[i]; }
            uint32_t above = 0;   // keys in bins of higher lanes
            uint32_t run = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_down_sync(0xffffffffu, run, o);
                if (lane + o < 32) run += v;
            }
            above = run - tot;    // run = inclusive suffix sum over lanes >= lane
            const bool mine = above < remaining && remaining <= above + tot;
            if (mine) {
                uint32_t cum = above;
                int b = 7;
                for (; b > 0; --b) {
                    if (cum + c[b] >= remaining) break;
                    cum += c[b];
                }
                s_prefix = prefix | ((uint32_t)(lane * 8 + b) << shift);
                s_remaining = remaining - cum;
            }
        }
        __syncthreads();
        prefix = s_prefix;
        remaining = s_remaining;
        mask |= dmask << shift;
    }
    if (threadIdx.x == 0) {
        float v = key_to_float(prefix);                // unresolved low bits are zero: the lower edge of the bin
        thr[q] = (v == v) ? v : -CUDART_INF_F;
    }
}

// One CTA per query.  Walks the query's hit queue (chunks of 32 TF32 scores dumped by rowpanel_kernel<kCollect>), keeps every
// column with a_j + kappa*||c_j|| >= lambda, computes its exact canonical fp32 score from the authoritative corpus, drops
// rows with s < lambda (the exact K-th best score is >= lambda, see the file header), and sorts the survivors by
// (score desc, index asc).  Overflow anywhere hands the query to the exact fallback.
template <int E>
__global__ void __launch_bounds__(256) collect_rescore_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ C, int ldc, int K,
                                                              int64_t n, const float* __restrict__ queue, const int32_t* __restrict__ segcnt,
                                                              int nseg, int cap_seg, const float* __restrict__ norms,
                                                              const float* __restrict__ eps, const float* __restrict__ thr, int cap, int P, Perm pm,
                                                              int64_t idx_base, float* __restrict__ out_s, int32_t* __restrict__ out_i,
                                                              int32_t* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sq = reinterpret_cast<float*>(smem_raw);          // 128 floats: the query row
    float* ss = sq + 128;                                      // P scores
    int32_t* si = reinterpret_cast<int32_t*>(ss + P);          // P indices
    int32_t* spos = si + P;                                    // cap permuted positions (first-stage list)
    __shared__ int s_n1, s_n2;
    const int q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_n1 = 0; s_n2 = 0; }
    const int my_cnt = tid < nseg ? segcnt[(int64_t)q * nseg + tid] : 0;
    for (int k = tid; k < E; k += 256) sq[k] = Q[(int64_t)q * ldq + k];
    if (__syncthreads_or(my_cnt > cap_seg)) {   // a queue segment overflowed
        if (tid == 0) flags[q] = 1;
        return;
    }
    const float kappa = eps[q], lambda = thr[q];
    const float* qbase = queue + (int64_t)q * nseg * cap_seg * kHitWords;
    // stage 1: per-column upper-bound test; a warp takes whole segments, four entries in flight
    for (int s = warp; s < nseg; s += 8) {
        const int cnt = __shfl_sync(0xffffffffu, segcnt[(int64_t)q * nseg + s], 0);
        const float* sb = qbase + (int64_t)s * cap_seg * kHitWords;
        for (int e0 = 0; e0 < cnt; e0 += 4) {
            float a[4];
            int nb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int e = min(e0 + u, cnt - 1);
                nb[u] = __float_as_int(__ldg(sb + e * kHitWords));
                a[u] = __ldg(sb + e * kHitWords + 4 + lane);
            }
            float nr[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) nr[u] = (nb[u] + lane < n) ? __ldg(norms + nb[u] + lane) : -CUDART_INF_F;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (e0 + u < cnt && fmaf(kappa, nr[u], a[u]) >= lambda) {   // -inf norm: column beyond the corpus
                    const int slot = atomicAdd(&s_n1, 1);
                    if (slot < cap) spos[slot] = nb[u] + lane;
                }
            }
        }
    }
    __syncthreads();
    const int m1 = s_n1;
    if (m1 > cap) {
        if (tid == 0) flags[q] = 1;
        return;
    }
    // stage 2: exact canonical score (k ascending, one fmaf per term), exact prune against lambda
    for (int t = tid; t < m1; t += 256) {
        const int32_t id = (int32_t)perm_orig(pm, spos[t]);   // permuted position -> original corpus row
        const float4* row = reinterpret_cast<const float4*>(C + (int64_t)id * ldc);
        float4 cv[E / 4];
#pragma unroll
        for (int k4 = 0; k4 < E / 4; ++k4) cv[k4] = __ldg(row + k4);
        float acc = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < E / 4; ++k4) {
            const float4 qv = *reinterpret_cast<const float4*>(sq + 4 * k4);
            acc = fmaf(qv.x, cv[k4].x, acc);
            acc = fmaf(qv.y, cv[k4].y, acc);
            acc = fmaf(qv.z, cv[k4].z, acc);
            acc = fmaf(qv.w, cv[k4].w, acc);
        }
        if (acc >= lambda) {
            const int slot = atomicAdd(&s_n2, 1);
            ss[slot] = acc;
            si[slot] = id;
        }
    }
    __syncthreads();
    const int m2 = s_n2;
    int need = m2 > K ? m2 : K, Ps = 2;
    while (Ps < need) Ps <<= 1;
    Ps = Ps < P ? Ps : P;
    for (int t = m2 + tid; t < Ps; t += 256) { ss[t] = -CUDART_INF_F; si[t] = 0x7fffffff; }
    __syncthreads();
    for (int k = 2; k <= Ps; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < Ps / 2; t += 256) {
                int i = 2 * t - (t & (j - 1));
                int p2 = i + j;
                bool up = ((i & k) == 0);
                float a = ss[i], b = ss[p2];
                int32_t ia = si[i], ib = si[p2];
                bool swap = up ? ranks_before(b, ib, a, ia) : ranks_before(a, ia, b, ib);
                if (swap) { ss[i] = b; ss[p2] = a; si[i] = ib; si[p2] = ia; }
            }
            __syncthreads();
        }
    }
    for (int t = tid; t < K; t += 256) {
        const bool pad = si[t] == 0x7fffffff;
        out_s[(int64_t)q * K + t] = pad ? -CUDART_INF_F : ss[t];
        out_i[(int64_t)q * K + t] = pad ? -1 : (int32_t)(si[t] + idx_base);
    }
}

// C32p[pos][:] = tf32_rn(C[orig(pos)][:]); norms[pos] = ||C[orig(pos)]|| (rounded up a hair); norms[n..n_pad) = 0.
__global__ void __launch_bounds__(256) prepare_corpus_kernel(const float* __restrict__ C, int ldc, int64_t n, int E, Perm pm,
                                                             float* __restrict__ C32p, float* __restrict__ norms, int64_t n_pad) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t pos = warp; pos < n_pad; pos += nwarps) {
        float s = 0.f;
        if (pos < n) {
            const int64_t o = perm_orig(pm, pos);
            for (int k = lane; k < E; k += 32) {
                float v = C[o * ldc + k];
                C32p[pos * E + k] = tf32_rn(v);
                s = fmaf(v, v, s);
            }
        }
        s = warp_sum(s);
        if (lane == 0) norms[pos] = pos < n ? sqrtf(s) * 1.0001f : 0.f;
    }
}

// gn[c] = max over the 32 rows of chunk c of norms[] (norms is zero beyond n)
__global__ void chunk_max_kernel(const float* __restrict__ norms, int64_t nchunks, float* __restrict__ gn) {
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    float m = 0.f;
    for (int i = 0; i < 32; ++i) m = fmaxf(m, norms[c * 32 + i]);
    gn[c] = m;
}

// norms buffer layout: [n_pad per-row norms][n_pad/32 per-chunk maxima]
int launch_prepare(const float* C, int ldc, int64_t n, int E, float* C32p, float* norms, int64_t n_pad, cudaStream_t st) {
    int64_t g = ceil_div(n_pad * 32, 256);
    int64_t cap = (int64_t)sm_count() * 16;
    prepare_corpus_kernel<<<(unsigned)(g > cap ? cap : (g < 1 ? 1 : g)), 256, 0, st>>>(C, ldc, n, E, make_perm(n), C32p, norms, n_pad);
    TT_LAUNCH_OK("prepare_corpus_kernel");
    const int64_t nchunks = n_pad / 32;
    chunk_max_kernel<<<(unsigned)ceil_div(nchunks, 256), 256, 0, st>>>(norms, nchunks, norms + n_pad);
    TT_LAUNCH_OK("chunk_max_kernel");
    return TT_OK;
}

template <int MODE, int E>
static int launch_idx(const CUtensorMap& tmQ, const CUtensorMap& tmC, const RowPanelParams& p, int m_tiles, int splits, cudaStream_t st,
                      const char* name) {
    constexpr int BN = (E <= 64) ? 256 : 128;
    using Cfg = RowPanelCfg<MODE, E, BN>;
    static bool attr_done = false;
    if (!attr_done) {
        TT_CUDA_OK(cudaFuncSetAttribute(rowpanel_kernel<MODE, E, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_done = true;
    }
    dim3 grid((unsigned)m_tiles, (unsigned)splits);
    rowpanel_kernel<MODE, E, BN><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tmQ, tmC, tmC, p);
    TT_LAUNCH_OK(name);
    return TT_OK;
}

template <int MODE>
static int launch_idx_e(int E, const CUtensorMap& tmQ, const CUtensorMap& tmC, const RowPanelParams& p, int m_tiles, int splits, cudaStream_t st,
                        const char* name) {
    switch (E) {
        case 32: return launch_idx<MODE, 32>(tmQ, tmC, p, m_tiles, splits, st, name);
        case 64: return launch_idx<MODE, 64>(tmQ, tmC, p, m_tiles, splits, st, name);
        default: return launch_idx<MODE, 128>(tmQ, tmC, p, m_tiles, splits, st, name);
    }
}

struct IdxLayout {
    size_t q32, eps, thr, gmax, cnt, flags, queue, c32, norms, exact, total;
    int ngroups, n_tiles, cap;
    int m_tiles, splits, tps, nseg, cap_seg;
};

static IdxLayout layout(int nq, int64_t n, int E, int K, bool need_corpus_copy) {
    IdxLayout L;
    L.n_tiles = (int)ceil_div(n, idx_bn(E));
    L.ngroups = L.n_tiles * (idx_bn(E) / kGroup);
    L.cap = cand_cap(K);
    L.m_tiles = (int)ceil_div(nq, 128);
    choose_splits(L.m_tiles, L.n_tiles, 2, 64, &L.splits, &L.tps);
    L.nseg = L.splits * (idx_bn(E) / 32 >= 2 ? 2 : 1);   // RowPanelCfg::kHalves
    // hit-queue segment of one (row, split, half): ~1.25 K / nseg qualifying chunks are expected; generous slack because an
    // overflow costs a trip through the exact CUDA-core fallback
    L.cap_seg = g_cap_override > 0 ? 1 + g_cap_override / L.nseg : (int)ceil_div(2 * K, L.nseg) + 16;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    L.q32 = take((size_t)nq * E * 4);
    L.eps = take((size_t)nq * 4);
    L.thr = take((size_t)nq * 4);
    L.gmax = take((size_t)nq * L.ngroups * 4);
    L.cnt = take((size_t)nq * L.nseg * 4);
    L.flags = take((size_t)nq * 4);
    L.queue = take((size_t)nq * L.nseg * L.cap_seg * kHitWords * 4);
    L.c32 = take(need_corpus_copy ? (size_t)n * E * 4 : 0);
    L.norms = take(need_corpus_copy ? (size_t)TT_INDEX_NORM_PAD(n) * 4 : 0);
    L.exact = off;
    L.total = off;
    return L;
}

}  // namespace tc

using namespace tc;

bool index_tc_supported(int ldq, int ldc, int E, int K, int64_t n, const void* Q, const void* C) {
    if (!softmax_tc_supported(ldq, ldc, E, Q, C)) return false;
    if (K < 1 || K > 1024) return false;
    if (n < 4096) return false;                                   // tiny corpora: the exact kernel is already fast
    if (ceil_div(n, kGroup) < 4 * (int64_t)K) return false;      // the group filter needs many more groups than K
    return true;
}

size_t index_tc_workspace(int nq, int64_t n, int E, int K, bool need_corpus_copy) {
    if (!(E == 32 || E == 64 || E == 128)) return 0;
    return layout(nq, n, E, K, need_corpus_copy).total + 512;
}

void debug_index_cap(int cap) { tc::g_cap_override = cap; }

int index_tc(const float* Q, int ldq, const float* C, int ldc, const float* C32_in, const float* norms_in, int nq, int64_t n, int E, int K,
             int64_t idx_base, float* out_s, int32_t* out_i, void* ws, size_t ws_bytes, cudaStream_t st) {
    TT_REQUIRE((C32_in == nullptr) == (norms_in == nullptr), "tt_index_topk: corpus_prepared and corpus_norms must be given together");
    TT_REQUIRE(C32_in == nullptr || (reinterpret_cast<uintptr_t>(C32_in) & 15) == 0, "tt_index_topk: corpus_prepared must be 16-byte aligned");
    const bool need_copy = (C32_in == nullptr);
    IdxLayout L = layout(nq, n, E, K, need_copy);
    const size_t exact_bytes = index_exact_workspace(nq, n, K);
    TT_REQUIRE(ws && ws_bytes >= L.total + exact_bytes, "tt_index_topk: workspace too small (%zu < %zu)", ws_bytes, L.total + exact_bytes);
    char* base = reinterpret_cast<char*>(ws);
    float* q32 = reinterpret_cast<float*>(base + L.q32);
    float* eps = reinterpret_cast<float*>(base + L.eps);
    float* thr = reinterpret_cast<float*>(base + L.thr);
    float* gmax = reinterpret_cast<float*>(base + L.gmax);
    int32_t* cnt = reinterpret_cast<int32_t*>(base + L.cnt);
    int32_t* flags = reinterpret_cast<int32_t*>(base + L.flags);
    float* queue = reinterpret_cast<float*>(base + L.queue);
    const float* c32 = C32_in;
    const float* norms = norms_in;
    if (need_copy) {
        float* dst = reinterpret_cast<float*>(base + L.c32);
        float* nd = reinterpret_cast<float*>(base + L.norms);
        int rcn = launch_prepare(C, ldc, n, E, dst, nd, (int64_t)TT_INDEX_ROWS_PAD(n), st);
        if (rcn) return rcn;
        c32 = dst;
        norms = nd;
    }
    prep_queries_kernel<<<(unsigned)ceil_div((int64_t)nq * 32, 256), 256, 0, st>>>(Q, ldq, nq, E, q32, eps);
    TT_LAUNCH_OK("prep_queries_kernel");

    CUtensorMap tmQ, tmC;
    int rc = make_tmap_2d(&tmQ, q32, nq, E, E, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmC, c32, n, E, E, idx_bn(E));   // the prepared copy is dense (ld = E)
    if (rc) return rc;
    const int m_tiles = L.m_tiles, splits = L.splits, tps = L.tps;
    TT_REQUIRE(L.nseg <= 256, "tt_index_topk: too many column splits");
    RowPanelParams p{};
    p.nR = nq; p.nT = (int)n; p.n_tiles = L.n_tiles; p.tiles_per_split = tps; p.rowv = eps; p.rowv2 = thr; p.colv2 = norms; p.gnorm = norms + TT_INDEX_ROWS_PAD(n); p.d = -(1 << 30);
    p.out0 = gmax; p.out1 = nullptr; p.out2 = nullptr; p.ld_out = L.ngroups; p.trace = nullptr;
    rc = launch_idx_e<kIndex>(E, tmQ, tmC, p, m_tiles, splits, st, "rowpanel_kernel<index>");
    if (rc) return rc;
    select_threshold_kernel<<<(unsigned)nq, 256, 0, st>>>(gmax, L.ngroups, (int)ceil_div(n, kGroup), K, thr, flags);
    TT_LAUNCH_OK("select_threshold_kernel");
    p.out0 = queue;
    p.out1 = reinterpret_cast<float*>(cnt);
    p.ld_out = L.cap_seg;
    rc = launch_idx_e<kCollect>(E, tmQ, tmC, p, m_tiles, splits, st, "rowpanel_kernel<collect>");
    if (rc) return rc;
    const int P = next_pow2(L.cap < 2 ? 2 : L.cap);
    const size_t smem = 128 * 4 + (size_t)P * 8 + (size_t)L.cap * 4;
    auto launch_rescore = [&](auto kern) -> int {
        if (smem > 48 * 1024) TT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)nq, 256, smem, st>>>(Q, ldq, C, ldc, K, n, queue, cnt, L.nseg, L.cap_seg, norms, eps, thr, L.cap, P, make_perm(n), idx_base,
                                             out_s, out_i, flags);
        return TT_OK;
    };
    rc = E == 32 ? launch_rescore(collect_rescore_kernel<32>) : E == 64 ? launch_rescore(collect_rescore_kernel<64>) : launch_rescore(collect_rescore_kernel<128>);
    if (rc) return rc;
    TT_LAUNCH_OK("collect_rescore_kernel");
    // queries whose candidate list overflowed: exact CUDA-core path, decided on the device (no host sync)
    return index_exact(Q, ldq, C, ldc, nq, n, E, K, idx_base, out_s, out_i, base + L.exact, exact_bytes, st, flags);
}

}  // namespace tt
