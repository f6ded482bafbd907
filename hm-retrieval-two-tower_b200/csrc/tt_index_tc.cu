// tt_index_tc.cu -- brute-force index on the tensor cores with EXACT results.
//
// Replaces (reference file:line): brute_force.py:75-78 (scores = Q.C^T) and :81 (top_k, sorted, lower index
// first on ties).  The (nq x n) score matrix never reaches HBM, and the returned (score, index) pairs are the
// canonical fp32 values -- bit-identical to the exact path and to oracle/tt_oracle.c -- although the
// heavy lifting runs in TF32 on tcgen05:
//
//   1. rowpanel_kernel<kIndex>   on a SAMPLE of the corpus (corpora of 400 k rows and more: the first quarter of the permuted copy,
//                                i.e. a pseudo-random quarter of the rows; smaller ones: all of it): TF32 scores a_ij with the error bound eps_ij = kappa_i*||c_j||, kappa_i = 2^-9*||q_i||;
//                                only a lower bound of the best score of every group of consecutive corpus rows
//                                (128 rows; 32 for small corpora) is kept: max_j a_ij - kappa_i * max_j ||c_j||.
//   2. select_threshold_kernel   lambda_i = (a lower bound, tight to 16 bits, of) the r-th largest group value of the sample,
//                                r = K f + 6 sqrt(K f (1 - f)) + 1 for a sample fraction f: with probability 1 - 1e-9 fewer than r
//                                of the true top-K rows fall into the sample, so lambda_i <= the exact K-th best score.  (One and a
//                                quarter tensor passes instead of two; lambda sits near rank r / f, so ~2x more columns are rescored.)
//   3. rowpanel_kernel<kCollect> same TF32 contraction; every chunk of 32 columns whose maximum can reach lambda_i
//                                (~1.25 K chunks per query) is appended whole to the dumping warp's hit log.
//   4. column_test_kernel        per logged column: a_ij >= lambda_i - kappa_i * max_chunk ||c|| -> the query's list.
//   5. exact_score_kernel        exact canonical fp32 score of every listed column from the authoritative corpus; rows
//                                with s < lambda_i are dropped.
//   6. sort_topk_kernel          exact top-K by (score desc, index asc).
//   7. a query whose log or list overflowed, OR that ends with fewer than K rescored rows at or above lambda_i (the sampled
//      threshold was too high: never observed, probability ~1e-9), is redone by the exact CUDA-core kernel (tt_index.cu), on
//      the device.
//
// Why it is exact: |a_ij - s_ij| <= eps_ij (TF32 operand rounding 2^-11 each and Cauchy-Schwarz, plus the fp32
// accumulation error of both evaluations, with a 2x margin), so every row with s >= lambda has a + eps >= lambda and is
// listed.  When at least K listed rows have an exact score >= lambda, the exact K-th best score s_K is >= lambda, every true
// top-K row (s >= s_K) is among them, and steps 5-6 order them by the exact (score desc, index asc) rule; otherwise step 7.
#include <cuda_fp16.h>

#include "tt_tc_rowpanel.cuh"

namespace tt {

// tt_index.cu
size_t index_exact_workspace(int nq, int64_t n, int K);
int index_exact(const float* Q, int ldq, const float* C, int ldc, int nq, int64_t n, int E, int K, int64_t idx_base, float* out_s,
                int32_t* out_i, void* ws, size_t ws_bytes, cudaStream_t st, const int32_t* flags);
// tt_softmax_tc.cu
bool softmax_tc_supported(int ldq, int ldc, int E, const void* Q, const void* C);

namespace tc {

static inline int idx_bn(int E) { return E <= 64 ? 256 : 128; }   // shared-memory budget of the T ring
static inline int idx_halves(int E) { return idx_bn(E) / 32 >= 2 ? 2 : 1; }   // RowPanelCfg::kHalves
constexpr float kEpsCoef = 1.0f / 512.0f;   // 2^-9
static int g_cap_override = 0;              // tests: force tiny candidate lists to exercise the fallback

static inline int cand_cap(int rank) {   // listed columns per query for a threshold near `rank` (a power of two: it is also the largest sort size)
    // 2x the rank plus slack: a trained, popularity-dominated corpus packs ~1.4-1.9 K rows within the filter's error bound of the
    // K-th best score (measured), and one overflowing query costs a trip through the exact path
    int c = 128;
    while (c < 2 * rank + 64) c <<= 1;
    return c < 1024 ? c : 1024;
}

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

// Operand format of the filter passes.  E >= 64: fp16 tiles (tcgen05 kind::f16, twice the TF32 rate and half the bytes) of
// power-of-two SCALED operands -- every query row by s_q (its own: only scores of one row are ever compared), the corpus by one
// s_c -- chosen so the largest magnitude lands in [2^14, 2^15): no overflow, and an element only becomes an fp16 subnormal when it
// is < 2^-28 of the largest one.  fp16 and TF32 both keep 11 significant bits, so the relative part of the error bound is the same:
//     |a'_ij - S s_ij| <= kappa'_i ||c_j|| + alpha'_i,   S = s_q s_c,  a' = the fp16 tensor-core score of the scaled operands,
//     kappa'_i = (2^-9 ||q_i|| s_q + 2^-25 sqrt(E)) s_c      (relative rounding of both operands with a 2x margin for the fp32
//                                                             accumulation; + query elements that fell into the subnormal range)
//     alpha'_i = 2^-25 sqrt(E) ||q_i|| s_q                   (corpus elements that fell into the subnormal range: |delta| <= 2^-25)
// The passes work on scaled scores throughout (group bounds, lambda', hit logs); only the exact rescoring compares unscaled scores
// and takes lambda = (lambda' - alpha') / S.  E = 32 keeps TF32 tiles of the unscaled operands (alpha' = 0, S = 1).
__device__ __forceinline__ float pow2_scale_for(float amax) {   // s = 2^k with amax * s in [2^14, 2^15); 1 for zero / non-finite input
    if (!(amax > 0.f) || !(amax < CUDART_INF_F)) return 1.f;
    int k = 14 - ilogbf(amax);
    k = k > 60 ? 60 : (k < -60 ? -60 : k);
    return exp2f((float)k);
}
__host__ __device__ inline bool idx_half_operands(int E) { return E >= 64; }

// prepared queries; kappa[q], alpha[q], inv_s[q] as above.  One warp per query row.
template <bool H>
__global__ void __launch_bounds__(256) prep_queries_kernel(const float* __restrict__ Q, int ldq, int nq, int E, void* __restrict__ Qp,
                                                           float* __restrict__ eps, float* __restrict__ alpha, float* __restrict__ inv_s,
                                                           const float* __restrict__ corpus_scale) {
    const int lane = threadIdx.x & 31;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    float s = 0.f, amax = 0.f;
    for (int k = lane; k < E; k += 32) {
        const float v = Q[(int64_t)q * ldq + k];
        if (!H) reinterpret_cast<float*>(Qp)[(int64_t)q * E + k] = tf32_rn(v);
        s = fmaf(v, v, s);
        amax = fmaxf(amax, fabsf(v));
    }
    s = warp_sum(s);
    if (!H) {
        if (lane == 0) { eps[q] = kEpsCoef * sqrtf(s) * 1.0001f + 1e-30f; alpha[q] = 0.f; inv_s[q] = 1.f; }
        return;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float sq = pow2_scale_for(amax), sc = __ldg(corpus_scale);
    for (int k = lane; k < E; k += 32)
        reinterpret_cast<__half*>(Qp)[(int64_t)q * E + k] = __float2half_rn(Q[(int64_t)q * ldq + k] * sq);
    if (lane == 0) {
        const float nq2 = sqrtf(s) * 1.0001f * sq;                    // ||q'|| (rounded up a hair)
        const float sub = 2.98023224e-8f * sqrtf((float)E) * 1.0001f;  // 2^-25 sqrt(E)
        eps[q] = (kEpsCoef * nq2 + sub) * sc * 1.0001f + 1e-30f;
        alpha[q] = sub * nq2;
        inv_s[q] = 1.f / (sq * sc);                                    // a power of two: exact
    }
}

// One CTA per query: lambda = a lower bound, tight to 16 significant bits, of the K-th largest of gmax[q][0..ngroups).
// Any lambda <= the K-th largest group value keeps the filter exact (a smaller lambda only admits more candidates), so
// the radix select skips the bits every key shares (group values of one query differ only from about the 9th bit on:
// without the skip nearly all keys land in one histogram bin and the shared-memory atomics serialise) and stops after
// two 8-bit digits, returning the lower edge of the bin that holds the K-th largest key.
constexpr int kSelectThreads = 128;
constexpr int kSelectCache = 8;                        // keys held in registers per thread (rows up to 1024 groups; longer rows are re-read)
// Writes thr[q] = lambda' - 2 alpha' (what the collect pass and the column test compare scaled tensor-core scores with) and
// thr_exact[q] = (lambda' - alpha') / S (what the exact rescoring compares unscaled fp32 scores with), both rounded down.
__global__ void __launch_bounds__(kSelectThreads) select_threshold_kernel(const float* __restrict__ gmax, int ld, int ngroups, int K,
                                                                          float* __restrict__ thr, float* __restrict__ thr_exact,
                                                                          const float* __restrict__ alpha, const float* __restrict__ inv_s,
                                                                          int32_t* __restrict__ flags, int32_t* __restrict__ cand_cnt) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_remaining, s_red[2 * kSelectThreads / 32];
    const int q = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* row = gmax + (int64_t)q * ld;
    if (tid == 0) { flags[q] = 0; cand_cnt[q] = 0; }
    if (ngroups < K) {   // fewer groups than K: everything is a candidate (the list will overflow unless n is tiny)
        if (tid == 0) { thr[q] = -CUDART_INF_F; thr_exact[q] = -CUDART_INF_F; }
        return;
    }
    uint32_t cache[kSelectCache];
    const bool cached = ngroups <= kSelectCache * kSelectThreads;
    uint32_t kmax = 0u, kmin = 0xFFFFFFFFu;
    if (cached) {
#pragma unroll
        for (int i = 0; i < kSelectCache; ++i) {
            const int g = tid + i * kSelectThreads;
            cache[i] = g < ngroups ? ordered_key(__ldg(row + g)) : 0u;
            if (g < ngroups) { kmax = max(kmax, cache[i]); kmin = min(kmin, cache[i]); }
        }
    } else {
        for (int g = tid; g < ngroups; g += kSelectThreads) {
            const uint32_t k = ordered_key(__ldg(row + g));
            kmax = max(kmax, k); kmin = min(kmin, k);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    }
    if (lane == 0) { s_red[warp] = kmax; s_red[kSelectThreads / 32 + warp] = kmin; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kSelectThreads / 32; ++w) { kmax = max(kmax, s_red[w]); kmin = min(kmin, s_red[kSelectThreads / 32 + w]); }
    const uint32_t diff = kmax ^ kmin;
    if (diff == 0u) {   // every group has the same value
        if (tid == 0) thr[q] = key_to_float(kmax);
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
        hist[tid] = 0; hist[tid + kSelectThreads] = 0;
        __syncthreads();
        if (cached) {
#pragma unroll
            for (int i = 0; i < kSelectCache; ++i) {
                const int g = tid + i * kSelectThreads;
                if (g < ngroups && (cache[i] & mask) == prefix) atomicAdd(&hist[(cache[i] >> shift) & dmask], 1u);
            }
        } else {
            for (int g = tid; g < ngroups; g += kSelectThreads) {
                const uint32_t k = ordered_key(__ldg(row + g));
                if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & dmask], 1u);
            }
        }
        __syncthreads();
        if (tid < 32) {   // warp 0: find the bin holding the `remaining`-th largest key (suffix scan over 256 bins)
            uint32_t c[8], tot = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; tot += c[i]; }
            uint32_t run = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t v = __shfl_down_sync(0xffffffffu, run, o);
                if (lane + o < 32) run += v;
            }
            const uint32_t above = run - tot;    // keys in bins of higher lanes (run = inclusive suffix sum over lanes >= lane)
            if (above < remaining && remaining <= above + tot) {
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
    if (tid == 0) {
        float v = key_to_float(prefix);                // unresolved low bits are zero: the lower edge of the bin
        if (!(v == v)) v = -CUDART_INF_F;
        const float a = alpha[q];
        thr[q] = __fsub_rd(v, 2.0002f * a);
        thr_exact[q] = __fmul_rd(__fsub_rd(v, a), inv_s[q]) ;
    }
}

// ---- after the collect pass: three small kernels, each with whole-chip parallelism and no intra-query barrier -------
// (one CTA per query doing all three steps is bound by its own chain of dependent L2 round trips times the few CTAs an
//  SM can hold: measured 47 us for 2048 queries against ~20 us for the split form.)
//
// column_test_kernel    a quarter-warp (8 lanes x 16 B = one dumped chunk of 32 TF32 scores) per hit-queue segment
//                       (rowpanel_kernel<kCollect> dumps qualifying chunks into per-(query, split, half) segments);
//                       column j is listed when a_j >= lambda - kappa*max_chunk||c|| (the threshold the dumping lane
//                       stored in the entry; it is implied by the per-column bound a_j + kappa*||c_j|| >= lambda, so no
//                       true top-K row is lost).  Lists are appended with one global atomic per passing quad.
// exact_score_kernel    one thread per listed column: canonical fp32 score from the authoritative corpus (k ascending,
//                       one fmaf per term) -> 64-bit key (ordered score, ~index); rows with s < lambda get key 0 (the
//                       exact K-th best score is >= lambda, see the file header).
// sort_topk_kernel      one warp per query: bitonic network over the keys in registers, descending key == (score desc,
//                       index asc); writes the first K.
// A queue-segment overflow or more than `cap` (<= 1024) listed columns hands the query to the exact fallback.
constexpr int kTestThreads = 256;

// grid (ceil(cap_log / 32), logs of the collect pass): a quarter-warp per log entry
__global__ void __launch_bounds__(kTestThreads) column_test_kernel(const float* __restrict__ logs, const int32_t* __restrict__ log_cnt, int cap_log,
                                                                   int64_t n, int cap, int32_t* __restrict__ cand, int32_t* __restrict__ cand_cnt) {
    const int log = blockIdx.y;
    const int e = blockIdx.x * (kTestThreads / 8) + (threadIdx.x >> 3);
    const int sub = threadIdx.x & 7;
    if (e >= __ldg(log_cnt + log)) return;
    const float4* ent = reinterpret_cast<const float4*>(logs + ((int64_t)log * cap_log + e) * kHitWords);
    const float4 hd = __ldg(ent);
    const float4 a = __ldg(ent + 2 + sub);
    const int q = __float_as_int(hd.x);
    const float tc = hd.z;                                     // lambda - kappa * max ||c|| over the chunk
    const int col0 = __float_as_int(hd.y) + 4 * sub;
    uint32_t pass = (a.x >= tc ? 1u : 0u) | (a.y >= tc ? 2u : 0u) | (a.z >= tc ? 4u : 0u) | (a.w >= tc ? 8u : 0u);
    if ((int64_t)col0 + 4 > n) pass &= (col0 < n) ? (0xFu >> (4 - (int)(n - col0))) : 0u;   // columns beyond the corpus
    if (pass) {
        int slot = atomicAdd(cand_cnt + q, __popc(pass));
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if ((pass >> t) & 1u) { if (slot < cap) cand[(int64_t)q * cap + slot] = col0 + t; ++slot; }
    }
}

// grid (ceil(cap / 128), queries); a warp takes 32 listed columns: the rows are copied to shared memory with coalesced
// 128-bit loads (16-byte requests from 32 different rows would cost a 32-byte L2 sector access each -- the kernel is
// bound by L2 sector accesses), then lane j computes row j's canonical score.
template <int E>
struct ScoreCfg {
    static constexpr int kThreads = E >= 128 ? 64 : 128;   // static shared memory budget (32 staged rows per warp)
};
template <int E>
__global__ void __launch_bounds__(ScoreCfg<E>::kThreads) exact_score_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ C, int ldc,
                                                                    const int32_t* __restrict__ cand, const int32_t* __restrict__ cand_cnt, int cap,
                                                                    const float* __restrict__ thr, const int32_t* __restrict__ orig_of,
                                                                    unsigned long long* __restrict__ keys) {
    constexpr int kScoreThreads = ScoreCfg<E>::kThreads;
    constexpr int kRowLd = E + 4;                              // padded row stride (floats): conflict-free 128-bit row reads
    constexpr int kF4 = E / 4;                                 // float4s per row
    __shared__ __align__(16) float sq[E];
    __shared__ __align__(16) float srows[kScoreThreads / 32][32 * kRowLd];
    const int q = blockIdx.y;
    const int m1 = min(cand_cnt[q], cap);
    if ((int)(blockIdx.x * kScoreThreads) >= m1) return;
    for (int k = threadIdx.x; k < E; k += kScoreThreads) sq[k] = Q[(int64_t)q * ldq + k];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot0 = blockIdx.x * kScoreThreads + warp * 32;
    if (slot0 >= m1) return;
    const int nrow = min(32, m1 - slot0);
    const int32_t my_id = (lane < nrow) ? __ldg(orig_of + cand[(int64_t)q * cap + slot0 + lane]) : 0;   // permuted position -> original corpus row
    float* mine = srows[warp];
    float4 v[kF4];
#pragma unroll
    for (int i = 0; i < kF4; ++i) {                            // 32 * kF4 pieces, 32 per step: consecutive lanes, consecutive 16 bytes
        const int piece = i * 32 + lane;
        const int r = piece / kF4, c4 = piece % kF4;
        const int32_t id = __shfl_sync(0xffffffffu, my_id, r);
        if (r < nrow) v[i] = __ldg(reinterpret_cast<const float4*>(C + (int64_t)id * ldc) + c4);
    }
#pragma unroll
    for (int i = 0; i < kF4; ++i) {
        const int piece = i * 32 + lane;
        const int r = piece / kF4, c4 = piece % kF4;
        if (r < nrow) *reinterpret_cast<float4*>(mine + r * kRowLd + 4 * c4) = v[i];
    }
    __syncwarp();
    if (lane < nrow) {
        const float* row = mine + lane * kRowLd;
        float acc = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < kF4; ++k4) {
            const float4 cv = *reinterpret_cast<const float4*>(row + 4 * k4);
            const float4 qv = *reinterpret_cast<const float4*>(sq + 4 * k4);
            acc = fmaf(qv.x, cv.x, acc);
            acc = fmaf(qv.y, cv.y, acc);
            acc = fmaf(qv.z, cv.z, acc);
            acc = fmaf(qv.w, cv.w, acc);
        }
        keys[(int64_t)q * cap + slot0 + lane] =
            (acc >= thr[q]) ? (((unsigned long long)ordered_key(acc) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)my_id)) : 0ull;
    }
}

__device__ __forceinline__ void ce_desc(unsigned long long& a, unsigned long long& b, bool desc) {   // compare-exchange
    const bool sw = (a < b) == desc;
    const unsigned long long t = a;
    a = sw ? b : a;
    b = sw ? t : b;
}
// sorts the 32*S keys of a warp (element g = lane*S + i, the first m of them read from `src`) into descending order and
// writes the first K as (score, index)
template <int S>
__device__ __forceinline__ bool sort_write(const unsigned long long* __restrict__ src, int m, int K, int need_k, int64_t idx_base, float* __restrict__ os,
                                           int32_t* __restrict__ oi, int lane) {
    unsigned long long key[S];
    int alive = 0;
#pragma unroll
    for (int i = 0; i < S; ++i) {
        key[i] = (i * 32 + lane < m) ? src[i * 32 + lane] : 0ull;   // 0 sorts after every real key
        alive += key[i] != 0ull;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) alive += __shfl_xor_sync(0xffffffffu, alive, o);
    if (alive < need_k) return false;   // fewer than K rescored rows reach lambda: the threshold was too high (exact fallback)
    for (int k = 2; k <= 32 * S; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= S) {
                const int lj = j / S;
                const bool take_max = ((lane & lj) == 0) == (((lane * S) & k) == 0);
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const unsigned long long o = __shfl_xor_sync(0xffffffffu, key[i], lj);
                    key[i] = ((key[i] < o) == take_max) ? o : key[i];
                }
            } else {
#pragma unroll
                for (int jj = 1; jj < S; jj <<= 1) {
                    if (j == jj) {
#pragma unroll
                        for (int i = 0; i < S; ++i)
                            if ((i & jj) == 0) ce_desc(key[i], key[i | jj], (((lane * S + i) & k) == 0));
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < S; ++i) {
        const int g = lane * S + i;
        if (g < K) {
            const bool pad = key[i] == 0ull;                   // fewer survivors than K (only when n < K)
            os[g] = pad ? -CUDART_INF_F : key_to_float((uint32_t)(key[i] >> 32));
            oi[g] = pad ? -1 : (int32_t)((int64_t)(0xFFFFFFFFu - (uint32_t)key[i]) + idx_base);
        }
    }
    return true;
}

constexpr int kSortWarps = 4;             // queries per CTA
template <bool BIG>                       // BIG: up to 1024 listed columns (32 keys per lane); otherwise up to 256 (8 per lane)
__global__ void __launch_bounds__(32 * kSortWarps) sort_topk_kernel(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ cand_cnt,
                                                                    int nq, int cap, int K, int need_k, int64_t idx_base, float* __restrict__ out_s,
                                                                    int32_t* __restrict__ out_i, int32_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * kSortWarps + (threadIdx.x >> 5);
    if (q >= nq) return;
    const int m1 = cand_cnt[q];
    if (m1 > cap || flags[q]) {   // list overflow, or a queue segment overflowed: exact fallback
        if (lane == 0) flags[q] = 1;
        return;
    }
    const unsigned long long* src = keys + (int64_t)q * cap;
    float* os = out_s + (int64_t)q * K;
    int32_t* oi = out_i + (int64_t)q * K;
    const int need = m1 > K ? m1 : K;
    bool ok;
    if (need <= 128) ok = sort_write<4>(src, m1, K, need_k, idx_base, os, oi, lane);        // the network is sized per query, at run time
    else if (need <= 256) ok = sort_write<8>(src, m1, K, need_k, idx_base, os, oi, lane);
    else if constexpr (BIG) {
        if (need <= 512) ok = sort_write<16>(src, m1, K, need_k, idx_base, os, oi, lane);
        else ok = sort_write<32>(src, m1, K, need_k, idx_base, os, oi, lane);
    } else ok = false;
    if (!ok && lane == 0) flags[q] = 1;
}

// max |C| as float bits (non-negative floats order like their bit patterns)
__global__ void __launch_bounds__(256) corpus_amax_kernel(const float* __restrict__ C, int ldc, int64_t n, int E, uint32_t* __restrict__ bits) {
    float m = 0.f;
    const int64_t total = n * E;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / E;
        const float v = fabsf(C[r * ldc + (i - r * E)]);
        m = (v < CUDART_INF_F) ? fmaxf(m, v) : m;      // non-finite cells do not pick the scale
    }
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16)); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4)); m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));
}
__global__ void corpus_scale_kernel(float* __restrict__ slot) {   // slot: amax bits in, scale out
    slot[0] = pow2_scale_for(__uint_as_float(reinterpret_cast<uint32_t*>(slot)[0]));
}

// Cp[pos][:] = prepared(C[orig(pos)][:]) (H: fp16 of the scaled row; else TF32-rounded fp32); norms[pos] = ||C[orig(pos)]||
// (unscaled, rounded up a hair); norms[n..n_pad) = 0.
template <bool H>
__global__ void __launch_bounds__(256) prepare_corpus_kernel(const float* __restrict__ C, int ldc, int64_t n, int E, Perm pm,
                                                             void* __restrict__ Cp, float* __restrict__ norms, int32_t* __restrict__ orig_of, int64_t n_pad,
                                                             const float* __restrict__ scale) {
    int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float sc = H ? __ldg(scale) : 1.f;
    for (int64_t pos = warp; pos < n_pad; pos += nwarps) {
        float s = 0.f;
        if (pos < n) {
            const int64_t o = perm_orig(pm, pos);
            for (int k = lane; k < E; k += 32) {
                float v = C[o * ldc + k];
                if (H) reinterpret_cast<__half*>(Cp)[pos * E + k] = __float2half_rn(v * sc);
                else reinterpret_cast<float*>(Cp)[pos * E + k] = tf32_rn(v);
                s = fmaf(v, v, s);
            }
        }
        s = warp_sum(s);
        if (lane == 0) {
            norms[pos] = pos < n ? sqrtf(s) * 1.0001f : 0.f;
            orig_of[pos] = pos < n ? (int32_t)perm_orig(pm, pos) : 0;
        }
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

// norms buffer layout: [n_pad per-row norms][n_pad/32 per-chunk maxima][n_pad int32: original row of every permuted position]
// [32 floats: [0] = the corpus scale s_c (1 for TF32 tiles)]
static inline int64_t scale_slot(int64_t n_pad) { return n_pad + n_pad / 32 + n_pad; }
int launch_prepare(const float* C, int ldc, int64_t n, int E, float* Cp, float* norms, int64_t n_pad, cudaStream_t st) {
    int64_t g = ceil_div(n_pad * 32, 256);
    int64_t cap = (int64_t)sm_count() * 16;
    const unsigned grid = (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
    int32_t* orig_of = reinterpret_cast<int32_t*>(norms + n_pad + n_pad / 32);
    float* slot = norms + scale_slot(n_pad);
    if (idx_half_operands(E)) {
        TT_CUDA_OK(cudaMemsetAsync(slot, 0, 32 * sizeof(float), st));
        int64_t ga = ceil_div(n * E, 256 * 8);
        corpus_amax_kernel<<<(unsigned)(ga > cap ? cap : (ga < 1 ? 1 : ga)), 256, 0, st>>>(C, ldc, n, E, reinterpret_cast<uint32_t*>(slot));
        TT_LAUNCH_OK("corpus_amax_kernel");
        corpus_scale_kernel<<<1, 1, 0, st>>>(slot);
        TT_LAUNCH_OK("corpus_scale_kernel");
        prepare_corpus_kernel<true><<<grid, 256, 0, st>>>(C, ldc, n, E, make_perm(n), Cp, norms, orig_of, n_pad, slot);
    } else {
        prepare_corpus_kernel<false><<<grid, 256, 0, st>>>(C, ldc, n, E, make_perm(n), Cp, norms, orig_of, n_pad, slot);
    }
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
    constexpr bool H = (E >= 64);     // idx_half_operands(E)
    using Cfg = RowPanelCfg<MODE, E, BN, H>;
    { static SmemAttr smem_attr; TT_CUDA_OK(smem_attr.ensure(rowpanel_kernel<MODE, E, BN, H>, Cfg::kSmemBytes)); }
    dim3 grid((unsigned)m_tiles, (unsigned)splits);
    rowpanel_kernel<MODE, E, BN, H><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tmQ, tmC, tmC, p);
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
    size_t q32, eps, thr, thr_exact, alpha, inv_s, gmax, cnt, flags, queue, cand, ccnt, keys, c32, norms, exact, total;
    int ngroups, n_tiles, cap;
    int m_tiles, splits, tps, n_logs, cap_log, fine;
    int n_tiles_s, splits_s, tps_s, k_sel, rank;   // threshold pass: tiles of the sample, its launch shape, the order statistic taken, the rank lambda sits near
};

static IdxLayout layout(int nq, int64_t n, int E, int K, bool need_corpus_copy) {
    IdxLayout L;
    L.n_tiles = (int)ceil_div(n, idx_bn(E));
    // one filter group per (tile, warp half) = BN/2 consecutive rows; per 32-row chunk when that leaves fewer than 8 K groups
    constexpr int64_t kSampleMinRows = 400000;   // below: full threshold pass (measured at 105 k rows: 0.166 ms vs 0.185 ms sampled)
    L.fine = (int64_t)L.n_tiles * idx_halves(E) < (n < kSampleMinRows ? 8 : 32) * (int64_t)K;
    const int gpt = L.fine ? idx_bn(E) / 32 : idx_halves(E);   // groups per tile
    // sample for the threshold pass: a quarter of the tiles (half for large K), more when that leaves fewer than 8 r groups
    // (small corpora keep the full pass: there the rescoring tail, which grows with the rank lambda sits at, outweighs 3/4 of a 50 us pass)
    double f = n < kSampleMinRows ? 1.0 : (K <= 256 ? 0.25 : 0.5);
    for (;;) {
        L.n_tiles_s = (int)ceil_div((int64_t)(L.n_tiles * f), 1);
        if (L.n_tiles_s < 1) L.n_tiles_s = 1;
        if (L.n_tiles_s >= L.n_tiles) { L.n_tiles_s = L.n_tiles; f = 1.0; }
        const double fe = (double)L.n_tiles_s / L.n_tiles;
        L.k_sel = f >= 1.0 ? K : (int)(K * fe + 6.0 * sqrt(K * fe * (1.0 - fe)) + 1.0);
        if (L.k_sel > K) L.k_sel = K;
        L.rank = f >= 1.0 ? K : (int)(L.k_sel / fe) + 1;
        if (f >= 1.0 || ((int64_t)L.n_tiles_s * gpt >= 8 * (int64_t)L.k_sel && L.rank + L.rank / 4 <= 1000)) break;
        f *= 2.0;
    }
    L.ngroups = L.n_tiles_s * gpt;
    L.cap = cand_cap(L.rank);
    L.m_tiles = (int)ceil_div(nq, 128);
    choose_splits(L.m_tiles, L.n_tiles, 2, 64, &L.splits, &L.tps);
    choose_splits(L.m_tiles, L.n_tiles_s, 2, 64, &L.splits_s, &L.tps_s);
    // hit logs of the collect pass: a warp's 32 rows each see ~1.25 K / (splits * halves) qualifying chunks; generous slack because an overflow costs a trip through the exact CUDA-core fallback
    L.n_logs = L.m_tiles * L.splits * 4 * idx_halves(E);   // one log per epilogue warp (32 rows x its share of the columns)
    L.cap_log = g_cap_override > 0 ? 32 * (1 + g_cap_override / (4 * L.splits)) : 32 * (int)ceil_div(8 * (int64_t)L.rank, 2 * L.splits * idx_halves(E)) + 128;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    L.q32 = take((size_t)nq * E * 4);
    L.eps = take((size_t)nq * 4);
    L.thr = take((size_t)nq * 4);
    L.thr_exact = take((size_t)nq * 4);
    L.alpha = take((size_t)nq * 4);
    L.inv_s = take((size_t)nq * 4);
    L.gmax = take((size_t)nq * L.ngroups * 4);
    L.cnt = take((size_t)L.n_logs * 4);
    L.flags = take((size_t)nq * 4);
    L.queue = take((size_t)L.n_logs * L.cap_log * kHitWords * 4);
    L.cand = take((size_t)nq * L.cap * 4);
    L.ccnt = take((size_t)nq * 4);
    L.keys = take((size_t)nq * L.cap * 8);
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
    if (K < 1 || K > 640) return false;                           // sort_topk_kernel orders at most 1024 listed columns
    if (n < 4096) return false;                                   // tiny corpora: the exact kernel is already fast
    if (ceil_div(n, 32) < 4 * (int64_t)K) return false;           // the group filter needs many more groups than K
    return true;
}

size_t index_tc_workspace(int nq, int64_t n, int E, int K, bool need_corpus_copy) {
    if (!(E == 32 || E == 64 || E == 128)) return 0;
    return layout(nq, n, E, K, need_corpus_copy).total + 512;
}

void debug_index_cap(int cap) { tc::g_cap_override = cap; }

// debug: per-stage device times of the next tensor-core index calls (events on the caller's stream; reading them syncs)
static float* g_stage_ms = nullptr;   // host array of 8 floats, accumulated
void debug_index_stages(float* host_ms8) {
    g_stage_ms = host_ms8;
}

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
    float* thr_exact = reinterpret_cast<float*>(base + L.thr_exact);
    float* alpha = reinterpret_cast<float*>(base + L.alpha);
    float* inv_s = reinterpret_cast<float*>(base + L.inv_s);
    float* gmax = reinterpret_cast<float*>(base + L.gmax);
    int32_t* cnt = reinterpret_cast<int32_t*>(base + L.cnt);
    int32_t* flags = reinterpret_cast<int32_t*>(base + L.flags);
    float* queue = reinterpret_cast<float*>(base + L.queue);
    int32_t* cand = reinterpret_cast<int32_t*>(base + L.cand);
    int32_t* ccnt = reinterpret_cast<int32_t*>(base + L.ccnt);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(base + L.keys);
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
    cudaEvent_t ev[8];
    int nev = 0;
    auto mark = [&]() { if (g_stage_ms && nev < 8) { cudaEventCreate(&ev[nev]); cudaEventRecord(ev[nev], st); ++nev; } };
    mark();
    const bool half_ops = idx_half_operands(E);
    const float* corpus_scale = norms + scale_slot((int64_t)TT_INDEX_ROWS_PAD(n));
    const unsigned pgrid = (unsigned)ceil_div((int64_t)nq * 32, 256);
    if (half_ops) prep_queries_kernel<true><<<pgrid, 256, 0, st>>>(Q, ldq, nq, E, q32, eps, alpha, inv_s, corpus_scale);
    else prep_queries_kernel<false><<<pgrid, 256, 0, st>>>(Q, ldq, nq, E, q32, eps, alpha, inv_s, corpus_scale);
    TT_LAUNCH_OK("prep_queries_kernel");

    CUtensorMap tmQ, tmC;
    int rc = half_ops ? make_tmap_2d_f16(&tmQ, q32, nq, E, E, 128) : make_tmap_2d(&tmQ, q32, nq, E, E, 128);
    if (rc) return rc;
    rc = half_ops ? make_tmap_2d_f16(&tmC, c32, n, E, E, idx_bn(E)) : make_tmap_2d(&tmC, c32, n, E, E, idx_bn(E));   // the prepared copy is dense (ld = E)
    if (rc) return rc;
    const int m_tiles = L.m_tiles, splits = L.splits, tps = L.tps;
    RowPanelParams p{};
    p.nR = nq; p.nT = (int)n; p.n_tiles = L.n_tiles; p.tiles_per_split = tps; p.rowv = eps; p.rowv2 = thr; p.colv2 = norms; p.gnorm = norms + TT_INDEX_ROWS_PAD(n); p.d = -(1 << 30); p.fine_groups = L.fine;
    p.out0 = gmax; p.out1 = nullptr; p.out2 = nullptr; p.ld_out = L.ngroups; p.trace = nullptr;
    mark();
    {   // threshold pass over the sample: the first n_tiles_s tiles of the permuted copy
        RowPanelParams ps = p;
        const int64_t rows_s = (int64_t)L.n_tiles_s * idx_bn(E);
        ps.nT = (int)(rows_s < n ? rows_s : n); ps.n_tiles = L.n_tiles_s; ps.tiles_per_split = L.tps_s;
        rc = launch_idx_e<kIndex>(E, tmQ, tmC, ps, m_tiles, L.splits_s, st, "rowpanel_kernel<index>");
        if (rc) return rc;
    }
    mark();
    select_threshold_kernel<<<(unsigned)nq, kSelectThreads, 0, st>>>(gmax, L.ngroups, L.ngroups, L.k_sel, thr, thr_exact, alpha, inv_s, flags, ccnt);
    TT_LAUNCH_OK("select_threshold_kernel");
    mark();
    p.out0 = queue;
    p.out1 = reinterpret_cast<float*>(cnt);
    p.out2 = reinterpret_cast<float*>(flags);
    p.ld_out = L.cap_log;
    rc = launch_idx_e<kCollect>(E, tmQ, tmC, p, m_tiles, splits, st, "rowpanel_kernel<collect>");
    if (rc) return rc;
    mark();
    {
        const int32_t* orig_of = reinterpret_cast<const int32_t*>(norms + TT_INDEX_ROWS_PAD(n) + TT_INDEX_ROWS_PAD(n) / 32);
        const dim3 cgrid((unsigned)ceil_div(L.cap_log, kTestThreads / 8), (unsigned)L.n_logs);
        column_test_kernel<<<cgrid, kTestThreads, 0, st>>>(queue, cnt, L.cap_log, n, L.cap, cand, ccnt);
        TT_LAUNCH_OK("column_test_kernel");
#define TT_SCORE(EE)                                                                                                        \
    exact_score_kernel<EE><<<dim3((unsigned)ceil_div(L.cap, ScoreCfg<EE>::kThreads), (unsigned)nq), ScoreCfg<EE>::kThreads, 0, st>>>( \
        Q, ldq, C, ldc, cand, ccnt, L.cap, thr_exact, orig_of, keys)
        if (E == 32) TT_SCORE(32); else if (E == 64) TT_SCORE(64); else TT_SCORE(128);
#undef TT_SCORE
        TT_LAUNCH_OK("exact_score_kernel");
        const unsigned tgrid = (unsigned)ceil_div(nq, kSortWarps);
        const int need_k = (int)((int64_t)K < n ? K : n);   // rows a query must end with (fewer: the sampled threshold was too high -> exact fallback)
        if (L.cap <= 256) sort_topk_kernel<false><<<tgrid, 32 * kSortWarps, 0, st>>>(keys, ccnt, nq, L.cap, K, need_k, idx_base, out_s, out_i, flags);
        else sort_topk_kernel<true><<<tgrid, 32 * kSortWarps, 0, st>>>(keys, ccnt, nq, L.cap, K, need_k, idx_base, out_s, out_i, flags);
        TT_LAUNCH_OK("sort_topk_kernel");
    }
    mark();
    // queries whose candidate list overflowed: exact CUDA-core path, decided on the device (no host sync)
    rc = index_exact(Q, ldq, C, ldc, nq, n, E, K, idx_base, out_s, out_i, base + L.exact, exact_bytes, st, flags);
    mark();
    if (g_stage_ms && nev > 1) {   // debug only: prep | filter | select | collect | rescore | fallback
        cudaEventSynchronize(ev[nev - 1]);
        for (int i = 0; i + 1 < nev; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            g_stage_ms[i] += ms;
        }
        for (int i = 0; i < nev; ++i) cudaEventDestroy(ev[i]);
    }
    return rc;
}

}  // namespace tt

extern "C" int tt_debug_index_layout(int nq, int64_t n, int E, int K, int have_corpus_prepared, int64_t* out8) {
    if (!out8 || nq <= 0 || n <= 0 || K <= 0) return TT_ERR_ARG;
    const tt::tc::IdxLayout L = tt::tc::layout(nq, n, E, K, !have_corpus_prepared);
    out8[0] = (int64_t)L.flags; out8[1] = (int64_t)L.ccnt; out8[2] = (int64_t)L.cnt; out8[3] = L.cap;
    out8[4] = L.cap_log; out8[5] = L.n_logs; out8[6] = L.rank; out8[7] = L.ngroups;
    return TT_OK;
}

