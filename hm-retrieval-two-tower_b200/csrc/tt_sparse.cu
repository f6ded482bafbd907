// tt_sparse.cu -- deterministic sparse embedding-gradient update.
//
// Replaces (reference file:line): two_tower_model.py:124 (optimizer.minimize) for the Embedding
// tables of input_layer.py:37-40, i.e. tf-keras 2.16 legacy OptimizerV2:
//   IndexedSlices (duplicates not merged) -> _deduplicate_indexed_slices (Unique + UnsortedSegmentSum)
//   -> ResourceSparseApplyAdagradV2 / Adam._resource_apply_sparse          (optimizer_factory.py:15-18)
//
// Determinism: no floating-point atomics anywhere.  (id, position) pairs are ordered by a stable LSD
// radix sort (8-bit digits, only as many passes as the table's row count needs); each run of equal
// ids is then summed by ONE warp in ascending position order and the row is written once.
// The summation order equals oracle/two_tower_oracle.py:dedup_indexed_slices, so the updated rows are
// bit-identical to the oracle given identical gradients.
#include "tt_common.cuh"

namespace tt {

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kItems = 8;                          // keys per thread
constexpr int kTile = kSortThreads * kItems;       // 2048 keys per CTA
constexpr int kRadix = 256;                        // 8-bit digits: the bandwidth regime (hundreds of thousands of ids)
// Training batches (<= 2^14 ids per table) are launch- and latency-bound instead: three kernels per pass, ~16 us per pass at 8192 ids
// however few the keys.  They use 11-bit digits: a 1.37 M-row table sorts in TWO passes instead of three, and each pass fits one of
// the two windows of a train step in which no persistent softmax kernel runs (under the tower forward / under the tower backward).
constexpr int kWideBits = 11;
constexpr int kWideMaxN = 1 << 14;   // beyond, the flat scan and the 2048-bin prologues cost more than the third pass saves (measured at 65536 ids: 0.30 -> 0.375 ms per 8-GPU step)
constexpr int kRadixMax = 1 << kWideBits;

struct JobArr {
    tt_sparse_job j[TT_MAX_JOBS];
    int n;
};

struct SortPlan {  // per-job workspace pointers
    uint32_t* keys[2];
    int32_t* vals[2];
    uint32_t* hist;  // [kRadix][ntiles]
    float* partL;    // [nblocks][e] piece sums of runs continuing from an earlier block
    float* partR;    // [nblocks][e] piece sums of runs that begin in the block and continue
    float* stage;    // row-sharded tables: [n][e] gradient rows of the owned entries in sorted order (pulled from peer memory), else null
    int32_t n;       // elements
    int32_t ntiles;
    int32_t npass;
    int32_t bits;    // digit width of this call: 8, or 11 for small batches (then the scan is flat: hist holds global offsets)
    int32_t vec;     // floats per lane access in the segmented reduce (1, 2 or 4): widest that divides e and the alignment of every row base
    int32_t gvec;    // 1: the gradient sources are aligned for `vec`-wide loads too (a feature's dX slice may start at any column)
    int32_t svec;    // 1: staging copies 16-byte pieces (sources, staging rows and e allow it), else 4-byte pieces
};
struct PlanArr {
    SortPlan p[TT_MAX_JOBS];
    int n;
};

static int passes_for_rows(int rows, int digit_bits) {
    int bits = 1;
    while (bits < 31 && (1ll << bits) < (long long)rows) ++bits;
    return (bits + digit_bits - 1) / digit_bits;
}

// row-sharded tables: number of local rows of every shard; also the key given to entries this rank does not own
__host__ __device__ __forceinline__ uint32_t skip_key(const tt_sparse_job& job) {
    return (uint32_t)((job.rows + job.shard_world - 1) / job.shard_world);
}

__device__ __forceinline__ uint32_t load_key(const tt_sparse_job& job, int i) {
    int src = i / job.n_per_src, r = i - src * job.n_per_src;
    int id = __ldg(job.ids[src] + r);
    if ((unsigned)id >= (unsigned)job.rows) id = 0;  // out-of-range -> OOV row, as in the forward gather
    if (job.shard_world > 1) {                       // row-sharded table: local row, or the skip key (sorts last) for other ranks' rows
        const unsigned g = (unsigned)job.shard_world, q = (unsigned)id / g;
        return ((unsigned)id - q * g == (unsigned)job.shard_rank) ? q : skip_key(job);
    }
    return (uint32_t)id;
}

// ---- pass kernels (blockIdx.y = job, blockIdx.x = tile) ---------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const __grid_constant__ JobArr jobs, const __grid_constant__ PlanArr plans,
                                                                 int pass) {
    constexpr int RADIX = 1 << BITS;
    const SortPlan& pl = plans.p[blockIdx.y];
    if (pass >= pl.npass || (int)blockIdx.x >= pl.ntiles) return;
    __shared__ uint32_t s_hist[RADIX];
    for (int d = threadIdx.x; d < RADIX; d += kSortThreads) s_hist[d] = 0;
    __syncthreads();
    const int base = blockIdx.x * kTile;
    const uint32_t* kin = pl.keys[pass & 1];
    const int shift = pass * BITS;
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        int i = base + it * kSortThreads + threadIdx.x;
        if (i < pl.n) {
            uint32_t key = pass == 0 ? load_key(jobs.j[blockIdx.y], i) : kin[i];
            atomicAdd(&s_hist[(key >> shift) & (RADIX - 1)], 1u);  // integer: order-independent
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < RADIX; d += kSortThreads) pl.hist[(size_t)d * pl.ntiles + blockIdx.x] = s_hist[d];
}

// Flat exclusive scan of hist viewed as one array of radix*ntiles entries (digit-major) by ONE CTA per job: small batches only
// (radix * ntiles <= 2^17 entries; 8192 at a batch of 8192 ids).  hist then holds global offsets: the scatter adds nothing.
__global__ void __launch_bounds__(1024) sort_scan_flat_kernel(const __grid_constant__ PlanArr plans, int pass, int radix) {
    const SortPlan& pl = plans.p[blockIdx.x];
    if (pass >= pl.npass || pl.ntiles == 0) return;
    __shared__ uint32_t s_warp[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = radix * pl.ntiles;                       // a multiple of 2048: 16-byte pieces, workspace slices are 256-byte aligned
    const int per = ((total + 1023) / 1024 + 3) / 4 * 4;       // entries per thread, a multiple of 4
    const int b = min(total, (int)threadIdx.x * per), e = min(total, b + per);
    uint4* h4 = reinterpret_cast<uint4*>(pl.hist);
    uint32_t sum = 0;
#pragma unroll 4
    for (int i = b; i < e; i += 4) {                           // independent loads: they pipeline
        const uint4 v = h4[i >> 2];
        sum += v.x + v.y + v.z + v.w;
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t v = s_warp[lane];
        uint32_t w = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += u;
        }
        s_warp[lane] = w - v;
    }
    __syncthreads();
    uint32_t run = s_warp[warp] + inc - sum;
#pragma unroll 4
    for (int i = b; i < e; i += 4) {
        const uint4 v = h4[i >> 2];
        uint4 w;
        w.x = run; w.y = run + v.x; w.z = w.y + v.y; w.w = w.z + v.z;
        h4[i >> 2] = w;
        run = w.w + v.w;
    }
}

// Exclusive scan of hist viewed as one array of kRadix*ntiles entries (digit-major), in two levels so that it is not one CTA's
// serial walk (round 1: 21 us per pass at 2^20 ids, as long as the scatter): CTA (digit, job) scans ITS row of ntiles counts in
// place and leaves the row total in hist[kRadix * ntiles + digit]; the scatter kernel adds the exclusive prefix of the 256 row
// totals (a 256-thread scan in its prologue).
__global__ void __launch_bounds__(256) sort_scan_kernel(const __grid_constant__ PlanArr plans, int pass) {
    const SortPlan& pl = plans.p[blockIdx.y];
    if (pass >= pl.npass || pl.ntiles == 0) return;
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_run;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* row = pl.hist + (size_t)blockIdx.x * pl.ntiles;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int t0 = 0; t0 < pl.ntiles; t0 += 256) {
        const int t = t0 + threadIdx.x;
        const uint32_t v = t < pl.ntiles ? row[t] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t before = s_run;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        if (t < pl.ntiles) row[t] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 255) s_run = before + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) pl.hist[(size_t)kRadix * pl.ntiles + blockIdx.x] = s_run;
}

template <int BITS>
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const __grid_constant__ JobArr jobs, const __grid_constant__ PlanArr plans,
                                                                    int pass) {
    constexpr int RADIX = 1 << BITS;
    const SortPlan& pl = plans.p[blockIdx.y];
    if (pass >= pl.npass || (int)blockIdx.x >= pl.ntiles) return;
    __shared__ uint16_t s_wcount[kSortWarps][RADIX];   // per warp: keys of each digit (<= 256 per warp), then the exclusive prefix over warps
    __shared__ uint32_t s_base[RADIX];                 // global position of the tile's first key of each digit
    for (int i = threadIdx.x; i < kSortWarps * RADIX; i += kSortThreads) (&s_wcount[0][0])[i] = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int base = blockIdx.x * kTile + warp * (kTile / kSortWarps);  // each warp owns 256 consecutive keys
    const uint32_t* kin = pl.keys[pass & 1];
    const int32_t* vin = pl.vals[pass & 1];
    uint32_t* kout = pl.keys[(pass + 1) & 1];
    int32_t* vout = pl.vals[(pass + 1) & 1];
    const int shift = pass * BITS;
    uint32_t key[kItems];
    int32_t val[kItems];
    uint32_t rank[kItems];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        int i = base + it * 32 + lane;
        bool ok = i < pl.n;
        key[it] = 0;
        val[it] = 0;
        if (ok) {
            if (pass == 0) { key[it] = load_key(jobs.j[blockIdx.y], i); val[it] = i; }
            else { key[it] = kin[i]; val[it] = vin[i]; }
        }
        uint32_t digit = ok ? ((key[it] >> shift) & (RADIX - 1)) : 0xffffffffu;  // invalid lanes group together
        uint32_t peers = __match_any_sync(0xffffffffu, digit);
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (ok && lane == leader) old = s_wcount[warp][digit];
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[it] = old + __popc(peers & lt);
        if (ok && lane == leader) s_wcount[warp][digit] = (uint16_t)(old + __popc(peers));
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over warps, and the global base of (digit, tile)
    uint32_t dbase = 0;
    if constexpr (BITS == 8) {
        if (pl.bits == 8) {   // two-level scan: the exclusive prefix of the digit totals is formed here (256 values, one per thread)
            __shared__ uint32_t s_dw[kSortWarps];
            const uint32_t tot = pl.hist[(size_t)RADIX * pl.ntiles + threadIdx.x];
            uint32_t inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            if (lane == 31) s_dw[warp] = inc;
            __syncthreads();
            dbase = inc - tot;
            for (int w = 0; w < warp; ++w) dbase += s_dw[w];
        }
    }
    for (int d = threadIdx.x; d < RADIX; d += kSortThreads) {
        s_base[d] = dbase + pl.hist[(size_t)d * pl.ntiles + blockIdx.x];
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = s_wcount[w][d];
            s_wcount[w][d] = (uint16_t)run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        int i = base + it * 32 + lane;
        if (i < pl.n) {
            const uint32_t d = (key[it] >> shift) & (RADIX - 1);
            uint32_t pos = s_base[d] + s_wcount[warp][d] + rank[it];
            kout[pos] = key[it];
            vout[pos] = val[it];
        }
    }
}

// ---- segmented reduce + row update: one warp per run of equal ids --------------------------------------
__device__ __forceinline__ const float* grad_row(const tt_sparse_job& job, int pos) {
    int src = pos / job.n_per_src, r = pos - src * job.n_per_src;
    return job.grad[src] + (int64_t)r * job.grad_ld[src];
}

enum { kModeAdagrad = 0, kModeAdamMoments = 1 };
constexpr int kMaxColsPerLane = 8;   // embedding widths up to 256

// The stably sorted (id, position) array is cut into blocks of 32 entries; one warp owns one block.
// A run of equal ids is summed piecewise: each block sums its piece in ascending position order
// (sequential fp32 adds), and pieces are added in block order.  Runs that stay inside one block (the
// common case) are applied immediately; a run crossing block boundaries leaves per-block partial sums
// (partL: the piece that continues a run begun in an earlier block; partR: the piece of a run that begins
// here and continues) which phase B adds up from the run's head block.  The order depends only on the
// sorted array, so the result is deterministic and equals oracle/two_tower_oracle.py:dedup_indexed_slices.
// Body of phase A for one block, NC columns per lane, RB runs in flight.  The loads of a batch of RB runs
// (first gradient row of each run plus the table / slot values their update will need) are issued together so a
// block of 32 distinct ids costs ~32/RB memory round trips instead of 64.
template <int NC, int RB>
__device__ __forceinline__ void block_body(const tt_sparse_job& job, int kMode, uint32_t key, int pos, uint32_t heads, int cnt, bool cont_in,
                                           bool cont_out, float* __restrict__ partL, float* __restrict__ partR, int lane, float lr, float eps,
                                           float omb1, float omb2, const float* __restrict__ stage) {
    // stage != null: the block's gradient rows were copied (from the producing GPUs' dX buffers) to stage[lane index], contiguous
    const int e = job.e;
    while (heads) {
        int s[RB], en[RB];
        uint32_t id[RB];
        int nb = 0;
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            s[u] = 0; en[u] = 0; id[u] = 0;
            if (heads) {
                s[u] = __ffs(heads) - 1;
                heads &= heads - 1;
                en[u] = heads ? (__ffs(heads) - 1) : cnt;
                nb = u + 1;
            }
            id[u] = __shfl_sync(0xffffffffu, key, s[u]);
        }
        float g[RB][NC], a[RB][NC], w[RB][NC];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int p0 = __shfl_sync(0xffffffffu, pos, s[u]);
            const float* r0 = stage ? stage + (int64_t)s[u] * e : grad_row(job, p0);
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) {
                const int c = lane + 32 * ci;
                g[u][ci] = 0.f; a[u][ci] = 0.f; w[u][ci] = 0.f;
                if (u < nb && c < e) {
                    const int64_t o = (int64_t)id[u] * e + c;
                    g[u][ci] = __ldg(r0 + c);
                    a[u][ci] = job.slot0[o];
                    w[u][ci] = (kMode == kModeAdagrad) ? job.table[o] : job.slot1[o];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            for (int j = s[u] + 1; j < en[u]; ++j) {   // duplicates of the id inside this block: sequential, position order
                const int pj = __shfl_sync(0xffffffffu, pos, j);
                const float* r0 = stage ? stage + (int64_t)j * e : grad_row(job, pj);
#pragma unroll
                for (int ci = 0; ci < NC; ++ci) {
                    const int c = lane + 32 * ci;
                    if (c < e) g[u][ci] = __fadd_rn(g[u][ci], __ldg(r0 + c));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            if (u >= nb) continue;
            const bool to_l = (s[u] == 0) && cont_in;
            const bool to_r = (en[u] == cnt) && cont_out && !to_l;
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) {
                const int c = lane + 32 * ci;
                if (c >= e) continue;
                const float gv = g[u][ci];
                if (to_l) partL[c] = gv;          // continuation piece (may also continue further)
                else if (to_r) partR[c] = gv;     // run begins here and continues
                else {                            // run lives entirely in this block: apply with the prefetched state
                    const int64_t o = (int64_t)id[u] * e + c;
                    if (kMode == kModeAdagrad) {
                        const float an = __fadd_rn(a[u][ci], __fmul_rn(gv, gv));
                        job.slot0[o] = an;
                        job.table[o] = __fsub_rn(w[u][ci], __fdiv_rn(__fmul_rn(gv, lr), __fadd_rn(__fsqrt_rn(an), eps)));
                    } else {
                        job.slot0[o] = __fadd_rn(a[u][ci], __fmul_rn(gv, omb1));
                        job.slot1[o] = __fadd_rn(w[u][ci], __fmul_rn(__fmul_rn(gv, gv), omb2));
                    }
                }
            }
        }
    }
}

// ---- the same body with 64 / 128-bit accesses -----------------------------------------------------------------------------------
// A lane owns VW CONSECUTIVE columns per chunk of 32 * VW columns, so a 64-wide row is one 8-byte access per lane and a 128-wide row
// one 16-byte access (ncu, 2^20 ids into a 1.37 M x 64 table: the scalar body above executes ~250 warp instructions per run -- address
// arithmetic, the position / n_per_src division, per-column predicates -- and is issue-bound at 41 % of DRAM throughput).  Every column
// is still summed sequentially in position order, so the result is bit-identical to the scalar body and to the oracle.
template <int VW>
__device__ __forceinline__ void vload(const float* p, float (&o)[VW]) {
    if constexpr (VW == 4) { const float4 t = *reinterpret_cast<const float4*>(p); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
    else if constexpr (VW == 2) { const float2 t = *reinterpret_cast<const float2*>(p); o[0] = t.x; o[1] = t.y; }
    else o[0] = *p;
}
template <int VW>
__device__ __forceinline__ void vload_nc(const float* p, float (&o)[VW]) {
    if constexpr (VW == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w; }
    else if constexpr (VW == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); o[0] = t.x; o[1] = t.y; }
    else o[0] = __ldg(p);
}
template <int VW>
__device__ __forceinline__ void vstore(float* p, const float (&v)[VW]) {
    if constexpr (VW == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (VW == 2) *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    else *p = v[0];
}

template <int VW, int NC, int RB>
__device__ __forceinline__ void block_body_vec(const tt_sparse_job& job, int kMode, uint32_t key, int pos, uint32_t heads, int cnt, bool cont_in,
                                               bool cont_out, float* __restrict__ partL, float* __restrict__ partR, int lane, float lr, float eps,
                                               float omb1, float omb2, const float* __restrict__ stage, bool gvec) {
    const int e = job.e;
    float* const s0 = job.slot0;                                              // Adagrad accumulator / Adam m
    float* const s1 = (kMode == kModeAdagrad) ? job.table : job.slot1;        // the second array of the update: weights / Adam v
    const bool single = job.nsrc == 1;                                        // one gradient source: no position / n_per_src division
    const float* const g0 = job.grad[0];
    const int64_t ld0 = job.grad_ld[0];
    int c0[NC];
    bool cv[NC];
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
        c0[ci] = (ci * 32 + lane) * VW;
        cv[ci] = c0[ci] < e;                                                  // e % VW == 0: the VW columns of a lane are valid together
    }
    auto grow = [&](int j, int p) -> const float* {
        return stage ? stage + (int64_t)j * e : (single ? g0 + (int64_t)p * ld0 : grad_row(job, p));
    };
    auto gload = [&](const float* r, float (&o)[VW]) {
        if (gvec) vload_nc<VW>(r, o);
        else {
#pragma unroll
            for (int v = 0; v < VW; ++v) o[v] = __ldg(r + v);
        }
    };
    while (heads) {
        int s[RB], en[RB];
        uint32_t id[RB];
        int nb = 0;
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            s[u] = 0; en[u] = 0;
            if (heads) {
                s[u] = __ffs(heads) - 1;
                heads &= heads - 1;
                en[u] = heads ? (__ffs(heads) - 1) : cnt;
                nb = u + 1;
            }
            id[u] = __shfl_sync(0xffffffffu, key, s[u]);
        }
        float g[RB][NC][VW], a[RB][NC][VW], w[RB][NC][VW];
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            const int p0 = __shfl_sync(0xffffffffu, pos, s[u]);
            const float* r0 = grow(s[u], p0);
            const int64_t ro = (int64_t)id[u] * e;
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) {
#pragma unroll
                for (int v = 0; v < VW; ++v) { g[u][ci][v] = 0.f; a[u][ci][v] = 0.f; w[u][ci][v] = 0.f; }
                if (u < nb && cv[ci]) {
                    gload(r0 + c0[ci], g[u][ci]);
                    vload<VW>(s0 + ro + c0[ci], a[u][ci]);
                    vload<VW>(s1 + ro + c0[ci], w[u][ci]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            for (int j = s[u] + 1; j < en[u]; ++j) {   // duplicates of the id inside this block: sequential, position order
                const int pj = __shfl_sync(0xffffffffu, pos, j);
                const float* r = grow(j, pj);
#pragma unroll
                for (int ci = 0; ci < NC; ++ci) {
                    if (!cv[ci]) continue;
                    float t[VW];
                    gload(r + c0[ci], t);
#pragma unroll
                    for (int v = 0; v < VW; ++v) g[u][ci][v] = __fadd_rn(g[u][ci][v], t[v]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            if (u >= nb) continue;
            const bool to_l = (s[u] == 0) && cont_in;
            const bool to_r = (en[u] == cnt) && cont_out && !to_l;
            const int64_t ro = (int64_t)id[u] * e;
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) {
                if (!cv[ci]) continue;
                if (to_l) vstore<VW>(partL + c0[ci], g[u][ci]);          // continuation piece (may also continue further)
                else if (to_r) vstore<VW>(partR + c0[ci], g[u][ci]);     // run begins here and continues
                else {                                                   // run lives entirely in this block: apply with the prefetched state
                    float x0[VW], x1[VW];
#pragma unroll
                    for (int v = 0; v < VW; ++v) {
                        const float gv = g[u][ci][v];
                        if (kMode == kModeAdagrad) {
                            x0[v] = __fadd_rn(a[u][ci][v], __fmul_rn(gv, gv));
                            x1[v] = __fsub_rn(w[u][ci][v], __fdiv_rn(__fmul_rn(gv, lr), __fadd_rn(__fsqrt_rn(x0[v]), eps)));
                        } else {
                            x0[v] = __fadd_rn(a[u][ci][v], __fmul_rn(gv, omb1));
                            x1[v] = __fadd_rn(w[u][ci][v], __fmul_rn(__fmul_rn(gv, gv), omb2));
                        }
                    }
                    vstore<VW>(s0 + ro + c0[ci], x0);
                    vstore<VW>(s1 + ro + c0[ci], x1);
                }
            }
        }
    }
}

// Row-sharded tables: the gradient row of every OWNED entry is copied from wherever it was produced (the dX block of another GPU,
// peer-mapped over NVLink) into the local staging array, in sorted order; the segmented reduce then streams contiguous local rows.
// One THREAD per piece of a row (16 bytes when the sources allow it, else 4): every thread walks the same short dependent chain
// (key, position, piece) and all of them are in flight at once, so the NVLink latency is paid once per launch, not once per row
// of a warp (the round-1 warp-per-entry form took 17 us for 8192 owned rows; a warp-per-32-entries form 37 us).
template <typename T>
__global__ void __launch_bounds__(256) sparse_stage_kernel(const __grid_constant__ JobArr jobs, const __grid_constant__ PlanArr plans) {
    const SortPlan& pl = plans.p[blockIdx.y];
    const tt_sparse_job& job = jobs.j[blockIdx.y];
    if (pl.stage == nullptr) return;
    constexpr int W = (int)(sizeof(T) / sizeof(float));
    if ((pl.svec != 0) != (W == 4)) return;          // this job is served by the other instantiation
    const int per_row = job.e / W;
    const uint32_t* ks = pl.keys[pl.npass & 1];
    const int32_t* vs = pl.vals[pl.npass & 1];
    const uint32_t skip = skip_key(job);
    const int64_t total = (int64_t)pl.n * per_row;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // the owned entries are a prefix of the sorted array (the skip key sorts last): a thread that meets a foreign entry is done
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += stride) {
        const int j = (int)(p / per_row), c = (int)(p - (int64_t)j * per_row);
        if (ks[j] == skip) break;
        reinterpret_cast<T*>(pl.stage + (int64_t)j * job.e)[c] = __ldg(reinterpret_cast<const T*>(grad_row(job, vs[j])) + c);
    }
}

// phase A: grid (ceil(max_blocks / 8), njobs), 8 warps per CTA, one warp per block of 32 sorted entries
// kWide: eight runs in flight per warp for narrow tables (more registers, fewer round trips) -- the latency-bound regime of a
// training step's batch; the bandwidth-bound regime (hundreds of thousands of entries) keeps four and the higher occupancy
template <bool kWide>
__global__ void __launch_bounds__(256, kWide ? 2 : 4) sparse_block_kernel(const __grid_constant__ JobArr jobs, const __grid_constant__ PlanArr plans, int kMode,
                                                           float lr, float eps, float omb1, float omb2) {
    const SortPlan& pl = plans.p[blockIdx.y];
    const tt_sparse_job& job = jobs.j[blockIdx.y];
    const uint32_t* ks = pl.keys[pl.npass & 1];
    const int32_t* vs = pl.vals[pl.npass & 1];
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nblk = (pl.n + 31) >> 5;
    if (b >= nblk) return;
    const int base = b << 5;
    int cnt = min(32, pl.n - base);
    const uint32_t key = lane < cnt ? ks[base + lane] : 0xffffffffu;
    if (job.shard_world > 1) {   // entries of rows other ranks own carry the skip key and sort last: keep the owned prefix
        cnt = __popc(__ballot_sync(0xffffffffu, lane < cnt && key != skip_key(job)));
        if (cnt == 0) return;
    }
    const int pos = lane < cnt ? vs[base + lane] : 0;
    const uint32_t kprev_lane = __shfl_up_sync(0xffffffffu, key, 1);
    const bool cont_in = base > 0 && ks[base - 1] == __shfl_sync(0xffffffffu, key, 0);           // first run began earlier
    const uint32_t last_key = __shfl_sync(0xffffffffu, key, cnt - 1);
    const bool cont_out = base + 32 < pl.n && cnt == 32 && ks[base + 32] == last_key;            // last run goes on
    const bool head = lane < cnt && (lane == 0 || key != kprev_lane);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const int e = job.e;
    float* partL = pl.partL + (int64_t)b * e;
    float* partR = pl.partR + (int64_t)b * e;
    const float* stage = pl.stage ? pl.stage + (int64_t)base * e : nullptr;
    const int vw = pl.vec, ncv = (e + 32 * vw - 1) / (32 * vw);   // chunks of 32 * vw columns
    const bool gvec = pl.gvec != 0;
    if (ncv == 1 && vw == 1) { block_body_vec<1, 1, kWide ? 8 : 4>(job, kMode, key, pos, heads, cnt, cont_in, cont_out, partL, partR, lane, lr, eps, omb1, omb2, stage, gvec); return; }
    if (ncv == 1 && vw == 2) { block_body_vec<2, 1, kWide ? 8 : 4>(job, kMode, key, pos, heads, cnt, cont_in, cont_out, partL, partR, lane, lr, eps, omb1, omb2, stage, gvec); return; }
    if (ncv == 1 && vw == 4) { block_body_vec<4, 1, 2>(job, kMode, key, pos, heads, cnt, cont_in, cont_out, partL, partR, lane, lr, eps, omb1, omb2, stage, gvec); return; }
    if (ncv == 2 && vw == 4) { block_body_vec<4, 2, 1>(job, kMode, key, pos, heads, cnt, cont_in, cont_out, partL, partR, lane, lr, eps, omb1, omb2, stage, gvec); return; }
    // rows whose width or alignment rules out vector accesses (e.g. e = 64 at an odd column offset of a 4-byte aligned table)
    const int nc = (e + 31) >> 5;
    if (nc <= 1) block_body<1, kWide ? 8 : 4>(job, kMode, key, pos, heads, cnt, cont_in, cont_out, partL, partR, lane, lr, eps, omb1, omb2, stage);
    else if (nc <= 2) block_body<2, kWide ? 8 : 4>(job, kMode, key, pos, heads, cnt, cont_in, cont_out, partL, partR, lane, lr, eps, omb1, omb2, stage);
    else if (nc <= 4) block_body<4, 2>(job, kMode, key, pos, heads, cnt, cont_in, cont_out, partL, partR, lane, lr, eps, omb1, omb2, stage);
    else block_body<8, 1>(job, kMode, key, pos, heads, cnt, cont_in, cont_out, partL, partR, lane, lr, eps, omb1, omb2, stage);
}

// pieces of one boundary-crossing run added in block order, all NC column groups of a lane at once, 16 / NC pieces per round of loads
template <int NC>
__device__ __forceinline__ void combine_body(const tt_sparse_job& job, const SortPlan& pl, int kMode, uint32_t id, int b, int last, int lane, float lr,
                                             float eps, float omb1, float omb2) {
    constexpr int T = 16 / NC;
    const int e = job.e;
    float g[NC];
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
        const int c = lane + 32 * ci;
        g[ci] = c < e ? pl.partR[(int64_t)b * e + c] : 0.f;
    }
    for (int nb0 = b + 1; nb0 <= last; nb0 += T) {
        float v[T][NC];
#pragma unroll
        for (int t = 0; t < T; ++t)
#pragma unroll
            for (int ci = 0; ci < NC; ++ci) {
                const int c = lane + 32 * ci;
                v[t][ci] = (nb0 + t <= last && c < e) ? pl.partL[(int64_t)(nb0 + t) * e + c] : 0.f;
            }
#pragma unroll
        for (int t = 0; t < T; ++t)
            if (nb0 + t <= last) {
#pragma unroll
                for (int ci = 0; ci < NC; ++ci) g[ci] = __fadd_rn(g[ci], v[t][ci]);
            }
    }
    // apply: state loads of all column groups first, then the update
    float a[NC], w[NC];
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
        const int c = lane + 32 * ci;
        const int64_t o = (int64_t)id * e + c;
        a[ci] = c < e ? job.slot0[o] : 0.f;
        w[ci] = c < e ? (kMode == kModeAdagrad ? job.table[o] : job.slot1[o]) : 0.f;
    }
#pragma unroll
    for (int ci = 0; ci < NC; ++ci) {
        const int c = lane + 32 * ci;
        if (c >= e) continue;
        const int64_t o = (int64_t)id * e + c;
        const float gv = g[ci];
        if (kMode == kModeAdagrad) {
            const float an = __fadd_rn(a[ci], __fmul_rn(gv, gv));
            job.slot0[o] = an;
            job.table[o] = __fsub_rn(w[ci], __fdiv_rn(__fmul_rn(gv, lr), __fadd_rn(__fsqrt_rn(an), eps)));
        } else {
            job.slot0[o] = __fadd_rn(a[ci], __fmul_rn(gv, omb1));
            job.slot1[o] = __fadd_rn(w[ci], __fmul_rn(__fmul_rn(gv, gv), omb2));
        }
    }
}

// phase B: the head block of every boundary-crossing run adds the pieces in block order and applies the update
__global__ void __launch_bounds__(256) sparse_combine_kernel(const __grid_constant__ JobArr jobs, const __grid_constant__ PlanArr plans, int kMode,
                                                             float lr, float eps, float omb1, float omb2) {
    const SortPlan& pl = plans.p[blockIdx.y];
    const tt_sparse_job& job = jobs.j[blockIdx.y];
    const uint32_t* ks = pl.keys[pl.npass & 1];
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nblk = (pl.n + 31) >> 5;
    if (b >= nblk) return;
    const int base = b << 5;
    if (base + 32 >= pl.n) return;                       // last block: nothing can continue
    const uint32_t id = ks[base + 31];
    if (ks[base + 32] != id) return;                     // last run does not continue
    if (job.shard_world > 1 && id == skip_key(job)) return;   // rows owned by other ranks
    if (ks[base] == id && base > 0 && ks[base - 1] == id) return;   // whole block is a middle piece: the head block owns it
    // this block holds the head of the run (either mid-block, or at lane 0 with a different predecessor)
    const int e = job.e;
    int last = b + 1;                                    // last block touched by the run
    const bool longer = (int64_t)(last + 1) * 32 < pl.n && ks[(int64_t)(last + 1) * 32] == id;
    while (longer) {                                     // a hot id spanning many blocks: the lanes probe 32 blocks at a time
        const int64_t idx = (int64_t)(last + 1 + lane) * 32;
        const bool same = idx < pl.n && ks[idx] == id;
        const uint32_t m = __ballot_sync(0xffffffffu, same);
        if (m == 0xffffffffu) { last += 32; continue; }
        last += __ffs(~m) - 1;
        break;
    }
    const int nc = (e + 31) >> 5;
    if (nc <= 1) combine_body<1>(job, pl, kMode, id, b, last, lane, lr, eps, omb1, omb2);
    else if (nc <= 2) combine_body<2>(job, pl, kMode, id, b, last, lane, lr, eps, omb1, omb2);
    else if (nc <= 4) combine_body<4>(job, pl, kMode, id, b, last, lane, lr, eps, omb1, omb2);
    else combine_body<8>(job, pl, kMode, id, b, last, lane, lr, eps, omb1, omb2);
}

// Adam whole-table sweeps (legacy Adam is not lazy): phase 0: m *= b1, v *= b2; phase 1: w -= (m*lr_t)/(sqrt(v)+eps)
__global__ void __launch_bounds__(256) adam_sweep_kernel(const __grid_constant__ JobArr jobs, int phase, float b1, float b2, float lr_t,
                                                         float eps) {
    const tt_sparse_job& job = jobs.j[blockIdx.y];
    const int64_t total = (int64_t)(job.shard_world > 1 ? (int)skip_key(job) : job.rows) * job.e;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
        if (phase == 0) {
            job.slot0[i] = __fmul_rn(job.slot0[i], b1);
            job.slot1[i] = __fmul_rn(job.slot1[i], b2);
        } else {
            job.table[i] = __fsub_rn(job.table[i], __fdiv_rn(__fmul_rn(job.slot0[i], lr_t), __fadd_rn(__fsqrt_rn(job.slot1[i]), eps)));
        }
    }
}

// ---- host ---------------------------------------------------------------------------------------------
static size_t part_bytes(int max_n, int max_e) {
    size_t n = (size_t)(max_n > 0 ? max_n : 1);
    return align_up(((n + 31) / 32) * (size_t)(max_e > 0 ? max_e : 1) * sizeof(float), 256);
}
static size_t stage_bytes(int max_n, int max_e) {
    return align_up((size_t)(max_n > 0 ? max_n : 1) * (size_t)(max_e > 0 ? max_e : 1) * sizeof(float), 256);
}
static size_t per_job_bytes(int max_n, int max_e) {
    size_t n = (size_t)(max_n > 0 ? max_n : 1);
    size_t ntiles = (n + kTile - 1) / kTile;
    const size_t radix = max_n <= kWideMaxN ? kRadixMax : kRadix;
    return 4 * align_up(n * 4, 256) + align_up((ntiles + 1) * radix * 4, 256) + 2 * part_bytes(max_n, max_e) + stage_bytes(max_n, max_e);
}

static int make_plans(const tt_sparse_job* jobs, int njobs, void* ws, size_t ws_bytes, JobArr* ja, PlanArr* pa, int* max_tiles,
                      int* max_pass, const char* who) {
    TT_REQUIRE(jobs != nullptr && njobs >= 1 && njobs <= TT_MAX_JOBS, "%s: njobs must be in [1,%d]", who, TT_MAX_JOBS);
    memset(ja, 0, sizeof(*ja));
    memset(pa, 0, sizeof(*pa));
    ja->n = pa->n = njobs;
    int max_n = 0, max_e = 0;
    for (int j = 0; j < njobs; ++j) {
        const tt_sparse_job& jb = jobs[j];
        if (jb.e > max_e) max_e = jb.e;
        TT_REQUIRE(jb.table && jb.slot0 && jb.rows >= 1 && jb.e >= 1, "%s: job %d malformed", who, j);
        TT_REQUIRE(jb.e <= 32 * kMaxColsPerLane, "%s: job %d embedding width %d exceeds %d", who, j, jb.e, 32 * kMaxColsPerLane);
        TT_REQUIRE(jb.nsrc >= 1 && jb.nsrc <= TT_MAX_SRC && jb.n_per_src >= 0, "%s: job %d nsrc/n_per_src out of range", who, j);
        TT_REQUIRE(jb.shard_world <= 1 || (jb.shard_rank >= 0 && jb.shard_rank < jb.shard_world), "%s: job %d shard rank/world out of range", who, j);
        TT_REQUIRE((int64_t)jb.nsrc * jb.n_per_src < (1ll << 31), "%s: job %d too many rows", who, j);
        for (int s = 0; s < jb.nsrc; ++s) TT_REQUIRE(jb.n_per_src == 0 || (jb.ids[s] && jb.grad[s] && jb.grad_ld[s] >= jb.e), "%s: job %d source %d malformed", who, j, s);
        ja->j[j] = jb;
        int n = jb.nsrc * jb.n_per_src;
        if (n > max_n) max_n = n;
    }
    size_t per = per_job_bytes(max_n, max_e);
    TT_REQUIRE(ws != nullptr && ws_bytes >= per * (size_t)njobs, "%s: workspace too small (%zu < %zu)", who, ws_bytes, per * (size_t)njobs);
    *max_tiles = 0;
    *max_pass = 0;
    for (int j = 0; j < njobs; ++j) {
        char* base = reinterpret_cast<char*>(ws) + per * (size_t)j;
        size_t seg = align_up((size_t)(max_n > 0 ? max_n : 1) * 4, 256);
        SortPlan& p = pa->p[j];
        p.keys[0] = reinterpret_cast<uint32_t*>(base);
        p.keys[1] = reinterpret_cast<uint32_t*>(base + seg);
        p.vals[0] = reinterpret_cast<int32_t*>(base + 2 * seg);
        p.vals[1] = reinterpret_cast<int32_t*>(base + 3 * seg);
        p.hist = reinterpret_cast<uint32_t*>(base + 4 * seg);
        const int bits = max_n <= kWideMaxN ? kWideBits : 8;
        size_t hist_bytes = align_up(((size_t)ceil_div(max_n > 0 ? max_n : 1, kTile) + 1) * (size_t)(1 << bits) * 4, 256);   // + the digit totals
        p.partL = reinterpret_cast<float*>(base + 4 * seg + hist_bytes);
        p.partR = reinterpret_cast<float*>(base + 4 * seg + hist_bytes + part_bytes(max_n, max_e));
        p.stage = jobs[j].shard_world > 1 ? reinterpret_cast<float*>(base + 4 * seg + hist_bytes + 2 * part_bytes(max_n, max_e)) : nullptr;
        p.n = jobs[j].nsrc * jobs[j].n_per_src;
        p.ntiles = (int)ceil_div(p.n, kTile);
        p.bits = bits;
        p.npass = passes_for_rows(jobs[j].shard_world > 1 ? (int)skip_key(jobs[j]) + 1 : jobs[j].rows, bits);
        {   // access width of the segmented reduce: e <= 32 -> 1 float per lane, <= 64 -> 2, wider -> 4, narrowed until it divides e
            // and the alignment of every base pointer the kernels add `row * e + column` to
            const tt_sparse_job& jb = jobs[j];
            int want = jb.e <= 32 ? 1 : (jb.e <= 64 ? 2 : 4);
            while (want > 1 && jb.e % want) want >>= 1;
            uintptr_t bits = (uintptr_t)jb.table | (uintptr_t)jb.slot0 | (uintptr_t)jb.slot1 | (uintptr_t)p.partL | (uintptr_t)p.partR | (uintptr_t)p.stage;
            while (want > 1 && (bits % (want * sizeof(float))) != 0) want >>= 1;
            uintptr_t gbits = 0;
            for (int s = 0; s < jb.nsrc; ++s) gbits |= (uintptr_t)jb.grad[s] | ((uintptr_t)(uint32_t)jb.grad_ld[s] * sizeof(float));
            p.vec = want;
            p.gvec = (p.stage != nullptr || (gbits % (want * sizeof(float))) == 0) ? 1 : 0;   // staged rows are read from the workspace
            p.svec = (p.stage != nullptr && jb.e % 4 == 0 && (gbits % 16) == 0 && ((uintptr_t)p.stage % 16) == 0) ? 1 : 0;
        }
        if (p.ntiles > *max_tiles) *max_tiles = p.ntiles;
        if (p.npass > *max_pass) *max_pass = p.npass;
    }
    return TT_OK;
}

static int apply_grid(const PlanArr& pa) {   // one warp per block of 32 sorted entries, 8 warps per CTA
    int max_n = 0;
    for (int j = 0; j < pa.n; ++j) max_n = pa.p[j].n > max_n ? pa.p[j].n : max_n;
    int64_t g = ceil_div(ceil_div((int64_t)max_n, 32), 8);
    return (int)(g < 1 ? 1 : g);
}

static int launch_apply(const JobArr& ja, const PlanArr& pa, int njobs, int mode, float lr, float eps, float omb1, float omb2, cudaStream_t st) {
    dim3 grid((unsigned)apply_grid(pa), (unsigned)njobs);
    bool staged = false;
    int max_n = 0;
    for (int j = 0; j < pa.n; ++j) { staged |= pa.p[j].stage != nullptr; max_n = pa.p[j].n > max_n ? pa.p[j].n : max_n; }
    if (staged) {
        int max_e = 1;
        bool any4 = false, any1 = false;
        for (int j = 0; j < pa.n; ++j) {
            if (pa.p[j].stage == nullptr) continue;
            max_e = ja.j[j].e > max_e ? ja.j[j].e : max_e;
            (pa.p[j].svec ? any4 : any1) = true;
        }
        // a thread per 16-byte (4-byte) piece of the entries the rank can own: about 1 / world of the sorted array, grid-stride beyond
        int64_t sg = ceil_div((int64_t)(max_n > 0 ? max_n : 1) * max_e / 4, 256);
        if (sg > 16 * (int64_t)sm_count()) sg = 16 * (int64_t)sm_count();
        if (sg < 1) sg = 1;
        if (any4) { sparse_stage_kernel<float4><<<dim3((unsigned)sg, (unsigned)njobs), 256, 0, st>>>(ja, pa); TT_LAUNCH_OK("sparse_stage_kernel<16B>"); }
        if (any1) { sparse_stage_kernel<float><<<dim3((unsigned)(4 * sg > 16 * (int64_t)sm_count() ? 16 * (int64_t)sm_count() : 4 * sg), (unsigned)njobs), 256, 0, st>>>(ja, pa); TT_LAUNCH_OK("sparse_stage_kernel<4B>"); }
    }
    if (max_n <= (1 << 17)) sparse_block_kernel<true><<<grid, 256, 0, st>>>(ja, pa, mode, lr, eps, omb1, omb2);
    else sparse_block_kernel<false><<<grid, 256, 0, st>>>(ja, pa, mode, lr, eps, omb1, omb2);
    TT_LAUNCH_OK("sparse_block_kernel");
    sparse_combine_kernel<<<grid, 256, 0, st>>>(ja, pa, mode, lr, eps, omb1, omb2);
    TT_LAUNCH_OK("sparse_combine_kernel");
    return TT_OK;
}

}  // namespace tt

using namespace tt;

extern "C" {

size_t tt_sparse_workspace_bytes(int njobs, int max_n, int max_e) {
    if (njobs < 1) njobs = 1;
    return per_job_bytes(max_n, max_e) * (size_t)njobs + 256;
}

int tt_sparse_sort(const tt_sparse_job* jobs, int njobs, void* ws, size_t ws_bytes, void* stream) {
    return tt_sparse_sort_passes(jobs, njobs, ws, ws_bytes, 0, 1 << 30, stream);
}

int tt_sparse_sort_passes(const tt_sparse_job* jobs, int njobs, void* ws, size_t ws_bytes, int first_pass, int end_pass, void* stream) {
    JobArr ja;
    PlanArr pa;
    int max_tiles = 0, max_pass = 0;
    int rc = make_plans(jobs, njobs, ws, ws_bytes, &ja, &pa, &max_tiles, &max_pass, "tt_sparse_sort");
    if (rc) return rc;
    TT_REQUIRE(first_pass >= 0 && end_pass >= first_pass, "tt_sparse_sort_passes: bad pass range");
    if (max_tiles == 0) return TT_OK;
    cudaStream_t st = as_stream(stream);
    if (end_pass < max_pass) max_pass = end_pass;
    const bool wide = pa.p[0].bits == kWideBits;
    for (int pass = first_pass; pass < max_pass; ++pass) {
        dim3 grid((unsigned)max_tiles, (unsigned)njobs);
        if (wide) {
            sort_hist_kernel<kWideBits><<<grid, kSortThreads, 0, st>>>(ja, pa, pass);
            TT_LAUNCH_OK("sort_hist_kernel");
            sort_scan_flat_kernel<<<(unsigned)njobs, 1024, 0, st>>>(pa, pass, kRadixMax);
            TT_LAUNCH_OK("sort_scan_flat_kernel");
            sort_scatter_kernel<kWideBits><<<grid, kSortThreads, 0, st>>>(ja, pa, pass);
            TT_LAUNCH_OK("sort_scatter_kernel");
        } else {
            sort_hist_kernel<8><<<grid, kSortThreads, 0, st>>>(ja, pa, pass);
            TT_LAUNCH_OK("sort_hist_kernel");
            sort_scan_kernel<<<dim3(kRadix, (unsigned)njobs), 256, 0, st>>>(pa, pass);
            TT_LAUNCH_OK("sort_scan_kernel");
            sort_scatter_kernel<8><<<grid, kSortThreads, 0, st>>>(ja, pa, pass);
            TT_LAUNCH_OK("sort_scatter_kernel");
        }
    }
    return TT_OK;
}

int tt_debug_sparse_plan(const tt_sparse_job* jobs, int njobs, void* ws, size_t ws_bytes, int32_t* vec, int32_t* gvec) {
    JobArr ja;
    PlanArr pa;
    int max_tiles = 0, max_pass = 0;
    TT_REQUIRE(vec != nullptr && gvec != nullptr, "tt_debug_sparse_plan: null pointer");
    int rc = make_plans(jobs, njobs, ws, ws_bytes, &ja, &pa, &max_tiles, &max_pass, "tt_debug_sparse_plan");
    if (rc) return rc;
    for (int j = 0; j < njobs; ++j) { vec[j] = pa.p[j].vec; gvec[j] = pa.p[j].gvec; }
    return TT_OK;
}

int tt_sparse_adagrad(const tt_sparse_job* jobs, int njobs, float lr, float eps, void* ws, size_t ws_bytes, void* stream) {
    JobArr ja;
    PlanArr pa;
    int max_tiles = 0, max_pass = 0;
    int rc = make_plans(jobs, njobs, ws, ws_bytes, &ja, &pa, &max_tiles, &max_pass, "tt_sparse_adagrad");
    if (rc) return rc;
    if (max_tiles == 0) return TT_OK;
    return launch_apply(ja, pa, njobs, kModeAdagrad, lr, eps, 0.f, 0.f, as_stream(stream));
}

int tt_sparse_adam(const tt_sparse_job* jobs, int njobs, float lr_t, float beta1, float beta2, float eps, void* ws, size_t ws_bytes,
                   void* stream) {
    JobArr ja;
    PlanArr pa;
    int max_tiles = 0, max_pass = 0;
    int rc = make_plans(jobs, njobs, ws, ws_bytes, &ja, &pa, &max_tiles, &max_pass, "tt_sparse_adam");
    if (rc) return rc;
    for (int j = 0; j < njobs; ++j) TT_REQUIRE(jobs[j].slot1 != nullptr, "tt_sparse_adam: job %d has no second moment", j);
    cudaStream_t st = as_stream(stream);
    int64_t max_elems = 0;
    for (int j = 0; j < njobs; ++j) {
        int64_t t = (int64_t)jobs[j].rows * jobs[j].e;
        if (t > max_elems) max_elems = t;
    }
    int64_t sg = ceil_div(max_elems, 256 * 4);
    int64_t cap = (int64_t)sm_count() * 8;
    if (sg > cap) sg = cap;
    if (sg < 1) sg = 1;
    dim3 sweep((unsigned)sg, (unsigned)njobs);
    adam_sweep_kernel<<<sweep, 256, 0, st>>>(ja, 0, beta1, beta2, lr_t, eps);
    TT_LAUNCH_OK("adam_sweep_kernel<decay>");
    if (max_tiles > 0) {
        rc = launch_apply(ja, pa, njobs, kModeAdamMoments, 0.f, eps, 1.0f - beta1, 1.0f - beta2, st);
        if (rc) return rc;
    }
    adam_sweep_kernel<<<sweep, 256, 0, st>>>(ja, 1, beta1, beta2, lr_t, eps);
    TT_LAUNCH_OK("adam_sweep_kernel<update>");
    return TT_OK;
}

}  // extern "C"
