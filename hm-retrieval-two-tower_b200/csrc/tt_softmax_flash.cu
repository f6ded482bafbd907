// tt_softmax_flash.cu -- host side of the two-pass in-batch softmax (tt_tc_flash.cuh) for E in {64, 128}.
//
// Replaces (reference file:line): two_tower_model.py:90-92 (Q.C^T), logq_correction.py:66-71, two_tower_model.py:119-122 +
// runner.py:78-83 (eye labels, CE from logits, SUM) and the autodiff of those (two_tower_model.py:110-124).
//
//   step     : amax -> per-tensor power-of-two scales | fp16 operand copies + scaled column terms | pass 1 (forward + dQ) |
//              combine 1 (lse, loss, dQ, lse column term) | pass 2 (dC) | combine 2          -- 6 launches + one 64-byte memset
//   forward  : the same up to combine 1 without the dQ output
//   backward : (lse given) pass 2 on one or both sides + combine 2
// Precision contract: operands are rounded to fp16 AFTER scaling each tensor so that its largest magnitude lies in [2^14, 2^15):
// 11 significant bits like TF32, no overflow for any finite input, entries below 2^-29 of the tensor's maximum lose relative
// precision only (absolute logit error < 2^-40 of the largest product).  Products are exact, accumulation is fp32.
#include <cuda_fp16.h>

#include <atomic>
#include <cstdlib>
#include <utility>

#include "tt_tc_flash.cuh"

namespace tt {
namespace tc {

int make_tmap_2d_f16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);   // tt_softmax_tc.cu

static std::atomic<int> g_fl_lbo{0}, g_fl_sbo{0};                  // debug knobs (process-wide, read once per call)
static std::atomic<unsigned long long*> g_fl_trace{nullptr};

// ---- programmatic dependent launch -----------------------------------------------------------------------------------------
// The six kernels of a step run back to back on one stream.  Each is launched with the programmatic-stream-serialization
// attribute and calls pdl_trigger() first thing (the next kernel may be launched: its blocks become resident as soon as ours leave
// room, and run their set-up) and pdl_wait() before it touches anything its predecessors wrote (returns once they have completed
// and their memory is visible).  That hides each kernel's launch latency and prologue behind the previous kernel's tail.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
static bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("TT_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- scales ---------------------------------------------------------------------------------------------------
// scal[0] sq  [1] sc  [2],[3] kmul = log2e / (sq*sc)  [4] 1/sq  [5] 1/sc ; [11] ticket of the loss reduction   (64 floats), followed
// by the per-block maxima of fl_amax_kernel: part[0][kAmaxBlocksMax] for Q, part[1][kAmaxBlocksMax] for C (float bit patterns)
constexpr int kScalFloats = 64;
constexpr int kAmaxBlocksMax = 512;

__device__ __forceinline__ float pow2_scale_for(uint32_t amax_bits) {
    const int e = (int)(amax_bits >> 23);                       // biased exponent of the largest magnitude
    if (e == 0 || e == 255) return 1.f;                         // all zero / denormal, or inf / NaN (the result is inf / NaN either way)
    int se = 268 - e;                                           // 2^(14 - floor(log2 amax)), biased
    se = se > 167 ? 167 : se;                                   // cap at 2^40
    return __uint_as_float((uint32_t)se << 23);
}

struct AmaxArgs { const float* X[2]; int ld[2]; int n[2]; int E; };
// Every block leaves the largest magnitudes it saw (as float bit patterns: non-negative floats order like integers) in part[t][block];
// fl_convert_kernel reduces the <= 512 values per tensor itself.  No atomics, no ticket, nothing to zero beforehand (a memset node
// and a last-block stage per call in the first version); block 0 also zeroes the loss reduction's ticket.
__global__ void __launch_bounds__(256) fl_amax_kernel(const AmaxArgs a, float* __restrict__ scal) {
    pdl_trigger();
    pdl_wait();   // (the inputs may come from a kernel launched the same way)
    uint32_t* part = reinterpret_cast<uint32_t*>(scal) + kScalFloats;
    __shared__ uint32_t s_max[2][8];
    uint32_t mx[2] = {0u, 0u};
    const int e4 = a.E >> 2;
    for (int t = 0; t < 2; ++t) {
        const int64_t cnt = (int64_t)a.n[t] * e4;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < cnt; i += (int64_t)gridDim.x * blockDim.x) {
            const int r = (int)(i / e4), c4 = (int)(i % e4);
            const float4 v = *reinterpret_cast<const float4*>(a.X[t] + (int64_t)r * a.ld[t] + 4 * c4);
            const uint32_t m0 = max(__float_as_uint(fabsf(v.x)), __float_as_uint(fabsf(v.y)));
            const uint32_t m1 = max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)));
            mx[t] = max(mx[t], max(m0, m1));
        }
    }
    for (int t = 0; t < 2; ++t) {
        for (int o = 16; o > 0; o >>= 1) mx[t] = max(mx[t], __shfl_xor_sync(0xffffffffu, mx[t], o));
        if ((threadIdx.x & 31) == 0) s_max[t][threadIdx.x >> 5] = mx[t];
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int t = threadIdx.x;
        uint32_t m = 0;
        for (int w = 0; w < 8; ++w) m = max(m, s_max[t][w]);
        part[t * kAmaxBlocksMax + blockIdx.x] = m;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) reinterpret_cast<uint32_t*>(scal)[11] = 0u;
}

// the two operand scales from the per-block maxima (every thread of the block returns the same pair)
__device__ __forceinline__ void fl_block_scales(const float* __restrict__ scal, int nblk, float& sq, float& sc) {
    __shared__ uint32_t s_red[2][8];
    const uint32_t* part = reinterpret_cast<const uint32_t*>(scal) + kScalFloats;
    uint32_t m0 = 0u, m1 = 0u;
    for (int i = threadIdx.x; i < nblk; i += 256) { m0 = max(m0, part[i]); m1 = max(m1, part[kAmaxBlocksMax + i]); }
    for (int o = 16; o > 0; o >>= 1) { m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, o)); m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, o)); }
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = m0; s_red[1][threadIdx.x >> 5] = m1; }
    __syncthreads();
    m0 = 0u; m1 = 0u;
    for (int w = 0; w < 8; ++w) { m0 = max(m0, s_red[0][w]); m1 = max(m1, s_red[1][w]); }
    sq = pow2_scale_for(m0);
    sc = pow2_scale_for(m1);
}

// ---- operand copies, column terms, the positives' logits ------------------------------------------------------------------
struct CvtItem {
    const float* X; int ld, n; __half* Xh; int which;            // Xh[n][E] = fp16(X * scal[which])
    const float* colv; float* c2; int n_col, n_pad; float cmul;  // c2[j] = colv[j] * cmul (0 when colv is null), zero padded to n_pad
    int xblocks, cblocks;
};
// z_ii = Q[i] . C[i + off] - bias[i + off] from the fp32 operands (the positive is left out of the tensor-core products):
// zd2[i] = z_ii * log2e for the pass-1 combine; with lse given (backward entry points) also pm1[i] = exp(z_ii - lse_i) - 1
struct DiagItem { const float* Q; int ldq; const float* C; int ldc; const float* bias; const float* lse; int Bq, off; float* zd2; float* pm1; int blocks; };
struct CvtArgs { CvtItem s[2]; int n; int E; DiagItem dg; int amax_blocks; };
__global__ void __launch_bounds__(256) fl_convert_kernel(const CvtArgs a, float* __restrict__ scal) {
    pdl_trigger();
    pdl_wait();
    int b = blockIdx.x;
    const int e8 = a.E >> 3;
    if (b == 0) {            // block 0 publishes the scales for the kernels that follow (they start after this grid has completed)
        float sq, sc;
        fl_block_scales(scal, a.amax_blocks, sq, sc);
        if (threadIdx.x == 0) {
            scal[0] = sq; scal[1] = sc;
            const float inv = (1.f / sq) * (1.f / sc);           // exact: powers of two, |exponent| <= 80
            scal[2] = kLog2e * inv; scal[3] = kLog2e * inv;
            scal[4] = 1.f / sq; scal[5] = 1.f / sc;
        }
    }
    if (b < a.dg.blocks) {   // e8 consecutive lanes per row
        const int64_t idx = (int64_t)b * 256 + threadIdx.x;
        const int r = (int)(idx / e8), c8 = (int)(idx % e8);
        float dot = 0.f;
        if (r < a.dg.Bq) {
            const float* q = a.dg.Q + (int64_t)r * a.dg.ldq + 8 * c8;
            const float* c = a.dg.C + (int64_t)(r + a.dg.off) * a.dg.ldc + 8 * c8;
            const float4 q0 = *reinterpret_cast<const float4*>(q), q1 = *reinterpret_cast<const float4*>(q + 4);
            const float4 c0 = *reinterpret_cast<const float4*>(c), c1 = *reinterpret_cast<const float4*>(c + 4);
            dot = fmaf(q0.x, c0.x, fmaf(q0.y, c0.y, fmaf(q0.z, c0.z, q0.w * c0.w))) + fmaf(q1.x, c1.x, fmaf(q1.y, c1.y, fmaf(q1.z, c1.z, q1.w * c1.w)));
        }
        for (int o = e8 >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (r < a.dg.Bq && c8 == 0) {
            const float z = dot - (a.dg.bias ? a.dg.bias[r + a.dg.off] : 0.f);
            a.dg.zd2[r] = z * kLog2e;
            if (a.dg.lse) a.dg.pm1[r] = expf(z - a.dg.lse[r]) - 1.f;
        }
        return;
    }
    b -= a.dg.blocks;
    for (int i = 0; i < a.n; ++i) {
        const CvtItem& it = a.s[i];
        if (b < it.xblocks) {
            float sq, sc;
            fl_block_scales(scal, a.amax_blocks, sq, sc);        // (uniform per block; scal[0..1] may not have been written yet)
            const int64_t idx = (int64_t)b * 256 + threadIdx.x;
            if (idx < (int64_t)it.n * e8) {
                const int r = (int)(idx / e8), c8 = (int)(idx % e8);
                const float s = it.which == 0 ? sq : sc;
                const float4 v0 = *reinterpret_cast<const float4*>(it.X + (int64_t)r * it.ld + 8 * c8);
                const float4 v1 = *reinterpret_cast<const float4*>(it.X + (int64_t)r * it.ld + 8 * c8 + 4);
                uint4 o;
                o.x = pack_f16x2(v0.x * s, v0.y * s); o.y = pack_f16x2(v0.z * s, v0.w * s);
                o.z = pack_f16x2(v1.x * s, v1.y * s); o.w = pack_f16x2(v1.z * s, v1.w * s);
                *reinterpret_cast<uint4*>(it.Xh + (int64_t)r * a.E + 8 * c8) = o;
            }
            return;
        }
        b -= it.xblocks;
        if (b < it.cblocks) {
            const int j = b * 256 + threadIdx.x;
            if (j < it.n_pad) it.c2[j] = (j < it.n_col && it.colv) ? it.colv[j] * it.cmul : 0.f;
            return;
        }
        b -= it.cblocks;
    }
}
static int cvt_item(CvtItem& it, const float* X, int ld, int n, int E, __half* Xh, int which, const float* colv, float* c2, int n_col, int n_pad, float cmul) {
    it.X = X; it.ld = ld; it.n = n; it.Xh = Xh; it.which = which; it.colv = colv; it.c2 = c2; it.n_col = n_col; it.n_pad = n_pad; it.cmul = cmul;
    it.xblocks = Xh ? (int)ceil_div((int64_t)n * (E / 8), 256) : 0;
    it.cblocks = c2 ? (int)ceil_div(n_pad, 256) : 0;
    return it.xblocks + it.cblocks;
}

// ---- plan ---------------------------------------------------------------------------------------------------------------
static inline int fl_bn(int E) { return E <= 64 ? 128 : 64; }
static inline int fl_split(int E) { return fl_bn(E) / 64; }   // partials per (panel pair, CTA slot): one per 64-column half of the tile
struct FlPlan { int grid, units, slots[2], unit0[2], m_pairs[2], n_tiles[2]; };
static FlPlan fl_plan(int n_pass, const int* nR, const int* nT, int E) {
    FlPlan pl{};
    const int bn = fl_bn(E);
    for (int i = 0; i < n_pass; ++i) {
        pl.m_pairs[i] = (int)ceil_div(nR[i], 256);
        pl.n_tiles[i] = (int)ceil_div(nT[i], bn);
        pl.unit0[i] = pl.units;
        pl.units += pl.m_pairs[i] * pl.n_tiles[i];
    }
    const int sms = sm_count();
    pl.grid = pl.units < sms ? (pl.units > 0 ? pl.units : 1) : sms;
    for (int i = 0; i < n_pass; ++i) {
        int mx = 1;
        for (int pr = 0; pr < pl.m_pairs[i]; ++pr) {
            const int first = pl.unit0[i] + pr * pl.n_tiles[i];
            const int s = sk_owner(first + pl.n_tiles[i] - 1, pl.units, pl.grid) - sk_owner(first, pl.units, pl.grid) + 1;
            mx = s > mx ? s : mx;
        }
        pl.slots[i] = mx;
    }
    return pl;
}

template <int MODE, int E>
static int launch_flash(const FlMaps& maps, const FlParams& p, int grid, cudaStream_t st, const char* name) {
    using Cfg = FlCfg<E>;
    { static SmemAttr smem_attr; TT_CUDA_OK(smem_attr.ensure(flash_kernel<MODE, E>, Cfg::kSmemBytes)); }
    TT_CUDA_OK(launch_pdl(flash_kernel<MODE, E>, dim3((unsigned)grid), dim3(Cfg::kThreads), (size_t)Cfg::kSmemBytes, st, maps, p));
    TT_LAUNCH_OK(name);
    return TT_OK;
}
template <int MODE>
static int launch_flash_e(int E, const FlMaps& maps, const FlParams& p, int grid, cudaStream_t st) {
    if (E == 64) return launch_flash<MODE, 64>(maps, p, grid, st, MODE == kP1 ? "flash_kernel<p1,64>" : "flash_kernel<p2,64>");
    return launch_flash<MODE, 128>(maps, p, grid, st, MODE == kP1 ? "flash_kernel<p1,128>" : "flash_kernel<p2,128>");
}

// Partial G blocks are stored [part][E/4][rows] as float4, so that the flush of a segment (thread = row) and the merges below
// (thread = row, blockIdx.y = float4 column) touch consecutive rows with consecutive lanes.

// ---- combine 1: per-row merge of the pass-1 partials in slot order -> lse, row loss, loss, dQ, lse column term, p_ii - 1 -------------
// The positive was left out of the sums: with L_off = sum_s l_s 2^(m_s - M), pd = 2^(zd - M) (zd from the fp32 operands),
//   L = L_off + pd,  lse = M + log2 L,  p_ii - 1 = -L_off / L  (no cancellation when the softmax is sharp),
//   row loss = ln(1 + L_off / pd),      dQ_i = (sum_s G_s 2^(m_s - M)) / (L * scale_C) + (p_ii - 1) C[i + d]
struct Comb1Args {
    const float* pm; const float* pl; const float4* pg; const float* zd2;
    int nR, rows_pad, n_tiles, units, grid, unit0, E, d, ksplit;
    const float* C; int ldc;
    const float* scal;
    float* lse; float* rowloss; float* c2_lse; int c2_pad; float* pm1;
    float* dQ; int lddq;
    double* block_sums; unsigned int* counter; float* loss;
};
__global__ void __launch_bounds__(256) fl_combine1_kernel(const Comb1Args a) {
    __shared__ double s_sum[256];
    __shared__ bool s_last;
    const int e4 = a.E >> 2;
    const int r = blockIdx.x * 256 + threadIdx.x, c4 = blockIdx.y;
    pdl_trigger();
    pdl_wait();
    double mine = 0.0;
    if (r < a.nR) {
        const int first = a.unit0 + (r >> 8) * a.n_tiles;
        const int parts = (sk_owner(first + a.n_tiles - 1, a.units, a.grid) - sk_owner(first, a.units, a.grid) + 1) * a.ksplit;
        const float zd = a.zd2[r];
        float M = zd;                                            // the positive takes part in the maximum
        // The partials are read kCB at a time with every load of a batch issued before the first use (a plain loop over a
        // run-time part count serialises one L2 round trip per part); the merge order -- part 0, 1, 2, ... -- is unchanged.
        constexpr int kCB = 6;
        const float* pm_r = a.pm + r;
        const float* pl_r = a.pl + r;
        const float4* pg_r = a.pg + (int64_t)c4 * a.rows_pad + r;
        for (int s0 = 0; s0 < parts; s0 += kCB) {
            float v[kCB];
#pragma unroll
            for (int j = 0; j < kCB; ++j) v[j] = (s0 + j < parts) ? pm_r[(int64_t)(s0 + j) * a.rows_pad] : -CUDART_INF_F;
#pragma unroll
            for (int j = 0; j < kCB; ++j) M = fmaxf(M, v[j]);
        }
        float Loff = 0.f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s0 = 0; s0 < parts; s0 += kCB) {
            float vm[kCB], vl[kCB];
            float4 vg[kCB];
#pragma unroll
            for (int j = 0; j < kCB; ++j) {
                const bool on = s0 + j < parts;
                vm[j] = on ? pm_r[(int64_t)(s0 + j) * a.rows_pad] : -CUDART_INF_F;
                vl[j] = on ? pl_r[(int64_t)(s0 + j) * a.rows_pad] : 0.f;
                vg[j] = (on && a.dQ) ? pg_r[(int64_t)(s0 + j) * e4 * a.rows_pad] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < kCB; ++j) {
                if (s0 + j < parts) {
                    const float w = exp2f(vm[j] - M);
                    Loff = fmaf(vl[j], w, Loff);
                    acc.x = fmaf(vg[j].x, w, acc.x); acc.y = fmaf(vg[j].y, w, acc.y); acc.z = fmaf(vg[j].z, w, acc.z); acc.w = fmaf(vg[j].w, w, acc.w);
                }
            }
        }
        const float pd = exp2f(zd - M);
        const float L = Loff + pd;
        const float lse2 = M + log2f(L);
        const float pm1 = -Loff / L;                             // p_ii - 1
        if (a.dQ) {
            const float inv = a.scal[5] / L;                     // 1 / (L * scale_C)
            const float4 cv = *reinterpret_cast<const float4*>(a.C + (int64_t)(r + a.d) * a.ldc + 4 * c4);
            *reinterpret_cast<float4*>(a.dQ + (int64_t)r * a.lddq + 4 * c4) =
                make_float4(fmaf(acc.x, inv, pm1 * cv.x), fmaf(acc.y, inv, pm1 * cv.y), fmaf(acc.z, inv, pm1 * cv.z), fmaf(acc.w, inv, pm1 * cv.w));
        }
        if (c4 == 0) {
            a.lse[r] = lse2 * kLn2;
            const float rl = (pd >= Loff) ? log1pf(Loff / pd) : (lse2 - zd) * kLn2;
            a.rowloss[r] = rl;
            mine = (double)rl;
            if (a.c2_lse) a.c2_lse[r] = lse2;
            if (a.pm1) a.pm1[r] = pm1;
        }
    } else if (a.c2_lse && c4 == 0 && r < a.c2_pad) {
        a.c2_lse[r] = 0.f;
    }
    if (c4 != 0) return;   // (uniform per block) the loss is summed by the blocks of float4 column 0
    s_sum[threadIdx.x] = mine;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.block_sums[blockIdx.x] = s_sum[0];
        __threadfence();
        s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {   // the last block adds the block sums: thread t takes blocks t, t + 256, ... in order, then a fixed tree (deterministic)
        __threadfence();
        double t = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += 256) t += reinterpret_cast<volatile double*>(a.block_sums)[b];
        s_sum[threadIdx.x] = t;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            a.loss[0] = (float)s_sum[0];
            *a.counter = 0u;
        }
    }
}

// ---- combine 2: G[r][:] = 2^-kOff2 / scale_T * sum over the pair's CTA slots, in slot order, + (p - 1) T[r + d] ----------------------
// The positive (row r <-> T row r + d) is left out of the tensor-core product; its term is added here in fp32.  p - 1 is indexed by
// the QUERY of the pair: the pass-1 combine wrote it (step), or the convert kernel formed it from the given lse (backward).
struct Comb2Side {
    const float4* part; float* G; int ldg, nR, rows_pad, n_tiles, unit0, scal_idx, rblocks;
    const float* T; int ldt, nT, d;
    const float* pm1_q; int q_is_row;   // query index of row r: r (dQ side) or r + d (dC side)
};
struct Comb2Args { Comb2Side s[2]; int n, E, units, grid, ksplit; const float* scal; };
__global__ void __launch_bounds__(256) fl_combine2_kernel(const Comb2Args a) {
    pdl_trigger();
    pdl_wait();
    int rb = blockIdx.x;
    const int e4 = a.E >> 2, c4 = blockIdx.y;
    for (int k = 0; k < a.n; ++k) {
        const Comb2Side& sd = a.s[k];
        if (rb < sd.rblocks) {
            const int r = rb * 256 + threadIdx.x;
            if (r >= sd.nR) return;
            const int first = sd.unit0 + (r >> 8) * sd.n_tiles;
            const int parts = (sk_owner(first + sd.n_tiles - 1, a.units, a.grid) - sk_owner(first, a.units, a.grid) + 1) * a.ksplit;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            constexpr int kCB = 8;                                // loads batched as in combine 1; the sum order (part 0, 1, ...) is unchanged
            const float4* pg_r = sd.part + (int64_t)c4 * sd.rows_pad + r;
            for (int z0 = 0; z0 < parts; z0 += kCB) {
                float4 v[kCB];
#pragma unroll
                for (int j = 0; j < kCB; ++j) v[j] = (z0 + j < parts) ? pg_r[(int64_t)(z0 + j) * e4 * sd.rows_pad] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < kCB; ++j) {
                    if (z0 + j < parts) {
                        acc.x = __fadd_rn(acc.x, v[j].x); acc.y = __fadd_rn(acc.y, v[j].y); acc.z = __fadd_rn(acc.z, v[j].z); acc.w = __fadd_rn(acc.w, v[j].w);
                    }
                }
            }
            const int tr = r + sd.d;
            float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
            float pm1 = 0.f;
            if (tr >= 0 && tr < sd.nT) {
                tv = *reinterpret_cast<const float4*>(sd.T + (int64_t)tr * sd.ldt + 4 * c4);
                pm1 = sd.pm1_q[sd.q_is_row ? r : tr];
            }
            const float f = a.scal[sd.scal_idx] * 6.103515625e-05f;   // 2^-14 / scale_T (exact)
            *reinterpret_cast<float4*>(sd.G + (int64_t)r * sd.ldg + 4 * c4) =
                make_float4(fmaf(acc.x, f, pm1 * tv.x), fmaf(acc.y, f, pm1 * tv.y), fmaf(acc.z, f, pm1 * tv.z), fmaf(acc.w, f, pm1 * tv.w));
            return;
        }
        rb -= sd.rblocks;
    }
}

// ---- workspace ------------------------------------------------------------------------------------------------------------
struct FlWs {
    float* scal; __half* Qh; __half* Ch; float* c2_bias; float* c2_lse; float* pm1; float* zd2; double* block_sums; unsigned int* counter;
    float* rowloss; float* p1_m; float* p1_l; float* p1_g; float* p2_g[2];
    size_t bytes;
};
// one layout for every entry point (step / forward / backward with one or both sides)
static FlWs fl_carve(void* ws, int Bq, int Bc, int E) {
    FlWs w{};
    Carver cv(ws ? ws : reinterpret_cast<void*>(256));           // ws == null: size query only
    const int bn = fl_bn(E);
    const int pad_c = (int)(ceil_div(Bc, bn) * bn) + 256, pad_q = (int)(ceil_div(Bq, bn) * bn) + 256;
    w.rowloss = cv.take<float>((size_t)(Bq > Bc ? Bq : Bc));      // first: the SIMT path keeps its row losses at the workspace base too
    w.scal = cv.take<float>(kScalFloats + 2 * kAmaxBlocksMax);
    w.Qh = cv.take<__half>((size_t)Bq * E);
    w.Ch = cv.take<__half>((size_t)Bc * E);
    w.c2_bias = cv.take<float>((size_t)pad_c);
    w.c2_lse = cv.take<float>((size_t)pad_q);
    w.pm1 = cv.take<float>((size_t)pad_q);
    w.zd2 = cv.take<float>((size_t)pad_q);
    w.block_sums = cv.take<double>((size_t)ceil_div(Bq, 256) + 64);
    w.counter = reinterpret_cast<unsigned int*>(w.scal) + 11;     // zeroed by block 0 of fl_amax_kernel
    // pass 1: R = Q, T = C
    int nR1[1] = {Bq}, nT1[1] = {Bc};
    FlPlan p1 = fl_plan(1, nR1, nT1, E);
    const size_t rp1 = (size_t)p1.m_pairs[0] * 256, ks = (size_t)fl_split(E);
    w.p1_m = cv.take<float>(ks * p1.slots[0] * rp1);
    w.p1_l = cv.take<float>(ks * p1.slots[0] * rp1);
    w.p1_g = cv.take<float>(ks * p1.slots[0] * rp1 * E);
    // pass 2: up to two sides in one launch (dQ side: R = Q; dC side: R = C); sized for the larger of {both, either alone}
    int nRb[2] = {Bq, Bc}, nTb[2] = {Bc, Bq};
    FlPlan pb = fl_plan(2, nRb, nTb, E);
    int nRq[1] = {Bq}, nTq[1] = {Bc}, nRc[1] = {Bc}, nTc[1] = {Bq};
    FlPlan pq = fl_plan(1, nRq, nTq, E), pc = fl_plan(1, nRc, nTc, E);
    const int sq = pb.slots[0] > pq.slots[0] ? pb.slots[0] : pq.slots[0], sc = pb.slots[1] > pc.slots[0] ? pb.slots[1] : pc.slots[0];
    w.p2_g[0] = cv.take<float>(ks * sq * pb.m_pairs[0] * 256 * E);
    w.p2_g[1] = cv.take<float>(ks * sc * pb.m_pairs[1] * 256 * E);
    w.bytes = align_up(cv.off, 256) + 256;
    return w;
}
size_t softmax_flash_workspace(int Bq, int Bc, int E) { return fl_carve(nullptr, Bq, Bc, E).bytes; }

// amax -> scales, fp16 copies, column terms
static int fl_prepare(const FlWs& w, const float* Q, int ldq, const float* C, int ldc, const float* bias, const float* lse, int Bq, int Bc, int E,
                      int off, cudaStream_t st) {
    const int bn = fl_bn(E);
    AmaxArgs aa{};
    aa.X[0] = Q; aa.ld[0] = ldq; aa.n[0] = Bq; aa.X[1] = C; aa.ld[1] = ldc; aa.n[1] = Bc; aa.E = E;
    int64_t work = ((int64_t)Bq + Bc) * (E / 4);
    int blocks = (int)ceil_div(work, 256 * 4);
    blocks = blocks > 2 * sm_count() ? 2 * sm_count() : (blocks < 1 ? 1 : blocks);
    blocks = blocks > kAmaxBlocksMax ? kAmaxBlocksMax : blocks;
    TT_CUDA_OK(launch_pdl(fl_amax_kernel, dim3((unsigned)blocks), dim3(256), 0, st, aa, w.scal));
    TT_LAUNCH_OK("fl_amax_kernel");
    CvtArgs ca{};
    ca.E = E;
    ca.amax_blocks = blocks;
    ca.dg = DiagItem{Q, ldq, C, ldc, bias, lse, Bq, off, w.zd2, w.pm1, (int)ceil_div((int64_t)Bq * (E / 8), 256)};
    int cb = ca.dg.blocks;
    cb += cvt_item(ca.s[ca.n++], Q, ldq, Bq, E, w.Qh, 0, lse, lse ? w.c2_lse : nullptr, Bq, (int)(ceil_div(Bq, bn) * bn), kLog2e);
    cb += cvt_item(ca.s[ca.n++], C, ldc, Bc, E, w.Ch, 1, bias, w.c2_bias, Bc, (int)(ceil_div(Bc, bn) * bn), kLog2e);
    TT_CUDA_OK(launch_pdl(fl_convert_kernel, dim3((unsigned)cb), dim3(256), 0, st, ca, w.scal));
    TT_LAUNCH_OK("fl_convert_kernel");
    return TT_OK;
}

static int fl_pass1(const FlWs& w, const float* C, int ldc, int Bq, int Bc, int E, int off, float* lse, float* loss, float* dQ, int lddq,
                    bool want_c2_lse, cudaStream_t st) {
    const int bn = fl_bn(E);
    int nR[1] = {Bq}, nT[1] = {Bc};
    FlPlan pl = fl_plan(1, nR, nT, E);
    FlMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc = make_tmap_2d_f16(&maps.r[0], w.Qh, Bq, E, E, 128);
    if (rc) return rc;
    rc = make_tmap_2d_f16(&maps.t[0], w.Ch, Bc, E, E, bn);
    if (rc) return rc;
    FlParams p{};
    p.n_pass = 1; p.units = pl.units; p.kmul = w.scal + 2; p.mn_lbo = g_fl_lbo.load(); p.mn_sbo = g_fl_sbo.load(); p.trace = g_fl_trace.load();
    FlPass& ps = p.pass[0];
    ps.nR = Bq; ps.nT = Bc; ps.m_pairs = pl.m_pairs[0]; ps.n_tiles = pl.n_tiles[0]; ps.d = off; ps.unit0 = 0; ps.rowv = nullptr;
    ps.colv2 = w.c2_bias; ps.out_g = w.p1_g; ps.out_m = w.p1_m; ps.out_l = w.p1_l;
    rc = launch_flash_e<kP1>(E, maps, p, pl.grid, st);
    if (rc) return rc;
    Comb1Args ca{};
    ca.pm = w.p1_m; ca.pl = w.p1_l; ca.pg = reinterpret_cast<const float4*>(w.p1_g); ca.zd2 = w.zd2;
    ca.nR = Bq; ca.rows_pad = pl.m_pairs[0] * 256; ca.n_tiles = pl.n_tiles[0]; ca.units = pl.units; ca.grid = pl.grid; ca.unit0 = 0; ca.E = E; ca.d = off; ca.ksplit = fl_split(E);
    ca.C = C; ca.ldc = ldc; ca.scal = w.scal; ca.lse = lse; ca.rowloss = w.rowloss;
    ca.c2_lse = want_c2_lse ? w.c2_lse : nullptr; ca.c2_pad = (int)(ceil_div(Bq, bn) * bn); ca.pm1 = want_c2_lse ? w.pm1 : nullptr;
    ca.dQ = dQ; ca.lddq = lddq; ca.block_sums = w.block_sums; ca.counter = w.counter; ca.loss = loss;
    const int rows_c = ca.c2_pad > Bq ? ca.c2_pad : Bq;
    TT_CUDA_OK(launch_pdl(fl_combine1_kernel, dim3((unsigned)ceil_div(rows_c, 256), (unsigned)(E / 4)), dim3(256), 0, st, ca));
    TT_LAUNCH_OK("fl_combine1_kernel");
    return TT_OK;
}

struct FlSide { int r_is_q; float* G; int ldg; };   // r_is_q: resident operand Q (gradient dQ) or C (gradient dC)
static int fl_pass2(const FlWs& w, const float* Q, int ldq, const float* C, int ldc, const float* bias, const float* lse, int Bq, int Bc, int E, int off,
                    const FlSide* sides, int n, cudaStream_t st) {
    const int bn = fl_bn(E);
    int nR[2], nT[2];
    for (int i = 0; i < n; ++i) { nR[i] = sides[i].r_is_q ? Bq : Bc; nT[i] = sides[i].r_is_q ? Bc : Bq; }
    FlPlan pl = fl_plan(n, nR, nT, E);
    if (pl.units == 0) return TT_OK;
    FlMaps maps;
    memset(&maps, 0, sizeof(maps));
    FlParams p{};
    p.n_pass = n; p.units = pl.units; p.kmul = w.scal + 2; p.mn_lbo = g_fl_lbo.load(); p.mn_sbo = g_fl_sbo.load(); p.trace = g_fl_trace.load();
    Comb2Args ca{};
    ca.n = n; ca.E = E; ca.units = pl.units; ca.grid = pl.grid; ca.ksplit = fl_split(E); ca.scal = w.scal;
    int rblocks = 0;
    for (int i = 0; i < n; ++i) {
        const bool rq = sides[i].r_is_q != 0;
        int rc = make_tmap_2d_f16(&maps.r[i], rq ? w.Qh : w.Ch, nR[i], E, E, 128);
        if (rc) return rc;
        rc = make_tmap_2d_f16(&maps.t[i], rq ? w.Ch : w.Qh, nT[i], E, E, bn);
        if (rc) return rc;
        FlPass& ps = p.pass[i];
        ps.nR = nR[i]; ps.nT = nT[i]; ps.m_pairs = pl.m_pairs[i]; ps.n_tiles = pl.n_tiles[i]; ps.d = rq ? off : -off; ps.unit0 = pl.unit0[i];
        ps.rowv = rq ? lse : bias; ps.colv2 = rq ? w.c2_bias : w.c2_lse;
        ps.out_g = w.p2_g[rq ? 0 : 1]; ps.out_m = nullptr; ps.out_l = nullptr;
        Comb2Side& cs = ca.s[i];
        cs.part = reinterpret_cast<const float4*>(ps.out_g); cs.G = sides[i].G; cs.ldg = sides[i].ldg; cs.nR = nR[i]; cs.rows_pad = pl.m_pairs[i] * 256;
        cs.n_tiles = pl.n_tiles[i]; cs.rblocks = (int)ceil_div(nR[i], 256);
        cs.unit0 = pl.unit0[i]; cs.scal_idx = rq ? 5 : 4;   // the streamed operand's scale: C for the dQ side, Q for the dC side
        cs.T = rq ? C : Q; cs.ldt = rq ? ldc : ldq; cs.nT = nT[i]; cs.d = ps.d;
        cs.pm1_q = w.pm1; cs.q_is_row = rq ? 1 : 0;
        rblocks += cs.rblocks;
    }
    int rc = launch_flash_e<kP2>(E, maps, p, pl.grid, st);
    if (rc) return rc;
    TT_CUDA_OK(launch_pdl(fl_combine2_kernel, dim3((unsigned)rblocks, (unsigned)(E / 4)), dim3(256), 0, st, ca));
    TT_LAUNCH_OK("fl_combine2_kernel");
    return TT_OK;
}

}  // namespace tc

using namespace tc;

bool softmax_flash_supported(int E) { return E == 64 || E == 128; }
size_t softmax_flash_workspace_bytes(int Bq, int Bc, int E) { return tc::softmax_flash_workspace(Bq, Bc, E); }

int softmax_step_flash(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse, float* loss,
                       float* dQ, int lddq, float* dC, int lddc, float* ws, cudaStream_t st) {
    FlWs w = fl_carve(ws, Bq, Bc, E);
    int rc = fl_prepare(w, Q, ldq, C, ldc, bias, nullptr, Bq, Bc, E, off, st);
    if (rc) return rc;
    rc = fl_pass1(w, C, ldc, Bq, Bc, E, off, lse, loss, dQ, lddq, true, st);
    if (rc) return rc;
    FlSide side{0, dC, lddc};
    return fl_pass2(w, Q, ldq, C, ldc, bias, lse, Bq, Bc, E, off, &side, 1, st);
}

int softmax_fwd_flash(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse, float* loss,
                      float* ws, cudaStream_t st) {
    FlWs w = fl_carve(ws, Bq, Bc, E);
    int rc = fl_prepare(w, Q, ldq, C, ldc, bias, nullptr, Bq, Bc, E, off, st);
    if (rc) return rc;
    return fl_pass1(w, C, ldc, Bq, Bc, E, off, lse, loss, nullptr, 0, false, st);
}

// which = 0 dQ only, 1 dC only, 2 both (G0 = dQ, G1 = dC)
int softmax_bwd_flash(const float* Q, int ldq, const float* C, int ldc, const float* bias, const float* lse, int Bq, int Bc, int E, int off, int which,
                      float* G0, int ldg0, float* G1, int ldg1, float* ws, cudaStream_t st) {
    FlWs w = fl_carve(ws, Bq, Bc, E);
    int rc = fl_prepare(w, Q, ldq, C, ldc, bias, lse, Bq, Bc, E, off, st);
    if (rc) return rc;
    FlSide sides[2];
    int n = 0;
    if (which == 0 || which == 2) sides[n++] = FlSide{1, G0, ldg0};
    if (which == 1) sides[n++] = FlSide{0, G0, ldg0};
    if (which == 2) sides[n++] = FlSide{0, G1, ldg1};
    return fl_pass2(w, Q, ldq, C, ldc, bias, lse, Bq, Bc, E, off, sides, n, st);
}

void debug_flash(void* trace, int mn_lbo, int mn_sbo) {
    tc::g_fl_trace.store(reinterpret_cast<unsigned long long*>(trace));
    tc::g_fl_lbo.store(mn_lbo);
    tc::g_fl_sbo.store(mn_sbo);
}

}  // namespace tt
