// tt_softmax_tc.cu -- in-batch sampled softmax on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces (reference file:line): two_tower_model.py:90-92 (Q.C^T), logq_correction.py:66-71,
// two_tower_model.py:119-122 + runner.py:78-83 (eye labels, CE from logits, SUM) and the autodiff of those.
// Operands are expected TF32-rounded (tt_round_tf32 / the Y_tf32 output of tt_dense_fwd); products are
// exact in fp32 and accumulate in fp32 in TMEM, so logits agree with the fp32 reference to ~2^-11 relative.
//
//   forward : scale/pad of the column term, ONE persistent streamk_kernel<kFwd> (tt_tc_streamk.cuh: every SM walks an
//             equal share of the (row panel, column tile) units), a combine kernel (per-row partials -> lse, row loss)
//             and the fixed-order loss sum
//   backward: one prep kernel (transposed copies of both operands for the second MMA's K-major B tiles; scaled column
//             terms), ONE persistent streamk_kernel<kBwd> covering the dQ and the dC pass, one fixed-order reduction of
//             the per-(panel, CTA) partial blocks (deterministic, no atomics)
#include <cuda_fp16.h>

#include <atomic>

#include "tt_tc_rowpanel.cuh"
#include "tt_tc_streamk.cuh"

namespace tt {

int sum_rows_launch(const float* v, int n, float* out, cudaStream_t st);  // tt_softmax_simt.cu
// tt_softmax_flash.cu: the two-pass kernels that serve E in {64, 128}; the stream-K kernels below keep E = 32
bool softmax_flash_supported(int E);
size_t softmax_flash_workspace_bytes(int Bq, int Bc, int E);
int softmax_step_flash(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse, float* loss,
                       float* dQ, int lddq, float* dC, int lddc, float* ws, cudaStream_t st);
int softmax_fwd_flash(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse, float* loss,
                      float* ws, cudaStream_t st);
int softmax_bwd_flash(const float* Q, int ldq, const float* C, int ldc, const float* bias, const float* lse, int Bq, int Bc, int E, int off, int which,
                      float* G0, int ldg0, float* G1, int ldg1, float* ws, cudaStream_t st);

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

bool tensor_maps_available() { return encode_fn() != nullptr; }

int make_tmap_2d(CUtensorMap* out, const float* base, int64_t rows, int cols, int ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return TT_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d ld=%d box_rows=%d", (int)r, (long long)rows, cols, ld, box_rows);
        return TT_ERR_CUDA;
    }
    return TT_OK;
}

// 2-D fp16 row-major matrix; box = 64 columns (128 B, one swizzle span) x box_rows rows; SWIZZLE_128B; OOB reads as zero
int make_tmap_2d_f16(CUtensorMap* out, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)"); return TT_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(fp16) failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows);
        return TT_ERR_CUDA;
    }
    return TT_OK;
}

// ---- small helper kernels ---------------------------------------------------------------------------------
// out[i] = colv[i] * log2(e) for i < n (0 when colv is null), zero padding up to n_pad
__global__ void scale_pad_kernel(const float* __restrict__ colv, int n, float* __restrict__ out, int n_pad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) out[i] = (i < n && colv) ? colv[i] * kLog2e : 0.f;
}

static int launch_scale_pad(const float* colv, int n, float* out, int n_pad, cudaStream_t st) {
    scale_pad_kernel<<<(unsigned)ceil_div(n_pad, 256), 256, 0, st>>>(colv, n, out, n_pad);
    TT_LAUNCH_OK("scale_pad_kernel");
    return TT_OK;
}

template <int MODE, int E, int BN>
static int launch_rowpanel(const CUtensorMap& tmR, const CUtensorMap& tmT, const CUtensorMap& tmTt, const RowPanelParams& p, int m_tiles,
                           int splits, cudaStream_t st, const char* name) {
    using Cfg = RowPanelCfg<MODE, E, BN>;
    { static SmemAttr smem_attr; TT_CUDA_OK(smem_attr.ensure(rowpanel_kernel<MODE, E, BN>, Cfg::kSmemBytes)); }
    dim3 grid((unsigned)m_tiles, (unsigned)splits);
    rowpanel_kernel<MODE, E, BN><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tmR, tmT, tmTt, p);
    TT_LAUNCH_OK(name);
    return TT_OK;
}

constexpr int kFwdBN = 256;     // forward tile width: the per-unit costs of the producer and MMA lanes amortise over 2x the columns
constexpr int kLogitsBN = 128;  // rowpanel_kernel<kLogits>
constexpr int kFwdHalves = 4;    // SkCfg<kFwd, E, 128>::kHalves
constexpr int kMaxSplits = 16;   // column splits per launch; the forward keeps 2 partials per split (one per warp half)

// debug knobs (tt_debug_tc): timeline buffer and a cap on the column splits
static std::atomic<unsigned long long*> g_trace{nullptr};   // (debug knobs: process-wide, read once per call)
static std::atomic<int> g_max_splits{kMaxSplits};

struct Plan {
    int m_tiles, n_tiles, splits, tps;
};
static Plan plan_for(int nR, int nT, int bn) {
    Plan pl;
    pl.m_tiles = (int)ceil_div(nR, 128);
    pl.n_tiles = (int)ceil_div(nT, bn);
    choose_splits(pl.m_tiles, pl.n_tiles, 2, g_max_splits.load(), &pl.splits, &pl.tps);
    return pl;
}

static inline size_t tt_ld(int n) { return align_up((size_t)n, 4); }




// ---- stream-K host side ---------------------------------------------------------------------------------
struct SkPlan {
    int grid;       // CTAs
    int units;
    int slots[2];   // partial slots per panel (max number of CTAs that share a panel), per pass
};
static int sk_slots(int unit0, int m_tiles, int n_tiles, int units, int G) {
    int mx = 1;
    for (int pnl = 0; pnl < m_tiles; ++pnl) {
        const int first = unit0 + pnl * n_tiles;
        const int s = sk_owner(first + n_tiles - 1, units, G) - sk_owner(first, units, G) + 1;
        mx = s > mx ? s : mx;
    }
    return mx;
}
static SkPlan sk_plan(int n_pass, const int* m_tiles, const int* n_tiles) {
    SkPlan pl{};
    pl.units = 0;
    int unit0[2] = {0, 0};
    for (int i = 0; i < n_pass; ++i) { unit0[i] = pl.units; pl.units += m_tiles[i] * n_tiles[i]; }
    const int sms = sm_count();
    pl.grid = pl.units < sms ? (pl.units > 0 ? pl.units : 1) : sms;
    for (int i = 0; i < n_pass; ++i) pl.slots[i] = sk_slots(unit0[i], m_tiles[i], n_tiles[i], pl.units, pl.grid);
    return pl;
}

template <int MODE, int E, int BN, bool H>
static int launch_streamk(const SkMaps& maps, const SkParams& p, int grid, cudaStream_t st, const char* name) {
    using Cfg = SkCfg<MODE, E, BN, H>;
    { static SmemAttr smem_attr; TT_CUDA_OK(smem_attr.ensure(streamk_kernel<MODE, E, BN, H>, Cfg::kSmemBytes)); }
    streamk_kernel<MODE, E, BN, H><<<(unsigned)grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(maps, p);
    TT_LAUNCH_OK(name);
    return TT_OK;
}

// per-row merge of the forward partials, slot order then warp-half order.  Also: the scaled, zero-padded lse column term
// of the dC pass (c2_lse, optional), and the loss = sum of the row losses -- every block leaves its partial sum (double),
// the last block to finish adds them in block order (a fixed order: deterministic) and re-arms the counter.
__global__ void __launch_bounds__(256) fwd_combine_sk_kernel(const float* __restrict__ m2, const float* __restrict__ l, const float* __restrict__ zd,
                                                             int nR, int n_tiles, int units, int G, int halves, float* __restrict__ lse,
                                                             float* __restrict__ rowloss, float* __restrict__ c2_lse, int c2_pad,
                                                             double* __restrict__ block_sums, unsigned int* __restrict__ counter,
                                                             float* __restrict__ loss) {
    __shared__ double s_sum[256];
    __shared__ bool s_last;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    double mine = 0.0;
    if (r < nR) {
        const int first = (r >> 7) * n_tiles;
        const int parts = (sk_owner(first + n_tiles - 1, units, G) - sk_owner(first, units, G) + 1) * halves;
        // the partials of a row are independent loads: fetch them eight at a time (the merge order is unchanged)
        float M = -CUDART_INF_F;
        for (int s0 = 0; s0 < parts; s0 += 8) {
            float mv[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) mv[t] = (s0 + t < parts) ? m2[(int64_t)(s0 + t) * nR + r] : -CUDART_INF_F;
#pragma unroll
            for (int t = 0; t < 8; ++t) M = fmaxf(M, mv[t]);
        }
        float L = 0.f;
        for (int s0 = 0; s0 < parts; s0 += 8) {
            float mv[8], lv[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const bool on = s0 + t < parts;
                mv[t] = on ? m2[(int64_t)(s0 + t) * nR + r] : -CUDART_INF_F;
                lv[t] = on ? l[(int64_t)(s0 + t) * nR + r] : 0.f;
            }
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (mv[t] > -CUDART_INF_F) L += lv[t] * exp2f(mv[t] - M);
        }
        const float v = (M + log2f(L)) * 0.6931471805599453f;
        lse[r] = v;
        const float rl = v - zd[r];
        rowloss[r] = rl;
        mine = (double)rl;
        if (c2_lse) c2_lse[r] = v * kLog2e;
    } else if (c2_lse && r < c2_pad) {
        c2_lse[r] = 0.f;
    }
    s_sum[threadIdx.x] = mine;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        block_sums[blockIdx.x] = s_sum[0];
        __threadfence();
        s_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (unsigned b = 0; b < gridDim.x; ++b) t += reinterpret_cast<volatile double*>(block_sums)[b];
        loss[0] = (float)t;
        *counter = 0u;
    }
}

// operand copies and column terms prepared once for a forward + backward pair (tt_inbatch_softmax_step)
struct SkPrepared {
    const __half* Qh; const __half* Ch;    // fp16 row-major copies (E >= 64), else null
    const __half* Qt; const __half* Ct;    // fp16 transposed copies (E x ld)
    int ldqt, ldct;
    const float* c2_bias;                  // bias * log2e, zero padded
    float* c2_lse;                         // lse * log2e, zero padded to c2_lse_pad (written by the forward's combine kernel)
    int c2_lse_pad;
    double* block_sums; unsigned int* counter;   // loss reduction scratch (counter zeroed by the prep kernel)
};

struct SkSide {   // one backward pass
    const float* R; int ldr;
    const float* T; int ldt;
    const float* rowv; const float* colv;
    int nR, nT, d;
    float* G; int ldg;
};
struct PrepItem {   // one matrix X (n x E, leading dimension ld) and/or one per-row vector of it
    const float* X; int ld, n;
    __half* Xh;               // fp16 row-major copy (n x E, dense), or null
    __half* Xt; int ldtt;     // fp16 transposed copy (E x ldtt), or null
    const float* colv; float* c2; int n_pad;   // c2[j] = colv[j] * log2 e, zero padded to n_pad (c2 null: skip)
    int tblocks_x, tblocks;   // 32x32 tiles: per row of tiles, total (0 when no copy is wanted)
    int cblocks;              // scale/pad blocks of 256
};
struct PrepArgs { PrepItem s[3]; int n; int E; unsigned int* zero_me; };
// one launch: fp16 copies (row-major: operands of the first MMA; transposed: B tiles of the second MMA) and scaled,
// padded column terms, for every pass
__global__ void __launch_bounds__(256) sk_prep_kernel(const PrepArgs a) {
    __shared__ float tile[32][33];
    int b = blockIdx.x;
    if (b == 0 && threadIdx.x == 0 && a.zero_me) *a.zero_me = 0u;
    for (int i = 0; i < a.n; ++i) {
        const PrepItem& sd = a.s[i];
        if (b < sd.tblocks) {
            const int r0 = (b % sd.tblocks_x) * 32, c0 = (b / sd.tblocks_x) * 32;
            const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
            for (int k = ty; k < 32; k += 8) {
                const int r = r0 + k, c = c0 + tx;
                const float v = (r < sd.n && c < a.E) ? sd.X[(int64_t)r * sd.ld + c] : 0.f;
                tile[k][tx] = v;
                if (sd.Xh && r < sd.n && c < a.E) sd.Xh[(int64_t)r * a.E + c] = __float2half_rn(v);
            }
            if (sd.Xt) {
                __syncthreads();
                for (int k = ty; k < 32; k += 8) {
                    const int c = c0 + k, r = r0 + tx;
                    if (c < a.E && r < sd.n) sd.Xt[(int64_t)c * sd.ldtt + r] = __float2half_rn(tile[tx][k]);
                }
            }
            return;
        }
        b -= sd.tblocks;
        if (b < sd.cblocks) {
            const int j = b * 256 + threadIdx.x;
            if (j < sd.n_pad) sd.c2[j] = (j < sd.n && sd.colv) ? sd.colv[j] * kLog2e : 0.f;
            return;
        }
        b -= sd.cblocks;
    }
}
static int prep_item(PrepItem& it, const float* X, int ld, int n, int E, __half* Xh, __half* Xt, int ldtt, const float* colv, float* c2, int n_pad) {
    it.X = X; it.ld = ld; it.n = n; it.Xh = Xh; it.Xt = Xt; it.ldtt = ldtt; it.colv = colv; it.c2 = c2; it.n_pad = n_pad;
    it.tblocks_x = (int)ceil_div(n, 32);
    it.tblocks = (Xh || Xt) ? it.tblocks_x * (int)ceil_div(E, 32) : 0;
    it.cblocks = c2 ? (int)ceil_div(n_pad, 256) : 0;
    return it.tblocks + it.cblocks;
}

struct RedSide { const float* part; float* G; int ldg, nR, n_tiles, unit0; };
struct RedArgs { RedSide s[2]; int n, E, units, grid; };
// G[r][:] = sum over the panel's CTA slots, in slot order (one thread per float4)
__global__ void __launch_bounds__(256) bwd_reduce_sk_kernel(const RedArgs a) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int e4 = a.E >> 2;
    for (int k = 0; k < a.n; ++k) {
        const RedSide& sd = a.s[k];
        const int64_t cnt = (int64_t)sd.nR * e4;
        if (i < cnt) {
            const int r = (int)(i / e4), c4 = (int)(i % e4);
            const int first = sd.unit0 + (r >> 7) * sd.n_tiles;
            const int slots = sk_owner(first + sd.n_tiles - 1, a.units, a.grid) - sk_owner(first, a.units, a.grid) + 1;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int z = 0; z < slots; ++z) {
                const float4 v = *reinterpret_cast<const float4*>(sd.part + ((int64_t)z * sd.nR + r) * a.E + 4 * c4);
                acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
            }
            float* dst = sd.G + (int64_t)r * sd.ldg + 4 * c4;
            dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w;
            return;
        }
        i -= cnt;
    }
}

static inline int sk_bwd_bn(int E) { return E <= 64 ? 128 : 64; }   // shared-memory budget (SkCfg)
static inline size_t sk_ldtt(int n) { return align_up((size_t)n, 8); }   // fp16 row stride: 16-byte multiples

static size_t bwd_sk_floats(int n_sides, const int* nR, const int* nT, int E) {
    const int bn = sk_bwd_bn(E);
    int mt[2], nt[2];
    for (int i = 0; i < n_sides; ++i) { mt[i] = (int)ceil_div(nR[i], 128); nt[i] = (int)ceil_div(nT[i], bn); }
    SkPlan pl = sk_plan(n_sides, mt, nt);
    size_t f = 0;
    for (int i = 0; i < n_sides; ++i) {
        f += align_up((size_t)pl.slots[i] * nR[i] * E, 64);          // partial blocks
        f += align_up((size_t)E * sk_ldtt(nT[i]) / 2, 64);           // T^T (fp16)
        f += align_up((size_t)nT[i] * E / 2, 64);                    // T (fp16 row-major)
        f += align_up((size_t)nR[i] * E / 2, 64);                    // R (fp16 row-major; shared with the other pass when both run)
        f += align_up((size_t)nt[i] * bn, 64);                       // scaled column term
    }
    return f + 512;
}

size_t softmax_step_floats(int Bq, int Bc, int E);
static size_t sk_fwd_bytes(int Bq, int Bc, int E) {
    size_t rows = (size_t)(Bq > Bc ? Bq : Bc);
    // forward: m2, l partials (slots x halves) + zdiag + rowloss + scaled column term + fp16 copies + loss scratch
    int mt = (int)ceil_div(Bq, 128), nt = (int)ceil_div(Bc, kFwdBN);
    SkPlan pf = sk_plan(1, &mt, &nt);
    size_t seg = align_up(rows * sizeof(float), 256);
    return (2 * kFwdHalves * (size_t)pf.slots[0] + 2) * seg + align_up(((size_t)nt * kFwdBN + 64) * sizeof(float), 256) +
           align_up((size_t)Bq * E * 2, 256) + align_up((size_t)Bc * E * 2, 256) + align_up(((size_t)ceil_div(rows, 256) + 64) * 8, 256) + 4096;
}
static size_t sk_bwd_bytes(int Bq, int Bc, int E) {
    // backward: both passes at once, or one pass alone (tt_inbatch_softmax_bwd_one)
    int nR[2] = {Bq, Bc}, nT[2] = {Bc, Bq};
    size_t both = bwd_sk_floats(2, nR, nT, E) * sizeof(float);
    int nRq[1] = {Bq}, nTq[1] = {Bc}, nRc[1] = {Bc}, nTc[1] = {Bq};
    size_t one_q = bwd_sk_floats(1, nRq, nTq, E) * sizeof(float), one_c = bwd_sk_floats(1, nRc, nTc, E) * sizeof(float);
    size_t bwd = both > one_q ? both : one_q;
    return align_up(bwd > one_c ? bwd : one_c, 256);
}
size_t softmax_tc_workspace(int Bq, int Bc, int E) {
    size_t rows = (size_t)(Bq > Bc ? Bq : Bc);
    size_t fwd = sk_fwd_bytes(Bq, Bc, E), bwd = sk_bwd_bytes(Bq, Bc, E);
    size_t step = fwd + bwd + softmax_step_floats(Bq, Bc, E) * sizeof(float);      // tt_inbatch_softmax_step keeps all three regions live
    (void)rows;
    return step + 1024;
}

}  // namespace tc

using namespace tc;

bool softmax_tc_supported(int ldq, int ldc, int E, const void* Q, const void* C) {
    static int dev_ok = -1;
    if (dev_ok < 0) {
        int dev = 0, major = 0;
        dev_ok = (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess &&
                  major == 10 && tensor_maps_available())
                     ? 1
                     : 0;
    }
    if (!dev_ok) return false;
    if (!(E == 32 || E == 64 || E == 128)) return false;
    if (ldq % 4 || ldc % 4) return false;
    if ((reinterpret_cast<uintptr_t>(Q) & 15) || (reinterpret_cast<uintptr_t>(C) & 15)) return false;
    return true;
}

static int softmax_fwd_sk(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse, float* loss,
                          float* ws, cudaStream_t st, const SkPrepared* pre) {
    int mt = (int)ceil_div(Bq, 128), nt = (int)ceil_div(Bc, kFwdBN);
    SkPlan pl = sk_plan(1, &mt, &nt);
    SkMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc = make_tmap_2d(&maps.r[0], Q, Bq, E, ldq, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&maps.t[0], C, Bc, E, ldc, kFwdBN);
    if (rc) return rc;
    size_t rows = (size_t)(Bq > Bc ? Bq : Bc);
    size_t seg = align_up(rows * sizeof(float), 256) / sizeof(float);
    float* rowloss = ws;                                   // (the SIMT path keeps its row losses at the workspace base too)
    float* zd = rowloss + seg;
    float* m2 = zd + seg;
    float* l = m2 + kFwdHalves * (size_t)pl.slots[0] * seg;
    float* c2 = l + kFwdHalves * (size_t)pl.slots[0] * seg;
    const bool h16 = E >= 64;                              // fp16 operand tiles (see SkCfg)
    __half* Qh = reinterpret_cast<__half*>(c2 + align_up((size_t)nt * kFwdBN + 64, 64));
    __half* Ch = Qh + align_up((size_t)Bq * E, 128);
    double* block_sums = reinterpret_cast<double*>(Ch + align_up((size_t)Bc * E, 128));
    unsigned int* counter = reinterpret_cast<unsigned int*>(block_sums + align_up((size_t)ceil_div(Bq, 256), 32));
    if (pre) {
        Qh = const_cast<__half*>(pre->Qh); Ch = const_cast<__half*>(pre->Ch); c2 = const_cast<float*>(pre->c2_bias);
        block_sums = pre->block_sums; counter = pre->counter;
    } else {
        PrepArgs pa{};
        pa.E = E; pa.zero_me = counter;
        int blocks = 0;
        blocks += prep_item(pa.s[pa.n++], C, ldc, Bc, E, h16 ? Ch : nullptr, nullptr, 0, bias, c2, nt * kFwdBN);
        if (h16) blocks += prep_item(pa.s[pa.n++], Q, ldq, Bq, E, Qh, nullptr, 0, nullptr, nullptr, 0);
        sk_prep_kernel<<<(unsigned)blocks, 256, 0, st>>>(pa);
        TT_LAUNCH_OK("sk_prep_kernel");
    }
    if (h16) {
        rc = make_tmap_2d_f16(&maps.r[0], Qh, Bq, E, E, 128);
        if (rc) return rc;
        rc = make_tmap_2d_f16(&maps.t[0], Ch, Bc, E, E, kFwdBN);
        if (rc) return rc;
    }
    SkParams p{};
    p.n_pass = 1; p.units = pl.units; p.trace = g_trace.load();
    SkPass& ps = p.pass[0];
    ps.nR = Bq; ps.nT = Bc; ps.m_tiles = mt; ps.n_tiles = nt; ps.d = off; ps.unit0 = 0; ps.rowv = nullptr; ps.colv2 = c2;
    ps.out0 = m2; ps.out1 = l; ps.out2 = zd;
    switch (E) {
        case 32: rc = launch_streamk<kFwd, 32, kFwdBN, false>(maps, p, pl.grid, st, "streamk_kernel<fwd,32>"); break;
        case 64: rc = launch_streamk<kFwd, 64, kFwdBN, true>(maps, p, pl.grid, st, "streamk_kernel<fwd,64>"); break;
        default: rc = launch_streamk<kFwd, 128, kFwdBN, true>(maps, p, pl.grid, st, "streamk_kernel<fwd,128>"); break;
    }
    if (rc) return rc;
    const int c2_pad = pre ? pre->c2_lse_pad : 0;
    const int rows_c = c2_pad > Bq ? c2_pad : Bq;
    fwd_combine_sk_kernel<<<(unsigned)ceil_div(rows_c, 256), 256, 0, st>>>(m2, l, zd, Bq, nt, pl.units, pl.grid, kFwdHalves, lse, rowloss,
                                                                            pre ? pre->c2_lse : nullptr, c2_pad, block_sums, counter, loss);
    TT_LAUNCH_OK("fwd_combine_sk_kernel");
    return TT_OK;
}

int softmax_fwd_tc(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse, float* loss,
                   float* ws, cudaStream_t st) {
    if (softmax_flash_supported(E)) return softmax_fwd_flash(Q, ldq, C, ldc, bias, Bq, Bc, E, off, lse, loss, ws, st);
    return softmax_fwd_sk(Q, ldq, C, ldc, bias, Bq, Bc, E, off, lse, loss, ws, st, nullptr);
}

// one launch for `n` backward passes (dQ and dC, or one of them)
int softmax_bwd_sk(const SkSide* sides, int n, int E, float* ws, cudaStream_t st, const SkPrepared* pre = nullptr) {
    const int bn = sk_bwd_bn(E);
    int mt[2] = {0, 0}, nt[2] = {0, 0};
    for (int i = 0; i < n; ++i) { mt[i] = (int)ceil_div(sides[i].nR, 128); nt[i] = (int)ceil_div(sides[i].nT, bn); }
    SkPlan pl = sk_plan(n, mt, nt);
    SkMaps maps;
    memset(&maps, 0, sizeof(maps));
    SkParams p{};
    p.n_pass = n; p.units = pl.units; p.trace = g_trace.load();
    PrepArgs pa{};
    pa.E = E;
    RedArgs ra{};
    ra.n = n; ra.E = E; ra.units = pl.units; ra.grid = pl.grid;
    const bool h16 = E >= 64;                              // fp16 operand tiles for the first MMA too (see SkCfg)
    float* cur = ws;
    int unit0 = 0, prep_blocks = 0;
    int64_t red_items = 0;
    __half* Th[2] = {nullptr, nullptr};                    // fp16 row-major copy of each side's streamed operand
    float* part[2]; __half* Tt[2]; float* c2[2]; int ldtt[2];
    for (int i = 0; i < n; ++i) {
        const SkSide& sd = sides[i];
        part[i] = cur; cur += align_up((size_t)pl.slots[i] * sd.nR * E, 64);
        Tt[i] = reinterpret_cast<__half*>(cur); ldtt[i] = (int)sk_ldtt(sd.nT); cur += align_up((size_t)E * ldtt[i] / 2, 64);
        c2[i] = cur; cur += align_up((size_t)nt[i] * bn, 64);
        if (h16) { Th[i] = reinterpret_cast<__half*>(cur); cur += align_up((size_t)sd.nT * E / 2, 64); }
        if (pre) {   // side 0 streams C (column term: bias), side 1 streams Q (column term: lse)
            Th[i] = const_cast<__half*>(i == 0 ? pre->Ch : pre->Qh);
            Tt[i] = const_cast<__half*>(i == 0 ? pre->Ct : pre->Qt);
            ldtt[i] = i == 0 ? pre->ldct : pre->ldqt;
            c2[i] = i == 0 ? const_cast<float*>(pre->c2_bias) : pre->c2_lse;
        } else {
            prep_blocks += prep_item(pa.s[pa.n++], sd.T, sd.ldt, sd.nT, E, Th[i], Tt[i], ldtt[i], sd.colv, c2[i], nt[i] * bn);
        }
    }
    // the resident operand of a pass is the streamed operand of the other one; a pass running alone converts its own
    __half* Rh[2] = {nullptr, nullptr};
    if (h16) {
        if (pre) { Rh[0] = const_cast<__half*>(pre->Qh); Rh[1] = const_cast<__half*>(pre->Ch); }
        else if (n == 2) { Rh[0] = Th[1]; Rh[1] = Th[0]; }
        else {
            Rh[0] = reinterpret_cast<__half*>(cur); cur += align_up((size_t)sides[0].nR * E / 2, 64);
            prep_blocks += prep_item(pa.s[pa.n++], sides[0].R, sides[0].ldr, sides[0].nR, E, Rh[0], nullptr, 0, nullptr, nullptr, 0);
        }
    }
    for (int i = 0; i < n; ++i) {
        const SkSide& sd = sides[i];
        int rc;
        if (h16) {
            rc = make_tmap_2d_f16(&maps.r[i], Rh[i], sd.nR, E, E, 128);
            if (rc) return rc;
            rc = make_tmap_2d_f16(&maps.t[i], Th[i], sd.nT, E, E, bn);
        } else {
            rc = make_tmap_2d(&maps.r[i], sd.R, sd.nR, E, sd.ldr, 128);
            if (rc) return rc;
            rc = make_tmap_2d(&maps.t[i], sd.T, sd.nT, E, sd.ldt, bn);
        }
        if (rc) return rc;
        rc = make_tmap_2d_f16(&maps.tt[i], Tt[i], E, sd.nT, ldtt[i], E);
        if (rc) return rc;
        SkPass& ps = p.pass[i];
        ps.nR = sd.nR; ps.nT = sd.nT; ps.m_tiles = mt[i]; ps.n_tiles = nt[i]; ps.d = sd.d; ps.unit0 = unit0; ps.rowv = sd.rowv; ps.colv2 = c2[i];
        ps.out0 = part[i]; ps.out1 = nullptr; ps.out2 = nullptr;
        RedSide& rs = ra.s[i];
        rs.part = part[i]; rs.G = sd.G; rs.ldg = sd.ldg; rs.nR = sd.nR; rs.n_tiles = nt[i]; rs.unit0 = unit0;
        red_items += (int64_t)sd.nR * (E / 4);
        unit0 += mt[i] * nt[i];
    }
    if (pl.units == 0) return TT_OK;
    if (!pre) {
        sk_prep_kernel<<<(unsigned)prep_blocks, 256, 0, st>>>(pa);
        TT_LAUNCH_OK("sk_prep_kernel");
    }
    int rc;
    switch (E) {
        case 32: rc = launch_streamk<kBwd, 32, 128, false>(maps, p, pl.grid, st, "streamk_kernel<bwd,32>"); break;
        case 64: rc = launch_streamk<kBwd, 64, 128, true>(maps, p, pl.grid, st, "streamk_kernel<bwd,64>"); break;
        default: rc = launch_streamk<kBwd, 128, 64, true>(maps, p, pl.grid, st, "streamk_kernel<bwd,128>"); break;
    }
    if (rc) return rc;
    bwd_reduce_sk_kernel<<<(unsigned)ceil_div(red_items, 256), 256, 0, st>>>(ra);
    TT_LAUNCH_OK("bwd_reduce_sk_kernel");
    return TT_OK;
}

// forward + backward of one training step: ONE prep launch (fp16 row-major and transposed copies of Q and C, scaled bias),
// forward, combine (lse, loss, scaled lse column term), backward (dQ and dC), reduction -- 5 launches.
size_t tc::softmax_step_floats(int Bq, int Bc, int E) {
    const int bn = sk_bwd_bn(E);
    size_t f = 0;
    f += 2 * align_up((size_t)Bq * E / 2, 64) + 2 * align_up((size_t)Bc * E / 2, 64);                   // Qh, Ch
    f += align_up((size_t)E * sk_ldtt(Bq) / 2, 64) + align_up((size_t)E * sk_ldtt(Bc) / 2, 64);         // Qt, Ct
    f += align_up((size_t)Bc + 512, 64) + align_up((size_t)Bq + 512, 64);                               // c2_bias, c2_lse (padded)
    f += align_up(2 * ((size_t)ceil_div(Bq + 512, 256) + 64), 64) + 64;                                 // block sums (double), counter
    (void)bn;
    return f;
}

int softmax_step_tc(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse, float* loss,
                    float* dQ, int lddq, float* dC, int lddc, float* ws, cudaStream_t st) {
    if (softmax_flash_supported(E)) return softmax_step_flash(Q, ldq, C, ldc, bias, Bq, Bc, E, off, lse, loss, dQ, lddq, dC, lddc, ws, st);
    const size_t ws_floats_fwd = sk_fwd_bytes(Bq, Bc, E) / sizeof(float), ws_floats_bwd = sk_bwd_bytes(Bq, Bc, E) / sizeof(float);
    const bool h16 = E >= 64;
    const int bn = sk_bwd_bn(E);
    // workspace: [forward region | backward region | prepared operands]
    float* ws_fwd = ws;
    float* ws_bwd = ws + ws_floats_fwd;
    float* cur = ws_bwd + ws_floats_bwd;
    SkPrepared pre{};
    __half* Qh = reinterpret_cast<__half*>(cur); cur += align_up((size_t)Bq * E / 2, 64);
    __half* Ch = reinterpret_cast<__half*>(cur); cur += align_up((size_t)Bc * E / 2, 64);
    pre.ldqt = (int)sk_ldtt(Bq); pre.ldct = (int)sk_ldtt(Bc);
    __half* Qt = reinterpret_cast<__half*>(cur); cur += align_up((size_t)E * pre.ldqt / 2, 64);
    __half* Ct = reinterpret_cast<__half*>(cur); cur += align_up((size_t)E * pre.ldct / 2, 64);
    float* c2_bias = cur; cur += align_up((size_t)Bc + 512, 64);
    float* c2_lse = cur; cur += align_up((size_t)Bq + 512, 64);
    double* block_sums = reinterpret_cast<double*>(cur); cur += align_up(2 * ((size_t)ceil_div(Bq + 512, 256) + 64), 64);
    unsigned int* counter = reinterpret_cast<unsigned int*>(cur);
    const int pad_bias = (int)(ceil_div(Bc, 256) * 256);                 // covers the forward (BN 256) and the dQ pass (BN <= 128)
    const int pad_lse = (int)(ceil_div(Bq, bn) * bn);
    PrepArgs pa{};
    pa.E = E; pa.zero_me = counter;
    int blocks = 0;
    blocks += prep_item(pa.s[pa.n++], C, ldc, Bc, E, h16 ? Ch : nullptr, Ct, pre.ldct, bias, c2_bias, pad_bias);
    blocks += prep_item(pa.s[pa.n++], Q, ldq, Bq, E, h16 ? Qh : nullptr, Qt, pre.ldqt, nullptr, nullptr, 0);
    sk_prep_kernel<<<(unsigned)blocks, 256, 0, st>>>(pa);
    TT_LAUNCH_OK("sk_prep_kernel");
    pre.Qh = h16 ? Qh : nullptr; pre.Ch = h16 ? Ch : nullptr; pre.Qt = Qt; pre.Ct = Ct;
    pre.c2_bias = c2_bias; pre.c2_lse = c2_lse; pre.c2_lse_pad = pad_lse; pre.block_sums = block_sums; pre.counter = counter;
    int rc = softmax_fwd_sk(Q, ldq, C, ldc, bias, Bq, Bc, E, off, lse, loss, ws_fwd, st, &pre);
    if (rc) return rc;
    SkSide sides[2] = {SkSide{Q, ldq, C, ldc, lse, bias, Bq, Bc, off, dQ, lddq}, SkSide{C, ldc, Q, ldq, bias, lse, Bc, Bq, -off, dC, lddc}};
    return softmax_bwd_sk(sides, 2, E, ws_bwd, st, &pre);
}

// (Q, C) backward: which = 0 dQ only, 1 dC only, 2 both (G0 = dQ, G1 = dC)
int softmax_bwd_tc(const float* Q, int ldq, const float* C, int ldc, const float* bias, const float* lse, int Bq, int Bc, int E, int off, int which,
                   float* G0, int ldg0, float* G1, int ldg1, float* ws, cudaStream_t st) {
    if (softmax_flash_supported(E)) return softmax_bwd_flash(Q, ldq, C, ldc, bias, lse, Bq, Bc, E, off, which, G0, ldg0, G1, ldg1, ws, st);
    SkSide sides[2];
    int n = 0;
    if (which == 0 || which == 2) sides[n++] = SkSide{Q, ldq, C, ldc, lse, bias, Bq, Bc, off, G0, ldg0};
    if (which == 1) sides[n++] = SkSide{C, ldc, Q, ldq, bias, lse, Bc, Bq, -off, G0, ldg0};
    if (which == 2) sides[n++] = SkSide{C, ldc, Q, ldq, bias, lse, Bc, Bq, -off, G1, ldg1};
    return softmax_bwd_sk(sides, n, E, ws, st);
}

int logits_tc(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, float* Z, int ldz, cudaStream_t st) {
    Plan pl = plan_for(Bq, Bc, kLogitsBN);
    CUtensorMap tmQ, tmC;
    int rc = make_tmap_2d(&tmQ, Q, Bq, E, ldq, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmC, C, Bc, E, ldc, kLogitsBN);
    if (rc) return rc;
    // (test / API-convenience path: the column bias is read from global memory by the epilogue, no scratch needed)
    RowPanelParams p{};
    p.nR = Bq; p.nT = Bc; p.n_tiles = pl.n_tiles; p.tiles_per_split = pl.tps; p.colv2 = nullptr; p.rowv2 = bias; p.out0 = Z; p.ld_out = ldz;
    p.d = -(1 << 30);
    switch (E) {
        case 32: return launch_rowpanel<kLogits, 32, kLogitsBN>(tmQ, tmC, tmC, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<logits,32>");
        case 64: return launch_rowpanel<kLogits, 64, kLogitsBN>(tmQ, tmC, tmC, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<logits,64>");
        default: return launch_rowpanel<kLogits, 128, kLogitsBN>(tmQ, tmC, tmC, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<logits,128>");
    }
}

size_t softmax_tc_workspace_bytes(int Bq, int Bc, int E) {
    return softmax_flash_supported(E) ? softmax_flash_workspace_bytes(Bq, Bc, E) : tc::softmax_tc_workspace(Bq, Bc, E);
}

void debug_tc(void* trace, int max_splits) {
    tc::g_trace.store(reinterpret_cast<unsigned long long*>(trace));
    tc::g_max_splits.store((max_splits >= 1 && max_splits <= tc::kMaxSplits) ? max_splits : tc::kMaxSplits);
}

}  // namespace tt
