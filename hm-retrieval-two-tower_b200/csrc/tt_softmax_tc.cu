// tt_softmax_tc.cu -- in-batch sampled softmax on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces (reference file:line): two_tower_model.py:90-92 (Q.C^T), logq_correction.py:66-71,
// two_tower_model.py:119-122 + runner.py:78-83 (eye labels, CE from logits, SUM) and the autodiff of those.
// Operands are expected TF32-rounded (tt_round_tf32 / the Y_tf32 output of tt_dense_fwd); products are
// exact in fp32 and accumulate in fp32 in TMEM, so logits agree with the fp32 reference to ~2^-11 relative.
//
//   forward : one rowpanel_kernel<kFwd> over (row panels x column splits) + a tiny combine kernel
//   backward: rowpanel_kernel<kBwd> twice (rows = queries -> dQ; rows = candidates -> dC), each followed by
//             a fixed-order reduction of the per-split partial sums (deterministic, no atomics).  The
//             second MMA of the backward kernel reads a K-major tile of T^T, so each pass first writes a
//             transposed copy of its streamed operand into the workspace.
#include "tt_tc_rowpanel.cuh"

namespace tt {

int sum_rows_launch(const float* v, int n, float* out, cudaStream_t st);  // tt_softmax_simt.cu

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

// ---- small helper kernels ---------------------------------------------------------------------------------
__global__ void fwd_combine_kernel(const float* __restrict__ m2, const float* __restrict__ l, const float* __restrict__ zd, int splits, int nR,
                                   float* __restrict__ lse, float* __restrict__ rowloss) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nR) return;
    float M = -CUDART_INF_F;
    for (int s = 0; s < splits; ++s) M = fmaxf(M, m2[(int64_t)s * nR + r]);
    float L = 0.f;
    for (int s = 0; s < splits; ++s) {
        float ms = m2[(int64_t)s * nR + r];
        if (ms > -CUDART_INF_F) L += l[(int64_t)s * nR + r] * exp2f(ms - M);
    }
    float v = (M + log2f(L)) * 0.6931471805599453f;
    lse[r] = v;
    rowloss[r] = v - zd[r];
}

__global__ void bwd_reduce_kernel(const float* __restrict__ part, int splits, int64_t n, int E, float* __restrict__ G, int ldg) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s = __fadd_rn(s, part[(int64_t)z * n + i]);
    int64_t r = i / E;
    int c = (int)(i % E);
    G[r * ldg + c] = s;
}

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

// out (cols x ldo) = in (rows x cols)^T, 32x32 shared-memory tiles, coalesced both ways
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int ld, int rows, int cols, float* __restrict__ out, int ldo) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = ty; i < 32; i += 8) {
        int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? in[(int64_t)r * ld + c] : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) out[(int64_t)c * ldo + r] = tile[tx][i];
    }
}

template <int MODE, int E, int BN>
static int launch_rowpanel(const CUtensorMap& tmR, const CUtensorMap& tmT, const CUtensorMap& tmTt, const RowPanelParams& p, int m_tiles,
                           int splits, cudaStream_t st, const char* name) {
    using Cfg = RowPanelCfg<MODE, E, BN>;
    static bool attr_done = false;
    if (!attr_done) {
        TT_CUDA_OK(cudaFuncSetAttribute(rowpanel_kernel<MODE, E, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
        attr_done = true;
    }
    dim3 grid((unsigned)m_tiles, (unsigned)splits);
    rowpanel_kernel<MODE, E, BN><<<grid, Cfg::kThreads, Cfg::kSmemBytes, st>>>(tmR, tmT, tmTt, p);
    TT_LAUNCH_OK(name);
    return TT_OK;
}

constexpr int kFwdBN = 128;
constexpr int kMaxSplits = 16;   // column splits per launch; the forward keeps 2 partials per split (one per warp half)

// debug knobs (tt_debug_tc): timeline buffer and a cap on the column splits
static unsigned long long* g_trace = nullptr;
static int g_max_splits = kMaxSplits;
static inline int bwd_bn(int E) { return E <= 64 ? 64 : 32; }  // shared-memory budget (see RowPanelCfg)

struct Plan {
    int m_tiles, n_tiles, splits, tps;
};
static Plan plan_for(int nR, int nT, int bn) {
    Plan pl;
    pl.m_tiles = (int)ceil_div(nR, 128);
    pl.n_tiles = (int)ceil_div(nT, bn);
    choose_splits(pl.m_tiles, pl.n_tiles, 2, g_max_splits, &pl.splits, &pl.tps);
    return pl;
}

static inline size_t tt_ld(int n) { return align_up((size_t)n, 4); }

static size_t bwd_pass_floats(int nR, int nT, int E) {
    Plan pl = plan_for(nR, nT, bwd_bn(E));
    return align_up((size_t)pl.splits * nR * E, 64) + align_up((size_t)E * tt_ld(nT), 64) + (size_t)nT + 512;
}

size_t softmax_tc_workspace(int Bq, int Bc, int E) {
    size_t rows = (size_t)(Bq > Bc ? Bq : Bc);
    size_t fwd = (4 * (size_t)kMaxSplits + 4) * align_up(rows * sizeof(float), 256) + 2048;
    size_t a = bwd_pass_floats(Bq, Bc, E) * sizeof(float), b = bwd_pass_floats(Bc, Bq, E) * sizeof(float);
    size_t bwd = a > b ? a : b;
    return (fwd > bwd ? fwd : bwd) + 1024;
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

int softmax_fwd_tc(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse, float* loss,
                   float* ws, cudaStream_t st) {
    Plan pl = plan_for(Bq, Bc, kFwdBN);
    CUtensorMap tmQ, tmC;
    int rc = make_tmap_2d(&tmQ, Q, Bq, E, ldq, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmC, C, Bc, E, ldc, kFwdBN);
    if (rc) return rc;
    size_t rows = (size_t)(Bq > Bc ? Bq : Bc);
    size_t seg = align_up(rows * sizeof(float), 256) / sizeof(float);
    float* m2 = ws;
    float* l = ws + 2 * (size_t)kMaxSplits * seg;
    float* zd = ws + 4 * (size_t)kMaxSplits * seg;
    float* rowloss = zd + seg;
    float* c2 = rowloss + seg;
    rc = launch_scale_pad(bias, Bc, c2, pl.n_tiles * kFwdBN, st);
    if (rc) return rc;
    RowPanelParams p{};
    p.nR = Bq; p.nT = Bc; p.n_tiles = pl.n_tiles; p.tiles_per_split = pl.tps; p.rowv = nullptr; p.colv2 = c2; p.d = off;
    p.out0 = m2; p.out1 = l; p.out2 = zd; p.ld_out = 0; p.trace = g_trace;
    switch (E) {
        case 32: rc = launch_rowpanel<kFwd, 32, kFwdBN>(tmQ, tmC, tmC, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<fwd,32>"); break;
        case 64: rc = launch_rowpanel<kFwd, 64, kFwdBN>(tmQ, tmC, tmC, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<fwd,64>"); break;
        default: rc = launch_rowpanel<kFwd, 128, kFwdBN>(tmQ, tmC, tmC, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<fwd,128>"); break;
    }
    if (rc) return rc;
    fwd_combine_kernel<<<(unsigned)ceil_div(Bq, 256), 256, 0, st>>>(m2, l, zd, pl.splits * 2 /* kHalves for BN=128 */, Bq, lse, rowloss);
    TT_LAUNCH_OK("fwd_combine_kernel");
    return sum_rows_launch(rowloss, Bq, loss, st);
}

int softmax_bwd_pass_tc(const float* R, int ldr, const float* T, int ldt, const float* rowv, const float* colv, int nR, int nT, int E, int d,
                        float* G, int ldg, float* ws, cudaStream_t st) {
    if (nR == 0) return TT_OK;
    const int bn = bwd_bn(E);
    Plan pl = plan_for(nR, nT, bn);
    float* part = ws;
    float* Tt = ws + align_up((size_t)pl.splits * nR * E, 64);
    const int ldtt = (int)tt_ld(nT);
    float* c2 = Tt + align_up((size_t)E * ldtt, 64);
    {
        int rc0 = launch_scale_pad(colv, nT, c2, pl.n_tiles * bn, st);
        if (rc0) return rc0;
        dim3 grid((unsigned)ceil_div(nT, 32), (unsigned)ceil_div(E, 32));
        transpose_kernel<<<grid, 256, 0, st>>>(T, ldt, nT, E, Tt, ldtt);
        TT_LAUNCH_OK("transpose_kernel");
    }
    CUtensorMap tmR, tmT, tmTt;
    int rc = make_tmap_2d(&tmR, R, nR, E, ldr, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmT, T, nT, E, ldt, bn);
    if (rc) return rc;
    rc = make_tmap_2d(&tmTt, Tt, E, nT, ldtt, E);
    if (rc) return rc;
    RowPanelParams p{};
    p.nR = nR; p.nT = nT; p.n_tiles = pl.n_tiles; p.tiles_per_split = pl.tps; p.rowv = rowv; p.colv2 = c2; p.d = d;
    p.out0 = part; p.out1 = nullptr; p.out2 = nullptr; p.ld_out = 0; p.trace = g_trace;
    switch (E) {
        case 32: rc = launch_rowpanel<kBwd, 32, 64>(tmR, tmT, tmTt, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<bwd,32>"); break;
        case 64: rc = launch_rowpanel<kBwd, 64, 64>(tmR, tmT, tmTt, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<bwd,64>"); break;
        default: rc = launch_rowpanel<kBwd, 128, 32>(tmR, tmT, tmTt, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<bwd,128>"); break;
    }
    if (rc) return rc;
    int64_t n = (int64_t)nR * E;
    bwd_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(part, pl.splits, n, E, G, ldg);
    TT_LAUNCH_OK("bwd_reduce_kernel");
    return TT_OK;
}

int logits_tc(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, float* Z, int ldz, cudaStream_t st) {
    Plan pl = plan_for(Bq, Bc, kFwdBN);
    CUtensorMap tmQ, tmC;
    int rc = make_tmap_2d(&tmQ, Q, Bq, E, ldq, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tmC, C, Bc, E, ldc, kFwdBN);
    if (rc) return rc;
    // test / API-convenience path: the scaled column term lives in a lazily grown static buffer
    static float* c2 = nullptr;
    static size_t c2_cap = 0;
    size_t need = (size_t)pl.n_tiles * kFwdBN;
    if (need > c2_cap) {
        if (c2) cudaFree(c2);
        TT_CUDA_OK(cudaMalloc(&c2, need * sizeof(float)));
        c2_cap = need;
    }
    rc = launch_scale_pad(bias, Bc, c2, (int)need, st);
    if (rc) return rc;
    RowPanelParams p{};
    p.nR = Bq; p.nT = Bc; p.n_tiles = pl.n_tiles; p.tiles_per_split = pl.tps; p.colv2 = c2; p.out0 = Z; p.ld_out = ldz;
    p.d = -(1 << 30);
    switch (E) {
        case 32: return launch_rowpanel<kLogits, 32, kFwdBN>(tmQ, tmC, tmC, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<logits,32>");
        case 64: return launch_rowpanel<kLogits, 64, kFwdBN>(tmQ, tmC, tmC, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<logits,64>");
        default: return launch_rowpanel<kLogits, 128, kFwdBN>(tmQ, tmC, tmC, p, pl.m_tiles, pl.splits, st, "rowpanel_kernel<logits,128>");
    }
}

size_t softmax_tc_workspace_bytes(int Bq, int Bc, int E) { return tc::softmax_tc_workspace(Bq, Bc, E); }

void debug_tc(void* trace, int max_splits) {
    tc::g_trace = reinterpret_cast<unsigned long long*>(trace);
    tc::g_max_splits = (max_splits >= 1 && max_splits <= tc::kMaxSplits) ? max_splits : tc::kMaxSplits;
}

}  // namespace tt
