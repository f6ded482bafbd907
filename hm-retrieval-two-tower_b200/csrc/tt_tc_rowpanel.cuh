// tt_tc_rowpanel.cuh -- one CTA per (row panel, column split): the tcgen05 skeleton behind the index filter and tt_logits.
//
//   S(128 x BN) = R_panel(128 x E) . T_tile(BN x E)^T      tcgen05.mma kind::tf32, fp32 accumulate in TMEM
//
// A CTA owns a 128-row panel of R (loaded once by TMA, resident in shared memory) and streams tiles of T
// through a TMA/mbarrier ring.  S is double-buffered in TMEM so the tensor core works on tile j+1 while
// the epilogue warps consume tile j; S never goes to HBM.  Warp roles:
//   warp 0: TMA producer (one elected lane)      warp 1: MMA issuer (one elected lane), TMEM owner
//   warps 2-9: epilogue; thread = one TMEM lane (= one R row); the two warps that share a lane quarter split
//              the columns of every tile between them
// Modes (compile-time):
//   kLogits Z = S - colv written out (tests / TwoTowerModel.call)
//   kIndex  lower bound of the best score in each group of BN/2 consecutive columns -> group values (filter stage 1)
//   kCollect every 32-column chunk that can hold a column reaching the row's threshold is appended (32 TF32 scores)
//           to the dumping warp's hit log (filter stage 2; the log is tested per column and rescored exactly afterwards)
// The in-batch softmax (kFwd / kBwd) moved to the persistent kernel in tt_tc_streamk.cuh, which reuses the helpers
// below (mbarrier/TMA/TMEM wrappers, the checked per-chunk epilogue bodies); the kFwd/kBwd branches that remain in this
// kernel body are compiled out (RowPanelCfg static_asserts the mode).
#pragma once
#include <math_constants.h>

#include "tt_tc_common.cuh"

namespace tt {
namespace tc {

enum RowPanelMode { kFwd = 0, kBwd = 1, kLogits = 2, kIndex = 3, kCollect = 4 };

struct RowPanelParams {
    int nR, nT;
    int n_tiles;          // ceil(nT / BN)
    int tiles_per_split;  // tiles handled by one blockIdx.y
    const float* rowv;    // kBwd: per-R-row term (lse or ln p), natural units | kIndex/kCollect: kappa_q = c*||q||
    const float* rowv2;   // kCollect: per-row threshold lambda_q | kLogits: per-COLUMN bias, natural units (may be null)
    const float* gnorm;   // kIndex/kCollect: max ||c_j|| over each chunk of 32 columns (colv2 holds the per-column norms)
    const float* colv2;   // per-T-row term * log2(e), padded with zeros to n_tiles*BN entries (never null)
    int d;                // diagonal: column == row + d
    int fine_groups;      // kIndex: 1 = one filter group per 32-column chunk (small corpora), 0 = one per (tile, warp half)
    float* out0;          // kFwd: m2 [split][nR] | kBwd: G partial [split][nR][E] | kLogits: Z | kIndex: gmax [nR][ld_out]
                          // kCollect: hit logs [CTA][epilogue warp][ld_out][kHitWords]
    float* out1;          // kFwd: l  [split][nR] | kCollect: (int32*) entries in each warp's log
    float* out2;          // kFwd: zdiag [nR] (natural units; written by the split that owns the diagonal column)
                          // kCollect: (int32*) per-row overflow flags (set when a warp's log is full)
    int ld_out;           // kLogits: ldz | kIndex: row stride of gmax (groups) | kCollect: capacity (entries) of one warp's hit log
    unsigned long long* trace;  // optional debug timeline: [cta][16] globaltimer stamps (ns); null in production
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TT_TRACE(slot)                                                                                        \
    do {                                                                                                      \
        if (p.trace) p.trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = gtime();          \
    } while (0)

constexpr int kHitWords = 40;   // hit-log entry: 32-byte header {row, first column, chunk threshold lambda - kappa*max||c||, 5 pad words} + 32 TF32 scores
                                // = 160 bytes = five whole 32-byte sectors, each written completely (a partially written sector costs a DRAM fill on read)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void bulk_copy_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// explicit shared-space accesses (a generic pointer would compile to slower generic LD/ST)
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// round-to-nearest (ties away from zero) to TF32 on the integer pipe: same result as cvt.rna.tf32.f32 for finite
// inputs, without occupying the transcendental/conversion unit the exponentials need
__device__ __forceinline__ float tf32_rn_int(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }

// 3-input maximum (one FMNMX3 on sm_100)
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// maximum of the 32 accumulator words of one chunk: 16 FMNMX3
__device__ __forceinline__ float chunk_max(const uint32_t (&r)[32]) {
    float m0 = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1]));
    float m1 = fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3]));
#pragma unroll
    for (int i = 4; i < 32; i += 4) {
        m0 = fmax3(m0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
        m1 = fmax3(m1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
    }
    return fmaxf(m0, m1);
}

// issue only (no wait): 32 lanes x 32 columns
__device__ __forceinline__ void tmem_ld_32x32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// H: fp16 operands (kind::f16, 64 columns per 128-byte slab) instead of TF32-rounded fp32 (kind::tf32, 32 per slab)
template <int MODE, int E, int BN, bool H = false>
struct RowPanelCfg {
    static constexpr int kSlabCols = H ? 64 : 32;
    static constexpr int kSlabs = E / kSlabCols;                // 128-byte K slabs per operand row
    static constexpr int kMma1 = H ? E / 16 : E / 8;            // MMA instructions per tile (32 bytes of K each)
    static_assert(!H || E >= 64, "fp16 operands need E >= 64");
    static constexpr int kRBytes = kSlabs * 128 * 128;
    static constexpr int kFit = (232448 - 6 * 1024 - kRBytes) / (kSlabs * BN * 128);   // T tiles that fit beside the R panel
    static constexpr int kStages = (MODE == kBwd) ? 3 : (kFit >= 4 ? 4 : (kFit >= 3 ? 3 : 2));   // T ring depth
    static constexpr int kT1Bytes = kSlabs * BN * 128;                       // T tile, K-major over E  (first MMA)
    static constexpr int kT2Bytes = (MODE == kBwd) ? (BN / 32) * E * 128 : 0;  // T^T tile, K-major over BN (second MMA)
    static constexpr int kTBytes = kT1Bytes + kT2Bytes;
    static constexpr int kPSlabs = BN / 32;
    static constexpr int kPBytes = (MODE == kBwd) ? kPSlabs * 128 * 128 : 0;
    static constexpr int kPBufs = (MODE == kBwd) ? 2 : 0;
    static constexpr bool kUsesC2 = (MODE != kIndex && MODE != kCollect && MODE != kLogits);   // staged per-column vector (softmax: colv * log2 e); kLogits reads the bias (rowv2, natural units) from global memory
    static constexpr int kC2Bytes = kUsesC2 ? BN * 4 : 0;
    static constexpr int kHalves = (BN / 32 >= 2) ? 2 : 1;                   // epilogue warps per TMEM lane quarter
    static constexpr int kEpiWarps = 4 * kHalves;
    static constexpr int kThreads = 64 + 32 * kEpiWarps;
    static constexpr int kTmemCols = (MODE == kBwd) ? (2 * BN + E <= 256 ? 256 : 512) : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512));
    static constexpr int kSmemBytes = kRBytes + kStages * kTBytes + kPBufs * kPBytes + 4 * 1024 /*c2 stages*/ + 1024 /*barriers*/ + 1024 /*align*/;
    static_assert(kSmemBytes <= 232448, "shared memory budget");
    static_assert(MODE == kLogits || MODE == kIndex || MODE == kCollect, "rowpanel_kernel: the softmax modes live in tt_tc_streamk.cuh");
    static_assert(E == 32 || E == 64 || E == 128, "E must be 32, 64 or 128");
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 up to 256");
    static_assert(MODE != kBwd || 2 * BN + E <= 512, "TMEM budget");
    static_assert(kC2Bytes <= 1024, "c2 stage");
};

struct RowPanelBars {
    uint64_t r_full;
    uint64_t t_full[4], t_empty[4];
    uint64_t s_full[2], s_empty[2];
    uint64_t p_full[2], p_empty[2];
    uint64_t g_full;
    uint32_t tmem_base;
};

// ---- per-chunk epilogue bodies (32 columns of one row), check-free when FAST ---------------------------
template <bool FAST>
__device__ __forceinline__ void fwd_chunk(const uint32_t (&r)[32], uint32_t c2s, int nb, int row, int nT, int d, float& m2, float& l, float& zd,
                                          bool& has_diag) {
    float z[32];
    float cm0 = -CUDART_INF_F, cm1 = -CUDART_INF_F;
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = g4 * 4 + t;
            float zz = fmaf(__uint_as_float(r[i]), kLog2e, -cv[t]);
            if (!FAST) {
                const int n = nb + i;
                if (n >= nT) zz = -CUDART_INF_F;
                else if (n == row + d) { zd = zz * kLn2; has_diag = true; }
            }
            z[i] = zz;
            if (t & 1) cm1 = fmaxf(cm1, zz); else cm0 = fmaxf(cm0, zz);
        }
    }
    const float cmax = fmaxf(cm0, cm1);
    if (FAST || cmax > -CUDART_INF_F) {
        const float mn = fmaxf(m2, cmax);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            s0 += ex2_approx(z[i] - mn); s1 += ex2_approx(z[i + 1] - mn);
            s2 += ex2_approx(z[i + 2] - mn); s3 += ex2_approx(z[i + 3] - mn);
        }
        l = l * ex2_approx(m2 - mn) + ((s0 + s1) + (s2 + s3));
        m2 = mn;
    }
}

template <bool FAST>
__device__ __forceinline__ void bwd_chunk(const uint32_t (&r)[32], uint32_t c2s, int nb, int row, int row_l, int nT, int nR, int d, float r2,
                                          uint32_t prow) {
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
        float o[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = g4 * 4 + t;
            float pv = ex2_approx(fmaf(__uint_as_float(r[i]), kLog2e, -r2) - cv[t]);
            if (!FAST) {
                const int n = nb + i;
                if (n >= nT || row >= nR) pv = 0.f;
                else if (n == row + d) pv -= 1.0f;
            }
            o[t] = tf32_rn_int(pv);
        }
        sts128(prow + ((g4 ^ (row_l & 7)) << 4), o[0], o[1], o[2], o[3]);
    }
}

template <int MODE, int E, int BN, bool H = false>
__global__ void __launch_bounds__(64 + 32 * 4 * ((BN / 32 >= 2) ? 2 : 1), 1)
rowpanel_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmTt,
                const RowPanelParams p) {
    using Cfg = RowPanelCfg<MODE, E, BN, H>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sR = smem;
    unsigned char* sT = sR + Cfg::kRBytes;
    unsigned char* sP = sT + Cfg::kStages * Cfg::kTBytes;
    unsigned char* sC2 = sP + Cfg::kPBufs * Cfg::kPBytes;       // kStages x 1 KB
    RowPanelBars* bars = reinterpret_cast<RowPanelBars*>(sC2 + 4 * 1024);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;
    const int tile_begin = blockIdx.y * p.tiles_per_split;
    const int my_tiles = min(p.tiles_per_split, p.n_tiles - tile_begin);
    if (threadIdx.x == 0) TT_TRACE(0);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmR);
        prefetch_tmap(&tmT);
        if (MODE == kBwd) prefetch_tmap(&tmTt);
        mbar_init(&bars->r_full, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&bars->t_full[i], 1); mbar_init(&bars->t_empty[i], (MODE == kBwd || !Cfg::kUsesC2) ? 1 : 1 + Cfg::kEpiWarps); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_empty[i], Cfg::kEpiWarps);
            mbar_init(&bars->p_full[i], Cfg::kEpiWarps); mbar_init(&bars->p_empty[i], 1);
        }
        mbar_init(&bars->g_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    if (threadIdx.x == 0) TT_TRACE(1);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars->r_full, Cfg::kRBytes);
            for (int s = 0; s < Cfg::kSlabs; ++s) tma_load_2d(sR + s * 128 * 128, &tmR, &bars->r_full, s * Cfg::kSlabCols, m0);
            for (int it = 0; it < my_tiles; ++it) {
                const int stage = it % Cfg::kStages;
                const uint32_t ph = (it / Cfg::kStages) & 1;
                const int n0 = (tile_begin + it) * BN;
                mbar_wait(&bars->t_empty[stage], ph ^ 1);
                mbar_arrive_expect_tx(&bars->t_full[stage], Cfg::kTBytes + Cfg::kC2Bytes);
                unsigned char* dst = sT + stage * Cfg::kTBytes;
                for (int s = 0; s < Cfg::kSlabs; ++s) tma_load_2d(dst + s * BN * 128, &tmT, &bars->t_full[stage], s * Cfg::kSlabCols, n0);
                if (MODE == kBwd) {
                    for (int s = 0; s < BN / 32; ++s) tma_load_2d(dst + Cfg::kT1Bytes + s * E * 128, &tmTt, &bars->t_full[stage], n0 + s * 32, 0);
                }
                if (Cfg::kUsesC2) bulk_copy_1d(sC2 + stage * 1024, p.colv2 + n0, Cfg::kC2Bytes, &bars->t_full[stage]);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc1 = H ? make_idesc_f16(128, BN) : make_idesc_tf32(128, BN, false, false);
        constexpr uint32_t idesc2 = make_idesc_tf32(128, E, false, false);
        const uint32_t sR_a = smem_u32(sR), sT_a = smem_u32(sT), sP_a = smem_u32(sP);
        auto issue_g1 = [&](int it) {
            const int stage = it % Cfg::kStages;
            const uint32_t tph = (it / Cfg::kStages) & 1;
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            mbar_wait(&bars->t_full[stage], tph);
            mbar_wait(&bars->s_empty[acc], aph ^ 1);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < Cfg::kMma1; ++k) {
                    const uint32_t off = (k >> 2) * 128 * 128 + (k & 3) * 32;
                    const uint32_t offT = (k >> 2) * BN * 128 + (k & 3) * 32;
                    uint64_t ad = make_smem_desc(sR_a + off, 16, 1024);
                    uint64_t bd = make_smem_desc(sT_a + stage * Cfg::kTBytes + offT, 16, 1024);
                    if (H) mma_f16(tmem + acc * BN, ad, bd, idesc1, k > 0 ? 1u : 0u);
                    else mma_tf32(tmem + acc * BN, ad, bd, idesc1, k > 0 ? 1u : 0u);
                }
                // non-bwd: the stage is released by this commit (T tile consumed) AND by the 4 epilogue warps (c2 consumed)
                if (MODE != kBwd) mma_commit(&bars->t_empty[stage]);
                mma_commit(&bars->s_full[acc]);
            }
            __syncwarp();
        };
        mbar_wait(&bars->r_full, 0);
        if (lane == 0) TT_TRACE(2);
        if constexpr (MODE != kBwd) {
            for (int it = 0; it < my_tiles; ++it) issue_g1(it);
        } else {
            issue_g1(0);
            for (int it = 0; it < my_tiles; ++it) {
                if (it + 1 < my_tiles) issue_g1(it + 1);  // keep the tensor core busy while the epilogue builds P(it)
                const int stage = it % Cfg::kStages;
                const int pb = it & 1;
                const uint32_t pph = (it >> 1) & 1;
                mbar_wait(&bars->p_full[pb], pph);
                tc_fence_after();
                if (lane == 0) {
#pragma unroll
                    for (int kk = 0; kk < BN / 8; ++kk) {
                        uint64_t ad = make_smem_desc(sP_a + pb * Cfg::kPBytes + (kk >> 2) * 128 * 128 + (kk & 3) * 32, 16, 1024);
                        uint64_t bd = make_smem_desc(sT_a + stage * Cfg::kTBytes + Cfg::kT1Bytes + (kk >> 2) * E * 128 + (kk & 3) * 32, 16, 1024);
                        mma_tf32(tmem + 2 * BN, ad, bd, idesc2, (it > 0 || kk > 0) ? 1u : 0u);
                    }
                    mma_commit(&bars->t_empty[stage]);  // T, T^T and c2 of this stage are all consumed by now
                    mma_commit(&bars->p_empty[pb]);
                }
                __syncwarp();
            }
            if (lane == 0) mma_commit(&bars->g_full);
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps (2..) =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;             // which share of each tile's columns this warp handles
        const int row_l = q * 32 + lane;              // row within the panel == TMEM lane
        const int row = m0 + row_l;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int wrow0 = m0 + q * 32;                // first row of this warp
        float m2 = -CUDART_INF_F, l = 0.f, zd = 0.f;
        bool has_diag = false;
        float r2 = 0.f;
        if (MODE == kBwd) r2 = (row < p.nR && p.rowv) ? __ldg(p.rowv + row) * kLog2e : 0.f;
        float r3 = CUDART_INF_F;
        if (MODE == kIndex || MODE == kCollect) r2 = (row < p.nR) ? __ldg(p.rowv + row) : 0.f;      // kappa_q
        if (MODE == kCollect) r3 = (row < p.nR) ? __ldg(p.rowv2 + row) : CUDART_INF_F;              // lambda_q
        constexpr int NC = BN / 32;
        constexpr int NCW = NC / Cfg::kHalves;        // 32-column chunks per warp per tile
        const int c_first = half * NCW;
        // kCollect: this warp's hit log (dense: the entries of its 32 lanes back to back, so the consumer reads whole lines;
        // slots come from a warp ballot -- no atomics)
        const int log_id = (blockIdx.y * gridDim.x + blockIdx.x) * Cfg::kEpiWarps + (warp - 2);
        float* qlog = (MODE == kCollect) ? p.out0 + (int64_t)log_id * p.ld_out * kHitWords : nullptr;
        int n_log = 0;                                // entries appended so far (warp-uniform)
        for (int it = 0; it < my_tiles; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int stage = it % Cfg::kStages;
            const int n0 = (tile_begin + it) * BN;
            const int pb = it & 1;
            // warp-uniform: tile fully in range and no diagonal element of this warp's rows inside it
            const bool fast = (n0 + BN <= p.nT) && (wrow0 + 32 <= p.nR) && (wrow0 + p.d + 32 <= n0 || wrow0 + p.d >= n0 + BN);
            float gn[NCW];                            // kIndex/kCollect: chunk norm maxima, fetched before the wait on the accumulator
            if constexpr (MODE == kIndex || MODE == kCollect) {
                const float* src = p.gnorm + (n0 >> 5) + c_first;
                if constexpr (NCW % 4 == 0) {
#pragma unroll
                    for (int c4 = 0; c4 < NCW; c4 += 4) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(src + c4));
                        gn[c4] = v.x; gn[c4 + 1] = v.y; gn[c4 + 2] = v.z; gn[c4 + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int c1 = 0; c1 < NCW; ++c1) gn[c1] = __ldg(src + c1);
                }
            }
            mbar_wait(&bars->s_full[acc], aph);
            tc_fence_after();
            if (Cfg::kUsesC2) mbar_wait(&bars->t_full[stage], (it / Cfg::kStages) & 1);   // completed long ago; acquires the staged column term
            if (threadIdx.x == 64 && it < 6) TT_TRACE(3 + 2 * it);
            if (MODE == kBwd) mbar_wait(&bars->p_empty[pb], ((it >> 1) & 1) ^ 1);
            const float4* c2v = reinterpret_cast<const float4*>(sC2 + stage * 1024);
            const uint32_t c2s = smem_u32(sC2 + stage * 1024);
            uint32_t rbuf[2][32];
            float gmx = -CUDART_INF_F;                // kIndex: the group (= this warp's share of the tile) lower bound
            tmem_ld_32x32_issue(tmem + lane_addr + acc * BN + c_first * 32, rbuf[0]);
#pragma unroll
            for (int cl = 0; cl < NCW; ++cl) {
                const int c = c_first + cl;
                tmem_ld_wait();
                if (cl + 1 < NCW) tmem_ld_32x32_issue(tmem + lane_addr + acc * BN + (c + 1) * 32, rbuf[(cl + 1) & 1]);
                uint32_t(&r)[32] = rbuf[cl & 1];
                const int nb = n0 + c * 32;
                if constexpr (MODE == kFwd) {
                    if (fast) fwd_chunk<true>(r, c2s + c * 128, nb, row, p.nT, p.d, m2, l, zd, has_diag);
                    else fwd_chunk<false>(r, c2s + c * 128, nb, row, p.nT, p.d, m2, l, zd, has_diag);
                } else if constexpr (MODE == kBwd) {
                    const uint32_t prow = smem_u32(sP + pb * Cfg::kPBytes + c * 128 * 128 + row_l * 128);
                    if (fast) bwd_chunk<true>(r, c2s + c * 128, nb, row, row_l, p.nT, p.nR, p.d, r2, prow);
                    else bwd_chunk<false>(r, c2s + c * 128, nb, row, row_l, p.nT, p.nR, p.d, r2, prow);
                } else if constexpr (MODE == kLogits) {
                    if (row < p.nR) {
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) {
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const int n = nb + g4 * 4 + t;
                                if (n < p.nT) p.out0[(int64_t)row * p.ld_out + n] = __uint_as_float(r[g4 * 4 + t]) - (p.rowv2 ? __ldg(p.rowv2 + n) : 0.f);
                            }
                        }
                    }
                } else if constexpr (MODE == kCollect) {
                    // A chunk can hold a column with a_j + kappa*||c_j|| >= lambda only if max_j a_j >= lambda - kappa*max_j||c_j||.
                    // The (rare, ~K per row) qualifying chunks are appended whole -- row, first column, threshold, 32 TF32 scores --
                    // to the CTA's hit log; column_test_kernel does the per-column test.  One shared-memory atomic per hit.
                    const float thr_c = fmaf(-r2, gn[cl], r3);
                    const bool hit = chunk_max(r) >= thr_c;
                    const unsigned hits = __ballot_sync(0xffffffffu, hit);
                    if (hit) {
                        const int slot = n_log + __popc(hits & ((1u << lane) - 1u));
                        if (slot < p.ld_out) {
                            float4* dst = reinterpret_cast<float4*>(qlog + (int64_t)slot * kHitWords);
                            dst[0] = make_float4(__int_as_float(row), __int_as_float(nb), thr_c, 0.f);
                            dst[1] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int g4 = 0; g4 < 8; ++g4)
                                dst[2 + g4] = make_float4(__uint_as_float(r[g4 * 4]), __uint_as_float(r[g4 * 4 + 1]), __uint_as_float(r[g4 * 4 + 2]),
                                                          __uint_as_float(r[g4 * 4 + 3]));
                        } else {
                            reinterpret_cast<int32_t*>(p.out2)[row] = 1;   // log full: this query goes to the exact fallback
                        }
                    }
                    n_log += __popc(hits);
                } else {  // kIndex: per chunk, max_j a_j - kappa * max_j ||c_j||  <=  max_j (a_j - kappa*||c_j||)  <=  max_j s_j
                    float mx = -CUDART_INF_F;
                    if (fast || nb + 32 <= p.nT) {
                        mx = chunk_max(r);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (nb + i < p.nT) ? __uint_as_float(r[i]) : -CUDART_INF_F);
                    }
                    const float lb = fmaf(-r2, gn[cl], mx);
                    gmx = fmaxf(gmx, lb);
                    if (p.fine_groups && row < p.nR) p.out0[(int64_t)row * p.ld_out + (int64_t)(tile_begin + it) * NC + c] = lb;   // small corpora
                }
            }
            // this S buffer may be overwritten by the MMA of tile it+2
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bars->s_empty[acc]);
                if (MODE != kBwd && Cfg::kUsesC2) mbar_arrive(&bars->t_empty[stage]);   // c2 of this stage consumed
            }
            if (threadIdx.x == 64 && it < 6) TT_TRACE(4 + 2 * it);
            if constexpr (MODE == kBwd) {
                fence_proxy_async_smem();   // P stores (generic proxy) -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->p_full[pb]);
            }
            if constexpr (MODE == kIndex) {
                if (!p.fine_groups && row < p.nR) p.out0[(int64_t)row * p.ld_out + (int64_t)(tile_begin + it) * Cfg::kHalves + half] = gmx;
            }
        }
        if constexpr (MODE == kCollect) {
            if (lane == 0) reinterpret_cast<int32_t*>(p.out1)[log_id] = min(n_log, p.ld_out);
        }
        if constexpr (MODE == kFwd) {
            if (row < p.nR) {
                // one partial per (column split, warp half): the combine kernel merges gridDim.y * kHalves of them
                p.out0[((int64_t)blockIdx.y * Cfg::kHalves + half) * p.nR + row] = m2;
                p.out1[((int64_t)blockIdx.y * Cfg::kHalves + half) * p.nR + row] = l;
                if (has_diag) p.out2[row] = zd;
            }
        }
        if constexpr (MODE == kBwd) {
            mbar_wait(&bars->g_full, 0);
            tc_fence_after();
#pragma unroll
            for (int c = half; c < E / 32; c += Cfg::kHalves) {
                float v[32];
                tmem_ld_32x32(tmem + lane_addr + 2 * BN + c * 32, v);
                if (row < p.nR) {
                    float4* dst = reinterpret_cast<float4*>(p.out0 + ((int64_t)blockIdx.y * p.nR + row) * E + c * 32);
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) dst[g4] = make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) TT_TRACE(15);
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, Cfg::kTmemCols);
    }
}

// split the n tiles over blockIdx.y so that (m_tiles * splits) fills whole waves of the SMs
inline void choose_splits(int m_tiles, int n_tiles, int min_tiles_per_split, int max_splits, int* splits, int* tiles_per_split) {
    const int sms = sm_count();
    int best = 1;
    double best_eff = -1.0;
    for (int s = 1; s <= max_splits && s <= n_tiles; ++s) {
        int tps = (n_tiles + s - 1) / s;
        if (tps < min_tiles_per_split && s > 1) break;
        int real = (n_tiles + tps - 1) / tps;  // non-empty splits
        if (real != s) continue;
        int64_t ctas = (int64_t)m_tiles * s;
        int64_t waves = (ctas + sms - 1) / sms;
        // per-CTA cost ~ tps tiles + ~2 tiles of fixed prologue; ideal ~ m_tiles * n_tiles / sms
        double eff = ((double)m_tiles * n_tiles / sms) / ((double)waves * (tps + 2.0));
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    }
    *splits = best;
    *tiles_per_split = (n_tiles + best - 1) / best;
}

}  // namespace tc
}  // namespace tt
