// tt_tc_rowpanel.cuh -- the one tcgen05 kernel skeleton behind the logits-shaped hot ops.
//
//   S(128 x BN) = R_panel(128 x E) . T_tile(BN x E)^T      tcgen05.mma kind::tf32, fp32 accumulate in TMEM
//
// A CTA owns a 128-row panel of R (loaded once by TMA, resident in shared memory) and streams tiles of T
// through a TMA/mbarrier ring.  S is double-buffered in TMEM so the tensor core works on tile j+1 while
// the epilogue warps consume tile j; S never goes to HBM.  Warp roles (192 threads):
//   warp 0: TMA producer (one elected lane)      warp 1: MMA issuer (one elected lane), TMEM owner
//   warps 2-5: epilogue, one TMEM lane (= one R row) per thread
// Modes (compile-time):
//   kFwd    online log-sum-exp per row (+ diagonal logit)                -> partial (m2, l, zdiag) per split
//   kBwd    P = exp(S - rowv - colv) - [col == row + d], written TF32 to swizzled smem as the A operand of a
//           second MMA  G(128 x E) += P(128 x BN) . T_tile(BN x E).  MN-major TF32 operands need a different
//           swizzle (128B_BASE32B) than the K-major tile of the first MMA, so the B operand of the second
//           MMA is a K-major tile of T^T (E x BN) loaded by TMA from a transposed copy of T.
//   kLogits Z = S - colv written out (tests / TwoTowerModel.call)
//   kIndex  max over each group of 32 consecutive columns                 -> group maxima (filter stage)
#pragma once
#include <math_constants.h>

#include "tt_tc_common.cuh"

namespace tt {
namespace tc {

enum RowPanelMode { kFwd = 0, kBwd = 1, kLogits = 2, kIndex = 3 };

struct RowPanelParams {
    int nR, nT;
    int n_tiles;          // ceil(nT / BN)
    int tiles_per_split;  // tiles handled by one blockIdx.y
    const float* rowv;    // kBwd: per-R-row term (lse or ln p)
    const float* colv;    // per-T-row term (ln p or lse); may be null
    int d;                // diagonal: column == row + d
    float* out0;          // kFwd: m2 [split][nR] | kBwd: G partial [split][nR][E] | kLogits: Z | kIndex: gmax [nR][ld_out]
    float* out1;          // kFwd: l  [split][nR]
    float* out2;          // kFwd: zdiag [nR] (written by the split that owns the diagonal column)
    int ld_out;           // kLogits: ldz | kIndex: row stride of gmax (groups)
};

constexpr float kLog2e = 1.4426950408889634f;

template <int MODE, int E, int BN>
struct RowPanelCfg {
    static constexpr int kSlabs = E / 32;                       // 128-byte K slabs per operand row
    static constexpr int kStages = (MODE == kBwd) ? 3 : 2;      // T ring depth
    static constexpr int kRBytes = kSlabs * 128 * 128;
    static constexpr int kT1Bytes = kSlabs * BN * 128;                       // T tile, K-major over E  (first MMA)
    static constexpr int kT2Bytes = (MODE == kBwd) ? (BN / 32) * E * 128 : 0;  // T^T tile, K-major over BN (second MMA)
    static constexpr int kTBytes = kT1Bytes + kT2Bytes;
    static constexpr int kPSlabs = BN / 32;
    static constexpr int kPBytes = (MODE == kBwd) ? kPSlabs * 128 * 128 : 0;
    static constexpr int kPBufs = (MODE == kBwd) ? 2 : 0;
    static constexpr int kTmemCols = (MODE == kBwd) ? (2 * BN + E <= 256 ? 256 : 512) : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512));
    static constexpr int kSmemBytes = kRBytes + kStages * kTBytes + kPBufs * kPBytes + 1024 /*barriers*/ + 1024 /*alignment slack*/;
    static_assert(E == 32 || E == 64 || E == 128, "E must be 32, 64 or 128");
    static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 up to 256");
    static_assert(MODE != kBwd || 2 * BN + E <= 512, "TMEM budget");
};

struct RowPanelBars {
    uint64_t r_full;
    uint64_t t_full[3], t_empty[3];
    uint64_t s_full[2], s_empty[2];
    uint64_t p_full[2], p_empty[2];
    uint64_t g_full;
    uint32_t tmem_base;
};

template <int MODE, int E, int BN>
__global__ void __launch_bounds__(192, 1)
rowpanel_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmTt,
                const RowPanelParams p) {
    using Cfg = RowPanelCfg<MODE, E, BN>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sR = smem;
    unsigned char* sT = sR + Cfg::kRBytes;
    unsigned char* sP = sT + Cfg::kStages * Cfg::kTBytes;
    RowPanelBars* bars = reinterpret_cast<RowPanelBars*>(sP + Cfg::kPBufs * Cfg::kPBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;
    const int tile_begin = blockIdx.y * p.tiles_per_split;
    const int my_tiles = min(p.tiles_per_split, p.n_tiles - tile_begin);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmR);
        prefetch_tmap(&tmT);
        if (MODE == kBwd) prefetch_tmap(&tmTt);
        mbar_init(&bars->r_full, 1);
        for (int i = 0; i < 3; ++i) { mbar_init(&bars->t_full[i], 1); mbar_init(&bars->t_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_empty[i], 4);
            mbar_init(&bars->p_full[i], 4); mbar_init(&bars->p_empty[i], 1);
        }
        mbar_init(&bars->g_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(&bars->tmem_base, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars->r_full, Cfg::kRBytes);
            for (int s = 0; s < Cfg::kSlabs; ++s) tma_load_2d(sR + s * 128 * 128, &tmR, &bars->r_full, s * 32, m0);
            for (int it = 0; it < my_tiles; ++it) {
                const int stage = it % Cfg::kStages;
                const uint32_t ph = (it / Cfg::kStages) & 1;
                mbar_wait(&bars->t_empty[stage], ph ^ 1);
                mbar_arrive_expect_tx(&bars->t_full[stage], Cfg::kTBytes);
                unsigned char* dst = sT + stage * Cfg::kTBytes;
                for (int s = 0; s < Cfg::kSlabs; ++s) tma_load_2d(dst + s * BN * 128, &tmT, &bars->t_full[stage], s * 32, (tile_begin + it) * BN);
                if constexpr (MODE == kBwd) {
                    for (int s = 0; s < BN / 32; ++s)
                        tma_load_2d(dst + Cfg::kT1Bytes + s * E * 128, &tmTt, &bars->t_full[stage], (tile_begin + it) * BN + s * 32, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc1 = make_idesc_tf32(128, BN, false, false);
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
                for (int k = 0; k < E / 8; ++k) {
                    const uint32_t off = (k >> 2) * 128 * 128 + (k & 3) * 32;
                    const uint32_t offT = (k >> 2) * BN * 128 + (k & 3) * 32;
                    uint64_t ad = make_smem_desc(sR_a + off, 16, 1024);
                    uint64_t bd = make_smem_desc(sT_a + stage * Cfg::kTBytes + offT, 16, 1024);
                    mma_tf32(tmem + acc * BN, ad, bd, idesc1, k > 0 ? 1u : 0u);
                }
                if (MODE != kBwd) mma_commit(&bars->t_empty[stage]);  // T tile is free once S is computed
                mma_commit(&bars->s_full[acc]);
            }
            __syncwarp();
        };
        mbar_wait(&bars->r_full, 0);
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
                    mma_commit(&bars->t_empty[stage]);
                    mma_commit(&bars->p_empty[pb]);
                }
                __syncwarp();
            }
            if (lane == 0) mma_commit(&bars->g_full);
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps (2..5) =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int row_l = q * 32 + lane;              // row within the panel == TMEM lane
        const int row = m0 + row_l;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        float m2 = -CUDART_INF_F, l = 0.f, zd = 0.f;
        bool has_diag = false;
        float rv = 0.f;
        if (MODE == kBwd) rv = (row < p.nR && p.rowv) ? __ldg(p.rowv + row) : 0.f;
        for (int it = 0; it < my_tiles; ++it) {
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int n0 = (tile_begin + it) * BN;
            mbar_wait(&bars->s_full[acc], aph);
            tc_fence_after();
            const int pb = it & 1;
            if (MODE == kBwd) mbar_wait(&bars->p_empty[pb], ((it >> 1) & 1) ^ 1);
            float gm[BN / 32];
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                float v[32];
                tmem_ld_32x32(tmem + lane_addr + acc * BN + c * 32, v);
                const int nb = n0 + c * 32;
                if constexpr (MODE == kFwd) {
                    float cmax = -CUDART_INF_F;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int n = nb + i;
                        float z = -CUDART_INF_F;
                        if (n < p.nT) {
                            z = v[i] - (p.colv ? __ldg(p.colv + n) : 0.f);
                            if (n == row + p.d) { zd = z; has_diag = true; }
                            z *= kLog2e;
                        }
                        v[i] = z;
                        cmax = fmaxf(cmax, z);
                    }
                    if (cmax > -CUDART_INF_F) {
                        const float mn = fmaxf(m2, cmax);
                        float sum = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum += exp2f(v[i] - mn);
                        l = l * exp2f(m2 - mn) + sum;
                        m2 = mn;
                    }
                } else if constexpr (MODE == kBwd) {
                    unsigned char* prow = sP + pb * Cfg::kPBytes + c * 128 * 128 + row_l * 128;
#pragma unroll
                    for (int g4 = 0; g4 < 8; ++g4) {
                        float o[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const int i = g4 * 4 + t;
                            const int n = nb + i;
                            float pv = 0.f;
                            if (n < p.nT && row < p.nR) {
                                const float cv = p.colv ? __ldg(p.colv + n) : 0.f;
                                pv = exp2f((v[i] - rv - cv) * kLog2e);
                                if (n == row + p.d) pv -= 1.0f;
                            }
                            o[t] = tf32_rn(pv);
                        }
                        *reinterpret_cast<float4*>(prow + ((g4 ^ (row_l & 7)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                } else if constexpr (MODE == kLogits) {
                    if (row < p.nR) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int n = nb + i;
                            if (n < p.nT) p.out0[(int64_t)row * p.ld_out + n] = v[i] - (p.colv ? __ldg(p.colv + n) : 0.f);
                        }
                    }
                } else {  // kIndex
                    float mx = -CUDART_INF_F;
#pragma unroll
                    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, (nb + i < p.nT) ? v[i] : -CUDART_INF_F);
                    gm[c] = mx;
                }
            }
            // this S buffer may be overwritten by the MMA of tile it+2
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->s_empty[acc]);
            if constexpr (MODE == kBwd) {
                fence_proxy_async_smem();   // P stores (generic proxy) -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->p_full[pb]);
            }
            if constexpr (MODE == kIndex) {
                if (row < p.nR) {
                    float* dst = p.out0 + (int64_t)row * p.ld_out + (int64_t)(tile_begin + it) * (BN / 32);
#pragma unroll
                    for (int c = 0; c < BN / 32; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(gm[c], gm[c + 1], gm[c + 2], gm[c + 3]);
                }
            }
        }
        if constexpr (MODE == kFwd) {
            if (row < p.nR) {
                p.out0[(int64_t)blockIdx.y * p.nR + row] = m2;
                p.out1[(int64_t)blockIdx.y * p.nR + row] = l;
                if (has_diag) p.out2[row] = zd;
            }
        }
        if constexpr (MODE == kBwd) {
            mbar_wait(&bars->g_full, 0);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < E / 32; ++c) {
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
        // work per CTA is tps tiles; time ~ waves * tps; ideal ~ m_tiles * n_tiles / sms
        double eff = ((double)m_tiles * n_tiles / sms) / ((double)waves * tps);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    }
    *splits = best;
    *tiles_per_split = (n_tiles + best - 1) / best;
}

}  // namespace tc
}  // namespace tt
