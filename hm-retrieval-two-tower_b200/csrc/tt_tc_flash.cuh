// tt_tc_flash.cuh -- the in-batch sampled softmax as two persistent tcgen05 passes over the (row panel, column tile) grid.
//
//   pass 1 (kP1)  forward AND dQ in one sweep:   S = Q_panel . C_tile^T  ->  online softmax (lazy rescale)  ->  P~ (fp16, in TMEM)
//                 G(128 x E) += P~ . C_tile;  per row: reference exponent, sum of P~, diagonal logit.  The combine kernel turns
//                 (G, sum) into lse, loss and dQ = G / sum - (1 - p_ii) c_ii.
//   pass 2 (kP2)  dC (or any side given lse):    S = C_panel . Q_tile^T  ->  P = 2^(z - lse) - [diagonal]  ->  G += P . Q_tile
//
// Two exponentials and four tile products per logit for forward + backward (the three-sweep form -- forward, dQ pass, dC pass --
// needs three and five).  One CTA per SM walks an equal share of the unit list (stream-K); a unit is a PAIR of 128-row panels
// times one BN-column tile: the streamed tile is loaded once and used by four MMAs, and the two panels belong to two epilogue
// warpgroups that alternate on the tensor pipe (one computes exponentials while the other's products run).
//
// Operands are fp16 copies scaled by a per-tensor power of two (tt_softmax_flash.cu: amax -> scale), so any finite fp32 input is in
// range; products are exact and accumulate in fp32 in TMEM.  The second product reads P from TENSOR MEMORY (A operand) and the
// SAME shared-memory tile as the first (as an MN-major B operand): no transposed copies exist anywhere.
//
// Thread = one row (TMEM lane); a tile is consumed in chunks of 32 columns.  Pass 1 keeps a per-row reference exponent that is only
// raised when a chunk exceeds it by more than 2^kTau (then the row's running sum, its G row in TMEM and the P chunks already
// written for this tile are rescaled -- rare after the first tile of a panel).  Every mbarrier wait is bounded.
#pragma once
#include "tt_tc_streamk.cuh"

namespace tt {
namespace tc {

enum FlashMode { kP1 = 0, kP2 = 1 };

constexpr float kTau = 8.f;      // pass 1: a row's reference exponent may lag its true maximum by up to 2^kTau
constexpr float kOff1 = 6.f;     // pass 1: P~ = 2^(z - ref + kOff1) <= 2^(kTau + kOff1) = 2^14 < 65504
constexpr float kOff2 = 14.f;    // pass 2: P' = 2^(z - lse + kOff2) <= 2^14; fp16 normals then reach down to p = 2^-28

struct FlPass {
    int nR, nT;
    int m_pairs, n_tiles;   // pairs of 128-row panels of R; BN-row tiles of T
    int d;                  // diagonal: column == row + d
    int unit0;
    const float* rowv;      // kP2: per-R-row term, natural units (lse or ln p); may be null
    const float* colv2;     // per-T-row term * log2(e), zero padded to n_tiles*BN entries
    float* out_g;           // G partials [slot][m_pairs*256][E]
    float* out_m;           // kP1: reference exponent (log2 units) [slot][m_pairs*256]
    float* out_l;           // kP1: sum of P~                      [slot][m_pairs*256]
    float* out_zd;          // kP1: diagonal logit, log2 units     [m_pairs*256]
    const float* diag_pm1;  // kP2: (p - 1) of the positive in T-row (column) n, computed without cancellation by the pass-1 combine; may be null
};
struct FlParams {
    FlPass pass[2];
    int n_pass;
    int units;
    const float* kmul;      // device scalar per pass: log2(e) / (scale_R * scale_T)   (kmul[pass])
    int mn_lbo, mn_sbo;     // debug: descriptor fields (bytes) of the MN-major B operand; 0 = defaults
    unsigned long long* trace;
};
struct FlMaps {
    CUtensorMap r[2], t[2];   // per pass: R panels (box 128 rows x 64 fp16), T tiles (box BN rows x 64 fp16)
};

struct FlCursor {
    int pass, pair, tile;
    __device__ __forceinline__ void init(const FlParams& p, int u) {
        pass = (p.n_pass > 1 && u >= p.pass[1].unit0) ? 1 : 0;
        const int local = u - p.pass[pass].unit0;
        pair = local / p.pass[pass].n_tiles;
        tile = local - pair * p.pass[pass].n_tiles;
    }
    __device__ __forceinline__ void next(const FlParams& p) {
        if (++tile == p.pass[pass].n_tiles) {
            tile = 0;
            if (++pair == p.pass[pass].m_pairs) { pair = 0; ++pass; }
        }
    }
};

template <int MODE, int E, int BN>
struct FlCfg {
    static_assert(E == 64 || E == 128, "flash softmax: E must be 64 or 128 (fp16 slabs of 64 columns)");
    static_assert(BN == 64 || BN == 128, "BN must be 64 or 128");
    static constexpr int kSlabs = E / 64;
    static constexpr int kPanelBytes = kSlabs * 128 * 128;        // one R panel
    static constexpr int kRBytes = 2 * kPanelBytes;               // the pair
    static constexpr int kTBytes = kSlabs * BN * 128;             // one T tile (K-major over E for MMA1 == MN-major over E for MMA2)
    static constexpr int kC2Bytes = BN * 4;
    static constexpr int kFixed = kRBytes + 8 * kC2Bytes + 1024 /*barriers*/ + 1024 /*align*/;
    static constexpr int kFit = (232448 - kFixed) / kTBytes;
    static constexpr int kStages = kFit >= 8 ? 8 : kFit;
    static constexpr int kSmemBytes = kFixed + kStages * kTBytes;
    static constexpr int kMma1 = E / 16;                          // K = 16 per instruction
    static constexpr int kMma2 = BN / 16;
    static constexpr int kChunks = BN / 32;
    static constexpr int kPCols = BN / 2;
    // TMEM map per warpgroup (256 columns each): S | P | G
    static constexpr int kSCol = 0, kPCol = BN, kGCol = BN + BN / 2;
    static_assert(kGCol + E <= 256, "TMEM budget");
    static constexpr int kThreads = 32 * 11;                      // 2 x 4 epilogue warps, producer, MMA1 issuer, MMA2 issuer
    static_assert(kStages >= 3, "shared memory budget");
};

// S and P of a panel are handed over in two HALVES of BN/2 columns, each with its own full/empty pair: the first product of the
// next tile refills the first half of S while the epilogue is still working on the second, which is what a second S buffer
// would buy (tensor memory has no room for one: 2 x (S + P + G) = 512 columns)
struct FlBars {
    uint64_t r_full, r_empty;
    uint64_t t_full[8], t_empty[8];
    uint64_t s_full[2][2], s_empty[2][2];   // [panel of the pair][half]
    uint64_t p_full[2][2], p_empty[2][2];
    uint64_t g_full[2], g_empty[2];
    uint32_t tmem_base;
};

// kind::f16, fp16 operands, fp32 accumulate; B operand MN-major (the streamed tile as stored: rows of T, E contiguous)
__host__ __device__ constexpr uint32_t make_idesc_f16_bmn(int M, int N) { return make_idesc_f16(M, N) | (1u << 16); }

__device__ __forceinline__ void tmem_ld_32x16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&w)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]),
        "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]),
        "r"(w[23]), "r"(w[24]), "r"(w[25]), "r"(w[26]), "r"(w[27]), "r"(w[28]), "r"(w[29]), "r"(w[30]), "r"(w[31])
        : "memory");
}
__device__ __forceinline__ void unpack_f16x2(uint32_t w, float& lo, float& hi) {
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(w));
}

// ---- pass 1, phase A: negated log2-domain logits zn = c2_j - s*kmul of one 32-column chunk, and their minimum -----------------
__device__ __forceinline__ float p1_zn_fast(const uint32_t (&r)[32], uint32_t c2s, float kmul, f32x2 (&zn)[16]) {
    const f32x2 nk = pk2(-kmul, -kmul);
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        zn[2 * g4] = fma2(pk2(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1])), nk, pk2(cc.x, cc.y));
        zn[2 * g4 + 1] = fma2(pk2(__uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3])), nk, pk2(cc.z, cc.w));
    }
    float a0, a1, b0, b1;
    upk2(zn[0], a0, a1);
    upk2(zn[1], b0, b1);
    float mn0 = fminf(a0, a1), mn1 = fminf(b0, b1);
#pragma unroll
    for (int i = 2; i < 16; i += 2) {
        upk2(zn[i], a0, a1);
        upk2(zn[i + 1], b0, b1);
        mn0 = fmin3(mn0, a0, a1);
        mn1 = fmin3(mn1, b0, b1);
    }
    return fminf(mn0, mn1);
}
// checked form: columns >= nT become +inf (weight 0); returns the chunk-local index of the row's diagonal column or -1
__device__ __forceinline__ float p1_zn_checked(const uint32_t (&r)[32], uint32_t c2s, float kmul, int nb, int nT, int dcol_abs, f32x2 (&zn)[16],
                                               int& dloc) {
    float mn = CUDART_INF_F;
    dloc = -1;
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
        float z[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = g4 * 4 + t;
            float v = fmaf(__uint_as_float(r[i]), -kmul, cv[t]);
            if (nb + i >= nT) v = CUDART_INF_F;
            if (nb + i == dcol_abs) dloc = i;
            z[t] = v;
            mn = fminf(mn, v);
        }
        zn[2 * g4] = pk2(z[0], z[1]);
        zn[2 * g4 + 1] = pk2(z[2], z[3]);
    }
    return mn;
}
// ---- pass 1, phase B: P~ = 2^(a - zn) as packed fp16, running sum --------------------------------------------------------------
__device__ __forceinline__ void p1_exp_fast(const f32x2 (&zn)[16], float a, f32x2& lsum, uint32_t (&w)[16]) {
    const f32x2 mone = pk2(-1.f, -1.f), aa = pk2(a, a);
    f32x2 s1 = pk2(0.f, 0.f);                                   // second accumulator: halves the dependent FADD2 chain
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float x0, x1;
        upk2(fma2(zn[i], mone, aa), x0, x1);
        const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
        if (i & 1) s1 = add2(s1, pk2(p0, p1));
        else lsum = add2(lsum, pk2(p0, p1));
        w[i] = pack_f16x2(p0, p1);
    }
    lsum = add2(lsum, s1);
}
__device__ __forceinline__ void p1_exp_checked(const f32x2 (&zn)[16], float a, int dloc, f32x2& lsum, uint32_t (&w)[16], float& zd) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float z0, z1;
        upk2(zn[i], z0, z1);
        float p0 = ex2_approx(a - z0), p1 = ex2_approx(a - z1);     // +inf -> 0
        // the positive is left out of the sum and of the product: the combine kernel forms p_ii - 1 = -sum_offdiag / sum without
        // cancellation and adds (p_ii - 1) c_ii in fp32
        if (dloc == 2 * i) { p0 = 0.f; zd = -z0; }
        if (dloc == 2 * i + 1) { p1 = 0.f; zd = -z1; }
        lsum = add2(lsum, pk2(p0, p1));
        w[i] = pack_f16x2(p0, p1);
    }
}
// ---- pass 2: P' = 2^(s*kmul - c2_j - r2 + kOff2) - [diagonal] 2^kOff2 -----------------------------------------------------------
__device__ __forceinline__ void p2_chunk_fast(const uint32_t (&r)[32], uint32_t c2s, float kmul, float rowc, uint32_t (&w)[16]) {
    const f32x2 km = pk2(kmul, kmul), mone = pk2(-1.f, -1.f), rc = pk2(rowc, rowc);
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        const f32x2 ad0 = fma2(pk2(cc.x, cc.y), mone, rc), ad1 = fma2(pk2(cc.z, cc.w), mone, rc);
        float a0, a1, b0, b1;
        upk2(fma2(pk2(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1])), km, ad0), a0, a1);
        upk2(fma2(pk2(__uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3])), km, ad1), b0, b1);
        w[2 * g4] = pack_f16x2(ex2_approx(a0), ex2_approx(a1));
        w[2 * g4 + 1] = pack_f16x2(ex2_approx(b0), ex2_approx(b1));
    }
}
__device__ __forceinline__ void p2_chunk_checked(const uint32_t (&r)[32], uint32_t c2s, float kmul, float rowc, int nb, int nT, bool row_ok,
                                                 int dcol_abs, const float* __restrict__ diag_pm1, uint32_t (&w)[16]) {
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
        float pv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = g4 * 4 + t;
            float v = ex2_approx(fmaf(__uint_as_float(r[i]), kmul, rowc - cv[t]));
            if (nb + i >= nT || !row_ok) v = 0.f;
            else if (nb + i == dcol_abs) v = diag_pm1 ? __ldg(diag_pm1 + nb + i) * 16384.f : v - 16384.f;   // (p - 1) 2^kOff2
            pv[t] = v;
        }
        w[2 * g4] = pack_f16x2(pv[0], pv[1]);
        w[2 * g4 + 1] = pack_f16x2(pv[2], pv[3]);
    }
}

// pass 1, rare path: raise the reference exponent of the rows whose chunk exceeds it by more than 2^kTau; rescale their running
// sums, their G rows (valid once a second product of this segment has been issued; the caller has waited for all of them) and
// the P chunks of the current half that are already written (earlier halves are already inside G)
struct P1State { float a; f32x2 l; };   // returned by value: a by-reference state would live in local memory in the hot loop
template <int E>
__device__ __noinline__ P1State p1_raise(bool need, float cmin, bool g_valid, int chunks_done, uint32_t tG, uint32_t tP, float a_run, f32x2 lsum) {
    const float a_new = need ? cmin + kOff1 : a_run;
    const float sc = need ? ex2_approx(a_new - a_run) : 1.f;   // a_run = +inf (first chunk of a segment) -> 0
    lsum = fma2(lsum, pk2(sc, sc), pk2(0.f, 0.f));
    if (g_valid) {
#pragma unroll 1
        for (int cg = 0; cg < E / 32; ++cg) {
            uint32_t gv[32];
            tmem_ld_32x32_issue(tG + cg * 32, gv);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) gv[i] = __float_as_uint(__uint_as_float(gv[i]) * sc);
            tmem_st_32x32(tG + cg * 32, gv);
        }
    }
#pragma unroll 1
    for (int pc = 0; pc < chunks_done; ++pc) {
        uint32_t pw[16];
        tmem_ld_32x16_issue(tP + pc * 16, pw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float lo, hi;
            unpack_f16x2(pw[i], lo, hi);
            pw[i] = pack_f16x2(lo * sc, hi * sc);
        }
        tmem_st_32x16(tP + pc * 16, pw);
    }
    tmem_st_wait();
    return P1State{a_new, lsum};
}

#define FL_TRACE(it, ev)                                                                                     \
    do {                                                                                                     \
        if (p.trace && (it) < 64) p.trace[((size_t)blockIdx.x * 64 + (it)) * 8 + (ev)] = gtime();             \
    } while (0)

template <int MODE, int E, int BN>
__global__ void __launch_bounds__(FlCfg<MODE, E, BN>::kThreads, 1)
flash_kernel(const __grid_constant__ FlMaps maps, const __grid_constant__ FlParams p) {
    using Cfg = FlCfg<MODE, E, BN>;
    constexpr int kHC = Cfg::kChunks / 2;      // 32-column chunks per half
    constexpr int kHN = BN / 2;                // columns per half
    const int u_begin = sk_begin(blockIdx.x, p.units, gridDim.x), u_end = sk_begin(blockIdx.x + 1, p.units, gridDim.x);
    const int my_units = u_end - u_begin;
    if (my_units <= 0) return;

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sR = smem;
    unsigned char* sT = sR + Cfg::kRBytes;
    unsigned char* sC2 = sT + Cfg::kStages * Cfg::kTBytes;
    FlBars* bars = reinterpret_cast<FlBars*>(sC2 + 8 * Cfg::kC2Bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kProducerWarp = 8, kMma1Warp = 9, kMma2Warp = 10;
    if (warp == kProducerWarp && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            if (i < p.n_pass) { prefetch_tmap(&maps.r[i]); prefetch_tmap(&maps.t[i]); }
        }
        mbar_init(&bars->r_full, 1);
        mbar_init(&bars->r_empty, 1);
        for (int i = 0; i < 8; ++i) { mbar_init(&bars->t_full[i], 1); mbar_init(&bars->t_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            for (int h = 0; h < 2; ++h) {
                mbar_init(&bars->s_full[i][h], 1); mbar_init(&bars->s_empty[i][h], 4);
                mbar_init(&bars->p_full[i][h], 4); mbar_init(&bars->p_empty[i][h], 1);
            }
            mbar_init(&bars->g_full[i], 1); mbar_init(&bars->g_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == kMma1Warp) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == kProducerWarp) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            FlCursor c;
            c.init(p, u_begin);
            int k = 0;
            for (int it = 0; it < my_units; ++it, c.next(p)) {
                const FlPass& ps = p.pass[c.pass];
                if (it == 0 || c.tile == 0) {   // a new pair of panels
                    mbar_wait(&bars->r_empty, (k & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->r_full, Cfg::kRBytes);
                    for (int g = 0; g < 2; ++g)
                        for (int s = 0; s < Cfg::kSlabs; ++s)
                            tma_load_2d(sR + g * Cfg::kPanelBytes + s * 128 * 128, &maps.r[c.pass], &bars->r_full, s * 64, (c.pair * 2 + g) * 128);
                    ++k;
                }
                const int stage = it % Cfg::kStages;
                const int n0 = c.tile * BN;
                mbar_wait(&bars->t_empty[stage], ((it / Cfg::kStages) & 1) ^ 1);
                mbar_arrive_expect_tx(&bars->t_full[stage], Cfg::kTBytes + Cfg::kC2Bytes);
                unsigned char* dst = sT + stage * Cfg::kTBytes;
                for (int s = 0; s < Cfg::kSlabs; ++s) tma_load_2d(dst + s * BN * 128, &maps.t[c.pass], &bars->t_full[stage], s * 64, n0);
                bulk_copy_1d(sC2 + stage * Cfg::kC2Bytes, ps.colv2 + n0, Cfg::kC2Bytes, &bars->t_full[stage]);
                FL_TRACE(it, 0);
            }
        }
    } else if (warp == kMma1Warp) {
        // ===================== first product, per panel and half: S_g[:, half] = R_g . T[half]^T =====================
        constexpr uint32_t idesc1 = make_idesc_f16(128, kHN);
        const uint32_t sR_a = smem_u32(sR), sT_a = smem_u32(sT);
        FlCursor c;
        c.init(p, u_begin);
        int k = -1;
        for (int it = 0; it < my_units; ++it, c.next(p)) {
            const bool seg_start = (it == 0 || c.tile == 0);
            const bool seg_end = (it == my_units - 1 || c.tile == p.pass[c.pass].n_tiles - 1);
            if (seg_start) {
                ++k;
                mbar_wait(&bars->r_full, k & 1);
            }
            const int stage = it % Cfg::kStages;
            mbar_wait(&bars->t_full[stage], (it / Cfg::kStages) & 1);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(&bars->s_empty[g][h], (it & 1) ^ 1);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t ad0 = make_smem_desc(sR_a + g * Cfg::kPanelBytes, 16, 1024);
                        const uint64_t bd0 = make_smem_desc(sT_a + stage * Cfg::kTBytes + h * kHN * 128, 16, 1024);   // rows [h*BN/2, +BN/2) of every slab
#pragma unroll
                        for (int kk = 0; kk < Cfg::kMma1; ++kk) {
                            const uint64_t ad = ad0 + (uint64_t)(((kk >> 2) * 128 * 128 + (kk & 3) * 32) >> 4);
                            const uint64_t bd = bd0 + (uint64_t)(((kk >> 2) * BN * 128 + (kk & 3) * 32) >> 4);
                            mma_f16(tmem + g * 256 + Cfg::kSCol + h * kHN, ad, bd, idesc1, kk > 0 ? 1u : 0u);
                        }
                        mma_commit(&bars->s_full[g][h]);
                        if (g == 1 && h == 1 && seg_end) mma_commit(&bars->r_empty);
                        if (g == 0 && h == 0) FL_TRACE(it, 1);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == kMma2Warp) {
        // ===================== second product, per panel and half: G_g += P_g[:, half] . T[half]  (A = P from tensor memory,
        // B = the same tile, MN-major) =====================
        constexpr uint32_t idesc2 = make_idesc_f16_bmn(128, E);
        const uint32_t sT_a = smem_u32(sT);
        const uint32_t lbo = p.mn_lbo ? (uint32_t)p.mn_lbo : (uint32_t)(BN * 128), sbo = p.mn_sbo ? (uint32_t)p.mn_sbo : 1024u;
        FlCursor c;
        c.init(p, u_begin);
        int k = -1;
        for (int it = 0; it < my_units; ++it, c.next(p)) {
            const bool seg_start = (it == 0 || c.tile == 0);
            const bool seg_end = (it == my_units - 1 || c.tile == p.pass[c.pass].n_tiles - 1);
            if (seg_start) ++k;
            const int stage = it % Cfg::kStages;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                if (seg_start) mbar_wait(&bars->g_empty[g], (k & 1) ^ 1);   // the epilogue has drained the previous segment's G
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    mbar_wait(&bars->p_full[g][h], it & 1);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t bd0 = make_smem_desc(sT_a + stage * Cfg::kTBytes, lbo, sbo);
                        const uint32_t d_t = tmem + g * 256 + Cfg::kGCol, a_t = tmem + g * 256 + Cfg::kPCol;
#pragma unroll
                        for (int k2 = 0; k2 < Cfg::kMma2 / 2; ++k2) {   // 16 rows of T (2048 bytes of every slab) per instruction
                            const int kk = h * (Cfg::kMma2 / 2) + k2;
                            mma_f16_ts(d_t, a_t + kk * 8, bd0 + (uint64_t)((kk * 2048) >> 4), idesc2, (!seg_start || kk > 0) ? 1u : 0u);
                        }
                        mma_commit(&bars->p_empty[g][h]);
                        if (g == 1 && h == 1) mma_commit(&bars->t_empty[stage]);   // both panels are done with this tile (and its column term)
                        if (h == 1 && seg_end) mma_commit(&bars->g_full[g]);
                        if (g == 0 && h == 1) FL_TRACE(it, 7);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ===================== epilogue warpgroups: warps 0-3 -> panel 0 of the pair, warps 4-7 -> panel 1 =====================
        const int g = warp >> 2, q = warp & 3;
        const int row_l = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t tS = tmem + lane_addr + g * 256 + Cfg::kSCol, tP = tmem + lane_addr + g * 256 + Cfg::kPCol,
                       tG = tmem + lane_addr + g * 256 + Cfg::kGCol;
        FlCursor c;
        c.init(p, u_begin);
        int k = -1;
        int row = 0, slot = 0, rows_pad = 0;
        bool row_ok = false;
        float kmul = 0.f;
        // pass 1 state: x = a_run - zn;  a_run = +inf until the first chunk
        float a_run = CUDART_INF_F, zd = 0.f;
        f32x2 lsum = pk2(0.f, 0.f);
        bool has_diag = false;
        float rowc = 0.f;   // pass 2: kOff2 - rowv*log2e
        for (int it = 0; it < my_units; ++it, c.next(p)) {
            const FlPass& ps = p.pass[c.pass];
            const bool seg_start = (it == 0 || c.tile == 0);
            const bool seg_end = (it == my_units - 1 || c.tile == ps.n_tiles - 1);
            if (seg_start) {
                ++k;
                row = (c.pair * 2 + g) * 128 + row_l;
                row_ok = row < ps.nR;
                rows_pad = ps.m_pairs * 256;
                slot = blockIdx.x - sk_owner(ps.unit0 + c.pair * ps.n_tiles, p.units, gridDim.x);
                kmul = __ldg(p.kmul + c.pass);
                a_run = CUDART_INF_F; zd = 0.f; has_diag = false; lsum = pk2(0.f, 0.f);
                if (MODE == kP2) rowc = kOff2 - ((row_ok && ps.rowv) ? __ldg(ps.rowv + row) * kLog2e : 0.f);
            }
            const int wrow0 = (c.pair * 2 + g) * 128 + q * 32;
            const int stage = it % Cfg::kStages;
            const int n0 = c.tile * BN;
            const bool fast = (n0 + BN <= ps.nT) && (wrow0 + 32 <= ps.nR) && (wrow0 + ps.d + 32 <= n0 || wrow0 + ps.d >= n0 + BN);
            const int dcol_abs = row_ok ? row + ps.d : -1;
            const uint32_t ph = it & 1;
            mbar_wait(&bars->t_full[stage], (it / Cfg::kStages) & 1);   // the staged column term (landed long ago)
            mbar_wait(&bars->s_full[g][0], ph);
            tc_fence_after();
            if (lane == 0 && q == 0 && g == 0) FL_TRACE(it, 2);
            const uint32_t c2s = smem_u32(sC2 + stage * Cfg::kC2Bytes);
            uint32_t rbuf[2][32];
            tmem_ld_32x32_issue(tS, rbuf[0]);
#pragma unroll
            for (int cc = 0; cc < Cfg::kChunks; ++cc) {
                const int h = cc / kHC;
                const bool first_of_half = (cc % kHC) == 0, last_of_half = (cc % kHC) == kHC - 1;
                tmem_ld_wait();
                if (cc + 1 < Cfg::kChunks) {
                    if ((cc + 1) % kHC == 0) {   // the next chunk opens the second half of S
                        mbar_wait(&bars->s_full[g][1], ph);
                        tc_fence_after();
                    }
                    tmem_ld_32x32_issue(tS + (cc + 1) * 32, rbuf[(cc + 1) & 1]);
                }
                if (last_of_half) {   // every column of this half of S is in registers: the next tile's first product may refill it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->s_empty[g][h]);
                }
                if (first_of_half) {  // the second product of the previous tile has consumed this half of P
                    mbar_wait(&bars->p_empty[g][h], ph ^ 1);
                    tc_fence_after();
                }
                uint32_t(&r)[32] = rbuf[cc & 1];
                const int nb = n0 + cc * 32;
                if constexpr (MODE == kP1) {
                    // (the fast and the checked form are separate branches end to end: joining them after phase A costs a register
                    // move per logit)
                    auto raise = [&](bool need, float cmin) {
                        // G must be quiescent: every second product issued so far has to be complete (the previous tile's second half;
                        // this tile's first half when we are in the second)
                        mbar_wait(&bars->p_empty[g][1], ph ^ 1);
                        if (h == 1) mbar_wait(&bars->p_empty[g][0], ph);
                        tc_fence_after();
                        const P1State ns = p1_raise<E>(need, cmin, !seg_start || h == 1, cc % kHC, tG, tP + h * kHC * 16, a_run, lsum);
                        a_run = ns.a; lsum = ns.l;
                    };
                    if (fast) {
                        f32x2 zn[16];
                        uint32_t w[16];
                        const float cmin = p1_zn_fast(r, c2s + cc * 128, kmul, zn);
                        const bool need = cmin < a_run - (kOff1 + kTau);   // this chunk exceeds the row's reference by more than 2^kTau
                        if (__any_sync(0xffffffffu, need)) raise(need, cmin);
                        p1_exp_fast(zn, a_run, lsum, w);
                        tmem_st_32x16(tP + cc * 16, w);
                    } else {
                        f32x2 zn[16];
                        uint32_t w[16];
                        int dloc = -1;
                        const float cmin = p1_zn_checked(r, c2s + cc * 128, kmul, nb, ps.nT, dcol_abs, zn, dloc);
                        const bool need = cmin < a_run - (kOff1 + kTau);
                        if (__any_sync(0xffffffffu, need)) raise(need, cmin);
                        p1_exp_checked(zn, a_run, dloc, lsum, w, zd);
                        if (dloc >= 0) has_diag = true;
                        tmem_st_32x16(tP + cc * 16, w);
                    }
                } else {
                    uint32_t w[16];
                    if (fast) p2_chunk_fast(r, c2s + cc * 128, kmul, rowc, w);
                    else p2_chunk_checked(r, c2s + cc * 128, kmul, rowc, nb, ps.nT, row_ok, dcol_abs, ps.diag_pm1, w);
                    tmem_st_32x16(tP + cc * 16, w);
                }
                if (last_of_half) {   // this half of P is complete in tensor memory
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->p_full[g][h]);
                }
            }
            if (lane == 0 && q == 0 && g == 0) FL_TRACE(it, 3);
            if (seg_end) {
                mbar_wait(&bars->g_full[g], k & 1);
                tc_fence_after();
#pragma unroll
                for (int cg = 0; cg < E / 32; ++cg) {
                    float v[32];
                    tmem_ld_32x32(tG + cg * 32, v);
                    if (row_ok) {
                        float4* dst = reinterpret_cast<float4*>(ps.out_g + ((int64_t)slot * rows_pad + row) * E + cg * 32);
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) dst[g4] = make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->g_empty[g]);
                if (MODE == kP1 && row_ok) {
                    float l0, l1;
                    upk2(lsum, l0, l1);
                    ps.out_m[(int64_t)slot * rows_pad + row] = -a_run;
                    ps.out_l[(int64_t)slot * rows_pad + row] = l0 + l1;
                    if (has_diag) ps.out_zd[row] = zd;
                }
            }
        }
    }
    __syncthreads();
    if (warp == kMma1Warp) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace tc
}  // namespace tt
