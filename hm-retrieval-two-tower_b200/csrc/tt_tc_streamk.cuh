// tt_tc_streamk.cuh -- persistent ("stream-K") form of the logits-shaped softmax kernels.
//
// The in-batch softmax is a grid of work units, one per (128-row panel of the resident operand R, BN-row tile of the
// streamed operand T):  S(128 x BN) = R_panel . T_tile^T on tcgen05 (kind::tf32, fp32 accumulators in TMEM), consumed by
// the epilogue warps straight from TMEM.  rowpanel_kernel (tt_tc_rowpanel.cuh) launches one CTA per (panel, column
// split); with 64 panels that fills 128 of the 148 SMs and the backward runs as two such launches.  Here ONE launch of
// (at most) one CTA per SM walks a contiguous range of the unit list -- all panels of the dQ pass followed by all panels
// of the dC pass -- so every SM gets the same number of units (+-1).  A CTA whose range crosses a panel boundary
// switches panels in flight: the producer reloads R once the last MMA that reads it has completed, the MMA warp flips
// to the second G accumulator, the epilogue warps flush the finished panel (per-row log-sum-exp partials forward, the
// 128 x E gradient block backward) into a per-(panel, CTA slot) partial buffer.  A panel is covered by consecutive CTAs;
// its partials are merged in slot order by a small kernel (fixed order: deterministic, no atomics).
//
//   kFwd  z = s*log2e - colv2_j; online max / sum 2^(z - max) per row; diagonal logit          (tt_inbatch_softmax_fwd)
//   kBwd  P = 2^(z - rowv_i*log2e) - [j == i + d] -> fp16 pairs -> tensor memory (tcgen05.st) = A operand of the 2nd MMA (kind::f16)
//         G(128 x E) += P . T_tile; the B operand is a K-major fp16 tile of T^T (TMA from a transposed fp16 copy)
//                                                                                                (tt_inbatch_softmax_bwd)
#pragma once
#include "tt_tc_rowpanel.cuh"

#define SK_TRACE(it, ev)                                                                                         \
    do {                                                                                                     \
        if (p.trace && (it) < 64) p.trace[((size_t)blockIdx.x * 64 + (it)) * 8 + (ev)] = gtime();             \
    } while (0)

namespace tt {
namespace tc {

struct SkPass {
    int nR, nT;          // rows of the resident operand / of the streamed operand
    int m_tiles, n_tiles;
    int d;               // diagonal: column == row + d
    int unit0;           // first work unit of this pass
    const float* rowv;   // kBwd: per-R-row term (lse or ln p), natural units; may be null (zero)
    const float* colv2;  // per-T-row term * log2(e), zero padded to n_tiles*BN entries (never null)
    float* out0;         // kFwd: m2 [slot][half][nR] | kBwd: G partial [slot][nR][E]
    float* out1;         // kFwd: l  [slot][half][nR]
    float* out2;         // kFwd: zdiag [nR] (natural units)
};
struct SkParams {
    SkPass pass[2];
    int n_pass;
    int units;           // work units of all passes
    unsigned long long* trace;   // optional debug timeline [cta][64 units][8]: TMA issued, MMA1 issued, S seen by the epilogue, S released, MMA warp: T landed, S buffer free, P ready, MMA2 issued
};
struct SkMaps {
    CUtensorMap r[2], t[2], tt[2];   // per pass: R panels (box 128 rows), T tiles (box BN rows), T^T tiles (box E rows)
};

// CTA b of G owns units [sk_begin(b), sk_begin(b + 1)); sk_owner inverts it
__host__ __device__ inline int sk_begin(int b, int units, int G) { return (int)(((long long)b * units) / G); }
__host__ __device__ inline int sk_owner(int u, int units, int G) { return (int)((((long long)u + 1) * G + units - 1) / units) - 1; }

struct SkCursor {   // position in the unit list
    int pass, panel, tile;
    __device__ __forceinline__ void init(const SkParams& p, int u) {
        pass = (p.n_pass > 1 && u >= p.pass[1].unit0) ? 1 : 0;
        const int local = u - p.pass[pass].unit0;
        panel = local / p.pass[pass].n_tiles;
        tile = local - panel * p.pass[pass].n_tiles;
    }
    __device__ __forceinline__ void next(const SkParams& p) {
        if (++tile == p.pass[pass].n_tiles) {
            tile = 0;
            if (++panel == p.pass[pass].m_tiles) { panel = 0; ++pass; }
        }
    }
};

// H: the first MMA's operands are fp16 (kind::f16) instead of TF32-rounded fp32 (kind::tf32).  A TF32-rounded value inside
// fp16's normal range [2^-14, 65504] converts exactly, so the logits are the same; the operand tiles are half the size, and the
// kernels are bound by streaming those tiles from L2 (E >= 64 only: a 128-byte swizzle span holds 64 fp16).
template <int MODE, int E, int BN, bool H>
struct SkCfg {
    static constexpr int kSlabCols = H ? 64 : 32;                       // columns per 128-byte K slab of an operand row
    static constexpr int kSlabs = E / kSlabCols;
    static constexpr int kRBytes = kSlabs * 128 * 128;                  // R panel (K-major over E)
    static constexpr int kT1Bytes = kSlabs * BN * 128;                  // T tile (K-major over E): first MMA
    static constexpr int kMma1 = H ? E / 16 : E / 8;                    // first-MMA instructions per tile (32 bytes of K each)
    static_assert(!H || E >= 64, "fp16 operands need E >= 64");
    static constexpr int kT2Bytes = (MODE == kBwd) ? (BN / 64) * E * 128 : 0;     // T^T tile (fp16, K-major over BN): second MMA
    static constexpr int kTBytes = kT1Bytes + kT2Bytes;
    static constexpr int kPBytes = 0;                                   // P lives in TMEM (A operand of the second MMA)
    static constexpr int kPBufs = 0;
    static constexpr int kPCols = BN / 2;                               // TMEM columns of one P buffer: fp16 pairs packed per 32-bit column
    static constexpr int kC2Bytes = BN * 4;
    static constexpr int kFixed = kRBytes + kPBufs * kPBytes + 4 * 1024 /*c2 stages*/ + 1024 /*barriers*/ + 1024 /*align*/;
    static constexpr int kFit = (232448 - kFixed) / kTBytes;
    static constexpr int kStages = kFit >= 8 ? 8 : kFit;                // T ring depth (TMA latency under load is several unit times)
    static constexpr int kC2Slots = 4096 / kC2Bytes >= 8 ? 8 : 4096 / kC2Bytes;   // ring of staged column terms (its own barriers)
    static constexpr int kSmemBytes = kFixed + kStages * kTBytes;
    // epilogue warps per TMEM lane quarter: the forward runs 4 (one 32-column chunk each per tile: 4 warps per scheduler hide the
    // dependent FMNMX / MUFU / FADD latencies better than 2 warps with two chunks each)
    static constexpr int kHalves = (BN / 32 >= 4) ? 4 : ((BN / 32 >= 2) ? 2 : 1);
    static constexpr int kEpiWarps = 4 * kHalves;
    static constexpr int kThreads = 32 * kEpiWarps + 64 + (MODE == kBwd ? 32 : 0);   // epilogue warps, producer, MMA issuer(s)
    // S accumulators in flight: the MMA -> epilogue -> MMA hand-over latency is amortised over kAcc units
    static constexpr int kAcc = (MODE == kFwd) ? (4 * BN <= 512 ? 4 : 2) : 2;
    // backward TMEM map: S[kAcc] | P[2] (fp16 pairs) | G[2]
    static constexpr int kPBase = kAcc * BN, kGBase = kAcc * BN + 2 * (BN / 2);
    static constexpr int kTmemNeed = (MODE == kBwd) ? kGBase + 2 * E : kAcc * BN;
    static constexpr int kTmemCols = kTmemNeed <= 128 ? 128 : (kTmemNeed <= 256 ? 256 : 512);
    static_assert(kStages >= 2, "shared memory budget");
    static_assert(kTmemNeed <= 512, "TMEM budget");
    static_assert(MODE == kFwd || MODE == kBwd, "stream-K kernel: forward or backward");
    static_assert(MODE != kBwd || BN % 64 == 0, "backward: BN must be a multiple of 64 (fp16 slabs of 64 columns)");
    static_assert(E == 32 || E == 64 || E == 128, "E must be 32, 64 or 128");
};

// D[tmem] (+)= A[tmem] . B[smem]: the A operand (M x K, fp16 pairs packed per 32-bit column, row m in lane m) is read from
// tensor memory -- the epilogue writes P there with tcgen05.st, so P never touches shared memory
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        :
        : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp writes lane (lane_base + t)
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&w)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]),
        "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {   // round-to-nearest-even, lo in the low half
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2: two lanes of the FMA pipe per issue slot) --------------------------
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// forward epilogue body for a chunk that is fully in range and off the diagonal: works on the NEGATED logits
// zn = colv2 - s*log2e (one FFMA2 per two columns), running minimum with FMNMX3, 2^(min - zn) with one FFMA2 + two MUFU,
// packed accumulation.  ~115 instructions per 32 columns against ~200 for the scalar form.
__device__ __forceinline__ void fwd_chunk_packed(const uint32_t (&r)[32], uint32_t c2s, float& m2, float& l) {
    f32x2 zn[16];
    const f32x2 nl2e = pk2(-kLog2e, -kLog2e);
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        zn[2 * g4] = fma2(pk2(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1])), nl2e, pk2(cc.x, cc.y));
        zn[2 * g4 + 1] = fma2(pk2(__uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3])), nl2e, pk2(cc.z, cc.w));
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
    const float mneg = fminf(-m2, fminf(mn0, mn1));   // minus the new running maximum
    const f32x2 mone = pk2(-1.f, -1.f), mm = pk2(mneg, mneg);
    f32x2 s0 = pk2(0.f, 0.f), s1 = s0;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        upk2(fma2(zn[i], mone, mm), a0, a1);           // mneg - zn = z - max  (<= 0)
        upk2(fma2(zn[i + 1], mone, mm), b0, b1);
#if defined(TT_EXPERIMENT) && TT_EXPERIMENT == 1
        s0 = add2(s0, pk2(a0, a1));
        s1 = add2(s1, pk2(b0, b1));
#else
        s0 = add2(s0, pk2(ex2_approx(a0), ex2_approx(a1)));
        s1 = add2(s1, pk2(ex2_approx(b0), ex2_approx(b1)));
#endif
    }
    upk2(add2(s0, s1), a0, a1);
    l = l * ex2_approx(m2 + mneg) + (a0 + a1);        // m2 - max_new; m2 = -inf gives 0
    m2 = -mneg;
}

// backward epilogue bodies: 32 columns of one row -> P = 2^(s*log2e - r2 - c2_j) - [j == i + d] as 16 packed fp16 pairs
// (fp16 keeps the 11 significant bits TF32 would; P lies in [-1, 1], values below 6e-5 lose relative precision only).
// Fast path (chunk fully in range, off the diagonal): one FFMA2 per two columns (r2 folded into the addend), two MUFU,
// one F2FP pack.
__device__ __forceinline__ void bwd_chunk_packed(const uint32_t (&r)[32], uint32_t c2s, float r2, uint32_t (&w)[16]) {
    const f32x2 l2e = pk2(kLog2e, kLog2e), mone = pk2(-1.f, -1.f), nr2 = pk2(-r2, -r2);
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        // t = s*log2e - (c2 + r2):  addend = c2*(-1) + (-r2), then fma(s, log2e, addend)
        const f32x2 ad0 = fma2(pk2(cc.x, cc.y), mone, nr2), ad1 = fma2(pk2(cc.z, cc.w), mone, nr2);
        float a0, a1, b0, b1;
        upk2(fma2(pk2(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1])), l2e, ad0), a0, a1);
        upk2(fma2(pk2(__uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3])), l2e, ad1), b0, b1);
        w[2 * g4] = pack_f16x2(ex2_approx(a0), ex2_approx(a1));
        w[2 * g4 + 1] = pack_f16x2(ex2_approx(b0), ex2_approx(b1));
    }
}
__device__ __forceinline__ void bwd_chunk_checked(const uint32_t (&r)[32], uint32_t c2s, int nb, int row, int nT, int nR, int d, float r2,
                                                  uint32_t (&w)[16]) {
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
        float pv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = g4 * 4 + t;
            float v = ex2_approx(fmaf(__uint_as_float(r[i]), kLog2e, -r2) - cv[t]);
            const int n = nb + i;
            if (n >= nT || row >= nR) v = 0.f;
            else if (n == row + d) v -= 1.0f;
            pv[t] = v;
        }
        w[2 * g4] = pack_f16x2(pv[0], pv[1]);
        w[2 * g4 + 1] = pack_f16x2(pv[2], pv[3]);
    }
}

struct SkBars {
    uint64_t r_full, r_empty;
    uint64_t t_full[8], t_empty[8];
    uint64_t c_full[8], c_empty[8];
    uint64_t s_full[4], s_empty[4];
    uint64_t p_full[2], p_empty[2];
    uint64_t g_full[2], g_empty[2];
    uint32_t tmem_base;
};

template <int MODE, int E, int BN, bool H>
__global__ void __launch_bounds__(SkCfg<MODE, E, BN, H>::kThreads, 1)
streamk_kernel(const __grid_constant__ SkMaps maps, const __grid_constant__ SkParams p) {
    using Cfg = SkCfg<MODE, E, BN, H>;
    using Sk = SkCfg<MODE, E, BN, H>;
    const int u_begin = sk_begin(blockIdx.x, p.units, gridDim.x), u_end = sk_begin(blockIdx.x + 1, p.units, gridDim.x);
    const int my_units = u_end - u_begin;
    if (my_units <= 0) return;   // (uniform) more CTAs than work units

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sR = smem;
    unsigned char* sT = sR + Cfg::kRBytes;
    unsigned char* sP = sT + Cfg::kStages * Cfg::kTBytes;
    unsigned char* sC2 = sP;                                    // kC2Slots x kC2Bytes (4 KB)
    SkBars* bars = reinterpret_cast<SkBars*>(sC2 + 4 * 1024);

    // roles: warps 0..kEpiWarps-1 epilogue, then the TMA producer, then the MMA issuer(s).  The backward has TWO issuers (first
    // MMA / second MMA): one lane issuing all 12 MMAs of a unit plus its three barrier waits was the per-unit critical path
    // (~100 cycles per tcgen05.mma issue: descriptor math, R2UR, commit; 1.5 us per unit measured).  The single-lane warps get the
    // HIGHEST warp ids: the scheduler arbitrates highest-id-first, and behind 4 always-eligible epilogue warps per scheduler a
    // low-id MMA warp was issuing one unit per ~0.9 us (measured) although every barrier it waits on had long completed.
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kProducerWarp = Cfg::kEpiWarps, kMmaWarp = Cfg::kEpiWarps + 1, kMma2Warp = Cfg::kEpiWarps + 2;
    if (warp == kProducerWarp && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            if (i < p.n_pass) {
                prefetch_tmap(&maps.r[i]);
                prefetch_tmap(&maps.t[i]);
                if (MODE == kBwd) prefetch_tmap(&maps.tt[i]);
            }
        }
        mbar_init(&bars->r_full, 1);
        mbar_init(&bars->r_empty, 1);
        for (int i = 0; i < 8; ++i) {
            mbar_init(&bars->t_full[i], 1); mbar_init(&bars->t_empty[i], 1);
            mbar_init(&bars->c_full[i], 1); mbar_init(&bars->c_empty[i], Cfg::kEpiWarps);
        }
        for (int i = 0; i < 4; ++i) { mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_empty[i], Cfg::kEpiWarps); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->p_full[i], Cfg::kEpiWarps); mbar_init(&bars->p_empty[i], 1);
            mbar_init(&bars->g_full[i], 1); mbar_init(&bars->g_empty[i], Cfg::kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(&bars->tmem_base, Sk::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == kProducerWarp) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            SkCursor c;
            c.init(p, u_begin);
            int k = 0;   // panels loaded so far
            for (int it = 0; it < my_units; ++it, c.next(p)) {
                const SkPass& ps = p.pass[c.pass];
                if (it == 0 || c.tile == 0) {   // a new panel: R may be overwritten once the last MMA reading it has completed
                    mbar_wait(&bars->r_empty, (k & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->r_full, Cfg::kRBytes);
                    for (int s = 0; s < Cfg::kSlabs; ++s) tma_load_2d(sR + s * 128 * 128, &maps.r[c.pass], &bars->r_full, s * Cfg::kSlabCols, c.panel * 128);
                    ++k;
                }
                const int stage = it % Cfg::kStages;
                const uint32_t ph = (it / Cfg::kStages) & 1;
                const int n0 = c.tile * BN;
                const int cs = it % Cfg::kC2Slots;
                mbar_wait(&bars->c_empty[cs], ((it / Cfg::kC2Slots) & 1) ^ 1);
                mbar_arrive_expect_tx(&bars->c_full[cs], Cfg::kC2Bytes);
                bulk_copy_1d(sC2 + cs * Cfg::kC2Bytes, ps.colv2 + n0, Cfg::kC2Bytes, &bars->c_full[cs]);
                mbar_wait(&bars->t_empty[stage], ph ^ 1);
                mbar_arrive_expect_tx(&bars->t_full[stage], Cfg::kTBytes);
                unsigned char* dst = sT + stage * Cfg::kTBytes;
                for (int s = 0; s < Cfg::kSlabs; ++s) tma_load_2d(dst + s * BN * 128, &maps.t[c.pass], &bars->t_full[stage], s * Cfg::kSlabCols, n0);
                if (MODE == kBwd) {
                    for (int s = 0; s < BN / 64; ++s)   // fp16 T^T: boxes of 64 columns (128 bytes) x E rows
                        tma_load_2d(dst + Cfg::kT1Bytes + s * E * 128, &maps.tt[c.pass], &bars->t_full[stage], n0 + s * 64, 0);
                }
                SK_TRACE(it, 0);
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc1 = H ? make_idesc_f16(128, BN) : make_idesc_tf32(128, BN, false, false);
        const uint32_t sR_a = smem_u32(sR), sT_a = smem_u32(sT);
        SkCursor c1;   // unit whose first MMA is issued next
        c1.init(p, u_begin);
        int k1 = -1;   // panel sequence number of c1's unit
        auto issue_g1 = [&](int it) {
            const bool panel_start = (it == 0 || c1.tile == 0);
            const bool panel_end = (it == my_units - 1 || c1.tile == p.pass[c1.pass].n_tiles - 1);
            if (panel_start) {
                ++k1;
                mbar_wait(&bars->r_full, k1 & 1);
            }
            const int stage = it % Cfg::kStages;
            const uint32_t tph = (it / Cfg::kStages) & 1;
            const int acc = it % Cfg::kAcc;
            const uint32_t aph = (it / Cfg::kAcc) & 1;
            mbar_wait(&bars->t_full[stage], tph);
            if (lane == 0) SK_TRACE(it, 4);
            mbar_wait(&bars->s_empty[acc], aph ^ 1);
            if (lane == 0) SK_TRACE(it, 5);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t ad0 = make_smem_desc(sR_a, 16, 1024), bd0 = make_smem_desc(sT_a + stage * Cfg::kTBytes, 16, 1024);
#pragma unroll
                for (int k = 0; k < Cfg::kMma1; ++k) {   // K steps differ only in the start-address field (units of 16 bytes)
                    const uint64_t ad = ad0 + (uint64_t)(((k >> 2) * 128 * 128 + (k & 3) * 32) >> 4);
                    const uint64_t bd = bd0 + (uint64_t)(((k >> 2) * BN * 128 + (k & 3) * 32) >> 4);
                    if (H) mma_f16(tmem + acc * BN, ad, bd, idesc1, k > 0 ? 1u : 0u);
                    else mma_tf32(tmem + acc * BN, ad, bd, idesc1, k > 0 ? 1u : 0u);
                }
                if (MODE != kBwd) mma_commit(&bars->t_empty[stage]);   // T tile consumed
                mma_commit(&bars->s_full[acc]);
                if (panel_end) mma_commit(&bars->r_empty);             // the last MMA that reads this R panel
                SK_TRACE(it, 1);
            }
            __syncwarp();
            c1.next(p);
        };
        for (int it = 0; it < my_units; ++it) issue_g1(it);   // (backward: runs ahead of the second-MMA issuer by up to kAcc units)
    } else if (MODE == kBwd && warp == kMma2Warp) {
        // ===================== second-MMA issuer (backward): G += P . T_tile, A = P from tensor memory =====================
        constexpr uint32_t idesc2 = make_idesc_f16(128, E);
        const uint32_t sT_a = smem_u32(sT);
        SkCursor c2;
        c2.init(p, u_begin);
        int k2 = -1;
        for (int it = 0; it < my_units; ++it, c2.next(p)) {
            const bool panel_start = (it == 0 || c2.tile == 0);
            const bool panel_end = (it == my_units - 1 || c2.tile == p.pass[c2.pass].n_tiles - 1);
            if (panel_start) {
                ++k2;
                mbar_wait(&bars->g_empty[k2 & 1], ((k2 >> 1) & 1) ^ 1);   // the epilogue has drained this G buffer
            }
            const int gb = k2 & 1;
            const int stage = it % Cfg::kStages;
            const int pb = it & 1;
            mbar_wait(&bars->t_full[stage], (it / Cfg::kStages) & 1);     // T^T tile landed (long ago: the first MMA of this unit is done)
            mbar_wait(&bars->p_full[pb], (it >> 1) & 1);
            if (lane == 0) SK_TRACE(it, 6);
            tc_fence_after();
            if (lane == 0) {
                // descriptor of the first K step; the others differ only in the 14-bit start-address field (units of 16 bytes)
                const uint64_t bd0 = make_smem_desc(sT_a + stage * Cfg::kTBytes + Cfg::kT1Bytes, 16, 1024);
                const uint32_t d_t = tmem + Cfg::kGBase + gb * E, a_t = tmem + Cfg::kPBase + pb * Cfg::kPCols;
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk)   // fp16: 16 columns (32 bytes) per MMA, 64 per swizzle slab
                    mma_f16_ts(d_t, a_t + kk * 8, bd0 + (uint64_t)(((kk >> 2) * E * 128 + (kk & 3) * 32) >> 4), idesc2, (!panel_start || kk > 0) ? 1u : 0u);
                mma_commit(&bars->t_empty[stage]);  // T and T^T of this stage are consumed by now (the first MMA completed before P existed)
                mma_commit(&bars->p_empty[pb]);
                if (panel_end) mma_commit(&bars->g_full[gb]);
                SK_TRACE(it, 7);
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps (0 .. kEpiWarps-1) =====================
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = warp >> 2;                   // which share of each tile's columns this warp handles
        const int row_l = q * 32 + lane;              // row within the panel == TMEM lane
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        constexpr int NC = BN / 32;
        constexpr int NCW = NC / Cfg::kHalves;        // 32-column chunks per warp per tile
        const int c_first = half * NCW;
        SkCursor c;
        c.init(p, u_begin);
        int k = -1;                                   // panel sequence number
        int row = 0, slot = 0;
        float m2 = -CUDART_INF_F, l = 0.f, zd = 0.f, r2 = 0.f;
        bool has_diag = false;
        for (int it = 0; it < my_units; ++it, c.next(p)) {
            const SkPass& ps = p.pass[c.pass];
            const bool panel_start = (it == 0 || c.tile == 0);
            const bool panel_end = (it == my_units - 1 || c.tile == ps.n_tiles - 1);
            if (panel_start) {
                ++k;
                row = c.panel * 128 + row_l;
                slot = blockIdx.x - sk_owner(ps.unit0 + c.panel * ps.n_tiles, p.units, gridDim.x);
                m2 = -CUDART_INF_F; l = 0.f; zd = 0.f; has_diag = false;
                if (MODE == kBwd) r2 = (row < ps.nR && ps.rowv) ? __ldg(ps.rowv + row) * kLog2e : 0.f;
            }
            const int wrow0 = c.panel * 128 + q * 32;     // first row of this warp
            const int acc = it % Cfg::kAcc;
            const uint32_t aph = (it / Cfg::kAcc) & 1;
            const int stage = it % Cfg::kStages;
            const int n0 = c.tile * BN;
            const int pb = it & 1;
            // warp-uniform: tile fully in range and no diagonal element of this warp's rows inside it
            const bool fast = (n0 + BN <= ps.nT) && (wrow0 + 32 <= ps.nR) && (wrow0 + ps.d + 32 <= n0 || wrow0 + ps.d >= n0 + BN);
            mbar_wait(&bars->s_full[acc], aph);
            tc_fence_after();
            if (threadIdx.x == 0) SK_TRACE(it, 2);
            const int cs = it % Cfg::kC2Slots;
            mbar_wait(&bars->c_full[cs], (it / Cfg::kC2Slots) & 1);     // the staged column term (landed long ago)
            if (MODE == kBwd) mbar_wait(&bars->p_empty[pb], ((it >> 1) & 1) ^ 1);
            const uint32_t c2s = smem_u32(sC2 + cs * Cfg::kC2Bytes);
            uint32_t rbuf[2][32];
            tmem_ld_32x32_issue(tmem + lane_addr + acc * BN + c_first * 32, rbuf[0]);
#pragma unroll
            for (int cl = 0; cl < NCW; ++cl) {
                const int cc = c_first + cl;
                tmem_ld_wait();
                if (cl + 1 < NCW) tmem_ld_32x32_issue(tmem + lane_addr + acc * BN + (cc + 1) * 32, rbuf[(cl + 1) & 1]);
                uint32_t(&r)[32] = rbuf[cl & 1];
                const int nb = n0 + cc * 32;
                if constexpr (MODE == kFwd) {
#if defined(TT_EXPERIMENT) && TT_EXPERIMENT == 2
                    if (fast) { l += __uint_as_float(r[0]) + __uint_as_float(r[31]); }
#else
                    if (fast) fwd_chunk_packed(r, c2s + cc * 128, m2, l);
#endif
                    else fwd_chunk<false>(r, c2s + cc * 128, nb, row, ps.nT, ps.d, m2, l, zd, has_diag);
                } else {
                    uint32_t w[16];
                    if (fast) bwd_chunk_packed(r, c2s + cc * 128, r2, w);
                    else bwd_chunk_checked(r, c2s + cc * 128, nb, row, ps.nT, ps.nR, ps.d, r2, w);
                    tmem_st_32x16(tmem + lane_addr + Cfg::kPBase + pb * Cfg::kPCols + cc * 16, w);
                }
            }
            // this S buffer may be overwritten by the MMA of a later unit; backward: P is complete in tensor memory
            if (MODE == kBwd) tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bars->s_empty[acc]);
                mbar_arrive(&bars->c_empty[cs]);                        // staged column term consumed
            }
            if (threadIdx.x == 0) SK_TRACE(it, 3);
            if constexpr (MODE == kBwd) {
                if (lane == 0) mbar_arrive(&bars->p_full[pb]);   // (the P stores were completed and fenced above)
            }
            if (panel_end) {
                if constexpr (MODE == kFwd) {
                    if (row < ps.nR) {
                        // one partial per (CTA slot of the panel, warp half): the combine kernel merges them in slot order
                        ps.out0[((int64_t)slot * Cfg::kHalves + half) * ps.nR + row] = m2;
                        ps.out1[((int64_t)slot * Cfg::kHalves + half) * ps.nR + row] = l;
                        if (has_diag) ps.out2[row] = zd;
                    }
                } else {
                    const int gb = k & 1;
                    mbar_wait(&bars->g_full[gb], (k >> 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int cg = half; cg < E / 32; cg += Cfg::kHalves) {
                        float v[32];
                        tmem_ld_32x32(tmem + lane_addr + Cfg::kGBase + gb * E + cg * 32, v);
                        if (row < ps.nR) {
                            float4* dst = reinterpret_cast<float4*>(ps.out0 + ((int64_t)slot * ps.nR + row) * E + cg * 32);
#pragma unroll
                            for (int g4 = 0; g4 < 8; ++g4) dst[g4] = make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
                        }
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->g_empty[gb]);
                }
            }
        }
    }
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem, Sk::kTmemCols);
    }
}

}  // namespace tc
}  // namespace tt
