// tt_tc_flash.cuh -- the in-batch sampled softmax as two persistent tcgen05 passes over the (row panel, column tile) grid.
//
//   pass 1 (kP1)  forward AND dQ in one sweep:   S = Q_panel . C_tile^T  ->  online softmax (lazy rescale)  ->  P~ (fp16, in TMEM)
//                 G(128 x E) += P~ . C_tile;  per row: reference exponent, sum of P~, diagonal logit.  The combine kernel turns
//                 (G, sum) into lse, loss and dQ = G / sum + (p_ii - 1) c_ii.
//   pass 2 (kP2)  dC (or any side given lse):    S = C_panel . Q_tile^T  ->  P = 2^(z - lse) - [diagonal]  ->  G += P . Q_tile
//
// Two exponentials and four tile products per logit for forward + backward (the three-sweep form -- forward, dQ pass, dC pass --
// needs three and five).  One CTA per SM walks an equal share of the unit list (stream-K); a unit is a PAIR of 128-row panels
// times one BN-column tile, so the streamed tile is loaded once and used by eight MMAs.
//
// A unit is worked on as NS = 2 * kSplit independent STREAMS (panel of the pair, 64-column half of the tile).  A stream owns
// 64 + E columns of tensor memory -- TWO 32-column S buffers (its half of a tile is two sub-tiles), with the fp16 P written over
// the S columns that are already in registers, and its OWN accumulator G -- four epilogue warps (thread = one row = one TMEM
// lane), its own MMA issuer warp and a private online-softmax state.  Streams never talk to each other: the partial (reference
// exponent, sum, G) of every stream is merged by the combine kernel exactly like the partials of two CTAs that share a panel.
// Per sub-tile:  first product (S)  ->  epilogue (exponentials, P)  ->  second product (G += P . T); with two S buffers the first
// product of sub-tile v+1 is complete before the epilogue has finished v, so the 16 epilogue warps (four per scheduler) never
// wait for the tensor pipe and the MUFU pipe is the only limit.
//
// Operands are fp16 copies scaled by a per-tensor power of two (tt_softmax_flash.cu: amax -> scale), so any finite fp32 input is in
// range; products are exact and accumulate in fp32 in TMEM.  The second product reads P from TENSOR MEMORY (A operand) and the
// SAME shared-memory tile as the first (as an MN-major B operand): no transposed copies exist anywhere.
//
// Pass 1 keeps a per-row reference exponent that is only raised when a chunk of 32 logits exceeds it by more than 2^kTau (then the
// row's running sum, its G row and the P chunk already written for this tile are rescaled -- rare after the first tile of a
// panel; G is quiescent whenever the epilogue runs, because the stream's next product is only issued after the epilogue).
// Every mbarrier wait is bounded.
#pragma once
#include <cstddef>
#include <type_traits>

#include "tt_tc_streamk.cuh"

namespace tt {
namespace tc {

enum FlashMode { kP1 = 0, kP2 = 1 };

constexpr float kTau = 8.f;      // pass 1: a row's reference exponent may lag its true maximum by up to 2^kTau
constexpr float kOff1 = 6.f;     // pass 1: P~ = 2^(z - ref + kOff1) <= 2^(kTau + kOff1) = 2^14 < 65504
#ifndef TT_FLASH_SBUF
#define TT_FLASH_SBUF 1   // S buffers per stream: 1 = one 64-column buffer (N = 64 products), 2 = two 32-column buffers
#endif
#ifndef TT_FLASH_POLY_PAIRS
#define TT_FLASH_POLY_PAIRS 0   // measured: any share on the polynomial is slower -- the passes are issue-bound, not MUFU-bound
#endif
constexpr float kOff2 = 14.f;    // pass 2: P' = 2^(z - lse + kOff2) <= 2^14; fp16 normals then reach down to p = 2^-28

struct FlPass {
    int nR, nT;
    int m_pairs, n_tiles;   // pairs of 128-row panels of R; BN-row tiles of T
    int d;                  // diagonal: column == row + d
    int unit0;
    const float* rowv;      // kP2: per-R-row term, natural units (lse or ln p); may be null
    const float* colv2;     // per-T-row term * log2(e), zero padded to n_tiles*BN entries
    float* out_g;           // G partials [slot * kSplit + half][E / 4][m_pairs*256] as float4 (the flush writes consecutive rows per float4 column)
    float* out_m;           // kP1: reference exponent (log2 units) [slot * kSplit + half][m_pairs*256]
    float* out_l;           // kP1: sum of P~ (positive excluded)    [slot * kSplit + half][m_pairs*256]
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

// position in the unit list with the constants of the current pass in registers (p.pass[] is read only when the pass changes)
struct FlWalk {
    int pass, pair, tile, n_tiles, m_pairs, n_pass;
    __device__ __forceinline__ void init(const FlParams& p, int u) {
        n_pass = p.n_pass;
        pass = (p.n_pass > 1 && u >= p.pass[1].unit0) ? 1 : 0;
        n_tiles = p.pass[pass].n_tiles;
        m_pairs = p.pass[pass].m_pairs;
        const int local = u - p.pass[pass].unit0;
        pair = local / n_tiles;
        tile = local - pair * n_tiles;
    }
    __device__ __forceinline__ bool last_tile() const { return tile == n_tiles - 1; }
    __device__ __forceinline__ void next(const FlParams& p) {
        if (++tile == n_tiles) {
            tile = 0;
            if (++pair == m_pairs) {
                pair = 0;
                if (++pass < n_pass) { n_tiles = p.pass[pass].n_tiles; m_pairs = p.pass[pass].m_pairs; }
            }
        }
    }
};

template <int E>
struct FlCfg {
    static_assert(E == 64 || E == 128, "flash softmax: E must be 64 or 128 (fp16 slabs of 64 columns)");
    static constexpr int BN = (E == 64) ? 128 : 64;               // tile width
    static constexpr int kSplit = BN / 64;                        // 64-column halves per tile
    static constexpr int NS = 2 * kSplit;                         // streams: (panel of the pair, half)
    static constexpr int kEpiWarps = 4 * NS;
    static constexpr int kSlabs = E / 64;
    static constexpr int kPanelBytes = kSlabs * 128 * 128;        // one R panel
    static constexpr int kRBytes = 2 * kPanelBytes;               // the pair
    static constexpr int kTBytes = kSlabs * BN * 128;             // one T tile (K-major over E for MMA1 == MN-major over E for MMA2)
    static constexpr int kC2Bytes = BN * 4;
    static constexpr int kFixed = 2 * kRBytes /*two pairs: the next one is prefetched*/ + 8 * kC2Bytes + 1024 /*barriers*/ + 1024 /*align*/;
    static constexpr int kFit = (232448 - kFixed) / kTBytes;
    static constexpr int kStages = kFit >= 8 ? 8 : kFit;
    static constexpr int kSmemBytes = kFixed + kStages * kTBytes;
    static constexpr int kMma1 = E / 16;                          // K = 16 per instruction
    static constexpr int kStreamCols = 64 + E;                    // TMEM columns of a stream: S (64; P over its first 32) | G (E)
    static constexpr int kGCol = 64;
    static_assert(NS * kStreamCols <= 512, "TMEM budget");
    static constexpr int kThreads = 32 * (kEpiWarps + NS);        // + one issuer warp per stream (MMA, and its share of the TMA loads)
    static constexpr int kLook = kStages - 2;                     // tiles prefetched ahead of the one being multiplied
    static_assert(kStages >= 3, "shared memory budget");
};

struct FlBars {
    uint64_t r_full[2], r_empty[2];
    uint64_t t_full[8], t_empty[8];
    uint64_t s_full[4][2];  // [stream][S buffer] first product of the sub-tile complete
    uint64_t p_full[4][2];  // P of the sub-tile complete in tensor memory (4 epilogue warps)
    uint64_t p_empty[4][2]; // second product complete: the buffer's columns may be refilled
    uint64_t g_full[4], g_empty[4];
    uint32_t tmem_base;
};

// ---- mbarrier by shared-space address: one try_wait inline, the spin + watchdog out of line -----------------------------
__device__ __forceinline__ bool mbar_try_a(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t addr, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_a(addr, parity)) {
        if (++spins > (1u << 24)) {
            printf("libtt: mbarrier watchdog (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, addr, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
    if (!mbar_try_a(addr, parity)) mbar_wait_slow(addr, parity);
}
// one lane of a converged warp (the same lane every time for the same mask)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void mma_commit_a(uint32_t addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}

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
__device__ __forceinline__ void tmem_ld_32x8_issue(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&w)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
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

// A stream's 64 columns are consumed in four CHUNKS of 16 (one tcgen05.ld.x16 each, the next one in flight while the current is
// worked on; P of chunk j -- 8 packed columns -- lands on S columns [8j, 8j+8), which are in registers by then).  16 keeps the
// epilogue at ~90 registers, so 16 epilogue warps fit an SM without spilling.
constexpr int kCW = 16;
constexpr int kSBuf = TT_FLASH_SBUF;          // S buffers (sub-tiles) per stream and tile
constexpr int kSW = 64 / kSBuf;               // columns per sub-tile

// ---- pass 1 state of one row ----------------------------------------------------------------------------------------------
// P~ = 2^(a - zn) with zn = c2_j - s*kmul the negated log2-domain logit; a = +inf until the first chunk.  Passed and returned
// by value: a by-reference state would live in local memory in the hot loop.
struct P1State { float a; f32x2 l; };

// The positive of a row (column dcol_abs) takes no part in maximum, sum or product: its logit is formed from the fp32 operands by
// the combine kernels, which add its term there.  Inside the kernel it is masked on the raw accumulator (s = -inf gives weight 0 in
// both passes); the block below is only entered for the (at most three) chunks per panel that hold positives of this warp's rows.
__device__ __forceinline__ void mask_positive(uint32_t (&r)[16], int nb, int dcol_abs) {
    const int dl = dcol_abs - nb;   // chunk-local column of the positive (outside [0, 16) when it is elsewhere)
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (dl == i) r[i] = 0xff800000u;
}

// pass 1, fast chunk, phase A: zn of 16 columns and their minimum
__device__ __forceinline__ float p1_zn_fast(const uint32_t (&r)[16], uint32_t c2s, float kmul, f32x2 (&zn)[8]) {
    const f32x2 nk = pk2(-kmul, -kmul);
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        zn[2 * g4] = fma2(pk2(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1])), nk, pk2(cc.x, cc.y));
        zn[2 * g4 + 1] = fma2(pk2(__uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3])), nk, pk2(cc.z, cc.w));
    }
    float a0, a1, b0, b1;
    upk2(zn[0], a0, a1);
    upk2(zn[1], b0, b1);
    float mn0 = fminf(a0, a1), mn1 = fminf(b0, b1);
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
        upk2(zn[i], a0, a1);
        upk2(zn[i + 1], b0, b1);
        mn0 = fmin3(mn0, a0, a1);
        mn1 = fmin3(mn1, b0, b1);
    }
    return fminf(mn0, mn1);
}
// 2^x of two values on the FMA pipe instead of the MUFU unit (which is the bottleneck of both passes: one exponential per logit
// against 16 per clock and SM): Cody-Waite split x = n + f with |f| <= 0.5 by magic-number rounding, degree-4 minimax polynomial for
// 2^f (relative error 2.7e-6 -- the result is rounded to fp16 for the second product and summed in fp32), exponent inserted with one
// integer multiply-add.  x is clamped at -125 (2^-125 stands in for anything smaller, including the -inf of masked columns).
__device__ __forceinline__ void ex2_poly2(f32x2 x, float& p0, float& p1) {
    float x0, x1;
    upk2(x, x0, x1);
    const f32x2 xc = pk2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
    const f32x2 t = add2(xc, pk2(12582912.f, 12582912.f));              // 1.5 * 2^23: the integer part lands in the low mantissa bits
    const f32x2 f = add2(xc, fma2(t, pk2(-1.f, -1.f), pk2(12582912.f, 12582912.f)));   // x - n
    f32x2 q = fma2(f, pk2(0.009570039808750153f, 0.009570039808750153f), pk2(0.05591772496700287f, 0.05591772496700287f));
    q = fma2(q, f, pk2(0.240247443318367f, 0.240247443318367f));
    q = fma2(q, f, pk2(0.6931218504905701f, 0.6931218504905701f));
    q = fma2(q, f, pk2(0.9999992847442627f, 0.9999992847442627f));
    float t0, t1, q0, q1;
    upk2(t, t0, t1);
    upk2(q, q0, q1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));   // (bits(1.5 * 2^23) << 23 == 0 mod 2^32)
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}
// of the eight column pairs of a chunk, the first kPolyPairs take the polynomial, the others the MUFU unit
constexpr int kPolyPairs = TT_FLASH_POLY_PAIRS;

// pass 1, fast chunk, phase B: P~ as packed fp16, running sum (two accumulators: halves the dependent FADD2 chain)
__device__ __forceinline__ void p1_exp_fast(const f32x2 (&zn)[8], float a, f32x2& lsum, uint32_t (&w)[8]) {
    const f32x2 mone = pk2(-1.f, -1.f), aa = pk2(a, a);
    f32x2 s1 = pk2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float p0, p1;
        if (i < kPolyPairs) {
            ex2_poly2(fma2(zn[i], mone, aa), p0, p1);
        } else {
            float x0, x1;
            upk2(fma2(zn[i], mone, aa), x0, x1);
            p0 = ex2_approx(x0);
            p1 = ex2_approx(x1);
        }
        if (i & 1) s1 = add2(s1, pk2(p0, p1));
        else lsum = add2(lsum, pk2(p0, p1));
        w[i] = pack_f16x2(p0, p1);
    }
    lsum = add2(lsum, s1);
}

// pass 1, rare path: raise the reference exponent of the rows whose chunk exceeds it by more than 2^kTau; rescale their running
// sums, their G rows (valid once a second product of this segment has completed) and the P chunks of this tile that are already
// written.  G and P are quiescent: the stream's second product for this tile is only issued after the epilogue.
template <int E>
__device__ __noinline__ P1State p1_raise(bool need, float cmin, bool g_valid, int chunks_done, uint32_t tG, uint32_t tP, P1State st) {
    const float a_new = need ? cmin + kOff1 : st.a;
    const float sc = need ? ex2_approx(a_new - st.a) : 1.f;   // a = +inf (first chunk of a segment) -> 0
    st.l = fma2(st.l, pk2(sc, sc), pk2(0.f, 0.f));
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
        uint32_t pw[8];
        tmem_ld_32x8_issue(tP + pc * 8, pw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float lo, hi;
            unpack_f16x2(pw[i], lo, hi);
            pw[i] = pack_f16x2(lo * sc, hi * sc);
        }
        tmem_st_32x8(tP + pc * 8, pw);
    }
    tmem_st_wait();
    st.a = a_new;
    return st;
}

// pass 1, checked chunk (tile edge, rows past the end, or the diagonal inside the chunk), out of line: reads its 16 columns of S
// itself, masks columns >= nT (weight 0), leaves the positive out of the sum and of the product (the combine kernel forms
// p_ii - 1 = -sum_offdiag / sum without cancellation) and records its logit.
template <int E>
__device__ __noinline__ P1State p1_chunk_checked(uint32_t tS_chunk, uint32_t c2s, float kmul, int nb, int nT, int dcol_abs, bool g_valid, int chunks_done,
                                                 uint32_t tG, uint32_t tP, P1State st) {
    uint32_t r[16];
    tmem_ld_32x16_issue(tS_chunk, r);
    tmem_ld_wait();
    float z[16];
    float mn = CUDART_INF_F;
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = g4 * 4 + t;
            float v = fmaf(__uint_as_float(r[i]), -kmul, cv[t]);
            if (nb + i >= nT) v = CUDART_INF_F;
            if (nb + i == dcol_abs) v = CUDART_INF_F;   // the positive takes no part in max, sum or product
            z[i] = v;
            mn = fminf(mn, v);
        }
    }
    const bool need = mn < st.a - (kOff1 + kTau);
    if (__any_sync(0xffffffffu, need)) st = p1_raise<E>(need, mn, g_valid, chunks_done, tG, tP, st);
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float p0 = (z[2 * i] == CUDART_INF_F) ? 0.f : ex2_approx(st.a - z[2 * i]);
        const float p1 = (z[2 * i + 1] == CUDART_INF_F) ? 0.f : ex2_approx(st.a - z[2 * i + 1]);
        st.l = add2(st.l, pk2(p0, p1));
        w[i] = pack_f16x2(p0, p1);
    }
    tmem_st_32x8(tP + chunks_done * 8, w);
    return st;
}

// ---- pass 2: P' = 2^(s*kmul - c2_j - r2 + kOff2), the positive left out (weight 0) ------------------------------------------------
// pass 2, fast chunk, exponent arguments only (the exponentials are taken one chunk later: software pipeline in the epilogue)
__device__ __forceinline__ void p2_x_fast(const uint32_t (&r)[16], uint32_t c2s, float kmul, float rowc, f32x2 (&x)[8]) {
    const f32x2 km = pk2(kmul, kmul), mone = pk2(-1.f, -1.f), rc = pk2(rowc, rowc);
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        const f32x2 ad0 = fma2(pk2(cc.x, cc.y), mone, rc), ad1 = fma2(pk2(cc.z, cc.w), mone, rc);
        x[2 * g4] = fma2(pk2(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1])), km, ad0);
        x[2 * g4 + 1] = fma2(pk2(__uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3])), km, ad1);
    }
}
// pass 2, fast chunk: P' of 16 columns as packed fp16 (unpipelined form)
__device__ __forceinline__ void p2_chunk_fast(const uint32_t (&r)[16], uint32_t c2s, float kmul, float rowc, uint32_t (&w)[8]) {
    const f32x2 km = pk2(kmul, kmul), mone = pk2(-1.f, -1.f), rc = pk2(rowc, rowc);
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        const f32x2 ad0 = fma2(pk2(cc.x, cc.y), mone, rc), ad1 = fma2(pk2(cc.z, cc.w), mone, rc);
        const f32x2 xa = fma2(pk2(__uint_as_float(r[4 * g4]), __uint_as_float(r[4 * g4 + 1])), km, ad0);
        const f32x2 xb = fma2(pk2(__uint_as_float(r[4 * g4 + 2]), __uint_as_float(r[4 * g4 + 3])), km, ad1);
        float p0, p1, p2, p3;
        if (2 * g4 < kPolyPairs) ex2_poly2(xa, p0, p1);
        else { upk2(xa, p0, p1); p0 = ex2_approx(p0); p1 = ex2_approx(p1); }
        if (2 * g4 + 1 < kPolyPairs) ex2_poly2(xb, p2, p3);
        else { upk2(xb, p2, p3); p2 = ex2_approx(p2); p3 = ex2_approx(p3); }
        w[2 * g4] = pack_f16x2(p0, p1);
        w[2 * g4 + 1] = pack_f16x2(p2, p3);
    }
}
__device__ __noinline__ void p2_chunk_checked(uint32_t tS_chunk, uint32_t c2s, float kmul, float rowc, int nb, int nT, bool row_ok, int dcol_abs,
                                              uint32_t tP_chunk) {
    uint32_t r[16];
    tmem_ld_32x16_issue(tS_chunk, r);
    tmem_ld_wait();
    uint32_t w[8];
#pragma unroll
    for (int g4 = 0; g4 < 4; ++g4) {
        const float4 cc = lds128(c2s + g4 * 16);
        const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
        float pv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int i = g4 * 4 + t;
            float v = ex2_approx(fmaf(__uint_as_float(r[i]), kmul, rowc - cv[t]));
            if (nb + i >= nT || !row_ok) v = 0.f;
            else if (nb + i == dcol_abs) v = 0.f;   // the positive: added by the combine kernel in fp32
            pv[t] = v;
        }
        w[2 * g4] = pack_f16x2(pv[0], pv[1]);
        w[2 * g4 + 1] = pack_f16x2(pv[2], pv[3]);
    }
    tmem_st_32x8(tP_chunk, w);
}

#define FL_TRACE(it, ev)                                                                                     \
    do {                                                                                                     \
        if (p.trace && (it) < 64) p.trace[((size_t)blockIdx.x * 64 + (it)) * 8 + (ev)] = gtime();             \
    } while (0)

template <int MODE, int E>
__global__ void __launch_bounds__(FlCfg<E>::kThreads, 1)
flash_kernel(const __grid_constant__ FlMaps maps, const __grid_constant__ FlParams p) {
    using Cfg = FlCfg<E>;
    constexpr int BN = Cfg::BN, NS = Cfg::NS, kSplit = Cfg::kSplit;
    const int u_begin = sk_begin(blockIdx.x, p.units, gridDim.x), u_end = sk_begin(blockIdx.x + 1, p.units, gridDim.x);
    const int my_units = u_end - u_begin;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // (programmatic dependent launch, see tt_softmax_flash.cu)
    if (my_units <= 0) return;
    if (p.trace && threadIdx.x == 0) p.trace[((size_t)blockIdx.x * 64 + 63) * 8 + 0] = gtime();   // CTA entry

    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    unsigned char* sR = smem;
    unsigned char* sT = sR + 2 * Cfg::kRBytes;                    // two R pairs (double buffered)
    unsigned char* sC2 = sT + Cfg::kStages * Cfg::kTBytes;
    FlBars* bars = reinterpret_cast<FlBars*>(sC2 + 8 * Cfg::kC2Bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kIssuerWarp0 = Cfg::kEpiWarps;
    if (warp == kIssuerWarp0 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            if (i < p.n_pass) { prefetch_tmap(&maps.r[i]); prefetch_tmap(&maps.t[i]); }
        }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->r_full[i], 1); mbar_init(&bars->r_empty[i], NS); }
        for (int i = 0; i < 8; ++i) { mbar_init(&bars->t_full[i], 1); mbar_init(&bars->t_empty[i], NS); }
        for (int i = 0; i < 4; ++i) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(&bars->s_full[i][b], 1);
                mbar_init(&bars->p_full[i][b], 4); mbar_init(&bars->p_empty[i][b], 1);
            }
            mbar_init(&bars->g_full[i], 1); mbar_init(&bars->g_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == kIssuerWarp0) tmem_alloc(&bars->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    asm volatile("griddepcontrol.wait;" ::: "memory");   // the kernels before this one have completed: operand copies, scales, lse are visible
    const uint32_t tmem = bars->tmem_base;
    const uint32_t b_rfull = smem_u32(&bars->r_full[0]), b_rempty = smem_u32(&bars->r_empty[0]), b_tfull = smem_u32(&bars->t_full[0]),
                   b_tempty = smem_u32(&bars->t_empty[0]), b_sfull = smem_u32(&bars->s_full[0][0]), b_pfull = smem_u32(&bars->p_full[0][0]),
                   b_pempty = smem_u32(&bars->p_empty[0][0]), b_gfull = smem_u32(&bars->g_full[0]), b_gempty = smem_u32(&bars->g_empty[0]);

    if (warp >= kIssuerWarp0) {
        // ===================== issuer of stream s (one lane) =====================
        //   first product   S_s[sub] = R_g . T[32 rows]^T     (N = 32)
        //   second product  G_s += P_s[sub] . T[32 rows]       (A = P from tensor memory, B = the same tile, MN-major; K = 32 = 2 instructions)
        // The issuers also share the TMA loads: issuer s fetches every tile j with j % NS == s, kLook tiles ahead, and the
        // pair of R panels when that tile opens a new pair (R is double buffered, so the switch does not drain the pipeline).
        // The warp runs converged with warp-uniform values (descriptors then live in uniform registers and a tcgen05.mma costs a few
        // issue slots; from inside a one-lane branch every operand goes through an R2UR election loop, ~100 cycles per MMA); only
        // the tcgen05 / TMA / arrive instructions themselves are predicated on the elected lane.
        {
            const int s = __shfl_sync(0xffffffffu, warp, 0) - kIssuerWarp0;
            const bool lead = elect_one();
            const int g = s / kSplit, h = s % kSplit;
            constexpr uint32_t idesc2 = make_idesc_f16_bmn(128, E);
            const uint32_t sR_a = smem_u32(sR), sT_a = smem_u32(sT);
            const uint32_t lbo = p.mn_lbo > 0 ? (uint32_t)p.mn_lbo : (uint32_t)(BN * 128), sbo = p.mn_sbo > 0 ? (uint32_t)p.mn_sbo : 1024u;
            const uint32_t tS = tmem + s * Cfg::kStreamCols, tG = tS + Cfg::kGCol;
            FlWalk c, pc;
            c.init(p, u_begin);
            pc = c;
            int pk = -1;   // pairs of panels met by the prefetch cursor - 1
            auto prefetch = [&](int j) {   // tile j of this CTA's range; pc stands on it
                const bool newseg = (j == 0 || pc.tile == 0);
                if (newseg) ++pk;
                if (j % NS == s) {
                    if (newseg) {   // every stream's last first-product on the pair that used this buffer is complete
                        mbar_wait_a(b_rempty + (pk & 1) * 8, ((pk >> 1) & 1) ^ 1);
                        if (lead) {
                            mbar_arrive_expect_tx(&bars->r_full[pk & 1], Cfg::kRBytes);
                            for (int gg = 0; gg < 2; ++gg)
                                for (int sl = 0; sl < Cfg::kSlabs; ++sl)
                                    tma_load_2d(sR + (pk & 1) * Cfg::kRBytes + gg * Cfg::kPanelBytes + sl * 128 * 128, &maps.r[pc.pass], &bars->r_full[pk & 1],
                                                sl * 64, (pc.pair * 2 + gg) * 128);
                        }
                    }
                    const int stage = j % Cfg::kStages;
                    const int n0 = pc.tile * BN;
                    mbar_wait_a(b_tempty + stage * 8, ((j / Cfg::kStages) & 1) ^ 1);
                    if (lead) {
                        mbar_arrive_expect_tx(&bars->t_full[stage], Cfg::kTBytes + Cfg::kC2Bytes);
                        unsigned char* dst = sT + stage * Cfg::kTBytes;
                        for (int sl = 0; sl < Cfg::kSlabs; ++sl) tma_load_2d(dst + sl * BN * 128, &maps.t[pc.pass], &bars->t_full[stage], sl * 64, n0);
                        bulk_copy_1d(sC2 + stage * Cfg::kC2Bytes, p.pass[pc.pass].colv2 + n0, Cfg::kC2Bytes, &bars->t_full[stage]);
                        if (s == 0) FL_TRACE(j, 0);
                    }
                    __syncwarp();
                }
                pc.next(p);
            };
            // sub-tile v = 2 * tile + sub lives in S buffer `sub` of the stream (columns [sub * 32, +32) of its tensor-memory block)
            constexpr uint32_t idesc1s = make_idesc_f16(128, kSW);
            auto first_product = [&](int it, int sub, int kseg) {   // tile `it` belongs to pair number kseg of this CTA
                const int stage = it % Cfg::kStages;
                const uint64_t ad0 = make_smem_desc(sR_a + (kseg & 1) * Cfg::kRBytes + g * Cfg::kPanelBytes, 16, 1024);
                const uint64_t bd0 = make_smem_desc(sT_a + stage * Cfg::kTBytes + (h * 64 + sub * kSW) * 128, 16, 1024);   // kSW rows of every slab
                if (lead) {
#pragma unroll
                    for (int kk = 0; kk < Cfg::kMma1; ++kk) {
                        const uint64_t ad = ad0 + (uint64_t)(((kk >> 2) * 128 * 128 + (kk & 3) * 32) >> 4);
                        const uint64_t bd = bd0 + (uint64_t)(((kk >> 2) * BN * 128 + (kk & 3) * 32) >> 4);
                        mma_f16(tS + sub * kSW, ad, bd, idesc1s, kk > 0 ? 1u : 0u);
                    }
                    mma_commit_a(b_sfull + (s * 2 + sub) * 8);
                    if (s == 0 && sub == 0) FL_TRACE(it, 1);
                }
                __syncwarp();
                if (p.trace && s == 0 && sub == 0) {   // debug: completion latency of the first product (blocks the issuer)
                    mbar_wait_a(b_sfull + (s * 2 + sub) * 8, it & 1);
                    if (lead) FL_TRACE(it, 4);
                }
            };
            int pf = 0;
            for (; pf < Cfg::kLook && pf < my_units; ++pf) prefetch(pf);
            // the first tile: both sub-tiles
            mbar_wait_a(b_rfull, 0);
            mbar_wait_a(b_tfull, 0);
            tc_fence_after();
            for (int sub = 0; sub < kSBuf; ++sub) first_product(0, sub, 0);
            const bool inorder = p.mn_lbo < 0;   // debug: 0 = wait for the second product's completion before refilling its buffer
            int k = -1;
            for (int it = 0; it < my_units; ++it) {
                if (pf < my_units) prefetch(pf++);
                const bool seg_start = (it == 0 || c.tile == 0);
                const bool pair_end = c.last_tile();
                const bool seg_end = (it == my_units - 1 || pair_end);
                if (seg_start) {
                    ++k;
                    mbar_wait_a(b_gempty + s * 8, (k & 1) ^ 1);   // the epilogue has drained the previous segment's G
                }
                const int kn = k + (pair_end ? 1 : 0);            // pair number of the next tile
                const int stage = it % Cfg::kStages;
                const uint64_t bd0 = make_smem_desc(sT_a + stage * Cfg::kTBytes, lbo, sbo);
#pragma unroll
                for (int sub = 0; sub < kSBuf; ++sub) {
                    mbar_wait_a(b_pfull + (s * 2 + sub) * 8, it & 1);
                    tc_fence_after();
                    if (lead) {
#pragma unroll
                        for (int k2 = 0; k2 < kSW / 16; ++k2)   // 16 rows of T (2048 bytes of every slab) per instruction
                            mma_f16_ts(tG, tS + sub * kSW + k2 * 8, bd0 + (uint64_t)(((h * 64 + sub * kSW + k2 * 16) * 128) >> 4), idesc2,
                                       (!seg_start || sub > 0 || k2 > 0) ? 1u : 0u);
                        if (sub == kSBuf - 1) {
                            mma_commit_a(b_tempty + stage * 8);   // (one of NS arrivals) this stream is done with the tile and its column term
                            if (seg_end) {
                                mma_commit_a(b_gfull + s * 8);
                                mma_commit_a(b_rempty + (k & 1) * 8);   // (one of NS arrivals) ... and with this pair of panels
                            }
                        }
                        mma_commit_a(b_pempty + (s * 2 + sub) * 8);
                        if (s == 0 && sub == kSBuf - 1) FL_TRACE(it, 7);
                    }
                    __syncwarp();
                    if (p.trace && s == 0 && sub == kSBuf - 1) {   // debug: completion latency of the second product (blocks the issuer)
                        mbar_wait_a(b_pempty + (s * 2 + sub) * 8, it & 1);
                        if (lead) FL_TRACE(it, 5);
                    }
                    if (it + 1 < my_units) {   // refill this S buffer with the same sub-tile of the next tile
                        if (sub == 0) {
                            if (kn != k) mbar_wait_a(b_rfull + (kn & 1) * 8, (kn >> 1) & 1);
                            mbar_wait_a(b_tfull + ((it + 1) % Cfg::kStages) * 8, ((it + 1) / Cfg::kStages) & 1);
                        }
                        // The second product reads P from this buffer.  One thread's tcgen05.mma execute in issue order, so the refill may
                        // follow directly (inorder); otherwise wait until the second product has completed.
                        if (!inorder) mbar_wait_a(b_pempty + (s * 2 + sub) * 8, it & 1);
                        tc_fence_after();
                        first_product(it + 1, sub, kn);
                    }
                }
                c.next(p);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: stream s = warp / 4 (panel g, half h), TMEM lane quarter q = warp % 4 =====================
        const int s = warp >> 2, q = warp & 3;
        const int g = s / kSplit, h = s % kSplit;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t tS0 = tmem + lane_addr + s * Cfg::kStreamCols, tG = tS0 + Cfg::kGCol;
        // (barrier addresses are formed from one base at the use sites, and what only the ends of a segment need -- row, partial slot,
        // padded row count -- is recomputed there: the loop is register-bound at 96)
        const uint32_t bs_full = b_sfull + s * 16;
#define bp_full (bs_full + (uint32_t)(offsetof(FlBars, p_full) - offsetof(FlBars, s_full)))
#define bp_empty (bs_full + (uint32_t)(offsetof(FlBars, p_empty) - offsetof(FlBars, s_full)))
        const uint32_t bg_full = b_gfull + s * 8, bg_empty = b_gempty + s * 8;
        FlWalk c;
        c.init(p, u_begin);
        int k = -1;
        int nT = 0, dlo = 0, dcol_abs = -1;
        bool rows_in = false;
        float kmul = 0.f;
        P1State st{CUDART_INF_F, pk2(0.f, 0.f)};
        float rowc = 0.f;   // pass 2: kOff2 - rowv*log2e
        const uint32_t c2s0 = smem_u32(sC2) + h * 256;
        for (int it = 0; it < my_units; ++it, c.next(p)) {
            const bool seg_start = (it == 0 || c.tile == 0);
            const bool seg_end = (it == my_units - 1 || c.last_tile());
            const FlPass& ps = p.pass[c.pass];   // (only dereferenced at the ends of a segment)
            const int wrow0 = (c.pair * 2 + g) * 128 + q * 32, row = wrow0 + lane;
            const bool row_ok = row < ps.nR;
            if (seg_start) {
                ++k;
                nT = ps.nT;
                rows_in = wrow0 + 32 <= ps.nR;
                dlo = wrow0 + ps.d;                                  // positives of this warp's rows sit in columns [dlo, dlo + 32)
                dcol_abs = row_ok ? row + ps.d : -1;
                kmul = __ldg(p.kmul + c.pass);
                st = P1State{CUDART_INF_F, pk2(0.f, 0.f)};
                if (MODE == kP2) rowc = kOff2 - ((row_ok && ps.rowv) ? __ldg(ps.rowv + row) * kLog2e : 0.f);
            }
            const int stage = it % Cfg::kStages;
            // (the staged column term needs no wait of its own: S of this tile exists, so the issuer had seen the tile's TMA barrier)
#pragma unroll
            for (int sub = 0; sub < kSBuf; ++sub) {
                constexpr int NCH = kSW / kCW;                        // 16-column chunks per sub-tile
                const int n0 = c.tile * BN + h * 64 + sub * kSW;      // first column of this sub-tile
                const uint32_t tS = tS0 + sub * kSW, tP = tS;
                const uint32_t c2s = c2s0 + stage * Cfg::kC2Bytes + sub * kSW * 4;
                // warp-uniform: the sub-tile is fully in range (else: the checked path) / holds a positive of this warp's rows
                const bool in_range = (n0 + kSW <= nT) && rows_in && p.mn_sbo >= 0;   // (mn_sbo < 0: debug, checked path everywhere)
                const bool has_d = dlo < n0 + kSW && dlo + 32 > n0;
                const bool g_valid = !seg_start || sub > 0;           // a second product of this segment has been issued
                // pass 1, rare: before G is rescaled every second product issued so far must be complete (with one S buffer that is
                // implied by S being there; with two, the previous sub-tile's may still run)
                auto g_quiet = [&]() {
                    if (kSBuf > 1 && g_valid) {
                        mbar_wait_a(bp_empty + (sub ^ 1) * 8, (sub == 1 ? it : it - 1) & 1);
                        tc_fence_after();
                    }
                };
                mbar_wait_a(bs_full + sub * 8, it & 1);
                tc_fence_after();
                if (sub == 0 && lane == 0 && warp == 0) FL_TRACE(it, 2);
                // software pipeline over the chunks: the address arithmetic of chunk j+1 (FFMA2, min) sits in the same basic block as the
                // exponentials of chunk j, so the scheduler overlaps them.  DIAG: chunks may hold positives of this warp's rows.
                auto fast_sub = [&](auto diag_tag) {
                    constexpr bool DIAG = decltype(diag_tag)::value;
                    uint32_t rb[2][kCW];
                    tmem_ld_32x16_issue(tS, rb[0]);
                    tmem_ld_wait();
                    if (NCH > 1) tmem_ld_32x16_issue(tS + kCW, rb[1]);
                    if (DIAG) mask_positive(rb[0], n0, dcol_abs);
                    if constexpr (MODE == kP1) {
                        f32x2 zn[2][kCW / 2];
                        float cmin = p1_zn_fast(rb[0], c2s, kmul, zn[0]);
                        {
                            const bool need = cmin < st.a - (kOff1 + kTau);   // this chunk exceeds the row's reference by more than 2^kTau
                            if (__any_sync(0xffffffffu, need)) { g_quiet(); st = p1_raise<E>(need, cmin, g_valid, 0, tG, tP, st); }
                        }
#pragma unroll
                        for (int j = 0; j < NCH; ++j) {
                            uint32_t w[kCW / 2];
                            if (j + 1 < NCH) {
                                tmem_ld_wait();                                                   // chunk j+1 is in registers
                                if (DIAG) mask_positive(rb[(j + 1) & 1], n0 + (j + 1) * kCW, dcol_abs);
                                cmin = p1_zn_fast(rb[(j + 1) & 1], c2s + (j + 1) * kCW * 4, kmul, zn[(j + 1) & 1]);
                                if (j + 2 < NCH) tmem_ld_32x16_issue(tS + (j + 2) * kCW, rb[j & 1]);
                            }
                            p1_exp_fast(zn[j & 1], st.a, st.l, w);
                            tmem_st_32x8(tP + j * (kCW / 2), w);
                            if (j + 1 < NCH) {
                                const bool need = cmin < st.a - (kOff1 + kTau);
                                if (__any_sync(0xffffffffu, need)) { g_quiet(); st = p1_raise<E>(need, cmin, g_valid, j + 1, tG, tP, st); }
                            }
                        }
                    } else {
                        f32x2 x[2][kCW / 2];
                        p2_x_fast(rb[0], c2s, kmul, rowc, x[0]);
#pragma unroll
                        for (int j = 0; j < NCH; ++j) {
                            uint32_t w[kCW / 2];
                            if (j + 1 < NCH) {
                                tmem_ld_wait();
                                if (DIAG) mask_positive(rb[(j + 1) & 1], n0 + (j + 1) * kCW, dcol_abs);
                                p2_x_fast(rb[(j + 1) & 1], c2s + (j + 1) * kCW * 4, kmul, rowc, x[(j + 1) & 1]);
                                if (j + 2 < NCH) tmem_ld_32x16_issue(tS + (j + 2) * kCW, rb[j & 1]);
                            }
#pragma unroll
                            for (int i = 0; i < kCW / 2; ++i) {
                                float x0, x1;
                                upk2(x[j & 1][i], x0, x1);
                                w[i] = pack_f16x2(ex2_approx(x0), ex2_approx(x1));
                            }
                            tmem_st_32x8(tP + j * (kCW / 2), w);
                        }
                    }
                };
                if (in_range && !has_d) {
                    fast_sub(std::false_type{});
                } else if (in_range) {
                    fast_sub(std::true_type{});
                } else {
#pragma unroll 1
                    for (int j = 0; j < NCH; ++j) {
                        if constexpr (MODE == kP1) {
                            g_quiet();   // (the checked chunk may raise the reference: G has to be quiescent)
                            st = p1_chunk_checked<E>(tS + j * kCW, c2s + j * kCW * 4, kmul, n0 + j * kCW, nT, dcol_abs, g_valid, j, tG, tP, st);
                        } else {
                            p2_chunk_checked(tS + j * kCW, c2s + j * kCW * 4, kmul, rowc, n0 + j * kCW, nT, row_ok, dcol_abs, tP + j * (kCW / 2));
                        }
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(bp_full + sub * 8);
            }
            if (lane == 0 && warp == 0) FL_TRACE(it, 3);
            if (seg_end) {
                const int rows_pad = c.m_pairs * 256;
                const int part = (blockIdx.x - sk_owner(ps.unit0 + c.pair * c.n_tiles, p.units, gridDim.x)) * kSplit + h;
                mbar_wait_a(bg_full, k & 1);
                tc_fence_after();
#pragma unroll
                for (int cg = 0; cg < E / 32; ++cg) {
                    float v[32];
                    tmem_ld_32x32(tG + cg * 32, v);
                    if (row_ok) {   // 32 lanes = 32 consecutive rows of one float4 column: 512 contiguous bytes per store instruction
                        float4* dst = reinterpret_cast<float4*>(ps.out_g) + ((int64_t)part * (E / 4) + cg * 8) * rows_pad + row;
#pragma unroll
                        for (int g4 = 0; g4 < 8; ++g4) dst[(int64_t)g4 * rows_pad] = make_float4(v[g4 * 4], v[g4 * 4 + 1], v[g4 * 4 + 2], v[g4 * 4 + 3]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(bg_empty);
                if (MODE == kP1 && row_ok) {
                    float l0, l1;
                    upk2(st.l, l0, l1);
                    ps.out_m[(int64_t)part * rows_pad + row] = -st.a;
                    ps.out_l[(int64_t)part * rows_pad + row] = l0 + l1;
                }
            }
        }
    }
#undef bp_full
#undef bp_empty
    __syncthreads();
    if (p.trace && threadIdx.x == 0) p.trace[((size_t)blockIdx.x * 64 + 63) * 8 + 1] = gtime();   // CTA exit
    if (warp == kIssuerWarp0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace tc
}  // namespace tt
