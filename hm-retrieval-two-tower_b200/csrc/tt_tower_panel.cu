// tt_tower_panel.cu -- tower Dense kernels for the common case where a whole reduction extent fits in shared
// memory (K <= 256 inputs, N <= 256 units): both operands are staged once per CTA, the inner loop touches only
// shared memory, and the embedding gather writes straight into the staged A operand.
//
// Same arithmetic contract as tt_tower.cu (reference: input_layer.py:37-41,61-68; tower.py:41-49,72-75):
// forward accumulates k-ascending with one fmaf per term (bit-identical to oracle/tt_oracle.c:tto_dense_fmaf);
// dW/db are summed over fixed 64-row chunks and then over chunks in order (deterministic).
#include "tt_common.cuh"

namespace tt {

struct FeatArr {   // same layout as in tt_tower.cu
    tt_feature f[TT_MAX_FEATURES];
    int n;
};

constexpr int PT = 64;        // tile edge
constexpr int PS = PT + 4;    // padded shared-memory row (floats): float4-aligned, spreads banks
constexpr int kPanelMaxK = 256;

// ---- forward: Y = relu?(X.W + b), X either dense or gathered from the embedding tables ----------------------
// grid (ceil(B/64), ceil(N/64)), 256 threads, dynamic smem = 2 * K * PS floats (+ ids for the gather)
// PM: batch rows per CTA (64, or 32 when that is what it takes to give every SM two CTAs: these launches are latency-bound)
template <bool kGather, int PM>
__global__ void __launch_bounds__(256) dense_fwd_panel_kernel(const __grid_constant__ FeatArr fa, const float* __restrict__ X, int ldx,
                                                              const float* __restrict__ W, const float* __restrict__ bias,
                                                              float* __restrict__ Xout, float* __restrict__ Y, int ldy,
                                                              float* __restrict__ Ytf32, int B, int K, int N, int relu) {
    extern __shared__ __align__(16) float sm[];
    constexpr int PMS = PM + 4;      // padded row of the staged A operand
    constexpr int RM = PM / 16;      // batch rows per thread
    float* XsT = sm;                 // [K][PMS]  XsT[k][r]
    float* Ws = sm + (size_t)K * PMS;  // [K][PS]  Ws[k][n]
    const int m0 = blockIdx.x * PM, n0 = blockIdx.y * PT;
    const int tid = threadIdx.x;
    // All staging loops are flat and free of loop-carried dependences, so each thread keeps many independent
    // global loads in flight (these kernels are latency-bound: one CTA per SM, one staging phase per CTA).
#pragma unroll 8
    for (int idx = tid; idx < K * PT; idx += 256) {   // W tile, coalesced along n
        const int k = idx >> 6, n = idx & 63;
        Ws[k * PS + n] = (n0 + n < N) ? __ldg(W + (int64_t)k * N + n0 + n) : 0.f;
    }
    if constexpr (kGather) {
        __shared__ int32_t s_ids[PM * TT_MAX_FEATURES];
        __shared__ uint8_t s_feat[kPanelMaxK];
        __shared__ uint16_t s_off[kPanelMaxK];
        for (int c = tid; c < K; c += 256) {
            uint8_t ff = 255;
            uint16_t oo = 0;
            for (int f = 0; f < fa.n; ++f)
                if (c >= fa.f[f].col && c < fa.f[f].col + fa.f[f].e) { ff = (uint8_t)f; oo = (uint16_t)(c - fa.f[f].col); }
            s_feat[c] = ff;
            s_off[c] = oo;
        }
        for (int i = tid; i < PM * fa.n; i += 256) {
            const int r = i / fa.n, f = i - r * fa.n;
            int id = 0;
            if (m0 + r < B && fa.f[f].table != nullptr) {
                id = __ldg(reinterpret_cast<const int32_t*>(fa.f[f].src) + m0 + r);
                if ((unsigned)id >= (unsigned)fa.f[f].rows) id = 0;
            }
            s_ids[r * TT_MAX_FEATURES + f] = id;
        }
        __syncthreads();
        const bool write_x = (Xout != nullptr) && blockIdx.y == 0;
#pragma unroll 8
        for (int idx = tid; idx < PM * K; idx += 256) {   // consecutive threads -> consecutive columns of one row
            const int r = idx / K, k = idx - r * K;
            const int row = m0 + r;
            const int f = s_feat[k];
            float v = 0.f;
            if (row < B && f != 255) {
                const tt_feature& ft = fa.f[f];
                v = (ft.table == nullptr) ? __ldg(reinterpret_cast<const float*>(ft.src) + row)
                                          : __ldg(feature_row(ft, s_ids[r * TT_MAX_FEATURES + f]) + s_off[k]);
            }
            XsT[k * PMS + r] = v;
            if (write_x && row < B) Xout[(int64_t)row * ldx + k] = v;
        }
        if (write_x) {
            const int padc = ldx - K;
            for (int idx = tid; idx < PM * padc; idx += 256) {   // zero the padding columns
                const int r = idx / padc, c = K + idx - r * padc;
                if (m0 + r < B) Xout[(int64_t)(m0 + r) * ldx + c] = 0.f;
            }
        }
    } else {
#pragma unroll 8
        for (int idx = tid; idx < PM * K; idx += 256) {
            const int r = idx / K, k = idx - r * K;
            XsT[k * PMS + r] = (m0 + r < B) ? __ldg(X + (int64_t)(m0 + r) * ldx + k) : 0.f;
        }
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    float acc[RM][4] = {};
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        float a[RM];
        if constexpr (RM == 4) {
            const float4 a4 = *reinterpret_cast<const float4*>(XsT + k * PMS + ty * 4);
            a[0] = a4.x; a[1] = a4.y; a[2] = a4.z; a[3] = a4.w;
        } else {
            const float2 a2 = *reinterpret_cast<const float2*>(XsT + k * PMS + ty * 2);
            a[0] = a2.x; a[1] = a2.y;
        }
        const float4 b4 = *reinterpret_cast<const float4*>(Ws + k * PS + tx * 4);
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int m = m0 + ty * RM + i;
        if (m >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float y = acc[i][j];
            if (bias) y = __fadd_rn(y, __ldg(bias + n));
            if (relu && !(y > 0.f)) y = 0.f;
            Y[(int64_t)m * ldy + n] = y;
            if (Ytf32) Ytf32[(int64_t)m * ldy + n] = tf32_rn(y);
        }
    }
}

// ---- forward, common tower layer (N == 64 units, K <= 96 inputs): vectorised staging ------------------------------------------
// Same arithmetic as dense_fwd_panel_kernel (k ascending, one fmaf per term: bit-identical), different data movement: the input
// block is staged ROW-major with one 128-bit load per four columns of a table row (a quarter-warp reads a whole 64-wide row:
// coalesced; the scalar version issued four times the loads and a division per element), W with 128-bit loads, the saved input
// and both outputs with 128-bit stores where the column offset allows.  The W tile, the ids and the piece table are fetched
// in one phase, the table rows in the next: two dependent round trips instead of three.
constexpr int kFwdKMax = 96;
constexpr int kFwdXs = kFwdKMax + 4;            // row stride of the staged input block (floats)
constexpr int kFwdPM = 32;                      // batch rows per CTA
constexpr int kFwdMaxPieces = 64;               // float4 pieces / scalars of one input row
template <bool kGather>
__global__ void __launch_bounds__(256) dense_fwd_fused64_kernel(const __grid_constant__ FeatArr fa, const float* __restrict__ X, int ldx,
                                                                const float* __restrict__ W, const float* __restrict__ bias,
                                                                float* __restrict__ Xout, float* __restrict__ Y, int ldy,
                                                                float* __restrict__ Ytf32, int B, int K, int relu) {
    constexpr int N = PT;
    __shared__ __align__(16) float Xs[kFwdPM * kFwdXs];
    __shared__ __align__(16) float Ws[kFwdKMax * PS];
    __shared__ int32_t s_ids[kFwdPM * TT_MAX_FEATURES];
    __shared__ uint8_t s_pf[kFwdMaxPieces];      // piece -> feature
    __shared__ uint8_t s_po[kFwdMaxPieces];      // piece -> first column inside the feature (multiple of 4; 0 for a numeric)
    __shared__ int s_np;
    const int tid = threadIdx.x, m0 = blockIdx.x * kFwdPM;
    for (int q = tid; q < K * 16; q += 256) {
        const int k = q >> 4, c4 = q & 15;
        *reinterpret_cast<float4*>(Ws + k * PS + 4 * c4) = __ldg(reinterpret_cast<const float4*>(W + (int64_t)k * N) + c4);
    }
    if constexpr (kGather) {
        if (tid == 0) {   // pieces of one input row: a numeric feature is one scalar, a table feature e/4 float4s (e % 4 == 0 checked on the host)
            int np = 0;
            for (int f = 0; f < fa.n; ++f) {
                const int cnt = fa.f[f].table == nullptr ? 1 : fa.f[f].e / 4;
                for (int c = 0; c < cnt; ++c) { s_pf[np] = (uint8_t)f; s_po[np] = (uint8_t)(4 * c); ++np; }
            }
            s_np = np;
        }
        for (int i = tid; i < kFwdPM * fa.n; i += 256) {
            const int r = i / fa.n, f = i - r * fa.n;
            int id = 0;
            if (m0 + r < B && fa.f[f].table != nullptr) {
                id = __ldg(reinterpret_cast<const int32_t*>(fa.f[f].src) + m0 + r);
                if ((unsigned)id >= (unsigned)fa.f[f].rows) id = 0;
            }
            s_ids[r * TT_MAX_FEATURES + f] = id;
        }
        __syncthreads();
        const int np = s_np;
        const bool write_x = Xout != nullptr;
        // consecutive threads -> consecutive pieces of one row.  Four pieces per thread and round: all their loads are issued before
        // the first store, so a row that lives in another GPU's HBM (row-sharded tables) costs one NVLink round trip per round
        constexpr int kGB = 4;
        for (int q0 = tid; q0 < kFwdPM * np; q0 += 256 * kGB) {
            float4 v[kGB];
            int rr[kGB], ff[kGB], oo[kGB];
#pragma unroll
            for (int u = 0; u < kGB; ++u) {
                const int q = q0 + 256 * u;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                rr[u] = -1;
                if (q < kFwdPM * np) {
                    const int r = q / np, pc = q - r * np;
                    rr[u] = r; ff[u] = s_pf[pc]; oo[u] = s_po[pc];
                    const tt_feature& ft = fa.f[ff[u]];
                    if (m0 + r < B) {
                        if (ft.table == nullptr) v[u].x = __ldg(reinterpret_cast<const float*>(ft.src) + m0 + r);
                        else v[u] = __ldg(reinterpret_cast<const float4*>(feature_row(ft, s_ids[r * TT_MAX_FEATURES + ff[u]]) + oo[u]));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kGB; ++u) {
                if (rr[u] < 0) continue;
                const tt_feature& ft = fa.f[ff[u]];
                const int row = m0 + rr[u], col = ft.col + oo[u];
                float* xs = Xs + rr[u] * kFwdXs + col;
                float* xo = Xout + (int64_t)row * ldx + col;
                const bool wx = write_x && row < B;
                if (ft.table == nullptr) {
                    xs[0] = v[u].x;
                    if (wx) xo[0] = v[u].x;
                } else if ((col & 3) == 0) {
                    *reinterpret_cast<float4*>(xs) = v[u];
                    if (wx) *reinterpret_cast<float4*>(xo) = v[u];
                } else {
                    xs[0] = v[u].x; xs[1] = v[u].y; xs[2] = v[u].z; xs[3] = v[u].w;
                    if (wx) { xo[0] = v[u].x; xo[1] = v[u].y; xo[2] = v[u].z; xo[3] = v[u].w; }
                }
            }
        }
        if (write_x) {
            const int padc = ldx - K;
            for (int idx = tid; idx < kFwdPM * padc; idx += 256) {   // zero the padding columns
                const int r = idx / padc, c = K + idx - r * padc;
                if (m0 + r < B) Xout[(int64_t)(m0 + r) * ldx + c] = 0.f;
            }
        }
    } else {
        const int k4 = (K + 3) >> 2;   // ldx >= 4 * k4 (checked on the host)
        for (int q = tid; q < kFwdPM * k4; q += 256) {
            const int r = q / k4, c4 = q - r * k4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + r < B) v = __ldg(reinterpret_cast<const float4*>(X + (int64_t)(m0 + r) * ldx) + c4);
            *reinterpret_cast<float4*>(Xs + r * kFwdXs + 4 * c4) = v;
        }
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;      // rows 2 ty, 2 ty + 1; columns 4 tx .. 4 tx + 3
    float acc[2][4] = {};
    const float* x0 = Xs + (ty * 2) * kFwdXs;
    const float* x1 = x0 + kFwdXs;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float a0 = x0[k], a1 = x1[k];
        const float4 b4 = *reinterpret_cast<const float4*>(Ws + k * PS + tx * 4);
        acc[0][0] = fmaf(a0, b4.x, acc[0][0]); acc[0][1] = fmaf(a0, b4.y, acc[0][1]); acc[0][2] = fmaf(a0, b4.z, acc[0][2]); acc[0][3] = fmaf(a0, b4.w, acc[0][3]);
        acc[1][0] = fmaf(a1, b4.x, acc[1][0]); acc[1][1] = fmaf(a1, b4.y, acc[1][1]); acc[1][2] = fmaf(a1, b4.z, acc[1][2]); acc[1][3] = fmaf(a1, b4.w, acc[1][3]);
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias) + tx);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty * 2 + i;
        if (m >= B) continue;
        float y[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
        if (bias) { y[0] = __fadd_rn(y[0], bv.x); y[1] = __fadd_rn(y[1], bv.y); y[2] = __fadd_rn(y[2], bv.z); y[3] = __fadd_rn(y[3], bv.w); }
        if (relu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (!(y[j] > 0.f)) y[j] = 0.f;
        }
        *reinterpret_cast<float4*>(Y + (int64_t)m * ldy + 4 * tx) = make_float4(y[0], y[1], y[2], y[3]);
        if (Ytf32) *reinterpret_cast<float4*>(Ytf32 + (int64_t)m * ldy + 4 * tx) = make_float4(tf32_rn(y[0]), tf32_rn(y[1]), tf32_rn(y[2]), tf32_rn(y[3]));
    }
}

// ---- backward dX = dpre.W^T  (reduction over N <= 256) ----------------------------------------------------------
// grid (ceil(B/64), ceil(K/64)); smem: DsT[N][PS] (dpre^T tile) + WsT[N][PS] (WsT[n][kk] = W[k0+kk][n])
__device__ __forceinline__ void dense_bwd_dx_tile(float* sm, int bx, int by, const float* __restrict__ dY, int lddy, const float* __restrict__ Yv,
                                                  int ldy, const float* __restrict__ W, float* __restrict__ dX, int lddx, int B, int K, int N,
                                                  int relu) {
    float* DsT = sm;
    float* WsT = sm + (size_t)N * PS;
    const int m0 = bx * PT, k0 = by * PT;
    const int tid = threadIdx.x;
#pragma unroll 8
    for (int idx = tid; idx < PT * N; idx += 256) {
        const int r = idx / N, n = idx - r * N;
        const int row = m0 + r;
        float g = 0.f;
        if (row < B) {
            g = __ldg(dY + (int64_t)row * lddy + n);
            if (relu && !(__ldg(Yv + (int64_t)row * ldy + n) > 0.f)) g = 0.f;
        }
        DsT[n * PS + r] = g;
    }
#pragma unroll 8
    for (int idx = tid; idx < PT * N; idx += 256) {
        const int kk = idx / N, n = idx - kk * N;
        WsT[n * PS + kk] = (k0 + kk < K) ? __ldg(W + (int64_t)(k0 + kk) * N + n) : 0.f;
    }
    __syncthreads();
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4] = {};
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
        const float4 a4 = *reinterpret_cast<const float4*>(DsT + n * PS + ty * 4);
        const float4 b4 = *reinterpret_cast<const float4*>(WsT + n * PS + tx * 4);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= B) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < K) dX[(int64_t)m * lddx + k] = acc[i][j];
        }
    }
}

// ---- backward dW / db: partial[chunk][k][n] = sum_{b in chunk} X[b][k] * dpre[b][n]; row k == K is the bias gradient --
// grid (nchunk, ceil((K+1)/96), ceil(N/64)); a chunk is `rows` batch rows processed 64 at a time.
// thread (ty 0..15, tx 0..15): rows k0 + ty + 16*i (i < 6), cols n0 + 4*tx .. +3.
constexpr int kDwRows = 96;
__device__ __forceinline__ void dense_bwd_dw_tile(float* sm, int bx, int by, int bz, const float* __restrict__ X, int ldx,
                                                  const float* __restrict__ dY, int lddy, const float* __restrict__ Yv, int ldy,
                                                  float* __restrict__ partial, int B, int K, int N, int rows, int relu) {
    float(*Xs)[kDwRows + 1] = reinterpret_cast<float(*)[kDwRows + 1]>(sm);                       // [PT][kDwRows + 1]
    float(*Ds)[PS] = reinterpret_cast<float(*)[PS]>(sm + (PT * (kDwRows + 1) + 3) / 4 * 4);     // [PT][PS], 16-byte aligned
    const int b_begin = bx * rows, b_end = min(B, b_begin + rows);
    const int k0 = by * kDwRows, n0 = bz * PT;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[6][4] = {};
    for (int bb = b_begin; bb < b_end; bb += PT) {
#pragma unroll 8
        for (int idx = tid; idx < PT * kDwRows; idx += 256) {
            const int r = idx / kDwRows, kk = idx - r * kDwRows;
            const int row = bb + r, k = k0 + kk;
            float v = 0.f;
            if (row < b_end) v = k < K ? __ldg(X + (int64_t)row * ldx + k) : (k == K ? 1.0f : 0.f);
            Xs[r][kk] = v;
        }
#pragma unroll 8
        for (int idx = tid; idx < PT * PT; idx += 256) {
            const int r = idx >> 6, n = idx & 63;
            const int row = bb + r;
            float g = 0.f;
            if (row < b_end && n0 + n < N) {
                g = __ldg(dY + (int64_t)row * lddy + n0 + n);
                if (relu && !(__ldg(Yv + (int64_t)row * ldy + n0 + n) > 0.f)) g = 0.f;
            }
            Ds[r][n] = g;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < PT; ++r) {
            const float4 d4 = *reinterpret_cast<const float4*>(&Ds[r][tx * 4]);
            const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const float a = Xs[r][ty + 16 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a, d[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    float* out = partial + (int64_t)bx * (K + 1) * N;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int k = k0 + ty + 16 * i;
        if (k > K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) out[(int64_t)k * N + n] = acc[i][j];
        }
    }
}

// ---- one launch for the whole layer backward: dX tiles and dW/db chunk partials are independent -----------------------
// blocks [0, n_dx) compute dX tiles (grid_dx = (gx, gy)); the rest compute dW partials (grid_dw = (nchunk, hy, hz)).
constexpr int kDwSmemFloats = (PT * (kDwRows + 1) + 3) / 4 * 4 + PT * PS;
// three CTAs per SM (<= 85 registers): the 384 CTAs of a B = 8192 layer then fit one wave instead of 1.3
__global__ void __launch_bounds__(256, 3) dense_bwd_panel_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                              const float* __restrict__ Yv, int ldy, const float* __restrict__ dY, int lddy,
                                                              float* __restrict__ dX, int lddx, float* __restrict__ partial, int B, int K, int N,
                                                              int relu, int rows, int n_dx, int gx, int nchunk, int hy) {
    extern __shared__ __align__(16) float sm[];
    const int b = blockIdx.x;
    if (b < n_dx) {
        dense_bwd_dx_tile(sm, b % gx, b / gx, dY, lddy, Yv, ldy, W, dX, lddx, B, K, N, relu);
    } else {
        const int d = b - n_dx;
        dense_bwd_dw_tile(sm, d % nchunk, (d / nchunk) % hy, d / (nchunk * hy), X, ldx, dY, lddy, Yv, ldy, partial, B, K, N, rows, relu);
    }
}

// ---- fused backward for the common tower layer (N == 64 units, K + 1 <= 96 inputs incl. the bias row) --------------------------
// One CTA per chunk of `rows` batch rows (the SAME chunks as the split kernel above, so dW / db come out bit-identical): per 64
// rows it stages dpre = dY.[Y > 0] ONCE, row-major and with 128-bit loads (no transposing stores: the split kernel's are 8-way
// bank-conflicted), then computes that block's dX rows (n ascending, one fmaf per term) and adds the block to the chunk's dW / db
// accumulators (rows ascending).  One launch of B/rows CTAs replaces 3x as many CTAs that each staged their own copies; the inner
// loops read shared memory with conflict-free 128-bit loads (thread tx owns rows tx, tx+16, ... of the W tile: 16-byte slots
// 17 tx mod 8 are all different).
constexpr int kFusedKMax = 96;                  // K + 1 <= 96
constexpr int kFusedXs = kFusedKMax + 4;        // row stride of the staged X block (floats)
constexpr int kFusedSmemFloats = PT * PS + kFusedKMax * PS + PT * kFusedXs;
__global__ void __launch_bounds__(256, 2) dense_bwd_fused64_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                                  const float* __restrict__ Yv, int ldy, const float* __restrict__ dY, int lddy,
                                                                  float* __restrict__ dX, int lddx, float* __restrict__ partial, int B, int K,
                                                                  int relu, int rows) {
    extern __shared__ __align__(16) float sm[];
    constexpr int N = PT;
    float* Ds = sm;                              // [64][PS]   dpre block, row-major
    float* Ws = Ds + PT * PS;                    // [96][PS]   W, rows >= K zero
    float* Xs = Ws + kFusedKMax * PS;            // [64][100]  X block, column K = 1 (bias row), columns > K zero
    // (a 512-thread variant with the dX and the dW product on separate warp groups was measured: 12 us alone, but one CTA then
    //  fills an SM's register file and the two towers' launches no longer overlap -- 0.2118 vs 0.2102 ms per step)
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int b_begin = blockIdx.x * rows, b_end = min(B, b_begin + rows);
    const int k4 = (K + 3) >> 2;                 // float4 pieces of an X row that hold its K columns (ldx >= 4 * k4)
    for (int q = tid; q < kFusedKMax * 16; q += 256) {
        const int k = q >> 4, c4 = q & 15;
        const float4 w = k < K ? __ldg(reinterpret_cast<const float4*>(W + (int64_t)k * N) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4*>(Ws + k * PS + 4 * c4) = w;
    }
    float accw[6][4] = {};
    for (int bb = b_begin; bb < b_end; bb += PT) {
#pragma unroll
        for (int q = tid; q < PT * 16; q += 256) {
            const int r = q >> 4, c4 = q & 15;
            const int row = bb + r;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < b_end) {
                g = __ldg(reinterpret_cast<const float4*>(dY + (int64_t)row * lddy) + c4);
                if (relu) {
                    const float4 y = __ldg(reinterpret_cast<const float4*>(Yv + (int64_t)row * ldy) + c4);
                    g.x = y.x > 0.f ? g.x : 0.f; g.y = y.y > 0.f ? g.y : 0.f; g.z = y.z > 0.f ? g.z : 0.f; g.w = y.w > 0.f ? g.w : 0.f;
                }
            }
            *reinterpret_cast<float4*>(Ds + r * PS + 4 * c4) = g;
        }
        for (int q = tid; q < PT * (kFusedXs / 4); q += 256) {
            const int r = q / (kFusedXs / 4), c4 = q - r * (kFusedXs / 4);
            const int row = bb + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < b_end && c4 < k4) v = __ldg(reinterpret_cast<const float4*>(X + (int64_t)row * ldx) + c4);
            float* e = reinterpret_cast<float*>(&v);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int k = 4 * c4 + t;
                if (k > K) e[t] = 0.f;                                   // padding of the input buffer is not trusted
                else if (k == K) e[t] = row < b_end ? 1.0f : 0.f;       // the bias row of dW
            }
            *reinterpret_cast<float4*>(Xs + r * kFusedXs + 4 * c4) = v;
        }
        __syncthreads();
        if (dX != nullptr) {   // dX[m][k] = sum_n dpre[m][n] W[k][n]: rows m = 4 ty + i, columns k = tx + 16 j
            float acc[4][6] = {};
#pragma unroll 2
            for (int n4 = 0; n4 < N / 4; ++n4) {
                float4 a[4], b[6];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(Ds + (ty * 4 + i) * PS + 4 * n4);
#pragma unroll
                for (int j = 0; j < 6; ++j) b[j] = *reinterpret_cast<const float4*>(Ws + (tx + 16 * j) * PS + 4 * n4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        float t = acc[i][j];
                        t = fmaf(a[i].x, b[j].x, t); t = fmaf(a[i].y, b[j].y, t); t = fmaf(a[i].z, b[j].z, t); t = fmaf(a[i].w, b[j].w, t);
                        acc[i][j] = t;
                    }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = bb + ty * 4 + i;
                if (m >= b_end) continue;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const int k = tx + 16 * j;
                    if (k < K) dX[(int64_t)m * lddx + k] = acc[i][j];
                }
            }
        }
        // dW[k][n] += sum_r X[r][k] dpre[r][n]: rows k = ty + 16 i, columns n = 4 tx .. 4 tx + 3 (the split kernel's mapping and order)
#pragma unroll 4
        for (int r = 0; r < PT; ++r) {
            const float4 d4 = *reinterpret_cast<const float4*>(Ds + r * PS + 4 * tx);
            const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const float a = Xs[r * kFusedXs + ty + 16 * i];
#pragma unroll
                for (int j = 0; j < 4; ++j) accw[i][j] = fmaf(a, d[j], accw[i][j]);
            }
        }
        __syncthreads();
    }
    float* out = partial + (int64_t)blockIdx.x * (K + 1) * N;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int k = ty + 16 * i;
        if (k <= K) *reinterpret_cast<float4*>(out + (int64_t)k * N + 4 * tx) = make_float4(accw[i][0], accw[i][1], accw[i][2], accw[i][3]);
    }
}

// dW / db = sum over chunks: four interleaved partial sums per element (chunks z = j mod 4, ascending), combined as
// ((s0 + s1) + (s2 + s3)) -- a fixed order, so the result is deterministic; four times shorter dependent chains than one sum
__global__ void __launch_bounds__(256) dense_bwd_reduce4_kernel(const float* __restrict__ partial, int nchunk, int K, int N, float* __restrict__ dW,
                                                                float* __restrict__ db) {
    const int64_t total = (int64_t)(K + 1) * N;
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t i = t >> 2;
    const int j = (int)(t & 3);
    float s = 0.f;
    if (i < total) {
#pragma unroll 8
        for (int z = j; z < nchunk; z += 4) s = __fadd_rn(s, __ldg(partial + (int64_t)z * total + i));
    }
    const float s1 = __shfl_xor_sync(0xffffffffu, s, 1);
    const float a = (j & 1) ? __fadd_rn(s1, s) : __fadd_rn(s, s1);     // lanes j=0,1 both hold s0 + s1; lanes 2,3 hold s2 + s3
    const float a2 = __shfl_xor_sync(0xffffffffu, a, 2);
    if (i < total && j == 0) {
        const float v = __fadd_rn(a, a2);
        const int k = (int)(i / N), n = (int)(i % N);
        if (k < K) dW[(int64_t)k * N + n] = v;
        else if (db) db[n] = v;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
bool panel_fwd_ok(int K, int N) { return K >= 1 && K <= kPanelMaxK && N >= 1; }
bool panel_bwd_ok(int K, int N) { return K >= 1 && K <= 1024 && N >= 1 && N <= kPanelMaxK; }

template <bool G, int PM>
static int launch_fwd_pm(const FeatArr& fa, const float* X, int ldx, const float* W, const float* b, float* Xout, float* Y, int ldy, float* Y32, int B,
                         int K, int N, int relu, cudaStream_t st) {
    const size_t smem = (size_t)K * (PM + 4 + PS) * sizeof(float);
    { static SmemAttr smem_attr; TT_CUDA_OK(smem_attr.ensure(dense_fwd_panel_kernel<G, PM>, (int)(2 * kPanelMaxK * PS * sizeof(float)))); }
    dim3 grid((unsigned)ceil_div(B, PM), (unsigned)ceil_div(N, PT));
    dense_fwd_panel_kernel<G, PM><<<grid, 256, smem, st>>>(fa, X, ldx, W, b, Xout, Y, ldy, Y32, B, K, N, relu);
    TT_LAUNCH_OK("dense_fwd_panel_kernel");
    return TT_OK;
}

template <bool G>
static int launch_fwd(const FeatArr& fa, const float* X, int ldx, const float* W, const float* b, float* Xout, float* Y, int ldy, float* Y32, int B,
                      int K, int N, int relu, cudaStream_t st) {
    const auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    bool fused = N == PT && K >= 1 && K <= kFwdKMax && ldy % 4 == 0 && al16(W) && al16(b) && al16(Y) && al16(Y32);
    if (fused && G) {
        int pieces = 0;
        for (int f = 0; f < fa.n; ++f) {
            const tt_feature& ft = fa.f[f];
            if (ft.table == nullptr) { pieces += 1; continue; }
            pieces += ft.e / 4;
            fused = fused && ft.e % 4 == 0 && (ft.shards > 1 || al16(ft.table));
        }
        fused = fused && pieces <= kFwdMaxPieces && (Xout == nullptr || (ldx % 4 == 0 && al16(Xout)));
    } else if (fused) {
        fused = ldx % 4 == 0 && ldx >= (K + 3) / 4 * 4 && al16(X);
    }
    if (fused) {
        dense_fwd_fused64_kernel<G><<<(unsigned)ceil_div(B, kFwdPM), 256, 0, st>>>(fa, X, ldx, W, b, Xout, Y, ldy, Y32, B, K, relu);
        TT_LAUNCH_OK("dense_fwd_fused64_kernel");
        return TT_OK;
    }
    // 32-row tiles while 64-row tiles would leave SMs with fewer than two CTAs
    if (ceil_div(B, PT) * ceil_div(N, PT) < 2 * (int64_t)sm_count())
        return launch_fwd_pm<G, 32>(fa, X, ldx, W, b, Xout, Y, ldy, Y32, B, K, N, relu, st);
    return launch_fwd_pm<G, 64>(fa, X, ldx, W, b, Xout, Y, ldy, Y32, B, K, N, relu, st);
}

int panel_dense_fwd(const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, float* Y32, int B, int K, int N, int relu,
                    cudaStream_t st) {
    FeatArr fa;
    memset(&fa, 0, sizeof(fa));
    return launch_fwd<false>(fa, X, ldx, W, b, nullptr, Y, ldy, Y32, B, K, N, relu, st);
}

int panel_input_dense_fwd(const tt_feature* feats, int nfeat, int D, const float* W, const float* b, float* Xout, int ldx, float* Y, int ldy,
                          float* Y32, int B, int N, int relu, cudaStream_t st) {
    FeatArr fa;
    memset(&fa, 0, sizeof(fa));
    fa.n = nfeat;
    for (int f = 0; f < nfeat; ++f) fa.f[f] = feats[f];
    return launch_fwd<true>(fa, nullptr, ldx, W, b, Xout, Y, ldy, Y32, B, D, N, relu, st);
}

static int dw_plan(int B, int* rows) {   // batch rows per CTA: a multiple of 64, about one wave of CTAs (the dX tiles share the launch)
    int64_t want = (int64_t)sm_count();
    int64_t r = ceil_div(ceil_div(B, want), PT) * PT;
    if (r < PT) r = PT;
    *rows = (int)r;
    return (int)ceil_div(B, r);
}

size_t panel_bwd_workspace(int B, int K, int N) {
    int rows = 0;
    int nchunk = dw_plan(B, &rows);
    return align_up((size_t)nchunk * (K + 1) * N * sizeof(float), 256) + 256;
}

// dX (optional) and dW/db in one launch + the fixed-order chunk reduction
int panel_dense_bwd(const float* X, int ldx, const float* W, const float* Y, int ldy, const float* dY, int lddy, float* dX, int lddx, float* dW,
                    float* db, float* partial, int B, int K, int N, int relu, cudaStream_t st) {
    int rows = 0;
    const int nchunk = dw_plan(B, &rows);
    const auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    if (N == PT && K + 1 <= kFusedKMax && ldx % 4 == 0 && ldx >= (K + 3) / 4 * 4 && ldy % 4 == 0 && lddy % 4 == 0 && al16(X) && al16(W) && al16(Y) && al16(dY) &&
        al16(partial)) {
        const size_t smem = (size_t)kFusedSmemFloats * sizeof(float);
        { static SmemAttr smem_attr; TT_CUDA_OK(smem_attr.ensure(dense_bwd_fused64_kernel, (int)smem)); }
        dense_bwd_fused64_kernel<<<(unsigned)nchunk, 256, smem, st>>>(X, ldx, W, Y, ldy, dY, lddy, dX, lddx, partial, B, K, relu, rows);
        TT_LAUNCH_OK("dense_bwd_fused64_kernel");
        const int64_t total = (int64_t)(K + 1) * N;
        dense_bwd_reduce4_kernel<<<(unsigned)ceil_div(total * 4, 256), 256, 0, st>>>(partial, nchunk, K, N, dW, db);
        TT_LAUNCH_OK("dense_bwd_reduce4_kernel");
        return TT_OK;
    }
    const int gx = (int)ceil_div(B, PT), gy = (int)ceil_div(K, PT);
    const int n_dx = dX ? gx * gy : 0;
    const int hy = (int)ceil_div(K + 1, kDwRows), hz = (int)ceil_div(N, PT);
    const size_t smem_dx = 2 * (size_t)N * PS * sizeof(float), smem_dw = (size_t)kDwSmemFloats * sizeof(float);
    const size_t smem = smem_dx > smem_dw ? smem_dx : smem_dw;
            const size_t mx = 2 * (size_t)kPanelMaxK * PS * sizeof(float);
        { static SmemAttr smem_attr; TT_CUDA_OK(smem_attr.ensure(dense_bwd_panel_kernel, (int)(mx > smem_dw ? mx : smem_dw))); }
    dense_bwd_panel_kernel<<<(unsigned)(n_dx + nchunk * hy * hz), 256, smem, st>>>(X, ldx, W, Y, ldy, dY, lddy, dX, lddx, partial, B, K, N, relu, rows,
                                                                                   n_dx, gx, nchunk, hy);
    TT_LAUNCH_OK("dense_bwd_panel_kernel");
    const int64_t total = (int64_t)(K + 1) * N;
    dense_bwd_reduce4_kernel<<<(unsigned)ceil_div(total * 4, 256), 256, 0, st>>>(partial, nchunk, K, N, dW, db);
    TT_LAUNCH_OK("dense_bwd_reduce4_kernel");
    return TT_OK;
}

}  // namespace tt
