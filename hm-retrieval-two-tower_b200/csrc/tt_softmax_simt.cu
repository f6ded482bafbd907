// tt_softmax_simt.cu -- exact fp32 (CUDA-core) in-batch sampled softmax: the reference-precision
// path and the cross-check for the tcgen05 kernels in tt_softmax_tc.cu.
//
// Replaces (reference file:line): two_tower_model.py:90-92 (Q.C^T), logq_correction.py:66-71
// (Z = S - ln p_j), two_tower_model.py:119-122 + runner.py:78-83 (labels = eye, CE from logits,
// reduction SUM) and their autodiff (dZ = softmax(Z) - I, dQ = dZ.C, dC = dZ^T.Q).
// The (Bq x Bc) matrix lives only in registers / shared memory, one 64 x 64 tile at a time.
#include <math_constants.h>

#include "tt_common.cuh"
#include "tt_simt_gemm.cuh"

namespace tt {

struct RowsOf {  // element (m, k) of a row-major matrix
    const float* P;
    int ld, rows, cols;
    __device__ __forceinline__ float operator()(int m, int k) const {
        return (m < rows && k < cols) ? __ldg(P + (int64_t)m * ld + k) : 0.f;
    }
};
struct TransOf {  // element (k, n) of P^T, i.e. P[n][k]
    const float* P;
    int ld, rows, cols;
    __device__ __forceinline__ float operator()(int k, int n) const {
        return (n < rows && k < cols) ? __ldg(P + (int64_t)n * ld + k) : 0.f;
    }
};

// reductions across the 16 threads (tx = 0..15) that share one group of 4 rows
__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- forward: lse_i and rowloss_i = lse_i - z_{i,i+off} --------------------------------------------
__global__ void __launch_bounds__(256) softmax_fwd_simt_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ C, int ldc,
                                                               const float* __restrict__ bias, int Bq, int Bc, int E, int off,
                                                               float* __restrict__ lse, float* __restrict__ rowloss) {
    __shared__ TileSmem sm;
    const int m0 = blockIdx.x * BM;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    RowsOf la{Q, ldq, Bq, E};
    TransOf lb{C, ldc, Bc, E};
    float run_m[TM], run_l[TM], zdiag[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) { run_m[i] = -CUDART_INF_F; run_l[i] = 0.f; zdiag[i] = 0.f; }
    for (int n0 = 0; n0 < Bc; n0 += BN) {
        float acc[TM][TN] = {};
        tile_gemm<true, false>(acc, la, lb, m0, n0, 0, E, sm);
        float bj[TN];
        bool ok[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            ok[j] = n < Bc;
            bj[j] = (ok[j] && bias) ? __ldg(bias + n) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int m = m0 + ty * TM + i;
            float z[TN];
            float tmax = -CUDART_INF_F;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                z[j] = ok[j] ? __fsub_rn(acc[i][j], bj[j]) : -CUDART_INF_F;
                tmax = fmaxf(tmax, z[j]);
                if (ok[j] && (n0 + tx * TN + j) == m + off) zdiag[i] = z[j];
            }
            tmax = group16_max(tmax);
            float new_m = fmaxf(run_m[i], tmax);
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < TN; ++j) part += ok[j] ? expf(z[j] - new_m) : 0.f;
            part = group16_sum(part);
            float scale = (run_m[i] == -CUDART_INF_F) ? 0.f : expf(run_m[i] - new_m);
            run_l[i] = run_l[i] * scale + part;
            run_m[i] = new_m;
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        float zd = group16_sum(zdiag[i]);  // exactly one thread of the group holds it
        if (tx == 0 && m < Bq) {
            float l = run_m[i] + logf(run_l[i]);
            lse[m] = l;
            rowloss[m] = l - zd;
        }
    }
}

// fixed-order (tree) sum of n floats into out[0]: single block, deterministic
__global__ void __launch_bounds__(1024) sum_rows_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
    __shared__ double s[1024];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) acc += (double)v[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = (float)s[0];
}

// ---- backward pass: G[r,:] = sum_c (exp(s_rc - rowv_r - colv_c) - [c == r + d]) * T[c,:] -----------
//   pass A (dQ): R = Q, T = C, rowv = lse,  colv = bias, d = +off
//   pass B (dC): R = C, T = Q, rowv = bias, colv = lse,  d = -off
template <int NE>
__global__ void __launch_bounds__(256) softmax_bwd_simt_kernel(const float* __restrict__ R, int ldr, const float* __restrict__ T, int ldt,
                                                               const float* __restrict__ rowv, const float* __restrict__ colv, int nR,
                                                               int nT, int E, int d, float* __restrict__ G, int ldg) {
    __shared__ TileSmem sm;
    __shared__ float Ps[BM][BN + PAD];
    const int m0 = blockIdx.x * BM;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    RowsOf la{R, ldr, nR, E};
    TransOf lb{T, ldt, nT, E};
    float rv[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        rv[i] = (m < nR && rowv) ? __ldg(rowv + m) : 0.f;
    }
    float g[NE][TM][TN] = {};
    for (int n0 = 0; n0 < nT; n0 += BN) {
        float acc[TM][TN] = {};
        tile_gemm<true, false>(acc, la, lb, m0, n0, 0, E, sm);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            bool ok = n < nT;
            float cv = (ok && colv) ? __ldg(colv + n) : 0.f;
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                int m = m0 + ty * TM + i;
                float p = 0.f;
                if (ok && m < nR) {
                    p = expf(acc[i][j] - rv[i] - cv);
                    if (n == m + d) p -= 1.0f;
                }
                Ps[ty * TM + i][tx * TN + j] = p;
            }
        }
        __syncthreads();
        auto pa = [&](int m, int kk) -> float { return Ps[m - m0][kk - n0]; };
        auto tb = [&](int kk, int n) -> float { return (kk < nT && n < E) ? __ldg(T + (int64_t)kk * ldt + n) : 0.f; };
#pragma unroll
        for (int e = 0; e < NE; ++e) tile_gemm<true, true>(g[e], pa, tb, m0, e * BN, n0, n0 + BN, sm);
    }
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int m = m0 + ty * TM + i;
            if (m >= nR) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                int n = e * BN + tx * TN + j;
                if (n < E) G[(int64_t)m * ldg + n] = g[e][i][j];
            }
        }
}

__global__ void __launch_bounds__(256) logits_simt_kernel(const float* __restrict__ Q, int ldq, const float* __restrict__ C, int ldc,
                                                          const float* __restrict__ bias, int Bq, int Bc, int E, float* __restrict__ Z,
                                                          int ldz) {
    __shared__ TileSmem sm;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    RowsOf la{Q, ldq, Bq, E};
    TransOf lb{C, ldc, Bc, E};
    float acc[TM][TN] = {};
    tile_gemm<true, false>(acc, la, lb, m0, n0, 0, E, sm);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= Bq) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int n = n0 + tx * TN + j;
            if (n < Bc) Z[(int64_t)m * ldz + n] = bias ? __fsub_rn(acc[i][j], __ldg(bias + n)) : acc[i][j];
        }
    }
}

int sum_rows_launch(const float* v, int n, float* out, cudaStream_t st) {
    sum_rows_kernel<<<1, 1024, 0, st>>>(v, n, out);
    TT_LAUNCH_OK("sum_rows_kernel");
    return TT_OK;
}

// ---- host-side entry points used by the dispatchers in tt_softmax.cu -------------------------------
int softmax_fwd_simt(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse,
                     float* loss, float* rowloss, cudaStream_t st) {
    softmax_fwd_simt_kernel<<<(unsigned)ceil_div(Bq, BM), 256, 0, st>>>(Q, ldq, C, ldc, bias, Bq, Bc, E, off, lse, rowloss);
    TT_LAUNCH_OK("softmax_fwd_simt_kernel");
    sum_rows_kernel<<<1, 1024, 0, st>>>(rowloss, Bq, loss);
    TT_LAUNCH_OK("sum_rows_kernel");
    return TT_OK;
}

template <int NE>
static int launch_bwd(const float* R, int ldr, const float* T, int ldt, const float* rowv, const float* colv, int nR, int nT, int E, int d,
                      float* G, int ldg, cudaStream_t st) {
    softmax_bwd_simt_kernel<NE><<<(unsigned)ceil_div(nR, BM), 256, 0, st>>>(R, ldr, T, ldt, rowv, colv, nR, nT, E, d, G, ldg);
    TT_LAUNCH_OK("softmax_bwd_simt_kernel");
    return TT_OK;
}

int softmax_bwd_pass_simt(const float* R, int ldr, const float* T, int ldt, const float* rowv, const float* colv, int nR, int nT, int E,
                          int d, float* G, int ldg, cudaStream_t st) {
    if (nR == 0) return TT_OK;
    int ne = (int)ceil_div(E, BN);
    switch (ne) {
        case 1: return launch_bwd<1>(R, ldr, T, ldt, rowv, colv, nR, nT, E, d, G, ldg, st);
        case 2: return launch_bwd<2>(R, ldr, T, ldt, rowv, colv, nR, nT, E, d, G, ldg, st);
        case 3: return launch_bwd<3>(R, ldr, T, ldt, rowv, colv, nR, nT, E, d, G, ldg, st);
        case 4: return launch_bwd<4>(R, ldr, T, ldt, rowv, colv, nR, nT, E, d, G, ldg, st);
        default: set_error("softmax bwd (SIMT): E=%d > 256 unsupported", E); return TT_ERR_UNSUPPORTED;
    }
}

int logits_simt(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, float* Z, int ldz,
                cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(Bq, BM), (unsigned)ceil_div(Bc, BN));
    logits_simt_kernel<<<grid, 256, 0, st>>>(Q, ldq, C, ldc, bias, Bq, Bc, E, Z, ldz);
    TT_LAUNCH_OK("logits_simt_kernel");
    return TT_OK;
}

}  // namespace tt
