// tt_simt_gemm.cuh -- exact fp32 tile GEMM on the CUDA cores (canonical k-ascending fmaf order).
#pragma once
#include "tt_common.cuh"

namespace tt {

// ------------------------------------------------------------------------------------------------
// Generic 64x64x16 SIMT tile GEMM with functor operand loaders.  256 threads, 4x4 outputs each.
// Accumulation is strictly k-ascending with one fmaf per term (the canonical order).
// ------------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4, PAD = 4;

struct TileSmem {
    float As[BK][BM + PAD];
    float Bs[BK][BN + PAD];
};

// LA: float a(int m, int k) -- element of the (M x K) left operand, 0 outside bounds.
// LB: float b(int k, int n) -- element of the (K x N) right operand, 0 outside bounds.
// kAContigK / kBContigN choose the thread->element mapping so global loads coalesce.
template <bool kAContigK, bool kBContigN, class LA, class LB>
__device__ __forceinline__ void tile_gemm(float (&acc)[TM][TN], const LA& la, const LB& lb, int m0, int n0, int kbeg, int kend,
                                          TileSmem& sm) {
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
        for (int i = 0; i < (BM * BK) / 256; ++i) {
            int e = tid + i * 256;
            int m, k;
            if (kAContigK) { m = e / BK; k = e % BK; } else { k = e / BM; m = e % BM; }
            float v = (k0 + k < kend) ? la(m0 + m, k0 + k) : 0.f;
            sm.As[k][m] = v;
        }
#pragma unroll
        for (int i = 0; i < (BN * BK) / 256; ++i) {
            int e = tid + i * 256;
            int n, k;
            if (kBContigN) { k = e / BN; n = e % BN; } else { n = e / BK; k = e % BK; }
            float v = (k0 + k < kend) ? lb(k0 + k, n0 + n) : 0.f;
            sm.Bs[k][n] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a4 = *reinterpret_cast<const float4*>(&sm.As[k][ty * TM]);
            float4 b4 = *reinterpret_cast<const float4*>(&sm.Bs[k][tx * TN]);
            float a[TM] = {a4.x, a4.y, a4.z, a4.w};
            float b[TN] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
}


}  // namespace tt
