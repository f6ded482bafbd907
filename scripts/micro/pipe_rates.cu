// Pipe-rate micro-benchmarks behind the softmax epilogue design (development tool):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
// Per SM and clock: MUFU.EX2 (f32 and f16x2), FFMA2, F2FP pack, FMNMX3, and the epilogue's instruction mix
// (FFMA2 + 2 MUFU + FADD2 + F2FP per two logits) at 1, 2 and 4 warps per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { uint32_t r; asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

template <int MODE>
__global__ void k(int iters, float seed, float* out, long long* cyc) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = seed + 0.001f * (threadIdx.x + i);
    f32x2 acc0 = pk2(0.f, 0.f), acc1 = acc0;
    uint32_t hacc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {   // 16 independent MUFU.EX2
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = ex2(x[i]);
        } else if (MODE == 1) {   // 16 packed-half exponentials (32 results)
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = __uint_as_float(ex2h2(__float_as_uint(x[i])));
        } else if (MODE == 2) {   // 16 FFMA2 (32 results)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                f32x2 v = fma2(pk2(x[i], x[i + 1]), pk2(1.0001f, 0.9999f), pk2(0.5f, 0.25f));
                f32x2 u = fma2(v, pk2(0.9999f, 1.0001f), pk2(-0.5f, -0.25f));
                upk2(u, x[i], x[i + 1]);
            }
        } else if (MODE == 3) {   // 8 F2FP packs (16 inputs)
#pragma unroll
            for (int i = 0; i < 16; i += 2) hacc ^= pack(x[i], x[i + 1]) + it;
        } else if (MODE == 4) {   // 8 FMNMX3
#pragma unroll
            for (int i = 0; i < 16; i += 2) x[i] = fmin3(x[i], x[i + 1], seed + it);
        } else if (MODE == 5) {   // the pass-1 epilogue mix per 16 logits: 8 FFMA2 (zn) + 8 FMNMX3 + 8 FFMA2 (x) + 16 MUFU + 8 FADD2 + 8 F2FP
            f32x2 zn[8];
            float mn = 1e30f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                zn[i] = fma2(pk2(x[2 * i], x[2 * i + 1]), pk2(-1.44f, -1.44f), pk2(0.1f * i, 0.2f * i));
                float a, b; upk2(zn[i], a, b);
                mn = fmin3(mn, a, b);
            }
            const f32x2 aa = pk2(mn, mn), mone = pk2(-1.f, -1.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float a, b; upk2(fma2(zn[i], mone, aa), a, b);
                const float p0 = ex2(a), p1 = ex2(b);
                if (i & 1) acc1 = add2(acc1, pk2(p0, p1)); else acc0 = add2(acc0, pk2(p0, p1));
                hacc ^= pack(p0, p1);
                x[2 * i] = p0 * 1e-3f; x[2 * i + 1] = p1 * 1e-3f;   // (keeps the chain alive; 16 FMUL extra)
            }
        } else if (MODE == 6) {   // the pass-2 mix per 16 logits: 8 FFMA2 (addend) + 8 FFMA2 + 16 MUFU + 8 F2FP
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                f32x2 ad = fma2(pk2(0.1f * i, 0.2f * i), pk2(-1.f, -1.f), pk2(seed, seed));
                float a, b; upk2(fma2(pk2(x[2 * i], x[2 * i + 1]), pk2(1.44f, 1.44f), ad), a, b);
                const float p0 = ex2(a), p1 = ex2(b);
                hacc ^= pack(p0, p1);
                x[2 * i] = p0 * 1e-3f; x[2 * i + 1] = p1 * 1e-3f;
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    float a, b; upk2(add2(acc0, acc1), a, b);
    if (s + a + b == 12345.678f) out[0] = s + hacc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
    float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    const char* names[] = {"MUFU.EX2 f32", "MUFU.EX2 f16x2 (results)", "FFMA2 (results)", "F2FP (inputs)", "FMNMX3 (inputs)", "pass-1 mix (logits)",
                           "pass-2 mix (logits)"};
    const double per_iter[] = {16, 32, 32, 16, 16, 16, 16};
    for (int mode = 0; mode < 7; ++mode)
        for (int warps : {4, 8, 16}) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                switch (mode) {
                    case 0: k<0><<<148, warps * 32>>>(iters, -0.5f, out, cyc); break;
                    case 1: k<1><<<148, warps * 32>>>(iters, -0.5f, out, cyc); break;
                    case 2: k<2><<<148, warps * 32>>>(iters, -0.5f, out, cyc); break;
                    case 3: k<3><<<148, warps * 32>>>(iters, -0.5f, out, cyc); break;
                    case 4: k<4><<<148, warps * 32>>>(iters, -0.5f, out, cyc); break;
                    case 5: k<5><<<148, warps * 32>>>(iters, -0.5f, out, cyc); break;
                    default: k<6><<<148, warps * 32>>>(iters, -0.5f, out, cyc); break;
                }
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-28s warps/SM %2d: %8lld cycles  %7.2f per clk per SM   err=%s\n", names[mode], warps, h,
                   per_iter[mode] * iters * warps * 32 / (double)h, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
