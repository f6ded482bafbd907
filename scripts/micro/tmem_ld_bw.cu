// TMEM read bandwidth micro-benchmark (development tool): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr) {
    uint32_t acc = 0;
    if constexpr (X == 32) {
        uint32_t r[32];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
                       "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= r[i];
    } else {
        uint32_t r[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= r[i];
    }
    return acc;
}
// two loads in flight before the wait
__device__ __forceinline__ uint32_t ld2x32(uint32_t t0, uint32_t t1) {
    uint32_t r[64];
#define LD32(base, T) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
                     : "=r"(r[base+0]),"=r"(r[base+1]),"=r"(r[base+2]),"=r"(r[base+3]),"=r"(r[base+4]),"=r"(r[base+5]),"=r"(r[base+6]),"=r"(r[base+7]),"=r"(r[base+8]),"=r"(r[base+9]),"=r"(r[base+10]),"=r"(r[base+11]),"=r"(r[base+12]),"=r"(r[base+13]),"=r"(r[base+14]),"=r"(r[base+15]), \
                       "=r"(r[base+16]),"=r"(r[base+17]),"=r"(r[base+18]),"=r"(r[base+19]),"=r"(r[base+20]),"=r"(r[base+21]),"=r"(r[base+22]),"=r"(r[base+23]),"=r"(r[base+24]),"=r"(r[base+25]),"=r"(r[base+26]),"=r"(r[base+27]),"=r"(r[base+28]),"=r"(r[base+29]),"=r"(r[base+30]),"=r"(r[base+31]) \
                     : "r"(T) : "memory")
    LD32(0, t0); LD32(32, t1);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) acc ^= r[i];
    return acc;
}
template <int MODE>
__global__ void k(int iters, uint32_t* out, long long* cyc) {
    __shared__ uint32_t holder;
    int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t tmem = holder;
    uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        uint32_t col = ((i * 64) + (warp >> 2) * 32) & 511;
        if (MODE == 0) acc ^= ld<32>(tmem + lane_addr + (col & 480));
        else if (MODE == 1) acc ^= ld<16>(tmem + lane_addr + (col & 496));
        else acc ^= ld2x32(tmem + lane_addr + (col & 448), tmem + lane_addr + ((col & 448) + 32));
    }
    long long t1 = clock64();
    if (acc == 0x12345u) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}
int main() {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {4, 8, 16}) {
            long long h = 0;
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<148, warps * 32>>>(iters, out, cyc);
                else if (mode == 1) k<1><<<148, warps * 32>>>(iters, out, cyc);
                else k<2><<<148, warps * 32>>>(iters, out, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double bytes = (double)iters * warps * 32 * 4 * (mode == 0 ? 32 : mode == 1 ? 16 : 64);
            printf("mode %d (%s) warps %2d: %lld cycles, %.1f B/clk/SM  err=%s\n", mode, mode == 0 ? "x32+wait" : mode == 1 ? "x16+wait" : "2*x32+wait", warps, h,
                   bytes / h, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
