// tt_peer.cu -- exchanges between the one-process-per-GPU ranks done by our own kernels over NVLink peer memory
// (no reference counterpart: the reference is single-process).
//
// With row-sharded embedding tables every bulk transfer of a data-parallel step is a peer read inside the kernel that
// consumes the data (table rows in the gather, batch ids in the radix sort, gradient rows in the staging kernel, dense
// gradients in peer_sum_kernel).  What is left of the collectives is ordering: "every rank has finished phase X".
// peer_barrier_kernel provides it with one flag store per peer and a bounded spin on local memory, so a whole training
// step is ONE CUDA graph with no NCCL call inside.
#include "tt_common.cuh"

namespace tt {

// Flag block of a rank (uint32, in its own peer-shareable memory): [slot] local epoch counters (TT_PEER_SLOTS of them), then
// arrive[slot][src rank] = last epoch at which `src` reached barrier `slot`.
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(const unsigned long long* __restrict__ blocks, int rank, int world, int slot,
                                                          long long max_cycles) {
    uint32_t* mine = reinterpret_cast<uint32_t*>(blocks[rank]);
    const uint32_t epoch = mine[slot] + 1;          // every lane reads it before lane 0 bumps it (after the __syncwarp below)
    const int t = threadIdx.x;
    if (t < world) {
        __threadfence_system();
        uint32_t* theirs = reinterpret_cast<uint32_t*>(blocks[t]);
        st_release_sys(theirs + TT_PEER_SLOTS + slot * world + rank, epoch);            // "rank has arrived", written into peer t's block
        const uint32_t* src = mine + TT_PEER_SLOTS + slot * world + t;                   // wait for peer t's arrival in MY block (local spin)
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(src) - epoch) < 0) {
            if (clock64() - t0 > max_cycles) {     // a rank died or the ranks left lockstep -- fail loudly, do not hang
                printf("tt_peer_barrier: rank %d timed out waiting for rank %d (slot %d, epoch %u)\n", rank, t, slot, epoch);
                __trap();
            }
            __nanosleep(64);
        }
        __threadfence_system();
    }
    __syncwarp();
    if (t == 0) mine[slot] = epoch;
}

// ring[0] = number of completed steps; ring[1 + (step % ring_len) * nk + k] = %globaltimer (ns) when stamp k of that step ran
__global__ void __launch_bounds__(32) stamp_kernel(unsigned long long* __restrict__ ring, int ring_len, int k, int nk) {
    if (threadIdx.x != 0) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long step = ring[0];
    ring[1 + (step % (unsigned long long)ring_len) * nk + k] = t;
    if (k == nk - 1) ring[0] = step + 1;
}

// out[i] = src[0][i] + src[1][i] + ... in rank order (fixed order: every rank computes the same bits)
__global__ void __launch_bounds__(256) peer_sum_kernel(const unsigned long long* __restrict__ srcs, int world, int64_t n, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        float acc = __ldg(reinterpret_cast<const float*>(srcs[0]) + i);
        for (int r = 1; r < world; ++r) acc = __fadd_rn(acc, __ldg(reinterpret_cast<const float*>(srcs[r]) + i));
        out[i] = acc;
    }
}

// out[r][i] = src_r[i] for every rank r: all-gather by peer reads (16-byte loads when n is a multiple of 4 floats)
__global__ void __launch_bounds__(256) peer_gather_kernel(const unsigned long long* __restrict__ srcs, int world, int64_t n4, float4* __restrict__ out) {
    const int64_t total = n4 * world;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += stride) {
        const int r = (int)(i / n4);
        out[i] = __ldg(reinterpret_cast<const float4*>(srcs[r]) + (i - (int64_t)r * n4));
    }
}

}  // namespace tt

using namespace tt;

extern "C" {

int tt_peer_barrier(const void* flag_blocks, int rank, int world, int slot, void* stream) {
    TT_REQUIRE(flag_blocks != nullptr, "tt_peer_barrier: null pointer");
    TT_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world && slot >= 0 && slot < TT_PEER_SLOTS, "tt_peer_barrier: bad rank/world/slot");
    // how long a rank may wait for its peers before the kernel traps: TT_PEER_TIMEOUT_S seconds (default 120; a host-side stall of
    // one rank -- loading a file, a checkpoint, the first eager step -- must stay below it), counted at ~2 GHz
    static const long long max_cycles = [] {
        const char* e = getenv("TT_PEER_TIMEOUT_S");
        double sec = e ? atof(e) : 120.0;
        if (!(sec >= 1.0)) sec = 1.0;
        return (long long)(sec * 2.0e9);
    }();
    peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned long long*>(flag_blocks), rank, world, slot, max_cycles);
    TT_LAUNCH_OK("peer_barrier_kernel");
    return TT_OK;
}

int tt_stamp(uint64_t* ring, int ring_len, int k, int nk, void* stream) {
    TT_REQUIRE(ring != nullptr && ring_len >= 1 && nk >= 1 && k >= 0 && k < nk, "tt_stamp: bad arguments");
    stamp_kernel<<<1, 32, 0, as_stream(stream)>>>(reinterpret_cast<unsigned long long*>(ring), ring_len, k, nk);
    TT_LAUNCH_OK("stamp_kernel");
    return TT_OK;
}

int tt_peer_sum_f32(const void* src_ptrs, int world, int64_t n, float* out, void* stream) {
    TT_REQUIRE(src_ptrs != nullptr && out != nullptr && world >= 1 && n >= 0, "tt_peer_sum_f32: bad arguments");
    if (n == 0) return TT_OK;
    int64_t g = ceil_div(n, 256);
    if (g > 4 * (int64_t)sm_count()) g = 4 * (int64_t)sm_count();
    peer_sum_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned long long*>(src_ptrs), world, n, out);
    TT_LAUNCH_OK("peer_sum_kernel");
    return TT_OK;
}

int tt_peer_gather_f32(const void* src_ptrs, int world, int64_t n, float* out, void* stream) {
    TT_REQUIRE(src_ptrs != nullptr && out != nullptr && world >= 1 && n >= 0, "tt_peer_gather_f32: bad arguments");
    TT_REQUIRE(n % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "tt_peer_gather_f32: n must be a multiple of 4 floats, buffers 16-byte aligned");
    if (n == 0) return TT_OK;
    int64_t g = ceil_div(n / 4 * world, 256);
    if (g > 8 * (int64_t)sm_count()) g = 8 * (int64_t)sm_count();
    peer_gather_kernel<<<(unsigned)g, 256, 0, as_stream(stream)>>>(reinterpret_cast<const unsigned long long*>(src_ptrs), world, n / 4,
                                                                     reinterpret_cast<float4*>(out));
    TT_LAUNCH_OK("peer_gather_kernel");
    return TT_OK;
}

}  // extern "C"
