// Shared device helpers for libmgp (sm_100a).  FP64 tensor-core path = mma.sync m8n8k4 f64, which
// lowers to DMMA.8x8x4 on sm_100a (tcgen05 has no f64 kind).  Measured issue-rate peak on B200:
// 37.0 TFLOP/s (profiles/r01_fp64_peak_microbench.txt).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mgp {

// ---- fragment conventions (PTX ISA, mma.m8n8k4 .f64) -------------------------------------------
//   lane = g*4 + t,  g = lane>>2 (0..7), t = lane&3 (0..3)
//   A (8x4, row)  : a  = A[g][t]
//   B (4x8, col)  : b  = B[t][g]           (k = t, n = g)
//   C/D (8x8)     : c0 = C[g][2t], c1 = C[g][2t+1]
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// ---- "fragment-major" layout of a left operand W [R x C] (R % 8 == 0, C % 4 == 0) ----------------
// 8x4 blocks are stored contiguously (32 doubles, element (g,t) at g*4+t), block (rb, kb) at
// (rb*(C/4) + kb)*32.  A warp reads one A-fragment with a single fully coalesced 256-byte load.
__host__ __device__ __forceinline__ size_t wf_index(int r, int c, int C) {
    return ((size_t)(r >> 3) * (size_t)(C >> 2) + (size_t)(c >> 2)) * 32 + (size_t)((r & 7) * 4 + (c & 3));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the 8 "g" groups (lanes with equal t): xor 4, 8, 16
__device__ __forceinline__ double sum_over_g(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

// Sums of 8 per-lane values over the 8 "g" groups with 7 shuffles instead of 24: at each butterfly level a lane hands
// over the half of the values it will not keep.  Lane (g, t) returns the total of v[g] over all g' (same t).
__device__ __forceinline__ double reduce8_over_g(const double (&v)[8], int lane) {
    const bool b2 = (lane >> 4) & 1, b1 = (lane >> 3) & 1, b0 = (lane >> 2) & 1;
    double w[4], x[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double keep = b2 ? v[i + 4] : v[i], send = b2 ? v[i] : v[i + 4];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double keep = b1 ? w[i + 2] : w[i], send = b1 ? w[i] : w[i + 2];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const double keep = b0 ? x[1] : x[0], send = b0 ? x[0] : x[1];
    return keep + __shfl_xor_sync(0xffffffffu, send, 4);
}

// sum over the 4 "t" lanes of a group: xor 1, 2
__device__ __forceinline__ double sum_over_t(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- mbarrier + bulk-copy (TMA, UBLKCP) pipeline primitives ---------------------------------------
// Producer/consumer rings: a producer warp issues cp.async.bulk row copies that complete on a "full" mbarrier
// (transaction bytes); consumer warps wait on it, multiply, and arrive on the stage's "empty" mbarrier.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// arrive and report how many arrivals the phase was still waiting for BEFORE this one (1: this was the last)
__device__ __forceinline__ unsigned mbar_arrive_pending(uint64_t* bar) {
    unsigned pending;
    asm volatile(
        "{\n"
        ".reg .b64 st;\n"
        "mbarrier.arrive.shared::cta.b64 st, [%1];\n"
        "mbarrier.pending_count.b64 %0, st;\n"
        "}\n"
        : "=r"(pending)
        : "r"(smem_u32(bar))
        : "memory");
    return pending;
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned), completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// shared -> global bulk copy of this thread's bulk-async group (the shared-memory source must have been made visible to
// the async proxy: writers execute fence_proxy_async() before the barrier that precedes the copy)
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
// the issuing thread's bulk stores have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// pull `bytes` (multiple of 16) at a 16-byte aligned global address into L2
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// deterministic block-wide sum of one double per thread; result valid in thread 0.  `red` >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = lane < nw ? red[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

}  // namespace mgp
