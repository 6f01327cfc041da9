// Where do the last 5 % of the DMMA inner loop go?  (tools/wloop_bench.cu: 35.0-35.3 of 37.0 TFLOP/s however many warps.)
// The product loop shape — 16 x 32 warp tile, groups of 4 k4-blocks = 32 DMMA over 16 accumulators — with the operands
// coming from   MODE 0: registers only (distinct registers, loaded once)      MODE 1: B from shared memory (16 LDS.64 / group)
//               MODE 2: A from L2 (8 LDG.64 / group, one group ahead), B regs   MODE 3: both (= the product loop)
// ORDER 0: (j, nf, mf) as in the product; ORDER 1: (j, mf, nf) — four consecutive DMMAs share the A register pair.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/loop_bisect tools/loop_bisect.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
template <int MODE, int ORDER>
__global__ void __launch_bounds__(512) loop_kernel(const double* W, int Mp, int iters, double* out) {
    constexpr int NT = 32, NF = 4, STR = NT + 4;
    extern __shared__ double T[];
    for (int i = threadIdx.x; i < Mp * STR; i += blockDim.x) T[i] = 1e-3 * (i % 13);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int C4 = Mp / 4, nb16 = Mp / 16;
    const double* tb = T + t * STR + g;
    double acc[2][NF][2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
    double a0[4], a1[4], n0[4], n1[4], br[4][NF];
    const int b = warp % nb16;
    const double* w0 = W + ((size_t)(2 * b) * C4) * 32 + lane;
    const double* w1 = w0 + (size_t)C4 * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a0[j] = __ldg(w0 + j * 32); a1[j] = __ldg(w1 + j * 32);
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) br[j][nf] = tb[(size_t)j * 4 * STR + nf * 8];
    }
    for (int it = 0; it < iters; ++it) {
        for (int kb = 0; kb < C4; kb += 4) {
            if (MODE & 2) {
                const int nk = (kb + 4 < C4) ? kb + 4 : 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) { n0[j] = __ldg(w0 + (size_t)(nk + j) * 32); n1[j] = __ldg(w1 + (size_t)(nk + j) * 32); }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double bb[NF];
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) bb[nf] = (MODE & 1) ? tb[(size_t)(kb + j) * 4 * STR + nf * 8] : br[j][nf];
                if (ORDER == 0) {
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) { dmma(acc[0][nf], a0[j], bb[nf]); dmma(acc[1][nf], a1[j], bb[nf]); }
                } else {
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) dmma(acc[0][nf], a0[j], bb[nf]);
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) dmma(acc[1][nf], a1[j], bb[nf]);
                }
            }
            if (MODE & 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { a0[j] = n0[j]; a1[j] = n1[j]; }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) s += acc[mf][nf][0] + acc[mf][nf][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int ORDER>
static void run(const double* W, double* out, int sms, int threads) {
    const int Mp = 256, iters = 400;
    const size_t smem = (size_t)Mp * 36 * 8;
    CK(cudaFuncSetAttribute(loop_kernel<MODE, ORDER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    loop_kernel<MODE, ORDER><<<sms, threads, smem>>>(W, Mp, 20, out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        loop_kernel<MODE, ORDER><<<sms, threads, smem>>>(W, Mp, iters, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double flops = (double)sms * (threads / 32) * iters * (Mp / 16.0) * 32 * 512;
    printf("MODE %d (%s%s) ORDER %d  %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", MODE, (MODE & 2) ? "A:L2 " : "A:reg ", (MODE & 1) ? "B:smem" : "B:reg",
           ORDER, threads / 32, best, flops / best / 1e9);
}

// MODE 4: A through a PER-WARP shared-memory ring of S groups (2 KB each) fed by the warp's own bulk copies (TMA) S - 1
// groups ahead; B from shared memory.  No registers spent on prefetched fragments, L2 latency hidden S - 1 groups deep.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int S>
__global__ void __launch_bounds__(512) ring_kernel(const double* W, int Mp, int iters, double* out) {
    constexpr int NT = 32, NF = 4, STR = NT + 4;
    extern __shared__ __align__(16) double T[];
    const int nw = blockDim.x >> 5;
    double* rings = T + (size_t)Mp * STR;
    uint64_t* bars = reinterpret_cast<uint64_t*>(rings + (size_t)nw * S * 256);
    for (int i = threadIdx.x; i < Mp * STR; i += blockDim.x) T[i] = 1e-3 * (i % 13);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    double* ring = rings + (size_t)warp * S * 256;
    uint64_t* bar = bars + warp * S;
    if (lane == 0) { for (int i = 0; i < S; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
    __syncthreads();
    const int C4 = Mp / 4, nb16 = Mp / 16, G = C4 / 4;   // groups per row block
    const double* tb = T + t * STR + g;
    double acc[2][NF][2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
    const int b = warp % nb16;
    const double* w0 = W + ((size_t)(2 * b) * C4) * 32;
    const double* w1 = w0 + (size_t)C4 * 32;
    const int total = iters * G;
    // producer cursor (lane 0): next group to issue, its stage and k-offset; consumer cursor: group being multiplied
    int pi = 0, pst = 0, pkb = 0;
    auto issue = [&]() {   // lane 0
        mbar_expect(&bar[pst], 2048);
        bulk_g2s(ring + pst * 256, w0 + (size_t)pkb * 32, 1024, &bar[pst]);
        bulk_g2s(ring + pst * 256 + 128, w1 + (size_t)pkb * 32, 1024, &bar[pst]);
        ++pi; if (++pst == S) pst = 0; pkb += 4; if (pkb == C4) pkb = 0;
    };
    if (lane == 0) for (int i = 0; i < S; ++i) issue();
    int st = 0, kb = 0; unsigned ph = 0;       // stage / k-offset / parity of the group whose fragments are loaded NEXT
    double a0[4], a1[4], n0[4], n1[4];
    auto fetch = [&](double (&x0)[4], double (&x1)[4]) {   // wait for the stage, pull its fragments into registers
        mbar_wait(&bar[st], ph);
        const double* fr = ring + st * 256 + lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) { x0[j] = fr[j * 32]; x1[j] = fr[128 + j * 32]; }
    };
    auto advance = [&]() { if (++st == S) { st = 0; ph ^= 1u; } };
    auto mul = [&](const double (&x0)[4], const double (&x1)[4], int kbm) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) {
                const double bb = tb[(size_t)(kbm + j) * 4 * STR + nf * 8];
                dmma(acc[0][nf], x0[j], bb); dmma(acc[1][nf], x1[j], bb);
            }
        }
    };
    fetch(a0, a1); advance();
    for (int gi = 0; gi < total; gi += 2) {
        // group gi in (a0, a1): its stage is free again -> refill; fetch gi + 1 while gi multiplies
        __syncwarp();
        if (lane == 0 && pi < total) issue();
        if (gi + 1 < total) { fetch(n0, n1); advance(); }
        mul(a0, a1, kb); kb += 4; if (kb == C4) kb = 0;
        if (gi + 1 >= total) break;
        __syncwarp();
        if (lane == 0 && pi < total) issue();
        if (gi + 2 < total) { fetch(a0, a1); advance(); }
        mul(n0, n1, kb); kb += 4; if (kb == C4) kb = 0;
    }
    double s = 0;
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) s += acc[mf][nf][0] + acc[mf][nf][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int S>
static void run_ring(const double* W, double* out, int sms, int threads) {
    const int Mp = 256, iters = 400;
    const size_t smem = (size_t)Mp * 36 * 8 + (size_t)(threads / 32) * S * (2048 + 8);
    CK(cudaFuncSetAttribute(ring_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    ring_kernel<S><<<sms, threads, smem>>>(W, Mp, 20, out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        ring_kernel<S><<<sms, threads, smem>>>(W, Mp, iters, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double flops = (double)sms * (threads / 32) * iters * (Mp / 16.0) * 32 * 512;
    printf("MODE 4 (A: per-warp smem ring of %d groups fed by bulk copies, B:smem)  %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", S, threads / 32, best, flops / best / 1e9);
}

// MODE 5: A from L2, loaded PD groups ahead into PD + 1 rotating register sets (asm volatile loads: ptxas may not sink
// them to their first use), B from shared memory.
__device__ __forceinline__ double ldg_v(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];\n" : "=d"(v) : "l"(p));
    return v;
}
template <int PD>
__global__ void __launch_bounds__(256) deep_kernel(const double* W, int Mp, int iters, double* out) {
    constexpr int NT = 32, NF = 4, STR = NT + 4, NS = PD + 1;
    extern __shared__ double T[];
    for (int i = threadIdx.x; i < Mp * STR; i += blockDim.x) T[i] = 1e-3 * (i % 13);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int C4 = Mp / 4, nb16 = Mp / 16;
    const double* tb = T + t * STR + g;
    double acc[2][NF][2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
    const int b = warp % nb16;
    const double* w0 = W + ((size_t)(2 * b) * C4) * 32 + lane;
    const double* w1 = w0 + (size_t)C4 * 32;
    double a0[NS][4], a1[NS][4];
    int pk = 0;   // k4-block of the next group to load
    auto load = [&](double (&x0)[4], double (&x1)[4]) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { x0[j] = ldg_v(w0 + (size_t)(pk + j) * 32); x1[j] = ldg_v(w1 + (size_t)(pk + j) * 32); }
        pk += 4; if (pk == C4) pk = 0;
    };
#pragma unroll
    for (int s = 0; s < PD; ++s) load(a0[s], a1[s]);
    int kb = 0;
    const int total = iters * (C4 / 4);
    for (int gi = 0; gi < total; gi += NS) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            load(a0[(s + PD) % NS], a1[(s + PD) % NS]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) {
                    const double bb = tb[(size_t)(kb + j) * 4 * STR + nf * 8];
                    dmma(acc[0][nf], a0[s][j], bb); dmma(acc[1][nf], a1[s][j], bb);
                }
            }
            kb += 4; if (kb == C4) kb = 0;
        }
    }
    double sres = 0;
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) sres += acc[mf][nf][0] + acc[mf][nf][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sres;
}
template <int PD>
static void run_deep(const double* W, double* out, int sms) {
    const int Mp = 256, iters = 402, threads = 256;   // (402 * 16 groups: a multiple of 2, 3 and 4 register sets)
    const size_t smem = (size_t)Mp * 36 * 8;
    CK(cudaFuncSetAttribute(deep_kernel<PD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    deep_kernel<PD><<<sms, threads, smem>>>(W, Mp, 24, out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0));
        deep_kernel<PD><<<sms, threads, smem>>>(W, Mp, iters, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    const double flops = (double)sms * (threads / 32) * iters * (Mp / 16.0) * 32 * 512;
    printf("MODE 5 (A: L2, %d groups ahead in %d register sets, B:smem)   8 warps/SM: %.3f ms  %.2f TFLOP/s\n", PD, PD + 1, best, flops / best / 1e9);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double *W, *out;
    CK(cudaMalloc(&W, 256 * 256 * 8)); CK(cudaMemset(W, 0, 256 * 256 * 8));
    CK(cudaMalloc(&out, sizeof(double) * sms * 512));
    for (int threads : {256, 512}) {
        run<0, 0>(W, out, sms, threads); run<0, 1>(W, out, sms, threads);
        run<1, 0>(W, out, sms, threads); run<1, 1>(W, out, sms, threads);
        run<2, 0>(W, out, sms, threads); run<2, 1>(W, out, sms, threads);
        run<3, 0>(W, out, sms, threads); run<3, 1>(W, out, sms, threads);
    }
    run_deep<1>(W, out, sms); run_deep<2>(W, out, sms); run_deep<3>(W, out, sms);
    run_ring<2>(W, out, sms, 256); run_ring<3>(W, out, sms, 256); run_ring<4>(W, out, sms, 256); run_ring<6>(W, out, sms, 256);
    run_ring<4>(W, out, sms, 512);
    return 0;
}
