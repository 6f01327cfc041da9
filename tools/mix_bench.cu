// What does a scalar FP64 instruction cost next to a saturated DMMA stream on sm_100a?
// Each warp loops over { 32 independent DMMA.8x8x4 ; N scalar instructions of one kind }, 2 or 4 warps per SM
// sub-partition, and the marginal cost per scalar warp-instruction is reported in sub-partition clocks:
//     cost = (t(N) - t(0)) * f_clk / (iters * N * warps_per_subpartition)
// Kinds: DFMA (independent chains), DADD, DMUL, FFMA, IMAD, and "DFMA issued by OTHER warps" (half the warps only
// multiply, the other half only run DFMAs) to separate pipe sharing from in-warp issue effects.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mix_bench tools/mix_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

enum { K_DFMA = 0, K_DADD = 1, K_DMUL = 2, K_FFMA = 3, K_IMAD = 4 };

template <int KIND, int N>
__device__ __forceinline__ void scalar_ops(double (&s)[8], float (&f)[8], int (&q)[8], double a, double b) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        if (KIND == K_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;\n" : "+d"(s[i & 7]) : "d"(a), "d"(b));
        if (KIND == K_DADD) asm volatile("add.rn.f64 %0, %0, %1;\n" : "+d"(s[i & 7]) : "d"(b));
        if (KIND == K_DMUL) asm volatile("mul.rn.f64 %0, %0, %1;\n" : "+d"(s[i & 7]) : "d"(a));
        if (KIND == K_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;\n" : "+f"(f[i & 7]) : "f"((float)a), "f"((float)b));
        if (KIND == K_IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;\n" : "+r"(q[i & 7]) : "r"(q[(i + 1) & 7] | 3), "r"(i + 1));
    }
}

// SPLIT: warps with odd index only run the scalar instructions (32 * N / 32 ... the same N per iteration), even warps
// only the DMMAs
template <int KIND, int N, bool SPLIT>
__global__ void __launch_bounds__(512) mix_kernel(double* out, int iters, double a, double b) {
    double c0[16], c1[16], s[8];
    float f[8];
    int q[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 1.0 + i * 1e-3; f[i] = 1.0f + i; q[i] = threadIdx.x + i; }
    const int warp = threadIdx.x >> 5;
    // warps of one sub-partition: warp % 4 is the sub-partition; (warp / 4) & 1 says which role in SPLIT mode
    const bool mma_role = !SPLIT || (((warp >> 2) & 1) == 0);
    const bool sc_role = !SPLIT || (((warp >> 2) & 1) == 1);
    for (int it = 0; it < iters; ++it) {
        if (mma_role) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) dmma884(c0[i], c1[i], a, b);
        }
        if (sc_role) scalar_ops<KIND, N>(s, f, q, a, b);
    }
    double t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i] + f[i] + q[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

static double g_clk_khz;
static int g_sms;

template <int KIND, int N, bool SPLIT>
static double run(double* out, int threads, int iters) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    mix_kernel<KIND, N, SPLIT><<<g_sms, threads>>>(out, iters / 10, 1.0000001, 1e-9);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        mix_kernel<KIND, N, SPLIT><<<g_sms, threads>>>(out, iters, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

template <int KIND, bool SPLIT>
static void sweep(const char* name, double* out, int threads, int iters) {
    const int wps = threads / 32 / 4;                       // warps per sub-partition
    const int mma_wps = SPLIT ? wps / 2 : wps, sc_wps = SPLIT ? wps / 2 : wps;
    const double t0 = run<KIND, 0, SPLIT>(out, threads, iters);
    const double ideal = (double)iters * 32 * 16 * mma_wps / (g_clk_khz * 1e3) * 1e3;   // ms at 16 clocks per DMMA
    printf("%-22s warps/SP=%d  N=0: %.3f ms (DMMA pipe %.1f %%)", name, wps, t0, 100.0 * ideal / t0);
    const double ts[4] = {run<KIND, 4, SPLIT>(out, threads, iters), run<KIND, 8, SPLIT>(out, threads, iters),
                          run<KIND, 16, SPLIT>(out, threads, iters), run<KIND, 32, SPLIT>(out, threads, iters)};
    const int ns[4] = {4, 8, 16, 32};
    for (int i = 0; i < 4; ++i) {
        const double clk = (ts[i] - t0) * 1e-3 * g_clk_khz * 1e3 / ((double)iters * ns[i] * sc_wps);
        printf("  N=%d: %.3f ms (+%.1f clk/instr)", ns[i], ts[i], clk);
    }
    printf("\n");
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0;
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    g_clk_khz = clk; g_sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, g_sms, clk);
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * g_sms * 512));
    const int iters = 20000;
    for (int threads : {256, 512}) {
        sweep<K_DFMA, false>("DFMA in-warp", out, threads, iters);
        sweep<K_DADD, false>("DADD in-warp", out, threads, iters);
        sweep<K_DMUL, false>("DMUL in-warp", out, threads, iters);
        sweep<K_FFMA, false>("FFMA in-warp", out, threads, iters);
        sweep<K_IMAD, false>("IMAD in-warp", out, threads, iters);
        sweep<K_DFMA, true>("DFMA other warps", out, threads, iters);
        sweep<K_FFMA, true>("FFMA other warps", out, threads, iters);
    }
    return 0;
}
