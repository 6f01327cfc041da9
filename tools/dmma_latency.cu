// Dependent-issue latency of DMMA.8x8x4 on sm_100a: one warp per SM sub-partition runs C independent accumulator chains
// (each DMMA of a chain consumes the previous one's result).  clocks per DMMA = max(issue interval, latency / C).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_latency tools/dmma_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int C>
__global__ void __launch_bounds__(512) chains(double* out, int iters, double a, double b, long long* clk) {
    double c0[C], c1[C];
#pragma unroll
    for (int i = 0; i < C; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 32 / C; ++r)
#pragma unroll
            for (int i = 0; i < C; ++i) dmma884(c0[i], c1[i], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
template <int C>
static void run(double* out, long long* clk, int sms, int threads) {
    const int iters = 4000;
    chains<C><<<sms, threads>>>(out, iters, 1.0000001, 1e-9, clk);
    CK(cudaDeviceSynchronize());
    long long h; CK(cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost));
    const double per = (double)h / ((double)iters * 32);
    printf("%2d warps/SM, %2d chains per warp: %6.1f clocks per DMMA per warp  -> %5.1f %% of the pipe per sub-partition\n", threads / 32, C, per,
           100.0 * 16.0 * (threads / 128.0) / per);
}
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    double* out; long long* clk;
    CK(cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 512)); CK(cudaMalloc(&clk, 8));
    for (int threads : {128, 256, 512}) {
        run<1>(out, clk, p.multiProcessorCount, threads); run<2>(out, clk, p.multiProcessorCount, threads);
        run<4>(out, clk, p.multiProcessorCount, threads); run<8>(out, clk, p.multiProcessorCount, threads);
        run<16>(out, clk, p.multiProcessorCount, threads); run<32>(out, clk, p.multiProcessorCount, threads);
    }
    return 0;
}
