// FP64 issue-rate microbenchmark for sm_100a: DFMA vs DMMA (mma.sync f64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double a, double b) {
    double c[NACC][4];
    double af[4] = {a, a * 0.5, a * 0.25, a * 0.125}, bf[2] = {b, b * 0.5};
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma1688(c[i], af, bf);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA fed from shared memory: warp tile (8*MI) x (8*NI), B from smem (LDS.64), A from smem.
template <int MI, int NI>
__global__ void __launch_bounds__(256) dmma_smem_kernel(double* out, int iters) {
    __shared__ double sa[MI * 8 * 36];
    __shared__ double sb[32 * 68];
    for (int i = threadIdx.x; i < MI * 8 * 36; i += blockDim.x) sa[i] = 1e-3 * (i % 7);
    for (int i = threadIdx.x; i < 32 * 68; i += blockDim.x) sb[i] = 1e-3 * (i % 5);
    __syncthreads();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double c0[MI][NI], c1[MI][NI];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) { c0[i][j] = 0; c1[i][j] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 8 k4-steps over 32 k
            double af[MI], bf[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) af[i] = sa[(i * 8 + g) * 36 + k * 4 + t] + warp;
#pragma unroll
            for (int j = 0; j < NI; ++j) bf[j] = sb[(k * 4 + t) * 68 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(c0[i][j], c1[i][j], af[i], bf[j]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) s += c0[i][j] + c1[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_it(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, nsm, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 256 * 4));
    const int iters = 4096;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = nsm * bps;
        {
            float ms = time_it([&] { dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * grid * 256.0 * iters * 8;
            printf("DFMA        bps=%d warps/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, bps * 8, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dmma884_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * grid * 8.0 * iters * 8 * 256;
            printf("DMMA m8n8k4 bps=%d warps/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, bps * 8, ms, fl / ms * 1e-9);
        }
        {
            float ms = time_it([&] { dmma1688_kernel<4><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double fl = 2.0 * grid * 8.0 * iters * 4 * 1024;
            printf("DMMA m16n8k8 bps=%d warps/SM=%d : %.3f ms  %.2f TFLOP/s\n", bps, bps * 8, ms, fl / ms * 1e-9);
        }
    }
    // fewer warps: 4 warps per SM
    {
        float ms = time_it([&] { dmma884_kernel<8><<<nsm, 128>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * nsm * 4.0 * iters * 8 * 256;
        printf("DMMA m8n8k4 4 warps/SM : %.3f ms  %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    for (int bps = 1; bps <= 2; ++bps) {
        int grid = nsm * bps;
        float ms = time_it([&] { dmma_smem_kernel<2, 4><<<grid, 256>>>(out, 512); }, 5);
        double fl = 2.0 * grid * 8.0 * 512 * 8 * 8 * 256;
        printf("DMMA smem-fed 16x32 tile bps=%d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        ms = time_it([&] { dmma_smem_kernel<4, 4><<<grid, 256>>>(out, 512); }, 5);
        fl = 2.0 * grid * 8.0 * 512 * 8 * 16 * 256;
        printf("DMMA smem-fed 32x32 tile bps=%d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
        ms = time_it([&] { dmma_smem_kernel<4, 8><<<grid, 256>>>(out, 512); }, 5);
        fl = 2.0 * grid * 8.0 * 512 * 8 * 32 * 256;
        printf("DMMA smem-fed 32x64 tile bps=%d: %.3f ms  %.2f TFLOP/s\n", bps, ms, fl / ms * 1e-9);
    }
    // sustained: 2 seconds of DMMA
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0;
        for (; n < 400; ++n) dmma884_kernel<8><<<nsm * 2, 256>>>(out, iters * 4, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double fl = 2.0 * nsm * 2 * 8.0 * iters * 4 * 8 * 256 * n;
        printf("DMMA m8n8k4 sustained %.0f ms: %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    CK(cudaFree(out));
    return 0;
}
