// Ceiling of the "left operand from L2 (fragment-major LDG), right operand from shared memory" DMMA loop used by the
// cond_* kernels, as a function of warps per SM.  No tile switching, no epilogue: what the inner loop itself can reach.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/wloop_bench tools/wloop_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
struct WFrag { double a0[4], a1[4]; };
__device__ __forceinline__ void wfrag_load(WFrag& f, const double* w, int rb8, int C4, int kb, int lane) {
    const double* w0 = w + ((size_t)rb8 * C4 + kb) * 32 + lane;
    const double* w1 = w0 + (size_t)C4 * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) { f.a0[j] = __ldg(w0 + j * 32); f.a1[j] = __ldg(w1 + j * 32); }
}
// each warp: repeatedly  C(16 x 32) += W[rows of block b, all k] * T   for blocks b = warp, warp + nwarps, ...
template <int NT>
__global__ void wloop(const double* W, int Mp, int nmat, int iters, double* out) {
    constexpr int NF = NT / 8, STR = NT + 4;
    extern __shared__ double T[];
    for (int i = threadIdx.x; i < Mp * STR; i += blockDim.x) T[i] = 1e-3 * (i % 13);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const double* tb = T + t * STR + g;
    double acc[2][NF][2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
    WFrag f, n;
    int b = warp % nb16, m = 0;
    wfrag_load(f, W, 2 * b, C4, 0, lane);
    for (int it = 0; it < iters; ++it) {
        const double* w = W + (size_t)m * Mp * Mp;
        int nb = b + nw; int nm = m;
        if (nb >= nb16) { nb -= nb16; nm = (m + 1) % nmat; }
        const double* wn = W + (size_t)nm * Mp * Mp;
        for (int kb = 0; kb < C4; kb += 4) {
            if (kb + 4 < C4) wfrag_load(n, w, 2 * b, C4, kb + 4, lane);
            else wfrag_load(n, wn, 2 * nb, C4, 0, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double* tr = tb + (size_t)(kb + j) * 4 * STR;
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) {
                    const double bb = tr[nf * 8];
                    dmma(acc[0][nf], f.a0[j], bb);
                    dmma(acc[1][nf], f.a1[j], bb);
                }
            }
            f = n;
        }
        b = nb; m = nm;
    }
    double s = 0;
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) s += acc[mf][nf][0] + acc[mf][nf][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


struct Seg { const double* w; int rb8; int kb0; };
__device__ __forceinline__ void wfrag_load2(WFrag& f, const Seg& sg, int C4, int kb, int lane) { wfrag_load(f, sg.w, sg.rb8, C4, kb, lane); }
template <int NT, int NFW>
__device__ __forceinline__ void wgemm_group(const WFrag& f, const double* tb, int kb, int kb_last, double (&bq)[NFW], double (&acc)[2][NFW][2]) {
    constexpr int STR = NT + 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double b[NFW];
#pragma unroll
        for (int nf = 0; nf < NFW; ++nf) b[nf] = bq[nf];
        if (kb + j < kb_last) {
            const double* tr = tb + (size_t)(kb + j + 1) * 4 * STR;
#pragma unroll
            for (int nf = 0; nf < NFW; ++nf) bq[nf] = tr[nf * 8];
        }
#pragma unroll
        for (int nf = 0; nf < NFW; ++nf) { dmma(acc[0][nf], f.a0[j], b[nf]); dmma(acc[1][nf], f.a1[j], b[nf]); }
    }
}
template <int NT, int NFW>
__device__ __forceinline__ void wgemm_seg(const Seg& cur, int kb1, int C4, const double* Tsm, double (&acc)[2][NFW][2], int lane, WFrag& f, const Seg& nxt) {
    constexpr int STR = NT + 4;
    const int g = lane >> 2, t = lane & 3;
    const double* tb = Tsm + t * STR + g;
    const int kb_last = kb1 - 1;
    double bq[NFW];
    { const double* tr = tb + (size_t)cur.kb0 * 4 * STR;
#pragma unroll
      for (int nf = 0; nf < NFW; ++nf) bq[nf] = tr[nf * 8]; }
    int kb = cur.kb0;
    WFrag n;
    if (((kb1 - kb) >> 2) & 1) {
        if (kb + 4 < kb1) wfrag_load2(n, cur, C4, kb + 4, lane); else wfrag_load2(n, nxt, C4, nxt.kb0, lane);
        wgemm_group<NT, NFW>(f, tb, kb, kb_last, bq, acc);
        f = n; kb += 4;
    }
    for (; kb < kb1; kb += 8) {
        wfrag_load2(n, cur, C4, kb + 4, lane);
        wgemm_group<NT, NFW>(f, tb, kb, kb_last, bq, acc);
        if (kb + 8 < kb1) wfrag_load2(f, cur, C4, kb + 8, lane); else wfrag_load2(f, nxt, C4, nxt.kb0, lane);
        wgemm_group<NT, NFW>(n, tb, kb + 4, kb_last, bq, acc);
    }
}
// same schedule as wloop but through the product's wgemm_seg; CS column parts per row set (warp = h * 8 + ws)
template <int NT, int CS>
__global__ void wloop_pp(const double* W, int Mp, int nmat, int iters, double* out) {
    constexpr int NFW = NT / 8 / CS, STR = NT + 4;
    extern __shared__ double T[];
    for (int i = threadIdx.x; i < Mp * STR; i += blockDim.x) T[i] = 1e-3 * (i % 13);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ws = warp % 8, h = warp / 8;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    double acc[2][NFW][2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NFW; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
    WFrag f;
    int b = ws, m = 0;
    Seg cur{W, 2 * b, 0};
    wfrag_load2(f, cur, C4, 0, lane);
    for (int it = 0; it < iters; ++it) {
        int nb = b + 8, nm = m;
        if (nb >= nb16) { nb -= nb16; nm = (m + 1) % nmat; }
        Seg nxt{W + (size_t)nm * Mp * Mp, 2 * nb, 0};
        wgemm_seg<NT, NFW>(cur, C4, C4, T + h * (NT / CS), acc, lane, f, nxt);
        cur = nxt; b = nb; m = nm;
    }
    double s = 0;
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NFW; ++nf) s += acc[mf][nf][0] + acc[mf][nf][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// simple loop + one prefetch.global.L1 instruction per group, PD groups ahead (16 lanes cover the 16 lines of a group)
template <int NT, int PD>
__global__ void wloop_pf(const double* W, int Mp, int nmat, int iters, double* out) {
    constexpr int NF = NT / 8, STR = NT + 4;
    extern __shared__ double T[];
    for (int i = threadIdx.x; i < Mp * STR; i += blockDim.x) T[i] = 1e-3 * (i % 13);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const double* tb = T + t * STR + g;
    double acc[2][NF][2];
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
    WFrag f, n;
    int b = warp % nb16, m = 0;
    wfrag_load(f, W, 2 * b, C4, 0, lane);
    for (int it = 0; it < iters; ++it) {
        const double* w = W + (size_t)m * Mp * Mp;
        int nb = b + nw; int nm = m;
        if (nb >= nb16) { nb -= nb16; nm = (m + 1) % nmat; }
        const double* wn = W + (size_t)nm * Mp * Mp;
        for (int kb = 0; kb < C4; kb += 4) {
            {   // prefetch group kb + 4 * PD of this row block (wraps into the next block's start: good enough for a ceiling test)
                int pk = kb + 4 * PD;
                const double* pw = w; int pb = b;
                if (pk >= C4) { pk -= C4; pw = wn; pb = nb; }
                if (lane < 16) {
                    const double* a = pw + ((size_t)(2 * pb + (lane >> 3)) * C4 + pk) * 32 + (lane & 7) * 16;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
                }
            }
            if (kb + 4 < C4) wfrag_load(n, w, 2 * b, C4, kb + 4, lane);
            else wfrag_load(n, wn, 2 * nb, C4, 0, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double* tr = tb + (size_t)(kb + j) * 4 * STR;
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) {
                    const double bb = tr[nf * 8];
                    dmma(acc[0][nf], f.a0[j], bb);
                    dmma(acc[1][nf], f.a1[j], bb);
                }
            }
            f = n;
        }
        b = nb; m = nm;
    }
    double s = 0;
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) s += acc[mf][nf][0] + acc[mf][nf][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// two 16-row blocks (e.g. the same rows of two components' matrices) multiplied against the SAME right operand at once:
// 4 A fragments + 4 B fragments feed 16 DMMAs per k4-step (16 accumulator chains, half the LDS per DMMA)
struct WFrag4 { double a[4][4]; };
__device__ __forceinline__ void wfrag4_load(WFrag4& f, const double* wA, const double* wB, int rb8, int C4, int kb, int lane) {
    const double* p0 = wA + ((size_t)rb8 * C4 + kb) * 32 + lane;
    const double* p1 = wB + ((size_t)rb8 * C4 + kb) * 32 + lane;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        f.a[0][j] = __ldg(p0 + j * 32); f.a[1][j] = __ldg(p0 + (size_t)C4 * 32 + j * 32);
        f.a[2][j] = __ldg(p1 + j * 32); f.a[3][j] = __ldg(p1 + (size_t)C4 * 32 + j * 32);
    }
}
template <int NT>
__global__ void __launch_bounds__(256, 1) wloop_dual(const double* W, int Mp, int nmat, int iters, double* out) {
    constexpr int NF = NT / 8, STR = NT + 4;
    extern __shared__ double T[];
    for (int i = threadIdx.x; i < Mp * STR; i += blockDim.x) T[i] = 1e-3 * (i % 13);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const double* tb = T + t * STR + g;
    double acc[4][NF][2];
#pragma unroll
    for (int mf = 0; mf < 4; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
    WFrag4 f, n;
    int b = warp % nb16, m = 0;
    wfrag4_load(f, W, W + (size_t)Mp * Mp, 2 * b, C4, 0, lane);
    for (int it = 0; it < iters; ++it) {
        const double* wA = W + (size_t)m * Mp * Mp;
        const double* wB = W + (size_t)((m + 1) % nmat) * Mp * Mp;
        int nb = b + 8, nm = m;
        if (nb >= nb16) { nb -= nb16; nm = (m + 2) % nmat; }
        const double* nA = W + (size_t)nm * Mp * Mp;
        const double* nB = W + (size_t)((nm + 1) % nmat) * Mp * Mp;
        for (int kb = 0; kb < C4; kb += 8) {
            wfrag4_load(n, wA, wB, 2 * b, C4, kb + 4, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double* tr = tb + (size_t)(kb + j) * 4 * STR;
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) {
                    const double bb = tr[nf * 8];
#pragma unroll
                    for (int mf = 0; mf < 4; ++mf) dmma(acc[mf][nf], f.a[mf][j], bb);
                }
            }
            if (kb + 8 < C4) wfrag4_load(f, wA, wB, 2 * b, C4, kb + 8, lane);
            else wfrag4_load(f, nA, nB, 2 * nb, C4, 0, lane);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double* tr = tb + (size_t)(kb + 4 + j) * 4 * STR;
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) {
                    const double bb = tr[nf * 8];
#pragma unroll
                    for (int mf = 0; mf < 4; ++mf) dmma(acc[mf][nf], n.a[mf][j], bb);
                }
            }
        }
        b = nb; m = nm;
    }
    double s = 0;
#pragma unroll
    for (int mf = 0; mf < 4; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) s += acc[mf][nf][0] + acc[mf][nf][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount, Mp = 256, nmat = 4;
    double *W, *out;
    CK(cudaMalloc(&W, sizeof(double) * nmat * Mp * Mp));
    CK(cudaMemset(W, 0, sizeof(double) * nmat * Mp * Mp));
    auto fill = [&](bool random) {
        static double* h = (double*)malloc(sizeof(double) * nmat * Mp * Mp);
        unsigned long long x = 88172645463325252ull;
        for (size_t i = 0; i < (size_t)nmat * Mp * Mp; ++i) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            h[i] = random ? ((double)(x >> 11) / 9007199254740992.0 - 0.5) * 1e-2 : 0.0;
        }
        CK(cudaMemcpy(W, h, sizeof(double) * nmat * Mp * Mp, cudaMemcpyHostToDevice));
    };
    CK(cudaMalloc(&out, sizeof(double) * nsm * 4 * 1024));
    const int iters = 512;
    auto run = [&](auto kernel, int NT, int ctas_per_sm, int threads, int csplit = 1) {
        const size_t smem = (size_t)Mp * (NT + 4) * 8;
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        kernel<<<nsm * ctas_per_sm, threads, smem>>>(W, Mp, nmat, iters, out);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 3; ++r) kernel<<<nsm * ctas_per_sm, threads, smem>>>(W, Mp, nmat, iters, out);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= 3;
        const double fl = 2.0 * nsm * ctas_per_sm * (threads / 32) * (double)iters * 16.0 * NT * Mp / (double)csplit;
        printf("NT=%d  %d CTA/SM x %2d warps (%2d warps/SM): %.3f ms  %.2f TFLOP/s\n", NT, ctas_per_sm, threads / 32,
               ctas_per_sm * threads / 32, ms, fl / ms * 1e-9);
    };
    run(wloop<32>, 32, 1, 128);
    run(wloop<32>, 32, 1, 256);
    run(wloop<32>, 32, 1, 384);
    run(wloop<32>, 32, 1, 512);
    run(wloop<32>, 32, 2, 256);
    run(wloop<32>, 32, 3, 256);
    run(wloop<16>, 16, 1, 256);
    run(wloop<16>, 16, 1, 512);
    run(wloop<16>, 16, 2, 512);
    printf("product wgemm_seg (ping-pong fragments, B prefetch), W = 0:\n");
    run(wloop_pp<32, 1>, 32, 1, 256);
    run(wloop_pp<32, 2>, 32, 1, 512, 2);
    run(wloop_pp<32, 1>, 32, 2, 256);
    printf("simple loop + prefetch.global.L1 PD groups ahead, 8 warps:\n");
    run(wloop_pf<32, 2>, 32, 1, 256);
    run(wloop_pf<32, 3>, 32, 1, 256);
    run(wloop_pf<32, 4>, 32, 1, 256);
    printf("two 16-row blocks per warp against one right operand (16 chains, 8 warps); flops = 2x per iteration:\n");
    run(wloop_dual<32>, 32, 1, 256, 1);
    fill(true);
    printf("same with random W (data-dependent power / clocks?):\n");
    run(wloop_pp<32, 1>, 32, 1, 256);
    run(wloop_pp<32, 2>, 32, 1, 512, 2);
    run(wloop_pp<32, 1>, 32, 2, 256);
    run(wloop<32>, 32, 1, 512);
    return 0;
}
