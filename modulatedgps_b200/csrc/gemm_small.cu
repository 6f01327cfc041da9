// Batched strided FP64 GEMM on the M x M operands of the replicated pre/post-compute, DMMA.8x8x4.
// C[b] = alpha * opA(A[b]) * opB(B[b]) + beta * C[b], row-major.  64x64 CTA tile, 8 warps of 16x32,
// K staged through shared memory in chunks of 32.  These GEMMs are O(M^3), latency-bound and replicated
// on every rank; the streaming kernels (stream_kernels.cu, syrk.cu) carry the O(N M^2) work.
#include "common.cuh"
#include "kernels.h"

namespace mgp {

constexpr int GS_BM = 64, GS_BN = 64, GS_BK = 32;
constexpr int GS_AS = GS_BK + 4;  // As[m][k] stride  (== 4 mod 16 -> conflict-free A-fragment loads)
constexpr int GS_BS = GS_BN + 4;  // Bs[k][n] stride  (== 4 mod 16 -> conflict-free B-fragment loads)

__global__ void __launch_bounds__(256) gemm_small_kernel(int m, int n, int k, double alpha, const double* A, int lda,
                                                         int64_t strideA, int transA, const double* B, int ldb,
                                                         int64_t strideB, int transB, double beta, double* C, int ldc,
                                                         int64_t strideC) {
    __shared__ double As[GS_BM * GS_AS];
    __shared__ double Bs[GS_BK * GS_BS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * GS_BM, n0 = blockIdx.x * GS_BN;
    A += (int64_t)blockIdx.z * strideA;
    B += (int64_t)blockIdx.z * strideB;
    C += (int64_t)blockIdx.z * strideC;
    const int wr = (warp & 3) * 16, wc = (warp >> 2) * 32;
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int k0 = 0; k0 < k; k0 += GS_BK) {
        // stage opA tile [64 x 32]: thread mapping follows the contiguous source direction
        for (int idx = tid; idx < GS_BM * GS_BK; idx += 256) {
            int mm, kk;
            if (transA) { mm = idx % GS_BM; kk = idx / GS_BM; } else { kk = idx % GS_BK; mm = idx / GS_BK; }
            const int gm = m0 + mm, gk = k0 + kk;
            double v = 0.0;
            if (gm < m && gk < k) v = transA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
            As[mm * GS_AS + kk] = v;
        }
        for (int idx = tid; idx < GS_BK * GS_BN; idx += 256) {
            int nn, kk;
            if (transB) { kk = idx % GS_BK; nn = idx / GS_BK; } else { nn = idx % GS_BN; kk = idx / GS_BN; }
            const int gn = n0 + nn, gk = k0 + kk;
            double v = 0.0;
            if (gn < n && gk < k) v = transB ? B[(size_t)gn * ldb + gk] : B[(size_t)gk * ldb + gn];
            Bs[kk * GS_BS + nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < GS_BK / 4; ++ks) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = As[(wr + i * 8 + g) * GS_AS + ks * 4 + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(ks * 4 + t) * GS_BS + wc + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j], a[i], b[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int gm = m0 + wr + i * 8 + g, gn = n0 + wc + j * 8 + 2 * t + e;
                if (gm < m && gn < n) {
                    double* p = C + (size_t)gm * ldc + gn;
                    const double v = alpha * acc[i][j][e];
                    *p = (beta == 0.0) ? v : v + beta * (*p);
                }
            }
}

void gemm_small(int m, int n, int k, double alpha, const double* A, int lda, int64_t strideA, bool transA,
                const double* B, int ldb, int64_t strideB, bool transB, double beta, double* C, int ldc,
                int64_t strideC, int batch, const Launch& ln) {
    dim3 grid((n + GS_BN - 1) / GS_BN, (m + GS_BM - 1) / GS_BM, batch);
    gemm_small_kernel<<<grid, 256, 0, ln.stream>>>(m, n, k, alpha, A, lda, strideA, transA ? 1 : 0, B, ldb, strideB,
                                                   transB ? 1 : 0, beta, C, ldc, strideC);
    ln.tick();
}

}  // namespace mgp
