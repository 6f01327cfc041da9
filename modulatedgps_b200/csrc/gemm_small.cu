// Batched strided FP64 GEMM on the M x M operands of the replicated pre/post-compute, DMMA.8x8x4.
// C[b] = alpha * opA(A[b]) * opB(B[b]) + beta * C[b], row-major.
// These GEMMs are O(M^3), replicated on every rank and LATENCY-bound (a 256^3 product is 3.6 us of DMMA time on the
// whole chip): the tile is small (32x32, 4 warps of 16x16) so that even one 256 x 256 output spreads over 64 CTAs, and
// the k-chunks of 32 are double-buffered with cp.async, so a chunk costs one shared-memory round trip instead of a
// global one.  Aligned, full-tile operands (every call of the replicated pre/post-compute: Mp is a multiple of 32) are
// staged with 16-byte copies along the contiguous source direction — a transposed operand is stored transposed and its
// fragments are read the other way round; the general path keeps 8-byte copies with range checks (the LSU retires
// 8-byte asynchronous copies element by element: 2048 per chunk cost ~2.7 k clocks, a 256^3 product 14 us).  The streaming kernels (stream_kernels.cu, syrk.cu) carry the
// O(N M^2) work.
#include "common.cuh"
#include "kernels.h"

namespace mgp {

constexpr int GS_BM = 32, GS_BN = 32, GS_BK = 32, GS_THREADS = 128;
constexpr int GS_AS = GS_BK + 4;  // As[m][k] stride  (== 4 mod 16 -> conflict-free A-fragment loads)
constexpr int GS_BS = GS_BN + 4;  // Bs[k][n] stride  (== 4 mod 16 -> conflict-free B-fragment loads)

__global__ void __launch_bounds__(GS_THREADS) gemm_small_kernel(int m, int n, int k, double alpha, const double* A, int lda,
                                                                int64_t strideA, int transA, const double* B, int ldb,
                                                                int64_t strideB, int transB, double beta, double* C, int ldc,
                                                                int64_t strideC) {
    __shared__ __align__(16) double As[2][GS_BM * GS_AS];
    __shared__ __align__(16) double Bs[2][GS_BK * GS_BS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * GS_BM, n0 = blockIdx.x * GS_BN;
    A += (int64_t)blockIdx.z * strideA;
    B += (int64_t)blockIdx.z * strideB;
    C += (int64_t)blockIdx.z * strideC;
    const int wr = (warp & 1) * 16, wc = (warp >> 1) * 16;
    double acc[2][2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // vector path: 16-byte aligned bases and leading dimensions, no partial tiles
    const bool vec = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0 && (lda & 1) == 0 &&
                     (ldb & 1) == 0 && m % GS_BM == 0 && n % GS_BN == 0 && k % GS_BK == 0;
    const bool aT = vec && transA;   // A chunk stored [k][m]
    const bool bT = vec && transB;   // B chunk stored [n][k]
    // stage the k-chunk starting at k0 into buffer `buf`; out-of-range elements are written as zeros directly
    auto stage = [&](int k0, int buf) {
        if (vec) {
            for (int idx = tid; idx < GS_BM * GS_BK / 2; idx += GS_THREADS) {
                const int r = idx / (GS_BK / 2), c = (idx % (GS_BK / 2)) * 2;   // smem row r, columns c, c + 1
                // !transA: row = m, col = k, source A[m][k];   transA: row = k, col = m, source A[k][m]
                const double* src = transA ? A + (size_t)(k0 + r) * lda + m0 + c : A + (size_t)(m0 + r) * lda + k0 + c;
                cp_async16(&As[buf][r * GS_AS + c], src);
            }
            for (int idx = tid; idx < GS_BK * GS_BN / 2; idx += GS_THREADS) {
                const int r = idx / (GS_BN / 2), c = (idx % (GS_BN / 2)) * 2;
                // !transB: row = k, col = n, source B[k][n];   transB: row = n, col = k, source B[n][k]
                const double* src = transB ? B + (size_t)(n0 + r) * ldb + k0 + c : B + (size_t)(k0 + r) * ldb + n0 + c;
                cp_async16(&Bs[buf][r * GS_BS + c], src);
            }
            cp_async_commit();
            return;
        }
        for (int idx = tid; idx < GS_BM * GS_BK; idx += GS_THREADS) {
            int mm, kk;   // thread mapping follows the contiguous source direction
            if (transA) { mm = idx % GS_BM; kk = idx / GS_BM; } else { kk = idx % GS_BK; mm = idx / GS_BK; }
            const int gm = m0 + mm, gk = k0 + kk;
            double* d = &As[buf][mm * GS_AS + kk];
            if (gm < m && gk < k) cp_async8(d, transA ? A + (size_t)gk * lda + gm : A + (size_t)gm * lda + gk);
            else *d = 0.0;
        }
        for (int idx = tid; idx < GS_BK * GS_BN; idx += GS_THREADS) {
            int nn, kk;
            if (transB) { kk = idx % GS_BK; nn = idx / GS_BK; } else { nn = idx % GS_BN; kk = idx / GS_BN; }
            const int gn = n0 + nn, gk = k0 + kk;
            double* d = &Bs[buf][kk * GS_BS + nn];
            if (gn < n && gk < k) cp_async8(d, transB ? B + (size_t)gn * ldb + gk : B + (size_t)gk * ldb + gn);
            else *d = 0.0;
        }
        cp_async_commit();
    };
    const int nchunk = (k + GS_BK - 1) / GS_BK;
    stage(0, 0);
    for (int c = 0; c < nchunk; ++c) {
        const int buf = c & 1;
        if (c + 1 < nchunk) { stage((c + 1) * GS_BK, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const double* as = As[buf];
        const double* bs = Bs[buf];
#pragma unroll
        for (int ks = 0; ks < GS_BK / 4; ++ks) {
            double a[2], b[2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
                a[i] = aT ? as[(ks * 4 + t) * GS_AS + wr + i * 8 + g] : as[(wr + i * 8 + g) * GS_AS + ks * 4 + t];
#pragma unroll
            for (int j = 0; j < 2; ++j)
                b[j] = bT ? bs[(wc + j * 8 + g) * GS_BS + ks * 4 + t] : bs[(ks * 4 + t) * GS_BS + wc + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma(acc[i][j], a[i], b[j]);
        }
        __syncthreads();   // this buffer is refilled two chunks from now
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
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
    gemm_small_kernel<<<grid, GS_THREADS, 0, ln.stream>>>(m, n, k, alpha, A, lda, strideA, transA ? 1 : 0, B, ldb, strideB,
                                                          transB ? 1 : 0, beta, C, ldc, strideC);
    ln.tick();
}

}  // namespace mgp
