// S_k += A diag(vbar_k) A^T : the only large reduction over points in the backward pass.
//
// With S_k and A mubar in hand, every M x M gradient of the layer follows from replicated O(M^3) algebra
// (precompute.cu::finish_layer):   dLq_k = tril(2 S_k Lq_k),   Abar A^T = sum_k Q_k S_k + q_mu (A mubar)^T,
// Lbar = -tril(L^-T Abar A^T)  — i.e. the N-reductions  tril(A Bbar_k^T)  and  tril(Kuf_bar A^T)  of SURVEY.md
// Appendix B are folded into K weighted Gram matrices of the materialised A.
//
// Output-stationary DMMA kernel: CTA = (64x64 tile pair I >= J, split of the point range, group of 4
// components); warp = (component, 32-row half) with a 32x64 accumulator tile; A slabs of 32 points are
// double-buffered through shared memory with cp.async.  Each CTA accumulates into its own slot of `part`
// (no atomics; deterministic), the slots are summed by reduce_partials.
#include "common.cuh"
#include "kernels.h"

namespace mgp {

constexpr int SY_KC = 32;           // points per stage
constexpr int SY_STR = SY_KC + 4;   // == 4 mod 16: conflict-free fragment loads
constexpr int SY_THREADS = 256;
constexpr int SY_NSTAGE = 3;
constexpr int SY_MUB = 12;          // [n][k] stride of the mubar slab (conflict-free B-fragment loads)
constexpr int SY_STAGE = 2 * 64 * SY_STR + 4 * SY_KC + SY_KC * SY_MUB;  // doubles per stage: AI, AJ, weights[4][KC], mubar[KC][12]

__device__ __forceinline__ void syrk_load_stage(double* st, const double* A, const double* vbar, const double* mubar,
                                                int Mp, int K, int64_t ldn, int I, int J, int kbase, int64_t p0) {
    double* AI = st;
    double* AJ = st + 64 * SY_STR;
    double* wt = st + 2 * 64 * SY_STR;
    for (int idx = threadIdx.x; idx < 64 * (SY_KC / 2); idx += SY_THREADS) {
        const int row = idx / (SY_KC / 2), c2 = idx % (SY_KC / 2);
        const int ri = I * 64 + row, rj = J * 64 + row;
        if (ri < Mp) cp_async16(AI + row * SY_STR + 2 * c2, A + (size_t)ri * ldn + p0 + 2 * c2);
        if (I != J && rj < Mp) cp_async16(AJ + row * SY_STR + 2 * c2, A + (size_t)rj * ldn + p0 + 2 * c2);
    }
    // weights / mubar slabs: 8-byte async copies (entries for k >= K are never written and stay zero)
    for (int idx = threadIdx.x; idx < 4 * SY_KC; idx += SY_THREADS) {
        const int kl = idx / SY_KC, n = idx % SY_KC;
        const int k = kbase + kl;
        if (k < K) cp_async8(wt + kl * SY_KC + n, vbar + (size_t)(p0 + n) * K + k);
    }
    if (mubar != nullptr) {
        double* mb = st + 2 * 64 * SY_STR + 4 * SY_KC;
        for (int idx = threadIdx.x; idx < SY_KC * K; idx += SY_THREADS) {
            const int n = idx / K, k = idx - n * K;
            cp_async8(mb + n * SY_MUB + k, mubar + (size_t)(p0 + n) * K + k);
        }
    }
    cp_async_commit();
}

__global__ void __launch_bounds__(SY_THREADS, 1) syrk_kernel(const double* A, const double* vbar, const double* mubar,
                                                             double* part, double* mraw_part, int Mp, int K, int64_t ldn,
                                                             int64_t n, int64_t per_split) {
    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    // tile pair (I >= J) from blockIdx.x
    int I = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
    while (I * (I + 1) / 2 > (int)blockIdx.x) --I;
    while ((I + 1) * (I + 2) / 2 <= (int)blockIdx.x) ++I;
    const int J = blockIdx.x - I * (I + 1) / 2;
    const int split = blockIdx.y, kbase = blockIdx.z * 4;
    const int kl = warp >> 1, half = warp & 1, k = kbase + kl;
    const int64_t pbeg = (int64_t)split * per_split;
    int64_t pend = pbeg + per_split;
    const int64_t nround = (n + SY_KC - 1) / SY_KC * SY_KC;
    if (pend > nround) pend = nround;
    if (pbeg >= pend) return;
    const int nstage = (int)((pend - pbeg) / SY_KC);
    // the J == 0 tile column of component group 0 sees every row block of A exactly once: it also forms A mubar
    const bool do_mraw = (J == 0 && blockIdx.z == 0);
    const double* mub_src = do_mraw ? mubar : nullptr;
    double am[2] = {0.0, 0.0};

    for (int idx = threadIdx.x; idx < SY_NSTAGE * SY_STAGE; idx += SY_THREADS) smem[idx] = 0.0;   // rows >= Mp stay zero
    __syncthreads();

    double acc[4][8][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    // SY_NSTAGE-deep cp.async ring, ONE barrier per stage: after the barrier of iteration s every warp has finished
    // multiplying stage s-1, so its buffer can be refilled (with stage s + SY_NSTAGE - 1) right away.
#pragma unroll
    for (int p = 0; p < SY_NSTAGE - 1; ++p) {
        if (p < nstage) syrk_load_stage(smem + p * SY_STAGE, A, vbar, mub_src, Mp, K, ldn, I, J, kbase, pbeg + (int64_t)p * SY_KC);
        else cp_async_commit();
    }
    for (int s = 0; s < nstage; ++s) {
        double* cur = smem + (s % SY_NSTAGE) * SY_STAGE;
        cp_async_wait<SY_NSTAGE - 2>();   // stage s has landed (only the newest SY_NSTAGE-2 groups may be pending)
        __syncthreads();
        if (s + SY_NSTAGE - 1 < nstage)
            syrk_load_stage(smem + ((s + SY_NSTAGE - 1) % SY_NSTAGE) * SY_STAGE, A, vbar, mub_src, Mp, K, ldn, I, J, kbase,
                            pbeg + (int64_t)(s + SY_NSTAGE - 1) * SY_KC);
        else
            cp_async_commit();   // empty group keeps the wait count uniform
        if (k < K) {
            const double* AI = cur + (half * 32 + g) * SY_STR + t;
            const double* AJ = (I == J ? cur : cur + 64 * SY_STR) + g * SY_STR + t;
            const double* wt = cur + 2 * 64 * SY_STR + kl * SY_KC + t;
#pragma unroll
            for (int ks = 0; ks < SY_KC / 4; ++ks) {
                double a[4], b[8];
                const double wv = wt[ks * 4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = AI[mi * 8 * SY_STR + ks * 4];
#pragma unroll
                for (int ni = 0; ni < 8; ++ni) b[ni] = AJ[ni * 8 * SY_STR + ks * 4] * wv;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 8; ++ni) dmma(acc[mi][ni], a[mi], b[ni]);
            }
        }
        if (do_mraw) {   // warp w: rows I*64 + 8w .. +8 ; columns = components
            const double* AIw = cur + (warp * 8 + g) * SY_STR + t;
            const double* mb = cur + 2 * 64 * SY_STR + 4 * SY_KC + t * SY_MUB + g;
#pragma unroll
            for (int ks = 0; ks < SY_KC / 4; ++ks) dmma(am, AIw[ks * 4], mb[ks * 4 * SY_MUB]);
        }
    }
    if (do_mraw) {
        const int row = I * 64 + warp * 8 + g;
        if (row < Mp) {
            double* p = mraw_part + ((size_t)split * Mp + row) * KP + 2 * t;
            p[0] += am[0];
            p[1] += am[1];
        }
    }
    if (k < K) {
        double* P = part + ((size_t)split * K + k) * Mp * Mp;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni) {
                const int row = I * 64 + half * 32 + mi * 8 + g, col = J * 64 + ni * 8 + 2 * t;
                if (row < Mp && col < Mp) {
                    double* p = P + (size_t)row * Mp + col;
                    p[0] += acc[mi][ni][0];
                    p[1] += acc[mi][ni][1];
                }
            }
    }
}

int syrk_num_splits(int Mp, int K, const Launch& ln) {
    const int nb = (Mp + 63) / 64, npairs = nb * (nb + 1) / 2, kgroups = (K + 3) / 4;
    int ns = ln.num_sms / (npairs * kgroups);
    return ns < 1 ? 1 : ns;
}

void syrk_accumulate(const LayerDev& ly, const ChunkBuffers& cb, double* part, double* mraw_part, int nsplit,
                     const Launch& ln) {
    const int nb = (ly.Mp + 63) / 64, npairs = nb * (nb + 1) / 2, kgroups = (ly.K + 3) / 4;
    const int64_t nchunks = (cb.n + SY_KC - 1) / SY_KC;
    int ns = nsplit;
    if (ns > nchunks) ns = (int)nchunks;
    const int64_t per_split = (nchunks + ns - 1) / ns * SY_KC;
    const size_t smem = (size_t)SY_NSTAGE * SY_STAGE * sizeof(double);
    cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(npairs, ns, kgroups);
    syrk_kernel<<<grid, SY_THREADS, smem, ln.stream>>>(cb.A, cb.vbar, cb.mubar, part, mraw_part, ly.Mp, ly.K, cb.ldn, cb.n,
                                                       per_split);
    ln.tick();
}

}  // namespace mgp
