// The N-streaming FP64 tensor-core kernels of the SVGP conditional, forward and backward.
//
// Every kernel has the same shape: a persistent CTA walks over tiles of NT points; per tile it stages one
// [Mp x NT] right operand T in shared memory (generated RBF cross-covariances, or a slab of a materialised
// [Mp, N] array), and each warp computes 16-row blocks  C = W[rows, k-range] * T  on DMMA.8x8x4 with the
// left operand W streamed from L2 in fragment-major order (one coalesced 256-byte load per fragment).
// Triangular left operands only visit their non-zero k-range; 16-row blocks are dealt to the 8 warps in
// snake order so that the triangular work is balanced.
//
//   cond_fwd_a : T = Kuf tile (r^2 contraction on DMMA + exp)   A = L^-1 T            -> A, |a_n|^2
//   cond_fwd_b : T = A tile        B_k = Lq_k^T T (norms only), mean = q_mu^T T       -> fmean, fvar
//   cond_bwd_a : T = B_k tiles     Abar = sum_k Lq_k T_k diag(2 vbar_k) + q_mu mubar^T - 2 A diag(sum vbar)  -> Abar (over A)
//   cond_bwd_b : T = Abar tile     Kuf_bar = L^-T T ; E = Kuf_bar .* Kuf              -> sums E [1, xs, xs^2]
//
// Reference arithmetic replaced: gpflow SquaredExponential.K + base_conditional as called from
// IndependentPosteriorSingleOutputModified._conditional_fused (MixtureGPs/models.py:129-144), evaluated once
// per point instead of once per (sample, point) (SGP.integrate only tiles X, models.py:35-36), and TF's
// reverse pass through it (utils/training_utils.py:8-10).  Math: SURVEY.md Appendix B.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace mgp {

constexpr int SK_WARPS = 8;
constexpr int SK_THREADS = SK_WARPS * 32;

__host__ __device__ inline int xs_stride(int Dp) { return ((Dp - 4 + 15) / 16) * 16 + 4; }

// 16-row block dealt to warp w in round r (snake order); returns -1 past the end
__device__ __forceinline__ int snake_block(int round, int warp, int nb16) {
    const int b = round * SK_WARPS + ((round & 1) ? (SK_WARPS - 1 - warp) : warp);
    return b < nb16 ? b : -1;
}

// acc[2][NF][2] += W[rows 8*rb8 .. +16, k4-blocks kb0..kb1) * T[(kb*4 ..), :]
// Wf points at k4-block 0 of this segment for row-block 0; C4 = total k4-blocks per row-block of W.
// SCALED: right-operand column (nf*8+g) is multiplied by sc[nf] (diag(vbar_k) folded into the B fragment).
template <int NT, bool SCALED>
__device__ __forceinline__ void wgemm_block(const double* __restrict__ Wf, int C4, int rb8, int kb0, int kb1,
                                            const double* Tsm, double (&acc)[2][NT / 8][2],
                                            const double (&sc)[NT / 8], int lane) {
    constexpr int NF = NT / 8, STR = NT + 4;
    const int g = lane >> 2, t = lane & 3;
    const double* w0 = Wf + ((size_t)rb8 * C4) * 32 + lane;
    const double* w1 = w0 + (size_t)C4 * 32;
    const double* tb = Tsm + t * STR + g;
    int kb = kb0;
    const int ngroups = (kb1 - kb0) >> 2;
    double a0[4], a1[4];
    if (ngroups > 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a0[j] = __ldg(w0 + (size_t)(kb + j) * 32);
            a1[j] = __ldg(w1 + (size_t)(kb + j) * 32);
        }
    }
    for (int it = 0; it < ngroups; ++it) {
        double n0[4], n1[4];
        if (it + 1 < ngroups) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                n0[j] = __ldg(w0 + (size_t)(kb + 4 + j) * 32);
                n1[j] = __ldg(w1 + (size_t)(kb + 4 + j) * 32);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double* tr = tb + (size_t)(kb + j) * 4 * STR;
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) {
                double b = tr[nf * 8];
                if (SCALED) b *= sc[nf];
                dmma(acc[0][nf], a0[j], b);
                dmma(acc[1][nf], a1[j], b);
            }
        }
        if (it + 1 < ngroups) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a0[j] = n0[j];
                a1[j] = n1[j];
            }
        }
        kb += 4;
    }
    for (; kb < kb1; ++kb) {
        const double x0 = __ldg(w0 + (size_t)kb * 32), x1 = __ldg(w1 + (size_t)kb * 32);
        const double* tr = tb + (size_t)kb * 4 * STR;
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
            double b = tr[nf * 8];
            if (SCALED) b *= sc[nf];
            dmma(acc[0][nf], x0, b);
            dmma(acc[1][nf], x1, b);
        }
    }
}

template <int NF>
__device__ __forceinline__ void zero_acc(double (&acc)[2][NF][2]) {
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
}

// number of 16-row blocks dealt to this warp (only the last snake round can be short)
__device__ __forceinline__ int my_block_count(int warp, int nb16) {
    const int R = (nb16 + SK_WARPS - 1) / SK_WARPS;
    return R == 0 ? 0 : (snake_block(R - 1, warp, nb16) >= 0 ? R : R - 1);
}

// ---- cross-segment software pipelining of the left-operand fragments ------------------------------------------
// A warp's work is a fixed sequence of segments (pass, 16-row block, k-range).  The first fragment group of the
// NEXT segment is fetched while the last group of the current one is multiplied, so no L2 round trip is exposed at
// block / pass / tile boundaries.
struct Seg {
    const double* w;   // fragment-major base of this pass's left operand (k4-block 0 of row-block 0)
    int rb8;           // first 8-row block of the 16-row block
    int kb0;           // first k4-block of the segment
};
struct WFrag {
    double a0[4], a1[4];
};
__device__ __forceinline__ void wfrag_load(WFrag& f, const Seg& sg, int C4, int kb, int lane) {
    const double* w0 = sg.w + ((size_t)sg.rb8 * C4 + kb) * 32 + lane;
    const double* w1 = w0 + (size_t)C4 * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        f.a0[j] = __ldg(w0 + j * 32);
        f.a1[j] = __ldg(w1 + j * 32);
    }
}
// acc += W[segment rows, kb0..kb1) * T ; (kb1 - kb0) must be a positive multiple of 4.  On entry `f` holds the
// first group of `cur`; on exit it holds the first group of `nxt`.
template <int NT>
__device__ __forceinline__ void wgemm_seg(const Seg& cur, int kb1, int C4, const double* Tsm, double (&acc)[2][NT / 8][2],
                                          int lane, WFrag& f, const Seg& nxt) {
    constexpr int NF = NT / 8, STR = NT + 4;
    const int g = lane >> 2, t = lane & 3;
    const double* tb = Tsm + t * STR + g;
    for (int kb = cur.kb0; kb < kb1; kb += 4) {
        WFrag n;
        if (kb + 4 < kb1) wfrag_load(n, cur, C4, kb + 4, lane);
        else wfrag_load(n, nxt, C4, nxt.kb0, lane);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double* tr = tb + (size_t)(kb + j) * 4 * STR;
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) {
                const double b = tr[nf * 8];
                dmma(acc[0][nf], f.a0[j], b);
                dmma(acc[1][nf], f.a1[j], b);
            }
        }
        f = n;
    }
}

// Xs[n][d] = X[n0+n][d] / lengthscale_d (0 outside the chunk / padding); xs2[n] = |Xs_n|^2
template <int NT>
__device__ __forceinline__ void stage_x(const LayerDev& ly, const ChunkBuffers& cb, int64_t n0, double* Xs,
                                        double* xs2) {
    const int Dp = ly.Dp, D = ly.D, XSTR = xs_stride(Dp);
    for (int idx = threadIdx.x; idx < NT * Dp; idx += SK_THREADS) {
        const int n = idx / Dp, d = idx % Dp;
        double v = 0.0;
        if (n0 + n < cb.n && d < D) v = cb.X[(size_t)(n0 + n) * D + d] / ly.lengthscales[ly.n_ls == 1 ? 0 : d];
        Xs[n * XSTR + d] = v;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < NT; n += SK_THREADS) {
        double s = 0.0;
        for (int d = 0; d < Dp; ++d) s += Xs[n * XSTR + d] * Xs[n * XSTR + d];
        xs2[n] = s;
    }
    __syncthreads();
}

// Kuf values of the 8-row block rb8 in C-fragment layout: kv[nf][e] = k(z_{8 rb8+g}, x_{nf*8+2t+e}).
// The -2 Zs.Xs contraction runs on DMMA (north_star: "squared-distance term on FP64 DMMA").
template <int NT>
__device__ __forceinline__ void gen_kuf_block(const LayerDev& ly, int rb8, const double* Xs, const double* xs2,
                                              double variance, double (&kv)[NT / 8][2], int lane) {
    constexpr int NF = NT / 8;
    const int g = lane >> 2, t = lane & 3;
    const int Dp = ly.Dp, XSTR = xs_stride(Dp), D4 = Dp >> 2;
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) kv[nf][0] = kv[nf][1] = 0.0;
    for (int kd = 0; kd < D4; ++kd) {
        const double a = __ldg(ly.Zs_fm + ((size_t)rb8 * D4 + kd) * 32 + lane);
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) dmma(kv[nf], a, Xs[(nf * 8 + g) * XSTR + kd * 4 + t]);
    }
    const int i = rb8 * 8 + g;
    const double zi = __ldg(ly.zs2 + i);
    const bool live = i < ly.M;
#pragma unroll
    for (int nf = 0; nf < NF; ++nf)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double r2 = -2.0 * kv[nf][e] + (zi + xs2[nf * 8 + 2 * t + e]);   // square_distance(X, X2)
            kv[nf][e] = live ? variance * exp(-0.5 * r2) : 0.0;                    // K_r2
        }
}

// T[:, 0..NT) <- G[:, n0 .. n0+NT)  for a row-major [Mp, ldn] array (16-byte cp.async)
template <int NT>
__device__ __forceinline__ void load_tile_async(double* T, const double* G, int Mp, int64_t ldn, int64_t n0) {
    constexpr int STR = NT + 4, C2 = NT / 2;
    for (int idx = threadIdx.x; idx < Mp * C2; idx += SK_THREADS) {
        const int row = idx / C2, c2 = idx % C2;
        cp_async16(T + (size_t)row * STR + 2 * c2, G + (size_t)row * ldn + n0 + 2 * c2);
    }
    cp_async_commit();
}

// ==================================================================================================
// cond_fwd_a
// ==================================================================================================
template <int NT>
__global__ void __launch_bounds__(SK_THREADS) cond_fwd_a_kernel(LayerDev ly, ChunkBuffers cb, int ntiles) {
    constexpr int NF = NT / 8, STR = NT + 4;
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, XSTR = xs_stride(ly.Dp);
    double* T = smem;
    double* Xs = T + (size_t)Mp * STR;
    double* xs2 = Xs + NT * XSTR;
    double* colpart = xs2 + NT;  // [SK_WARPS][NT]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, nb8 = Mp / 8, C4 = Mp / 4;
    const double variance = ly.variance[0];
    const int nmy = my_block_count(warp, nb16);
    auto seg_of = [&](int i) { const int b = snake_block(i, warp, nb16); return Seg{ly.W_Linv, 2 * b, 0}; };
    WFrag wf;
    if (nmy > 0) wfrag_load(wf, seg_of(0), C4, 0, lane);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = (int64_t)tile * NT;
        stage_x<NT>(ly, cb, n0, Xs, xs2);
        for (int rb = warp; rb < nb8; rb += SK_WARPS) {
            double kv[NF][2];
            gen_kuf_block<NT>(ly, rb, Xs, xs2, variance, kv, lane);
#pragma unroll
            for (int nf = 0; nf < NF; ++nf)
                *reinterpret_cast<double2*>(T + (size_t)(rb * 8 + g) * STR + nf * 8 + 2 * t) = make_double2(kv[nf][0], kv[nf][1]);
        }
        __syncthreads();
        double colsq[NF][2];
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) colsq[nf][0] = colsq[nf][1] = 0.0;
        for (int i = 0; i < nmy; ++i) {
            const int b = snake_block(i, warp, nb16);
            double acc[2][NF][2];
            zero_acc<NF>(acc);
            wgemm_seg<NT>(seg_of(i), (b + 1) * 4, C4, T, acc, lane, wf, seg_of(i + 1 < nmy ? i + 1 : 0));   // lower triangular
#pragma unroll
            for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) {
                    const size_t row = (size_t)(b * 16 + mf * 8 + g);
                    *reinterpret_cast<double2*>(cb.A + row * cb.ldn + n0 + nf * 8 + 2 * t) =
                        make_double2(acc[mf][nf][0], acc[mf][nf][1]);
                    colsq[nf][0] += acc[mf][nf][0] * acc[mf][nf][0];
                    colsq[nf][1] += acc[mf][nf][1] * acc[mf][nf][1];
                }
        }
#pragma unroll
        for (int nf = 0; nf < NF; ++nf)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const double s = sum_over_g(colsq[nf][e]);
                if (g == 0) colpart[warp * NT + nf * 8 + 2 * t + e] = s;
            }
        __syncthreads();
        for (int n = threadIdx.x; n < NT; n += SK_THREADS) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < SK_WARPS; ++w) s += colpart[w * NT + n];
            cb.asq[n0 + n] = s;
        }
        __syncthreads();
    }
}

// ==================================================================================================
// cond_fwd_b
// ==================================================================================================
template <int NT>
__global__ void __launch_bounds__(SK_THREADS) cond_fwd_b_kernel(LayerDev ly, ChunkBuffers cb, int ntiles) {
    constexpr int NF = NT / 8, STR = NT + 4;
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, K = ly.K;
    double* T = smem;
    double* colpart = T + (size_t)Mp * STR;  // [SK_WARPS][KP][NT]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const double variance = ly.variance[0];
    const int nmy = my_block_count(warp, nb16);
    const bool does_mean = (warp == SK_WARPS - 1);
    const Seg mean_seg{ly.W_mT, 0, 0};
    auto seg_of = [&](int k, int i) {
        const int b = snake_block(i, warp, nb16);
        return Seg{ly.W_LqT + (size_t)k * Mp * Mp, 2 * b, 4 * b};
    };
    auto first_seg = [&]() { return nmy > 0 ? seg_of(0, 0) : mean_seg; };
    WFrag wf;
    if (nmy > 0 || does_mean) { const Seg s0 = first_seg(); wfrag_load(wf, s0, C4, s0.kb0, lane); }

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = (int64_t)tile * NT;
        load_tile_async<NT>(T, cb.A, Mp, cb.ldn, n0);
        cp_async_wait<0>();
        __syncthreads();
        for (int k = 0; k < K; ++k) {
            double colsq[NF][2];
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) colsq[nf][0] = colsq[nf][1] = 0.0;
            for (int i = 0; i < nmy; ++i) {
                const int b = snake_block(i, warp, nb16);
                double acc[2][NF][2];
                zero_acc<NF>(acc);
                const Seg nxt = (i + 1 < nmy) ? seg_of(k, i + 1) : (k + 1 < K ? seg_of(k + 1, 0) : (does_mean ? mean_seg : seg_of(0, 0)));
                wgemm_seg<NT>(seg_of(k, i), C4, C4, T, acc, lane, wf, nxt);   // upper triangular
                double* Bk = cb.Bk ? cb.Bk + (size_t)k * Mp * cb.ldn : nullptr;
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) {
                        if (Bk)   // kept for the backward: Abar needs Lq_k (2 B_k diag(vbar_k))
                            *reinterpret_cast<double2*>(Bk + (size_t)(b * 16 + mf * 8 + g) * cb.ldn + n0 + nf * 8 + 2 * t) =
                                make_double2(acc[mf][nf][0], acc[mf][nf][1]);
                        colsq[nf][0] += acc[mf][nf][0] * acc[mf][nf][0];
                        colsq[nf][1] += acc[mf][nf][1] * acc[mf][nf][1];
                    }
            }
#pragma unroll
            for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double s = sum_over_g(colsq[nf][e]);
                    if (g == 0) colpart[((size_t)warp * KP + k) * NT + nf * 8 + 2 * t + e] = s;
                }
        }
        if (does_mean) {  // fmean^T [K x NT] = q_mu^T [K x Mp] * A tile
            double acc[2][NF][2];
            zero_acc<NF>(acc);
            wgemm_seg<NT>(mean_seg, C4, C4, T, acc, lane, wf, first_seg());
            if (g < K) {
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        cb.fmean[(size_t)(n0 + nf * 8 + 2 * t + e) * K + g] = acc[0][nf][e];
            }
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < NT * K; idx += SK_THREADS) {
            const int n = idx / K, k = idx % K;
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < SK_WARPS; ++w) s += colpart[((size_t)w * KP + k) * NT + n];
            cb.fvar[(size_t)(n0 + n) * K + k] = (variance - cb.asq[n0 + n]) + s;   // Knn - sum A^2 + sum LTA^2
        }
        __syncthreads();
    }
}

// ==================================================================================================
// cond_bwd_a  (triangular route)
//   Abar = sum_k Lq_k * (B_k diag(2 vbar_k))  +  q_mu * mubar^T  -  2 A diag(sum_k vbar_k)
// from fvar = variance - |a|^2 + sum |b_k|^2, b_k = Lq_k^T a, fmean = a^T q_mu.  The right operand changes with k
// (B_k tiles, re-staged K times per point tile), so each warp keeps the accumulators of ALL its row blocks
// (NBW of them) in registers across the k loop.  Executed flops = algorithmic K M^2 (x17/16).
// ==================================================================================================
template <int NT, int NBW>
__global__ void __launch_bounds__(SK_THREADS) cond_bwd_a_kernel(LayerDev ly, ChunkBuffers cb, int ntiles) {
    constexpr int NF = NT / 8, STR = NT + 4;
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, K = ly.K;
    // double-buffered right operand (B_k tiles) and per-tile adjoint slabs: the next (tile, k) operand streams in
    // with cp.async while the current one is being multiplied
    double* Tb = smem;                               // [2][Mp][STR]
    double* mubTb = Tb + 2 * (size_t)Mp * STR;       // [2][KP][STR]  mubar^T (right operand of the q_mu segment)
    double* vbb = mubTb + 2 * KP * STR;              // [2][KP][NT]   vbar^T
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const double sc_dummy[NF] = {};

    for (int idx = threadIdx.x; idx < 2 * KP * STR + 2 * KP * NT; idx += SK_THREADS) mubTb[idx] = 0.0;   // k >= K rows stay 0
    __syncthreads();
    const int nmy = my_block_count(warp, nb16);
    auto seg_of = [&](int k, int i) {
        const int b = snake_block(i, warp, nb16);
        return Seg{ly.W_Lq + (size_t)k * Mp * Mp, 2 * b, 0};
    };
    WFrag wf;
    if (nmy > 0) wfrag_load(wf, seg_of(0, 0), C4, 0, lane);

    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * K;
    auto issue = [&](int j) {   // stage operand j = (tile iteration, component) into buffer j & 1
        const int ti = j / K, k = j - ti * K;
        const int64_t n0 = (int64_t)(blockIdx.x + ti * gridDim.x) * NT;
        load_tile_async<NT>(Tb + (size_t)(j & 1) * Mp * STR, cb.Bk + (size_t)k * Mp * cb.ldn, Mp, cb.ldn, n0);
        if (k == 0) {
            double* mT = mubTb + (size_t)(ti & 1) * KP * STR;
            double* vT = vbb + (size_t)(ti & 1) * KP * NT;
            for (int idx = threadIdx.x; idx < NT * K; idx += SK_THREADS) {
                const int n = idx / K, kk = idx - n * K;
                cp_async8(mT + kk * STR + n, cb.mubar + (size_t)(n0 + n) * K + kk);
                cp_async8(vT + kk * NT + n, cb.vbar + (size_t)(n0 + n) * K + kk);
            }
            cp_async_commit();
        }
    };
    if (total > 0) issue(0);
    double acc[NBW][2][NF][2];
    for (int j = 0; j < total; ++j) {
        const int ti = j / K, k = j - ti * K;
        const int64_t n0 = (int64_t)(blockIdx.x + ti * gridDim.x) * NT;
        cp_async_wait<0>();    // operand j has landed ...
        __syncthreads();       // ... for everyone, and every warp is done with operand j - 1
        if (j + 1 < total) issue(j + 1);   // refill the buffer operand j - 1 used
        const double* T = Tb + (size_t)(j & 1) * Mp * STR;
        const double* mubT = mubTb + (size_t)(ti & 1) * KP * STR;
        const double* vb = vbb + (size_t)(ti & 1) * KP * NT;
        if (k == 0) {
#pragma unroll
            for (int r = 0; r < NBW; ++r) zero_acc<NF>(acc[r]);
        }
        // column weights 2 vbar_k of this lane's columns; applied once per (block, k) to the finished product
        // Lq_k B_k instead of to every B fragment (keeps DMUL out of the DMMA loop)
        double sc[NF][2];
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
            sc[nf][0] = 2.0 * vb[k * NT + nf * 8 + 2 * t];
            sc[nf][1] = 2.0 * vb[k * NT + nf * 8 + 2 * t + 1];
        }
#pragma unroll
        for (int r = 0; r < NBW; ++r) {
            if (r >= nmy) continue;
            const int b = snake_block(r, warp, nb16);
            double ck[2][NF][2];
            zero_acc<NF>(ck);
            const Seg nxt = (r + 1 < nmy) ? seg_of(k, r + 1) : seg_of(k + 1 < K ? k + 1 : 0, 0);
            wgemm_seg<NT>(seg_of(k, r), (b + 1) * 4, C4, T, ck, lane, wf, nxt);   // lower triangular
#pragma unroll
            for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) {
                    acc[r][mf][nf][0] = fma(ck[mf][nf][0], sc[nf][0], acc[r][mf][nf][0]);
                    acc[r][mf][nf][1] = fma(ck[mf][nf][1], sc[nf][1], acc[r][mf][nf][1]);
                }
        }
        if (k == K - 1) {   // tile epilogue: + q_mu mubar^T - 2 A diag(sum_k vbar_k), in place over A
            double vs[NF][2];
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) {
                vs[nf][0] = vs[nf][1] = 0.0;
                for (int kk = 0; kk < K; ++kk) {
                    vs[nf][0] += vb[kk * NT + nf * 8 + 2 * t];
                    vs[nf][1] += vb[kk * NT + nf * 8 + 2 * t + 1];
                }
            }
#pragma unroll
            for (int r = 0; r < NBW; ++r) {
                const int b = snake_block(r, warp, nb16);
                if (b < 0) continue;
                wgemm_block<NT, false>(ly.W_m, KP / 4, 2 * b, 0, KP / 4, mubT, acc[r], sc_dummy, lane);
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) {
                        double2* p = reinterpret_cast<double2*>(cb.A + (size_t)(b * 16 + mf * 8 + g) * cb.ldn + n0 + nf * 8 + 2 * t);
                        const double2 av = *p;
                        *p = make_double2(acc[r][mf][nf][0] - 2.0 * vs[nf][0] * av.x, acc[r][mf][nf][1] - 2.0 * vs[nf][1] * av.y);
                    }
            }
        }
    }
}

// ==================================================================================================
// cond_bwd_b
// ==================================================================================================
template <int NT>
__global__ void __launch_bounds__(SK_THREADS) cond_bwd_b_kernel(LayerDev ly, ChunkBuffers cb, int ntiles,
                                                                double* esum_part) {
    constexpr int NF = NT / 8, STR = NT + 4;
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, Dp = ly.Dp, D = ly.D, XSTR = xs_stride(Dp), E = 1 + 2 * Dp;
    double* T = smem;
    double* Xs = T + (size_t)Mp * STR;
    double* xs2 = Xs + NT * XSTR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const double variance = ly.variance[0];
    double* my_part = esum_part + (size_t)blockIdx.x * Mp * E;
    const int nmy = my_block_count(warp, nb16);
    auto seg_of = [&](int i) { const int b = snake_block(i, warp, nb16); return Seg{ly.W_LinvT, 2 * b, 4 * b}; };
    WFrag wf;
    if (nmy > 0) wfrag_load(wf, seg_of(0), C4, seg_of(0).kb0, lane);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = (int64_t)tile * NT;
        load_tile_async<NT>(T, cb.A, Mp, cb.ldn, n0);
        stage_x<NT>(ly, cb, n0, Xs, xs2);
        cp_async_wait<0>();
        __syncthreads();
        for (int i = 0; i < nmy; ++i) {
            const int b = snake_block(i, warp, nb16);
            double acc[2][NF][2];
            zero_acc<NF>(acc);
            wgemm_seg<NT>(seg_of(i), C4, C4, T, acc, lane, wf, seg_of(i + 1 < nmy ? i + 1 : 0));   // upper triangular
#pragma unroll
            for (int mf = 0; mf < 2; ++mf) {
                double kv[NF][2];
                gen_kuf_block<NT>(ly, 2 * b + mf, Xs, xs2, variance, kv, lane);
                double e0 = 0.0;
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        acc[mf][nf][e] *= kv[nf][e];   // E = Kuf_bar .* Kuf
                        e0 += acc[mf][nf][e];
                    }
                double* p = my_part + (size_t)(b * 16 + mf * 8 + g) * E;
                // fire-and-forget reductions: each address has exactly ONE writer (this lane, this CTA's private slot), so
                // the accumulation order is fixed and the result deterministic; RED avoids the load-add-store round trip
                e0 = sum_over_t(e0);
                if (t == 0) atomicAdd(p, e0);
                for (int d = 0; d < D; ++d) {
                    double e1 = 0.0, e2 = 0.0;
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double x = Xs[(nf * 8 + 2 * t + e) * XSTR + d];
                            const double ex = acc[mf][nf][e] * x;
                            e1 += ex;
                            e2 += ex * x;
                        }
                    e1 = sum_over_t(e1);
                    e2 = sum_over_t(e2);
                    if (t == 0) {
                        atomicAdd(p + 1 + d, e1);
                        atomicAdd(p + 1 + Dp + d, e2);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ==================================================================================================
// host side
// ==================================================================================================
int stream_max_parts(const Launch& ln) { return ln.num_sms * 4; }

static int pick_nt(int Mp, size_t extra_bytes_nt32) {
    // widest supported tile whose [Mp x (NT+4)] operand fits in 227 KB (Mp=256, NT=32: 74 KB -> 3 CTAs/SM)
    const size_t cap = 227 * 1024;
    if ((size_t)Mp * 36 * 8 + extra_bytes_nt32 <= cap) return 32;
    if ((size_t)Mp * 20 * 8 + extra_bytes_nt32 <= cap) return 16;
    return 0;
}

template <typename KernelT>
static int persistent_grid(KernelT kernel, size_t smem, int ntiles, int cap, const Launch& ln) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, SK_THREADS, smem);
    if (occ < 1) occ = 1;
    int grid = ln.num_sms * occ;
    if (grid > ntiles) grid = ntiles;
    if (cap > 0 && grid > cap) grid = cap;
    return grid < 1 ? 1 : grid;
}

void cond_fwd_a(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln) {
    const int XSTR = xs_stride(ly.Dp);
    const int nt = pick_nt(ly.Mp, (size_t)(32 * XSTR + 32 + SK_WARPS * 32) * 8);
    auto launch = [&](auto kernel, int NT) {
        const size_t smem = ((size_t)ly.Mp * (NT + 4) + NT * XSTR + NT + SK_WARPS * NT) * sizeof(double);
        const int ntiles = (int)((cb.n + NT - 1) / NT);
        const int grid = persistent_grid(kernel, smem, ntiles, 0, ln);
        kernel<<<grid, SK_THREADS, smem, ln.stream>>>(ly, cb, ntiles);
        ln.tick();
    };
    if (nt == 32) launch(cond_fwd_a_kernel<32>, 32); else launch(cond_fwd_a_kernel<16>, 16);
}

void cond_fwd_b(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln) {
    const int nt = pick_nt(ly.Mp, (size_t)(SK_WARPS * KP * 32) * 8);
    auto launch = [&](auto kernel, int NT) {
        const size_t smem = ((size_t)ly.Mp * (NT + 4) + (size_t)SK_WARPS * KP * NT) * sizeof(double);
        const int ntiles = (int)((cb.n + NT - 1) / NT);
        const int grid = persistent_grid(kernel, smem, ntiles, 0, ln);
        kernel<<<grid, SK_THREADS, smem, ln.stream>>>(ly, cb, ntiles);
        ln.tick();
    };
    if (nt == 32) launch(cond_fwd_b_kernel<32>, 32); else launch(cond_fwd_b_kernel<16>, 16);
}

void cond_bwd_a(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln) {
    // double-buffered operand: two [Mp x (NT+4)] tiles must fit
    int nt = 0;
    if (2 * ((size_t)ly.Mp * 36 + KP * 36 + KP * 32) * 8 <= 227 * 1024) nt = 32;
    else if (2 * ((size_t)ly.Mp * 20 + KP * 20 + KP * 16) * 8 <= 227 * 1024) nt = 16;
    const int nbw = (ly.Mp / 16 + SK_WARPS - 1) / SK_WARPS;   // 16-row blocks per warp
    auto launch = [&](auto kernel, int NT) {
        const size_t smem = 2 * ((size_t)ly.Mp * (NT + 4) + KP * (NT + 4) + KP * NT) * sizeof(double);
        const int ntiles = (int)((cb.n + NT - 1) / NT);
        const int grid = persistent_grid(kernel, smem, ntiles, 0, ln);
        kernel<<<grid, SK_THREADS, smem, ln.stream>>>(ly, cb, ntiles);
        ln.tick();
    };
    if (nt == 32 && nbw <= 4) {   // accumulators of all row blocks live in registers: NBW*2*NF*2 doubles per lane
        if (nbw <= 1) launch(cond_bwd_a_kernel<32, 1>, 32);
        else if (nbw <= 2) launch(cond_bwd_a_kernel<32, 2>, 32);
        else launch(cond_bwd_a_kernel<32, 4>, 32);
    } else {
        if (nbw <= 4) launch(cond_bwd_a_kernel<16, 4>, 16);
        else if (nbw <= 8) launch(cond_bwd_a_kernel<16, 8>, 16);
        else launch(cond_bwd_a_kernel<16, 11>, 16);
    }
}

void cond_bwd_b(const LayerDev& ly, const ChunkBuffers& cb, double* esum_part, int nparts_cap, int* nparts,
                const Launch& ln) {
    const int XSTR = xs_stride(ly.Dp);
    const int nt = pick_nt(ly.Mp, (size_t)(32 * XSTR + 32) * 8);
    auto launch = [&](auto kernel, int NT) {
        const size_t smem = ((size_t)ly.Mp * (NT + 4) + NT * XSTR + NT) * sizeof(double);
        const int ntiles = (int)((cb.n + NT - 1) / NT);
        const int grid = persistent_grid(kernel, smem, ntiles, nparts_cap, ln);
        kernel<<<grid, SK_THREADS, smem, ln.stream>>>(ly, cb, ntiles, esum_part);
        ln.tick();
        if (grid > *nparts) *nparts = grid;
    };
    if (nt == 32) launch(cond_bwd_b_kernel<32>, 32); else launch(cond_bwd_b_kernel<16>, 16);
}

}  // namespace mgp
