// Replicated (per-rank) small-matrix work of one SVGP layer: Kuu, Cholesky, L^-1, operand packing,
// and the backward of all of it (Cholesky backward, kernel backward on Kuu, KL terms).
//
// Reference arithmetic being replaced:
//   Kuu            gpflow covariances.Kuu + SquaredExponential.K      (call site MixtureGPs/models.py:135)
//   chol / solve   gpflow base_conditional: cholesky, triangular_solve (call site models.py:141-143)
//   KL             gpflow gauss_kl, whitened                           (call site models.py:79)
//   backward       TF autodiff of the above                            (utils/training_utils.py:8-10)
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "exp_tab.h"
#include "kernels.h"

namespace mgp {

constexpr double JITTER = 1e-6;  // gpflow.config.default_jitter(), MixtureGPs/models.py:17,135

// --------------------------------------------------------------------------------------------------
// Zs = Z / lengthscales (row-major and fragment-major), |Zs_i|^2, 1/lengthscale
// --------------------------------------------------------------------------------------------------
__global__ void prep_z_kernel(LayerDev ly) {
    const int Mp = ly.Mp, Dp = ly.Dp, M = ly.M, D = ly.D;
    for (int d = threadIdx.x; d < Dp; d += blockDim.x)
        ly.inv_ls[d] = d < D ? 1.0 / ly.lengthscales[ly.n_ls == 1 ? 0 : d] : 0.0;
    for (int idx = threadIdx.x; idx < Mp * Dp; idx += blockDim.x) {
        const int i = idx / Dp, d = idx % Dp;
        double v = 0.0;
        if (i < M && d < D) v = ly.Z[(size_t)i * D + d] / ly.lengthscales[ly.n_ls == 1 ? 0 : d];  // Stationary.scale
        ly.Zs_rm[idx] = v;
        // left operand of the Kuf exponent's contraction, in units of ln2/64 (stream_kernels.cu::exp2_tab)
        ly.Zs_fm[wf_index(i, d, Dp)] = v * EXP_TAB_L;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Mp; i += blockDim.x) {
        double s = 0.0;
        for (int d = 0; d < Dp; ++d) {
            const double v = ly.Zs_rm[(size_t)i * Dp + d];
            s += v * v;
        }
        ly.zs2[i] = s;
        const double zh = (log(ly.variance[0]) - 0.5 * s) * EXP_TAB_L;
        ly.zh[i] = zh;
        if (D + 2 <= Dp) {   // kuf_fold (stream_kernels.cu): the exponent's row term rides in the padding columns of Zs_fm
            ly.Zs_fm[wf_index(i, D, Dp)] = i < M ? zh : 0.0;
            ly.Zs_fm[wf_index(i, D + 1, Dp)] = i < M ? EXP_TAB_L : 0.0;
        }
    }
}

// Kuu = variance * exp(-r2/2) + jitter I, r2 = -2 Zs Zs^T + (|zs_i|^2 + |zs_j|^2)   (square_distance, X2=None)
__global__ void kuu_kernel(LayerDev ly) {
    const int Mp = ly.Mp, Dp = ly.Dp, M = ly.M;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Mp * Mp) return;
    const int i = idx / Mp, j = idx % Mp;
    double v;
    if (i < M && j < M) {
        double dot = 0.0;
        for (int d = 0; d < Dp; ++d) dot += ly.Zs_rm[(size_t)i * Dp + d] * ly.Zs_rm[(size_t)j * Dp + d];
        const double r2 = -2.0 * dot + (ly.zs2[i] + ly.zs2[j]);
        v = ly.variance[0] * exp(-0.5 * r2);
        if (i == j) v += JITTER;
    } else {
        v = (i == j) ? 1.0 : 0.0;  // padding: decoupled unit block
    }
    ly.Kuu[idx] = v;
}

// --------------------------------------------------------------------------------------------------
// Blocked right-looking Cholesky, one CTA per matrix (matrix stays in L2; 32-wide panels).
//   (1) diagonal 32x32 block D: one warp, one row per lane in registers, warp shuffles for the pivots; the same warp
//       then forms V = D^-1 (lane j owns column j, forward substitution against D in shared memory) and keeps it for
//       the panel and for trinv_kernel (`Dinv`, [Mp/32][32][32])
//   (2) panel: P <- P V^T on DMMA, one 8-row block per warp step (was: one row per thread, a 528-FMA dependent chain
//       per row — the longest phase of a panel step)
//   (3) trailing update on DMMA, one 32x32 tile per warp.  The panel's rows come from a shared-memory copy the panel
//       phase leaves behind ([rows][36]: conflict-free 256-byte fragment reads) whenever it fits (Mp <= ~700): read from
//       global memory, a fragment load touches 8 cache lines and the update spent as many L1 wavefronts as DMMA clocks
//       (16 k clocks per round of 16 tiles for 8.2 k of DMMA).
// The diagonal blocks' explicit inverses cost cond(D) eps <= cond(L) eps in the panel, the same order as the explicit
// L^-1 every consumer of this factor uses anyway.
// --------------------------------------------------------------------------------------------------
constexpr int CHOL_THREADS = 512;

constexpr int CHOL_PSTR = 36;   // row stride of the shared-memory panel copy (== 4 mod 16 doubles)
template <bool panel_in_smem>
__global__ void __launch_bounds__(CHOL_THREADS, 1) chol_kernel(double* L, double* Dinv, int Mp, int* status) {
    extern __shared__ __align__(16) double Ps[];   // [Mp - 32][CHOL_PSTR] panel rows below the diagonal block (panel_in_smem)
    __shared__ double Dg[32][33];
    __shared__ double Vg[32][36];   // row stride == 4 mod 16 doubles: conflict-free B-fragment reads
    __shared__ double invd[32];     // 1 / D[r][r]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nblk = Mp / 32;
    for (int jb = 0; jb < nblk; ++jb) {
        const int j0 = jb * 32;
        if (warp == 0) {
            {
                double row[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) row[c] = L[(size_t)(j0 + lane) * Mp + j0 + c];
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const double piv = __shfl_sync(0xffffffffu, row[c], c);
                    if (!(piv > 0.0) && lane == 0) atomicOr(status, 1);
                    // one reciprocal square root instead of sqrt + division on the pivot chain (both are ~20-instruction
                    // sequences): d = piv * rsqrt(piv), l = row * rsqrt(piv), <= 2 ulp from the correctly rounded values
                    const double ri = rsqrt(piv);
                    const double l = (lane == c) ? piv * ri : row[c] * ri;
                    if (lane == c) invd[c] = ri;
                    row[c] = l;
#pragma unroll
                    for (int cc = c + 1; cc < 32; ++cc) {
                        const double lcc = __shfl_sync(0xffffffffu, l, cc);
                        row[cc] -= l * lcc;
                    }
                }
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const double v = (c <= lane) ? row[c] : 0.0;
                    Dg[lane][c] = v;
                    L[(size_t)(j0 + lane) * Mp + j0 + c] = v;
                }
            }
            __syncwarp();
            {   // V = D^-1, column `lane`:  y_r = (delta_{r,lane} - sum_{q<r} D[r][q] y_q) / D[r][r]   (y_q = 0 for q < lane).
                // Column-oriented: once y_q is final, every later row's sum takes its term — the dependent chain is one
                // FMA + one multiply per row instead of a whole row sum + a division.
                double y[32];
#pragma unroll
                for (int r = 0; r < 32; ++r) y[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    y[q] *= invd[q];
#pragma unroll
                    for (int r = q + 1; r < 32; ++r) y[r] = fma(-Dg[r][q], y[q], y[r]);
                }
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    Vg[r][lane] = y[r];
                    Dinv[((size_t)j0 + r) * 32 + lane] = y[r];
                }
            }
        }
        __syncthreads();
        // panel rows i >= j0 + 32:  X[i][c] = sum_p P[i][p] V[c][p]
        const int prow0 = j0 + 32;
        for (int u = warp; u < (Mp - prow0) / 8; u += CHOL_THREADS / 32) {
            double* rowp = L + (size_t)(prow0 + u * 8 + g) * Mp + j0;
            double a[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) a[kk] = rowp[kk * 4 + t];
            double acc[4][2];
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc[ni][0] = acc[ni][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma(acc[ni], a[kk], Vg[ni * 8 + g][kk * 4 + t]);
            __syncwarp();   // every lane of the warp has read its part of the 8 rows before they are overwritten
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                rowp[ni * 8 + 2 * t] = acc[ni][0];
                rowp[ni * 8 + 2 * t + 1] = acc[ni][1];
                if (panel_in_smem)
                    *reinterpret_cast<double2*>(Ps + (size_t)(u * 8 + g) * CHOL_PSTR + ni * 8 + 2 * t) = make_double2(acc[ni][0], acc[ni][1]);
            }
        }
        __syncthreads();
        const int nb = nblk - jb - 1;
        const int ntile = nb * (nb + 1) / 2;
        for (int tile = warp; tile < ntile; tile += CHOL_THREADS / 32) {
            int ti = (int)((sqrt(8.0 * tile + 1.0) - 1.0) * 0.5);
            while (ti * (ti + 1) / 2 > tile) --ti;
            while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti;
            const int tj = tile - ti * (ti + 1) / 2;
            const int r0 = j0 + 32 + ti * 32, c0 = j0 + 32 + tj * 32;
            double acc[4][4][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            if (panel_in_smem) {
                const double* pa = Ps + (size_t)(ti * 32 + g) * CHOL_PSTR + t;
                const double* pb = Ps + (size_t)(tj * 32 + g) * CHOL_PSTR + t;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    double a[4], b[4];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) a[mi] = pa[mi * 8 * CHOL_PSTR + kk * 4];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) b[ni] = pb[ni * 8 * CHOL_PSTR + kk * 4];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni], a[mi], b[ni]);
                }
            } else {
#pragma unroll 4
            for (int kk = 0; kk < 8; ++kk) {
                double a[4], b[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = L[(size_t)(r0 + mi * 8 + g) * Mp + j0 + kk * 4 + t];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) b[ni] = L[(size_t)(c0 + ni * 8 + g) * Mp + j0 + kk * 4 + t];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni], a[mi], b[ni]);
            }
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    double* p = L + (size_t)(r0 + mi * 8 + g) * Mp + c0 + ni * 8 + 2 * t;
                    p[0] -= acc[mi][ni][0];
                    p[1] -= acc[mi][ni][1];
                }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < Mp * Mp; idx += CHOL_THREADS) {
        const int i = idx / Mp, j = idx % Mp;
        if (j > i) L[idx] = 0.0;
    }
}

// --------------------------------------------------------------------------------------------------
// X = L^-1 by block forward substitution, one CTA per 32-column block of X (X pre-zeroed):
//   X_ij = V_i (delta_ij I - sum_{j <= p < i} L_ip X_pj),   V_i = (L_ii)^-1 from chol_kernel.
// The chain over i is serial, so a step has to be short: the block products L_ip X_pj of a step are dealt to the 8
// warps by p (the k-range is split, every warp forms a full 32 x 32 partial on DMMA from operands fetched in ONE
// round trip), the partials are summed through shared memory, and V_i R is another 16 DMMAs per warp.
// (A thread-per-output FMA loop over staged blocks took 4.9 k clocks per 32-deep chunk and 99 us in all: four
// dependent FP64 FMA chains per thread, ~35 clocks per link.)
// --------------------------------------------------------------------------------------------------
constexpr int TI_WARPS = 8;
constexpr int TI_SMEM_DOUBLES = TI_WARPS * 1024 + 32 * 36;

__global__ void __launch_bounds__(TI_WARPS * 32, 1) trinv_kernel(const double* L, const double* Dinv, double* X, int Mp) {
    extern __shared__ __align__(16) double ti_sm[];
    double* part = ti_sm;                    // [TI_WARPS][16 fragments][32 lanes][2]   C-fragment layout
    double* R = ti_sm + TI_WARPS * 1024;     // [32][36]
    const int jb = blockIdx.x, nblk = Mp / 32, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int smi = warp >> 1, sni = (warp & 1) * 2;   // the two output fragments (smi, sni), (smi, sni + 1) of the V_i R product
    for (int ib = jb; ib < nblk; ++ib) {
        double v[8];   // row block smi of V_i (lower triangular: k4-blocks past the diagonal are zero)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
            v[kk] = (kk < 2 * (smi + 1)) ? Dinv[((size_t)ib * 32 + smi * 8 + g) * 32 + kk * 4 + t] : 0.0;
        const int nch = ib - jb;
        if (warp < nch) {
            double acc[4][4][2];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
            for (int pb = jb + warp; pb < ib; pb += TI_WARPS) {
                double a[4][8], b[4][8];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        a[mi][kk] = L[(size_t)(ib * 32 + mi * 8 + g) * Mp + pb * 32 + kk * 4 + t];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        b[ni][kk] = X[(size_t)(pb * 32 + kk * 4 + t) * Mp + jb * 32 + ni * 8 + g];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni], a[mi][kk], b[ni][kk]);
            }
            double* mine = part + (size_t)warp * 1024 + lane * 2;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    *reinterpret_cast<double2*>(mine + (mi * 4 + ni) * 64) = make_double2(acc[mi][ni][0], acc[mi][ni][1]);
        }
        __syncthreads();
        const int np = nch < TI_WARPS ? nch : TI_WARPS;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int pos = tid + 256 * q, f = pos >> 6, l = (pos >> 1) & 31, e = pos & 1;
            double s = 0.0;
            for (int w = 0; w < np; ++w) s += part[(size_t)w * 1024 + pos];
            const int row = (f >> 2) * 8 + (l >> 2), col = (f & 3) * 8 + 2 * (l & 3) + e;
            R[row * 36 + col] = ((ib == jb && row == col) ? 1.0 : 0.0) - s;
        }
        __syncthreads();
        double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
            if (kk < 2 * (smi + 1)) {
                dmma(o[0], v[kk], R[(kk * 4 + t) * 36 + sni * 8 + g]);
                dmma(o[1], v[kk], R[(kk * 4 + t) * 36 + (sni + 1) * 8 + g]);
            }
        double* xo = X + (size_t)(ib * 32 + smi * 8 + g) * Mp + jb * 32 + sni * 8 + 2 * t;
        xo[0] = o[0][0]; xo[1] = o[0][1];
        xo[8] = o[1][0]; xo[9] = o[1][1];
        __syncthreads();   // the block just written is read (from global memory) by this CTA's later steps
    }
}

// --------------------------------------------------------------------------------------------------
// packing into fragment-major operands
// --------------------------------------------------------------------------------------------------
// W[r][c_off + c] = (transpose ? src[c*ld + r] : src[r*ld + c]) inside the valid source range, else 0
__global__ void pack_fm_kernel(double* dst, int Rp, int Cp, int c_off, int Ctot, const double* src, int ld,
                               int Rvalid, int Cvalid, int transpose, int64_t dst_bstride, int64_t src_bstride) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)Rp * Cp) return;
    const int r = (int)(idx / Cp), c = (int)(idx % Cp);
    const double* s = src + (int64_t)blockIdx.y * src_bstride;
    double* d = dst + (int64_t)blockIdx.y * dst_bstride;
    double v = 0.0;
    if (r < Rvalid && c < Cvalid) v = transpose ? s[(size_t)c * ld + r] : s[(size_t)r * ld + c];
    d[wf_index(r, c_off + c, Ctot)] = v;
}

static void pack_fm(double* dst, int Rp, int Cp, int c_off, int Ctot, const double* src, int ld, int Rvalid,
                    int Cvalid, bool transpose, int batch, int64_t dst_bstride, int64_t src_bstride,
                    const Launch& ln) {
    const int64_t n = (int64_t)Rp * Cp;
    dim3 grid((unsigned)((n + 255) / 256), batch);
    pack_fm_kernel<<<grid, 256, 0, ln.stream>>>(dst, Rp, Cp, c_off, Ctot, src, ld, Rvalid, Cvalid, transpose ? 1 : 0,
                                                dst_bstride, src_bstride);
    ln.tick();
}

// Lq_rm[k] = tril(q_sqrt[k]) zero padded to Mp x Mp   (band_part(q_sqrt, -1, 0))
__global__ void lq_clean_kernel(LayerDev ly) {
    const int Mp = ly.Mp, M = ly.M;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)ly.K * Mp * Mp) return;
    const int k = (int)(idx / ((int64_t)Mp * Mp));
    const int rem = (int)(idx % ((int64_t)Mp * Mp));
    const int i = rem / Mp, j = rem % Mp;
    ly.Lq_rm[idx] = (i < M && j <= i) ? ly.q_sqrt[((size_t)k * M + i) * M + j] : 0.0;
}

// Q_rm = [Q_0 | ... | Q_{K-1}]  ([Mp, K*Mp] row-major),  Q_k = 2 (T1[k] - I_M)
__global__ void q_finish_kernel(LayerDev ly) {
    const int Mp = ly.Mp, M = ly.M, K = ly.K;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)K * Mp * Mp) return;
    const int k = (int)(idx / ((int64_t)Mp * Mp));
    const int rem = (int)(idx % ((int64_t)Mp * Mp));
    const int i = rem / Mp, j = rem % Mp;
    ly.Q_rm[(size_t)i * K * Mp + (size_t)k * Mp + j] = 2.0 * (ly.T1[idx] - ((i == j && i < M) ? 1.0 : 0.0));
}

// The replicated per-step pre-compute of a layer is two independent chains (api.cu runs them on different streams):
//   precompute_chol : Z / lengthscales, Kuu, its Cholesky factor, L^-1 and its packed forms — a latency-bound chain of
//                     single-CTA kernels, the critical path in front of the first streaming kernel;
//   precompute_lq   : everything that depends on q_mu / q_sqrt only (packed Lq_k^T, Lq_k, q_mu, Q_k = 2 (Lq_k Lq_k^T - I)).
void precompute_chol(const LayerDev& ly, bool need_bwd, int* d_status, const Launch& ln) {
    const int Mp = ly.Mp;
    const int64_t mm = (int64_t)Mp * Mp;
    prep_z_kernel<<<1, 1024, 0, ln.stream>>>(ly);
    kuu_kernel<<<(unsigned)((mm + 255) / 256), 256, 0, ln.stream>>>(ly);
    ln.tick(2);
    cudaMemcpyAsync(ly.L, ly.Kuu, sizeof(double) * mm, cudaMemcpyDeviceToDevice, ln.stream);
    {
        const size_t panel_bytes = (size_t)(Mp - 32) * CHOL_PSTR * sizeof(double);
        const int panel_in_smem = panel_bytes <= (size_t)200 * 1024 && getenv("MGP_CHOL_PANEL_GLOBAL") == nullptr;
        if (panel_in_smem) {
            cudaFuncSetAttribute(chol_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)panel_bytes);
            chol_kernel<true><<<1, CHOL_THREADS, panel_bytes, ln.stream>>>(ly.L, ly.Dinv, Mp, d_status);
        } else {
            chol_kernel<false><<<1, CHOL_THREADS, 0, ln.stream>>>(ly.L, ly.Dinv, Mp, d_status);
        }
    }
    cudaMemsetAsync(ly.Linv, 0, sizeof(double) * mm, ln.stream);
    cudaFuncSetAttribute(trinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TI_SMEM_DOUBLES * sizeof(double)));
    trinv_kernel<<<Mp / 32, TI_WARPS * 32, TI_SMEM_DOUBLES * sizeof(double), ln.stream>>>(ly.L, ly.Dinv, ly.Linv, Mp);
    ln.tick(2);
    pack_fm(ly.W_Linv, Mp, Mp, 0, Mp, ly.Linv, Mp, Mp, Mp, false, 1, 0, 0, ln);
    if (need_bwd) pack_fm(ly.W_LinvT, Mp, Mp, 0, Mp, ly.Linv, Mp, Mp, Mp, true, 1, 0, 0, ln);
}

void precompute_lq(const LayerDev& ly, bool need_bwd, const Launch& ln) {
    const int Mp = ly.Mp, K = ly.K;
    const int64_t mm = (int64_t)Mp * Mp;
    lq_clean_kernel<<<(unsigned)((K * mm + 255) / 256), 256, 0, ln.stream>>>(ly);
    ln.tick();
    pack_fm(ly.W_LqT, Mp, Mp, 0, Mp, ly.Lq_rm, Mp, Mp, Mp, true, K, mm, mm, ln);
    pack_fm(ly.W_mT, 16, Mp, 0, Mp, ly.q_mu, K, K, ly.M, true, 1, 0, 0, ln);
    if (need_bwd) {
        gemm_small(Mp, Mp, Mp, 1.0, ly.Lq_rm, Mp, mm, false, ly.Lq_rm, Mp, mm, true, 0.0, ly.T1, Mp, mm, K, ln);
        q_finish_kernel<<<(unsigned)((K * mm + 255) / 256), 256, 0, ln.stream>>>(ly);
        ln.tick();
        pack_fm(ly.W_Lq, Mp, Mp, 0, Mp, ly.Lq_rm, Mp, Mp, Mp, false, K, mm, mm, ln);
        pack_fm(ly.W_m, Mp, KP, 0, KP, ly.q_mu, K, ly.M, K, false, 1, 0, 0, ln);
    }
}

void precompute_layer(const LayerDev& ly, bool need_bwd, int* d_status, const Launch& ln) {
    precompute_chol(ly, need_bwd, d_status, ln);
    precompute_lq(ly, need_bwd, ln);
}

// --------------------------------------------------------------------------------------------------
// deterministic reduction of per-CTA partial sums
// --------------------------------------------------------------------------------------------------
__global__ void reduce_partials_kernel(double* dst, const double* src, int64_t n, int nparts, int64_t stride,
                                       int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = accumulate ? dst[i] : 0.0;
    for (int p = 0; p < nparts; ++p) s += src[(int64_t)p * stride + i];
    dst[i] = s;
}

// Many parts, few elements (the E-sums: one part per cond_bwd_b CTA, Mp (1 + 2 Dp) elements): a thread per element
// would walk hundreds of strided loads one after another.  Here a CTA takes 32 elements x 8 part lanes; lane py sums
// parts py, py + 8, ... and the 8 lane sums are added in a fixed order, so the result stays deterministic.
__global__ void __launch_bounds__(256) reduce_partials_tall_kernel(double* dst, const double* src, int64_t n, int nparts,
                                                                   int64_t stride, int accumulate) {
    __shared__ double part[8][33];
    const int ex = threadIdx.x & 31, py = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 32 + ex;
    double s0 = 0.0, s1 = 0.0;
    if (i < n) {
        int p = py;
        for (; p + 8 < nparts; p += 16) {
            s0 += src[(int64_t)p * stride + i];
            s1 += src[(int64_t)(p + 8) * stride + i];
        }
        if (p < nparts) s0 += src[(int64_t)p * stride + i];
    }
    part[py][ex] = s0 + s1;
    __syncthreads();
    if (py == 0 && i < n) {
        double s = accumulate ? dst[i] : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += part[q][ex];
        dst[i] = s;
    }
}

void reduce_partials(double* dst, const double* src, int64_t n, int nparts, int64_t stride, bool accumulate,
                     const Launch& ln) {
    if (n <= 0) return;
    if (nparts >= 64 && n <= 65536) {
        reduce_partials_tall_kernel<<<(unsigned)((n + 31) / 32), 256, 0, ln.stream>>>(dst, src, n, nparts, stride,
                                                                                      accumulate ? 1 : 0);
        ln.tick();
        return;
    }
    reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ln.stream>>>(dst, src, n, nparts, stride,
                                                                               accumulate ? 1 : 0);
    ln.tick();
}

// --------------------------------------------------------------------------------------------------
// replicated backward
// --------------------------------------------------------------------------------------------------
// Sfull[k][i][j] = S[k][max(i,j)][min(i,j)]
__global__ void symmetrize_kernel(double* Sfull, const double* S, int Mp, int K) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)K * Mp * Mp) return;
    const int64_t mm = (int64_t)Mp * Mp;
    const int k = (int)(idx / mm);
    const int rem = (int)(idx % mm);
    const int i = rem / Mp, j = rem % Mp;
    Sfull[idx] = (j <= i) ? S[idx] : S[(int64_t)k * mm + (size_t)j * Mp + i];
}

// dELBO/dq_sqrt[k] = tril(2 S_k Lq_k) + c (Lq_k - diag(1/diag Lq_k)),   c = -1/num_data   (T1 = S_k Lq_k)
__global__ void gqsqrt_kernel(LayerDev ly, double kl_coef, double* gqsqrt) {
    const int M = ly.M, Mp = ly.Mp;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)ly.K * M * M) return;
    const int k = (int)(idx / ((int64_t)M * M));
    const int rem = (int)(idx % ((int64_t)M * M));
    const int i = rem / M, j = rem % M;
    double v = 0.0;
    if (j <= i) {
        const size_t p = ((size_t)k * Mp + i) * Mp + j;
        const double lq = ly.Lq_rm[p];
        v = 2.0 * ly.T1[p] + kl_coef * (lq - (i == j ? 1.0 / lq : 0.0));
    }
    gqsqrt[idx] = v;
}

__global__ void gqmu_kernel(LayerDev ly, const double* mraw, double kl_coef, double* gqmu) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ly.M * ly.K) return;
    const int i = idx / ly.K, k = idx % ly.K;
    gqmu[idx] = mraw[(size_t)i * KP + k] + kl_coef * ly.q_mu[idx];
}

// T2 = tril(sum_k T1[k] + q_mu mraw^T)     (T1[k] = Q_k S_k, one batched GEMM instead of one with a K Mp long k-loop)
__global__ void t_finish_kernel(LayerDev ly, const double* mraw) {
    const int Mp = ly.Mp, M = ly.M, K = ly.K;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Mp * Mp) return;
    const int i = idx / Mp, j = idx % Mp;
    double v = 0.0;
    if (j <= i) {
        for (int k = 0; k < K; ++k) v += ly.T1[(size_t)k * Mp * Mp + idx];
        if (i < M && j < M)
            for (int k = 0; k < K; ++k) v += ly.q_mu[(size_t)i * K + k] * mraw[(size_t)j * KP + k];
    }
    ly.T2[idx] = v;
}

// in place: X = -tril(X)
__global__ void neg_tril_kernel(double* X, int Mp) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Mp * Mp) return;
    const int i = idx / Mp, j = idx % Mp;
    X[idx] = (j <= i) ? -X[idx] : 0.0;
}

// dst = Phi(src) + Phi(src)^T, Phi = lower triangle with halved diagonal
__global__ void phi_sym_kernel(double* dst, const double* src, int Mp) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Mp * Mp) return;
    const int i = idx / Mp, j = idx % Mp;
    dst[idx] = (i >= j) ? src[(size_t)i * Mp + j] : src[(size_t)j * Mp + i];
}

// kernel backward on Kuu: one CTA per inducing row a.  Kbar2 = Linv^T (P + P^T) Linv = 2 * Kuu_bar.
// rowout[a] = [ dZs_a[0..Dp) | dls_a[0..Dp) | dvar_a ]   (scaled coordinates)
__global__ void __launch_bounds__(128) kuu_bwd_kernel(LayerDev ly, const double* Kbar2, double* rowout) {
    __shared__ double red[32];
    const int a = blockIdx.x, M = ly.M, Mp = ly.Mp, Dp = ly.Dp, D = ly.D;
    const double var = ly.variance[0];
    const int W = 2 * Dp + 1;
    for (int q = 0; q < W; ++q) {
        double s = 0.0;
        for (int j = threadIdx.x; j < M; j += blockDim.x) {
            double kc = ly.Kuu[(size_t)a * Mp + j];
            if (j == a) kc -= JITTER;
            const double gk = 0.5 * Kbar2[(size_t)a * Mp + j] * kc;  // Kuu_bar_aj * K_aj
            if (q < Dp) {
                if (q < D) {
                    const double dz = ly.Zs_rm[(size_t)a * Dp + q] - ly.Zs_rm[(size_t)j * Dp + q];
                    s += -2.0 * gk * dz;  // both arguments of k(Z,Z); Kuu_bar symmetric
                }
            } else if (q < 2 * Dp) {
                const int d = q - Dp;
                if (d < D) {
                    const double dz = ly.Zs_rm[(size_t)a * Dp + d] - ly.Zs_rm[(size_t)j * Dp + d];
                    s += gk * dz * dz;
                }
            } else {
                s += gk / var;
            }
        }
        s = block_sum(s, red);
        if (threadIdx.x == 0) rowout[(size_t)a * W + q] = s;
        __syncthreads();
    }
}

// gauss_kl (whitened): 0.5 [ sum q_mu^2 - M K - sum log diag(Lq)^2 + sum Lq^2 ]
// stage 1: CTA (k, c) sums the terms of rows i = c mod KL_NC of Lq_k (k == K: the q_mu term); stage 2: fixed-order total.
constexpr int KL_NC = 4;
__global__ void __launch_bounds__(1024) kl_part_kernel(LayerDev ly, double* part) {
    __shared__ double red[32];
    const int M = ly.M, Mp = ly.Mp, K = ly.K, k = blockIdx.x, c = blockIdx.y;
    double s = 0.0;
    if (k < K) {
        const double* Lq = ly.Lq_rm + (size_t)k * Mp * Mp;
        // flat walk over the M x M square (fixed thread -> element map, so the sum order is fixed): a row-per-iteration
        // loop kept a quarter of the CTA busy and paid one L2 round trip per row (38 us at M = 256)
#pragma unroll 4
        for (int idx = c * blockDim.x + threadIdx.x; idx < M * M; idx += KL_NC * blockDim.x) {
            const int i = idx / M, j = idx - i * M;
            if (j <= i) {
                const double v = Lq[(size_t)i * Mp + j];
                s += v * v;
                if (i == j) s -= log(v * v);
            }
        }
    } else {
        for (int idx = c * blockDim.x + threadIdx.x; idx < M * K; idx += KL_NC * blockDim.x) {
            const double v = ly.q_mu[idx];
            s += v * v;
        }
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) part[k * KL_NC + c] = s;
}

__global__ void kl_total_kernel(LayerDev ly, const double* part, double* kl_out) {
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < (ly.K + 1) * KL_NC; ++k) s += part[k];
        kl_out[0] = 0.5 * (s - (double)ly.M * (double)ly.K);
    }
}

static void kl_launch(const LayerDev& ly, double* kl_out, const Launch& ln) {
    double* part = ly.zs2 + ly.Mp;   // zs2 is allocated with 2 * Mp + 16 doubles (api.cu); second half = scratch (>= 36)
    kl_part_kernel<<<dim3(ly.K + 1, KL_NC), 1024, 0, ln.stream>>>(ly, part);
    kl_total_kernel<<<1, 32, 0, ln.stream>>>(ly, part, kl_out);
    ln.tick(2);
}

// final assembly of dZ, dlengthscales, dvariance from the streamed sums (esum), the Kuu path (rowout) and Knn
__global__ void __launch_bounds__(256) assemble_kernel(LayerDev ly, const double* esum, const double* rowout,
                                                       const double* sumv, double* gZ, double* gvar, double* gls) {
    __shared__ double red[32];
    const int M = ly.M, D = ly.D, Dp = ly.Dp;
    const int E = 1 + 2 * Dp, W = 2 * Dp + 1;
    for (int idx = threadIdx.x; idx < M * D; idx += blockDim.x) {
        const int a = idx / D, d = idx % D;
        const double zs = ly.Zs_rm[(size_t)a * Dp + d];
        const double dzs = (esum[(size_t)a * E + 1 + d] - zs * esum[(size_t)a * E]) + rowout[(size_t)a * W + d];
        gZ[idx] = dzs * ly.inv_ls[d];
    }
    double total_ls = 0.0;
    for (int d = 0; d < D; ++d) {
        double s = 0.0;
        for (int a = threadIdx.x; a < M; a += blockDim.x) {
            const double zs = ly.Zs_rm[(size_t)a * Dp + d];
            s += esum[(size_t)a * E + 1 + Dp + d] - 2.0 * zs * esum[(size_t)a * E + 1 + d] +
                 zs * zs * esum[(size_t)a * E] + rowout[(size_t)a * W + Dp + d];
        }
        s = block_sum(s, red);
        if (threadIdx.x == 0) {
            const double g = s * ly.inv_ls[d];
            if (ly.n_ls == 1) total_ls += g; else gls[d] = g;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && ly.n_ls == 1) gls[0] = total_ls;
    double s = 0.0;
    for (int a = threadIdx.x; a < M; a += blockDim.x) s += esum[(size_t)a * E] / ly.variance[0] + rowout[(size_t)a * W + 2 * Dp];
    s = block_sum(s, red);
    if (threadIdx.x == 0) gvar[0] = s + sumv[0];
}

void prior_kl_layer(const LayerDev& ly, double* kl_out, const Launch& ln) {
    const int64_t mm = (int64_t)ly.Mp * ly.Mp;
    lq_clean_kernel<<<(unsigned)((ly.K * mm + 255) / 256), 256, 0, ln.stream>>>(ly);
    ln.tick();
    kl_launch(ly, kl_out, ln);
}

void finish_layer(const LayerDev& ly, const double* S_lower, const double* mraw, const double* esum,
                  const double* sumv, double kl_coef, double* gZ, double* gqmu, double* gqsqrt, double* gvar,
                  double* gls, double* kl_out, const Launch& ln) {   // kl_out == nullptr: the KL term is already there
    const int Mp = ly.Mp, K = ly.K, M = ly.M;
    const int64_t mm = (int64_t)Mp * Mp;
    const unsigned gmm = (unsigned)((mm + 255) / 256);
    symmetrize_kernel<<<(unsigned)((K * mm + 255) / 256), 256, 0, ln.stream>>>(ly.Sfull, S_lower, Mp, K);
    ln.tick();
    // dq_sqrt
    gemm_small(Mp, Mp, Mp, 1.0, ly.Sfull, Mp, mm, false, ly.Lq_rm, Mp, mm, false, 0.0, ly.T1, Mp, mm, K, ln);
    gqsqrt_kernel<<<(unsigned)(((int64_t)K * M * M + 255) / 256), 256, 0, ln.stream>>>(ly, kl_coef, gqsqrt);
    gqmu_kernel<<<(M * K + 255) / 256, 256, 0, ln.stream>>>(ly, mraw, kl_coef, gqmu);
    ln.tick(2);
    // T = tril( sum_k Q_k S_k + q_mu mraw^T )  ( = Abar A^T )
    // (T1 = S_k Lq_k has been consumed by gqsqrt_kernel above; it now takes the K partial products Q_k S_k)
    gemm_small(Mp, Mp, Mp, 1.0, ly.Q_rm, K * Mp, Mp, false, ly.Sfull, Mp, mm, false, 0.0, ly.T1, Mp, mm, K, ln);
    t_finish_kernel<<<gmm, 256, 0, ln.stream>>>(ly, mraw);
    // Lbar = -tril(L^-T T)
    gemm_small(Mp, Mp, Mp, 1.0, ly.Linv, Mp, 0, true, ly.T2, Mp, 0, false, 0.0, ly.T3, Mp, 0, 1, ln);
    neg_tril_kernel<<<gmm, 256, 0, ln.stream>>>(ly.T3, Mp);
    // Cholesky backward (Murray 2016): Kbar = 1/2 L^-T (P + P^T) L^-1, P = Phi(L^T Lbar)
    gemm_small(Mp, Mp, Mp, 1.0, ly.L, Mp, 0, true, ly.T3, Mp, 0, false, 0.0, ly.T2, Mp, 0, 1, ln);
    phi_sym_kernel<<<gmm, 256, 0, ln.stream>>>(ly.T3, ly.T2, Mp);
    gemm_small(Mp, Mp, Mp, 1.0, ly.Linv, Mp, 0, true, ly.T3, Mp, 0, false, 0.0, ly.T2, Mp, 0, 1, ln);
    gemm_small(Mp, Mp, Mp, 1.0, ly.T2, Mp, 0, false, ly.Linv, Mp, 0, false, 0.0, ly.T3, Mp, 0, 1, ln);
    ln.tick(3);
    // kernel backward on Kuu
    kuu_bwd_kernel<<<M, 128, 0, ln.stream>>>(ly, ly.T3, ly.rowout);
    assemble_kernel<<<1, 256, 0, ln.stream>>>(ly, esum, ly.rowout, sumv, gZ, gvar, gls);
    ln.tick(2);
    if (kl_out) kl_launch(ly, kl_out, ln);
}

// KL term of a layer whose Lq_rm is current (precompute_layer has run)
void prior_kl_precomputed(const LayerDev& ly, double* kl_out, const Launch& ln) { kl_launch(ly, kl_out, ln); }

}  // namespace mgp
