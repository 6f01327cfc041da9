// Host-side launchers of the libmgp kernels (one translation unit per group).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "layer.h"

namespace mgp {

struct Launch {
    cudaStream_t stream;
    int64_t* count;  // kernels launched so far (mgp_launch_count)
    int num_sms;
    inline void tick(int n = 1) const { *count += n; }
};

// ---- precompute.cu : replicated O(M^2 D + M^3) work ------------------------------------------------
// Kuu, Cholesky, L^-1 and the fragment-major operands.  need_bwd adds Q_k = 2(Lq_k Lq_k^T - I).
void precompute_layer(const LayerDev& ly, bool need_bwd, int* d_status, const Launch& ln);
void precompute_chol(const LayerDev& ly, bool need_bwd, int* d_status, const Launch& ln);   // Kuu, L, L^-1 chain
void precompute_lq(const LayerDev& ly, bool need_bwd, const Launch& ln);                    // q_mu / q_sqrt chain
// gauss_kl(q_mu, q_sqrt) of a whitened layer (only needs Lq_rm allocated)
void prior_kl_layer(const LayerDev& ly, double* kl_out, const Launch& ln);
void prior_kl_precomputed(const LayerDev& ly, double* kl_out, const Launch& ln);
// dst[i] = sum_s src[s*stride + i], i < n (deterministic order)
void reduce_partials(double* dst, const double* src, int64_t n, int nparts, int64_t stride, bool accumulate,
                     const Launch& ln);
// replicated backward: from S_k, mraw, esum (reduce buffer) to gradients w.r.t. constrained values
void finish_layer(const LayerDev& ly, const double* S_lower, const double* mraw, const double* esum,
                  const double* sumv, double kl_coef, double* gZ, double* gqmu, double* gqsqrt, double* gvar,
                  double* gls, double* kl_out, const Launch& ln);

// ---- gemm_small.cu : batched strided FP64 DMMA GEMM on M x M operands ------------------------------
// C[b] = alpha * opA(A[b]) * opB(B[b]) + beta * C[b];  all row-major with leading dimension ld.
void gemm_small(int m, int n, int k, double alpha, const double* A, int lda, int64_t strideA, bool transA,
                const double* B, int ldb, int64_t strideB, bool transB, double beta, double* C, int ldc,
                int64_t strideC, int batch, const Launch& ln);

// ---- stream_kernels.cu : the N-streaming DMMA kernels ------------------------------------------------
struct ChunkBuffers {
    int64_t n;        // valid points in this chunk
    int64_t ldn;      // padded length (multiple of 64) of the per-point arrays
    int tw;           // tile width of the tile-major workspace (layer_tile_width): 32 or 16 points
    int64_t tiles_cap;  // ldn / tw: tiles per component in Bk
    const double* X;  // [n, D] this chunk's rows
    double* A;        // tile-major [ldn/tw][Mp][tw+4]  A = L^-1 Kuf  (overwritten by Abar in the backward)
    double* Bk;       // [K] x tile-major: B_k = Lq_k^T A, kept for the backward (null on forward-only paths)
    double* fmean;    // [ldn, K]
    double* fvar;     // [ldn, K]
    double* mubar;    // [ldn, K]   d/d fmean
    double* vbar;     // [ldn, K]   d/d fvar
};
// A = L^-1 Kuf(tile) -> cb.A, cb.asq
void cond_fwd_a(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln);
// fmean = A^T q_mu, fvar = variance - |a|^2 + |Lq_k^T a|^2
void cond_fwd_b(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln);
// Abar = sum_k Lq_k (2 B_k diag(vbar_k)) + q_mu mubar^T - 2 A diag(sum_k vbar_k)   (in place over cb.A)
void cond_bwd_a(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln);
// Kuf_bar = L^-T Abar; E = Kuf_bar .* Kuf; esum partial += E [1, xs, xs^2]
void cond_bwd_b(const LayerDev& ly, const ChunkBuffers& cb, double* esum_part, int nparts_cap, int* nparts,
                const Launch& ln);
int stream_max_parts(const Launch& ln);
// tile width (points) of a layer's tile-major workspace: 32 while two [Mp x 36] tiles fit in shared memory, else 16
int layer_tile_width(int Mp, int Dp, int K);

// ---- syrk.cu : S_k += A diag(vbar_k) A^T ----------------------------------------------------------------
// One CTA per work item: 64x64 tile pair (I >= J) of S, component group kbase..kbase+3, points [pbeg, pend), and the
// slot of `part` it accumulates into.  The plan is built on the host (syrk_make_plan) whenever (Mp, K, n) change.
struct SyrkWork {
    int I, J, kbase, slot;
    int64_t pbeg, pend;
};
// part: [nslots, K, Mp, Mp]; each CTA accumulates into its own slot (deterministic); fragments with col <= row of the
// lower 64x64 tiles only.  mraw_part: [nslots, Mp, KP] += A mubar (computed by the J == 0 tile column, which sees
// every row of A).
void syrk_accumulate(const LayerDev& ly, const ChunkBuffers& cb, const SyrkWork* d_plan, int nwork, double* part,
                     double* mraw_part, const Launch& ln);
int syrk_num_splits(int Mp, int K, const Launch& ln);   // number of slots
}  // namespace mgp
#include <vector>
namespace mgp {
int syrk_make_plan(int Mp, int K, int64_t n, int kc, int num_sms, std::vector<SyrkWork>& plan);

// ---- mc_pass.cu : fused Monte-Carlo likelihood pass, forward + adjoints ------------------------------
struct McArgs {
    int model, lik, S, K;
    double temperature;
    double squash;        // RobustMax CDF squash (MGP_ROBUSTMAX_CDF_SQUASH unless mgp_set_robustmax_squash changed it)
    double inv_n_global;
    int64_t n;            // valid points in the chunk
    int64_t ldn;          // padded points (adjoints of padded points are written as 0)
    int64_t n_local;      // points in this shard (stride of the noise arrays)
    int64_t chunk_offset; // first point of the chunk within the shard
    const double* Y;      // [n] chunk rows
    const double *fmean_p, *fvar_p, *fmean_a, *fvar_a;  // [ldn, K]
    double *mubar_p, *vbar_p, *mubar_a, *vbar_a;        // [ldn, K]
    const double* lik_var;         // [K] or null
    const double* assign_lik_var;  // [K] or null
    const double *z, *u;           // [S, n_local, K] or null (Philox)
    uint64_t seed;
    int64_t point_offset;          // global index of the shard's first point
};
// per-block partial sums: [nblocks, MC_NPART]
constexpr int MC_NPART = 20;  // data, likvar[8], alikvar[8], sumv_p, sumv_a, pad
int mc_num_blocks(int64_t ldn);
void mc_pass(const McArgs& a, double* block_part, const Launch& ln);
// fold block partials into the reduce-buffer header (accumulating)
void mc_fold(const double* block_part, int nblocks, double* rb_header, const Launch& ln);

// ---- predict paths ------------------------------------------------------------------------------------
void predict_assign_kernel(const double* fmean, int64_t n, int K, double* probs, int64_t* argmax, const Launch& ln);
void predict_y_kernel(const double* fmean, const double* fvar, int64_t n, int K, int lik, const double* lik_var,
                      double squash, double* mean, double* var, const Launch& ln);
struct SampleArgs {
    int S, K, lik;
    double temperature;
    double squash;
    int64_t n;
    const double *fmean_p, *fvar_p, *fmean_a, *fvar_a;  // [n, K]
    const double* lik_var;
    const double *z, *u, *z_pred;  // [S, n, K]
    uint64_t seed;
    int64_t point_offset;
    double *samples_y, *samples_f;  // [S, n]
};
void predict_samples_kernel(const SampleArgs& a, const Launch& ln);
// W [S, n, K]: relaxed one-hot sample of the assign layer's logits (uses fmean_a, fvar_a, z, u / Philox of SampleArgs)
void w_sample_kernel(const SampleArgs& a, double* W_out, const Launch& ln);
// out [n] = logsumexp_S(sum_k W ve) - log S for the expert likelihood `lik` (0 Gaussian, 1 MultiClass RobustMax)
void e_log_p_y_kernel(const double* fmean, const double* fvar, const double* Y, const double* lik_var, int lik,
                      double squash, const double* W, int S, int64_t n, int K, double* out, const Launch& ln);

// stand-alone likelihood methods (mode 0 variational expectations, 1 Gaussian log-prob, 2 Gaussian predictive log density)
void lik_eval_kernel(int mode, int lik, const double* lik_var, double squash, const double* Fmu, const double* Fvar,
                     const double* Y, int64_t rows, int64_t y_period, int K, double* out, const Launch& ln);
// Philox test entries: ctr_key [n][6] = counter[4], key[2] -> out [n][4]; raw z / u draws [S, n, K] of throughput mode
void philox_kat_kernel(const uint32_t* ctr_key, int n, uint32_t* out, const Launch& ln);
void philox_draws_kernel(uint64_t seed, int64_t point_offset, int64_t n, int S, int K, int stream, double* z, double* u,
                         const Launch& ln);

}  // namespace mgp
