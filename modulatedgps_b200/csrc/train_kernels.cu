// The callers either side of the ELBO step, on the device:
//   f1  fused Adam update on the UNCONSTRAINED variables, bijector chain rule included
//       (utils/training_utils.py:6-10: tf.optimizers.Adam(lr).minimize over model.trainable_variables; the
//       gradients w.r.t. CONSTRAINED values come straight from mgp_elbo_fwd_bwd)
//   f2  minibatch gather for a shuffled epoch (demos/demo_tf2.py:53-56: shuffle(N).batch(B).repeat())
//   f4  Lloyd k-means for the inducing-point initialisation (demos/demo_tf2.py:39: scipy.cluster.vq.kmeans)
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>

#include "../../include/mgp.h"
#include "common.cuh"
#include "kernels.h"

namespace mgp {

// ---- f1: Adam --------------------------------------------------------------------------------------------------
// TF 2.10 Keras Adam (non-amsgrad), _resource_apply_dense:
//   lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t);  m += (g - m)(1 - beta1);  v += (g^2 - v)(1 - beta2);
//   theta -= lr_t * m / (sqrt(v) + eps)
// g = grad_scale * d(ELBO)/d(unconstrained) with the chain rule of the slot's bijector applied here:
//   identity: g_c;  softplus: g_c * sigmoid(theta);  fill-triangular: g_c[gather[i]] (a permutation).
struct AdamTable {
    mgp_adam_slot s[MGP_ADAM_MAX_SLOTS];
};

// guard: device scalar (the step's ELBO) — a step whose ELBO is not finite (failed Cholesky, stale precompute) must not
// reach theta, m, v: the reference's tf.linalg.cholesky raises before the optimiser runs.
__global__ void __launch_bounds__(256) adam_kernel(AdamTable tab, double grad_scale, double lr_t, double beta1, double beta2,
                                                   double eps, const double* guard) {
    if (guard != nullptr && !isfinite(*guard)) return;
    const mgp_adam_slot sl = tab.s[blockIdx.y];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sl.n; i += (int64_t)gridDim.x * blockDim.x) {
        const double th = sl.theta[i];
        double g = grad_scale * (sl.gather ? sl.grad[sl.gather[i]] : sl.grad[i]);
        if (sl.transform == MGP_TRANSFORM_SOFTPLUS) g *= 1.0 / (1.0 + exp(-th));   // d softplus / d theta = sigmoid
        const double m = sl.m[i] + (g - sl.m[i]) * (1.0 - beta1);
        const double v = sl.v[i] + (g * g - sl.v[i]) * (1.0 - beta2);
        sl.m[i] = m;
        sl.v[i] = v;
        sl.theta[i] = th - lr_t * m / (sqrt(v) + eps);
    }
}

// ---- f2: gather -----------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const double* X, const double* Y, const int64_t* idx, int64_t B, int D, double* Xb,
                                   double* Yb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * (D + 1)) return;
    const int64_t r = i / (D + 1);
    const int c = (int)(i - r * (D + 1));
    const int64_t src = idx[r];
    if (c < D) Xb[r * D + c] = X[src * D + c];
    else Yb[r] = Y[src];
}

// ---- gpflow.utilities.triangular() = TFP FillTriangular ----------------------------------------
// vector x of length n = m(m+1)/2  ->  concat([x[m:], reversed(x)]) reshaped [m, m], lower band.
__device__ __forceinline__ int64_t fill_tri_src(int i, int j, int m, int64_t n) {
    const int64_t t = (int64_t)i * m + j;
    return t < n - m ? m + t : n - 1 - (t - (n - m));
}
// forward: every element of the [batch, m, m] output is written (zeros above the diagonal) — no separate fill
__global__ void fill_tri_fwd_kernel(const double* x, int64_t batch, int m, double* out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t mm = (int64_t)m * m, n = (int64_t)m * (m + 1) / 2;
    if (idx >= batch * mm) return;
    const int64_t b = idx / mm;
    const int r = (int)(idx - b * mm), i = r / m, j = r - i * m;
    out[idx] = (j <= i) ? x[b * n + fill_tri_src(i, j, m, n)] : 0.0;
}
// inverse (also the adjoint of the forward map, it is a permutation of the lower triangle): x[src(i, j)] = mat[i][j]
__global__ void fill_tri_inv_kernel(const double* mat, int64_t batch, int m, double* x) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t mm = (int64_t)m * m, n = (int64_t)m * (m + 1) / 2;
    if (idx >= batch * mm) return;
    const int64_t b = idx / mm;
    const int r = (int)(idx - b * mm), i = r / m, j = r - i * m;
    if (j <= i) x[b * n + fill_tri_src(i, j, m, n)] = mat[idx];
}

// ---- f4: k-means ----------------------------------------------------------------------------------------------
// assignment: nearest centroid (squared Euclidean, lowest index wins ties); centroids staged in shared memory
__global__ void __launch_bounds__(256) kmeans_assign_kernel(const double* X, int64_t N, int D, const double* C, int M,
                                                            int32_t* label, double* dist_part) {
    extern __shared__ double Cs[];   // [M][D]
    __shared__ double red[32];
    for (int i = threadIdx.x; i < M * D; i += blockDim.x) Cs[i] = C[i];
    __syncthreads();
    double local = 0.0;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        double best = DBL_MAX;
        int arg = 0;
        for (int m = 0; m < M; ++m) {
            double s = 0.0;
            for (int d = 0; d < D; ++d) {
                const double diff = X[n * D + d] - Cs[m * D + d];
                s = fma(diff, diff, s);
            }
            if (s < best) { best = s; arg = m; }
        }
        label[n] = arg;
        local += sqrt(best);   // scipy's distortion: mean Euclidean distance to the nearest code
    }
    local = block_sum(local, red);
    if (threadIdx.x == 0) dist_part[blockIdx.x] = local;
}
// update: one CTA per centroid walks all labels in a fixed order (deterministic; N M work, fine for an initialiser)
__global__ void __launch_bounds__(256) kmeans_update_kernel(const double* X, int64_t N, int D, const int32_t* label,
                                                            double* C, int32_t* count) {
    __shared__ double red[32];
    const int m = blockIdx.x;
    double cnt = 0.0;
    for (int64_t n = threadIdx.x; n < N; n += blockDim.x) cnt += (label[n] == m) ? 1.0 : 0.0;
    cnt = block_sum(cnt, red);
    __shared__ double cnt_s;
    if (threadIdx.x == 0) { cnt_s = cnt; count[m] = (int32_t)cnt; }
    __syncthreads();
    if (cnt_s == 0.0) return;   // empty cluster keeps its centroid (the host drops it at the end, as scipy does)
    for (int d = 0; d < D; ++d) {
        double s = 0.0;
        for (int64_t n = threadIdx.x; n < N; n += blockDim.x) s += (label[n] == m) ? X[n * D + d] : 0.0;
        s = block_sum(s, red);
        if (threadIdx.x == 0) C[m * D + d] = s / cnt_s;
        __syncthreads();
    }
}
__global__ void kmeans_fold_kernel(const double* part, int nparts, int64_t N, double* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nparts; ++i) s += part[i];
        out[0] = s / (double)N;
    }
}

}  // namespace mgp

using namespace mgp;

extern "C" {

int mgp_adam_step(void* cuda_stream, const mgp_adam_slot* slots, int32_t nslots, double grad_scale, double lr, double beta1,
                  double beta2, double eps, int64_t step, const double* guard) {
    if (!slots || nslots < 0 || nslots > MGP_ADAM_MAX_SLOTS || step < 1) return MGP_ERR_BAD_ARG;
    if (nslots == 0) return MGP_OK;
    AdamTable tab;
    int64_t nmax = 0;
    for (int i = 0; i < nslots; ++i) {
        if (!slots[i].theta || !slots[i].grad || !slots[i].m || !slots[i].v || slots[i].n < 0) return MGP_ERR_BAD_ARG;
        tab.s[i] = slots[i];
        if (slots[i].n > nmax) nmax = slots[i].n;
    }
    const double lr_t = lr * sqrt(1.0 - pow(beta2, (double)step)) / (1.0 - pow(beta1, (double)step));
    int gx = (int)((nmax + 255) / 256);
    if (gx > 1184) gx = 1184;   // 8 waves of 148 SMs; grid-stride beyond
    if (gx < 1) gx = 1;
    adam_kernel<<<dim3(gx, nslots), 256, 0, (cudaStream_t)cuda_stream>>>(tab, grad_scale, lr_t, beta1, beta2, eps, guard);
    return cudaGetLastError() == cudaSuccess ? MGP_OK : MGP_ERR_CUDA;
}

int mgp_gather_rows(void* cuda_stream, const double* X, const double* Y, const int64_t* idx, int64_t B, int32_t D, double* Xb,
                    double* Yb) {
    if (B < 0 || D < 1 || (B > 0 && (!X || !Y || !idx || !Xb || !Yb))) return MGP_ERR_BAD_ARG;
    if (B == 0) return MGP_OK;
    const int64_t total = B * (D + 1);
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(X, Y, idx, B, D, Xb, Yb);
    return cudaGetLastError() == cudaSuccess ? MGP_OK : MGP_ERR_CUDA;
}

int mgp_fill_triangular(void* cuda_stream, const double* src, int64_t batch, int32_t m, double* dst, int32_t inverse) {
    if (batch < 0 || m < 1 || (batch > 0 && (!src || !dst))) return MGP_ERR_BAD_ARG;
    if (batch == 0) return MGP_OK;
    const int64_t total = batch * m * m;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (inverse) fill_tri_inv_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(src, batch, m, dst);
    else fill_tri_fwd_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(src, batch, m, dst);
    return cudaGetLastError() == cudaSuccess ? MGP_OK : MGP_ERR_CUDA;
}

int mgp_kmeans_iterate(void* cuda_stream, const double* X, int64_t N, int32_t D, double* centroids, int32_t M, int32_t iters,
                       int32_t* label, int32_t* count, double* scratch, double* distortion) {
    if (N < 1 || D < 1 || M < 1 || iters < 0 || !X || !centroids || !label || !count || !scratch || !distortion) return MGP_ERR_BAD_ARG;
    if ((size_t)M * D * sizeof(double) > 200 * 1024) return MGP_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int nblk = 592;   // scratch holds >= 592 doubles
    const size_t smem = (size_t)M * D * sizeof(double);
    cudaFuncSetAttribute(kmeans_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int it = 0; it < iters; ++it) {
        kmeans_assign_kernel<<<nblk, 256, smem, st>>>(X, N, D, centroids, M, label, scratch);
        kmeans_update_kernel<<<M, 256, 0, st>>>(X, N, D, label, centroids, count);
    }
    kmeans_assign_kernel<<<nblk, 256, smem, st>>>(X, N, D, centroids, M, label, scratch);   // labels / distortion of the result
    kmeans_fold_kernel<<<1, 32, 0, st>>>(scratch, nblk, N, distortion);
    return cudaGetLastError() == cudaSuccess ? MGP_OK : MGP_ERR_CUDA;
}

}  // extern "C"
