/* libmgp — C-ABI of the B200-native SVGP-mixture (data-association GP) hot path.
 *
 * The reference (LouieMiddle/ModulatedGPs) has no FFI: its boundary is the Python class API of the
 * MixtureGPs package.  Each entry point below replaces the arithmetic behind one of those methods
 * (reference file:line cited per function); modulatedgps_b200/ binds them with ctypes and re-exposes the
 * reference's class names.  See INTEGRATION.md for the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to float64 (or int64 where stated), row-major, contiguous,
 *     caller-owned; the library never frees or retains caller memory beyond the call;
 *   - all work is enqueued on the stream given to mgp_ctx_create; calls are asynchronous with respect to
 *     the host unless stated; one ctx per (process, device); a ctx is not thread-safe;
 *   - every function returns 0 on success, non-zero (MGP_ERR_*) on failure; mgp_last_error(ctx) gives text;
 *     no exception or abort crosses this boundary.  There is NO CPU fallback: without a CUDA device
 *     mgp_ctx_create fails.
 */
#ifndef MGP_H_
#define MGP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGP_OK 0
#define MGP_ERR_BAD_ARG 1
#define MGP_ERR_CUDA 2
#define MGP_ERR_NOT_PD 3   /* Cholesky of Kuu hit a non-positive pivot (TF: InvalidArgumentError) */
#define MGP_ERR_NOMEM 4
#define MGP_ERR_NCCL 6     /* a communicator is attached but NCCL failed or cannot be resolved */
#define MGP_ERR_STALE_PRECOMPUTE 5   /* parameter values changed in place between mgp_elbo_local and mgp_elbo_finish */

/* gpflow RobustMax.prob_is_largest squashes every Gaussian CDF into (s, 1 - s) before the product over classes:
 * cdfs = cdfs * (1 - 2 s) + s with the literal s = 1e-4 (gpflow/likelihoods/multiclass.py, 1.x through 2.x; the
 * attribute RobustMax._squash = 1e-6 is set in __init__ and never read).  oracle/svgp_mixture.py and oracle/shim
 * carry the same named constant; tests/test_host_logic.py checks the three agree.  DESIGN.md section 3 records the
 * evidence for 1e-4 over 1e-6 (both were replayed against the reference's published config-#2 trajectory). */
#define MGP_ROBUSTMAX_CDF_SQUASH 1e-4

#define MGP_MAX_K 8        /* latent GPs (components) per layer */
#define MGP_MAX_D 32       /* input dimensions */

#define MGP_MODEL_SMGP 0          /* MixtureGPs/models.py:44-103  */
#define MGP_MODEL_SMGP_MODIFIED 1 /* MixtureGPs/models.py:106-123 */
#define MGP_LIK_GAUSSIAN 0        /* MixtureGPs/likelihoods.py:12-41 (GaussianModified) */
#define MGP_LIK_MULTICLASS 1      /* gpflow MultiClass(RobustMax), via broadcasting_lik.py:26-37 */

typedef struct mgp_ctx mgp_ctx;

/* One whitened SVGP layer (SVGPModified, MixtureGPs/models.py:147-160), CONSTRAINED parameter values. */
typedef struct {
    int32_t M, D, K;             /* inducing points, input dims, latent GPs */
    int32_t n_lengthscales;      /* 1 (isotropic, every demo) or D (ARD) */
    const double* Z;             /* [M, D]   inducing_variable.Z */
    const double* q_mu;          /* [M, K] */
    const double* q_sqrt;        /* [K, M, M] dense; only the lower triangle is read (band_part(-1,0)) */
    const double* variance;      /* [1]  kernel.variance */
    const double* lengthscales;  /* [n_lengthscales] */
} mgp_layer;

/* Gradient of the ELBO w.r.t. the constrained values of one layer; same shapes as mgp_layer.
 * q_sqrt gradient is dense [K, M, M] with zeros above the diagonal. */
typedef struct {
    double* Z;
    double* q_mu;
    double* q_sqrt;
    double* variance;
    double* lengthscales;
} mgp_layer_grad;

/* Noise for the Monte-Carlo pass.  Parity mode: explicit arrays (what tf.random.normal, models.py:57, and
 * TFP's uniform draw, models.py:73, would have returned).  Throughput mode: z = u = NULL and a Philox
 * counter-based stream keyed by (seed, global point index, sample, component), independent of sharding. */
typedef struct {
    const double* z;        /* [S, N_local, K] standard normal, or NULL */
    const double* u;        /* [S, N_local, K] uniform on (DBL_MIN, 1), or NULL */
    uint64_t seed;          /* Philox key when z/u are NULL */
    int64_t point_offset;   /* global index of this shard's first point (Philox counter offset) */
} mgp_noise;

typedef struct {
    int32_t model;            /* MGP_MODEL_* */
    int32_t lik;              /* MGP_LIK_*: the EXPERT likelihood (pred layer) */
    int32_t S;                /* num_samples */
    int32_t reserved;
    double temperature;       /* 1e-2 in the reference (models.py:60) */
    double num_data;          /* SGP.num_data: divisor of the KL term (models.py:79) */
    int64_t n_global;         /* global minibatch size: divisor of the per-point mean (models.py:76) */
} mgp_elbo_cfg;

int mgp_ctx_create(int device, void* cuda_stream, mgp_ctx** out);
void mgp_ctx_destroy(mgp_ctx* ctx);
const char* mgp_last_error(const mgp_ctx* ctx);
/* number of CUDA kernels this ctx has launched since creation (bench.py's gpu_launches) */
int64_t mgp_launch_count(const mgp_ctx* ctx);
/* Synchronises the stream and reports deferred device-side failures (MGP_ERR_NOT_PD when a Cholesky pivot
 * was not positive since the last check; MGP_ERR_CUDA for asynchronous CUDA errors). */
int mgp_check_status(mgp_ctx* ctx);
/* Optional per-stage device timing (CUDA events on the launch stream around each kernel group).  Off by default;
 * bench.py switches it on for the timed region to report the live roofline of the dominant kernel.
 * mgp_timing_read synchronises the stream; ms/calls have mgp_num_stages() entries, named by mgp_stage_name. */
int mgp_timing_enable(mgp_ctx* ctx, int on);
int mgp_timing_read(mgp_ctx* ctx, double* ms, int64_t* calls, int reset);
int mgp_num_stages(void);
const char* mgp_stage_name(int i);
/* Data parallelism without torch (mgp_ctx_create(device, nccl_comm_or_null, stream)): hand the
 * context an initialised ncclComm_t (caller-owned; one rank per process; NULL detaches).  While one is attached,
 * mgp_elbo_local leaves `reduce_buf` already summed over the ranks (ncclAllReduce, double, sum, on the context's stream:
 * the header and the pred layer's part first, the assign layer's part behind it), so that
 * mgp_elbo_local -> mgp_elbo_finish, and therefore mgp_elbo_fwd_bwd, are complete data-parallel steps: X / Y are this
 * rank's shard, cfg.n_global the global batch size, noise->point_offset the shard's first global row.  NCCL is not
 * linked: ncclAllReduce is looked up in the libnccl.so.2 already loaded in the process that created the communicator.
 * mgp_all_reduce is the same collective on a caller buffer (e.g. for a caller that reduces its own statistics). */
int mgp_ctx_set_comm(mgp_ctx* ctx, void* nccl_comm);
int mgp_all_reduce(mgp_ctx* ctx, double* buf, int64_t n);
/* Override MGP_ROBUSTMAX_CDF_SQUASH for this context (0 <= s < 0.5) — for the A/B replay of config #2 only. */
int mgp_set_robustmax_squash(mgp_ctx* ctx, double squash);
/* cap for the per-call point chunk (0 = automatic: whole shard if the materialised A fits the budget) */
int mgp_set_chunk_points(mgp_ctx* ctx, int64_t max_points);

/* SVGPModified.predict_f(Xnew, full_cov=False)  — MixtureGPs/models.py:129-144 (one of the S identical
 * slices the reference returns).  X [N, D] -> fmean [N, K], fvar [N, K]. */
int mgp_svgp_predict_f(mgp_ctx* ctx, const mgp_layer* layer, const double* X, int64_t N,
                       double* fmean, double* fvar);

/* SVGP.prior_kl() with whiten=True — gpflow gauss_kl(q_mu, q_sqrt, None), call site models.py:79.  kl [1]. */
int mgp_prior_kl(mgp_ctx* ctx, const mgp_layer* layer, double* kl);

/* SGP.predict_y — models.py:38-41 + likelihoods.py:31-32 (Gaussian: Fvar + variance_k) or gpflow
 * MultiClass._predict_mean_and_var.  lik_var [K] (ignored for MULTICLASS). */
int mgp_predict_y(mgp_ctx* ctx, const mgp_layer* pred, int32_t lik, const double* lik_var,
                  const double* X, int64_t N, double* mean, double* var);

/* SMGP.predict_assign — models.py:85-89: probs [N, K] = softmax_K(mu_assign); argmax int64 [N]
 * (lowest index wins ties, as np.argmax / tf.argmax). */
int mgp_predict_assign(mgp_ctx* ctx, const mgp_layer* assign, const double* X, int64_t N,
                       double* probs, int64_t* argmax);

/* SMGP.predict_samples — models.py:91-103.  noise->z is the z of W_dist (models.py:57), noise->u the relaxed
 * one-hot uniform, z_pred [S, N, K] the second normal draw (models.py:98).  Outputs [S, N] each. */
int mgp_predict_samples(mgp_ctx* ctx, const mgp_layer* pred, const mgp_layer* assign, int32_t lik,
                        const double* lik_var, const double* X, int64_t N, int32_t S, double temperature,
                        const mgp_noise* noise, const double* z_pred, double* samples_y, double* samples_f);

/* SMGP.W_dist(X).sample(1)[0] reshaped [S, N, K] — models.py:55-61,73-74: relaxed one-hot (Gumbel-softmax at
 * `temperature`) sample of the assign layer's reparameterised logits.  noise as in mgp_elbo_local. */
int mgp_w_sample(mgp_ctx* ctx, const mgp_layer* assign, const double* X, int64_t N, int32_t S, double temperature,
                 const mgp_noise* noise, double* W);

/* SMGP.E_log_p_Y(X, Y, W_SND) — models.py:63-67: out [N] = logsumexp_S(sum_k W ve) - log S with the expert
 * likelihood's variational expectation ve (likelihoods.py:39-41 / gpflow MultiClass).  W [S, N, K]. */
int mgp_e_log_p_y(mgp_ctx* ctx, const mgp_layer* pred, int32_t lik, const double* lik_var, const double* X,
                  const double* Y, int64_t N, int32_t S, const double* W, double* out);

/* The likelihood methods the reference exposes as stand-alone calls; Fmu / Fvar are [S, N, K] (row r = s N + n reads
 * Y[n]: the reference tiles Y over the sample axis, MixtureGPs/broadcasting_lik.py:22-37), Y [N] float64 (class index
 * for MULTICLASS).
 *   mgp_lik_variational_expectations  BroadcastingLikelihood.variational_expectations (broadcasting_lik.py:39-42):
 *        GAUSSIAN -> out [S, N, K], GaussianModified._variational_expectations, NOT reduced over k (likelihoods.py:39-41);
 *        MULTICLASS -> out [S, N] (the reference's [S, N, 1]), gpflow MultiClass._variational_expectations.
 *   mgp_lik_predict_mean_and_var      BroadcastingLikelihood.predict_mean_and_var (broadcasting_lik.py:44-46):
 *        GaussianModified._predict_mean_and_var (likelihoods.py:31-32) / MultiClass._predict_mean_and_var; rows = S N.
 *   mgp_lik_log_prob                  GaussianModified._scalar_log_prob (likelihoods.py:21-22) -> [S, N, K].
 *   mgp_lik_predict_log_density       GaussianModified._predict_log_density (likelihoods.py:34-35) -> [S, N]. */
int mgp_lik_variational_expectations(mgp_ctx* ctx, int32_t lik, const double* lik_var, const double* Fmu,
                                     const double* Fvar, const double* Y, int64_t S, int64_t N, int32_t K, double* out);
int mgp_lik_predict_mean_and_var(mgp_ctx* ctx, int32_t lik, const double* lik_var, const double* Fmu,
                                 const double* Fvar, int64_t rows, int32_t K, double* mean, double* var);
int mgp_lik_log_prob(mgp_ctx* ctx, const double* lik_var, const double* F, const double* Y, int64_t S, int64_t N,
                     int32_t K, double* out);
int mgp_lik_predict_log_density(mgp_ctx* ctx, const double* lik_var, const double* Fmu, const double* Fvar,
                                const double* Y, int64_t S, int64_t N, int32_t K, double* out);

/* Size (in doubles) of the flat reduction buffer of mgp_elbo_local for these layers. */
int64_t mgp_reduce_buffer_len(const mgp_layer* pred, const mgp_layer* assign);

/* ELBO forward + backward, phase 1 (per shard): SMGP._build_likelihood / SMGPModified.E_log_p_Y —
 * models.py:55-79,112-123 — and TF's reverse pass through it (utils/training_utils.py:8-10), up to the
 * point where every per-shard quantity is a SUM over this shard's points.  Writes those sums to
 * `reduce_buf` (mgp_reduce_buffer_len doubles): this is the single buffer a data-parallel caller
 * all-reduces (sum) across ranks.  X [N_local, D], Y [N_local]. */
int mgp_elbo_local(mgp_ctx* ctx, const mgp_elbo_cfg* cfg, const mgp_layer* pred, const mgp_layer* assign,
                   const double* lik_var, const double* assign_lik_var,
                   const double* X, const double* Y, int64_t N_local, const mgp_noise* noise,
                   double* reduce_buf);

/* Phase 2 (replicated): from the (all-reduced) buffer, the KL terms (models.py:79), the Cholesky and
 * kernel backward, and the final gradients w.r.t. constrained values.  elbo [1]; lik_var_grad [K] and
 * assign_lik_var_grad [K] may be NULL when the corresponding likelihood has no variance.
 * The factorisations mgp_elbo_local formed are reused only when `pred` / `assign` are the same layers (dimensions
 * and parameter pointers); otherwise they are formed again here, so two models may be interleaved on one context.
 * Parameter VALUES must not change between the two calls: a fingerprint of them is compared on the device and a
 * mismatch sets elbo = NaN and makes the next mgp_check_status return MGP_ERR_STALE_PRECOMPUTE. */
int mgp_elbo_finish(mgp_ctx* ctx, const mgp_elbo_cfg* cfg, const mgp_layer* pred, const mgp_layer* assign,
                    const double* lik_var, const double* assign_lik_var, const double* reduce_buf,
                    double* elbo, mgp_layer_grad* pred_grad, mgp_layer_grad* assign_grad,
                    double* lik_var_grad, double* assign_lik_var_grad);

/* Convenience: phase 1 + phase 2 on one device (no exchange). */
int mgp_elbo_fwd_bwd(mgp_ctx* ctx, const mgp_elbo_cfg* cfg, const mgp_layer* pred, const mgp_layer* assign,
                     const double* lik_var, const double* assign_lik_var,
                     const double* X, const double* Y, int64_t N_local, const mgp_noise* noise,
                     double* elbo, mgp_layer_grad* pred_grad, mgp_layer_grad* assign_grad,
                     double* lik_var_grad, double* assign_lik_var_grad);

/* ---- the callers either side of the ELBO step, stream-ordered, no ctx needed ---------- */
#define MGP_TRANSFORM_IDENTITY 0
#define MGP_TRANSFORM_SOFTPLUS 1      /* gpflow positive(): d constrained / d theta = sigmoid(theta) */
#define MGP_ADAM_MAX_SLOTS 16
typedef struct {
    double* theta;          /* [n] unconstrained variable, updated in place */
    const double* grad;     /* d ELBO / d CONSTRAINED value as mgp_elbo_finish wrote it */
    double* m;              /* [n] first moment */
    double* v;              /* [n] second moment */
    const int64_t* gather;  /* NULL, or [n] positions in `grad` (fill-triangular: vector entry i <- tril entry gather[i]) */
    int64_t n;
    int32_t transform;      /* MGP_TRANSFORM_* */
    int32_t reserved;
} mgp_adam_slot;

/* One fused Adam update of every trainable variable — replaces tf.optimizers.Adam(lr).minimize(training_loss,
 * model.trainable_variables) of utils/training_utils.py:6-10, TF 2.10 Keras defaults beta1 .9, beta2 .999, eps 1e-7.
 * grad_scale = -1 turns ELBO gradients into gradients of the training loss (models.py:81-83).  step >= 1.
 * guard: NULL, or a device scalar (the step's ELBO): when it is not finite the whole update is skipped on the device,
 * without a host synchronisation — a failed Cholesky (NaN ELBO and gradients) never reaches theta, m or v. */
int mgp_adam_step(void* cuda_stream, const mgp_adam_slot* slots, int32_t nslots, double grad_scale, double lr,
                  double beta1, double beta2, double eps, int64_t step, const double* guard);

/* Minibatch gather: Xb[r] = X[idx[r]], Yb[r] = Y[idx[r]] — the device half of
 * tf.data.Dataset.from_tensor_slices((X, Y)).shuffle(N).batch(B).repeat() (demos/demo_tf2.py:53-56). */
int mgp_gather_rows(void* cuda_stream, const double* X, const double* Y, const int64_t* idx, int64_t B, int32_t D,
                    double* Xb, double* Yb);

/* gpflow.utilities.triangular() — the TFP FillTriangular bijector behind SVGP.q_sqrt (gpflow/models/svgp.py, pinned
 * 2.7.0): inverse == 0: src [batch, m(m+1)/2] -> dst [batch, m, m] (lower band, zeros above);
 * inverse != 0: src [batch, m, m] -> dst [batch, m(m+1)/2] (also the adjoint: the map is a permutation). */
int mgp_fill_triangular(void* cuda_stream, const double* src, int64_t batch, int32_t m, double* dst, int32_t inverse);

/* `iters` Lloyd iterations from the given centroids [M, D] (in place), then labels int32 [N], cluster sizes int32 [M]
 * and the mean Euclidean distance to the nearest centroid (scipy.cluster.vq.kmeans' distortion; call site
 * demos/demo_tf2.py:39).  scratch: >= 592 doubles.  Empty clusters keep their centroid and report count 0. */
int mgp_kmeans_iterate(void* cuda_stream, const double* X, int64_t N, int32_t D, double* centroids, int32_t M,
                       int32_t iters, int32_t* label, int32_t* count, double* scratch, double* distortion);

/* Stage-level entry points (used by the parity tests to localise a failure; same kernels as above).
 * L, Linv: [M, M] row-major lower-triangular outputs for one layer. */
int mgp_debug_kuu_chol(mgp_ctx* ctx, const mgp_layer* layer, double* Kuu, double* L, double* Linv);
/* The throughput-mode noise generator, exposed for known-answer and distribution tests.
 * mgp_debug_philox: ctr_key uint32 [n][6] = {counter[4], key[2]} -> out uint32 [n][4] = Philox4x32-10 (Random123 KAT).
 * mgp_debug_noise:  the z (standard normal) and u (uniform) draws [S, N, K] the fused MC pass would use for noise->seed /
 *                   noise->point_offset; stream 0 = W_dist draws (models.py:57,73), 1 = predict_samples' second z (:98). */
int mgp_debug_philox(mgp_ctx* ctx, const uint32_t* ctr_key, int32_t n, uint32_t* out);
int mgp_debug_noise(mgp_ctx* ctx, const mgp_noise* noise, int64_t N, int32_t S, int32_t K, int32_t stream, double* z,
                    double* u);

#ifdef __cplusplus
}
#endif
#endif /* MGP_H_ */
