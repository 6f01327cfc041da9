// Fused Monte-Carlo likelihood pass (forward + adjoints) and the small per-point prediction kernels.
//
// One thread per point.  Replaces, for S samples at once and without materialising any [S,N,K] temporary:
//   SMGP.W_dist + reparameterize          MixtureGPs/models.py:55-61, MixtureGPs/utils.py:27
//   tfp RelaxedOneHotCategorical.sample   models.py:60,73  (gumbel = -log(-log u); exp(log_softmax((g+logits)/T)))
//   GaussianModified._variational_expectations   MixtureGPs/likelihoods.py:39-41 (per component, not summed)
//   gpflow MultiClass(RobustMax) expectation via BroadcastingLikelihood   broadcasting_lik.py:26-42
//   SMGP.E_log_p_Y / SMGPModified.E_log_p_Y      models.py:63-67 / 112-123  (logsumexp over samples)
// and TF's reverse pass through them.  The pass is HBM-bound: per point it reads 4K conditionals + y and (parity
// mode) 2 S K noise values, coalesced and vectorised by the [.., N, K] layouts, and writes 4K adjoints.
// One sweep over the samples: the softmax-over-samples weighted adjoints are accumulated online (running maximum +
// rescale), so no per-thread [S] storage and no recomputation of the noise / Gumbel-softmax is needed.
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "gh20.h"
#include "kernels.h"

namespace mgp {

constexpr double REPARAM_JITTER = 1e-6;  // config.default_jitter() in reparameterize, MixtureGPs/utils.py:27
constexpr int MC_THREADS = 128;
#ifndef MC_MIN_CTAS
#define MC_MIN_CTAS 4
#endif
constexpr double HALF_LOG_2PI = 0.91893853320467274178;

// ---- Philox4x32-10 (throughput mode) --------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// K standard normals and K uniforms on (0,1) for (global point, sample, stream): one Philox block per component
// (counter = point, sample, component | stream << 8) gives the component's 53-bit uniform; the normals come in
// Box-Muller PAIRS from the even component's block — radius and angle once, cosine branch for component k, sine branch
// for k + 1 (a log, a sqrt and a cospi per component would cost half as much FP64-pipe time again).
template <int K>
__device__ __forceinline__ void philox_draw_all(uint64_t seed, int64_t point, int s, int stream, double (&z)[K],
                                                double (&u)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        uint32_t c[4] = {(uint32_t)point, (uint32_t)((uint64_t)point >> 32), (uint32_t)s, (uint32_t)(k | (stream << 8))};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        const uint64_t r53 = (((uint64_t)c[2] << 32) | c[3]) >> 11;
        u[k] = ((double)r53 + 0.5) * (1.0 / 9007199254740992.0);
        if ((k & 1) == 0) {
            const double ua = ((double)c[0] + 0.5) * (1.0 / 4294967296.0);
            const double ub = ((double)c[1] + 0.5) * (1.0 / 4294967296.0);
            const double r = sqrt(-2.0 * log(ua));
            double sn, cs;
            sincospi(2.0 * ub, &sn, &cs);
            z[k] = r * cs;
            if (k + 1 < K) z[k + 1] = r * sn;
        }
    }
}

// ---- RobustMax Gauss-Hermite quadrature (gpflow MultiClass) ----------------------------
// c_gh_x / c_gh_w come statically initialised from gh20.h: a __constant__ symbol exists once per device, and a
// first-use symbol upload behind a process-wide flag would fill only the device current at that moment.

// epsilon after GPflow's Sigmoid-bijector round trip of 1e-3
__device__ __forceinline__ double robustmax_eps() {
    const double e = 1e-3;
    const double x = log(e) - log1p(-e);
    return 1.0 / (1.0 + exp(-x));
}

// p = P(f_c is largest) and its gradient w.r.t. mu[.] and var[.]
template <int K, bool GRAD>
__device__ __forceinline__ double robustmax_prob(int c, const double (&mu)[K], const double (&var)[K],
                                                 double (&dmu)[K], double (&dvar)[K], double squash) {
    // gpflow RobustMax.prob_is_largest: cdfs = cdfs * (1 - 2 squash) + squash, squash = MGP_ROBUSTMAX_CDF_SQUASH
    const double c1 = 1.0 - 2.0 * squash, INV_SQRT2 = 0.70710678118654752440, INV_SQRT_2PI = 0.39894228040143267794;
    double sd[K];
    double mu_c = 0.0, var_c = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        sd[j] = sqrt(fmax(var[j], 1e-10));
        if (j == c) { mu_c = mu[j]; var_c = var[j]; }
        if (GRAD) dmu[j] = dvar[j] = 0.0;
    }
    const double sdc2 = sqrt(fmax(2.0 * var_c, 1e-10));
    double p = 0.0, dmu_c = 0.0, dvar_c = 0.0;
    for (int q = 0; q < 20; ++q) {
        const double X = mu_c + c_gh_x[q] * sdc2;
        double cdf[K], dist[K];
        double prod = 1.0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            dist[j] = (X - mu[j]) / sd[j];
            cdf[j] = (0.5 * (1.0 + erf(dist[j] * INV_SQRT2))) * c1 + squash;
            if (j == c) cdf[j] = 1.0;
            prod *= cdf[j];
        }
        p += prod * c_gh_w[q];
        if (GRAD) {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                if (j == c) continue;
                // d prod / d dist_j
                const double dd = c_gh_w[q] * (prod / cdf[j]) * c1 * INV_SQRT_2PI * exp(-0.5 * dist[j] * dist[j]);
                dmu[j] -= dd / sd[j];
                if (var[j] > 1e-10) dvar[j] -= dd * (X - mu[j]) * 0.5 / (sd[j] * sd[j] * sd[j]);
                dmu_c += dd / sd[j];
                if (2.0 * var_c > 1e-10) dvar_c += dd / sd[j] * c_gh_x[q] / sdc2;
            }
        }
    }
    if (GRAD) {
#pragma unroll
        for (int j = 0; j < K; ++j)
            if (j == c) { dmu[j] = dmu_c; dvar[j] = dvar_c; }
    }
    return p;
}

// relaxed one-hot weights for one sample: W = exp(log_softmax((gumbel + logits)/T))
template <int K>
__device__ __forceinline__ void sample_weights(const double (&mu_a)[K], const double (&sd_a)[K], const double (&z)[K],
                                               const double (&u)[K], double inv_temperature, double (&W)[K]) {
    double x[K];
    double mx = -DBL_MAX;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double logit = mu_a[k] + z[k] * sd_a[k];          // reparameterize: mean + z * (var + jitter)**0.5
        const double gumbel = -log(-log(u[k]));
        x[k] = (gumbel + logit) * inv_temperature;   // (TFP divides: <= 1 ulp apart in x, ~1e-14 relative in W)
        mx = fmax(mx, x[k]);
    }
    // exp(log_softmax(x))_k = exp(x_k - mx) / sum_j exp(x_j - mx): the K exponentials are formed once and normalised
    // (TFP evaluates exp(x - mx - log(sum)); the two differ by a few ulp, far inside the 1e-9 parity bar)
    double ex[K];
    double se = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) { ex[k] = exp(x[k] - mx); se += ex[k]; }
    const double inv = 1.0 / se;
#pragma unroll
    for (int k = 0; k < K; ++k) W[k] = ex[k] * inv;
}

// softmax-over-samples weighted accumulators of one logsumexp term, online (flash-attention style)
template <int K>
struct OnlineTerm {
    double m = -DBL_MAX, s = 0.0;
    double mu[K], v[K], e[K];
    __device__ __forceinline__ OnlineTerm() {
#pragma unroll
        for (int k = 0; k < K; ++k) mu[k] = v[k] = e[k] = 0.0;
    }
    // one sample: weights W, per-component values c (t = sum_k W_k c_k), the normal draw z, hisd = 0.5 / sqrt(var + jitter)
    // (the divisions by the temperature and by sd are loop-invariant: one reciprocal each, outside the sample loop —
    // an FP64 division is ~25 instructions on the pipe this pass is bound by in throughput mode)
    __device__ __forceinline__ void add(const double (&W)[K], const double (&c)[K], const double (&z)[K],
                                        const double (&hisd)[K], double inv_temperature) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) t += W[k] * c[k];
        double w;
        if (t > m) {
            const double scale = exp(m - t);   // exp(-huge) = 0 on the first sample
            s = s * scale + 1.0;
#pragma unroll
            for (int k = 0; k < K; ++k) { mu[k] *= scale; v[k] *= scale; e[k] *= scale; }
            m = t;
            w = 1.0;
        } else {
            w = exp(t - m);
            s += w;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double xb = w * (c[k] * W[k] - W[k] * t) * inv_temperature;   // d/d logits_k (up to 1/(s n))
            mu[k] += xb;
            v[k] += xb * z[k] * hisd[k];
            e[k] += w * W[k];
        }
    }
};

template <int K>
__device__ __forceinline__ void load_noise(const McArgs& a, int64_t i, int s, double (&z)[K], double (&u)[K]) {
    if (a.z != nullptr) {
        const size_t base = ((size_t)s * a.n_local + a.chunk_offset + i) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            z[k] = __ldg(a.z + base + k);
            u[k] = __ldg(a.u + base + k);
        }
    } else {
        const int64_t point = a.point_offset + a.chunk_offset + i;
        philox_draw_all<K>(a.seed, point, s, 0, z, u);
    }
}

template <int K, int MODEL, int LIK>
__global__ void __launch_bounds__(MC_THREADS, MC_MIN_CTAS) mc_pass_kernel(McArgs a, double* block_part) {
    __shared__ double red[32];
    const int64_t i = (int64_t)blockIdx.x * MC_THREADS + threadIdx.x;
    const bool live = i < a.n;
    double data = 0.0, sumv_p = 0.0, sumv_a = 0.0;
    double glik[K], galik[K];
#pragma unroll
    for (int k = 0; k < K; ++k) glik[k] = galik[k] = 0.0;

    if (live) {
        double mu_p[K], var_p[K], mu_a[K], var_a[K], sd_a[K], e[K], eA[K];
        const double y = a.Y[i];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            mu_p[k] = a.fmean_p[(size_t)i * K + k];
            var_p[k] = a.fvar_p[(size_t)i * K + k];
            mu_a[k] = a.fmean_a[(size_t)i * K + k];
            var_a[k] = a.fvar_a[(size_t)i * K + k];
            sd_a[k] = sqrt(var_a[k] + REPARAM_JITTER);
        }
        double dp_dmu[K], dp_dvar[K];
        if (LIK == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double s = a.lik_var[k], r = y - mu_p[k];
                e[k] = -HALF_LOG_2PI - 0.5 * log(s) - 0.5 * (r * r + var_p[k]) / s;
            }
        } else {
            const double eps = robustmax_eps();
            const double p = robustmax_prob<K, true>((int)y, mu_p, var_p, dp_dmu, dp_dvar, a.squash);
            const double ve = p * log(1.0 - eps) + (1.0 - p) * log(eps / (K - 1.0));
#pragma unroll
            for (int k = 0; k < K; ++k) e[k] = ve;
        }
        if (MODEL == 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double s = a.assign_lik_var[k], r = y - mu_a[k];
                eA[k] = -HALF_LOG_2PI - 0.5 * log(s) - 0.5 * (r * r + var_a[k]) / s;
            }
        }
        // ONE sweep over the samples.  The adjoints are softmax_s(t_s)-weighted sums, so they are accumulated with a
        // running maximum and rescaled whenever it moves (online softmax), then normalised by the final sum:
        //   w_s = exp(t_s - lse) / n_global ;  d/dW_k = w_s c_k ;  through exp(log_softmax(x)):
        //   xbar_k = Wbar_k W_k - W_k sum_j Wbar_j W_j = w_s (c_k W_k - W_k t_s)          (c = e or eA, t = sum_k W_k c_k)
        OnlineTerm<K> ty_acc, tA_acc;
        const double inv_temperature = 1.0 / a.temperature;
        double hisd_a[K];
#pragma unroll
        for (int k = 0; k < K; ++k) hisd_a[k] = 0.5 / sd_a[k];
        for (int s = 0; s < a.S; ++s) {
            double z[K], u[K], W[K];
            load_noise<K>(a, i, s, z, u);
            sample_weights<K>(mu_a, sd_a, z, u, inv_temperature, W);
            ty_acc.add(W, e, z, hisd_a, inv_temperature);
            if (MODEL == 1) tA_acc.add(W, eA, z, hisd_a, inv_temperature);
        }
        const double logS = log((double)a.S);
        const double lse_y = log(ty_acc.s) + ty_acc.m;
        double lse_A = 0.0;
        if (MODEL == 1) lse_A = log(tA_acc.s) + tA_acc.m;
        // SMGP: logsumexp_s(t) - log S ; Modified: logsumexp_s(tA - log S) + logsumexp_s(ty - log S)
        data = (MODEL == 0) ? (lse_y - logS) : ((lse_A - logS) + (lse_y - logS));
        data *= a.inv_n_global;
        double mub_a[K], vb_a[K], ebar[K], eAbar[K];
        {
            const double ny = a.inv_n_global / ty_acc.s;
            const double nA = (MODEL == 1) ? a.inv_n_global / tA_acc.s : 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                mub_a[k] = ty_acc.mu[k] * ny;
                vb_a[k] = ty_acc.v[k] * ny;
                ebar[k] = ty_acc.e[k] * ny;
                eAbar[k] = 0.0;
                if (MODEL == 1) {
                    mub_a[k] += tA_acc.mu[k] * nA;
                    vb_a[k] += tA_acc.v[k] * nA;
                    eAbar[k] = tA_acc.e[k] * nA;
                }
            }
        }
        // through the likelihood terms
        double mub_p[K], vb_p[K];
        if (LIK == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double s = a.lik_var[k], r = y - mu_p[k];
                mub_p[k] = ebar[k] * r / s;
                vb_p[k] = -0.5 * ebar[k] / s;
                glik[k] = ebar[k] * (-0.5 / s + 0.5 * (r * r + var_p[k]) / (s * s));
            }
        } else {
            const double eps = robustmax_eps();
            double esum = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) esum += ebar[k];
            const double pbar = esum * (log(1.0 - eps) - log(eps / (K - 1.0)));
#pragma unroll
            for (int k = 0; k < K; ++k) {
                mub_p[k] = pbar * dp_dmu[k];
                vb_p[k] = pbar * dp_dvar[k];
            }
        }
        if (MODEL == 1) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double s = a.assign_lik_var[k], r = y - mu_a[k];
                mub_a[k] += eAbar[k] * r / s;
                vb_a[k] += -0.5 * eAbar[k] / s;
                galik[k] = eAbar[k] * (-0.5 / s + 0.5 * (r * r + var_a[k]) / (s * s));
            }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            a.mubar_p[(size_t)i * K + k] = mub_p[k];
            a.vbar_p[(size_t)i * K + k] = vb_p[k];
            a.mubar_a[(size_t)i * K + k] = mub_a[k];
            a.vbar_a[(size_t)i * K + k] = vb_a[k];
            sumv_p += vb_p[k];
            sumv_a += vb_a[k];
        }
    } else if (i < a.ldn) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            a.mubar_p[(size_t)i * K + k] = 0.0;
            a.vbar_p[(size_t)i * K + k] = 0.0;
            a.mubar_a[(size_t)i * K + k] = 0.0;
            a.vbar_a[(size_t)i * K + k] = 0.0;
        }
    }
    // deterministic block partials
    double* out = block_part + (size_t)blockIdx.x * MC_NPART;
    double r = block_sum(data, red);
    if (threadIdx.x == 0) out[0] = r;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        r = block_sum(k < K ? glik[k < K ? k : 0] : 0.0, red);
        if (threadIdx.x == 0) out[1 + k] = r;
        r = block_sum(k < K ? galik[k < K ? k : 0] : 0.0, red);
        if (threadIdx.x == 0) out[9 + k] = r;
    }
    r = block_sum(sumv_p, red);
    if (threadIdx.x == 0) out[17] = r;
    r = block_sum(sumv_a, red);
    if (threadIdx.x == 0) out[18] = r;
    if (threadIdx.x == 0) out[19] = 0.0;
}

int mc_num_blocks(int64_t ldn) { return (int)((ldn + MC_THREADS - 1) / MC_THREADS); }

template <int K>
static void mc_dispatch(const McArgs& a, double* block_part, int nblocks, cudaStream_t st) {
    if (a.model == 0 && a.lik == 0) mc_pass_kernel<K, 0, 0><<<nblocks, MC_THREADS, 0, st>>>(a, block_part);
    else if (a.model == 0 && a.lik == 1) mc_pass_kernel<K, 0, 1><<<nblocks, MC_THREADS, 0, st>>>(a, block_part);
    else if (a.model == 1 && a.lik == 0) mc_pass_kernel<K, 1, 0><<<nblocks, MC_THREADS, 0, st>>>(a, block_part);
    else mc_pass_kernel<K, 1, 1><<<nblocks, MC_THREADS, 0, st>>>(a, block_part);
}

void mc_pass(const McArgs& a, double* block_part, const Launch& ln) {
    const int nblocks = mc_num_blocks(a.ldn);
    switch (a.K) {
        case 1: mc_dispatch<1>(a, block_part, nblocks, ln.stream); break;
        case 2: mc_dispatch<2>(a, block_part, nblocks, ln.stream); break;
        case 3: mc_dispatch<3>(a, block_part, nblocks, ln.stream); break;
        case 4: mc_dispatch<4>(a, block_part, nblocks, ln.stream); break;
        case 5: mc_dispatch<5>(a, block_part, nblocks, ln.stream); break;
        case 6: mc_dispatch<6>(a, block_part, nblocks, ln.stream); break;
        case 7: mc_dispatch<7>(a, block_part, nblocks, ln.stream); break;
        default: mc_dispatch<8>(a, block_part, nblocks, ln.stream); break;
    }
    ln.tick();
}

// header[q] += sum over blocks, one CTA per entry q (fixed order => deterministic)
__global__ void __launch_bounds__(256) mc_fold_kernel(const double* block_part, int nblocks, double* hdr) {
    __shared__ double red[32];
    const int q = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += block_part[(size_t)b * MC_NPART + q];
    s = block_sum(s, red);
    if (threadIdx.x == 0) hdr[q] += s;
}

void mc_fold(const double* block_part, int nblocks, double* rb_header, const Launch& ln) {
    mc_fold_kernel<<<MC_NPART - 1, 256, 0, ln.stream>>>(block_part, nblocks, rb_header);
    ln.tick();
}

// ==================================================================================================
// prediction kernels
// ==================================================================================================
// SMGP.predict_assign: softmax_K(mean_S(mu)) — the S tiled copies are identical, so the mean is mu itself
__global__ void predict_assign_k(const double* fmean, int64_t n, int K, double* probs, int64_t* argmax) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mx = -DBL_MAX;
    for (int k = 0; k < K; ++k) mx = fmax(mx, fmean[(size_t)i * K + k]);
    double se = 0.0;
    for (int k = 0; k < K; ++k) se += exp(fmean[(size_t)i * K + k] - mx);
    double best = -1.0;
    int64_t arg = 0;
    for (int k = 0; k < K; ++k) {
        const double p = exp(fmean[(size_t)i * K + k] - mx) / se;
        probs[(size_t)i * K + k] = p;
        if (p > best) { best = p; arg = k; }   // first maximum wins, as np.argmax / tf.argmax
    }
    argmax[i] = arg;
}

void predict_assign_kernel(const double* fmean, int64_t n, int K, double* probs, int64_t* argmax, const Launch& ln) {
    if (n <= 0) return;
    predict_assign_k<<<(unsigned)((n + 127) / 128), 128, 0, ln.stream>>>(fmean, n, K, probs, argmax);
    ln.tick();
}

template <int K>
__global__ void predict_y_k(const double* fmean, const double* fvar, int64_t n, int lik, const double* lik_var,
                            double squash, double* mean, double* var) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mu[K], v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { mu[k] = fmean[(size_t)i * K + k]; v[k] = fvar[(size_t)i * K + k]; }
    if (lik == 0) {   // GaussianModified._predict_mean_and_var: (Fmu, Fvar + variance)
#pragma unroll
        for (int k = 0; k < K; ++k) { mean[(size_t)i * K + k] = mu[k]; var[(size_t)i * K + k] = v[k] + lik_var[k]; }
    } else {          // MultiClass._predict_mean_and_var
        const double eps = robustmax_eps();
        double d0[K], d1[K];
        for (int c = 0; c < K; ++c) {
            const double p = robustmax_prob<K, false>(c, mu, v, d0, d1, squash);
            const double ps = p * (1.0 - eps) + (1.0 - p) * (eps / (K - 1.0));
            mean[(size_t)i * K + c] = ps;
            var[(size_t)i * K + c] = ps - ps * ps;
        }
    }
}

void predict_y_kernel(const double* fmean, const double* fvar, int64_t n, int K, int lik, const double* lik_var,
                      double squash, double* mean, double* var, const Launch& ln) {
    if (n <= 0) return;
    const unsigned grid = (unsigned)((n + 127) / 128);
#define PY(KK) case KK: predict_y_k<KK><<<grid, 128, 0, ln.stream>>>(fmean, fvar, n, lik, lik_var, squash, mean, var); break;
    switch (K) { PY(1) PY(2) PY(3) PY(4) PY(5) PY(6) PY(7) default: predict_y_k<8><<<grid, 128, 0, ln.stream>>>(fmean, fvar, n, lik, lik_var, squash, mean, var); }
#undef PY
    ln.tick();
}

// SMGP.predict_samples (models.py:91-103): one thread per (sample, point)
template <int K>
__global__ void predict_samples_k(SampleArgs a) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)a.S * a.n) return;
    const int s = (int)(idx / a.n);
    const int64_t i = idx % a.n;
    double mu_a[K], sd_a[K], z[K], u[K], zp[K], W[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        mu_a[k] = a.fmean_a[(size_t)i * K + k];
        sd_a[k] = sqrt(a.fvar_a[(size_t)i * K + k] + REPARAM_JITTER);
        if (a.z != nullptr) {
            z[k] = a.z[((size_t)s * a.n + i) * K + k];
            u[k] = a.u[((size_t)s * a.n + i) * K + k];
            zp[k] = a.z_pred[((size_t)s * a.n + i) * K + k];
        }
    }
    if (a.z == nullptr) {
        double dummy[K];
        philox_draw_all<K>(a.seed, a.point_offset + i, s, 0, z, u);
        philox_draw_all<K>(a.seed, a.point_offset + i, s, 1, zp, dummy);
    }
    sample_weights<K>(mu_a, sd_a, z, u, 1.0 / a.temperature, W);
    double mu[K], v[K], my[K], vy[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { mu[k] = a.fmean_p[(size_t)i * K + k]; v[k] = a.fvar_p[(size_t)i * K + k]; }
    if (a.lik == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) { my[k] = mu[k]; vy[k] = v[k] + a.lik_var[k]; }
    } else {
        const double eps = robustmax_eps();
        const double squash = a.squash;
        double d0[K], d1[K];
        for (int c = 0; c < K; ++c) {
            const double p = robustmax_prob<K, false>(c, mu, v, d0, d1, squash);
            my[c] = p * (1.0 - eps) + (1.0 - p) * (eps / (K - 1.0));
            vy[c] = my[c] - my[c] * my[c];
        }
    }
    double sy = 0.0, sf = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        sy += (my[k] + zp[k] * sqrt(vy[k] + REPARAM_JITTER)) * W[k];   // reparameterize(mean, var, z) * W
        sf += (mu[k] + zp[k] * sqrt(v[k] + REPARAM_JITTER)) * W[k];
    }
    a.samples_y[idx] = sy;
    a.samples_f[idx] = sf;
}

void predict_samples_kernel(const SampleArgs& a, const Launch& ln) {
    const int64_t total = (int64_t)a.S * a.n;
    if (total <= 0) return;
    const unsigned grid = (unsigned)((total + 127) / 128);
#define PS(KK) case KK: predict_samples_k<KK><<<grid, 128, 0, ln.stream>>>(a); break;
    switch (a.K) { PS(1) PS(2) PS(3) PS(4) PS(5) PS(6) PS(7) default: predict_samples_k<8><<<grid, 128, 0, ln.stream>>>(a); }
#undef PS
    ln.tick();
}

// ==================================================================================================
// SMGP.W_dist(...).sample and SMGP.E_log_p_Y as stand-alone calls (models.py:55-67): the ELBO path fuses them into
// mc_pass; these exist so that code written against the reference's intermediate methods keeps working.
// ==================================================================================================
// W [S, n, K] = relaxed one-hot sample at temperature T of logits mu_a + z sqrt(var_a + jitter)
template <int K>
__global__ void w_sample_k(SampleArgs a, double* W_out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)a.S * a.n) return;
    const int s = (int)(idx / a.n);
    const int64_t i = idx % a.n;
    double mu_a[K], sd_a[K], z[K], u[K], W[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        mu_a[k] = a.fmean_a[(size_t)i * K + k];
        sd_a[k] = sqrt(a.fvar_a[(size_t)i * K + k] + REPARAM_JITTER);
        if (a.z != nullptr) {
            z[k] = a.z[((size_t)s * a.n + i) * K + k];
            u[k] = a.u[((size_t)s * a.n + i) * K + k];
        }
    }
    if (a.z == nullptr) philox_draw_all<K>(a.seed, a.point_offset + i, s, 0, z, u);
    sample_weights<K>(mu_a, sd_a, z, u, 1.0 / a.temperature, W);
#pragma unroll
    for (int k = 0; k < K; ++k) W_out[idx * K + k] = W[k];
}

void w_sample_kernel(const SampleArgs& a, double* W_out, const Launch& ln) {
    const int64_t total = (int64_t)a.S * a.n;
    if (total <= 0) return;
    const unsigned grid = (unsigned)((total + 127) / 128);
#define WS(KK) case KK: w_sample_k<KK><<<grid, 128, 0, ln.stream>>>(a, W_out); break;
    switch (a.K) { WS(1) WS(2) WS(3) WS(4) WS(5) WS(6) WS(7) default: w_sample_k<8><<<grid, 128, 0, ln.stream>>>(a, W_out); }
#undef WS
    ln.tick();
}

// out[n] = logsumexp_s( sum_k W[s,n,k] ve[n,k] ) - log S,  ve = the expert likelihood's variational expectation
template <int K>
__global__ void e_log_p_y_k(const double* fmean, const double* fvar, const double* Y, const double* lik_var, int lik,
                            double squash, const double* W, int S, int64_t n, double* out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mu[K], v[K], ve[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { mu[k] = fmean[(size_t)i * K + k]; v[k] = fvar[(size_t)i * K + k]; }
    const double y = Y[i];
    if (lik == 0) {   // GaussianModified._variational_expectations, likelihoods.py:39-41 (per component)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double s2 = lik_var[k], r = y - mu[k];
            ve[k] = -HALF_LOG_2PI - 0.5 * log(s2) - 0.5 * (r * r + v[k]) / s2;
        }
    } else {          // gpflow MultiClass(RobustMax): one value per point, broadcast over the components
        const double eps = robustmax_eps();
        double d0[K], d1[K];
        int c = (int)y;
        c = c < 0 ? 0 : (c >= K ? K - 1 : c);
        const double p = robustmax_prob<K, false>(c, mu, v, d0, d1, squash);
        const double val = p * log(1.0 - eps) + (1.0 - p) * log(eps / (K - 1.0));
#pragma unroll
        for (int k = 0; k < K; ++k) ve[k] = val;
    }
    double m = -DBL_MAX, acc = 0.0;
    for (int s = 0; s < S; ++s) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) t += W[((size_t)s * n + i) * K + k] * ve[k];
        if (t > m) { acc = acc * exp(m - t) + 1.0; m = t; }
        else acc += exp(t - m);
    }
    out[i] = m + log(acc) - log((double)S);
}

void e_log_p_y_kernel(const double* fmean, const double* fvar, const double* Y, const double* lik_var, int lik,
                      double squash, const double* W, int S, int64_t n, int K, double* out, const Launch& ln) {
    if (n <= 0) return;
    const unsigned grid = (unsigned)((n + 127) / 128);
#define EL(KK) case KK: e_log_p_y_k<KK><<<grid, 128, 0, ln.stream>>>(fmean, fvar, Y, lik_var, lik, squash, W, S, n, out); break;
    switch (K) { EL(1) EL(2) EL(3) EL(4) EL(5) EL(6) EL(7) default: e_log_p_y_k<8><<<grid, 128, 0, ln.stream>>>(fmean, fvar, Y, lik_var, lik, squash, W, S, n, out); }
#undef EL
    ln.tick();
}

// ==================================================================================================
// The likelihood methods the reference exposes as stand-alone calls (broadcasting_lik.py:39-46, likelihoods.py:21-41,
// gpflow MultiClass): elementwise over R = S N rows of [.., K] arrays; Y is [N] and row r reads Y[r % y_period]
// (the reference tiles Y over the S axis, broadcasting_lik.py:27-31).
//   mode 0  variational_expectations: Gaussian -> [R, K] (not reduced over k, likelihoods.py:39-41); MultiClass -> [R]
//   mode 1  GaussianModified._scalar_log_prob: logdensities.gaussian(Y, F, variance) -> [R, K]   (likelihoods.py:21-22)
//   mode 2  GaussianModified._predict_log_density: sum_k gaussian(Y, Fmu, Fvar + variance) -> [R] (likelihoods.py:34-35)
// ==================================================================================================
template <int K>
__global__ void lik_eval_k(int mode, int lik, const double* lik_var, double squash, const double* Fmu, const double* Fvar,
                           const double* Y, int64_t rows, int64_t y_period, double* out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    double mu[K], v[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        mu[k] = Fmu[(size_t)r * K + k];
        v[k] = (mode == 1 || Fvar == nullptr) ? 0.0 : Fvar[(size_t)r * K + k];
    }
    const double y = Y[r % y_period];
    if (lik == 0) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double d = y - mu[k];
            double val;
            if (mode == 2) {
                const double s2 = v[k] + lik_var[k];
                val = -HALF_LOG_2PI - 0.5 * log(s2) - 0.5 * d * d / s2;
                acc += val;
            } else {
                const double s2 = lik_var[k];
                val = -HALF_LOG_2PI - 0.5 * log(s2) - 0.5 * (d * d + v[k]) / s2;
                out[(size_t)r * K + k] = val;
            }
        }
        if (mode == 2) out[r] = acc;
    } else {
        const double eps = robustmax_eps();
        double d0[K], d1[K];
        int c = (int)y;
        c = c < 0 ? 0 : (c >= K ? K - 1 : c);
        const double p = robustmax_prob<K, false>(c, mu, v, d0, d1, squash);
        out[r] = p * log(1.0 - eps) + (1.0 - p) * log(eps / (K - 1.0));
    }
}

void lik_eval_kernel(int mode, int lik, const double* lik_var, double squash, const double* Fmu, const double* Fvar,
                     const double* Y, int64_t rows, int64_t y_period, int K, double* out, const Launch& ln) {
    if (rows <= 0) return;
    const unsigned grid = (unsigned)((rows + 127) / 128);
#define LE(KK) case KK: lik_eval_k<KK><<<grid, 128, 0, ln.stream>>>(mode, lik, lik_var, squash, Fmu, Fvar, Y, rows, y_period, out); break;
    switch (K) { LE(1) LE(2) LE(3) LE(4) LE(5) LE(6) LE(7) default: lik_eval_k<8><<<grid, 128, 0, ln.stream>>>(mode, lik, lik_var, squash, Fmu, Fvar, Y, rows, y_period, out); }
#undef LE
    ln.tick();
}

// Philox4x32-10 known-answer entry (tests: Random123's kat_vectors) and the raw throughput-mode draws of one block of
// points, so that the uniform -> normal / Gumbel mapping and the counter packing can be tested statistically.
__global__ void philox_kat_k(const uint32_t* ctr_key, int n, uint32_t* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c[4] = {ctr_key[i * 6 + 0], ctr_key[i * 6 + 1], ctr_key[i * 6 + 2], ctr_key[i * 6 + 3]};
    philox4x32_10(c, ctr_key[i * 6 + 4], ctr_key[i * 6 + 5]);
    for (int j = 0; j < 4; ++j) out[i * 4 + j] = c[j];
}
void philox_kat_kernel(const uint32_t* ctr_key, int n, uint32_t* out, const Launch& ln) {
    if (n <= 0) return;
    philox_kat_k<<<(n + 127) / 128, 128, 0, ln.stream>>>(ctr_key, n, out);
    ln.tick();
}

template <int K>
__global__ void philox_draws_k(uint64_t seed, int64_t point_offset, int64_t n, int S, int stream, double* z, double* u) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)S * n) return;
    const int s = (int)(idx / n);
    const int64_t i = idx % n;
    double zz[K], uu[K];
    philox_draw_all<K>(seed, point_offset + i, s, stream, zz, uu);
#pragma unroll
    for (int k = 0; k < K; ++k) { z[idx * K + k] = zz[k]; u[idx * K + k] = uu[k]; }
}
void philox_draws_kernel(uint64_t seed, int64_t point_offset, int64_t n, int S, int K, int stream, double* z, double* u,
                         const Launch& ln) {
    const int64_t total = (int64_t)S * n;
    if (total <= 0) return;
    const unsigned grid = (unsigned)((total + 127) / 128);
#define PD(KK) case KK: philox_draws_k<KK><<<grid, 128, 0, ln.stream>>>(seed, point_offset, n, S, stream, z, u); break;
    switch (K) { PD(1) PD(2) PD(3) PD(4) PD(5) PD(6) PD(7) default: philox_draws_k<8><<<grid, 128, 0, ln.stream>>>(seed, point_offset, n, S, stream, z, u); }
#undef PD
    ln.tick();
}

}  // namespace mgp
