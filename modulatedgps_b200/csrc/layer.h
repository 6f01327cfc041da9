// Per-layer device workspace shared by the kernels of libmgp.
#pragma once
#include <stdint.h>

namespace mgp {

constexpr int KP = 8;  // components padded to one MMA fragment (MGP_MAX_K)

// One whitened SVGP layer on the device.  Mp = M rounded up to 32, Dp = D rounded up to 4.
// "fm" = fragment-major (common.cuh::wf_index), "rm" = row-major.
struct LayerDev {
    int M, D, K, Mp, Dp, n_ls;
    // borrowed inputs (constrained parameter values, caller-owned)
    const double *Z, *q_mu, *q_sqrt, *variance, *lengthscales;
    // precomputed once per step (replicated on every rank)
    double* inv_ls;   // [Dp]            1/lengthscale_d, 0 on padding
    double* Zs_rm;    // [Mp, Dp]   rm   Z / lengthscales, 0 on padding
    double* Zs_fm;    // [Mp, Dp]   fm   same, left operand of the r^2 contraction
    double* zs2;      // [Mp]            |Zs_i|^2
    double* zh;       // [Mp]            log(variance) - |Zs_i|^2 / 2  (row term of the Kuf exponent)
    double* Kuu;      // [Mp, Mp]   rm   k(Z,Z) + jitter I   (identity on padding)
    double* L;        // [Mp, Mp]   rm   chol(Kuu), strict upper = 0
    double* Linv;     // [Mp, Mp]   rm   L^-1
    double* Dinv;     // [Mp/32][32][32] rm  inverses of the 32 x 32 diagonal blocks of L (chol_kernel -> trinv_kernel)
    double* W_Linv;   // [Mp, Mp]   fm   L^-1            (lower)
    double* W_LinvT;  // [Mp, Mp]   fm   L^-T            (upper)
    double* Lq_rm;    // [K, Mp, Mp] rm  tril(q_sqrt), zero padded
    double* W_LqT;    // [K][Mp, Mp] fm  Lq_k^T          (upper)
    double* Q_rm;     // [Mp, K*Mp]  rm  [Q_0 | ... | Q_{K-1}],  Q_k = 2 (Lq_k Lq_k^T - I)
    double* W_Lq;     // [K][Mp, Mp] fm  Lq_k            (lower)
    double* W_m;      // [Mp, KP]   fm   q_mu padded to KP columns
    double* W_mT;     // [16, Mp]   fm   row k = q_mu[:, k] (k < K), 0 otherwise
    // backward accumulators / scratch
    double* T1;       // [K, Mp, Mp] rm scratch
    double* T2;       // [Mp, Mp] rm scratch
    double* T3;       // [Mp, Mp] rm scratch
    double* Sfull;    // [K, Mp, Mp] rm symmetric S_k
    double* rowout;   // [Mp, 2 Dp + 1] per-row sums of the Kuu kernel backward
};

// layout of the flat reduction buffer (doubles) -------------------------------------------------
constexpr int RB_DATA = 0;        // sum_n l_n / n_global
constexpr int RB_LIKVAR = 1;      // [KP] d/d lik_var (data term)
constexpr int RB_ALIKVAR = 9;     // [KP] d/d assign_lik_var
constexpr int RB_SUMV_PRED = 17;  // sum_{n,k} vbar (d/d variance through Knn)
constexpr int RB_SUMV_ASSIGN = 18;
constexpr int RB_HEADER = 32;

struct LayerRB {  // offsets into the reduce buffer for one layer
    int64_t S;     // [K, Mp, Mp]  lower 64x64 tiles valid
    int64_t mraw;  // [Mp, KP]     A mubar
    int64_t esum;  // [Mp, 1 + 2 Dp]
    int64_t end;
};

inline LayerRB layer_rb(int64_t base, int Mp, int Dp, int K) {
    LayerRB r;
    r.S = base;
    r.mraw = r.S + (int64_t)K * Mp * Mp;
    r.esum = r.mraw + (int64_t)Mp * KP;
    r.end = r.esum + (int64_t)Mp * (1 + 2 * Dp);
    return r;
}

}  // namespace mgp
