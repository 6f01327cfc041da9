// The N-streaming FP64 tensor-core kernels of the SVGP conditional, forward and backward.
//
// Layout.  The [Mp x N] workspace arrays A = L^-1 Kuf and B_k = Lq_k^T A are stored TILE-MAJOR: tile ti (NT points)
// is one contiguous [Mp][NT + 4] block — exactly the shared-memory image the kernels multiply from (row stride
// NT + 4 == 4 mod 16 doubles => conflict-free B-fragment LDS.64).  A whole right operand is therefore fetched by ONE
// cp.async.bulk (TMA, UBLKCP) completing on an mbarrier.
//
// Structure.  One persistent CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Right operands move
// through a ring of NBUF shared-memory buffers guarded by full[] / done[] mbarriers; there is no CTA-wide barrier in
// the steady state.  Warps 0-7 are DMMA consumers: each computes 16-row blocks  C = W[rows, k-range] * T  on
// DMMA.8x8x4 with the left operand W streamed from L2 in fragment-major order (one coalesced 256-byte load per
// fragment, software-pipelined across block / component / tile boundaries).  Triangular left operands only visit
// their non-zero k-range; 16-row blocks are dealt to the 8 warps in snake order so the triangular work balances.
// Extra warps feed the ring and take the per-tile epilogues off the consumers:
//
//   cond_fwd_a : warps 8-11 generate the Kuf tile (r^2 contraction on DMMA + exp)   A = L^-1 Kuf        -> A
//   cond_fwd_b : warp 8 loads the A tile, reduces the per-warp partial norms         B_k = Lq_k^T A      -> B_k, fmean, fvar
//   cond_bwd_a : warp 8 loads B_0..B_{K-1}, A and the adjoint slabs of the tile       Abar (see below)    -> Abar (over A)
//   cond_bwd_b : warp 8 loads the Abar tile and stages X                              E = (L^-T Abar) .* Kuf -> sums E [1, xs, xs^2]
//
// Reference arithmetic replaced: gpflow SquaredExponential.K + base_conditional as called from
// IndependentPosteriorSingleOutputModified._conditional_fused (MixtureGPs/models.py:129-144), evaluated once
// per point instead of once per (sample, point) (SGP.integrate only tiles X, models.py:35-36), and TF's
// reverse pass through it (utils/training_utils.py:8-10).  Math: SURVEY.md Appendix B.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "exp_tab.h"
#include "kernels.h"

namespace mgp {

// Per-warp phase timers (clock64 sums) for TEMPORARY profiling builds only: nvcc -DMGP_PHASE_TIMERS, read back with
// mgp_debug_phase_dump (tools/phase_timers.py).  Compiled out of the product.
#ifdef MGP_PHASE_TIMERS
__device__ long long g_phase[4][160][17][8];   // [kernel][cta][warp][phase]
#define PH_DECL long long ph_t = clock64(); long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PH_MARK(i) { const long long ph_now = clock64(); ph_acc[i] += ph_now - ph_t; ph_t = ph_now; }
#define PH_STORE(kid) { if ((threadIdx.x & 31) == 0 && blockIdx.x < 160) for (int ph_i = 0; ph_i < 8; ++ph_i) g_phase[kid][blockIdx.x][threadIdx.x >> 5][ph_i] = ph_acc[ph_i]; }
#else
#define PH_DECL
#define PH_MARK(i)
#define PH_STORE(kid)
#endif

// exp2_tab(y) = exp(y ln2 / 64) from a 64-entry table of 2^(j/64) and a degree-5 polynomial on |rr| <= 1/2 (rr in
// units of ln2/64): 9 FP64 instructions instead of libdevice's 16-18 for exp, at most 1.1 ulp from expl
// (tools/exp_tab_check.c, tests/test_host_logic.py).  The argument arrives ALREADY in units of ln2/64: the left operand
// of the z.x contraction (Zs_fm, prep_z_kernel) is pre-scaled by 64/ln2, so the DMMA result needs no multiplication for
// the range reduction, and k = round(y), rr = y - k are two exact additions.  Every scalar FP64 instruction here matters
// out of proportion to its pipe time: measured in the fused forward kernel, a warp-wide scalar
// FP64 instruction issued beside a saturated DMMA stream costs that sub-partition ~10 clocks, not 2.  The underflow test
// runs on the integer pipe (sign-and-exponent word of y).  `tab` is the shared-memory copy of d_exp_tab64.
__device__ const double d_exp_tab64[64] = {EXP_TAB64_VALUES};
__device__ __forceinline__ double exp2_tab(double y, const double* tab) {
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to the nearest integer in the low word
    const double tt = y + MAGIC;
    const int k = __double2loint(tt);
    const double rr = y - (tt - MAGIC);
    double p = fma(rr, EXP2_C5, EXP2_C4);
    p = fma(p, rr, EXP2_C3);
    p = fma(p, rr, EXP2_C2);
    p = fma(p, rr, EXP2_C1);
    p *= rr;
    const double tj = tab[k & 63];
    const double res = fma(tj, p, tj);
    const double scaled = __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
    // y < -700 * 64 / ln2 (results below 1e-304 are flushed; the exponent add would leave the normal range):
    // for negative doubles the high word grows with the magnitude
    return (unsigned)__double2hiint(y) > 0xC0EF8F17u ? 0.0 : scaled;
}

// DMMA consumer warps per CTA: 8 with 32-point tiles (two warps per SM sub-partition keep the pipe ~90 % busy when a
// fragment group is 32 DMMAs long), 16 with 16-point tiles (M > 352: a group is only 16 DMMAs = ~512 clocks of a shared
// pipe, shorter than the L2 latency of the next group's fragments — ncu at config #5: stall_long_scoreboard 5.0 per
// issue, DMMA pipe 69 %; tools/wloop_bench.cu: 27.0 TFLOP/s with 8 warps vs 34.9 with 16 at NT = 16).
constexpr int SK_WARPS = 8;
constexpr int SK_WARPS_NT16 = 16;
constexpr int SK_CTHREADS = SK_WARPS * 32;
__host__ __device__ constexpr int sk_warps(int nt) { return nt == 32 ? SK_WARPS : SK_WARPS_NT16; }

__host__ __device__ inline int xs_stride(int Dp) { return ((Dp - 4 + 15) / 16) * 16 + 4; }
// 8-wide feature blocks of the kernel-gradient sums: features {1, xs_d, xs_d^2}, 1 + 2 D of them
__host__ __device__ inline int esum_feature_blocks(int D) { return (1 + 2 * D + 7) / 8; }

// 16-row block dealt to warp w in round r (snake order); returns -1 past the end
template <int NW = SK_WARPS>
__device__ __forceinline__ int snake_block(int round, int warp, int nb16) {
    const int b = round * NW + ((round & 1) ? (NW - 1 - warp) : warp);
    return b < nb16 ? b : -1;
}
// number of 16-row blocks dealt to this warp (only the last snake round can be short)
template <int NW = SK_WARPS>
__device__ __forceinline__ int my_block_count(int warp, int nb16) {
    const int R = (nb16 + NW - 1) / NW;
    return R == 0 ? 0 : (snake_block<NW>(R - 1, warp, nb16) >= 0 ? R : R - 1);
}

template <int NF>
__device__ __forceinline__ void zero_acc(double (&acc)[2][NF][2]) {
#pragma unroll
    for (int mf = 0; mf < 2; ++mf)
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) acc[mf][nf][0] = acc[mf][nf][1] = 0.0;
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
// Groups are processed in PAIRS with two fragment register sets in ping-pong (no register copies on the loop
// back-edge); tools/wloop_bench.cu: 35.0 vs 32.6 TFLOP/s for the copy-based loop at 8 warps per SM.
// TRI says which fragments of a DIAGONAL group (the 16 x 16 block on the diagonal of a triangular left operand) are
// identically zero and skipped: TRI_LOWER — rows 0-7 x columns 8-15 (first row block, k4-blocks 2, 3); TRI_UPPER — rows
// 8-15 x columns 0-7 (second row block, k4-blocks 0, 1).  2 of the 8 (row block, k4-block) DMMA sets of that group, i.e.
// 32 of the 1088 sets of a 256-row triangular operand: executed / algorithmic work 17/16 -> 33/32.  `diag` is
// warp-uniform; the skipped DMMAs are predicated off (no pipe time).
constexpr int TRI_NONE = 0, TRI_LOWER = 1, TRI_UPPER = 2;
template <int NT, int TRI = TRI_NONE>
__device__ __forceinline__ void wgemm_group(const WFrag& f, const double* tb, int kb, double (&acc)[2][NT / 8][2],
                                            bool diag = false) {
    constexpr int NF = NT / 8, STR = NT + 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double* tr = tb + (size_t)(kb + j) * 4 * STR;
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) {
            const double b = tr[nf * 8];
            if (TRI == TRI_LOWER && j >= 2) { if (!diag) dmma(acc[0][nf], f.a0[j], b); }
            else dmma(acc[0][nf], f.a0[j], b);
            if (TRI == TRI_UPPER && j < 2) { if (!diag) dmma(acc[1][nf], f.a1[j], b); }
            else dmma(acc[1][nf], f.a1[j], b);
        }
    }
}
// TRI_LOWER: the segment ENDS with the diagonal group; TRI_UPPER: it STARTS with it.
template <int NT, int TRI = TRI_NONE>
__device__ __forceinline__ void wgemm_seg(const Seg& cur, int kb1, int C4, const double* Tsm, double (&acc)[2][NT / 8][2],
                                          int lane, WFrag& f, const Seg& nxt) {
    constexpr int STR = NT + 4;
    const int g = lane >> 2, t = lane & 3;
    const double* tb = Tsm + t * STR + g;
    // fragment pointers of this lane: group at k4-block kb of `cur` is at wc + kb * 32 (second row block + C4 * 32)
    const double* wc = cur.w + (size_t)cur.rb8 * C4 * 32 + lane;
    const double* wn = nxt.w + ((size_t)nxt.rb8 * C4 + nxt.kb0) * 32 + lane;
    const size_t rstride = (size_t)C4 * 32;
    auto load = [&](WFrag& d, const double* w0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            d.a0[j] = __ldg(w0 + j * 32);
            d.a1[j] = __ldg(w0 + rstride + j * 32);
        }
    };
    int kb = cur.kb0;
    WFrag n;
    if (((kb1 - kb) >> 2) & 1) {   // odd number of groups: one single step first
        load(n, (kb + 4 < kb1) ? wc + (size_t)(kb + 4) * 32 : wn);
        wgemm_group<NT, TRI>(f, tb, kb, acc, TRI == TRI_UPPER || kb + 4 == kb1);
        f = n;
        kb += 4;
    }
    for (; kb < kb1; kb += 8) {
        load(n, wc + (size_t)(kb + 4) * 32);
        wgemm_group<NT, TRI == TRI_UPPER ? TRI_UPPER : TRI_NONE>(f, tb, kb, acc, kb == cur.kb0);
        load(f, (kb + 8 < kb1) ? wc + (size_t)(kb + 8) * 32 : wn);
        wgemm_group<NT, TRI == TRI_LOWER ? TRI_LOWER : TRI_NONE>(n, tb, kb + 4, acc, kb + 8 == kb1);
    }
}

// Copy-free variant (used by cond_fwd_b, where it gains 1.3 %; the two-CTA kernels and cond_bwd_a lose with the second
// set live across their epilogues).  `a` and `b` are the two fragment register sets.  On entry `a` holds the first group of `cur`.  Returns true if, on
// exit, the first group of `nxt` sits in `b` (odd number of groups: the sets have swapped roles) and false if it
// sits in `a` — there is NO register copy: after a copy-based odd step every DMMA of the next group waited for the
// loads the copy had to wait for.  Callers keep both sets alive and alternate the argument order (WPair::run).
template <int NT, int TRI = TRI_NONE>
__device__ __forceinline__ bool wgemm_seg_sw(const Seg& cur, int kb1, int C4, const double* Tsm, double (&acc)[2][NT / 8][2],
                                          int lane, WFrag& a, WFrag& b, const Seg& nxt) {
    constexpr int STR = NT + 4;
    const int g = lane >> 2, t = lane & 3;
    const double* tb = Tsm + t * STR + g;
    // fragment pointers of this lane: group at k4-block kb of `cur` is at wc + kb * 32 (second row block + C4 * 32)
    const double* wc = cur.w + (size_t)cur.rb8 * C4 * 32 + lane;
    const double* wn = nxt.w + ((size_t)nxt.rb8 * C4 + nxt.kb0) * 32 + lane;
    const size_t rstride = (size_t)C4 * 32;
    auto load = [&](WFrag& d, const double* w0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            d.a0[j] = __ldg(w0 + j * 32);
            d.a1[j] = __ldg(w0 + rstride + j * 32);
        }
    };
    int kb = cur.kb0;
    const bool odd = ((kb1 - kb) >> 2) & 1;
    if (!odd) {
        for (; kb < kb1; kb += 8) {
            load(b, wc + (size_t)(kb + 4) * 32);
            wgemm_group<NT, TRI == TRI_UPPER ? TRI_UPPER : TRI_NONE>(a, tb, kb, acc, kb == cur.kb0);
            load(a, (kb + 8 < kb1) ? wc + (size_t)(kb + 8) * 32 : wn);
            wgemm_group<NT, TRI == TRI_LOWER ? TRI_LOWER : TRI_NONE>(b, tb, kb + 4, acc, kb + 8 == kb1);
        }
        return false;
    }
    load(b, (kb + 4 < kb1) ? wc + (size_t)(kb + 4) * 32 : wn);
    wgemm_group<NT, TRI>(a, tb, kb, acc, TRI == TRI_UPPER || kb + 4 == kb1);
    kb += 4;
    for (; kb < kb1; kb += 8) {   // roles swapped: b is current
        load(a, wc + (size_t)(kb + 4) * 32);
        wgemm_group<NT>(b, tb, kb, acc);
        load(b, (kb + 8 < kb1) ? wc + (size_t)(kb + 8) * 32 : wn);
        wgemm_group<NT, TRI == TRI_LOWER ? TRI_LOWER : TRI_NONE>(a, tb, kb + 4, acc, kb + 8 == kb1);
    }
    return true;
}
// the two fragment sets of a warp and which of them currently holds the next group
struct WPair {
    WFrag f, n;
    bool sw = false;
    template <int NT, int TRI = TRI_NONE>
    __device__ __forceinline__ void run(const Seg& cur, int kb1, int C4, const double* Tsm, double (&acc)[2][NT / 8][2],
                                        int lane, const Seg& nxt) {
        if (!sw) { if (wgemm_seg_sw<NT, TRI>(cur, kb1, C4, Tsm, acc, lane, f, n, nxt)) sw = true; }
        else { if (wgemm_seg_sw<NT, TRI>(cur, kb1, C4, Tsm, acc, lane, n, f, nxt)) sw = false; }
    }
};

// When the input dimension leaves two padding columns in the k4-blocks of the zs.xs contraction (D + 2 <= Dp: D = 1, 2,
// 5, 6, ...), the exponent's row term (log variance - |zs|^2/2) and column term (-|xs|^2/2) ride in them —
// Zs_fm[i][D] = row term, Zs_fm[i][D+1] = 1 (prep_z_kernel), Xs[n][D] = 1, Xs[n][D+1] = column term — and the DMMA
// returns the whole exponent: two FP64 additions per Kuf element less on the pipe the kernel is bound by.
__host__ __device__ inline bool kuf_fold(int D, int Dp) { return D + 2 <= Dp; }

// Xs[n][d] = X[n0+n][d] / lengthscale_d (0 outside the chunk / padding); xs2[n] = -|Xs_n|^2 / 2 * 64/ln2 (column term of
// the Kuf exponent for the path without free padding columns, in exp2_tab's units).  One warp.
template <int NT>
__device__ __forceinline__ void stage_x_warp(const LayerDev& ly, const ChunkBuffers& cb, int64_t n0, double* Xs,
                                             double* xs2, int lane) {
    const int Dp = ly.Dp, D = ly.D, XSTR = xs_stride(Dp);
    for (int idx = lane; idx < NT * Dp; idx += 32) {
        const int n = idx / Dp, d = idx % Dp;
        double v = 0.0;
        if (n0 + n < cb.n && d < D) v = cb.X[(size_t)(n0 + n) * D + d] / ly.lengthscales[ly.n_ls == 1 ? 0 : d];
        Xs[n * XSTR + d] = v;
    }
    __syncwarp();
    const bool fold = kuf_fold(D, Dp);
    for (int n = lane; n < NT; n += 32) {
        double s = 0.0;
        for (int d = 0; d < D; ++d) s += Xs[n * XSTR + d] * Xs[n * XSTR + d];
        xs2[n] = -0.5 * s * EXP_TAB_L;   // (units of ln2/64, like the contraction's result: see exp2_tab)
        if (fold) {   // the two free padding columns of the contraction carry the row and column terms of the exponent
            Xs[n * XSTR + D] = 1.0;
            Xs[n * XSTR + D + 1] = -0.5 * s;
        }
    }
    __syncwarp();
}

// Kuf values of the 8-row block rb8 in C-fragment layout: kv[nf][e] = k(z_{8 rb8+g}, x_{nf*8+2t+e}).
// K = variance exp(-r2 / 2), r2 = -2 zs.xs + (|zs|^2 + |xs|^2)  (gpflow square_distance + K_r2) is evaluated as
// exp(zs.xs + (log variance - |zs|^2/2) + (-|xs|^2/2)): the same cancellation as the reference's r2, two FP64
// instructions instead of four around the exponential.  The zs.xs contraction runs on DMMA (north_star:
// "squared-distance term on FP64 DMMA").  `etab`: shared-memory copy of d_exp_tab64.
// `zs`: the fragment-major scaled inducing inputs — ly.Zs_fm, or a shared-memory copy of it (the generation is a latency
// chain load -> DMMA -> exp -> store per 8-row block; with the copy its first link is a 30-clock LDS, not an L2 access).
template <int NT>
__device__ __forceinline__ void gen_kuf_block(const LayerDev& ly, int rb8, const double* Xs, const double* xs2,
                                              const double* etab, double (&kv)[NT / 8][2], int lane,
                                              const double* zs = nullptr, bool have_z0 = false, double z0 = 0.0) {
    constexpr int NF = NT / 8;
    const int g = lane >> 2, t = lane & 3;
    const int Dp = ly.Dp, XSTR = xs_stride(Dp), D4 = Dp >> 2;
#pragma unroll
    for (int nf = 0; nf < NF; ++nf) kv[nf][0] = kv[nf][1] = 0.0;
    for (int kd = 0; kd < D4; ++kd) {
        // (have_z0: the caller already holds k4-block 0 of this row block's Z fragments)
        const double a = (have_z0 && kd == 0) ? z0
                         : zs ? zs[((size_t)rb8 * D4 + kd) * 32 + lane] : __ldg(ly.Zs_fm + ((size_t)rb8 * D4 + kd) * 32 + lane);
#pragma unroll
        for (int nf = 0; nf < NF; ++nf) dmma(kv[nf], a, Xs[(nf * 8 + g) * XSTR + kd * 4 + t]);
    }
    const int i = rb8 * 8 + g;
    const bool live = i < ly.M;
    if (kuf_fold(ly.D, Dp)) {   // (warp-uniform) the contraction already holds the whole exponent
#pragma unroll
        for (int nf = 0; nf < NF; ++nf)
#pragma unroll
            for (int e = 0; e < 2; ++e) kv[nf][e] = live ? exp2_tab(kv[nf][e], etab) : 0.0;
        return;
    }
    const double zh = __ldg(ly.zh + i);
#pragma unroll
    for (int nf = 0; nf < NF; ++nf)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double arg = kv[nf][e] + (zh + xs2[nf * 8 + 2 * t + e]);
            kv[nf][e] = live ? exp2_tab(arg, etab) : 0.0;
        }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// shared-memory carve-up helper: mbarriers live at the front of dynamic smem (16-byte aligned region)
constexpr int SK_BAR_DOUBLES = 8;   // room for 2 x NBUF mbarriers (NBUF <= 2) + padding

// ==================================================================================================
// cond_fwd_a :  A tile = L^-1 * Kuf tile
// As many CTAs per SM as fit (three at config #4: 24 warps): every warp first generates rows of the Kuf tile (exp on
// the FP64 pipe), then multiplies.  The generation phase is scalar FP64 work that shares the pipe with DMMA, so it wants MANY warps to
// overlap with the other CTAs' DMMA phases — a dedicated generator-warp ring starves behind the consumers' DMMAs
// (measured: 9.1 ms vs 6.2 ms for this form at config #4).
// ==================================================================================================
template <int NT>
__global__ void __launch_bounds__(sk_warps(NT) * 32, NT == 32 ? 3 : 1) cond_fwd_a_kernel(LayerDev ly, ChunkBuffers cb, int ntiles) {
    constexpr int NF = NT / 8, STR = NT + 4, NW = sk_warps(NT);
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, XSTR = xs_stride(ly.Dp);
    double* T = smem;
    double* Xs = T + (size_t)Mp * STR;
    double* xs2 = Xs + NT * XSTR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, nb8 = Mp / 8, C4 = Mp / 4;
    const size_t tile_elems = (size_t)Mp * STR;
    __shared__ double etab[64];
    if (threadIdx.x < 64) etab[threadIdx.x] = d_exp_tab64[threadIdx.x];   // visible after the first CTA barrier below
    const int nmy = my_block_count<NW>(warp, nb16);
    auto seg_of = [&](int i) { const int b = snake_block<NW>(i, warp, nb16); return Seg{ly.W_Linv, 2 * b, 0}; };
    WFrag wf;
    if (nmy > 0) wfrag_load(wf, seg_of(0), C4, 0, lane);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t n0 = (int64_t)tile * NT;
        double* Aout = cb.A + (size_t)tile * tile_elems;
        if (warp == 0) stage_x_warp<NT>(ly, cb, n0, Xs, xs2, lane);
        // the first Z fragment of each 8-row block this warp generates, requested BEFORE the barrier: the generation is a
        // latency chain (L2 load -> DMMA -> exp -> store) per block, and this takes its first link out of the chain
        constexpr int ZQ = 6;   // blocks warp, warp + NW, ... (Mp <= 352 at 32-point tiles: at most 6 per warp)
        double z0[ZQ];
#pragma unroll
        for (int q = 0; q < ZQ; ++q) {
            const int rb = warp + q * NW;
            z0[q] = rb < nb8 ? __ldg(ly.Zs_fm + (size_t)rb * (ly.Dp >> 2) * 32 + lane) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < ZQ; ++q) {
            const int rb = warp + q * NW;
            if (rb >= nb8) break;
            double kv[NF][2];
            gen_kuf_block<NT>(ly, rb, Xs, xs2, etab, kv, lane, nullptr, true, z0[q]);
#pragma unroll
            for (int nf = 0; nf < NF; ++nf)
                *reinterpret_cast<double2*>(T + (size_t)(rb * 8 + g) * STR + nf * 8 + 2 * t) = make_double2(kv[nf][0], kv[nf][1]);
        }
        for (int rb = warp + ZQ * NW; rb < nb8; rb += NW) {   // (larger M: 16-point tiles)
            double kv[NF][2];
            gen_kuf_block<NT>(ly, rb, Xs, xs2, etab, kv, lane);
#pragma unroll
            for (int nf = 0; nf < NF; ++nf)
                *reinterpret_cast<double2*>(T + (size_t)(rb * 8 + g) * STR + nf * 8 + 2 * t) = make_double2(kv[nf][0], kv[nf][1]);
        }
        // The generated tile is kept for cond_bwd_b (E = Kuf_bar .* Kuf) instead of being generated a second time there:
        // ONE bulk store of the shared-memory image by one thread, read by the async proxy while the warps multiply.
        if (cb.Kuf) fence_proxy_async();
        __syncthreads();
        if (cb.Kuf && threadIdx.x == 0) bulk_s2g(cb.Kuf + (size_t)tile * tile_elems, T, (unsigned)(tile_elems * sizeof(double)));
        for (int i = 0; i < nmy; ++i) {
            const int b = snake_block<NW>(i, warp, nb16);
            double acc[2][NF][2];
            zero_acc<NF>(acc);
            // (TRI_LOWER measured slower here, 5.78 vs 5.66 ms at config #4: the two CTAs' phases drift apart)
            wgemm_seg<NT>(seg_of(i), (b + 1) * 4, C4, T, acc, lane, wf, seg_of(i + 1 < nmy ? i + 1 : 0));   // lower triangular
#pragma unroll
            for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
                    *reinterpret_cast<double2*>(Aout + (size_t)(b * 16 + mf * 8 + g) * STR + nf * 8 + 2 * t) =
                        make_double2(acc[mf][nf][0], acc[mf][nf][1]);
        }
        if (cb.Kuf && threadIdx.x == 0) bulk_wait_read();   // T is overwritten after the barrier
        __syncthreads();
    }
}

// ==================================================================================================
// cond_fwd_a, software-pipelined form (32-point tiles, two tile buffers):  ONE persistent CTA per SM.
// The barrier-phased form above leaves the DMMA pipe idle 22 % of the time (ncu: stall_barrier 4.4 per issue): its three
// CTAs per SM fall into step — they share the pipe, so CTAs that multiply together finish together and then generate
// together.  Here every warp generates ITS rows of tile i + 1 (into the other buffer) between the row blocks of tile i
// it multiplies, so a warp's generation phase (a latency-bound chain: L2 load -> 4 DMMA -> table exp -> store) always
// runs beside the other warp of its sub-partition multiplying; a dedicated generator-warp ring starved (DESIGN.md
// section 5), warps that alternate cannot.  One CTA-wide barrier per tile; warp 8 stages the X rows two tiles ahead.
// ==================================================================================================
template <int NT>
__global__ void __launch_bounds__(SK_WARPS * 32 + 32, 1) cond_fwd_a_pipe_kernel(LayerDev ly, ChunkBuffers cb, int ntiles,
                                                                               int zs_in_smem) {
    constexpr int NF = NT / 8, STR = NT + 4, NW = SK_WARPS;
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, XSTR = xs_stride(ly.Dp);
    const size_t tile_elems = (size_t)Mp * STR;
    double* Tb = smem;                                   // [2][Mp][STR]
    double* Xsb = Tb + 2 * tile_elems;                   // [2]{[NT][XSTR], [NT]}
    const int xs_elems = NT * XSTR + NT;
    double* zsm = Xsb + 2 * xs_elems;                    // [Mp * Dp] copy of Zs_fm (zs_in_smem)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, nb8 = Mp / 8, C4 = Mp / 4;
    __shared__ double etab[64];
    if (threadIdx.x < 64) etab[threadIdx.x] = d_exp_tab64[threadIdx.x];
    if (zs_in_smem)
        for (int i = threadIdx.x; i < Mp * ly.Dp; i += blockDim.x) zsm[i] = ly.Zs_fm[i];
    const double* zs = zs_in_smem ? zsm : nullptr;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto tile_of = [&](int i) { return (int64_t)blockIdx.x + (int64_t)i * gridDim.x; };
    auto stage = [&](int i) {   // warp NW: scaled X rows of this CTA's i-th tile
        double* Xs = Xsb + (size_t)(i & 1) * xs_elems;
        stage_x_warp<NT>(ly, cb, tile_of(i) * NT, Xs, Xs + NT * XSTR, lane);
    };
    // rows [8 rb, 8 rb + 8) of the Kuf tile i -> buffer i & 1
    auto gen = [&](int i, int rb) {
        const double* Xs = Xsb + (size_t)(i & 1) * xs_elems;
        double* T = Tb + (size_t)(i & 1) * tile_elems;
        double kv[NF][2];
        gen_kuf_block<NT>(ly, rb, Xs, Xs + NT * XSTR, etab, kv, lane, zs);
#pragma unroll
        for (int nf = 0; nf < NF; ++nf)
            *reinterpret_cast<double2*>(T + (size_t)(rb * 8 + g) * STR + nf * 8 + 2 * t) = make_double2(kv[nf][0], kv[nf][1]);
    };
    const int ngen = warp < NW ? (nb8 - warp + NW - 1) / NW : 0;   // 8-row blocks warp, warp + NW, ... of every tile
    if (warp == NW && my_tiles > 0) stage(0);
    __syncthreads();
    if (my_tiles > 0) {
        if (warp == NW) { if (my_tiles > 1) stage(1); }
        else for (int q = 0; q < ngen; ++q) gen(0, warp + q * NW);
    }
    if (cb.Kuf) fence_proxy_async();
    __syncthreads();

    const int nmy = warp < NW ? my_block_count<NW>(warp, nb16) : 0;
    auto seg_of = [&](int i) { const int b = snake_block<NW>(i, warp, nb16); return Seg{ly.W_Linv, 2 * b, 0}; };
    WFrag wf;
    if (nmy > 0) wfrag_load(wf, seg_of(0), C4, 0, lane);
    for (int i = 0; i < my_tiles; ++i) {
        const int64_t tile = tile_of(i);
        const double* T = Tb + (size_t)(i & 1) * tile_elems;
        const bool more = i + 1 < my_tiles;
        if (warp == NW) {
            // kept for cond_bwd_b: one bulk store of the finished Kuf tile (its writers fenced before the last barrier)
            if (cb.Kuf && lane == 0) bulk_s2g(cb.Kuf + (size_t)tile * tile_elems, T, (unsigned)(tile_elems * sizeof(double)));
            if (i + 2 < my_tiles) stage(i + 2);     // buffer i & 1 of the X rows: tile i's generation ended before the last barrier
            if (cb.Kuf && lane == 0) bulk_wait_read();   // buffer i & 1 of T is generated into again after the next barrier
        } else {
            double* Aout = cb.A + (size_t)tile * tile_elems;
            int q = 0;
            for (int r = 0; r < nmy; ++r) {
                const int b = snake_block<NW>(r, warp, nb16);
                double acc[2][NF][2];
                zero_acc<NF>(acc);
                wgemm_seg<NT>(seg_of(r), (b + 1) * 4, C4, T, acc, lane, wf, seg_of(r + 1 < nmy ? r + 1 : 0));   // lower triangular
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf)
                        *reinterpret_cast<double2*>(Aout + (size_t)(b * 16 + mf * 8 + g) * STR + nf * 8 + 2 * t) =
                            make_double2(acc[mf][nf][0], acc[mf][nf][1]);
                if (more) {   // this warp's share of the next tile's generation, spread over its row blocks
                    const int q1 = (r + 1) * ngen / nmy;
                    for (; q < q1; ++q) gen(i + 1, warp + q * NW);
                }
            }
            if (more) for (; q < ngen; ++q) gen(i + 1, warp + q * NW);
            if (cb.Kuf) fence_proxy_async();
        }
        __syncthreads();
    }
}

// ==================================================================================================
// cond_fwd_b :  B_k = Lq_k^T A (kept for the backward), fvar = variance - |a|^2 + sum_m B_k^2, fmean = q_mu^T A
// warps 0-7 multiply and leave per-warp partial column sums; warp 8 loads tiles and finishes the sums
// ==================================================================================================
template <int NT, int NBUF>
__global__ void __launch_bounds__(sk_warps(NT) * 32 + 32, 1) cond_fwd_b_kernel(LayerDev ly, ChunkBuffers cb, int ntiles) {
    constexpr int NF = NT / 8, STR = NT + 4, NW = sk_warps(NT);
    constexpr int MW = SK_WARPS;   // warps that take a k-slice of the fmean contraction (C4 = Mp / 4 is a multiple of 8)
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, K = ly.K;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* done = full + NBUF;
    double* Tb = smem + SK_BAR_DOUBLES;                          // [NBUF][Mp][STR]
    double* sqpart = Tb + (size_t)NBUF * Mp * STR;               // [NBUF][NW][K][NT]  partial sum_m B_k^2
    double* mnpart = sqpart + (size_t)NBUF * NW * K * NT;        // [NBUF][MW][K][NT]  partial q_mu^T A
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const size_t tile_elems = (size_t)Mp * STR;
    const unsigned tile_bytes = (unsigned)(tile_elems * sizeof(double));
    if (threadIdx.x == 0) {
        for (int i = 0; i < NBUF; ++i) { mbar_init(&full[i], 1); mbar_init(&done[i], NW); }
        mbar_fence_init();
    }
    __syncthreads();
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto tile_of = [&](int i) { return (int64_t)blockIdx.x + (int64_t)i * gridDim.x; };

    if (warp == NW) {   // ---- loader + finisher ----
        auto issue = [&](int i) {
            if (lane == 0) {
                const int buf = i % NBUF;
                mbar_arrive_expect_tx(&full[buf], tile_bytes);
                bulk_g2s(Tb + (size_t)buf * tile_elems, cb.A + (size_t)tile_of(i) * tile_elems, tile_bytes, &full[buf]);
            }
        };
        for (int i = 0; i < NBUF && i < my_tiles; ++i) issue(i);
        const double variance = ly.variance[0];
        for (int i = 0; i < my_tiles; ++i) {
            const int buf = i % NBUF;
            const int64_t n0 = tile_of(i) * NT;
            const double* T = Tb + (size_t)buf * tile_elems;
            // |a_n|^2 of this lane's column(s) while the consumers are still multiplying (the tile has landed once
            // full[buf] completes; waiting on it here is harmless)
            mbar_wait(&full[buf], (unsigned)((i / NBUF) & 1));
            double asq[(NT + 31) / 32];
#pragma unroll
            for (int c = 0; c < (NT + 31) / 32; ++c) asq[c] = 0.0;
            if (lane < NT) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                for (int m = 0; m < Mp; m += 4) {
                    const double v0 = T[(size_t)m * STR + lane], v1 = T[(size_t)(m + 1) * STR + lane];
                    const double v2 = T[(size_t)(m + 2) * STR + lane], v3 = T[(size_t)(m + 3) * STR + lane];
                    s0 = fma(v0, v0, s0); s1 = fma(v1, v1, s1); s2 = fma(v2, v2, s2); s3 = fma(v3, v3, s3);
                }
                asq[0] = (s0 + s1) + (s2 + s3);
            }
            mbar_wait(&done[buf], (unsigned)((i / NBUF) & 1));   // every consumer has left its partials and the tile
            const double* sq = sqpart + (size_t)buf * NW * K * NT;
            const double* mn = mnpart + (size_t)buf * MW * K * NT;
            if (lane < NT) {
                for (int k = 0; k < K; ++k) {
                    double s = 0.0, m = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) s += sq[((size_t)w * K + k) * NT + lane];
#pragma unroll
                    for (int w = 0; w < MW; ++w) m += mn[((size_t)w * K + k) * NT + lane];
                    cb.fvar[(size_t)(n0 + lane) * K + k] = (variance - asq[0]) + s;   // Knn - sum A^2 + sum LTA^2
                    cb.fmean[(size_t)(n0 + lane) * K + k] = m;
                }
            }
            __syncwarp();
            if (i + NBUF < my_tiles) issue(i + NBUF);
        }
        return;
    }
    // ---- consumers ----
    const int nmy = my_block_count<NW>(warp, nb16);
    // the fmean contraction is split over the warps by k-range: warp w takes k4-blocks [w C4/8, (w+1) C4/8)
    const int mkb0 = warp * (C4 / MW), mkb1 = mkb0 + C4 / MW;   // C4 = Mp/4 is a multiple of 8; warps >= MW skip it
    auto seg_of = [&](int k, int r) {
        const int b = snake_block<NW>(r, warp, nb16);
        return Seg{ly.W_LqT + (size_t)k * Mp * Mp, 2 * b, 4 * b};
    };
    WPair wp;
    if (nmy > 0) { const Seg s0 = seg_of(0, 0); wfrag_load(wp.f, s0, C4, s0.kb0, lane); }
    PH_DECL
    for (int i = 0; i < my_tiles; ++i) {
        const int buf = i % NBUF;
        const int64_t tile = tile_of(i);
        const double* T = Tb + (size_t)buf * tile_elems;
        double* sq = sqpart + ((size_t)buf * NW + warp) * K * NT;
        double* mn = mnpart + ((size_t)buf * MW + (warp < MW ? warp : 0)) * K * NT;
        PH_MARK(5)
        mbar_wait(&full[buf], (unsigned)((i / NBUF) & 1));
        PH_MARK(0)
        for (int k = 0; k < K; ++k) {
            double colsq[NF][2];
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) colsq[nf][0] = colsq[nf][1] = 0.0;
            double* Bk = cb.Bk ? cb.Bk + ((size_t)k * cb.tiles_cap + tile) * tile_elems : nullptr;
            for (int r = 0; r < nmy; ++r) {
                const int b = snake_block<NW>(r, warp, nb16);
                double acc[2][NF][2];
                zero_acc<NF>(acc);
                const Seg nxt = (r + 1 < nmy) ? seg_of(k, r + 1) : seg_of(k + 1 < K ? k + 1 : 0, 0);
                wp.template run<NT, TRI_UPPER>(seg_of(k, r), C4, C4, T, acc, lane, nxt);   // upper triangular
                PH_MARK(1)
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) {
                        if (Bk)   // kept for the backward: Abar needs Lq_k (2 B_k diag(vbar_k))
                            *reinterpret_cast<double2*>(Bk + (size_t)(b * 16 + mf * 8 + g) * STR + nf * 8 + 2 * t) =
                                make_double2(acc[mf][nf][0], acc[mf][nf][1]);
                        colsq[nf][0] = fma(acc[mf][nf][0], acc[mf][nf][0], colsq[nf][0]);
                        colsq[nf][1] = fma(acc[mf][nf][1], acc[mf][nf][1], colsq[nf][1]);
                    }
                PH_MARK(2)
            }
            if (NF == 4) {   // 8 values per lane: butterfly with hand-over (7 shuffles), lane (g, t) ends with column
                             // (g >> 1) * 8 + 2 t + (g & 1) — every lane stores one distinct column
                double v[8];
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) { v[2 * nf] = colsq[nf < NF ? nf : 0][0]; v[2 * nf + 1] = colsq[nf < NF ? nf : 0][1]; }
                sq[(size_t)k * NT + (g >> 1) * 8 + 2 * t + (g & 1)] = reduce8_over_g(v, lane);
            } else {
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double s = sum_over_g(colsq[nf][e]);
                        if (g == 0) sq[(size_t)k * NT + nf * 8 + 2 * t + e] = s;
                    }
            }
        }
        PH_MARK(3)
        if (warp < MW) {   // this warp's slice of fmean^T [K x NT] = q_mu^T [K x Mp] * A tile   (rows >= K of W_mT are zero).
            // Two accumulator sets (even / odd k4-blocks): the contraction is a short dependent DMMA chain.
            double acc[2][NF][2];
            zero_acc<NF>(acc);
            const double* wm = ly.W_mT + (size_t)mkb0 * 32 + lane;
            const double* tb = T + t * STR + g;
            for (int kb = mkb0; kb < mkb1; kb += 2) {   // (mkb1 - mkb0) = Mp / 32 ... a multiple of 2 when Mp % 64 == 0
                const double a0 = __ldg(wm + (size_t)(kb - mkb0) * 32);
                const double* tr0 = tb + (size_t)kb * 4 * STR;
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) dmma(acc[0][nf], a0, tr0[nf * 8]);
                if (kb + 1 < mkb1) {
                    const double a1 = __ldg(wm + (size_t)(kb + 1 - mkb0) * 32);
                    const double* tr1 = tb + (size_t)(kb + 1) * 4 * STR;
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) dmma(acc[1][nf], a1, tr1[nf * 8]);
                }
            }
            if (g < K) {
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
                    *reinterpret_cast<double2*>(mn + (size_t)g * NT + nf * 8 + 2 * t) =
                        make_double2(acc[0][nf][0] + acc[1][nf][0], acc[0][nf][1] + acc[1][nf][1]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[buf]);
        PH_MARK(4)
    }
    PH_STORE(0)
}

// ==================================================================================================
// cond_fwd_fused (32-point tiles, M <= ~300):  cond_fwd_a and cond_fwd_b in ONE persistent kernel.
//   generator warps (4)   Kuf tile of point tile i + 1 -> buffer G  (z.x contraction on DMMA + table exp), while
//   consumer warps (8)    phase 1: A = L^-1 G  -> buffer Abuf (+ global A, for SYRK and the backward)
//                         phase 2: K passes B_k = Lq_k^T Abuf -> global B_k, partial column norms / q_mu^T A
//   warp 8 (a generator)  also finishes fmean / fvar of the previous tile.
// Why: stand-alone, cond_fwd_a has only M^2 of DMMA per tile to hide its scalar-FP64 generation phase behind (three
// barrier-phased CTAs per SM, DMMA pipe 72 %); here the generation runs on its own warps under (1 + K) M^2 of DMMA and
// the L^-1 product runs at the rate of the other passes.  The A tile never travels HBM -> SM for the B_k passes, and
// one launch (fill + tail) per layer disappears.
// Hand-offs per tile (mbarriers, phase = tile parity):   g_full  G written (4 generator warps)
//   g_free  consumers done reading G (8)        a_ready  Abuf written (8)
//   a_free  consumers done reading Abuf + finisher done with |a|^2 (9)        p_done  partial sums written (8)
// A consumer computes the L^-1 product of the NEXT tile's first row block before it waits for a_free, so the skew
// between warps at the tile boundary is absorbed by work.
// ==================================================================================================
constexpr int FU_GEN_WARPS = 4;
template <int NT>
__global__ void __launch_bounds__((SK_WARPS + FU_GEN_WARPS) * 32, 1) cond_fwd_fused_kernel(LayerDev ly, ChunkBuffers cb, int ntiles, int dbg) {
    constexpr int NF = NT / 8, STR = NT + 4, NW = SK_WARPS;
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, K = ly.K, XSTR = xs_stride(ly.Dp);
    uint64_t* g_full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* g_free = g_full + 1;
    uint64_t* a_ready = g_full + 2;
    uint64_t* a_free = g_full + 3;
    uint64_t* p_done = g_full + 4;
    const size_t tile_elems = (size_t)Mp * STR;
    double* G = smem + SK_BAR_DOUBLES;                          // [Mp][STR]  Kuf tile
    double* Abuf = G + tile_elems;                              // [Mp][STR]  A tile
    double* sqpart = Abuf + tile_elems;                         // [2][NW][K][NT]  partial sum_m B_k^2 (by tile parity)
    double* mnpart = sqpart + (size_t)2 * NW * K * NT;          // [2][NW][K][NT]  partial q_mu^T A
    double* Xsb = mnpart + (size_t)2 * NW * K * NT;             // [2]{[NT][XSTR], [NT]} scaled X rows (by tile parity)
    const int xs_elems = NT * XSTR + NT;
    __shared__ double etab[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, nb8 = Mp / 8, C4 = Mp / 4;
    if (threadIdx.x < 64) etab[threadIdx.x] = d_exp_tab64[threadIdx.x];
    if (threadIdx.x == 0) {
        mbar_init(g_full, FU_GEN_WARPS); mbar_init(g_free, NW); mbar_init(a_ready, NW); mbar_init(a_free, NW + 1);
        mbar_init(p_done, NW);
        mbar_fence_init();
    }
    __syncthreads();
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto tile_of = [&](int i) { return (int64_t)blockIdx.x + (int64_t)i * gridDim.x; };

    if (warp >= NW) {   // ---- generators (+ finisher = warp NW) ----
        const int gw = warp - NW;
        const double variance = ly.variance[0];
        auto finish = [&](int i) {   // warp NW: fmean / fvar of tile i
            const int64_t n0 = tile_of(i) * NT;
            const unsigned ph = (unsigned)(i & 1);
            mbar_wait(a_ready, ph);
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int m = 0; m < Mp; m += 4) {
                const double v0 = Abuf[(size_t)m * STR + lane], v1 = Abuf[(size_t)(m + 1) * STR + lane];
                const double v2 = Abuf[(size_t)(m + 2) * STR + lane], v3 = Abuf[(size_t)(m + 3) * STR + lane];
                s0 = fma(v0, v0, s0); s1 = fma(v1, v1, s1); s2 = fma(v2, v2, s2); s3 = fma(v3, v3, s3);
            }
            const double asq = (s0 + s1) + (s2 + s3);
            __syncwarp();
            if (lane == 0) mbar_arrive(a_free);
            mbar_wait(p_done, ph);
            const double* sq = sqpart + (size_t)(i & 1) * NW * K * NT;
            const double* mn = mnpart + (size_t)(i & 1) * NW * K * NT;
            for (int k = 0; k < K; ++k) {
                double sv = 0.0, mv = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    sv += sq[((size_t)w * K + k) * NT + lane];
                    mv += mn[((size_t)w * K + k) * NT + lane];
                }
                cb.fvar[(size_t)(n0 + lane) * K + k] = (variance - asq) + sv;   // Knn - sum A^2 + sum LTA^2
                cb.fmean[(size_t)(n0 + lane) * K + k] = mv;
            }
        };
        for (int i = 0; i < my_tiles; ++i) {
            double* Xs = Xsb + (size_t)(i & 1) * xs_elems;
            double* xs2 = Xs + NT * XSTR;
            if (gw == 0) stage_x_warp<NT>(ly, cb, tile_of(i) * NT, Xs, xs2, lane);
            named_bar_sync(2, FU_GEN_WARPS * 32);                           // X rows visible to the four generator warps
            if (i > 0) mbar_wait(g_free, (unsigned)((i - 1) & 1));          // consumers have left G
            for (int rb = gw; rb < nb8; rb += FU_GEN_WARPS) {
                double kv[NF][2];
                if (dbg & 1) {   // timing experiment only: no generation arithmetic
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) kv[nf][0] = kv[nf][1] = 1e-3;
                } else {
                    gen_kuf_block<NT>(ly, rb, Xs, xs2, etab, kv, lane);
                }
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
                    *reinterpret_cast<double2*>(G + (size_t)(rb * 8 + g) * STR + nf * 8 + 2 * t) = make_double2(kv[nf][0], kv[nf][1]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(g_full);
            if (gw == 0 && i > 0) finish(i - 1);
        }
        if (gw == 0 && my_tiles > 0) finish(my_tiles - 1);
        return;
    }

    // ---- consumers ----
    const int nmy = my_block_count<NW>(warp, nb16);
    const int mkb0 = warp * (C4 / NW), mkb1 = mkb0 + C4 / NW;   // this warp's k-slice of the fmean contraction
    auto seg_a = [&](int r) { const int b = snake_block<NW>(r, warp, nb16); return Seg{ly.W_Linv, 2 * b, 0}; };
    auto seg_b = [&](int k, int r) {
        const int b = snake_block<NW>(r, warp, nb16);
        return Seg{ly.W_LqT + (size_t)k * Mp * Mp, 2 * b, 4 * b};
    };
    WPair wp;
    if (nmy > 0) wfrag_load(wp.f, seg_a(0), C4, 0, lane);
    for (int i = 0; i < my_tiles; ++i) {
        const int64_t tile = tile_of(i);
        const unsigned ph = (unsigned)(i & 1);
        double* Aout = cb.A + (size_t)tile * tile_elems;
        double* sq = sqpart + ((size_t)(i & 1) * NW + warp) * K * NT;
        double* mn = mnpart + ((size_t)(i & 1) * NW + warp) * K * NT;
        // ---- phase 1: rows of A = L^-1 G (lower triangular) ----
        mbar_wait(g_full, ph);
        for (int r = 0; r < nmy; ++r) {
            const int b = snake_block<NW>(r, warp, nb16);
            double acc[2][NF][2];
            zero_acc<NF>(acc);
            const Seg nxt = (r + 1 < nmy) ? seg_a(r + 1) : seg_b(0, 0);
            wp.template run<NT, TRI_LOWER>(seg_a(r), (b + 1) * 4, C4, G, acc, lane, nxt);
            if (r == 0 && i > 0) mbar_wait(a_free, (unsigned)((i - 1) & 1));   // Abuf of the previous tile is no longer read
#pragma unroll
            for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) {
                    const size_t off = (size_t)(b * 16 + mf * 8 + g) * STR + nf * 8 + 2 * t;
                    const double2 v = make_double2(acc[mf][nf][0], acc[mf][nf][1]);
                    *reinterpret_cast<double2*>(Abuf + off) = v;
                    if (!(dbg & 2)) *reinterpret_cast<double2*>(Aout + off) = v;
                }
        }
        if (nmy == 0 && i > 0) mbar_wait(a_free, (unsigned)((i - 1) & 1));
        __syncwarp();
        if (lane == 0) { mbar_arrive(g_free); mbar_arrive(a_ready); }
        mbar_wait(a_ready, ph);
        // ---- phase 2: B_k = Lq_k^T A (upper triangular), partial norms and means ----
        const double* T = Abuf;
        for (int k = 0; k < K; ++k) {
            double colsq[NF][2];
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) colsq[nf][0] = colsq[nf][1] = 0.0;
            double* Bk = cb.Bk ? cb.Bk + ((size_t)k * cb.tiles_cap + tile) * tile_elems : nullptr;
            for (int r = 0; r < nmy; ++r) {
                const int b = snake_block<NW>(r, warp, nb16);
                double acc[2][NF][2];
                zero_acc<NF>(acc);
                const Seg nxt = (r + 1 < nmy) ? seg_b(k, r + 1) : (k + 1 < K ? seg_b(k + 1, 0) : seg_a(0));
                wp.template run<NT, TRI_UPPER>(seg_b(k, r), C4, C4, T, acc, lane, nxt);
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) {
                        if (Bk)
                            *reinterpret_cast<double2*>(Bk + (size_t)(b * 16 + mf * 8 + g) * STR + nf * 8 + 2 * t) =
                                make_double2(acc[mf][nf][0], acc[mf][nf][1]);
                        colsq[nf][0] = fma(acc[mf][nf][0], acc[mf][nf][0], colsq[nf][0]);
                        colsq[nf][1] = fma(acc[mf][nf][1], acc[mf][nf][1], colsq[nf][1]);
                    }
            }
            double v[8];
#pragma unroll
            for (int nf = 0; nf < 4; ++nf) { v[2 * nf] = colsq[nf < NF ? nf : 0][0]; v[2 * nf + 1] = colsq[nf < NF ? nf : 0][1]; }
            sq[(size_t)k * NT + (g >> 1) * 8 + 2 * t + (g & 1)] = reduce8_over_g(v, lane);
        }
        {   // this warp's slice of fmean^T [K x NT] = q_mu^T [K x Mp] * A tile (rows >= K of W_mT are zero)
            double acc[2][NF][2];
            zero_acc<NF>(acc);
            const double* wm = ly.W_mT + (size_t)mkb0 * 32 + lane;
            const double* tb = T + t * STR + g;
            for (int kb = mkb0; kb < mkb1; kb += 2) {
                const double a0 = __ldg(wm + (size_t)(kb - mkb0) * 32);
                const double* tr0 = tb + (size_t)kb * 4 * STR;
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) dmma(acc[0][nf], a0, tr0[nf * 8]);
                if (kb + 1 < mkb1) {
                    const double a1 = __ldg(wm + (size_t)(kb + 1 - mkb0) * 32);
                    const double* tr1 = tb + (size_t)(kb + 1) * 4 * STR;
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) dmma(acc[1][nf], a1, tr1[nf * 8]);
                }
            }
            if (g < K) {
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
                    *reinterpret_cast<double2*>(mn + (size_t)g * NT + nf * 8 + 2 * t) =
                        make_double2(acc[0][nf][0] + acc[1][nf][0], acc[0][nf][1] + acc[1][nf][1]);
            }
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive(a_free); mbar_arrive(p_done); }
    }
}

// ==================================================================================================
// cond_bwd_a  (triangular route)
//   Abar = sum_k Lq_k * (B_k diag(2 vbar_k))  +  q_mu * mubar^T  -  2 A diag(sum_k vbar_k)
// from fvar = variance - |a|^2 + sum |b_k|^2, b_k = Lq_k^T a, fmean = a^T q_mu.  The right operand changes with k:
// per point tile the ring carries K stages (B_0 .. B_{K-1}); each warp keeps the accumulators of ALL its row blocks
// (NBW of them) in registers across the stages of a tile.  The A tile of the elementwise epilogue is NOT a ring stage:
// it used to be a (K + 1)-th, very short one, and in a two-deep ring the stage after a short stage has only that
// stage's duration to land (B_0 of the next tile: ~1.5 k clocks for a 74 KB copy, ~4 k clocks exposed per tile).  The
// tile is pulled into L2 when the tile's first stage is issued, and each lane reads the 16 bytes it overwrites
// straight from there in the epilogue that follows the last stage.
// PROD_WARP: warp 8 feeds the ring; otherwise (accumulators too large for the 168-register cap of a 9-warp CTA: three
// warps on one SM sub-partition) the last warp to leave a stage refills its buffer (shared-memory arrival counter).
// ==================================================================================================
template <int NT, int NBUF, int NBW, bool PROD_WARP>
__global__ void __launch_bounds__(sk_warps(NT) * 32 + (PROD_WARP ? 32 : 0), 1)
    cond_bwd_a_kernel(LayerDev ly, ChunkBuffers cb, int ntiles) {
    constexpr int NF = NT / 8, STR = NT + 4, NW = sk_warps(NT);
    constexpr bool SCALE_IN_SMEM = (NT == 16) && !PROD_WARP;
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, K = ly.K;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* done = full + NBUF;
    unsigned* left = reinterpret_cast<unsigned*>(done + NBUF);   // !PROD_WARP: warps that have left the stage in buffer b
    double* Tb = smem + SK_BAR_DOUBLES;                // [NBUF][Mp][STR]
    double* mub = Tb + (size_t)NBUF * Mp * STR;        // [2][NT][K]  mubar slab of the tile (by tile parity)
    double* vbs = mub + 2 * NT * KP;                   // [2][NT][K]  vbar slab
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const size_t tile_elems = (size_t)Mp * STR;
    const unsigned tile_bytes = (unsigned)(tile_elems * sizeof(double));
    const unsigned slab_bytes = (unsigned)(NT * K * sizeof(double));
    if (threadIdx.x == 0) {
        for (int i = 0; i < NBUF; ++i) { mbar_init(&full[i], 1); mbar_init(&done[i], NW); left[i] = 0u; }
        mbar_fence_init();
    }
    __syncthreads();
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * K;
    auto tile_of = [&](int ti) { return (int64_t)blockIdx.x + (int64_t)ti * gridDim.x; };
    // stage j = (tile iteration ti, s): the B_s tile.  One lane issues.
    auto issue = [&](int j) {
        const int ti = j / K, s = j - ti * K, buf = j % NBUF;
        const int64_t tile = tile_of(ti);
        if (PROD_WARP && j >= NBUF) mbar_wait(&done[buf], (unsigned)(((j / NBUF) - 1) & 1));
        mbar_arrive_expect_tx(&full[buf], tile_bytes + (s == 0 ? 2 * slab_bytes : 0u));
        bulk_g2s(Tb + (size_t)buf * tile_elems, cb.Bk + ((size_t)s * cb.tiles_cap + tile) * tile_elems, tile_bytes, &full[buf]);
        if (s == 0) {
            bulk_g2s(mub + (size_t)(ti & 1) * NT * KP, cb.mubar + (size_t)tile * NT * K, slab_bytes, &full[buf]);
            bulk_g2s(vbs + (size_t)(ti & 1) * NT * KP, cb.vbar + (size_t)tile * NT * K, slab_bytes, &full[buf]);
            bulk_prefetch_l2(cb.A + (size_t)tile * tile_elems, tile_bytes);   // read by the epilogue K stages from now
        }
    };
    if (PROD_WARP && warp == NW) {
        if (lane == 0)
            for (int j = 0; j < total; ++j) issue(j);
        return;
    }
    if (!PROD_WARP && warp == 0 && lane == 0)
        for (int j = 0; j < NBUF && j < total; ++j) issue(j);

    const int nmy = my_block_count<NW>(warp, nb16);
    auto seg_of = [&](int k, int r) {
        const int b = snake_block<NW>(r, warp, nb16);
        return Seg{ly.W_Lq + (size_t)k * Mp * Mp, 2 * b, 0};
    };
    WFrag wf;
    if (nmy > 0) wfrag_load(wf, seg_of(0, 0), C4, 0, lane);
    // q_mu fragments of this warp's row blocks: the same for every tile, so they are read once (when the warp has
    // few enough row blocks for them to stay in registers)
    constexpr bool HOIST_WM = NBW <= 2;
    double wm[HOIST_WM ? NBW : 1][2][KP / 4];
#pragma unroll
    for (int r = 0; r < (HOIST_WM ? NBW : 0); ++r) {
        const int b = snake_block<NW>(r, warp, nb16);
#pragma unroll
        for (int mf = 0; mf < 2; ++mf)
#pragma unroll
            for (int kb = 0; kb < KP / 4; ++kb)
                wm[r][mf][kb] = (r < nmy && b >= 0) ? __ldg(ly.W_m + ((size_t)(2 * b + mf) * (KP / 4) + kb) * 32 + lane) : 0.0;
    }
    double acc[NBW][2][NF][2];
    PH_DECL
    for (int j = 0; j < total; ++j) {
        const int ti = j / K, s = j - ti * K, buf = j % NBUF;
        const double* T = Tb + (size_t)buf * tile_elems;
        const double* mb = mub + (size_t)(ti & 1) * NT * KP;
        const double* vb = vbs + (size_t)(ti & 1) * NT * KP;
        PH_MARK(5)
        mbar_wait(&full[buf], (unsigned)((j / NBUF) & 1));
        PH_MARK(0)
        if (s == 0) {
#pragma unroll
            for (int r = 0; r < NBW; ++r) zero_acc<NF>(acc[r]);
        }
        {
            if (SCALE_IN_SMEM) {
                // 16-point tiles (16 consumer warps, 128-register cap): the column weights 2 vbar_s are applied to the
                // landed B_s stage in shared memory — one pass of Mp * NT multiplications per stage (< 0.5 % of its
                // DMMA time at M = 1024) and one consumer-wide barrier — so that the products accumulate straight into
                // `acc`: no second accumulator set, no per-block FMA epilogue, no spills.
                double* Tw = Tb + (size_t)buf * tile_elems;
                const int n = threadIdx.x % NT;   // NW * 32 threads, a multiple of NT: a thread stays on one column
                const double w2 = 2.0 * vb[n * K + s];
                for (int m = threadIdx.x / NT; m < Mp; m += (NW * 32) / NT) Tw[(size_t)m * STR + n] *= w2;
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic writes before the next bulk copy into this buffer
                named_bar_sync(1, NW * 32);
#pragma unroll
                for (int r = 0; r < NBW; ++r) {
                    if (r >= nmy) continue;
                    const int b = snake_block<NW>(r, warp, nb16);
                    const Seg nxt = (r + 1 < nmy) ? seg_of(s, r + 1) : seg_of(s + 1 < K ? s + 1 : 0, 0);
                    wgemm_seg<NT, TRI_LOWER>(seg_of(s, r), (b + 1) * 4, C4, T, acc[r], lane, wf, nxt);   // lower triangular
                }
            } else {
            // column weights 2 vbar_s of this lane's columns; applied once per (block, s) to the finished product
            // Lq_s B_s instead of to every B fragment (keeps DMUL out of the DMMA loop)
            double sc[NF][2];
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) {
                sc[nf][0] = 2.0 * vb[(nf * 8 + 2 * t) * K + s];
                sc[nf][1] = 2.0 * vb[(nf * 8 + 2 * t + 1) * K + s];
            }
#pragma unroll
            for (int r = 0; r < NBW; ++r) {
                if (r >= nmy) continue;
                const int b = snake_block<NW>(r, warp, nb16);
                double ck[2][NF][2];
                zero_acc<NF>(ck);
                const Seg nxt = (r + 1 < nmy) ? seg_of(s, r + 1) : seg_of(s + 1 < K ? s + 1 : 0, 0);
                PH_MARK(1)
                wgemm_seg<NT, TRI_LOWER>(seg_of(s, r), (b + 1) * 4, C4, T, ck, lane, wf, nxt);   // lower triangular
                PH_MARK(2)
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) {
                        acc[r][mf][nf][0] = fma(ck[mf][nf][0], sc[nf][0], acc[r][mf][nf][0]);
                        acc[r][mf][nf][1] = fma(ck[mf][nf][1], sc[nf][1], acc[r][mf][nf][1]);
                    }
                PH_MARK(3)
            }
            }
        }
        if (s == K - 1) {   // tile epilogue: + q_mu mubar^T - 2 A diag(sum_k vbar_k), written over A (tile-major, in place)
            double* Aout = cb.A + (size_t)tile_of(ti) * tile_elems;
            double vs[NF][2];
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) {
                vs[nf][0] = vs[nf][1] = 0.0;
                for (int kk = 0; kk < K; ++kk) {
                    vs[nf][0] += vb[(nf * 8 + 2 * t) * K + kk];
                    vs[nf][1] += vb[(nf * 8 + 2 * t + 1) * K + kk];
                }
            }
            // B fragments of mubar^T [KP x NT]: element (k = kb*4 + t, n = nf*8 + g)
            double bm[KP / 4][NF];
#pragma unroll
            for (int kb = 0; kb < KP / 4; ++kb)
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) bm[kb][nf] = (kb * 4 + t < K) ? mb[(nf * 8 + g) * K + kb * 4 + t] : 0.0;
#pragma unroll
            for (int r = 0; r < NBW; ++r) {
                const int b = snake_block<NW>(r, warp, nb16);
                if (r >= nmy || b < 0) continue;
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                    for (int kb = 0; kb < KP / 4; ++kb) {
                        const double a = HOIST_WM ? wm[HOIST_WM ? r : 0][mf][kb]
                                                  : __ldg(ly.W_m + ((size_t)(2 * b + mf) * (KP / 4) + kb) * 32 + lane);
#pragma unroll
                        for (int nf = 0; nf < NF; ++nf) dmma(acc[r][mf][nf], a, bm[kb][nf]);
                    }
                // this lane's A values (prefetched into L2 when the tile's first stage was issued): all loads of the
                // block first, then the in-place stores
                double2 av[2][NF];
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf)
                        av[mf][nf] = __ldcs(reinterpret_cast<const double2*>(Aout + (size_t)(b * 16 + mf * 8 + g) * STR + nf * 8 + 2 * t));
#pragma unroll
                for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) {
                        const size_t off = (size_t)(b * 16 + mf * 8 + g) * STR + nf * 8 + 2 * t;
                        *reinterpret_cast<double2*>(Aout + off) =
                            make_double2(acc[r][mf][nf][0] - 2.0 * vs[nf][0] * av[mf][nf].x, acc[r][mf][nf][1] - 2.0 * vs[nf][1] * av[mf][nf].y);
                    }
            }
        }
        PH_MARK(4)
        __syncwarp();
        if (lane == 0) {
            if (PROD_WARP) {
                mbar_arrive(&done[buf]);
            } else {
                // no producer warp: the LAST warp to leave the stage refills its buffer with stage j + NBUF at once —
                // nobody waits for the stragglers (a rotating duty that waited on `done` cost 3.3 % of the kernel)
                __threadfence_block();
                const unsigned before = atomicAdd(&left[buf], 1u);
                if (before == NW - 1) {
                    left[buf] = 0u;
                    __threadfence_block();
                    if (j + NBUF < total) issue(j + NBUF);
                }
            }
        }
    }
    PH_STORE(1)
}

// ==================================================================================================
// cond_bwd_b :  Kuf_bar = L^-T Abar ; E = Kuf_bar .* Kuf ; per-row sums of E [1, xs_d, xs_d^2]
// Two CTAs per SM (16 warps), like cond_fwd_a: the per-fragment epilogue regenerates Kuf (exp) and is scalar FP64
// work that overlaps best with many DMMA warps.  The Abar tile arrives by one cp.async.bulk on an mbarrier.
// ==================================================================================================
template <int NT>
__global__ void __launch_bounds__(sk_warps(NT) * 32) cond_bwd_b_kernel(LayerDev ly, ChunkBuffers cb, int ntiles,
                                                                      double* esum_part, int pf_next) {
    constexpr int NF = NT / 8, STR = NT + 4, NW = sk_warps(NT);
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, Dp = ly.Dp, D = ly.D, XSTR = xs_stride(Dp), E = 1 + 2 * Dp;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // Abar tile landed (bulk copy) AND X rows staged: 2 arrivals
    uint64_t* tfree = full + 1;                            // every warp has finished multiplying from T: 8 arrivals
    double* T = smem + SK_BAR_DOUBLES;
    double* Xsb = T + (size_t)Mp * STR;                    // [2]{[NT][XSTR], [NT], [NT][FS]} scaled X rows / features, by tile parity
    const int FB = esum_feature_blocks(D), FS = 8 * FB + 2;   // feature row stride == 2 mod 8: conflict-free B fragments
    const int xs_elems = NT * XSTR + NT + NT * FS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const size_t tile_elems = (size_t)Mp * STR;
    const unsigned tile_bytes = (unsigned)(tile_elems * sizeof(double));
    __shared__ double etab[64];
    if (threadIdx.x < 64) etab[threadIdx.x] = d_exp_tab64[threadIdx.x];   // visible after the barrier-init sync below
    double* my_part = esum_part + (size_t)blockIdx.x * Mp * E;
    const int nmy = my_block_count<NW>(warp, nb16);
    auto seg_of = [&](int i) { const int b = snake_block<NW>(i, warp, nb16); return Seg{ly.W_LinvT, 2 * b, 4 * b}; };
    WFrag wf;
    if (nmy > 0) wfrag_load(wf, seg_of(0), C4, seg_of(0).kb0, lane);
    if (threadIdx.x == 0) { mbar_init(full, 2); mbar_init(tfree, NW); mbar_fence_init(); }
    __syncthreads();
    // The per-fragment epilogue does not read T, so the NEXT tile's bulk copy is issued as soon as every warp has left
    // the multiply phase (tfree) and flies while the epilogues run; X rows are staged into the other Xs buffer by
    // warp 1 at the same point.  No CTA-wide barrier in the loop.
    auto load_tile = [&](int tile) {   // thread 0
        mbar_arrive_expect_tx(full, tile_bytes);
        bulk_g2s(T, cb.A + (size_t)tile * tile_elems, tile_bytes, full);
        // the Kuf tile cond_fwd_a kept: on its way into L2 while the warps multiply; the epilogues read it from there
        if (cb.Kuf) bulk_prefetch_l2(cb.Kuf + (size_t)tile * tile_elems, tile_bytes);
        // T is single-buffered (two CTAs per SM), so the NEXT Abar tile can only be copied once every warp has left the
        // multiply phase: have it waiting in L2 by then instead of in HBM
        if (pf_next && tile + (int)gridDim.x < ntiles) bulk_prefetch_l2(cb.A + (size_t)(tile + gridDim.x) * tile_elems, tile_bytes);
    };
    auto stage_x = [&](int tile, int it) {   // warp 1
        double* Xs = Xsb + (size_t)(it & 1) * xs_elems;
        stage_x_warp<NT>(ly, cb, (int64_t)tile * NT, Xs, Xs + NT * XSTR, lane);
        // features of the E-sums, Phi[n][f] = {1, xs_d, xs_d^2}: right operand of the small DMMA product below
        double* Ph = Xs + NT * XSTR + NT;
        for (int idx = lane; idx < NT * 8 * FB; idx += 32) {
            const int n = idx / (8 * FB), f = idx % (8 * FB);
            double v = 0.0;
            if (f == 0) v = 1.0;
            else if (f <= D) v = Xs[n * XSTR + f - 1];
            else if (f <= 2 * D) { const double x = Xs[n * XSTR + f - 1 - D]; v = x * x; }
            Ph[n * FS + f] = v;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(full);
    };
    if ((int)blockIdx.x < ntiles) {
        if (threadIdx.x == 0) load_tile(blockIdx.x);
        if (warp == 1) stage_x(blockIdx.x, 0);
    }
    int it = 0;
    PH_DECL
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const double* Xs = Xsb + (size_t)(it & 1) * xs_elems;
        const double* xs2 = Xs + NT * XSTR;
        const unsigned phase = (unsigned)(it & 1);
        PH_MARK(5)
        mbar_wait(full, phase);
        PH_MARK(0)
        if (nmy == 0) { __syncwarp(); if (lane == 0) mbar_arrive(tfree); }
        for (int i = 0; i < nmy; ++i) {
            const int b = snake_block<NW>(i, warp, nb16);
            double acc[2][NF][2];
            zero_acc<NF>(acc);
            wgemm_seg<NT, TRI_UPPER>(seg_of(i), C4, C4, T, acc, lane, wf, seg_of(i + 1 < nmy ? i + 1 : 0));   // upper triangular
            PH_MARK(1)
            if (i == nmy - 1) { __syncwarp(); if (lane == 0) mbar_arrive(tfree); }   // this warp is done with T
            const double* Ph = xs2 + NT;
#pragma unroll
            for (int mf = 0; mf < 2; ++mf) {
                double kv[NF][2];
                if (cb.Kuf) {   // (CTA-uniform) the values cond_fwd_a generated, in C-fragment order
                    const double* kt = cb.Kuf + (size_t)tile * tile_elems + (size_t)((2 * b + mf) * 8 + g) * STR + 2 * t;
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) {
                        const double2 q = __ldcs(reinterpret_cast<const double2*>(kt + nf * 8));
                        kv[nf][0] = q.x; kv[nf][1] = q.y;
                    }
                } else {
                    gen_kuf_block<NT>(ly, 2 * b + mf, Xs, xs2, etab, kv, lane);
                }
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                    for (int e = 0; e < 2; ++e) acc[mf][nf][e] *= kv[nf][e];   // E = Kuf_bar .* Kuf
                PH_MARK(2)
                // sum_n E[i][n] Phi[n][f] on DMMA: the C fragment of E is read as A fragments over the point subsets
                // {nf*8 + 2t + e : t = 0..3}, the matching rows of Phi as B fragments.  (Per element this replaces 2 + 3 D
                // scalar FP64 instructions, each of which costs the tensor pipe several DMMA issue slots, by 1/4 DMMA.)
                double* p = my_part + (size_t)(b * 16 + mf * 8 + g) * E;
                for (int fb = 0; fb < FB; ++fb) {
                    double R[2] = {0.0, 0.0};
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf)
#pragma unroll
                        for (int e = 0; e < 2; ++e) dmma(R, acc[mf][nf][e], Ph[(nf * 8 + 2 * t + e) * FS + fb * 8 + g]);
                    // fire-and-forget reductions: each address has exactly ONE writer (this lane, this CTA's private slot),
                    // so the accumulation order is fixed and the result deterministic
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int f = fb * 8 + 2 * t + j;
                        if (f == 0) atomicAdd(p, R[j]);
                        else if (f <= D) atomicAdd(p + f, R[j]);
                        else if (f <= 2 * D) atomicAdd(p + 1 + Dp + (f - 1 - D), R[j]);
                    }
                }
                PH_MARK(3)
            }
        }
        const int next = tile + gridDim.x;
        if (warp <= 1 && next < ntiles) {
            // tfree(it) also tells that every warp has finished the epilogue of tile it - 1 (it comes before its
            // arrival in program order), i.e. nobody still reads the Xs buffer about to be refilled.
            // (Fetching the next tile BEFORE these two warps' own last epilogue was measured: 5.34 vs 5.29 ms.)
            mbar_wait(tfree, phase);
            if (threadIdx.x == 0) load_tile(next);
            if (warp == 1) stage_x(next, it + 1);
            PH_MARK(4)
        }
    }
    PH_STORE(2)
}

// ==================================================================================================
// cond_bwd_b, ring form (32-point tiles, Kuf tiles kept by cond_fwd_a):  ONE persistent CTA per SM like cond_fwd_b.
// With the Kuf values read back instead of generated, the per-block epilogue is 8 loads (issued BEFORE the block's
// multiply: a one-CTA kernel has the registers), 16 multiplications, 16 DMMAs and 4 reductions — no exponentials, no
// dependent scalar chain — so the kernel no longer needs a second CTA to hide it, and the single-buffered tile of the
// two-CTA form (next copy only after every warp has left the multiply phase) becomes a two-deep ring fed by warp 8,
// which also stages the X rows / E-sum features of the tile.
// ==================================================================================================
template <int NT, int NBUF>
__global__ void __launch_bounds__(SK_WARPS * 32 + 32, 1) cond_bwd_b_ring_kernel(LayerDev ly, ChunkBuffers cb, int ntiles,
                                                                              double* esum_part) {
    constexpr int NF = NT / 8, STR = NT + 4, NW = SK_WARPS;
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, Dp = ly.Dp, D = ly.D, XSTR = xs_stride(Dp), E = 1 + 2 * Dp;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // Abar tile landed and X rows / features staged
    uint64_t* done = full + NBUF;                          // every consumer warp has finished the tile
    double* Tb = smem + SK_BAR_DOUBLES;                    // [NBUF][Mp][STR]
    const size_t tile_elems = (size_t)Mp * STR;
    double* Xsb = Tb + (size_t)NBUF * tile_elems;          // [NBUF]{[NT][XSTR], [NT], [NT][FS]}
    const int FB = esum_feature_blocks(D), FS = 8 * FB + 2;
    const int xs_elems = NT * XSTR + NT + NT * FS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const unsigned tile_bytes = (unsigned)(tile_elems * sizeof(double));
    if (threadIdx.x == 0) {
        for (int i = 0; i < NBUF; ++i) { mbar_init(&full[i], 1); mbar_init(&done[i], NW); }
        mbar_fence_init();
    }
    __syncthreads();
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto tile_of = [&](int i) { return (int64_t)blockIdx.x + (int64_t)i * gridDim.x; };

    if (warp == NW) {   // ---- producer: X rows, features, Abar tile ----
        for (int i = 0; i < my_tiles; ++i) {
            const int buf = i % NBUF;
            const int64_t tile = tile_of(i);
            if (i >= NBUF) mbar_wait(&done[buf], (unsigned)(((i / NBUF) - 1) & 1));
            double* Xs = Xsb + (size_t)buf * xs_elems;
            stage_x_warp<NT>(ly, cb, tile * NT, Xs, Xs + NT * XSTR, lane);
            double* Ph = Xs + NT * XSTR + NT;   // Phi[n][f] = {1, xs_d, xs_d^2}: right operand of the E-sum product
            for (int idx = lane; idx < NT * 8 * FB; idx += 32) {
                const int n = idx / (8 * FB), f = idx % (8 * FB);
                double v = 0.0;
                if (f == 0) v = 1.0;
                else if (f <= D) v = Xs[n * XSTR + f - 1];
                else if (f <= 2 * D) { const double x = Xs[n * XSTR + f - 1 - D]; v = x * x; }
                Ph[n * FS + f] = v;
            }
            __syncwarp();
            if (lane == 0) {
                bulk_prefetch_l2(cb.Kuf + (size_t)tile * tile_elems, tile_bytes);
                mbar_arrive_expect_tx(&full[buf], tile_bytes);
                bulk_g2s(Tb + (size_t)buf * tile_elems, cb.A + (size_t)tile * tile_elems, tile_bytes, &full[buf]);
            }
        }
        return;
    }
    // ---- consumers ----
    double* my_part = esum_part + (size_t)blockIdx.x * Mp * E;
    const int nmy = my_block_count<NW>(warp, nb16);
    auto seg_of = [&](int i) { const int b = snake_block<NW>(i, warp, nb16); return Seg{ly.W_LinvT, 2 * b, 4 * b}; };
    WFrag wf;
    if (nmy > 0) wfrag_load(wf, seg_of(0), C4, seg_of(0).kb0, lane);
    for (int i = 0; i < my_tiles; ++i) {
        const int buf = i % NBUF;
        const double* T = Tb + (size_t)buf * tile_elems;
        const double* Ph = Xsb + (size_t)buf * xs_elems + NT * XSTR + NT;
        const double* ktile = cb.Kuf + (size_t)tile_of(i) * tile_elems;
        mbar_wait(&full[buf], (unsigned)((i / NBUF) & 1));
        for (int r = 0; r < nmy; ++r) {
            const int b = snake_block<NW>(r, warp, nb16);
            // the block's Kuf values (C-fragment order), in flight while the block is multiplied
            double2 kq[2][NF];
#pragma unroll
            for (int mf = 0; mf < 2; ++mf)
#pragma unroll
                for (int nf = 0; nf < NF; ++nf)
                    kq[mf][nf] = __ldcs(reinterpret_cast<const double2*>(ktile + (size_t)((2 * b + mf) * 8 + g) * STR + nf * 8 + 2 * t));
            double acc[2][NF][2];
            zero_acc<NF>(acc);
            wgemm_seg<NT, TRI_UPPER>(seg_of(r), C4, C4, T, acc, lane, wf, seg_of(r + 1 < nmy ? r + 1 : 0));   // upper triangular
#pragma unroll
            for (int mf = 0; mf < 2; ++mf) {
#pragma unroll
                for (int nf = 0; nf < NF; ++nf) { acc[mf][nf][0] *= kq[mf][nf].x; acc[mf][nf][1] *= kq[mf][nf].y; }   // E = Kuf_bar .* Kuf
                double* p = my_part + (size_t)(b * 16 + mf * 8 + g) * E;
                for (int fb = 0; fb < FB; ++fb) {   // sum_n E[i][n] Phi[n][f] on DMMA (see cond_bwd_b_kernel), two chains
                    double R0[2] = {0.0, 0.0}, R1[2] = {0.0, 0.0};
#pragma unroll
                    for (int nf = 0; nf < NF; ++nf) {
                        dmma(R0, acc[mf][nf][0], Ph[(nf * 8 + 2 * t) * FS + fb * 8 + g]);
                        dmma(R1, acc[mf][nf][1], Ph[(nf * 8 + 2 * t + 1) * FS + fb * 8 + g]);
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) {   // one writer per address (this lane, this CTA's slot): deterministic
                        const int f = fb * 8 + 2 * t + j;
                        const double v = R0[j] + R1[j];
                        if (f == 0) atomicAdd(p, v);
                        else if (f <= D) atomicAdd(p + f, v);
                        else if (f <= 2 * D) atomicAdd(p + 1 + Dp + (f - 1 - D), v);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[buf]);
    }
}

// ==================================================================================================
// host side
// ==================================================================================================
int stream_max_parts(const Launch& ln) { return ln.num_sms * 4; }

// Tile width of a layer's tile-major workspace (all four kernels and the SYRK share it) and the ring depth:
//   NT = 32, 2 buffers  while two [Mp x 36] tiles fit next to the per-kernel extras;
//   NT = 16, 2 buffers  up to Mp = 672;  NT = 16, 1 buffer beyond (no load / multiply overlap inside a CTA).
static size_t extras_bytes(int Mp, int Dp, int K, int nt, int nbuf = 2) {
    const size_t fa = (size_t)(nt * xs_stride(Dp) + nt) * 8;
    const size_t fb = (size_t)nbuf * (sk_warps(nt) + SK_WARPS) * K * nt * 8;
    const size_t ba = (size_t)4 * nt * KP * 8;
    const size_t bb = (size_t)2 * (nt * xs_stride(Dp) + nt + nt * (8 * ((1 + 2 * Dp + 7) / 8) + 2)) * 8;
    size_t m = fa;
    if (fb > m) m = fb;
    if (ba > m) m = ba;
    if (bb > m) m = bb;
    (void)Mp;
    return m + SK_BAR_DOUBLES * 8;
}
int layer_tile_width(int Mp, int Dp, int K) {
    const size_t cap = 227 * 1024;
    if (getenv("MGP_TILE_W16")) return 16;   // (timing experiment: 16-point tiles / 16 consumer warps at any M)
    return (2 * (size_t)Mp * 36 * 8 + extras_bytes(Mp, Dp, K, 32) <= cap) ? 32 : 16;
}
static int ring_depth(int Mp, int Dp, int K, int nt) {
    return (2 * (size_t)Mp * (nt + 4) * 8 + extras_bytes(Mp, Dp, K, nt, 2) <= (size_t)227 * 1024) ? 2 : 1;
}

template <typename KernelT>
static int persistent_grid(KernelT kernel, int threads, size_t smem, int ntiles, int cap, const Launch& ln) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    (void)threads;
    int grid = ln.num_sms;
    if (grid > ntiles) grid = ntiles;
    if (cap > 0 && grid > cap) grid = cap;
    return grid < 1 ? 1 : grid;
}
// grid = SMs x resident CTAs (occupancy query) for the kernels that run several CTAs per SM
template <typename KernelT>
static int occupancy_grid(KernelT kernel, int threads, size_t smem, int ntiles, int cap, const Launch& ln) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
    if (occ < 1) occ = 1;
    int grid = ln.num_sms * occ;
    if (grid > ntiles) grid = ntiles;
    if (cap > 0 && grid > cap) grid = cap;
    return grid < 1 ? 1 : grid;
}

void cond_fwd_a(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln) {
    const int NT = cb.tw;
    const size_t smem = ((size_t)ly.Mp * (NT + 4) + NT * xs_stride(ly.Dp) + NT) * sizeof(double);
    const int ntiles = (int)((cb.n + NT - 1) / NT);
    const int threads = sk_warps(NT) * 32;
    // measured at config #4: software-pipelined form 5.94 ms (6.14 before its Z fragments moved to shared memory), the
    // barrier-phased form with three CTAs per SM 5.43 ms -> the pipelined form stays selectable only
    const bool pipe = getenv("MGP_FWD_A_PIPE") != nullptr;   // (A/B timing switch; tests cover both forms)
    if (NT == 32 && pipe) {   // software-pipelined form: one CTA per SM, two tile buffers (they fit whenever NT = 32)
        size_t psmem = ((size_t)2 * ly.Mp * (NT + 4) + 2 * (NT * xs_stride(ly.Dp) + NT)) * sizeof(double);
        const size_t zbytes = (size_t)ly.Mp * ly.Dp * sizeof(double);
        const int zs_in_smem = psmem + zbytes <= (size_t)226 * 1024;
        if (zs_in_smem) psmem += zbytes;
        const int grid = persistent_grid(cond_fwd_a_pipe_kernel<32>, SK_CTHREADS + 32, psmem, ntiles, 0, ln);
        cond_fwd_a_pipe_kernel<32><<<grid, SK_CTHREADS + 32, psmem, ln.stream>>>(ly, cb, ntiles, zs_in_smem);
        ln.tick();
        return;
    }
    auto launch = [&](auto kernel) {
        const int grid = occupancy_grid(kernel, threads, smem, ntiles, 0, ln);
        kernel<<<grid, threads, smem, ln.stream>>>(ly, cb, ntiles);
        ln.tick();
    };
    if (NT == 32) launch(cond_fwd_a_kernel<32>); else launch(cond_fwd_a_kernel<16>);
}

// bytes of dynamic shared memory of the fused forward kernel, or 0 when the layer does not qualify (32-point tiles only)
static size_t fused_fwd_smem(const LayerDev& ly, int NT) {
    // OFF by default: measured SLOWER than the two kernels (23.7 vs 23.1 ms at config #4) — the Kuf
    // generation's scalar FP64 instructions cost the consumers' DMMA stream ~1.5 ms wherever they run
    if (NT != 32 || getenv("MGP_FUSED_FWD") == nullptr) return 0;
    const size_t bytes = ((size_t)SK_BAR_DOUBLES + (size_t)2 * ly.Mp * (NT + 4) + (size_t)4 * SK_WARPS * ly.K * NT +
                          (size_t)2 * (NT * xs_stride(ly.Dp) + NT)) * sizeof(double);
    return bytes <= (size_t)227 * 1024 - 1024 ? bytes : 0;
}
bool cond_fwd_is_fused(const LayerDev& ly, const ChunkBuffers& cb) { return fused_fwd_smem(ly, cb.tw) != 0; }

void cond_fwd_fused(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln) {
    const int NT = cb.tw;
    const size_t smem = fused_fwd_smem(ly, NT);
    const int threads = (SK_WARPS + FU_GEN_WARPS) * 32;
    const int ntiles = (int)((cb.n + NT - 1) / NT);
    const int grid = persistent_grid(cond_fwd_fused_kernel<32>, threads, smem, ntiles, 0, ln);
    static const int dbg = getenv("MGP_FUSED_DBG") ? atoi(getenv("MGP_FUSED_DBG")) : 0;
    cond_fwd_fused_kernel<32><<<grid, threads, smem, ln.stream>>>(ly, cb, ntiles, dbg);
    ln.tick();
}

void cond_fwd_b(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln) {
    const int NT = cb.tw, nbuf = ring_depth(ly.Mp, ly.Dp, ly.K, NT);
    const int threads = sk_warps(NT) * 32 + 32;
    const size_t smem = ((size_t)SK_BAR_DOUBLES + (size_t)nbuf * ly.Mp * (NT + 4) + (size_t)nbuf * (sk_warps(NT) + SK_WARPS) * ly.K * NT) * sizeof(double);
    const int ntiles = (int)((cb.n + NT - 1) / NT);
    auto launch = [&](auto kernel) {
        const int grid = persistent_grid(kernel, threads, smem, ntiles, 0, ln);
        kernel<<<grid, threads, smem, ln.stream>>>(ly, cb, ntiles);
        ln.tick();
    };
    if (NT == 32) launch(cond_fwd_b_kernel<32, 2>);
    else if (nbuf == 2) launch(cond_fwd_b_kernel<16, 2>);
    else launch(cond_fwd_b_kernel<16, 1>);
}

void cond_bwd_a(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln) {
    const int NT = cb.tw, nbuf = ring_depth(ly.Mp, ly.Dp, ly.K, NT);
    const int nw = sk_warps(NT);
    const int nbw = (ly.Mp / 16 + nw - 1) / nw;   // 16-row blocks per warp
    const size_t smem = ((size_t)SK_BAR_DOUBLES + (size_t)nbuf * ly.Mp * (NT + 4) + (size_t)4 * NT * KP) * sizeof(double);
    const int ntiles = (int)((cb.n + NT - 1) / NT);
    auto launch = [&](auto kernel, int threads) {
        const int grid = persistent_grid(kernel, threads, smem, ntiles, 0, ln);
        kernel<<<grid, threads, smem, ln.stream>>>(ly, cb, ntiles);
        ln.tick();
    };
    // accumulators of all row blocks live in registers: NBW * 2 * NF * 2 doubles per lane
    if (NT == 32) {
        if (nbw <= 1) launch(cond_bwd_a_kernel<32, 2, 1, true>, SK_CTHREADS + 32);
        else launch(cond_bwd_a_kernel<32, 2, 2, false>, SK_CTHREADS);
    } else if (nbuf == 2) {   // 16 consumer warps, no producer warp (128-register cap): Mp <= ~600 -> at most 3 blocks per warp
        if (nbw <= 2) launch(cond_bwd_a_kernel<16, 2, 2, false>, nw * 32);
        else launch(cond_bwd_a_kernel<16, 2, 4, false>, nw * 32);
    } else {                  // up to Mp = 1312: 82 row blocks over 16 warps
        if (nbw <= 4) launch(cond_bwd_a_kernel<16, 1, 4, false>, nw * 32);
        else launch(cond_bwd_a_kernel<16, 1, 6, false>, nw * 32);
    }
}

void cond_bwd_b(const LayerDev& ly, const ChunkBuffers& cb, double* esum_part, int nparts_cap, int* nparts,
                const Launch& ln) {
    const int NT = cb.tw;
    const size_t smem = ((size_t)SK_BAR_DOUBLES + (size_t)ly.Mp * (NT + 4) +
                         2 * (NT * xs_stride(ly.Dp) + NT + NT * (8 * esum_feature_blocks(ly.D) + 2))) * sizeof(double);
    const int ntiles = (int)((cb.n + NT - 1) / NT);
    // measured at config #4: ring form 5.47 ms, two CTAs per SM 5.28 ms -> the ring form stays selectable only
    const bool ring = getenv("MGP_BWD_B_RING") != nullptr;   // (A/B timing switch; tests cover both forms)
    if (NT == 32 && cb.Kuf && ring) {   // ring form: one CTA per SM, two-deep tile ring
        const size_t rsmem = ((size_t)SK_BAR_DOUBLES + (size_t)2 * ly.Mp * (NT + 4) +
                              2 * (NT * xs_stride(ly.Dp) + NT + NT * (8 * esum_feature_blocks(ly.D) + 2))) * sizeof(double);
        const int grid = persistent_grid(cond_bwd_b_ring_kernel<32, 2>, SK_CTHREADS + 32, rsmem, ntiles, nparts_cap, ln);
        cond_bwd_b_ring_kernel<32, 2><<<grid, SK_CTHREADS + 32, rsmem, ln.stream>>>(ly, cb, ntiles, esum_part);
        ln.tick();
        if (grid > *nparts) *nparts = grid;
        return;
    }
    auto launch = [&](auto kernel) {
        const int grid = occupancy_grid(kernel, sk_warps(NT) * 32, smem, ntiles, nparts_cap, ln);
        const int pf_next = getenv("MGP_BWD_B_NO_PF") ? 0 : 1;   // (A/B timing switch)
        kernel<<<grid, sk_warps(NT) * 32, smem, ln.stream>>>(ly, cb, ntiles, esum_part, pf_next);
        ln.tick();
        if (grid > *nparts) *nparts = grid;
    };
    if (NT == 32) launch(cond_bwd_b_kernel<32>); else launch(cond_bwd_b_kernel<16>);
}

}  // namespace mgp

#ifdef MGP_PHASE_TIMERS
// (temporary profiling builds only) copies g_phase to the host: out[4][160][17][8]
extern "C" int mgp_debug_phase_dump(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, mgp::g_phase, sizeof(long long) * 4 * 160 * 17 * 8);
}
#endif
