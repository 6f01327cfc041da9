// Device helpers shared by the N-streaming kernels (stream_kernels.cu: the forms that run; stream_kernels_alt.cu: the
// alternative forms that were built, measured slower and stay selectable): the table exponential, the warp-level
// triangular GEMM loop with its software-pipelined left-operand fragments, Kuf generation, launch-grid helpers.
#pragma once
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "exp_tab.h"
#include "kernels.h"

namespace mgp {

// exp2_tab(y) = exp(y ln2 / 64) from a 64-entry table of 2^(j/64) and a degree-5 polynomial on |rr| <= 1/2 (rr in
// units of ln2/64): 9 FP64 instructions instead of libdevice's 16-18 for exp, at most 1.1 ulp from expl
// (tools/exp_tab_check.c, tests/test_host_logic.py).  The argument arrives ALREADY in units of ln2/64: the left operand
// of the z.x contraction (Zs_fm, prep_z_kernel) is pre-scaled by 64/ln2, so the DMMA result needs no multiplication for
// the range reduction, and k = round(y), rr = y - k are two exact additions.  Every scalar FP64 instruction here matters
// out of proportion to its pipe time: measured in the fused forward kernel, a warp-wide scalar
// FP64 instruction issued beside a saturated DMMA stream costs that sub-partition ~10 clocks, not 2.  The underflow test
// runs on the integer pipe (sign-and-exponent word of y).  `tab` is the shared-memory copy of d_exp_tab64.
static __device__ const double d_exp_tab64[64] = {EXP_TAB64_VALUES};
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

// ---- launch-grid helpers (host) ---------------------------------------------------------------------------------
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

}  // namespace mgp
