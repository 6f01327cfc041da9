// S_k += A diag(vbar_k) A^T : the only large reduction over points in the backward pass.
//
// With S_k and A mubar in hand, every M x M gradient of the layer follows from replicated O(M^3) algebra
// (precompute.cu::finish_layer):   dLq_k = tril(2 S_k Lq_k),   Abar A^T = sum_k Q_k S_k + q_mu (A mubar)^T,
// Lbar = -tril(L^-T Abar A^T)  — i.e. the N-reductions  tril(A Bbar_k^T)  and  tril(Kuf_bar A^T)  of SURVEY.md
// Appendix B are folded into K weighted Gram matrices of the materialised A.
//
// Output-stationary, warp-specialised DMMA kernel.  One CTA per SM; a CTA owns (64x64 tile pair I >= J, group of 4
// components, range of points) and keeps its accumulators in registers for the whole range.  32-point slabs of the A
// rows (cp.async.bulk row copies into padded rows => conflict-free fragment loads) and the vbar / mubar slabs stream
// through a 4-stage shared-memory ring, each stage completing on its "full" mbarrier.  The 8 warps (component, half)
// wait on "full", multiply, and release the stage on its "empty" mbarrier; the producer duty rotates: at stage s,
// warp s % 8 first refills the ring two stages ahead.  There is no CTA-wide barrier in the loop, so the DMMA pipe
// never drains at a stage boundary.  (A ninth, dedicated producer warp would cap the kernel at 168 registers per
// thread — three warps on one SM sub-partition — and spill the 32x64 accumulator tile.)
//
// Off-diagonal pairs: warp tile 32x64 (32 fragments).  Diagonal pairs need only fragments with col <= row (36 of
// 64); they are dealt 18 + 18 to the two halves, and the host plan gives diagonal pairs proportionally longer point
// ranges (fewer splits) so that every CTA finishes at the same time.  Each CTA accumulates into its own slot of
// `part` (no atomics; deterministic); reduce_partials sums the slots.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace mgp {

// KC = points per stage = the layer's tile width (32, or 16 for large M); row stride KC + 4 == 4 mod 16
constexpr int SY_CONSUMERS = 8;
constexpr int SY_THREADS = SY_CONSUMERS * 32;
constexpr int SY_NSTAGE = 4;
constexpr int SY_AHEAD = 2;         // stages in flight beyond the current one (leaves one stage of slack for warp skew)
template <int KC>
__host__ __device__ constexpr int sy_stage() { return 2 * 64 * (KC + 4) + 2 * KC * KP; }  // doubles: AI, AJ, vbar[KC][K], mubar[KC][K]

// A mubar needs every 64-row block of A exactly once.  It rides on OFF-diagonal pairs (the diagonal ones already are
// the slowest CTAs of the integer split): block b < nb - 1 is the J operand of pair (b + 1, b); block nb - 1 is the I
// operand of pair (nb - 1, 0); a single-block matrix uses its only pair.  0: none, 1: rows of I, 2: rows of J.
__host__ __device__ inline int syrk_mraw_role(int I, int J, int kbase, int nb) {
    if (kbase != 0) return 0;
    if (nb == 1) return 1;
    if (nb == 2) return (I == 1 && J == 0) ? 3 : 0;   // both blocks ride on the only off-diagonal pair
    if (I == J + 1) return 2;
    if (I == nb - 1 && J == 0) return 1;
    return 0;
}

// fragment (mi, ni) of the 64x64 tile (8x8 fragments) owned by a warp of the given mode
//   0 / 1 : off-diagonal pair, rows 0-31 / 32-63, all columns
//   2 / 3 : diagonal pair, the two balanced halves of the 36 fragments with ni <= mi
template <int MODE>
__host__ __device__ constexpr bool frag_on(int mi, int ni) {
    return MODE == 0   ? (mi < 4)
           : MODE == 1 ? (mi >= 4)
           : MODE == 2 ? ((mi < 4 && ni <= mi) || (mi >= 4 && ni < 2))
                       : (mi >= 4 && ((ni == 2 || ni == 3) || (ni >= 4 && ni <= mi)));
}
template <int MODE>
__host__ __device__ constexpr bool row_on(int mi) {
    for (int ni = 0; ni < 8; ++ni)
        if (frag_on<MODE>(mi, ni)) return true;
    return false;
}
template <int MODE>
__host__ __device__ constexpr bool col_on(int ni) {
    for (int mi = 0; mi < 8; ++mi)
        if (frag_on<MODE>(mi, ni)) return true;
    return false;
}
// the weight diag(vbar_k) is folded into whichever operand has fewer fragments
template <int MODE>
__host__ __device__ constexpr bool scale_rows() { return MODE != 2; }

template <int MODE, int KC>
__device__ __forceinline__ void syrk_stage(const double* AI, const double* AJ, const double* wt, int K,
                                           double (&acc)[8][8][2]) {
    constexpr int SY_STR = KC + 4;
#pragma unroll
    for (int ks = 0; ks < KC / 4; ++ks) {
        double a[8], b[8];
        const double wv = wt[ks * 4 * K];
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
            if (row_on<MODE>(mi)) {
                a[mi] = AI[mi * 8 * SY_STR + ks * 4];
                if (scale_rows<MODE>()) a[mi] *= wv;
            }
#pragma unroll
        for (int ni = 0; ni < 8; ++ni)
            if (col_on<MODE>(ni)) {
                b[ni] = AJ[ni * 8 * SY_STR + ks * 4];
                if (!scale_rows<MODE>()) b[ni] *= wv;
            }
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni)
                if (frag_on<MODE>(mi, ni)) dmma(acc[mi][ni], a[mi], b[ni]);
    }
}

// stage s of this CTA's point range -> ring slot s % SY_NSTAGE  (one warp, all lanes)
struct SyrkSrc {
    const double *A, *vbar, *mubar;
    int Mp, rows_i, rows_j;
    unsigned bytes;
};
// A is tile-major ([tile][Mp][KC + 4], stream_kernels.cu): the 64-row block of a tile is contiguous, so a stage is
// two bulk copies of A plus the vbar / mubar slabs.  One lane issues.
template <int KC>
__device__ __forceinline__ void syrk_produce(const SyrkWork& w, const SyrkSrc& src, double* ring, uint64_t* full,
                                             uint64_t* empty, int s, int K, int mrole) {
    constexpr int SY_STR = KC + 4;
    const int st = s % SY_NSTAGE;
    double* dst = ring + (size_t)st * sy_stage<KC>();
    if (s >= SY_NSTAGE) mbar_wait(&empty[st], (unsigned)(((s / SY_NSTAGE) - 1) & 1));
    const int64_t tile = w.pbeg / KC + s;
    const double* At = src.A + (size_t)tile * src.Mp * SY_STR;
    const unsigned slab = (unsigned)(KC * K * sizeof(double));
    mbar_arrive_expect_tx(&full[st], src.bytes);
    bulk_g2s(dst, At + (size_t)w.I * 64 * SY_STR, (unsigned)(src.rows_i * SY_STR * sizeof(double)), &full[st]);
    if (src.rows_j > 0)
        bulk_g2s(dst + 64 * SY_STR, At + (size_t)w.J * 64 * SY_STR, (unsigned)(src.rows_j * SY_STR * sizeof(double)), &full[st]);
    bulk_g2s(dst + 2 * 64 * SY_STR, src.vbar + (size_t)tile * KC * K, slab, &full[st]);
    if (mrole) bulk_g2s(dst + 2 * 64 * SY_STR + KC * KP, src.mubar + (size_t)tile * KC * K, slab, &full[st]);
}

template <int MODE, int KC>
__device__ __forceinline__ void syrk_consumer(const SyrkWork& w, const SyrkSrc& src, double* ring, uint64_t* full,
                                              uint64_t* empty, int nstage, double* part, double* mraw_part, int Mp, int K,
                                              int mrole) {
    constexpr int SY_STR = KC + 4, SY_STAGE = sy_stage<KC>();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int k = w.kbase + (warp >> 1);
    const bool live = k < K;
    double acc[8][8][2];
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
        for (int ni = 0; ni < 8; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    double am_i[2] = {0.0, 0.0}, am_j[2] = {0.0, 0.0};   // A mubar of this warp's 8 rows of block I / block J

    for (int s = 0; s < nstage; ++s) {
        const int st = s % SY_NSTAGE;
        const double* cur = ring + (size_t)st * SY_STAGE;
        if ((s % SY_CONSUMERS) == warp && s + SY_AHEAD < nstage && lane == 0)
            syrk_produce<KC>(w, src, ring, full, empty, s + SY_AHEAD, K, mrole);
        mbar_wait(&full[st], (unsigned)((s / SY_NSTAGE) & 1));
        if (live) {
            const double* AI = cur + g * SY_STR + t;
            const double* AJ = (MODE >= 2 ? cur : cur + 64 * SY_STR) + g * SY_STR + t;
            syrk_stage<MODE, KC>(AI, AJ, cur + 2 * 64 * SY_STR + t * K + k, K, acc);
        }
        if (mrole) {   // warp w: rows 8w .. 8w+8 of the 64-row block; columns = components
            const double* mb = cur + 2 * 64 * SY_STR + KC * KP + t * K + g;
            if (mrole & 1) {
                const double* Aw = cur + (warp * 8 + g) * SY_STR + t;
#pragma unroll
                for (int ks = 0; ks < KC / 4; ++ks) dmma(am_i, Aw[ks * 4], g < K ? mb[ks * 4 * K] : 0.0);
            }
            if (mrole & 2) {
                const double* Aw = cur + (64 + warp * 8 + g) * SY_STR + t;
#pragma unroll
                for (int ks = 0; ks < KC / 4; ++ks) dmma(am_j, Aw[ks * 4], g < K ? mb[ks * 4 * K] : 0.0);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (mrole & 1) {
        const int row = w.I * 64 + warp * 8 + g;
        if (row < Mp) {
            double* p = mraw_part + ((size_t)w.slot * Mp + row) * KP + 2 * t;
            p[0] += am_i[0];
            p[1] += am_i[1];
        }
    }
    if (mrole & 2) {
        const int row = w.J * 64 + warp * 8 + g;
        if (row < Mp) {
            double* p = mraw_part + ((size_t)w.slot * Mp + row) * KP + 2 * t;
            p[0] += am_j[0];
            p[1] += am_j[1];
        }
    }
    if (live) {
        double* P = part + ((size_t)w.slot * K + k) * Mp * Mp;
#pragma unroll
        for (int mi = 0; mi < 8; ++mi)
#pragma unroll
            for (int ni = 0; ni < 8; ++ni)
                if (frag_on<MODE>(mi, ni)) {
                    const int row = w.I * 64 + mi * 8 + g, col = w.J * 64 + ni * 8 + 2 * t;
                    if (row < Mp && col < Mp) {
                        double* p = P + (size_t)row * Mp + col;
                        p[0] += acc[mi][ni][0];
                        p[1] += acc[mi][ni][1];
                    }
                }
    }
}

template <int KC>
__global__ void __launch_bounds__(SY_THREADS, 1) syrk_kernel(const SyrkWork* plan, const double* A, const double* vbar,
                                                             const double* mubar, double* part, double* mraw_part, int Mp,
                                                             int K) {
    constexpr int SY_STR = KC + 4, SY_STAGE = sy_stage<KC>();
    extern __shared__ __align__(16) double smem[];
    __shared__ uint64_t full[SY_NSTAGE], empty[SY_NSTAGE];
    const SyrkWork w = plan[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nstage = (int)((w.pend - w.pbeg) / KC);
    if (nstage <= 0) return;
    const bool diag = (w.I == w.J);
    const int mrole = syrk_mraw_role(w.I, w.J, w.kbase, (Mp + 63) / 64);
    const int rows_i = min(64, Mp - w.I * 64), rows_j = diag ? 0 : min(64, Mp - w.J * 64);

    if (rows_i < 64 || (!diag && rows_j < 64)) {   // rows past Mp are never copied: they must read as zero
        for (int idx = threadIdx.x; idx < SY_NSTAGE * SY_STAGE; idx += SY_THREADS) smem[idx] = 0.0;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < SY_NSTAGE; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], SY_CONSUMERS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    SyrkSrc src;
    src.A = A; src.vbar = vbar; src.mubar = mubar; src.Mp = Mp; src.rows_i = rows_i; src.rows_j = rows_j;
    src.bytes = (unsigned)((rows_i + rows_j) * SY_STR * sizeof(double)) + (unsigned)(KC * K * sizeof(double)) * (mrole ? 2u : 1u);
    if (warp == 0 && lane == 0)   // prologue: the first SY_AHEAD stages
        for (int s = 0; s < SY_AHEAD && s < nstage; ++s) syrk_produce<KC>(w, src, smem, full, empty, s, K, mrole);
    // ---- consumers ----
    const int half = warp & 1;
    if (!diag) {
        if (half == 0) syrk_consumer<0, KC>(w, src, smem, full, empty, nstage, part, mraw_part, Mp, K, mrole);
        else syrk_consumer<1, KC>(w, src, smem, full, empty, nstage, part, mraw_part, Mp, K, mrole);
    } else {
        if (half == 0) syrk_consumer<2, KC>(w, src, smem, full, empty, nstage, part, mraw_part, Mp, K, mrole);
        else syrk_consumer<3, KC>(w, src, smem, full, empty, nstage, part, mraw_part, Mp, K, mrole);
    }
}

// ---- host plan ----------------------------------------------------------------------------------------------------
// Units = (tile pair, component group) with weight 32 (off-diagonal) or 18 (diagonal) DMMA per k4-step and warp.
// When there are fewer units than SMs, each unit's point range is split so that the CTA count is ~num_sms and the
// per-CTA work (weight x points) is as even as the integer split counts allow.
// relative cost of a diagonal-pair CTA per point (18 of 32 fragments per warp, but more operand loads per DMMA)
static double syrk_diag_weight() {
    static const double w = getenv("MGP_SYRK_DIAG_WT") ? atof(getenv("MGP_SYRK_DIAG_WT")) : 18.0;   // tuning hook
    return w;
}

static void syrk_split_counts(int Mp, int K, int num_sms, std::vector<int>& I_of, std::vector<int>& J_of,
                              std::vector<int>& kb_of, std::vector<int>& ns_of) {
    const int nb = (Mp + 63) / 64, kgroups = (K + 3) / 4;
    std::vector<double> wt;
    for (int I = 0; I < nb; ++I)
        for (int J = 0; J <= I; ++J)
            for (int kg = 0; kg < kgroups; ++kg) {
                I_of.push_back(I); J_of.push_back(J); kb_of.push_back(kg * 4);
                const int role = syrk_mraw_role(I, J, kg * 4, nb);
                wt.push_back((I == J ? syrk_diag_weight() : 32.0) + (double)((role & 1) + (role >> 1)));
            }
    const int nu = (int)wt.size();
    ns_of.assign(nu, 1);
    if (nu >= num_sms) return;
    double W = 0.0;
    for (double x : wt) W += x;
    int used = 0;
    for (int u = 0; u < nu; ++u) {
        ns_of[u] = std::max(1, (int)(num_sms * wt[u] / W));
        used += ns_of[u];
    }
    while (used > num_sms) {   // (only when the max(1, .) clamps pushed us over)
        int best = -1;
        for (int u = 0; u < nu; ++u)
            if (ns_of[u] > 1 && (best < 0 || wt[u] / ns_of[u] < wt[best] / ns_of[best])) best = u;
        if (best < 0) break;
        --ns_of[best]; --used;
    }
    while (used < num_sms) {   // hand the remaining CTAs to the most loaded units
        int best = 0;
        for (int u = 1; u < nu; ++u)
            if (wt[u] / ns_of[u] > wt[best] / ns_of[best]) best = u;
        ++ns_of[best]; ++used;
    }
}

int syrk_num_splits(int Mp, int K, const Launch& ln) {
    std::vector<int> I_of, J_of, kb_of, ns_of;
    syrk_split_counts(Mp, K, ln.num_sms, I_of, J_of, kb_of, ns_of);
    return *std::max_element(ns_of.begin(), ns_of.end());
}

int syrk_make_plan(int Mp, int K, int64_t n, int kc, int num_sms, std::vector<SyrkWork>& plan) {
    const int SY_KC = kc;
    std::vector<int> I_of, J_of, kb_of, ns_of;
    syrk_split_counts(Mp, K, num_sms, I_of, J_of, kb_of, ns_of);
    const int64_t nchunks = (n + SY_KC - 1) / SY_KC;
    plan.clear();
    struct Ord { double load; SyrkWork w; };
    std::vector<Ord> all;
    for (size_t u = 0; u < ns_of.size(); ++u) {
        const int ns = (int)std::min<int64_t>(ns_of[u], std::max<int64_t>(nchunks, 1));
        for (int j = 0; j < ns; ++j) {
            SyrkWork w;
            w.I = I_of[u]; w.J = J_of[u]; w.kbase = kb_of[u]; w.slot = j;
            w.pbeg = nchunks * j / ns * SY_KC;
            w.pend = nchunks * (j + 1) / ns * SY_KC;
            // (+1 DMMA per k4-step and warp for every row block whose A mubar rides on this pair)
            const int role = syrk_mraw_role(I_of[u], J_of[u], kb_of[u], (Mp + 63) / 64);
            const double wt = (I_of[u] == J_of[u] ? syrk_diag_weight() : 32.0) + (double)((role & 1) + (role >> 1));
            if (w.pend > w.pbeg) all.push_back({wt * (double)(w.pend - w.pbeg), w});
        }
    }
    std::stable_sort(all.begin(), all.end(), [](const Ord& a, const Ord& b) { return a.load > b.load; });   // heaviest first
    for (auto& o : all) plan.push_back(o.w);
    return (int)plan.size();
}

void syrk_accumulate(const LayerDev& ly, const ChunkBuffers& cb, const SyrkWork* d_plan, int nwork, double* part,
                     double* mraw_part, const Launch& ln) {
    if (nwork <= 0) return;
    auto launch = [&](auto kernel, int kc) {
        const size_t smem = (size_t)SY_NSTAGE * (2 * 64 * (kc + 4) + 2 * kc * KP) * sizeof(double);
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kernel<<<nwork, SY_THREADS, smem, ln.stream>>>(d_plan, cb.A, cb.vbar, cb.mubar, part, mraw_part, ly.Mp, ly.K);
        ln.tick();
    };
    if (cb.tw == 32) launch(syrk_kernel<32>, 32); else launch(syrk_kernel<16>, 16);
}

}  // namespace mgp
