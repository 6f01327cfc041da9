// The N-streaming FP64 tensor-core kernels of the SVGP conditional, forward and backward.
//
// Layout.  The [Mp x N] workspace arrays A = L^-1 Kuf and B_k = Lq_k^T A are stored TILE-MAJOR: tile ti (NT points)
// is one contiguous [Mp][NT + 4] block — exactly the shared-memory image the kernels multiply from (row stride
// NT + 4 == 4 mod 16 doubles => conflict-free B-fragment LDS.64).  A whole right operand is therefore fetched by ONE
// cp.async.bulk (TMA, UBLKCP) completing on an mbarrier.
//
// Structure.  Persistent CTAs walk tiles blockIdx.x, blockIdx.x + gridDim.x, ...  DMMA consumer warps (8 at 32-point
// tiles, 16 at 16-point tiles) each compute 16-row blocks  C = W[rows, k-range] * T  on DMMA.8x8x4 with the left operand
// W streamed from L2 in fragment-major order (one coalesced 256-byte load per fragment, software-pipelined across block /
// component / tile boundaries; helpers in stream_common.cuh).  Triangular left operands only visit their non-zero
// k-range; 16-row blocks are dealt to the warps in snake order so the triangular work balances.
//
//   cond_fwd_a : three CTAs per SM, barrier-phased: X rows -> Kuf tile (z.x contraction on DMMA + table exp, kept for
//                cond_bwd_b by one bulk store) -> A = L^-1 Kuf                                          -> A, Kuf
//   cond_fwd_b : ONE CTA per SM, two-deep ring of A tiles fed by warp 8 (bulk copies on mbarriers), which also finishes
//                the per-warp partial sums; K passes  B_k = Lq_k^T A                                    -> B_k, fmean, fvar
//   cond_bwd_a : ONE CTA per SM, two-deep ring of B_k stages (the last warp to leave a stage refills it); accumulators
//                of a warp's row blocks in registers across the K stages                                 -> Abar (over A)
//   cond_bwd_b : two CTAs per SM; Abar tile by bulk copy; E = (L^-T Abar) .* Kuf (read back)            -> sums E [1, xs, xs^2]
//
// The one-CTA pipelined / fused / ring forms of cond_fwd_a, cond_fwd_a + cond_fwd_b and cond_bwd_b that were built and
// measured slower live in stream_kernels_alt.cu.
//
// Reference arithmetic replaced: gpflow SquaredExponential.K + base_conditional as called from
// IndependentPosteriorSingleOutputModified._conditional_fused (MixtureGPs/models.py:129-144), evaluated once
// per point instead of once per (sample, point) (SGP.integrate only tiles X, models.py:35-36), and TF's
// reverse pass through it (utils/training_utils.py:8-10).  Math: SURVEY.md Appendix B.
#include "stream_common.cuh"

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
    double* Tb = smem + SK_BAR_DOUBLES;                // [NBUF][Mp][STR]
    double* mub = Tb + (size_t)NBUF * Mp * STR;        // [2][NT][K]  mubar slab of the tile (by tile parity)
    double* vbs = mub + 2 * NT * KP;                   // [2][NT][K]  vbar slab
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4;
    const size_t tile_elems = (size_t)Mp * STR;
    const unsigned tile_bytes = (unsigned)(tile_elems * sizeof(double));
    const unsigned slab_bytes = (unsigned)(NT * K * sizeof(double));
    if (threadIdx.x == 0) {
        for (int i = 0; i < NBUF; ++i) { mbar_init(&full[i], 1); mbar_init(&done[i], NW); }
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
                // (`done[buf]` counts the NW leavers; the arrival that finds one pending is the last: its release /
                //  the others' are acquired by the parity wait, which returns at once, before the refill is issued)
                if (mbar_arrive_pending(&done[buf]) == 1u) {
                    mbar_wait(&done[buf], (unsigned)((j / NBUF) & 1));
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

void cond_fwd_a(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln) {
    const int NT = cb.tw;
    const size_t smem = ((size_t)ly.Mp * (NT + 4) + NT * xs_stride(ly.Dp) + NT) * sizeof(double);
    const int ntiles = (int)((cb.n + NT - 1) / NT);
    const int threads = sk_warps(NT) * 32;
    if (cond_fwd_a_wants_pipe(NT)) { cond_fwd_a_pipe(ly, cb, ln); return; }   // (stream_kernels_alt.cu: measured slower)
    auto launch = [&](auto kernel) {
        const int grid = occupancy_grid(kernel, threads, smem, ntiles, 0, ln);
        kernel<<<grid, threads, smem, ln.stream>>>(ly, cb, ntiles);
        ln.tick();
    };
    if (NT == 32) launch(cond_fwd_a_kernel<32>); else launch(cond_fwd_a_kernel<16>);
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
    if (NT == 32) {
        if (cond_fwd_b_wants_16w(ly, NT)) { cond_fwd_b_16w(ly, cb, ln); return; }   // (stream_kernels_alt.cu: measured slower)
        launch(cond_fwd_b_kernel<32, 2>);
    }
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
    if (cond_bwd_b_wants_ring(NT, cb)) {   // (stream_kernels_alt.cu: measured slower)
        cond_bwd_b_ring(ly, cb, esum_part, nparts_cap, nparts, ln);
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
