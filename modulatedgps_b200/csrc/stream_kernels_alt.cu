// Alternative forms of cond_fwd_a / cond_fwd_b / cond_bwd_b that were BUILT, are parity-tested
// (tests/test_gpu_round2.py::test_alternative_kernel_forms_agree, ::test_fused_forward_kernel_matches_the_goldens) and
// MEASURED SLOWER than the forms in stream_kernels.cu at BASELINE config #4.  They stay selectable
// through environment switches for A/B timing and as the record of what was tried:
//   MGP_FWD_A_PIPE   software-pipelined one-CTA cond_fwd_a            5.94 ms vs 5.28 ms (barrier-phased, three CTAs per SM)
//   MGP_FUSED_FWD    cond_fwd_a + cond_fwd_b in one persistent kernel  23.7 ms vs 23.1 ms
//   MGP_BWD_B_RING   ring-form one-CTA cond_bwd_b                      5.47 ms vs 5.28 ms (two CTAs per SM)
//   MGP_FWD_B_16W    cond_fwd_b with 16 consumer warps (4 per sub-partition)  18.12 ms vs 17.57 ms (8 warps + loader warp)
#include "stream_common.cuh"

namespace mgp {

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
// cond_fwd_b, 16-warp form (32-point tiles, even K):  FOUR warps per sub-partition instead of two.
// The inner loop's ceiling is 35.4 TFLOP/s with two warps per sub-partition and 36.4 with four (tools/loop_bisect.cu:
// the L2 round trip of the left-operand fragments is covered by three other warps instead of one), and with four a
// warp's per-pass / per-tile epilogue always has other warps' DMMAs beside it.  MEASURED SLOWER all the same (18.12 vs
// 17.57 ms at config #4).  Sixteen warps own ONE 16-row block per
// pass each; the triangular imbalance (block b has 16 - b fragment groups) is evened out ACROSS passes — pass k deals
// the blocks in direction k & 1, so a warp's two consecutive passes sum to 17 groups — which works because each pass has
// its own output (cond_bwd_a's accumulators persist across its stages: no such form there).  No loader warp (the
// register file allows 128 registers per thread at 16 warps, 102 at 17): the last warp to leave a tile folds the
// partial sums into fmean / fvar and refills the buffer.
// ==================================================================================================
template <int NT, int NBUF>
__global__ void __launch_bounds__(SK_WARPS_NT16 * 32, 1) cond_fwd_b16_kernel(LayerDev ly, ChunkBuffers cb, int ntiles) {
    constexpr int NF = NT / 8, STR = NT + 4, NW = SK_WARPS_NT16, MW = SK_WARPS;
    extern __shared__ __align__(16) double smem[];
    const int Mp = ly.Mp, K = ly.K;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* done = full + NBUF;
    double* Tb = smem + SK_BAR_DOUBLES;                          // [NBUF][Mp][STR]
    double* sqpart = Tb + (size_t)NBUF * Mp * STR;               // [NBUF][NW][K][NT]  partial sum_m B_k^2
    double* mnpart = sqpart + (size_t)NBUF * NW * K * NT;        // [NBUF][MW][K][NT]  partial q_mu^T A
    double* aspart = mnpart + (size_t)NBUF * MW * K * NT;        // [NBUF][NW][NT]     partial |a_n|^2
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nb16 = Mp / 16, C4 = Mp / 4, R = (nb16 + NW - 1) / NW;
    const size_t tile_elems = (size_t)Mp * STR;
    const unsigned tile_bytes = (unsigned)(tile_elems * sizeof(double));
    if (threadIdx.x == 0) {
        for (int i = 0; i < NBUF; ++i) { mbar_init(&full[i], 1); mbar_init(&done[i], NW); }
        mbar_fence_init();
    }
    __syncthreads();
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto tile_of = [&](int i) { return (int64_t)blockIdx.x + (int64_t)i * gridDim.x; };
    auto issue = [&](int i) {   // one lane
        const int buf = i % NBUF;
        mbar_arrive_expect_tx(&full[buf], tile_bytes);
        bulk_g2s(Tb + (size_t)buf * tile_elems, cb.A + (size_t)tile_of(i) * tile_elems, tile_bytes, &full[buf]);
    };
    if (threadIdx.x == 0)
        for (int i = 0; i < NBUF && i < my_tiles; ++i) issue(i);
    // block of (pass k, round r) for this warp: direction alternates with r + k; -1 past the end
    auto blk = [&](int k, int r) {
        const int b = r * NW + (((r + k) & 1) ? (NW - 1 - warp) : warp);
        return b < nb16 ? b : -1;
    };
    // the unit after (k, r) in this warp's per-tile sequence that has a block (wraps to the next tile's first)
    auto next_unit = [&](int& k, int& r) {
        do {
            if (++r == R) { r = 0; if (++k == K) k = 0; }
        } while (blk(k, r) < 0);
    };
    auto seg_of = [&](int k, int r) {
        const int b = blk(k, r);
        return Seg{ly.W_LqT + (size_t)k * Mp * Mp, 2 * b, 4 * b};
    };
    const int mkb0 = warp * (C4 / MW), mkb1 = mkb0 + C4 / MW;   // fmean k-slice of warps < MW
    const double variance = ly.variance[0];
    int k0 = 0, r0 = 0;   // first unit of the sequence
    if (blk(0, 0) < 0) next_unit(k0, r0);
    WPair wp;
    { const Seg s0 = seg_of(k0, r0); wfrag_load(wp.f, s0, C4, s0.kb0, lane); }
    for (int i = 0; i < my_tiles; ++i) {
        const int buf = i % NBUF;
        const int64_t tile = tile_of(i);
        const double* T = Tb + (size_t)buf * tile_elems;
        double* sq = sqpart + ((size_t)buf * NW + warp) * K * NT;
        mbar_wait(&full[buf], (unsigned)((i / NBUF) & 1));
        for (int k = 0; k < K; ++k) {
            double colsq[NF][2];
#pragma unroll
            for (int nf = 0; nf < NF; ++nf) colsq[nf][0] = colsq[nf][1] = 0.0;
            double* Bk = cb.Bk ? cb.Bk + ((size_t)k * cb.tiles_cap + tile) * tile_elems : nullptr;
            for (int r = 0; r < R; ++r) {
                const int b = blk(k, r);
                if (b < 0) continue;
                double acc[2][NF][2];
                zero_acc<NF>(acc);
                int kn = k, rn = r;
                next_unit(kn, rn);
                wp.template run<NT, TRI_UPPER>(seg_of(k, r), C4, C4, T, acc, lane, seg_of(kn, rn));   // upper triangular
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
        if (warp < MW) {   // this warp's slice of fmean^T [K x NT] = q_mu^T [K x Mp] * A tile (rows >= K of W_mT are zero)
            double* mn = mnpart + ((size_t)buf * MW + warp) * K * NT;
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
        {   // |a_n|^2 over this warp's share of the rows (lane = column)
            double s0 = 0.0, s1 = 0.0;
            for (int m = warp * (Mp / NW); m < (warp + 1) * (Mp / NW); m += 2) {
                const double v0 = T[(size_t)m * STR + lane], v1 = T[(size_t)(m + 1) * STR + lane];
                s0 = fma(v0, v0, s0); s1 = fma(v1, v1, s1);
            }
            aspart[((size_t)buf * NW + warp) * NT + lane] = s0 + s1;
        }
        __syncwarp();
        unsigned last = 0;
        if (lane == 0) last = mbar_arrive_pending(&done[buf]) == 1u;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {   // every warp has left the tile: fold the partial sums, then refill the buffer
            mbar_wait(&done[buf], (unsigned)((i / NBUF) & 1));
            const int64_t n0 = tile * NT;
            const double* sqb = sqpart + (size_t)buf * NW * K * NT;
            const double* mnb = mnpart + (size_t)buf * MW * K * NT;
            double asq = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) asq += aspart[((size_t)buf * NW + w) * NT + lane];
            for (int k = 0; k < K; ++k) {
                double sv = 0.0, mv = 0.0;
#pragma unroll
                for (int w = 0; w < NW; ++w) sv += sqb[((size_t)w * K + k) * NT + lane];
#pragma unroll
                for (int w = 0; w < MW; ++w) mv += mnb[((size_t)w * K + k) * NT + lane];
                cb.fvar[(size_t)(n0 + lane) * K + k] = (variance - asq) + sv;   // Knn - sum A^2 + sum LTA^2
                cb.fmean[(size_t)(n0 + lane) * K + k] = mv;
            }
            __syncwarp();
            if (lane == 0 && i + NBUF < my_tiles) issue(i + NBUF);
        }
    }
}

// ==================================================================================================
// host side
// ==================================================================================================
bool cond_fwd_a_wants_pipe(int NT) { return NT == 32 && getenv("MGP_FWD_A_PIPE") != nullptr; }

void cond_fwd_a_pipe(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln) {
    // one CTA per SM, two tile buffers (they fit whenever NT = 32) and, when it fits too, a copy of the Z fragments
    const int NT = 32;
    const int ntiles = (int)((cb.n + NT - 1) / NT);
    size_t psmem = ((size_t)2 * ly.Mp * (NT + 4) + 2 * (NT * xs_stride(ly.Dp) + NT)) * sizeof(double);
    const size_t zbytes = (size_t)ly.Mp * ly.Dp * sizeof(double);
    const int zs_in_smem = psmem + zbytes <= (size_t)226 * 1024;
    if (zs_in_smem) psmem += zbytes;
    const int grid = persistent_grid(cond_fwd_a_pipe_kernel<32>, SK_CTHREADS + 32, psmem, ntiles, 0, ln);
    cond_fwd_a_pipe_kernel<32><<<grid, SK_CTHREADS + 32, psmem, ln.stream>>>(ly, cb, ntiles, zs_in_smem);
    ln.tick();
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

static size_t fwd_b_16w_smem(const LayerDev& ly, int NT) {
    return ((size_t)SK_BAR_DOUBLES + (size_t)2 * ly.Mp * (NT + 4) +
            (size_t)2 * ((SK_WARPS_NT16 + SK_WARPS) * ly.K * NT + SK_WARPS_NT16 * NT)) * sizeof(double);
}
// even K (a warp's two consecutive passes balance), every warp needs a block (Mp >= 256), and the larger partial-sum
// buffers must fit beside the two tiles
bool cond_fwd_b_wants_16w(const LayerDev& ly, int NT) {
    return NT == 32 && getenv("MGP_FWD_B_16W") != nullptr && ly.K % 2 == 0 && ly.Mp >= 256 &&
           fwd_b_16w_smem(ly, NT) <= (size_t)227 * 1024;
}

void cond_fwd_b_16w(const LayerDev& ly, const ChunkBuffers& cb, const Launch& ln) {
    const int NT = 32;
    const int ntiles = (int)((cb.n + NT - 1) / NT);
    const size_t smem16 = fwd_b_16w_smem(ly, NT);
    const int grid = persistent_grid(cond_fwd_b16_kernel<32, 2>, SK_WARPS_NT16 * 32, smem16, ntiles, 0, ln);
    cond_fwd_b16_kernel<32, 2><<<grid, SK_WARPS_NT16 * 32, smem16, ln.stream>>>(ly, cb, ntiles);
    ln.tick();
}

bool cond_bwd_b_wants_ring(int NT, const ChunkBuffers& cb) { return NT == 32 && cb.Kuf && getenv("MGP_BWD_B_RING") != nullptr; }

void cond_bwd_b_ring(const LayerDev& ly, const ChunkBuffers& cb, double* esum_part, int nparts_cap, int* nparts,
                     const Launch& ln) {
    const int NT = 32;
    const int ntiles = (int)((cb.n + NT - 1) / NT);
    const size_t rsmem = ((size_t)SK_BAR_DOUBLES + (size_t)2 * ly.Mp * (NT + 4) +
                          2 * (NT * xs_stride(ly.Dp) + NT + NT * (8 * esum_feature_blocks(ly.D) + 2))) * sizeof(double);
    const int grid = persistent_grid(cond_bwd_b_ring_kernel<32, 2>, SK_CTHREADS + 32, rsmem, ntiles, nparts_cap, ln);
    cond_bwd_b_ring_kernel<32, 2><<<grid, SK_CTHREADS + 32, rsmem, ln.stream>>>(ly, cb, ntiles, esum_part);
    ln.tick();
    if (grid > *nparts) *nparts = grid;
}

}  // namespace mgp
