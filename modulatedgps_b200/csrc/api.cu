// libmgp C-ABI (include/mgp.h): context, workspace and the orchestration of the kernels.
// No torch types, no exceptions across the boundary, no CPU fallback.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../include/mgp.h"
#include "kernels.h"

using namespace mgp;

static inline int64_t round_up64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

namespace {

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

struct LayerSlot {
    LayerDev dev{};
    Buf inv_ls, Zs_rm, Zs_fm, zs2, zh, Kuu, L, Linv, Dinv, W_Linv, W_LinvT, Lq_rm, W_LqT, Q_rm, W_Lq, W_m, W_mT, T1, T2, T3, Sfull, rowout;
    Buf A, Bk, Kuf, fmean, fvar, mubar, vbar;   // chunk buffers
    Buf syrk_part, mraw_part, esum_part;    // per-CTA partial sums
    Buf syrk_plan;                          // SyrkWork[] of the current (Mp, K, chunk length)
    int64_t plan_key[3] = {-1, -1, -1};
    int plan_len = 0;
    int nsplit = 0, esum_nparts = 0;
};

}  // namespace

enum Stage { ST_PRECOMPUTE = 0, ST_COND_FWD_A, ST_COND_FWD_B, ST_MC_PASS, ST_SYRK, ST_COND_BWD_A, ST_COND_BWD_B,
             ST_REDUCE, ST_FINISH, ST_COND_FWD, ST_COUNT };
static const char* const kStageNames[ST_COUNT] = {"precompute", "cond_fwd_a", "cond_fwd_b", "mc_pass", "syrk",
                                                  "cond_bwd_a", "cond_bwd_b", "reduce_partials", "finish", "cond_fwd"};

// Optional per-stage device timing with CUDA events on the launch stream (bench.py's live roofline numbers).
struct StageTimer {
    bool on = false;
    struct Rec { int stage; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    double ms[ST_COUNT] = {};
    int64_t calls[ST_COUNT] = {};
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void begin(int stage, cudaStream_t st) {
        if (!on) return;
        Rec r{stage, get(), get()};
        cudaEventRecord(r.a, st);
        recs.push_back(r);
    }
    void end(cudaStream_t st) {
        if (!on) return;
        cudaEventRecord(recs.back().b, st);
    }
    void collect(cudaStream_t st) {
        if (recs.empty()) return;
        cudaStreamSynchronize(st);
        for (auto& r : recs) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.stage] += t; calls[r.stage] += 1; }
            pool.push_back(r.a); pool.push_back(r.b);
        }
        recs.clear();
    }
};

struct mgp_ctx {
    StageTimer timer;
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;          // the assign layer's replicated work runs here, concurrently with the pred layer's
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t aux = nullptr;           // the q_mu / q_sqrt chains of both layers and the KL terms (independent of Kuu)
    cudaEvent_t ev_fork_aux = nullptr, ev_join_aux = nullptr;
    int num_sms = 0;
    size_t total_mem = 0;
    int64_t launches = 0;
    int64_t chunk_cap = 0;
    std::string err;
    LayerSlot slot[2];
    Buf mc_part, scratch_rb, kl, status;
    bool pre_valid = false;
    // what the precompute held in slot[0..1] was formed from: finish() re-forms it when it is handed other layers
    mgp_layer pre_layer[2] = {};
    double rm_squash = MGP_ROBUSTMAX_CDF_SQUASH;
    void* comm = nullptr;    // ncclComm_t handed over by mgp_ctx_set_comm (caller-owned), or NULL: the caller reduces
    bool serial_layers = false;   // MGP_SERIAL_LAYERS=1: both layers' streaming kernels on ONE stream (as round 1)
    Buf fprint;              // uint64 [4]: parameter fingerprints at mgp_elbo_local [0..1] and at mgp_elbo_finish [2..3]
    bool kl_valid = false;   // the KL terms in `kl` belong to the current precompute (formed by mgp_elbo_local)
    bool pick_valid = false;
    int64_t pick_key[5] = {0, 0, 0, 0, 0};
    int64_t pick_value = 0;
    std::vector<void*> owned;
};

namespace {

// ---- NCCL, bound at run time --------------------------------------------------------------------------------------
// libmgp does not link NCCL: the communicator comes from the caller (mgp_ctx_set_comm) and the one function used on it
// is looked up in the libnccl.so.2 the process already has loaded (the caller created the communicator with it).
typedef int (*nccl_all_reduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_error_string_fn)(int);
struct NcclApi {
    nccl_all_reduce_fn all_reduce = nullptr;
    nccl_error_string_fn error_string = nullptr;
    bool tried = false;
};
NcclApi& nccl_api() {
    static NcclApi api;
    if (!api.tried) {
        api.tried = true;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW);
        if (h) {
            api.all_reduce = (nccl_all_reduce_fn)dlsym(h, "ncclAllReduce");
            api.error_string = (nccl_error_string_fn)dlsym(h, "ncclGetErrorString");
        }
    }
    return api;
}
constexpr int NCCL_DOUBLE = 8, NCCL_SUM = 0;   // ncclFloat64, ncclSum (nccl.h, stable since 2.0)

// every entry point runs on the context's device and puts the caller's current device back on return
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) err = cudaSetDevice(device); else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define ON_CTX_DEVICE(c)                     \
    DeviceGuard dev_guard__((c)->device);    \
    CUDA_TRY(c, dev_guard__.err)

int fail(mgp_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

int cuda_fail(mgp_ctx* c, cudaError_t e, const char* where) {
    return fail(c, MGP_ERR_CUDA, std::string(where) + ": " + cudaGetErrorString(e));
}

#define CUDA_TRY(ctx, expr)                                          \
    do {                                                             \
        cudaError_t e__ = (expr);                                    \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #expr);   \
    } while (0)

int ensure(mgp_ctx* c, Buf& b, size_t bytes, bool zero = false) {
    if (bytes == 0) bytes = 8;
    if (b.cap >= bytes) return MGP_OK;
    if (b.p) {
        // stream-ordered with respect to our own work: wait before freeing
        cudaStreamSynchronize(c->stream);
        cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
    }
    const size_t want = (bytes + 255) / 256 * 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(c, MGP_ERR_NOMEM, "cudaMalloc of " + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e));
    }
    b.cap = want;
    if (zero) cudaMemsetAsync(b.p, 0, want, c->stream);
    return MGP_OK;
}

#define TRY(expr)                     \
    do {                              \
        int rc__ = (expr);            \
        if (rc__ != MGP_OK) return rc__; \
    } while (0)

// in-place sum over the ranks of the attached communicator, on `st` (stream-ordered; nothing when no communicator)
int all_reduce_sum(mgp_ctx* c, double* buf, int64_t n, cudaStream_t st) {
    if (!c->comm || n <= 0) return MGP_OK;
    NcclApi& api = nccl_api();
    if (!api.all_reduce) return fail(c, MGP_ERR_NCCL, "a communicator is attached but libnccl.so.2 / ncclAllReduce cannot be resolved");
    const int rc = api.all_reduce(buf, buf, (size_t)n, NCCL_DOUBLE, NCCL_SUM, c->comm, st);
    if (rc != 0) return fail(c, MGP_ERR_NCCL, std::string("ncclAllReduce: ") + (api.error_string ? api.error_string(rc) : "error"));
    return MGP_OK;
}

int check_layer(mgp_ctx* c, const mgp_layer* l) {
    if (!l) return fail(c, MGP_ERR_BAD_ARG, "layer is NULL");
    if (l->M < 1 || l->D < 1 || l->K < 1) return fail(c, MGP_ERR_BAD_ARG, "layer: M, D, K must be positive");
    if (l->K > MGP_MAX_K) return fail(c, MGP_ERR_BAD_ARG, "layer: K exceeds MGP_MAX_K (8)");
    if (l->D > MGP_MAX_D) return fail(c, MGP_ERR_BAD_ARG, "layer: D exceeds MGP_MAX_D (32)");
    if (l->n_lengthscales != 1 && l->n_lengthscales != l->D)
        return fail(c, MGP_ERR_BAD_ARG, "layer: n_lengthscales must be 1 or D");
    if (!l->Z || !l->q_mu || !l->q_sqrt || !l->variance || !l->lengthscales)
        return fail(c, MGP_ERR_BAD_ARG, "layer: NULL parameter pointer");
    const int Mp = (l->M + 31) / 32 * 32;
    if ((size_t)Mp * 20 * 8 + 16 * 1024 > 227 * 1024)
        return fail(c, MGP_ERR_BAD_ARG, "layer: M too large for the shared-memory tile (M <= 1312)");
    return MGP_OK;
}

// fill slot.dev for this layer and size its per-layer buffers
int setup_layer(mgp_ctx* c, LayerSlot& s, const mgp_layer* l, bool need_bwd) {
    TRY(check_layer(c, l));
    LayerDev& d = s.dev;
    d.M = l->M; d.D = l->D; d.K = l->K; d.n_ls = l->n_lengthscales;
    d.Mp = (l->M + 31) / 32 * 32;
    d.Dp = (l->D + 3) / 4 * 4;
    d.Z = l->Z; d.q_mu = l->q_mu; d.q_sqrt = l->q_sqrt; d.variance = l->variance; d.lengthscales = l->lengthscales;
    const size_t Mp = d.Mp, Dp = d.Dp, K = d.K, mm = Mp * Mp * sizeof(double);
    TRY(ensure(c, s.inv_ls, Dp * 8));
    TRY(ensure(c, s.Zs_rm, Mp * Dp * 8));
    TRY(ensure(c, s.Zs_fm, Mp * Dp * 8));
    TRY(ensure(c, s.zs2, (2 * Mp + 16) * 8));   // [Mp] |Zs_i|^2 + [Mp..] small scratch (KL partials)
    TRY(ensure(c, s.zh, Mp * 8));
    TRY(ensure(c, s.Kuu, mm));
    TRY(ensure(c, s.L, mm));
    TRY(ensure(c, s.Linv, mm));
    TRY(ensure(c, s.Dinv, Mp * 32 * 8));
    TRY(ensure(c, s.W_Linv, mm));
    TRY(ensure(c, s.W_LinvT, mm));
    TRY(ensure(c, s.Lq_rm, K * mm));
    TRY(ensure(c, s.W_LqT, K * mm));
    TRY(ensure(c, s.W_mT, 16 * Mp * 8));
    d.inv_ls = (double*)s.inv_ls.p; d.Zs_rm = (double*)s.Zs_rm.p; d.Zs_fm = (double*)s.Zs_fm.p; d.zs2 = (double*)s.zs2.p; d.zh = (double*)s.zh.p;
    d.Kuu = (double*)s.Kuu.p; d.L = (double*)s.L.p; d.Linv = (double*)s.Linv.p; d.Dinv = (double*)s.Dinv.p;
    d.W_Linv = (double*)s.W_Linv.p; d.W_LinvT = (double*)s.W_LinvT.p;
    d.Lq_rm = (double*)s.Lq_rm.p; d.W_LqT = (double*)s.W_LqT.p; d.W_mT = (double*)s.W_mT.p;
    if (need_bwd) {
        TRY(ensure(c, s.Q_rm, K * mm));
        TRY(ensure(c, s.W_Lq, K * mm));
        TRY(ensure(c, s.W_m, Mp * KP * 8));
        TRY(ensure(c, s.T1, K * mm));
        TRY(ensure(c, s.T2, mm));
        TRY(ensure(c, s.T3, mm));
        TRY(ensure(c, s.Sfull, K * mm));
        TRY(ensure(c, s.rowout, Mp * (2 * Dp + 1) * 8));
        d.Q_rm = (double*)s.Q_rm.p; d.W_Lq = (double*)s.W_Lq.p; d.W_m = (double*)s.W_m.p; d.T1 = (double*)s.T1.p; d.T2 = (double*)s.T2.p;
        d.T3 = (double*)s.T3.p; d.Sfull = (double*)s.Sfull.p; d.rowout = (double*)s.rowout.p;
    } else {
        // forward-only paths still use T1 as scratch of nothing; keep pointers null-safe
        d.Q_rm = d.W_Lq = d.W_m = d.T1 = d.T2 = d.T3 = d.Sfull = d.rowout = nullptr;
    }
    return MGP_OK;
}

// cond_fwd_a keeps the Kuf tiles it generates (one more tile-major array of A's size per layer) so that cond_bwd_b
// multiplies by them instead of generating them again; MGP_NO_KUF_STASH=1 restores the regeneration (A/B timing)
static bool kuf_stash_enabled() { return getenv("MGP_NO_KUF_STASH") == nullptr; }   // (read per call: tests toggle it)

int ensure_chunk(mgp_ctx* c, LayerSlot& s, int64_t ldn, bool need_bwd) {
    const size_t Mp = s.dev.Mp, K = s.dev.K;
    const size_t tw = layer_tile_width(s.dev.Mp, s.dev.Dp, s.dev.K);
    const size_t tiled = ((size_t)ldn / tw) * Mp * (tw + 4) * 8;   // tile-major [ldn/tw][Mp][tw+4]
    TRY(ensure(c, s.A, tiled, true));
    TRY(ensure(c, s.fmean, (size_t)ldn * K * 8, true));
    TRY(ensure(c, s.fvar, (size_t)ldn * K * 8, true));
    if (need_bwd) {
        TRY(ensure(c, s.Bk, K * tiled, true));
        if (kuf_stash_enabled()) TRY(ensure(c, s.Kuf, tiled, true));
        TRY(ensure(c, s.mubar, (size_t)ldn * K * 8, true));
        TRY(ensure(c, s.vbar, (size_t)ldn * K * 8, true));
    }
    return MGP_OK;
}

// (re)build the SYRK work plan of this layer for chunks of n points; uploaded only when (Mp, K, n) change
int ensure_syrk_plan(mgp_ctx* c, LayerSlot& s, int64_t n) {
    const int kc = layer_tile_width(s.dev.Mp, s.dev.Dp, s.dev.K);
    const int64_t key[3] = {s.dev.Mp, (int64_t)s.dev.K * 64 + kc, n};
    if (s.syrk_plan.p && key[0] == s.plan_key[0] && key[1] == s.plan_key[1] && key[2] == s.plan_key[2]) return MGP_OK;
    std::vector<SyrkWork> plan;
    s.plan_len = syrk_make_plan(s.dev.Mp, s.dev.K, n, kc, c->num_sms, plan);
    TRY(ensure(c, s.syrk_plan, plan.size() * sizeof(SyrkWork)));
    // pageable source: the copy is staged before the call returns, so `plan` may go out of scope
    CUDA_TRY(c, cudaMemcpyAsync(s.syrk_plan.p, plan.data(), plan.size() * sizeof(SyrkWork), cudaMemcpyHostToDevice, c->stream));
    for (int i = 0; i < 3; ++i) s.plan_key[i] = key[i];
    return MGP_OK;
}

ChunkBuffers chunk_of(const LayerSlot& s, const double* X, int64_t n, int64_t ldn) {
    ChunkBuffers cb;
    cb.n = n; cb.ldn = ldn; cb.X = X;
    cb.tw = layer_tile_width(s.dev.Mp, s.dev.Dp, s.dev.K);
    cb.tiles_cap = ldn / cb.tw;
    cb.A = (double*)s.A.p; cb.Bk = nullptr; cb.Kuf = nullptr; cb.fmean = (double*)s.fmean.p; cb.fvar = (double*)s.fvar.p;
    cb.mubar = (double*)s.mubar.p; cb.vbar = (double*)s.vbar.p;
    return cb;
}

Launch launch_of(mgp_ctx* c) { return Launch{c->stream, &c->launches, c->num_sms}; }

// fork the side stream off the main one / join it back (the two layers' replicated O(M^3) work is independent)
Launch fork_side(mgp_ctx* c) {
    cudaEventRecord(c->ev_fork, c->stream);
    cudaStreamWaitEvent(c->side, c->ev_fork, 0);
    return Launch{c->side, &c->launches, c->num_sms};
}
void join_side(mgp_ctx* c) {
    cudaEventRecord(c->ev_join, c->side);
    cudaStreamWaitEvent(c->stream, c->ev_join, 0);
}
Launch fork_aux(mgp_ctx* c) {
    cudaEventRecord(c->ev_fork_aux, c->stream);
    cudaStreamWaitEvent(c->aux, c->ev_fork_aux, 0);
    return Launch{c->aux, &c->launches, c->num_sms};
}
void join_aux(mgp_ctx* c) {
    cudaEventRecord(c->ev_join_aux, c->aux);
    cudaStreamWaitEvent(c->stream, c->ev_join_aux, 0);
}

// host-side stall finder (MGP_HOST_PROFILE=1): prints any bracketed host section that takes longer than 2 ms
struct HostTimer {
    const char* what;
    std::chrono::steady_clock::time_point t0;
    static bool on() { static const bool v = getenv("MGP_HOST_PROFILE") != nullptr; return v; }
    explicit HostTimer(const char* w) : what(w) { if (on()) t0 = std::chrono::steady_clock::now(); }
    ~HostTimer() {
        if (!on()) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 2.0) fprintf(stderr, "[mgp host] %s took %.2f ms\n", what, ms);
    }
};

struct Timed {   // RAII stage bracket
    mgp_ctx* c;
    Timed(mgp_ctx* ctx, int stage) : c(ctx) { c->timer.begin(stage, c->stream); }
    ~Timed() { c->timer.end(c->stream); }
};

// points per chunk so that the materialised A of `nlayers` layers fits the budget
int64_t pick_chunk_uncached(mgp_ctx* c, int64_t N, int Mp_max, int nlayers, int kcopies);
int64_t pick_chunk(mgp_ctx* c, int64_t N, int Mp_max, int nlayers, int kcopies) {
    // cudaMemGetInfo stalls the host for 3-80 ms every few calls (measured on the B200 box): ask once per distinct
    // request, not once per step
    const int64_t key[5] = {N, Mp_max, nlayers, kcopies, c->chunk_cap};
    if (c->pick_valid && memcmp(key, c->pick_key, sizeof(key)) == 0) return c->pick_value;
    const int64_t v = pick_chunk_uncached(c, N, Mp_max, nlayers, kcopies);
    memcpy(c->pick_key, key, sizeof(key));
    c->pick_value = v;
    c->pick_valid = true;
    return v;
}

int64_t pick_chunk_uncached(mgp_ctx* c, int64_t N, int Mp_max, int nlayers, int kcopies) {
    int64_t cap = c->chunk_cap;
    if (cap <= 0) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        double budget = fmin(0.4 * (double)c->total_mem, 0.7 * (double)free_b);
        for (int i = 0; i < 2; ++i) budget += (double)c->slot[i].A.cap + (double)c->slot[i].Bk.cap + (double)c->slot[i].Kuf.cap;   // what we already hold counts as available
        cap = (int64_t)(budget / ((double)nlayers * ((1.0 + kcopies) * Mp_max + 6 * MGP_MAX_K) * 8.0));
        if (cap < 4096) cap = 4096;
    }
    cap = cap / 64 * 64;
    if (cap < 64) cap = 64;
    return N < cap ? N : cap;
}

// Order-sensitive 64-bit fingerprint of a layer's parameter VALUES (Z, q_mu, variance, lengthscales and the diagonal
// of every q_sqrt_k — an optimiser step moves all of them): mgp_elbo_finish compares the one taken when
// mgp_elbo_local formed L, L^-1, Lq ... with the values it is handed, so parameters changed in place between the two
// calls are reported (MGP_ERR_STALE_PRECOMPUTE through mgp_check_status, ELBO = NaN) instead of silently mixing two
// parameter states in the Cholesky backward.  One CTA per layer.
struct FingerprintArgs { mgp_layer l[2]; };
__global__ void __launch_bounds__(256) fingerprint_kernel(FingerprintArgs a, unsigned long long* out) {
    __shared__ unsigned long long red[8];
    const mgp_layer& l = a.l[blockIdx.x];
    const int64_t nz = (int64_t)l.M * l.D, nm = (int64_t)l.M * l.K, nd = (int64_t)l.K * l.M;
    const int64_t total = nz + nm + 1 + l.n_lengthscales + nd;
    unsigned long long h = 0;
    for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
        int64_t j = i;
        double v;
        if (j < nz) v = l.Z[j];
        else if ((j -= nz) < nm) v = l.q_mu[j];
        else if ((j -= nm) < 1) v = l.variance[0];
        else if ((j -= 1) < l.n_lengthscales) v = l.lengthscales[j];
        else { j -= l.n_lengthscales; const int64_t k = j / l.M, r = j % l.M; v = l.q_sqrt[(k * l.M + r) * l.M + r]; }
        h += (unsigned long long)__double_as_longlong(v) * (unsigned long long)(2 * i + 1);
    }
    for (int o = 16; o; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[blockIdx.x] = t;
    }
}

__global__ void elbo_finalize_kernel(const double* rb, const double* kl, double num_data, int K, double* elbo,
                                     double* glik, double* galik, const unsigned long long* fprint, int* status) {
    const bool stale = fprint != nullptr && (fprint[0] != fprint[2] || fprint[1] != fprint[3]);
    // a Cholesky pivot was not positive since the last mgp_check_status: whatever the factorisation left behind, the
    // caller must see a non-finite ELBO (mgp_adam_step's guard keys on it; TF raises InvalidArgumentError here)
    const bool not_pd = (*status & 1) != 0;
    if (threadIdx.x == 0) {
        elbo[0] = (stale || not_pd) ? nan("") : rb[RB_DATA] - (kl[0] + kl[1]) / num_data;
        if (stale) atomicOr(status, 2);
    }
    if (threadIdx.x < K) {
        if (glik) glik[threadIdx.x] = rb[RB_LIKVAR + threadIdx.x];
        if (galik) galik[threadIdx.x] = rb[RB_ALIKVAR + threadIdx.x];
    }
}

int check_cfg(mgp_ctx* c, const mgp_elbo_cfg* cfg, const mgp_layer* pred, const mgp_layer* assign,
              const double* lik_var, const double* assign_lik_var) {
    if (!cfg) return fail(c, MGP_ERR_BAD_ARG, "cfg is NULL");
    if (cfg->model != MGP_MODEL_SMGP && cfg->model != MGP_MODEL_SMGP_MODIFIED) return fail(c, MGP_ERR_BAD_ARG, "cfg.model");
    if (cfg->lik != MGP_LIK_GAUSSIAN && cfg->lik != MGP_LIK_MULTICLASS) return fail(c, MGP_ERR_BAD_ARG, "cfg.lik");
    if (cfg->S < 1) return fail(c, MGP_ERR_BAD_ARG, "cfg.S must be >= 1");
    if (!(cfg->temperature > 0.0)) return fail(c, MGP_ERR_BAD_ARG, "cfg.temperature must be > 0");
    if (!(cfg->num_data > 0.0)) return fail(c, MGP_ERR_BAD_ARG, "cfg.num_data must be > 0");
    if (cfg->n_global < 1) return fail(c, MGP_ERR_BAD_ARG, "cfg.n_global must be >= 1");
    TRY(check_layer(c, pred));
    TRY(check_layer(c, assign));
    if (pred->K != assign->K || pred->D != assign->D) return fail(c, MGP_ERR_BAD_ARG, "pred/assign layers disagree on K or D");
    if (cfg->lik == MGP_LIK_GAUSSIAN && !lik_var) return fail(c, MGP_ERR_BAD_ARG, "Gaussian expert likelihood needs lik_var");
    if (cfg->lik == MGP_LIK_MULTICLASS && pred->K < 2) return fail(c, MGP_ERR_BAD_ARG, "MultiClass needs K >= 2");
    if (cfg->model == MGP_MODEL_SMGP_MODIFIED && !assign_lik_var)
        return fail(c, MGP_ERR_BAD_ARG, "SMGPModified needs assign_lik_var");
    return MGP_OK;
}

// forward conditional of one layer over [N, D] points into caller arrays (n*K doubles each)
int run_predict_f(mgp_ctx* c, LayerSlot& s, const double* X, int64_t N, double* fmean, double* fvar) {
    const Launch ln = launch_of(c);
    const int K = s.dev.K, D = s.dev.D;
    const int64_t Nc = pick_chunk(c, N, s.dev.Mp, 1, 0);
    const int64_t ldn = round_up64(Nc, 64);
    TRY(ensure_chunk(c, s, ldn, false));
    for (int64_t c0 = 0; c0 < N; c0 += Nc) {
        const int64_t n = (N - c0 < Nc) ? N - c0 : Nc;
        ChunkBuffers cb = chunk_of(s, X + c0 * D, n, ldn);
        if (cond_fwd_is_fused(s.dev, cb)) cond_fwd_fused(s.dev, cb, ln);
        else { cond_fwd_a(s.dev, cb, ln); cond_fwd_b(s.dev, cb, ln); }
        CUDA_TRY(c, cudaMemcpyAsync(fmean + c0 * K, cb.fmean, sizeof(double) * n * K, cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(fvar + c0 * K, cb.fvar, sizeof(double) * n * K, cudaMemcpyDeviceToDevice, c->stream));
    }
    return MGP_OK;
}

}  // namespace

// ====================================================================================================
extern "C" {

int mgp_ctx_create(int device, void* cuda_stream, mgp_ctx** out) {
    if (!out) return MGP_ERR_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1 || device < 0 || device >= ndev) return MGP_ERR_CUDA;   // no CPU fallback
    DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return MGP_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MGP_ERR_CUDA;
    mgp_ctx* c = new mgp_ctx();
    c->device = device;
    c->stream = (cudaStream_t)cuda_stream;
    c->num_sms = prop.multiProcessorCount;
    c->total_mem = prop.totalGlobalMem;
    { const char* e = getenv("MGP_SERIAL_LAYERS"); c->serial_layers = e && e[0] == '1'; }
    if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork_aux, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join_aux, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        delete c;
        return MGP_ERR_CUDA;
    }
    if (ensure(c, c->kl, 64) != MGP_OK || ensure(c, c->status, 64, true) != MGP_OK || ensure(c, c->fprint, 64, true) != MGP_OK) {
        delete c;
        return MGP_ERR_NOMEM;
    }
    *out = c;
    return MGP_OK;
}

void mgp_ctx_destroy(mgp_ctx* c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    cudaStreamSynchronize(c->stream);
    auto rel = [](Buf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; };
    for (auto& s : c->slot) {
        Buf* all[] = {&s.inv_ls, &s.Zs_rm, &s.Zs_fm, &s.zs2, &s.zh, &s.Kuu, &s.L, &s.Linv, &s.Dinv, &s.W_Linv, &s.W_LinvT, &s.Lq_rm,
                      &s.W_LqT, &s.Q_rm, &s.W_Lq, &s.W_m, &s.W_mT, &s.T1, &s.T2, &s.T3, &s.Sfull, &s.rowout, &s.A, &s.Bk, &s.Kuf,
                      &s.fmean, &s.fvar, &s.mubar, &s.vbar, &s.syrk_part, &s.mraw_part, &s.esum_part,
                      &s.syrk_plan};
        for (Buf* b : all) rel(*b);
    }
    rel(c->mc_part); rel(c->scratch_rb); rel(c->kl); rel(c->status); rel(c->fprint);
    if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
    if (c->aux) { cudaStreamSynchronize(c->aux); cudaStreamDestroy(c->aux); }
    if (c->ev_fork_aux) cudaEventDestroy(c->ev_fork_aux);
    if (c->ev_join_aux) cudaEventDestroy(c->ev_join_aux);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
}

const char* mgp_last_error(const mgp_ctx* c) { return c ? c->err.c_str() : "mgp: NULL context (no CUDA device?)"; }

int64_t mgp_launch_count(const mgp_ctx* c) { return c ? c->launches : 0; }

int mgp_set_chunk_points(mgp_ctx* c, int64_t max_points) {
    if (!c || max_points < 0) return MGP_ERR_BAD_ARG;
    c->chunk_cap = max_points;
    return MGP_OK;
}

int mgp_ctx_set_comm(mgp_ctx* c, void* nccl_comm) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (nccl_comm && !nccl_api().all_reduce)
        return fail(c, MGP_ERR_NCCL, "libnccl.so.2 (ncclAllReduce) cannot be resolved in this process");
    c->comm = nccl_comm;
    return MGP_OK;
}

int mgp_all_reduce(mgp_ctx* c, double* buf, int64_t n) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (n < 0 || (n > 0 && !buf)) return fail(c, MGP_ERR_BAD_ARG, "all_reduce: bad arguments");
    ON_CTX_DEVICE(c);
    return all_reduce_sum(c, buf, n, c->stream);
}

int mgp_set_robustmax_squash(mgp_ctx* c, double squash) {
    if (!c || !(squash >= 0.0) || !(squash < 0.5)) return c ? fail(c, MGP_ERR_BAD_ARG, "robustmax squash must be in [0, 0.5)") : MGP_ERR_BAD_ARG;
    c->rm_squash = squash;
    return MGP_OK;
}

int mgp_timing_enable(mgp_ctx* c, int on) {
    if (!c) return MGP_ERR_BAD_ARG;
    ON_CTX_DEVICE(c);
    c->timer.collect(c->stream);
    c->timer.on = on != 0;
    return MGP_OK;
}

int mgp_timing_read(mgp_ctx* c, double* ms, int64_t* calls, int reset) {
    if (!c || !ms || !calls) return MGP_ERR_BAD_ARG;
    ON_CTX_DEVICE(c);
    c->timer.collect(c->stream);
    for (int i = 0; i < ST_COUNT; ++i) { ms[i] = c->timer.ms[i]; calls[i] = c->timer.calls[i]; }
    if (reset) for (int i = 0; i < ST_COUNT; ++i) { c->timer.ms[i] = 0.0; c->timer.calls[i] = 0; }
    return MGP_OK;
}

int mgp_num_stages(void) { return ST_COUNT; }

const char* mgp_stage_name(int i) { return (i >= 0 && i < ST_COUNT) ? kStageNames[i] : ""; }

int mgp_check_status(mgp_ctx* c) {
    if (!c) return MGP_ERR_BAD_ARG;
    ON_CTX_DEVICE(c);
    int h = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&h, c->status.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaGetLastError());
    if (h) cudaMemsetAsync(c->status.p, 0, sizeof(int), c->stream);
    if (h & 1) return fail(c, MGP_ERR_NOT_PD, "Cholesky of Kuu + jitter I failed: matrix is not positive definite");
    if (h & 2)
        return fail(c, MGP_ERR_STALE_PRECOMPUTE, "mgp_elbo_finish: parameter values changed after mgp_elbo_local factorised "
                                                 "them (ELBO set to NaN; call mgp_elbo_local again)");
    return MGP_OK;
}

int mgp_svgp_predict_f(mgp_ctx* c, const mgp_layer* layer, const double* X, int64_t N, double* fmean, double* fvar) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (N < 0 || (N > 0 && (!X || !fmean || !fvar))) return fail(c, MGP_ERR_BAD_ARG, "predict_f: bad arguments");
    ON_CTX_DEVICE(c);
    c->pre_valid = false; c->kl_valid = false;
    TRY(setup_layer(c, c->slot[0], layer, false));
    if (N == 0) return MGP_OK;
    precompute_layer(c->slot[0].dev, false, (int*)c->status.p, launch_of(c));
    TRY(run_predict_f(c, c->slot[0], X, N, fmean, fvar));
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_prior_kl(mgp_ctx* c, const mgp_layer* layer, double* kl) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (!kl) return fail(c, MGP_ERR_BAD_ARG, "prior_kl: NULL output");
    ON_CTX_DEVICE(c);
    c->pre_valid = false; c->kl_valid = false;
    TRY(setup_layer(c, c->slot[0], layer, false));
    prior_kl_layer(c->slot[0].dev, kl, launch_of(c));
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_predict_y(mgp_ctx* c, const mgp_layer* pred, int32_t lik, const double* lik_var, const double* X, int64_t N,
                  double* mean, double* var) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (lik == MGP_LIK_GAUSSIAN && !lik_var) return fail(c, MGP_ERR_BAD_ARG, "predict_y: Gaussian likelihood needs lik_var");
    ON_CTX_DEVICE(c);
    TRY(mgp_svgp_predict_f(c, pred, X, N, mean, var));
    if (N == 0) return MGP_OK;
    predict_y_kernel(mean, var, N, pred->K, lik, lik_var, c->rm_squash, mean, var, launch_of(c));
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_predict_assign(mgp_ctx* c, const mgp_layer* assign, const double* X, int64_t N, double* probs, int64_t* argmax) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (N > 0 && (!probs || !argmax)) return fail(c, MGP_ERR_BAD_ARG, "predict_assign: NULL output");
    if (N == 0) return MGP_OK;
    ON_CTX_DEVICE(c);
    TRY(ensure(c, c->scratch_rb, (size_t)N * assign->K * 8));
    TRY(mgp_svgp_predict_f(c, assign, X, N, probs, (double*)c->scratch_rb.p));
    predict_assign_kernel(probs, N, assign->K, probs, argmax, launch_of(c));
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_predict_samples(mgp_ctx* c, const mgp_layer* pred, const mgp_layer* assign, int32_t lik, const double* lik_var,
                        const double* X, int64_t N, int32_t S, double temperature, const mgp_noise* noise,
                        const double* z_pred, double* samples_y, double* samples_f) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (N < 0 || S < 1 || !noise || !(temperature > 0.0)) return fail(c, MGP_ERR_BAD_ARG, "predict_samples: bad arguments");
    if ((noise->z == nullptr) != (noise->u == nullptr) || (noise->z != nullptr && !z_pred))
        return fail(c, MGP_ERR_BAD_ARG, "predict_samples: z, u and z_pred must be given together");
    if (lik == MGP_LIK_GAUSSIAN && !lik_var) return fail(c, MGP_ERR_BAD_ARG, "predict_samples: needs lik_var");
    if (N == 0) return MGP_OK;
    ON_CTX_DEVICE(c);
    c->pre_valid = false; c->kl_valid = false;
    TRY(check_layer(c, pred));
    TRY(check_layer(c, assign));
    if (pred->K != assign->K || pred->D != assign->D) return fail(c, MGP_ERR_BAD_ARG, "pred/assign layers disagree on K or D");
    const int K = pred->K;
    TRY(ensure(c, c->scratch_rb, (size_t)N * K * 8 * 4));
    double* fm_p = (double*)c->scratch_rb.p;
    double *fv_p = fm_p + N * K, *fm_a = fv_p + N * K, *fv_a = fm_a + N * K;
    const Launch ln = launch_of(c);
    TRY(setup_layer(c, c->slot[0], pred, false));
    precompute_layer(c->slot[0].dev, false, (int*)c->status.p, ln);
    TRY(run_predict_f(c, c->slot[0], X, N, fm_p, fv_p));
    TRY(setup_layer(c, c->slot[1], assign, false));
    precompute_layer(c->slot[1].dev, false, (int*)c->status.p, ln);
    TRY(run_predict_f(c, c->slot[1], X, N, fm_a, fv_a));
    SampleArgs a;
    a.S = S; a.K = K; a.lik = lik; a.temperature = temperature; a.squash = c->rm_squash; a.n = N;
    a.fmean_p = fm_p; a.fvar_p = fv_p; a.fmean_a = fm_a; a.fvar_a = fv_a; a.lik_var = lik_var;
    a.z = noise->z; a.u = noise->u; a.z_pred = z_pred; a.seed = noise->seed; a.point_offset = noise->point_offset;
    a.samples_y = samples_y; a.samples_f = samples_f;
    predict_samples_kernel(a, ln);
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_w_sample(mgp_ctx* c, const mgp_layer* assign, const double* X, int64_t N, int32_t S, double temperature,
                 const mgp_noise* noise, double* W) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (N < 0 || S < 1 || !noise || !(temperature > 0.0) || (N > 0 && (!X || !W)))
        return fail(c, MGP_ERR_BAD_ARG, "w_sample: bad arguments");
    if ((noise->z == nullptr) != (noise->u == nullptr)) return fail(c, MGP_ERR_BAD_ARG, "w_sample: z and u must be given together");
    if (N == 0) return MGP_OK;
    ON_CTX_DEVICE(c);
    c->pre_valid = false; c->kl_valid = false;
    TRY(check_layer(c, assign));
    const int K = assign->K;
    TRY(ensure(c, c->scratch_rb, (size_t)N * K * 8 * 2));
    double* fm = (double*)c->scratch_rb.p;
    double* fv = fm + N * K;
    const Launch ln = launch_of(c);
    TRY(setup_layer(c, c->slot[1], assign, false));
    precompute_layer(c->slot[1].dev, false, (int*)c->status.p, ln);
    TRY(run_predict_f(c, c->slot[1], X, N, fm, fv));
    SampleArgs a;
    memset(&a, 0, sizeof(a));
    a.S = S; a.K = K; a.lik = 0; a.temperature = temperature; a.squash = c->rm_squash; a.n = N;
    a.fmean_a = fm; a.fvar_a = fv;
    a.z = noise->z; a.u = noise->u; a.seed = noise->seed; a.point_offset = noise->point_offset;
    w_sample_kernel(a, W, ln);
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_e_log_p_y(mgp_ctx* c, const mgp_layer* pred, int32_t lik, const double* lik_var, const double* X, const double* Y,
                  int64_t N, int32_t S, const double* W, double* out) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (N < 0 || S < 1 || (N > 0 && (!X || !Y || !W || !out))) return fail(c, MGP_ERR_BAD_ARG, "e_log_p_y: bad arguments");
    if (lik != MGP_LIK_GAUSSIAN && lik != MGP_LIK_MULTICLASS) return fail(c, MGP_ERR_BAD_ARG, "e_log_p_y: lik");
    if (lik == MGP_LIK_GAUSSIAN && !lik_var) return fail(c, MGP_ERR_BAD_ARG, "e_log_p_y: Gaussian likelihood needs lik_var");
    if (N == 0) return MGP_OK;
    ON_CTX_DEVICE(c);
    c->pre_valid = false; c->kl_valid = false;
    TRY(check_layer(c, pred));
    const int K = pred->K;
    TRY(ensure(c, c->scratch_rb, (size_t)N * K * 8 * 2));
    double* fm = (double*)c->scratch_rb.p;
    double* fv = fm + N * K;
    const Launch ln = launch_of(c);
    TRY(setup_layer(c, c->slot[0], pred, false));
    precompute_layer(c->slot[0].dev, false, (int*)c->status.p, ln);
    TRY(run_predict_f(c, c->slot[0], X, N, fm, fv));
    e_log_p_y_kernel(fm, fv, Y, lik_var, lik, c->rm_squash, W, S, N, K, out, ln);
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

static int lik_call(mgp_ctx* c, int mode, int32_t lik, const double* lik_var, const double* Fmu, const double* Fvar,
                    const double* Y, int64_t S, int64_t N, int32_t K, double* out, const char* what) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (S < 0 || N < 0 || K < 1 || K > MGP_MAX_K) return fail(c, MGP_ERR_BAD_ARG, std::string(what) + ": bad sizes");
    if (lik != MGP_LIK_GAUSSIAN && lik != MGP_LIK_MULTICLASS) return fail(c, MGP_ERR_BAD_ARG, std::string(what) + ": lik");
    if (lik == MGP_LIK_GAUSSIAN && !lik_var) return fail(c, MGP_ERR_BAD_ARG, std::string(what) + ": Gaussian likelihood needs lik_var");
    if (lik == MGP_LIK_MULTICLASS && (mode != 0 || K < 2)) return fail(c, MGP_ERR_BAD_ARG, std::string(what) + ": not defined for MultiClass");
    if (S * N == 0) return MGP_OK;
    if (!Fmu || !Y || !out || (mode != 1 && !Fvar)) return fail(c, MGP_ERR_BAD_ARG, std::string(what) + ": NULL argument");
    ON_CTX_DEVICE(c);
    lik_eval_kernel(mode, lik, lik_var, c->rm_squash, Fmu, Fvar, Y, S * N, N, K, out, launch_of(c));
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_lik_variational_expectations(mgp_ctx* c, int32_t lik, const double* lik_var, const double* Fmu, const double* Fvar,
                                     const double* Y, int64_t S, int64_t N, int32_t K, double* out) {
    return lik_call(c, 0, lik, lik_var, Fmu, Fvar, Y, S, N, K, out, "lik_variational_expectations");
}

int mgp_lik_log_prob(mgp_ctx* c, const double* lik_var, const double* F, const double* Y, int64_t S, int64_t N, int32_t K,
                     double* out) {
    return lik_call(c, 1, MGP_LIK_GAUSSIAN, lik_var, F, nullptr, Y, S, N, K, out, "lik_log_prob");
}

int mgp_lik_predict_log_density(mgp_ctx* c, const double* lik_var, const double* Fmu, const double* Fvar, const double* Y,
                                int64_t S, int64_t N, int32_t K, double* out) {
    return lik_call(c, 2, MGP_LIK_GAUSSIAN, lik_var, Fmu, Fvar, Y, S, N, K, out, "lik_predict_log_density");
}

int mgp_lik_predict_mean_and_var(mgp_ctx* c, int32_t lik, const double* lik_var, const double* Fmu, const double* Fvar,
                                 int64_t rows, int32_t K, double* mean, double* var) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (rows < 0 || K < 1 || K > MGP_MAX_K) return fail(c, MGP_ERR_BAD_ARG, "lik_predict_mean_and_var: bad sizes");
    if (lik != MGP_LIK_GAUSSIAN && lik != MGP_LIK_MULTICLASS) return fail(c, MGP_ERR_BAD_ARG, "lik_predict_mean_and_var: lik");
    if (lik == MGP_LIK_GAUSSIAN && !lik_var) return fail(c, MGP_ERR_BAD_ARG, "lik_predict_mean_and_var: needs lik_var");
    if (lik == MGP_LIK_MULTICLASS && K < 2) return fail(c, MGP_ERR_BAD_ARG, "lik_predict_mean_and_var: MultiClass needs K >= 2");
    if (rows == 0) return MGP_OK;
    if (!Fmu || !Fvar || !mean || !var) return fail(c, MGP_ERR_BAD_ARG, "lik_predict_mean_and_var: NULL argument");
    ON_CTX_DEVICE(c);
    predict_y_kernel(Fmu, Fvar, rows, K, lik, lik_var, c->rm_squash, mean, var, launch_of(c));
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_debug_philox(mgp_ctx* c, const uint32_t* ctr_key, int32_t n, uint32_t* out) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (n < 0 || (n > 0 && (!ctr_key || !out))) return fail(c, MGP_ERR_BAD_ARG, "debug_philox: bad arguments");
    ON_CTX_DEVICE(c);
    philox_kat_kernel(ctr_key, n, out, launch_of(c));
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_debug_noise(mgp_ctx* c, const mgp_noise* noise, int64_t N, int32_t S, int32_t K, int32_t stream, double* z, double* u) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (!noise || N < 0 || S < 1 || K < 1 || K > MGP_MAX_K || stream < 0 || stream > 255 || (N > 0 && (!z || !u)))
        return fail(c, MGP_ERR_BAD_ARG, "debug_noise: bad arguments");
    ON_CTX_DEVICE(c);
    philox_draws_kernel(noise->seed, noise->point_offset, N, S, K, stream, z, u, launch_of(c));
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int64_t mgp_reduce_buffer_len(const mgp_layer* pred, const mgp_layer* assign) {
    if (!pred || !assign) return 0;
    const int Mp_p = (pred->M + 31) / 32 * 32, Mp_a = (assign->M + 31) / 32 * 32;
    const int Dp = (pred->D + 3) / 4 * 4;
    const LayerRB rp = layer_rb(RB_HEADER, Mp_p, Dp, pred->K);
    const LayerRB ra = layer_rb(rp.end, Mp_a, Dp, assign->K);
    return ra.end;
}

int mgp_elbo_local(mgp_ctx* c, const mgp_elbo_cfg* cfg, const mgp_layer* pred, const mgp_layer* assign,
                   const double* lik_var, const double* assign_lik_var, const double* X, const double* Y,
                   int64_t N_local, const mgp_noise* noise, double* reduce_buf) {
    if (!c) return MGP_ERR_BAD_ARG;
    TRY(check_cfg(c, cfg, pred, assign, lik_var, assign_lik_var));
    if (N_local < 0 || !noise || !reduce_buf || (N_local > 0 && (!X || !Y)))
        return fail(c, MGP_ERR_BAD_ARG, "elbo_local: bad arguments");
    if ((noise->z == nullptr) != (noise->u == nullptr)) return fail(c, MGP_ERR_BAD_ARG, "noise: z and u must be given together");
    ON_CTX_DEVICE(c);
    const Launch ln = launch_of(c);
    LayerSlot &sp = c->slot[0], &sa = c->slot[1];
    TRY(setup_layer(c, sp, pred, true));
    TRY(setup_layer(c, sa, assign, true));
    const int K = pred->K, D = pred->D;
    const LayerRB rp = layer_rb(RB_HEADER, sp.dev.Mp, sp.dev.Dp, K);
    const LayerRB ra = layer_rb(rp.end, sa.dev.Mp, sa.dev.Dp, K);
    CUDA_TRY(c, cudaMemsetAsync(reduce_buf, 0, sizeof(double) * ra.end, c->stream));
    const int maxparts = stream_max_parts(ln);
    LayerSlot* slots[2] = {&sp, &sa};
    for (LayerSlot* s : slots) {
        const size_t Mp = s->dev.Mp, E = 1 + 2 * s->dev.Dp;
        s->nsplit = syrk_num_splits(s->dev.Mp, K, ln);
        TRY(ensure(c, s->syrk_part, (size_t)s->nsplit * K * Mp * Mp * 8));
        TRY(ensure(c, s->mraw_part, (size_t)s->nsplit * Mp * KP * 8));
        TRY(ensure(c, s->esum_part, (size_t)maxparts * Mp * E * 8));
        s->esum_nparts = 0;
    }

    {
        HostTimer ht("precompute launches");
        Timed t(c, ST_PRECOMPUTE);
        // three concurrent chains: the Cholesky chain of each layer (single-CTA, latency-bound kernels: the critical
        // path) and, on the aux stream, everything that depends on q_mu / q_sqrt only, including the KL terms (they
        // used to sit on the serial tail of mgp_elbo_finish)
        const Launch ls = fork_side(c);
        const Launch lx = fork_aux(c);
        precompute_chol(sp.dev, true, (int*)c->status.p, ln);
        precompute_chol(sa.dev, true, (int*)c->status.p, ls);
        for (LayerSlot* s : slots) {   // the per-CTA partial-sum buffers (76 MB at config #4) are cleared beside the Cholesky chains
            const size_t Mp = s->dev.Mp, E = 1 + 2 * s->dev.Dp;
            CUDA_TRY(c, cudaMemsetAsync(s->syrk_part.p, 0, (size_t)s->nsplit * K * Mp * Mp * 8, c->aux));
            CUDA_TRY(c, cudaMemsetAsync(s->mraw_part.p, 0, (size_t)s->nsplit * Mp * KP * 8, c->aux));
            CUDA_TRY(c, cudaMemsetAsync(s->esum_part.p, 0, (size_t)maxparts * Mp * E * 8, c->aux));
        }
        precompute_lq(sp.dev, true, lx);
        precompute_lq(sa.dev, true, lx);
        prior_kl_precomputed(sp.dev, (double*)c->kl.p, lx);
        prior_kl_precomputed(sa.dev, (double*)c->kl.p + 1, lx);
        {
            FingerprintArgs fa;
            fa.l[0] = *pred; fa.l[1] = *assign;
            fingerprint_kernel<<<2, 256, 0, c->aux>>>(fa, (unsigned long long*)c->fprint.p);
            c->launches += 1;
        }
        join_side(c);
        join_aux(c);
    }
    c->pre_valid = true;
    c->kl_valid = true;
    c->pre_layer[0] = *pred;
    c->pre_layer[1] = *assign;
    if (N_local == 0) {                // an empty shard contributes zeros ...
        if (!c->comm) return MGP_OK;
        // ... but takes part in the collective: the other ranks are waiting in the same two all-reduces
        const LayerRB rp0 = layer_rb(RB_HEADER, sp.dev.Mp, sp.dev.Dp, K);
        const LayerRB ra0 = layer_rb(rp0.end, sa.dev.Mp, sa.dev.Dp, K);
        TRY(all_reduce_sum(c, reduce_buf, rp0.end, c->stream));
        TRY(all_reduce_sum(c, reduce_buf + rp0.end, ra0.end - rp0.end, c->stream));
        return MGP_OK;
    }

    const int Mp_max = sp.dev.Mp > sa.dev.Mp ? sp.dev.Mp : sa.dev.Mp;
    int64_t Nc;
    { HostTimer ht("pick_chunk (cudaMemGetInfo)"); Nc = pick_chunk(c, N_local, Mp_max, 2, K + (kuf_stash_enabled() ? 1 : 0)); }
    const int64_t ldn = round_up64(Nc, 64);
    HostTimer ht_rest("elbo_local after pick_chunk");
    TRY(ensure_chunk(c, sp, ldn, true));
    TRY(ensure_chunk(c, sa, ldn, true));
    const int nblocks_max = mc_num_blocks(ldn);
    TRY(ensure(c, c->mc_part, (size_t)nblocks_max * MC_NPART * 8));

    for (int64_t c0 = 0; c0 < N_local; c0 += Nc) {
        const int64_t n = (N_local - c0 < Nc) ? N_local - c0 : Nc;
        const int64_t ldc = round_up64(n, 64);   // padded extent of THIS chunk (<= ldn); leading dimension stays ldn
        ChunkBuffers cp = chunk_of(sp, X + c0 * D, n, ldn), ca = chunk_of(sa, X + c0 * D, n, ldn);
        cp.Bk = (double*)sp.Bk.p;
        ca.Bk = (double*)sa.Bk.p;
        if (kuf_stash_enabled()) {   // (the fused forward kernel does not keep its Kuf tiles)
            if (!cond_fwd_is_fused(sp.dev, cp)) cp.Kuf = (double*)sp.Kuf.p;
            if (!cond_fwd_is_fused(sa.dev, ca)) ca.Kuf = (double*)sa.Kuf.p;
        }
        // The two layers are independent until the Monte-Carlo pass and again after it: the assign layer's streaming
        // kernels go to the side stream, so that its CTAs fill the SMs the pred layer's persistent CTAs leave at their
        // tail (and the other way round) instead of every launch paying its own fill and tail — a per-launch loss of
        // 2-7 % on an eighth of the points (8 GPUs).  With the stage timers on, everything stays on one stream so that a
        // timed interval is one kernel.
        const bool dual = !c->serial_layers && !c->timer.on;
        {
            const Launch la = dual ? fork_side(c) : ln;
            if (cond_fwd_is_fused(sp.dev, cp)) { Timed t(c, ST_COND_FWD); cond_fwd_fused(sp.dev, cp, ln); }
            else {
                { HostTimer ht("fwd_a p"); Timed t(c, ST_COND_FWD_A); cond_fwd_a(sp.dev, cp, ln); }
                { HostTimer ht("fwd_b p"); Timed t(c, ST_COND_FWD_B); cond_fwd_b(sp.dev, cp, ln); }
            }
            if (cond_fwd_is_fused(sa.dev, ca)) { Timed t(c, ST_COND_FWD); cond_fwd_fused(sa.dev, ca, la); }
            else {
                { HostTimer ht("fwd_a a"); Timed t(c, ST_COND_FWD_A); cond_fwd_a(sa.dev, ca, la); }
                { HostTimer ht("fwd_b a"); Timed t(c, ST_COND_FWD_B); cond_fwd_b(sa.dev, ca, la); }
            }
            if (dual) join_side(c);
        }
        McArgs m;
        m.model = cfg->model; m.lik = cfg->lik; m.S = cfg->S; m.K = K;
        m.temperature = cfg->temperature;
        m.squash = c->rm_squash;
        m.inv_n_global = 1.0 / (double)cfg->n_global;
        m.n = n; m.ldn = ldc; m.n_local = N_local; m.chunk_offset = c0;
        m.Y = Y + c0;
        m.fmean_p = cp.fmean; m.fvar_p = cp.fvar; m.fmean_a = ca.fmean; m.fvar_a = ca.fvar;
        m.mubar_p = cp.mubar; m.vbar_p = cp.vbar; m.mubar_a = ca.mubar; m.vbar_a = ca.vbar;
        m.lik_var = lik_var; m.assign_lik_var = assign_lik_var;
        m.z = noise->z; m.u = noise->u; m.seed = noise->seed; m.point_offset = noise->point_offset;
        {
            HostTimer ht("mc_pass");
            Timed t(c, ST_MC_PASS);
            mc_pass(m, (double*)c->mc_part.p, ln);
            mc_fold((double*)c->mc_part.p, mc_num_blocks(ldc), reduce_buf, ln);
        }
        { HostTimer ht("syrk plan"); TRY(ensure_syrk_plan(c, sp, n)); TRY(ensure_syrk_plan(c, sa, n)); }
        HostTimer ht_bwd("syrk + bwd launches");
        {
            const Launch la = dual ? fork_side(c) : ln;
            { Timed t(c, ST_SYRK); syrk_accumulate(sp.dev, cp, (const SyrkWork*)sp.syrk_plan.p, sp.plan_len, (double*)sp.syrk_part.p, (double*)sp.mraw_part.p, ln); }
            if (!dual) { Timed t(c, ST_SYRK); syrk_accumulate(sa.dev, ca, (const SyrkWork*)sa.syrk_plan.p, sa.plan_len, (double*)sa.syrk_part.p, (double*)sa.mraw_part.p, la); }
            { Timed t(c, ST_COND_BWD_A); cond_bwd_a(sp.dev, cp, ln); }
            { Timed t(c, ST_COND_BWD_B); cond_bwd_b(sp.dev, cp, (double*)sp.esum_part.p, maxparts, &sp.esum_nparts, ln); }
            if (dual) syrk_accumulate(sa.dev, ca, (const SyrkWork*)sa.syrk_plan.p, sa.plan_len, (double*)sa.syrk_part.p, (double*)sa.mraw_part.p, la);
            { Timed t(c, ST_COND_BWD_A); cond_bwd_a(sa.dev, ca, la); }
            { Timed t(c, ST_COND_BWD_B); cond_bwd_b(sa.dev, ca, (double*)sa.esum_part.p, maxparts, &sa.esum_nparts, la); }
            if (dual) join_side(c);
        }
    }
    const LayerRB* rbs[2] = {&rp, &ra};
    {
        Timed t_reduce(c, ST_REDUCE);
        const Launch lside = fork_side(c);   // the two layers' partial sums are independent: one stream each
        for (int i = 0; i < 2; ++i) {
            LayerSlot* s = slots[i];
            const Launch& lr = i == 0 ? ln : lside;
            const int64_t Mp = s->dev.Mp, E = 1 + 2 * s->dev.Dp;
            reduce_partials(reduce_buf + rbs[i]->S, (const double*)s->syrk_part.p, (int64_t)K * Mp * Mp, s->nsplit,
                            (int64_t)K * Mp * Mp, false, lr);
            reduce_partials(reduce_buf + rbs[i]->mraw, (const double*)s->mraw_part.p, Mp * KP, s->nsplit, Mp * KP, false, lr);
            reduce_partials(reduce_buf + rbs[i]->esum, (const double*)s->esum_part.p, Mp * E, s->esum_nparts, Mp * E, false, lr);
            // with a communicator attached the buffer leaves this call already summed over the ranks: the header and
            // the pred layer's part go out while the assign layer's partial sums are still being folded
            if (i == 0) TRY(all_reduce_sum(c, reduce_buf, rp.end, c->stream));
        }
        join_side(c);
        TRY(all_reduce_sum(c, reduce_buf + rp.end, ra.end - rp.end, c->stream));
    }
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_elbo_finish(mgp_ctx* c, const mgp_elbo_cfg* cfg, const mgp_layer* pred, const mgp_layer* assign,
                    const double* lik_var, const double* assign_lik_var, const double* reduce_buf, double* elbo,
                    mgp_layer_grad* pg, mgp_layer_grad* ag, double* lik_var_grad, double* assign_lik_var_grad) {
    if (!c) return MGP_ERR_BAD_ARG;
    TRY(check_cfg(c, cfg, pred, assign, lik_var, assign_lik_var));
    if (!reduce_buf || !elbo || !pg || !ag) return fail(c, MGP_ERR_BAD_ARG, "elbo_finish: NULL argument");
    for (const mgp_layer_grad* g : {pg, ag})
        if (!g->Z || !g->q_mu || !g->q_sqrt || !g->variance || !g->lengthscales)
            return fail(c, MGP_ERR_BAD_ARG, "elbo_finish: NULL gradient pointer");
    ON_CTX_DEVICE(c);
    const Launch ln = launch_of(c);
    LayerSlot &sp = c->slot[0], &sa = c->slot[1];
    TRY(setup_layer(c, sp, pred, true));
    TRY(setup_layer(c, sa, assign, true));
    // the precompute held by the context must be THESE layers' (a caller may interleave two models on one context, or
    // call a predict path in between): same dimensions and parameter pointers, else it is formed again here
    const bool same_layers = memcmp(&c->pre_layer[0], pred, sizeof(mgp_layer)) == 0 &&
                             memcmp(&c->pre_layer[1], assign, sizeof(mgp_layer)) == 0;
    const bool from_local = c->pre_valid && same_layers;
    if (!from_local) {
        const Launch ls = fork_side(c);
        precompute_layer(sp.dev, true, (int*)c->status.p, ln);
        precompute_layer(sa.dev, true, (int*)c->status.p, ls);
        join_side(c);
        c->pre_valid = true;
        c->kl_valid = false;
        c->pre_layer[0] = *pred;
        c->pre_layer[1] = *assign;
    } else {
        // same pointers: were the VALUES behind them changed since mgp_elbo_local factorised them?
        FingerprintArgs fa;
        fa.l[0] = *pred; fa.l[1] = *assign;
        fingerprint_kernel<<<2, 256, 0, c->stream>>>(fa, (unsigned long long*)c->fprint.p + 2);
        c->launches += 1;
    }
    const int K = pred->K;
    const LayerRB rp = layer_rb(RB_HEADER, sp.dev.Mp, sp.dev.Dp, K);
    const LayerRB ra = layer_rb(rp.end, sa.dev.Mp, sa.dev.Dp, K);
    const double kl_coef = -1.0 / cfg->num_data;
    double* kl = (double*)c->kl.p;
    Timed t_finish(c, ST_FINISH);
    const Launch ls = fork_side(c);
    finish_layer(sp.dev, reduce_buf + rp.S, reduce_buf + rp.mraw, reduce_buf + rp.esum, reduce_buf + RB_SUMV_PRED, kl_coef,
                 pg->Z, pg->q_mu, pg->q_sqrt, pg->variance, pg->lengthscales, c->kl_valid ? nullptr : kl, ln);
    finish_layer(sa.dev, reduce_buf + ra.S, reduce_buf + ra.mraw, reduce_buf + ra.esum, reduce_buf + RB_SUMV_ASSIGN, kl_coef,
                 ag->Z, ag->q_mu, ag->q_sqrt, ag->variance, ag->lengthscales, c->kl_valid ? nullptr : kl + 1, ls);
    join_side(c);
    elbo_finalize_kernel<<<1, 32, 0, c->stream>>>(reduce_buf, kl, cfg->num_data, K, elbo,
                                                  cfg->lik == MGP_LIK_GAUSSIAN ? lik_var_grad : nullptr,
                                                  cfg->model == MGP_MODEL_SMGP_MODIFIED ? assign_lik_var_grad : nullptr,
                                                  from_local ? (const unsigned long long*)c->fprint.p : nullptr,
                                                  (int*)c->status.p);
    c->launches += 1;
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

int mgp_elbo_fwd_bwd(mgp_ctx* c, const mgp_elbo_cfg* cfg, const mgp_layer* pred, const mgp_layer* assign,
                     const double* lik_var, const double* assign_lik_var, const double* X, const double* Y,
                     int64_t N_local, const mgp_noise* noise, double* elbo, mgp_layer_grad* pg, mgp_layer_grad* ag,
                     double* lik_var_grad, double* assign_lik_var_grad) {
    if (!c) return MGP_ERR_BAD_ARG;
    const int64_t len = mgp_reduce_buffer_len(pred, assign);
    if (len <= 0) return fail(c, MGP_ERR_BAD_ARG, "elbo_fwd_bwd: NULL layer");
    ON_CTX_DEVICE(c);
    TRY(check_cfg(c, cfg, pred, assign, lik_var, assign_lik_var));
    TRY(ensure(c, c->scratch_rb, (size_t)len * 8));
    TRY(mgp_elbo_local(c, cfg, pred, assign, lik_var, assign_lik_var, X, Y, N_local, noise, (double*)c->scratch_rb.p));
    return mgp_elbo_finish(c, cfg, pred, assign, lik_var, assign_lik_var, (const double*)c->scratch_rb.p, elbo, pg, ag,
                           lik_var_grad, assign_lik_var_grad);
}

int mgp_debug_kuu_chol(mgp_ctx* c, const mgp_layer* layer, double* Kuu, double* L, double* Linv) {
    if (!c) return MGP_ERR_BAD_ARG;
    if (!Kuu || !L || !Linv) return fail(c, MGP_ERR_BAD_ARG, "debug_kuu_chol: NULL output");
    ON_CTX_DEVICE(c);
    c->pre_valid = false; c->kl_valid = false;
    TRY(setup_layer(c, c->slot[0], layer, false));
    precompute_layer(c->slot[0].dev, false, (int*)c->status.p, launch_of(c));
    const LayerDev& d = c->slot[0].dev;
    const size_t M = d.M, Mp = d.Mp;
    CUDA_TRY(c, cudaMemcpy2DAsync(Kuu, M * 8, d.Kuu, Mp * 8, M * 8, M, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpy2DAsync(L, M * 8, d.L, Mp * 8, M * 8, M, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(c, cudaMemcpy2DAsync(Linv, M * 8, d.Linv, Mp * 8, M * 8, M, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_TRY(c, cudaGetLastError());
    return MGP_OK;
}

}  // extern "C"
