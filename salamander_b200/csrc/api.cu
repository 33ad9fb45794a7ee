// C ABI of libsalamander_b200.so -- argument checking, workspace ownership, dispatch.
// Declarations and the reference sites each entry point replaces: include/salamander_b200.h
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "sal_common.cuh"

#include <algorithm>
#include <mutex>
#include <vector>

static thread_local char g_err[512] = "";

void sal_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static bool timed_flags(int flags) { return (flags & (SAL_PASS_UPDATE_H | SAL_PASS_WNUM)) == (SAL_PASS_UPDATE_H | SAL_PASS_WNUM); }

int sal_timing_begin(sal_ctx* c, int flags, cudaStream_t st) {
    if (!c->timing || !timed_flags(flags)) return 0;
    if (!c->ev) c->ev = new std::vector<cudaEvent_t>();
    while (c->ev->size() < c->ev_used + 2) {
        cudaEvent_t e;
        SAL_CUDA(cudaEventCreate(&e));
        c->ev->push_back(e);
    }
    SAL_CUDA(cudaEventRecord((*c->ev)[c->ev_used], st));
    return 0;
}

int sal_timing_end(sal_ctx* c, int flags, cudaStream_t st) {
    if (!c->timing || !timed_flags(flags)) return 0;
    SAL_CUDA(cudaEventRecord((*c->ev)[c->ev_used + 1], st));
    c->ev_used += 2;
    return 0;
}

// Per-device free list of workspace scratch buffers.  A fit creates and destroys one handle; cudaMalloc / cudaFree were
// measured at up to 200 ms a call on a busy context (profiles/r01_e2e_phases.md), so released buffers are kept and handed to
// the next handle that asks for the same size (sizes depend on the SM count and dtype only).  sal_trim_scratch() frees them.
struct Scratch {
    int device;
    size_t bytes;
    void* ptr;
    bool in_use;
};
static std::vector<Scratch> g_scratch;
static std::mutex g_scratch_mutex;

static void* scratch_take(int device, size_t bytes) {
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    for (Scratch& b : g_scratch)
        if (!b.in_use && b.device == device && b.bytes == bytes) {
            b.in_use = true;
            return b.ptr;
        }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    g_scratch.push_back({device, bytes, p, true});
    return p;
}

static void scratch_give(int device, void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    for (Scratch& b : g_scratch)
        if (b.ptr == p && b.device == device) b.in_use = false;
}

// Exchange buffers of the period kernel.  They hold sequence-tagged words; the tag counter is one device word per DEVICE,
// shared by every handle and only ever increased (atomicMax by the kernels), and a buffer is zeroed when a handle takes it
// from the free list (it may have held anything before), so a stale word can never carry a tag a later launch waits for.
static unsigned int* g_period_seq[64] = {nullptr};

int sal_period_scratch(sal_ctx* c, cudaStream_t st) {
    if (c->period_partials) return 0;
    const size_t nv = (size_t)SAL_KMAX * SAL_VMAX;
    const size_t b_part = (size_t)2 * c->n_sm * nv * 8, b_sums = (size_t)2 * nv * 8 + 256, b_obj = (size_t)2 * c->n_sm * 16 + 512;
    {
        std::lock_guard<std::mutex> lock(g_scratch_mutex);
        unsigned int*& seq = g_period_seq[c->device & 63];
        if (!seq) {
            SAL_CUDA(cudaMalloc(&seq, 64));
            const unsigned int one = 1;
            SAL_CUDA(cudaMemset(seq, 0, 64));
            SAL_CUDA(cudaMemcpy(seq, &one, sizeof(one), cudaMemcpyHostToDevice));
        }
        c->period_seq = seq;
    }
    void* p = scratch_take(c->device, b_part);
    void* s = scratch_take(c->device, b_sums);
    void* o = scratch_take(c->device, b_obj);
    if (!p || !s || !o) {
        scratch_give(c->device, p), scratch_give(c->device, s), scratch_give(c->device, o);
        sal_set_error("cudaMalloc of the period-kernel exchange buffers failed: %s", cudaGetErrorString(cudaGetLastError()));
        return (int)cudaErrorMemoryAllocation;
    }
    SAL_CUDA(cudaMemsetAsync(p, 0, b_part, st));
    SAL_CUDA(cudaMemsetAsync(s, 0, b_sums, st));
    SAL_CUDA(cudaMemsetAsync(o, 0, b_obj, st));
    c->period_partials = p, c->period_sums = s, c->period_obj = o;
    return 0;
}

bool sal_pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("SAL_B200_NO_PDL");
        return !(e && e[0] == '1');
    }();
    return on;
}

extern "C" {

const char* sal_last_error(void) { return g_err; }
int sal_version(void) { return 100; }

int sal_create(sal_handle_t* out, int V, int64_t D_local, int k, int dtype, int device) {
    SAL_CHECK_ARG(out != nullptr, "out handle is null");
    SAL_CHECK_ARG(V >= 1 && k >= 1 && D_local >= 0, "V, k must be >= 1 and D >= 0");
    SAL_CHECK_ARG(dtype == SAL_F32 || dtype == SAL_F64, "dtype must be SAL_F32 or SAL_F64");
    if (V > SAL_VMAX || k > SAL_KMAX) {
        sal_set_error("unsupported shape: V=%d (max %d), k=%d (max %d)", V, SAL_VMAX, k, SAL_KMAX);
        return SAL_EUNSUPPORTED;
    }
    SAL_CUDA(cudaSetDevice(device));
    int cc_major = 0, n_sm = 0;
    SAL_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
    SAL_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    if (cc_major < 10) {
        sal_set_error("libsalamander_b200 is built for sm_100a only; device %d has compute capability %d.x", device, cc_major);
        return SAL_EUNSUPPORTED;
    }
    sal_ctx* c = new sal_ctx();
    memset(c, 0, sizeof(*c));
    c->V = V, c->k = k, c->KP = sal_kpad(k), c->dtype = dtype, c->device = device, c->D = D_local;
    c->math = SAL_MATH_FMA;
    c->n_sm = n_sm;
    c->grid_pass = c->n_sm * 2;  // the largest persistent grid of any pass flavour (sizes the per-CTA partial buffers)
    const size_t es = dtype == SAL_F32 ? 4 : 8;
    c->partial_wnum = scratch_take(device, (size_t)c->grid_pass * SAL_KMAX * SAL_VMAX * es);
    c->partial_obj = (double*)scratch_take(device, (size_t)c->grid_pass * sizeof(double));
    c->partial_hsum = (double*)scratch_take(device, (size_t)c->grid_pass * SAL_KMAX * sizeof(double));
    if (!c->partial_wnum || !c->partial_obj || !c->partial_hsum) {
        sal_set_error("cudaMalloc of the workspace failed: %s", cudaGetErrorString(cudaGetLastError()));
        scratch_give(device, c->partial_wnum), scratch_give(device, c->partial_obj), scratch_give(device, c->partial_hsum);
        delete c;
        return (int)cudaErrorMemoryAllocation;
    }
    *out = c;
    return 0;
}

int sal_destroy(sal_handle_t h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    // the scratch buffers go back to the per-device free list (a later handle may reuse them on another stream), so
    // everything that used them has to be complete -- the synchronisation cudaFree used to imply
    cudaDeviceSynchronize();
    scratch_give(h->device, h->partial_wnum), scratch_give(h->device, h->partial_obj), scratch_give(h->device, h->partial_hsum);
    scratch_give(h->device, h->period_partials), scratch_give(h->device, h->period_sums), scratch_give(h->device, h->period_obj);
    delete h->map_cache;
    if (h->norm_counter) cudaFree(h->norm_counter);
    if (h->ev) {
        for (cudaEvent_t e : *h->ev) cudaEventDestroy(e);
        delete h->ev;
    }
    delete h;
    return 0;
}

int sal_trim_scratch(void) {
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    for (Scratch& b : g_scratch)
        if (!b.in_use) {
            cudaSetDevice(b.device);
            cudaFree(b.ptr);
            b.ptr = nullptr;
        }
    g_scratch.erase(std::remove_if(g_scratch.begin(), g_scratch.end(), [](const Scratch& b) { return b.ptr == nullptr; }),
                    g_scratch.end());
    return 0;
}

int sal_set_math(sal_handle_t h, int math_mode) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(math_mode == SAL_MATH_FMA || math_mode == SAL_MATH_TF32 || math_mode == SAL_MATH_TF32_ALWAYS, "unknown math mode");
    if (math_mode != SAL_MATH_FMA && h->dtype != SAL_F32) {
        sal_set_error("SAL_MATH_TF32 needs an fp32 handle");
        return SAL_EINVAL;
    }
    h->math = math_mode;
    return 0;
}

int64_t sal_launch_count(sal_handle_t h) { return h ? h->launches : -1; }

int sal_set_timing(sal_handle_t h, int on) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    h->timing = on != 0;
    h->ev_used = 0;
    return 0;
}

int sal_get_pass_timing(sal_handle_t h, double* total_ms, int64_t* n_launches) {
    SAL_CHECK_ARG(h && total_ms && n_launches, "null argument");
    SAL_CUDA(cudaSetDevice(h->device));
    double t = 0.0;
    for (size_t i = 0; i + 1 < h->ev_used; i += 2) {
        float ms = 0.f;
        SAL_CUDA(cudaEventSynchronize((*h->ev)[i + 1]));
        SAL_CUDA(cudaEventElapsedTime(&ms, (*h->ev)[i], (*h->ev)[i + 1]));
        t += ms;
    }
    *total_ms = t, *n_launches = (int64_t)(h->ev_used / 2);
    h->ev_used = 0;
    return 0;
}

int sal_set_debug_buffer(sal_handle_t h, void* buf) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    h->dbg = buf;
    return 0;
}

int sal_klnmf_pass(sal_handle_t h, const void* X, const void* W, const void* H_in, void* H_out,
                   const void* w_kl, const void* w_lhalf, const void* h_scale, int flags, void* Wnum,
                   double* objective, void* per_sample, void* hsum, void* stream) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(W && (h->D == 0 || (X && H_in)), "X, W, H_in must be non-null");
    const int partials_only = flags & SAL_PASS_PARTIALS_ONLY;
    flags &= ~SAL_PASS_PARTIALS_ONLY;
    SAL_CHECK_ARG(flags != 0, "flags == 0: nothing to do");
    SAL_CHECK_ARG(!(flags & SAL_PASS_UPDATE_H) || H_out || h->D == 0, "UPDATE_H needs H_out");
    SAL_CHECK_ARG(!(flags & SAL_PASS_WNUM) || Wnum || partials_only, "WNUM needs Wnum");
    SAL_CHECK_ARG(!(flags & (SAL_PASS_OBJECTIVE | SAL_PASS_POISSON)) || objective, "OBJECTIVE needs objective");
    SAL_CHECK_ARG(!((flags & SAL_PASS_OBJECTIVE) && (flags & SAL_PASS_POISSON)), "OBJECTIVE and POISSON are exclusive");
    SAL_CHECK_ARG(!((flags & SAL_PASS_SAMPLEWISE) && (flags & SAL_PASS_POISSON)), "SAMPLEWISE and POISSON are exclusive");
    SAL_CHECK_ARG(!(flags & SAL_PASS_SAMPLEWISE) || per_sample || h->D == 0, "SAMPLEWISE needs per_sample");
    SAL_CHECK_ARG(!(flags & SAL_PASS_HSUM) || hsum, "HSUM needs hsum");
    SAL_CHECK_ARG(!(flags & SAL_PASS_SCALED_UPDATE) || (h_scale && (flags & SAL_PASS_UPDATE_H) && !w_kl && !w_lhalf && !(flags & SAL_PASS_NOCLIP)),
                  "SCALED_UPDATE needs h_scale and UPDATE_H (no weights, no NOCLIP)");
    SAL_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (h->D == 0) {  // empty shard: outputs are exact zeros
        const size_t es = h->dtype == SAL_F32 ? 4 : 8;
        if (flags & SAL_PASS_WNUM) SAL_CUDA(cudaMemsetAsync(Wnum, 0, (size_t)h->k * h->V * es, st));
        if (flags & (SAL_PASS_OBJECTIVE | SAL_PASS_POISSON)) SAL_CUDA(cudaMemsetAsync(objective, 0, sizeof(double), st));
        if (flags & SAL_PASS_HSUM) SAL_CUDA(cudaMemsetAsync(hsum, 0, (size_t)h->k * es, st));
        return 0;
    }
    PassArgs a;
    a.X = X, a.W = W, a.H_in = H_in, a.w_kl = w_kl, a.w_lhalf = w_lhalf, a.h_scale = h_scale;
    a.H_out = H_out, a.Wnum = Wnum, a.per_sample = per_sample, a.hsum = hsum, a.objective = objective;
    a.flags = flags, a.partials_only = partials_only ? 1 : 0;
    if (h->math != SAL_MATH_FMA && sal_pass_tf32_supported(h, a)) return sal_launch_pass_tf32(h, a, st);
    return sal_launch_pass_fma(h, a, st);
}

int sal_klnmf_update(sal_handle_t h, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out,
                     const void* w_kl, const void* w_lhalf, int n_given, int clip_given, void* Wnum,
                     double* objective, void* stream) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(W_in && W_out && Wnum && (h->D == 0 || (X && H_in && H_out)), "null argument");
    SAL_CHECK_ARG(W_in != W_out, "W_out must not alias W_in (the H step reads the old W)");
    SAL_CHECK_ARG(n_given >= 0 && n_given <= h->k, "n_given out of range");
    SAL_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = h->dtype == SAL_F32 ? 4 : 8;
    if (h->D == 0) {
        SAL_CUDA(cudaMemcpyAsync(W_out, W_in, (size_t)h->k * h->V * es, cudaMemcpyDeviceToDevice, st));
        SAL_CUDA(cudaMemsetAsync(Wnum, 0, (size_t)h->k * h->V * es, st));
        if (objective) SAL_CUDA(cudaMemsetAsync(objective, 0, sizeof(double), st));
        return 0;
    }
    PassArgs a;
    a.X = X, a.W = W_in, a.H_in = H_in, a.w_kl = w_kl, a.w_lhalf = w_lhalf, a.h_scale = nullptr;
    a.H_out = H_out, a.Wnum = Wnum, a.per_sample = nullptr, a.hsum = nullptr, a.objective = objective;
    a.flags = SAL_PASS_UPDATE_H | (n_given < h->k ? SAL_PASS_WNUM : 0) | (objective ? SAL_PASS_OBJECTIVE : 0);
    a.fuse_epilogue = n_given < h->k, a.n_given = n_given, a.clip_given = clip_given, a.W_out = W_out;
    if (n_given >= h->k) SAL_CUDA(cudaMemcpyAsync(W_out, W_in, (size_t)h->k * h->V * es, cudaMemcpyDeviceToDevice, st));
    if (h->math != SAL_MATH_FMA && sal_pass_tf32_supported(h, a)) return sal_launch_pass_tf32(h, a, st);
    return sal_launch_pass_fma(h, a, st);
}

int sal_klnmf_update_p2p(sal_handle_t h, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out,
                         const void* w_kl, const void* w_lhalf, int n_given, int clip_given, void* Wnum, double* objective,
                         const void* peer_buffers, void* p2p_state, int n_ranks, int rank, void* stream) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(W_in && W_out && Wnum && (h->D == 0 || (X && H_in && H_out)), "null argument");
    SAL_CHECK_ARG(W_in != W_out, "W_out must not alias W_in (the H step reads the old W)");
    SAL_CHECK_ARG(n_given >= 0 && n_given < h->k, "n_given out of range (all signatures given: use sal_klnmf_pass)");
    SAL_CHECK_ARG(peer_buffers && p2p_state && n_ranks >= 1 && n_ranks <= SAL_VMAX && rank >= 0 && rank < n_ranks, "bad exchange arguments");
    SAL_CHECK_ARG(h->D > 0, "every rank needs at least one sample");
    SAL_CUDA(cudaSetDevice(h->device));
    PassArgs a;
    a.X = X, a.W = W_in, a.H_in = H_in, a.w_kl = w_kl, a.w_lhalf = w_lhalf, a.h_scale = nullptr;
    a.H_out = H_out, a.Wnum = Wnum, a.per_sample = nullptr, a.hsum = nullptr, a.objective = objective;
    a.flags = SAL_PASS_UPDATE_H | SAL_PASS_WNUM | (objective ? SAL_PASS_OBJECTIVE : 0);
    a.fuse_epilogue = 1, a.n_given = n_given, a.clip_given = clip_given, a.W_out = W_out;
    a.p2p_peers = peer_buffers, a.p2p_state = p2p_state, a.p2p_n_ranks = n_ranks, a.p2p_rank = rank;
    cudaStream_t st = (cudaStream_t)stream;
    if (h->math != SAL_MATH_FMA && sal_pass_tf32_supported(h, a)) return sal_launch_pass_tf32(h, a, st);
    return sal_launch_pass_fma(h, a, st);
}

static int period_args(sal_handle_t h, PeriodArgs& a, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out,
                       int n_given, int clip_given, int n_updates, int objective_every, int final_objective, double* objectives,
                       const void* peer_buffers, void* p2p_state, int n_ranks, int rank) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(X && W_in && H_in, "X, W_in, H_in must be non-null");
    SAL_CHECK_ARG(n_updates >= 0 && objective_every >= 0 && (n_updates > 0 || final_objective), "nothing to do");
    SAL_CHECK_ARG(n_updates == 0 || (W_out && H_out), "updates need W_out and H_out");
    SAL_CHECK_ARG(n_given >= 0 && n_given <= h->k, "n_given out of range");
    SAL_CHECK_ARG((objective_every == 0 && !final_objective) || objectives, "objectives requested but the output array is null");
    SAL_CHECK_ARG(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "bad rank / n_ranks");
    SAL_CHECK_ARG(n_ranks == 1 || (peer_buffers && p2p_state), "several ranks need the peer table and the shared tag counter");
    a.X = X, a.W_in = W_in, a.H_in = H_in, a.W_out = W_out, a.H_out = H_out ? H_out : const_cast<void*>(H_in);
    a.objectives = objectives, a.peers = n_ranks > 1 ? peer_buffers : nullptr, a.state = n_ranks > 1 ? p2p_state : nullptr;
    a.n_updates = n_updates, a.obj_every = objective_every, a.final_obj = final_objective ? 1 : 0;
    a.n_given = n_given, a.clip_given = clip_given, a.n_ranks = n_ranks, a.rank = rank;
    if (!sal_period_supported(h, a)) {
        sal_set_error("sal_klnmf_period: needs an fp32 handle in tf32 mode, V = 96, k <= 32, D_local >= %d (or TF32_ALWAYS), 16-byte aligned "
                      "X / H, at least one signature to update and at most 8 ranks", (int)SAL_TF32_MIN_SAMPLES);
        return SAL_EUNSUPPORTED;
    }
    return 0;
}

int sal_klnmf_period_supported(sal_handle_t h, int n_given, int n_ranks) {
    if (!h) return 0;
    PeriodArgs a = {};
    a.n_given = n_given, a.n_ranks = n_ranks, a.n_updates = 1;
    return sal_period_supported(h, a) ? 1 : 0;
}

int sal_klnmf_period(sal_handle_t h, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, int n_given,
                     int clip_given, int n_updates, int objective_every, int final_objective, double* objectives,
                     const void* peer_buffers, void* p2p_state, int n_ranks, int rank, void* stream) {
    PeriodArgs a;
    if (int e = period_args(h, a, X, W_in, W_out, H_in, H_out, n_given, clip_given, n_updates, objective_every, final_objective,
                            objectives, peer_buffers, p2p_state, n_ranks, rank))
        return e;
    SAL_CUDA(cudaSetDevice(h->device));
    h->x_dirty = 0;
    sal_ctx* cs[1] = {h};
    return sal_launch_period(cs, 1, &a, (cudaStream_t)stream);
}

int sal_klnmf_period_emulated(const sal_handle_t* hs, int n_virtual, const void* const* X, const void* const* W_in, void* const* W_out,
                              const void* const* H_in, void* const* H_out, int n_given, int clip_given, int n_updates,
                              int objective_every, int final_objective, double* const* objectives,
                              const void* const* peer_tables, void* const* states, void* stream) {
    SAL_CHECK_ARG(hs && n_virtual >= 1 && n_virtual <= 2, "1 or 2 emulated ranks");
    SAL_CHECK_ARG(X && W_in && W_out && H_in && H_out && objectives && peer_tables && states, "null argument array");
    PeriodArgs as[2];
    sal_ctx* cs[2];
    for (int v = 0; v < n_virtual; ++v) {
        if (int e = period_args(hs[v], as[v], X[v], W_in[v], W_out[v], H_in[v], H_out[v], n_given, clip_given, n_updates,
                                objective_every, final_objective, objectives[v], peer_tables[v], states[v], n_virtual, v))
            return e;
        cs[v] = hs[v];
        SAL_CHECK_ARG(hs[v]->device == hs[0]->device && hs[v]->k == hs[0]->k && hs[v]->math == hs[0]->math, "emulated ranks must share device, k and math mode");
    }
    SAL_CUDA(cudaSetDevice(hs[0]->device));
    return sal_launch_period(cs, n_virtual, as, (cudaStream_t)stream);
}

size_t sal_p2p_exchange_bytes(int k, int n_ranks) { return (size_t)2 * n_ranks * (k + 1) * SAL_VMAX * 16; }

int sal_klnmf_small_supported(sal_handle_t h) { return h && sal_small_supported(h) ? 1 : 0; }

int sal_klnmf_small_updates(sal_handle_t h, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out,
                            int n_given, int n_iterations, double* objective, void* stream) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(X && W_in && W_out && H_in && H_out, "null argument");
    SAL_CHECK_ARG(n_given >= 0 && n_given <= h->k && n_iterations >= 0, "n_given / n_iterations out of range");
    if (!sal_small_supported(h)) {
        sal_set_error("sal_klnmf_small_updates: problem does not fit one CTA (D <= 256 and ~220 KB of shared memory)");
        return SAL_EUNSUPPORTED;
    }
    SAL_CUDA(cudaSetDevice(h->device));
    return sal_launch_klnmf_small(h, X, W_in, W_out, H_in, H_out, n_given, n_iterations, objective, (cudaStream_t)stream);
}

int sal_mvnmf_small_supported(sal_handle_t h) { return h && sal_mvnmf_small_ok(h) ? 1 : 0; }

int sal_mvnmf_small_updates(sal_handle_t h, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, double lam,
                            double delta, int n_given, int n_iterations, const double* gamma_in, double* gamma_out,
                            double* objective, void* stream) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(X && W_in && W_out && H_in && H_out && gamma_in && gamma_out, "null argument");
    SAL_CHECK_ARG(n_given >= 0 && n_given <= h->k && n_iterations >= 0, "n_given / n_iterations out of range");
    SAL_CHECK_ARG(lam > 0.0, "lam must be positive");
    if (!sal_mvnmf_small_ok(h)) {
        sal_set_error("sal_mvnmf_small_updates: problem does not fit one CTA (D <= 256 and ~220 KB of shared memory)");
        return SAL_EUNSUPPORTED;
    }
    SAL_CUDA(cudaSetDevice(h->device));
    return sal_launch_mvnmf_small(h, X, W_in, W_out, H_in, H_out, lam, delta, n_given, n_iterations, gamma_in, gamma_out, objective,
                                  (cudaStream_t)stream);
}

int sal_w_epilogue(sal_handle_t h, const void* W_in, const void* Wnum, int n_given, int clip_given,
                   void* W_out, void* stream) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(W_in && W_out, "W_in, W_out must be non-null");
    SAL_CHECK_ARG(n_given >= 0 && n_given <= h->k, "n_given out of range");
    SAL_CHECK_ARG(n_given == h->k || Wnum, "Wnum must be non-null");
    SAL_CUDA(cudaSetDevice(h->device));
    return sal_launch_w_epilogue(h, W_in, Wnum, n_given, clip_given, W_out, (cudaStream_t)stream);
}

int sal_mark_counts_written(sal_handle_t h) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    h->x_dirty = 1;
    return 0;
}

int sal_clip_counts(sal_handle_t h, void* X, int64_t n, long long* n_changed, void* stream) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(n >= 0 && (n == 0 || (X && n_changed)), "X / n_changed must be non-null");
    if (n == 0) return 0;
    SAL_CUDA(cudaSetDevice(h->device));
    h->x_dirty = 1;  // the next tensor-core pass of this handle is launched in plain stream order (it reads X early otherwise)
    return sal_launch_clip_counts(h, X, n, n_changed, (cudaStream_t)stream);
}

int sal_scale_clip_rows(sal_handle_t h, void* H, const void* scale, void* stream) {
    SAL_CHECK_ARG(h && scale && (h->D == 0 || H), "null argument");
    if (h->D == 0) return 0;
    SAL_CUDA(cudaSetDevice(h->device));
    return sal_launch_scale_clip_rows(h, H, scale, (cudaStream_t)stream);
}

int sal_mvnmf_logdet(sal_handle_t h, const void* W, double delta, double* out, void* stream) {
    SAL_CHECK_ARG(h && W && out, "null argument");
    SAL_CUDA(cudaSetDevice(h->device));
    return sal_launch_mvnmf_logdet(h, W, delta, out, (cudaStream_t)stream);
}

int sal_mvnmf_w_unconstrained(sal_handle_t h, const void* W, const void* N, const void* hsum, double lam,
                              double delta, int n_given, void* W_unc, void* stream) {
    SAL_CHECK_ARG(h && W && N && hsum && W_unc, "null argument");
    SAL_CHECK_ARG(n_given >= 0 && n_given <= h->k, "n_given out of range");
    SAL_CUDA(cudaSetDevice(h->device));
    return sal_launch_mvnmf_w_unc(h, W, N, hsum, lam, delta, n_given, W_unc, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

int sal_mvnmf_w_unconstrained_trial(sal_handle_t h, const void* W, const void* N, const void* hsum, double lam, double delta,
                                    int n_given, void* W_unc, void* W_trial, void* h_scale, double* logdet_out, void* stream) {
    SAL_CHECK_ARG(h && W && N && hsum && W_unc && W_trial && h_scale && logdet_out, "null argument");
    SAL_CHECK_ARG(n_given >= 0 && n_given <= h->k, "n_given out of range");
    SAL_CUDA(cudaSetDevice(h->device));
    return sal_launch_mvnmf_w_unc(h, W, N, hsum, lam, delta, n_given, W_unc, W_trial, h_scale, logdet_out, (cudaStream_t)stream);
}

int sal_mvnmf_trial(sal_handle_t h, const void* W, const void* W_unc, double gamma_blend, double delta,
                    void* W_trial, void* h_scale, double* logdet_out, void* stream) {
    SAL_CHECK_ARG(h && W && W_unc && W_trial && h_scale && logdet_out, "null argument");
    SAL_CUDA(cudaSetDevice(h->device));
    return sal_launch_mvnmf_trial(h, W, W_unc, gamma_blend, delta, W_trial, h_scale, logdet_out,
                                  (cudaStream_t)stream);
}

#define SAL_CORR_COMMON(m)                                                        \
    SAL_CHECK_ARG(h != nullptr, "handle is null");                                \
    SAL_CHECK_ARG((m) >= 1, "dim_embeddings must be >= 1");                       \
    if ((m) > sal_corrnmf_max_dim()) {                                            \
        sal_set_error("dim_embeddings %d exceeds the supported %d", (m), sal_corrnmf_max_dim()); \
        return SAL_EUNSUPPORTED;                                                  \
    }                                                                             \
    SAL_CUDA(cudaSetDevice(h->device));

int sal_row_sums(sal_handle_t h, const void* X, void* out, void* stream) {
    SAL_CHECK_ARG(h && (h->D == 0 || (X && out)), "null argument");
    SAL_CUDA(cudaSetDevice(h->device));
    return sal_launch_row_sums(h, X, out, (cudaStream_t)stream);
}

int sal_corrnmf_exposures(sal_handle_t h, const void* a, const void* b, const void* L, const void* U, int m, void* H, void* stream) {
    SAL_CORR_COMMON(m);
    SAL_CHECK_ARG(a && L && (h->D == 0 || (b && U && H)), "null argument");
    return sal_launch_corrnmf_exposures(h, a, b, L, U, m, H, (cudaStream_t)stream);
}

int sal_corrnmf_sample_scalings(sal_handle_t h, const void* xsum, const void* a, const void* L, const void* U, int m, void* b_out,
                                void* stream) {
    SAL_CORR_COMMON(m);
    SAL_CHECK_ARG(a && L && (h->D == 0 || (xsum && U && b_out)), "null argument");
    return sal_launch_corrnmf_sample_scalings(h, xsum, a, L, U, m, b_out, (cudaStream_t)stream);
}

int sal_corrnmf_signature_scalings_sums(sal_handle_t h, const void* auxT, const void* b, const void* L, const void* U, int m,
                                        double* sums, void* stream) {
    SAL_CORR_COMMON(m);
    SAL_CHECK_ARG(L && sums && (h->D == 0 || (auxT && b && U)), "null argument");
    return sal_launch_corrnmf_signature_scalings_sums(h, auxT, b, L, U, m, sums, (cudaStream_t)stream);
}

int sal_corrnmf_signature_scalings_finish(sal_handle_t h, const double* sums, void* a_out, void* stream) {
    SAL_CHECK_ARG(h && sums && a_out, "null argument");
    SAL_CUDA(cudaSetDevice(h->device));
    return sal_launch_corrnmf_signature_scalings_finish(h, sums, a_out, (cudaStream_t)stream);
}

int sal_corrnmf_sample_embeddings(sal_handle_t h, const void* auxT, const void* a, const void* b, const void* L, void* U, int m,
                                  double variance, int maxiter, void* stream) {
    SAL_CORR_COMMON(m);
    SAL_CHECK_ARG(a && L && (h->D == 0 || (auxT && b && U)), "null argument");
    SAL_CHECK_ARG(variance > 0.0 && maxiter >= 0, "variance must be positive, maxiter >= 0");
    return sal_launch_corrnmf_sample_embeddings(h, auxT, a, b, 0, L, U, m, variance, maxiter, (cudaStream_t)stream);
}

int sal_corrnmf_sample_embeddings_mm(sal_handle_t h, const void* auxT, const void* a, const void* b_mat, const void* L, void* U, int m,
                                     double variance, int maxiter, void* stream) {
    SAL_CORR_COMMON(m);
    SAL_CHECK_ARG(a && L && (h->D == 0 || (auxT && b_mat && U)), "null argument");
    SAL_CHECK_ARG(variance > 0.0 && maxiter >= 0, "variance must be positive, maxiter >= 0");
    return sal_launch_corrnmf_sample_embeddings(h, auxT, a, b_mat, 1, L, U, m, variance, maxiter, (cudaStream_t)stream);
}

int sal_corrnmf_signature_embeddings(sal_handle_t h, const void* auxT, const void* a, const void* b, void* L, const void* U, int m,
                                     double variance, void* stream) {
    SAL_CORR_COMMON(m);
    SAL_CHECK_ARG(a && L && (h->D == 0 || (auxT && b && U)), "null argument");
    SAL_CHECK_ARG(variance > 0.0, "variance must be positive");
    return sal_launch_corrnmf_signature_embeddings(h, auxT, a, b, L, U, m, variance, 0, h->k, (cudaStream_t)stream);
}

int sal_corrnmf_signature_embeddings_range(sal_handle_t h, const void* auxT, const void* a, const void* b, void* L, const void* U,
                                           int m, double variance, int sig_begin, int sig_count, void* stream) {
    SAL_CORR_COMMON(m);
    SAL_CHECK_ARG(a && L && (h->D == 0 || (auxT && b && U)), "null argument");
    SAL_CHECK_ARG(variance > 0.0, "variance must be positive");
    SAL_CHECK_ARG(sig_begin >= 0 && sig_count >= 0 && sig_begin + sig_count <= h->k, "signature range out of bounds");
    return sal_launch_corrnmf_signature_embeddings(h, auxT, a, b, L, U, m, variance, sig_begin, sig_count, (cudaStream_t)stream);
}

size_t sal_corrnmf_sig_exchange_bytes(int k, int n_ranks) { return sal_corrnmf_sig_exchange_words(k, n_ranks) * 16; }

int sal_corrnmf_signature_embeddings_p2p(sal_handle_t h, const void* auxT, const void* a, const void* b, void* L, const void* U, int m,
                                         double variance, const void* peers, int n_ranks, int rank, unsigned int launch_id,
                                         void* stream) {
    SAL_CORR_COMMON(m);
    SAL_CHECK_ARG(a && L && (h->D == 0 || (auxT && b && U)), "null argument");
    SAL_CHECK_ARG(variance > 0.0, "variance must be positive");
    SAL_CHECK_ARG(n_ranks >= 1 && n_ranks <= 8 && rank >= 0 && rank < n_ranks, "1 .. 8 ranks, 0 <= rank < n_ranks");
    SAL_CHECK_ARG(n_ranks == 1 || peers, "peers is null");
    SAL_CHECK_ARG(launch_id >= 1 && launch_id < (1u << 15), "launch_id must be in 1 .. 32767 (zero the receive buffers and start over)");
    const SigLaunchRank r = {h, auxT, a, b, U, L, peers, rank};
    return sal_launch_corrnmf_signature_embeddings_v(&r, 1, m, variance, 0, h->k, n_ranks, launch_id, (cudaStream_t)stream);
}

int sal_corrnmf_signature_embeddings_emulated(const sal_handle_t* hs, int n_virtual, const void* const* auxT, const void* const* a,
                                              const void* const* b, void* const* L, const void* const* U, int m, double variance,
                                              const void* const* peer_tables, unsigned int launch_id, void* stream) {
    SAL_CHECK_ARG(hs && n_virtual >= 1 && n_virtual <= 2, "1 or 2 emulated ranks");
    SAL_CHECK_ARG(auxT && a && b && L && U && peer_tables, "null argument array");
    sal_handle_t h = hs[0];
    SAL_CORR_COMMON(m);
    SAL_CHECK_ARG(variance > 0.0, "variance must be positive");
    SAL_CHECK_ARG(launch_id >= 1 && launch_id < (1u << 15), "launch_id must be in 1 .. 32767");
    SigLaunchRank rs[2];
    for (int v = 0; v < n_virtual; ++v) {
        SAL_CHECK_ARG(hs[v] && hs[v]->device == h->device && hs[v]->k == h->k && hs[v]->dtype == h->dtype, "emulated ranks must share device, k and dtype");
        SAL_CHECK_ARG(a[v] && L[v] && peer_tables[v] && (hs[v]->D == 0 || (auxT[v] && b[v] && U[v])), "null argument");
        rs[v] = {hs[v], auxT[v], a[v], b[v], U[v], L[v], peer_tables[v], v};
    }
    return sal_launch_corrnmf_signature_embeddings_v(rs, n_virtual, m, variance, 0, h->k, n_virtual, launch_id, (cudaStream_t)stream);
}

size_t sal_p2p_allreduce_bytes(int n_ranks) { return sal_p2p_allreduce_words(n_ranks) * 16; }
int sal_p2p_allreduce_max_values(void) { return sal_p2p_allreduce_max(); }

int sal_p2p_allreduce_f64(sal_handle_t h, double* values, int n, const void* peers, int n_ranks, int rank, unsigned int launch_id,
                          void* stream) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(values && peers && n >= 0, "null argument");
    SAL_CHECK_ARG(n_ranks >= 2 && n_ranks <= 8 && rank >= 0 && rank < n_ranks, "2 .. 8 ranks, 0 <= rank < n_ranks");
    SAL_CHECK_ARG(launch_id >= 1 && launch_id < (1u << 15), "launch_id must be in 1 .. 32767 (zero the receive buffers and start over)");
    SAL_CUDA(cudaSetDevice(h->device));
    h->launches++;
    return sal_launch_p2p_allreduce(1, &values, &peers, &rank, n, n_ranks, launch_id, (cudaStream_t)stream);
}

int sal_p2p_allreduce_f64_emulated(sal_handle_t h, double* const* values, int n, const void* const* peer_tables, int n_virtual,
                                   unsigned int launch_id, void* stream) {
    SAL_CHECK_ARG(h != nullptr, "handle is null");
    SAL_CHECK_ARG(values && peer_tables && n >= 0 && n_virtual == 2, "two emulated ranks");
    SAL_CHECK_ARG(launch_id >= 1 && launch_id < (1u << 15), "launch_id must be in 1 .. 32767");
    SAL_CUDA(cudaSetDevice(h->device));
    const int gpus[2] = {0, 1};
    return sal_launch_p2p_allreduce(n_virtual, values, peer_tables, gpus, n, n_virtual, launch_id, (cudaStream_t)stream);
}

int sal_corrnmf_norms(sal_handle_t h, const void* L, const void* U, int m, const void* X_or_null, double* out, void* stream) {
    SAL_CORR_COMMON(m);
    SAL_CHECK_ARG(L && out, "null argument");
    return sal_launch_corrnmf_norms(h, L, U, m, X_or_null, out, (cudaStream_t)stream);
}

}  // extern "C"
