// Shared declarations of libsalamander_b200 (sm_100a).  See include/salamander_b200.h for the ABI.
#pragma once

#include <cuda_runtime.h>

#include <utility>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "../../include/salamander_b200.h"

#define SAL_VMAX 96   // features per sample handled by the kernels (SBS-96 / ID-83 / SV-32)
#define SAL_KMAX 32   // signatures handled by the kernels
#define SAL_EPS_F32 1.1920928955078125e-07  // np.finfo(np.float32).eps, the reference's clip constant

struct sal_ctx {
    int V, k, KP, dtype, device, math;
    int64_t D;
    int n_sm;
    int grid_pass;         // persistent grid of the fused pass
    void* partial_wnum;    // [grid_pass][KP][SAL_VMAX] real
    double* partial_obj;   // [grid_pass]
    double* partial_hsum;  // [grid_pass][SAL_KMAX]
    void* dbg;             // optional diagnostics buffer of the tensor-core pass (sal_set_debug_buffer)
    int64_t launches;
    unsigned int* norm_counter;                       // ticket counters of the CorrNMF norms reduction
    int timing;                                       // sal_set_timing
    std::vector<cudaEvent_t>* ev;                     // event pairs around the UPDATE_H | WNUM pass kernels
    size_t ev_used;
    // period kernel (klnmf_period_tf32.cu): sequence-tagged exchange buffers, taken from the scratch list at first use
    void* period_partials;      // uint2 [2][n_sm][SAL_KMAX * SAL_VMAX]
    void* period_sums;          // uint2 [2][SAL_KMAX * SAL_VMAX]
    void* period_obj;           // uint4 [2][n_sm]
    unsigned int* period_seq;   // per-DEVICE tag counter (shared by all handles of the device, only ever increases)
    std::vector<struct sal_map_entry>* map_cache;     // tensor maps by (pointer, kind): encoding one costs ~10 us of host time
    int x_dirty;                // X was written by a kernel of this handle (sal_clip_counts) and no pass has run since
};

struct sal_map_entry {
    const void* ptr;
    int kind;  // 0: X [D][96] swizzled boxes; 1: H [D][k] rows (2-D) ; 2: H as [tile][k][128] (3-D, k % 4 != 0)
    alignas(64) unsigned char map[128];
};
// Tensor map of kind `kind` over `ptr` for this handle's (D, k), from the handle's cache (encoded on a miss).
int sal_cached_map(sal_ctx* c, void* map_out, const void* ptr, int kind);
// exchange buffers + tag counter of the period kernel, zeroed on the launch stream at first use
int sal_period_scratch(sal_ctx* c, cudaStream_t st);

// bracket a pass kernel with events when timing is on (no-ops otherwise)
int sal_timing_begin(sal_ctx* c, int flags, cudaStream_t st);
int sal_timing_end(sal_ctx* c, int flags, cudaStream_t st);

void sal_set_error(const char* fmt, ...);

#define SAL_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            sal_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (int)e__;                                                                 \
        }                                                                                    \
    } while (0)

#define SAL_CHECK_ARG(cond, msg)            \
    do {                                    \
        if (!(cond)) {                      \
            sal_set_error("invalid argument: %s", msg); \
            return SAL_EINVAL;              \
        }                                   \
    } while (0)

static inline int sal_kpad(int k) { return k <= 8 ? 8 : k <= 16 ? 16 : k <= 24 ? 24 : 32; }

// ---- launchers implemented in the .cu files -------------------------------------------------
// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// The update is a chain pass -> finish -> pass -> ...; each kernel's set-up (barrier init, TMEM allocation, descriptor
// prefetch, block scheduling) does not depend on its predecessor.  Kernels launched through sal_launch_pdl may start
// while the previous kernel in the stream is still running; they MUST execute pdl_wait() before touching anything a
// predecessor wrote (it returns once the predecessor grid has completed and its writes are visible) and call
// pdl_trigger() early so that their own successor can be scheduled.  SAL_B200_NO_PDL=1 turns the attribute off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool sal_pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t sal_launch_pdl_if(bool overlap, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr, cfg.numAttrs = (overlap && sal_pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t sal_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    return sal_launch_pdl_if(true, kernel, grid, block, smem, st, std::forward<Args>(args)...);
}
#endif

struct PassArgs {
    const void *X, *W, *H_in, *w_kl, *w_lhalf, *h_scale;
    void *H_out, *Wnum, *per_sample, *hsum;
    double* objective;
    int flags;
    // optional W epilogue fused into the reduction of the partials (sal_klnmf_update)
    int fuse_epilogue = 0, n_given = 0, clip_given = 0;
    void* W_out = nullptr;
    // multi-GPU one-shot all-reduce fused into the reduction (sal_klnmf_update_p2p)
    const void* p2p_peers = nullptr;  // device array [n_ranks] of exchange-buffer pointers
    void* p2p_state = nullptr;        // device unsigned[2]: {sequence number (starts at 1), ticket (0)}
    int p2p_n_ranks = 0, p2p_rank = 0;
    int partials_only = 0;  // SAL_PASS_PARTIALS_ONLY: skip the reduction kernel
};
int sal_launch_pass_fma(sal_ctx* c, const PassArgs& a, cudaStream_t st);
int sal_launch_pass_tf32(sal_ctx* c, const PassArgs& a, cudaStream_t st);  // tcgen05 path (fp32 only)
bool sal_pass_tf32_supported(const sal_ctx* c, const PassArgs& a);         // shapes / flags the tcgen05 path covers
int sal_launch_pass_reduce(sal_ctx* c, const PassArgs& a, int n_part, cudaStream_t st);
int sal_pass_smem_bytes(int dtype, int KP);
int sal_launch_w_epilogue(sal_ctx* c, const void* W_in, const void* Wnum, int n_given, int clip_given,
                          void* W_out, cudaStream_t st);
int sal_launch_clip_counts(sal_ctx* c, void* X, int64_t n, long long* n_changed, cudaStream_t st);
int sal_launch_scale_clip_rows(sal_ctx* c, void* H, const void* scale, cudaStream_t st);
int sal_launch_mvnmf_logdet(sal_ctx* c, const void* W, double delta, double* out, cudaStream_t st);
int sal_launch_mvnmf_w_unc(sal_ctx* c, const void* W, const void* N, const void* hsum, double lam,
                           double delta, int n_given, void* W_unc, void* W_trial, void* h_scale, double* logdet_out, cudaStream_t st);
int sal_launch_mvnmf_trial(sal_ctx* c, const void* W, const void* W_unc, double gamma, double delta,
                           void* W_trial, void* h_scale, double* logdet_out, cudaStream_t st);

// ---- correlated NMF launchers (corrnmf.cu) --------------------------------------------------------------
int sal_launch_corrnmf_exposures(sal_ctx* c, const void* a, const void* b, const void* L, const void* U, int m, void* H, cudaStream_t st);
int sal_launch_row_sums(sal_ctx* c, const void* X, void* out, cudaStream_t st);
int sal_launch_corrnmf_sample_scalings(sal_ctx* c, const void* xsum, const void* a, const void* L, const void* U, int m, void* b,
                                       cudaStream_t st);
int sal_launch_corrnmf_signature_scalings_sums(sal_ctx* c, const void* auxT, const void* b, const void* L, const void* U, int m,
                                               double* sums, cudaStream_t st);
int sal_launch_corrnmf_signature_scalings_finish(sal_ctx* c, const double* sums, void* a, cudaStream_t st);
int sal_launch_corrnmf_sample_embeddings(sal_ctx* c, const void* auxT, const void* a, const void* b, int b_is_matrix, const void* L,
                                         void* U, int m, double variance, int maxiter, cudaStream_t st);
int sal_launch_corrnmf_signature_embeddings(sal_ctx* c, const void* auxT, const void* a, const void* b, void* L, const void* U, int m,
                                            double variance, int sig_begin, int sig_count, cudaStream_t st);
// (several GPUs / emulated ranks: one entry per rank served by this launch; peers = device array [n_gpus] of receive buffers)
struct SigLaunchRank {
    sal_ctx* c;
    const void *auxT, *a, *b, *U;
    void* L;
    const void* peers;
    int gpu;
};
int sal_launch_corrnmf_signature_embeddings_v(const SigLaunchRank* rs, int n_virtual, int m, double variance, int sig_begin, int sig_count,
                                              int n_gpus, unsigned int launch_id, cudaStream_t st);
size_t sal_corrnmf_sig_exchange_words(int k, int n_gpus);  // 16-byte words of one receive buffer
// small all-reduce of doubles over peer memory (corrnmf_sig.cu)
size_t sal_p2p_allreduce_words(int n_gpus);
int sal_p2p_allreduce_max(void);
int sal_launch_p2p_allreduce(int n_virtual, double* const* bufs, const void* const* peer_tables, const int* gpus, int n, int n_gpus,
                             unsigned int launch_id, cudaStream_t st);
int sal_launch_corrnmf_norms(sal_ctx* c, const void* L, const void* U, int m, const void* X_or_null, double* out, cudaStream_t st);

// ---- persistent period kernel (klnmf_period_tf32.cu) ------------------------------------------------------
struct PeriodArgs {
    const void *X, *W_in, *H_in;
    void *W_out, *H_out;
    double* objectives;   // [ceil(n_updates / obj_every) + final_obj]
    const void* peers;    // device array [n_ranks] of receive buffers, or null (single rank)
    void* state;          // device unsigned: tag counter shared with the peers, or null (the device's own counter)
    int n_updates, obj_every, final_obj, n_given, clip_given, n_ranks, rank;
};
bool sal_period_supported(const sal_ctx* c, const PeriodArgs& a);
int sal_launch_period(sal_ctx* const* cs, int n_virtual, const PeriodArgs* as, cudaStream_t st);

// ---- small-problem MvNMF (mvnmf_small.cu) ------------------------------------------------------------------
bool sal_mvnmf_small_ok(const sal_ctx* c);
int sal_launch_mvnmf_small(sal_ctx* c, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, double lam,
                           double delta, int n_given, int n_iter, const double* gamma_in, double* gamma_out, double* objective,
                           cudaStream_t st);

// ---- small-problem persistent kernel (klnmf_small.cu) ----------------------------------------------------
bool sal_small_supported(const sal_ctx* c);
int sal_launch_klnmf_small(sal_ctx* c, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, int n_given,
                           int n_iter, double* objective, cudaStream_t st);
