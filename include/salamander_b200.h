/*
 * salamander_b200.h -- C ABI of libsalamander_b200.so (sm_100a).
 *
 * The reference (parklab/Salamander v0.4.2) has no FFI layer: its compiled code is the
 * numba-JIT output of 17 @njit functions.  Each entry point below
 * replaces one or more of those functions; the reference site is cited per function.
 * A maintainer binds them with ctypes.
 *
 * Conventions
 *   - every function returns int: 0 ok, <0 invalid argument (SAL_E*), >0 a cudaError_t;
 *     sal_last_error() returns a thread-local message.  Nothing throws across the ABI.
 *   - all array arguments are DEVICE pointers unless the name ends in _host;
 *     element type `real` is float (dtype SAL_F32) or double (SAL_F64), fixed per handle.
 *   - layouts are the AnnData memory of the reference:
 *        X [D][V]  counts, sample-major  (adata.X,            signature_nmf.py:281)
 *        H [D][k]  exposures             (adata.obsm["exposures"], klnmf.py:106)
 *        W [k][V]  signatures            (asignatures.X,      klnmf.py:105)
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*); scalar
 *     outputs are device doubles.  A handle may be used from one stream at a time.
 *   - D is the number of samples LOCAL to this GPU (sample-sharded data parallel).
 */
#ifndef SALAMANDER_B200_H
#define SALAMANDER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sal_ctx* sal_handle_t;

enum { SAL_F32 = 0, SAL_F64 = 1 };

/* arithmetic used for the three thin contractions of the fused pass (fp32 handles only) */
enum {
    SAL_MATH_FMA = 0,  /* CUDA-core FMA in the handle's dtype (exact fp32 / fp64)          */
    SAL_MATH_TF32 = 1, /* tcgen05 kind::tf32 tensor-core contractions, fp32 accumulation:
                          used by sal_klnmf_pass when V == 96, k <= 32 and flags within
                          UPDATE_H | WNUM | OBJECTIVE and D_local >= SAL_TF32_MIN_SAMPLES (below that the pass is
                          latency bound and the tf32 noise of the numerator no longer averages out over samples);
                          every other call runs the exact FMA kernels (still on the GPU)                      */
    SAL_MATH_TF32_ALWAYS = 2 /* as SAL_MATH_TF32 without the D_local threshold (used by the parity tests)     */
};
enum { SAL_TF32_MIN_SAMPLES = 4096 };

enum {
    SAL_EINVAL = -1,      /* null pointer / bad size / bad enum            */
    SAL_EUNSUPPORTED = -2 /* V > 96 or k > 32 (not yet covered by kernels)  */
};

/* flags of sal_klnmf_pass */
enum {
    SAL_PASS_UPDATE_H = 1,   /* write H_out = clip(H * W^T A)  (or the l-half closed form)   */
    SAL_PASS_WNUM = 2,       /* write Wnum[k][V] = sum_d wkl_d A[v,d] H[k,d]  (RAW, no W *)  */
    SAL_PASS_OBJECTIVE = 4,  /* write *objective = KL(X||WH) (+ l-half term) of the INPUT W,H */
    SAL_PASS_SAMPLEWISE = 8, /* write per_sample[d] = unweighted KL of sample d              */
    SAL_PASS_HSUM = 16,      /* write hsum[k] = sum_d H_in[d][k]                             */
    SAL_PASS_POISSON = 32,   /* objective = sum x ln(wh) - wh  (Poisson llh w/o ln Gamma)    */
    SAL_PASS_NOCLIP = 64,    /* with UPDATE_H: H_out = H * W^T A without the clip -- this is CorrNMF's
                                aux^T [D][k] (compute_aux, _utils_corrnmf.py:28-52) when H_in holds the exposures */
    SAL_PASS_PARTIALS_ONLY = 128, /* measurement aid: run the streaming kernel only and leave its per-CTA partial sums in
                                the workspace (Wnum / objective / hsum are NOT written), so that a caller can time the
                                kernel back to back between one pair of events (bench.py roofline leg) */
    SAL_PASS_SCALED_UPDATE = 256 /* with h_scale and UPDATE_H: H_out = the multiplicative update of the rescaled exposures
                                Hs = clip(H_in * h_scale), i.e. MvNMF's next H step (mvnmf.py:197-203 after the
                                normalisation of :80-88) fused into the line-search trial's objective pass, instead of
                                Hs itself; the objective is still KL(X || W Hs) */
};

const char* sal_last_error(void);
int sal_version(void);

/* Create / destroy a workspace for problems of shape (V, D_local, k).  Owns only small
 * scratch (per-CTA partial sums).  `device` is the CUDA ordinal. */
int sal_create(sal_handle_t* out, int V, int64_t D_local, int k, int dtype, int device);
int sal_destroy(sal_handle_t h);
/* Free the scratch buffers that destroyed handles left on the per-device free list (sal_create reuses them otherwise). */
int sal_trim_scratch(void);
int sal_set_math(sal_handle_t h, int math_mode);
/* Diagnostics of the tensor-core pass: when buf != NULL (device, >= 32768 floats) CTA 0 dumps the quotient tile
 * R[128][96] and Hn[128][32] of its first tile, then a clock64 timeline [role 4][tile 48][phase 8] (uint32) of
 * its warp roles.  Pass NULL to switch off. */
int sal_set_debug_buffer(sal_handle_t h, void* buf);
/* Kernel timing for benchmarks: when on, every fused-pass kernel launched with UPDATE_H | WNUM is bracketed by CUDA
 * events on its stream.  sal_get_pass_timing synchronises on them and returns their summed duration and count
 * since the last call (bench.py "roofline"). */
int sal_set_timing(sal_handle_t h, int on);
int sal_get_pass_timing(sal_handle_t h, double* total_ms, int64_t* n_launches);
/* number of kernels this handle has launched since creation (bench.py "gpu_launches") */
int64_t sal_launch_count(sal_handle_t h);

/*
 * The fused KL-NMF pass: streams X once and, per sample, rebuilds (WH) on the fly,
 * forms A = X / (WH) and accumulates whatever `flags` asks for.  (WH) and A never touch HBM.
 *
 *   replaces  update_WH          models/_utils_klnmf.py:281-361   (UPDATE_H | WNUM)
 *             update_H           models/_utils_klnmf.py:220-278   (UPDATE_H)
 *             update_W           models/_utils_klnmf.py:164-217   (WNUM; then sal_w_epilogue)
 *             kl_divergence      models/_utils_klnmf.py:11-55     (OBJECTIVE)
 *             samplewise_kl_divergence  _utils_klnmf.py:58-97     (SAMPLEWISE)
 *             _poisson_llh_wo_factorial _utils_klnmf.py:100-135   (POISSON)
 *             KLNMF.objective_function  models/klnmf.py:64-80     (OBJECTIVE with w_lhalf)
 *
 *   w_kl, w_lhalf : [D] or NULL (per-sample KL weights / l-half penalty weights)
 *   h_scale       : [k] or NULL; when given the pass reads H as clip(H_in * h_scale[k])
 *                   (MvNMF line-search trial: normalize_WH + clip, utils.py:155-158,
 *                   mvnmf.py:80-81) and, with UPDATE_H, writes that H to H_out unchanged.
 *   H_out may alias H_in.  Wnum / objective / per_sample / hsum may be NULL if not flagged.
 *   Stream ordering: the tensor-core pass is launched as a programmatic dependent launch and requests its first tiles
 *   of X before it waits for the previous kernel of the stream (X is constant during a fit).  W, H, weights and every
 *   output are only touched after that wait.  Hence X must not be written by the KERNEL that directly precedes a pass on
 *   the same stream.  The library enforces this for its own writer: after sal_clip_counts (the only kernel of this library
 *   that writes X) the handle's next pass is launched in plain stream order.  A caller whose own kernel writes X right before
 *   a pass either synchronises / puts a memcpy or event in between, or calls sal_mark_counts_written(h) first.
 *   (sal_klnmf_period is an ordinary stream-ordered launch and has no such requirement.)
 */
int sal_klnmf_pass(sal_handle_t h, const void* X, const void* W, const void* H_in, void* H_out,
                   const void* w_kl, const void* w_lhalf, const void* h_scale, int flags,
                   void* Wnum, double* objective, void* per_sample, void* hsum, void* stream);

/*
 * One whole joint update in two launches (single-GPU fast path): the fused pass with UPDATE_H | WNUM
 * (| OBJECTIVE of the INCOMING iterate when objective != NULL), then the fixed-order reduction of the per-CTA
 * partials with the W epilogue of sal_w_epilogue fused into it.
 *   replaces  update_WH  models/_utils_klnmf.py:281-361  (clip_given = 1)  and, with separate calls, nothing else.
 * W_out must not alias W_in (H is updated with the OLD W, :345).  H_out may alias H_in.  Wnum [k][V] receives the
 * raw numerator (useful for diagnostics).  n_given == k only updates H and copies W.
 */
int sal_klnmf_update(sal_handle_t h, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out,
                     const void* w_kl, const void* w_lhalf, int n_given, int clip_given, void* Wnum,
                     double* objective, void* stream);

/*
 * Multi-GPU joint update (samples sharded over ranks): as sal_klnmf_update, but the kernel that
 * reduces this rank's partials also performs the all-reduce of the 96 x k numerator (and of the objective scalar):
 * every thread pushes its double as two sequence-tagged 8-byte words into all peers' receive buffers (NVLink
 * peer stores) and polls its own buffer; the contributions are summed in rank order so that W stays bit-identical on
 * all ranks; the W epilogue follows -- pass + ONE more kernel per iteration, no library collective.
 *   peer_buffers : device array [n_ranks] of pointers to every rank's receive buffer of
 *                  sal_p2p_exchange_bytes(k, n_ranks) bytes (zero-initialised symmetric / peer-mapped memory; entry
 *                  `rank` is this rank's own buffer)
 *   p2p_state    : device unsigned[2] = {1, 0} (sequence number, ticket), private to this rank
 * All ranks must call it the same number of times; a missing peer traps after a bounded wait.
 */
int sal_klnmf_update_p2p(sal_handle_t h, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out,
                         const void* w_kl, const void* w_lhalf, int n_given, int clip_given, void* Wnum, double* objective,
                         const void* peer_buffers, void* p2p_state, int n_ranks, int rank, void* stream);
size_t sal_p2p_exchange_bytes(int k, int n_ranks);

/*
 * A whole convergence-test period in ONE persistent launch: the body of the reference's fit loop
 * (models/signature_nmf.py:361-380) -- `n_updates` joint multiplicative updates (update_WH,
 * models/_utils_klnmf.py:281-361, unweighted) with the KL objective (kl_divergence, :11-55) of the incoming iterate of
 * every `objective_every`-th update (0: never) fused into that update, and, with `final_objective`, one objective-only
 * sweep over X after the last update (the value signature_nmf.py:365-380 evaluates at max_iterations):
 *     (W_in, H_in) -> (W_out, H_out);  objectives[i] = KL before update i * objective_every;
 *     objectives[ceil(n_updates / objective_every)] = KL of the final iterate (if asked for).
 * fp32 handles in SAL_MATH_TF32 mode, V = 96, k <= 32 (sal_klnmf_period_supported).  The CTAs stay resident
 * (cooperative launch) and exchange the per-CTA numerators through sequence-tagged words in global memory: no kernel
 * boundary, no reduction kernel and no grid barrier between two updates; every CTA applies the W epilogue
 * (_utils_klnmf.py:338-341) itself.  H_out also serves as the working copy of the exposures after the first update
 * (H_out == H_in: in place); W_out may alias W_in.
 * Several GPUs (n_ranks > 1, samples sharded as in sal_klnmf_update_p2p): every CTA exchanges ITS slice of the numerator
 * with the peers over NVLink inside the same kernel (tagged 16-byte words pushed into the peers' receive buffers, summed
 * in rank order: bit-identical W on all ranks); peer_buffers / p2p_state as for sal_klnmf_update_p2p (the tag counter in
 * p2p_state[0] advances by the number of sweeps).  With n_ranks == 1 both may be null.
 * sal_klnmf_period_emulated runs 1 or 2 "virtual ranks" (one handle, shard and receive buffer each) inside ONE launch on one
 * GPU, so that the exchange protocol is testable without a second GPU (kernels that wait on one another must not be
 * separate launches on one device).
 */
int sal_klnmf_period_supported(sal_handle_t h, int n_given, int n_ranks);
int sal_klnmf_period(sal_handle_t h, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, int n_given,
                     int clip_given, int n_updates, int objective_every, int final_objective, double* objectives,
                     const void* peer_buffers, void* p2p_state, int n_ranks, int rank, void* stream);
int sal_klnmf_period_emulated(const sal_handle_t* handles, int n_virtual, const void* const* X, const void* const* W_in,
                              void* const* W_out, const void* const* H_in, void* const* H_out, int n_given, int clip_given,
                              int n_updates, int objective_every, int final_objective, double* const* objectives,
                              const void* const* peer_tables, void* const* states, void* stream);

/*
 * Small problems (D_local <= 256, state fits the shared memory of one SM -- BASELINE config 0, 96 x 192): n_iterations
 * joint updates (update_WH, _utils_klnmf.py:281-361, unweighted) in ONE launch of a single persistent CTA; *objective
 * (optional) receives the KL divergence of the INCOMING iterate (kl_divergence, :11-55).  W_out / H_out may alias the
 * inputs.  sal_klnmf_small_supported returns 1 when the handle's shape qualifies.
 */
int sal_klnmf_small_supported(sal_handle_t h);
int sal_klnmf_small_updates(sal_handle_t h, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out,
                            int n_given, int n_iterations, double* objective, void* stream);

/*
 * W epilogue: W_out = clip(colnorm(W_in * Wnum)) with given signatures restored.
 *   replaces the W tail of update_WH (_utils_klnmf.py:338-341, clip_given = 1: ALL columns
 *   clipped) and of update_W (:212-215, clip_given = 0: only non-given columns clipped).
 *   n_given == k leaves W unchanged.  W_out may alias W_in.
 */
int sal_w_epilogue(sal_handle_t h, const void* W_in, const void* Wnum, int n_given, int clip_given,
                   void* W_out, void* stream);

/*
 * In-place clip of the count matrix to EPSILON = 2^-23 (the reference does this on the host before every
 * fit: SignatureNMF._setup_adata, models/signature_nmf.py:280-281).  n = number of elements of X;
 * *n_changed (device int64, must be zeroed by the caller) receives how many entries were raised.
 */
int sal_clip_counts(sal_handle_t h, void* X, int64_t n, long long* n_changed, void* stream);
/* X was (or is being) written by a kernel outside this library: the handle's next pass must not read it ahead of stream order. */
int sal_mark_counts_written(sal_handle_t h);

/* H[d][j] <- max(H[d][j] * scale[j], EPSILON) in place: the exposure half of the normalise-and-clip that ends every
 * initialisation (initialize_mat, initialization/initialize.py:116-118; normalize_WH, utils.py:155-158). */
int sal_scale_clip_rows(sal_handle_t h, void* H, const void* scale, void* stream);

/* ---- MvNMF single-CTA k x k steps (models/mvnmf.py) -------------------------------- */

/* out[0] = ln det(W^T W + delta I)   (volume_logdet, mvnmf.py:19-24; LU with pivoting) */
int sal_mvnmf_logdet(sal_handle_t h, const void* W, double delta, double* out, void* stream);

/*
 * W_unc = update_W_unconstrained (mvnmf.py:37-66) given the raw numerator
 * N = (X/(WH)) H^T (sal_klnmf_pass WNUM) and rowsums of H (HSUM).
 */
int sal_mvnmf_w_unconstrained(sal_handle_t h, const void* W, const void* N, const void* hsum,
                              double lam, double delta, int n_given, void* W_unc, void* stream);

/*
 * One line-search candidate (mvnmf.py:80-81, 85-88 + utils.py:155-158):
 *   Wb = gamma_blend < 0 ? W_unc : (1 - gamma) W + gamma W_unc ;  s = colsum(Wb)
 *   W_trial = clip(Wb / s) ;  h_scale[k] = s[k] ;  logdet_out = ln det(W_trial^T W_trial + delta I)
 * The trial objective is then sal_klnmf_pass(OBJECTIVE, h_scale) + lam * logdet.
 */
int sal_mvnmf_trial(sal_handle_t h, const void* W, const void* W_unc, double gamma_blend, double delta,
                    void* W_trial, void* h_scale, double* logdet_out, void* stream);
/* sal_mvnmf_w_unconstrained followed by the FIRST line-search candidate (sal_mvnmf_trial with gamma_blend < 0: the full step,
 * which the line search always tries first, mvnmf.py:80-88) in one launch; same results as the two calls. */
int sal_mvnmf_w_unconstrained_trial(sal_handle_t h, const void* W, const void* N, const void* hsum, double lam, double delta,
                                    int n_given, void* W_unc, void* W_trial, void* h_scale, double* logdet_out,
                                    void* stream);

/*
 * Small MvNMF problems (D_local <= 256, state fits the shared memory of one SM -- BASELINE config 1, MvNMF on 96 x 192):
 * n_iterations whole iterations of MvNMF._update_parameters (models/mvnmf.py:190-210: update_H, update_W_unconstrained
 * :37-66, line_search :69-92 with normalize_WH utils.py:155-158) in ONE launch of a single persistent CTA, back-tracking
 * included; *objective (optional) receives the penalised objective (kl_divergence_penalized, :27-34) of the INCOMING
 * iterate.  gamma (the step size that persists across iterations, mvnmf.py:177-188) is read from *gamma_in and left in
 * *gamma_out (device doubles; may alias).  W_out / H_out may alias the inputs.
 */
int sal_mvnmf_small_supported(sal_handle_t h);
int sal_mvnmf_small_updates(sal_handle_t h, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, double lam,
                            double delta, int n_given, int n_iterations, const double* gamma_in, double* gamma_out,
                            double* objective, void* stream);

/* ---- correlated NMF (models/_utils_corrnmf.py, models/corrnmf_det.py) -------------------------------------
 * a [k] signature scalings, b [D] sample scalings, L [k][m] signature embeddings, U [D][m] sample embeddings,
 * auxT [D][k] (= sal_klnmf_pass UPDATE_H | NOCLIP on the exposures), m = dim_embeddings <= sal_corrnmf_max_dim().
 * All in the handle's dtype; the arithmetic is float64.  Per-sample quantities are this rank's rows. */
int sal_corrnmf_max_dim(void);
/* out[d] = sum_v X[d][v]  (iteration-invariant part of update_sample_scalings, _utils_corrnmf.py:170) */
int sal_row_sums(sal_handle_t h, const void* X, void* out, void* stream);
/* H[d][j] = exp(a_j + b_d + l_j . u_d)   (compute_exposures, _utils_corrnmf.py:11-25) */
int sal_corrnmf_exposures(sal_handle_t h, const void* a, const void* b, const void* L, const void* U, int m, void* H,
                          void* stream);
/* b_d = ln xsum_d - ln sum_j exp(a_j + l_j . u_d)   (update_sample_scalings, :141-179) */
int sal_corrnmf_sample_scalings(sal_handle_t h, const void* xsum, const void* a, const void* L, const void* U, int m,
                                void* b_out, void* stream);
/* sums[0..k) = sum_d auxT[d][j], sums[k..2k) = sum_d exp(b_d + l_j . u_d): the two (per-rank) sums of
 * update_signature_scalings (:103-138); a_j = ln sums[j] - ln sums[k + j] after they were added over the ranks
 * (sal_corrnmf_signature_scalings_finish). */
int sal_corrnmf_signature_scalings_sums(sal_handle_t h, const void* auxT, const void* b, const void* L, const void* U,
                                        int m, double* sums, void* stream);
int sal_corrnmf_signature_scalings_finish(sal_handle_t h, const double* sums, void* a_out, void* stream);
/* Newton-CG updates of the embeddings (update_embedding, :354-410; SciPy's algorithm restated on the device):
 * U[d] for every sample (others = signatures, maxiter Newton iterations; corrnmf_det.py:115-141) and L[j] for every
 * signature (others = samples; :88-113; single-GPU: the sums over samples are taken over this rank's rows). */
int sal_corrnmf_sample_embeddings(sal_handle_t h, const void* auxT, const void* a, const void* b, const void* L, void* U,
                                  int m, double variance, int maxiter, void* stream);
/* Multimodal variant (MultimodalCorrNMF.update_sample_embeddings, models/mmcorrnmf.py:398-428): the handle's k is the
 * TOTAL number of signatures over all modalities; auxT [D][k], a [k], L [k][m] are the modalities' values concatenated
 * and b_mat [D][k] holds, for every sample, its scaling in the modality each signature belongs to. */
int sal_corrnmf_sample_embeddings_mm(sal_handle_t h, const void* auxT, const void* a, const void* b_mat, const void* L,
                                     void* U, int m, double variance, int maxiter, void* stream);
int sal_corrnmf_signature_embeddings(sal_handle_t h, const void* auxT, const void* a, const void* b, void* L,
                                     const void* U, int m, double variance, void* stream);
/* The same update for the signatures sig_begin .. sig_begin + sig_count - 1 only; the other rows of L are left alone.
 * Multi-GPU CorrNMF: the per-sample inputs are all-gathered and every rank solves its share of the signatures. */
int sal_corrnmf_signature_embeddings_range(sal_handle_t h, const void* auxT, const void* a, const void* b, void* L,
                                           const void* U, int m, double variance, int sig_begin, int sig_count,
                                           void* stream);
/* Several GPUs, samples sharded: every rank runs the Newton-CG of ALL signatures on ITS samples and the
 * totals of each objective / gradient / Hessian evaluation (the sums over samples of _utils_corrnmf.py:182-351) are
 * exchanged INSIDE the kernel: sequence-tagged 16-byte words pushed over NVLink into every peer's receive buffer
 * (peers = device array [n_ranks] of buffers of sal_corrnmf_sig_exchange_bytes(k, n_ranks) bytes, zeroed once, as mapped
 * into this rank's address space) and summed in rank order -- every rank obtains the same bits, so the solvers stay in lock
 * step and L stays bit-identical.  launch_id: 1, 2, ... < 32768, the same on every rank and larger for every later launch on
 * the same buffers (it prefixes the tags; zero the buffers behind a barrier before starting over). */
size_t sal_corrnmf_sig_exchange_bytes(int k, int n_ranks);
int sal_corrnmf_signature_embeddings_p2p(sal_handle_t h, const void* auxT, const void* a, const void* b, void* L,
                                         const void* U, int m, double variance, const void* peers, int n_ranks, int rank,
                                         unsigned int launch_id, void* stream);
/* The same protocol with 2 ranks emulated on ONE device in one launch (tests on a single GPU): arrays of n_virtual handles
 * / buffers, peer_tables[v] = rank v's device array of the receive buffers. */
int sal_corrnmf_signature_embeddings_emulated(const sal_handle_t* hs, int n_virtual, const void* const* auxT,
                                              const void* const* a, const void* const* b, void* const* L,
                                              const void* const* U, int m, double variance,
                                              const void* const* peer_tables, unsigned int launch_id, void* stream);
/* The few-KB sums of a sample-sharded iteration (W numerator, scaling sums, norms, likelihood: the places where the
 * reference sums over ALL samples, e.g. corrnmf_det.py:39-86) summed over the ranks through peer memory instead of a
 * library collective: values[0 .. n) (n <= sal_p2p_allreduce_max_values()) are replaced in place by the sum over ranks, added
 * in rank order (bit-identical everywhere).  peers = device array [n_ranks] of receive buffers of
 * sal_p2p_allreduce_bytes(n_ranks) bytes, zeroed once; launch_id as for sal_corrnmf_signature_embeddings_p2p (own counter
 * per set of buffers).  _emulated: two ranks on one device in one launch (tests). */
size_t sal_p2p_allreduce_bytes(int n_ranks);
int sal_p2p_allreduce_max_values(void);
int sal_p2p_allreduce_f64(sal_handle_t h, double* values, int n, const void* peers, int n_ranks, int rank,
                          unsigned int launch_id, void* stream);
int sal_p2p_allreduce_f64_emulated(sal_handle_t h, double* const* values, int n, const void* const* peer_tables,
                                   int n_virtual, unsigned int launch_id, void* stream);
/* out[0] = sum L^2, out[1] = sum U^2 (update_variance corrnmf_det.py:60-69, ELBO priors _utils_corrnmf.py:93-98),
 * out[2] = sum lnGamma(1 + X) when X != NULL (constant of poisson_llh, _utils_klnmf.py:159) */
int sal_corrnmf_norms(sal_handle_t h, const void* L, const void* U, int m, const void* X_or_null, double* out,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SALAMANDER_B200_H */
