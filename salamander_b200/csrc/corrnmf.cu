// Correlated NMF kernels (reference models/_utils_corrnmf.py, models/corrnmf_det.py).
//
//   exposures            H_dk = exp(a_k + b_d + l_k.u_d)                          _utils_corrnmf.py:11-25
//   sample scalings      b_d  = ln sum_v x_dv - ln sum_k exp(a_k + l_k.u_d)       :141-179
//   signature scalings   a_k  = ln sum_d aux_kd - ln sum_d exp(b_d + l_k.u_d)     :103-138
//   embedding updates    argmin_e -[sum_i (o_i.e) aux_i - sum_i exp(s + s_i + o_i.e) - |e|^2/(2 var)]   :182-410
//                        sample embeddings: one THREAD per sample (others = the k signatures, 3 Newton iterations,
//                        corrnmf_det.py:115-141); signature embeddings: one CTA per signature, every objective /
//                        gradient / Hessian evaluation is a fixed-order block reduction over the samples (:88-113)
//   variance             clip(mean([L;U]^2))                                      corrnmf_det.py:60-69
//   aux and the W numerator come from the fused pass (klnmf_pass.cu, SAL_PASS_NOCLIP): aux_kd = H_dk (W^T A)_kd.
//
// The minimiser is a device restatement of SciPy's Newton-CG -- truncated Newton, CG inner loop, More'-Thuente DCSRCH
// line search with SciPy's defaults -- the third-party algorithm the reference calls at _utils_corrnmf.py:400-407;
// oracle/corrnmf.py::newton_cg is the same restatement in numpy and is pinned against scipy itself.
// All arithmetic is float64 whatever the handle's storage dtype.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "corrnmf_newton.cuh"

#include <string.h>

namespace cg = cooperative_groups;

namespace {

// ---------------------------------------------------------------------------------------------------------
// sample embeddings: one thread per sample, the k signature embeddings / scalings in shared memory
// ---------------------------------------------------------------------------------------------------------
template <int M>
struct SampleProblem {
    const double* others;    // [k][m] shared
    const double* s_others;  // [k]    shared
    const double* aux;       // [k]    local
    const double* s_vec;     // [k]    local: the sample's scaling for every "other" (one value repeated for a single
                             //        modality; per-modality values in multimodal CorrNMF, mmcorrnmf.py:413-419)
    double inv_var;
    int k, m_rt;
    __device__ double f_grad_hess(const double* x, double* g, double* A) const {
        const int m = M > 0 ? M : m_rt;
        double acc = 0.0, nrm = 0.0;
        _Pragma("unroll") for (int j = 0; j < m * m; ++j) A[j] = 0.0;
        _Pragma("unroll") for (int j = 0; j < m; ++j) g[j] = x[j] * inv_var, nrm += x[j] * x[j];
        // one term: only the upper triangle of the (symmetric) Hessian is accumulated, mirrored at the end
        auto term = [&](int i, double sp, double e) {
            acc += sp * aux[i] - e;
            const double w = e - aux[i];
            _Pragma("unroll") for (int j = 0; j < m; ++j) g[j] += w * others[i * m + j];
            _Pragma("unroll") for (int p = 0; p < m; ++p) {
                const double eo = e * others[i * m + p];
                _Pragma("unroll") for (int q = p; q < m; ++q) A[p * m + q] += eo * others[i * m + q];
            }
        };
        // two "others" per trip: their exp() chains are independent (one Newton-CG per THREAD: the only parallelism a thread has)
        int i = 0;
        for (; i + 1 < k; i += 2) {
            double sp0 = 0.0, sp1 = 0.0;
            _Pragma("unroll") for (int j = 0; j < m; ++j) sp0 += others[i * m + j] * x[j], sp1 += others[(i + 1) * m + j] * x[j];
            const double e0 = exp(s_vec[i] + s_others[i] + sp0), e1 = exp(s_vec[i + 1] + s_others[i + 1] + sp1);
            term(i, sp0, e0);
            term(i + 1, sp1, e1);
        }
        if (i < k) {
            double sp = 0.0;
            _Pragma("unroll") for (int j = 0; j < m; ++j) sp += others[i * m + j] * x[j];
            term(i, sp, exp(s_vec[i] + s_others[i] + sp));
        }
        _Pragma("unroll") for (int p = 0; p < m; ++p) {
            _Pragma("unroll") for (int q = p + 1; q < m; ++q) A[q * m + p] = A[p * m + q];
            A[p * m + p] += inv_var;
        }
        return -(acc - 0.5 * nrm * inv_var);
    }
};

// b: [D] (b_is_matrix = 0, one scaling per sample) or [D][k] (b_is_matrix = 1, one per sample and signature)
template <typename T, int M>
__global__ void __launch_bounds__(128) sample_embeddings_kernel(const T* auxT, const T* a, const T* b, int b_is_matrix, const T* L, T* U,
                                                               int64_t D, int k, int m_rt, double variance, int maxiter) {
    constexpr int MM = M > 0 ? M : MAXM;
    const int m = M > 0 ? M : m_rt;
    __shared__ double sL[SAL_KMAX * MAXM];
    __shared__ double sa[SAL_KMAX];
    for (int i = threadIdx.x; i < k * m; i += blockDim.x) sL[i] = (double)L[i];
    for (int i = threadIdx.x; i < k; i += blockDim.x) sa[i] = (double)a[i];
    __syncthreads();
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    double aux[SAL_KMAX], sv[SAL_KMAX], x[MM];
    for (int i = 0; i < k; ++i) aux[i] = (double)auxT[d * k + i], sv[i] = (double)(b_is_matrix ? b[d * k + i] : b[d]);
#pragma unroll
    for (int j = 0; j < MM; ++j)
        if (j < m) x[j] = (double)U[d * m + j];
    SampleProblem<M> p{sL, sa, aux, sv, 1.0 / variance, k, m};
    newton_cg<M>(p, x, m, maxiter);
#pragma unroll
    for (int j = 0; j < MM; ++j)
        if (j < m) U[d * m + j] = (T)snap_eps(x[j]);
}

// ---------------------------------------------------------------------------------------------------------
// element-wise / reduction kernels
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) exposures_kernel(const T* a, const T* b, const T* L, const T* U, int64_t D, int k, int m, T* H) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D * k) return;
    const int64_t d = i / k;
    const int j = (int)(i - d * k);
    double sp = 0.0;
    for (int q = 0; q < m; ++q) sp += (double)L[j * m + q] * (double)U[d * m + q];
    H[i] = (T)exp((double)a[j] + (double)b[d] + sp);
}

template <typename T>
__global__ void __launch_bounds__(256) row_sums_kernel(const T* X, int64_t D, int V, T* out) {
    const int64_t d = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (d >= D) return;
    double s = 0.0;
    for (int v = threadIdx.x & 31; v < V; v += 32) s += (double)X[d * V + v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) out[d] = (T)s;
}

template <typename T>
__global__ void __launch_bounds__(256) sample_scalings_kernel(const T* xsum, const T* a, const T* L, const T* U, int64_t D, int k, int m,
                                                             T* b) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    double s = 0.0;
    for (int j = 0; j < k; ++j) {
        double sp = 0.0;
        for (int q = 0; q < m; ++q) sp += (double)L[j * m + q] * (double)U[d * m + q];
        s += exp((double)a[j] + sp);
    }
    b[d] = (T)(log((double)xsum[d]) - log(s));
}

__device__ double block_sum_256(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += s_red[w];
    return t;
}

// sums[j] = sum_d aux_dj, sums[k + j] = sum_d exp(b_d + l_j.u_d)  (this rank's samples).  SCAL_BLOCKS CTAs per signature
// (blockIdx.y = signature) each leave a partial pair; the last one to finish adds them in block order (deterministic).
constexpr int SCAL_BLOCKS = 128;  // <= 256: the last block adds the partials one per thread
template <typename T>
__global__ void __launch_bounds__(256) signature_scalings_kernel(const T* auxT, const T* b, const T* L, const T* U, int64_t D, int k, int m,
                                                                double* partial, unsigned int* counter, double* sums) {
    __shared__ double s_red[8];
    __shared__ bool last;
    const int j = blockIdx.y;
    double l[MAXM];
    for (int q = 0; q < m; ++q) l[q] = (double)L[j * m + q];
    double s1 = 0.0, s2 = 0.0;
#pragma unroll 2
    for (int64_t d = (int64_t)blockIdx.x * 256 + threadIdx.x; d < D; d += (int64_t)SCAL_BLOCKS * 256) {
        double sp = 0.0;
        for (int q = 0; q < m; ++q) sp += l[q] * (double)U[d * m + q];
        s1 += (double)auxT[d * k + j];
        s2 += exp((double)b[d] + sp);
    }
    s1 = block_sum_256(s1, s_red);
    s2 = block_sum_256(s2, s_red);
    if (threadIdx.x == 0) {
        partial[(2 * j) * SCAL_BLOCKS + blockIdx.x] = s1;
        partial[(2 * j + 1) * SCAL_BLOCKS + blockIdx.x] = s2;
        __threadfence();
        last = atomicAdd(&counter[j], 1u) == SCAL_BLOCKS - 1;
    }
    __syncthreads();
    if (last) {  // (block-uniform) one partial per thread, then the fixed-order tree of block_sum_256
        __threadfence();
        const int i = threadIdx.x;
        double t1 = i < SCAL_BLOCKS ? ((volatile double*)partial)[(2 * j) * SCAL_BLOCKS + i] : 0.0;
        double t2 = i < SCAL_BLOCKS ? ((volatile double*)partial)[(2 * j + 1) * SCAL_BLOCKS + i] : 0.0;
        t1 = block_sum_256(t1, s_red);
        t2 = block_sum_256(t2, s_red);
        if (threadIdx.x == 0) {
            sums[j] = t1, sums[k + j] = t2;
            counter[j] = 0;
        }
    }
}

template <typename T>
__global__ void signature_scalings_finish_kernel(const double* sums, int k, T* a) {
    const int j = threadIdx.x;
    if (j < k) a[j] = (T)(log(sums[j]) - log(sums[k + j]));
}

// out[0] = sum L^2, out[1] = sum U^2 (this rank's samples), out[2] = sum_x lnGamma(1 + x) when X != null.
// blockIdx.y selects the quantity, NORM_BLOCKS blocks each leave a partial; the last block to finish adds them in order.
constexpr int NORM_BLOCKS = 256;  // <= 256: the last block adds the partials one per thread
template <typename T>
__global__ void __launch_bounds__(256) norms_kernel(const T* L, int64_t nL, const T* U, int64_t nU, const T* X, int64_t nX,
                                                   double* partial, unsigned int* counter, double* out) {
    __shared__ double s_red[8];
    __shared__ bool last;
    const int q = blockIdx.y;
    const T* src = q == 0 ? L : q == 1 ? U : X;
    const int64_t n = q == 0 ? nL : q == 1 ? nU : nX;
    double s = 0.0;
#pragma unroll 4
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)NORM_BLOCKS * 256) {
        const double v = (double)src[i];
        s += q < 2 ? v * v : lgamma(1.0 + v);
    }
    s = block_sum_256(s, s_red);
    if (threadIdx.x == 0) {
        partial[q * NORM_BLOCKS + blockIdx.x] = s;
        __threadfence();
        last = atomicAdd(&counter[q], 1u) == NORM_BLOCKS - 1;
    }
    __syncthreads();
    if (last) {  // (block-uniform)
        __threadfence();
        double t = threadIdx.x < NORM_BLOCKS ? ((volatile double*)partial)[q * NORM_BLOCKS + threadIdx.x] : 0.0;
        t = block_sum_256(t, s_red);
        if (threadIdx.x == 0) {
            out[q] = t;
            counter[q] = 0;
        }
    }
}

}  // namespace

#define SAL_DISPATCH_T(c, call_f, call_d) \
    do {                                  \
        if ((c)->dtype == SAL_F32) {      \
            call_f;                       \
        } else {                          \
            call_d;                       \
        }                                 \
    } while (0)

int sal_launch_corrnmf_exposures(sal_ctx* c, const void* a, const void* b, const void* L, const void* U, int m, void* H, cudaStream_t st) {
    const int64_t n = c->D * c->k;
    if (n == 0) return 0;
    const int grid = (int)((n + 255) / 256);
    SAL_DISPATCH_T(c, (exposures_kernel<float><<<grid, 256, 0, st>>>((const float*)a, (const float*)b, (const float*)L, (const float*)U, c->D, c->k, m, (float*)H)),
                   (exposures_kernel<double><<<grid, 256, 0, st>>>((const double*)a, (const double*)b, (const double*)L, (const double*)U, c->D, c->k, m, (double*)H)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_row_sums(sal_ctx* c, const void* X, void* out, cudaStream_t st) {
    if (c->D == 0) return 0;
    const int grid = (int)((c->D + 7) / 8);
    SAL_DISPATCH_T(c, (row_sums_kernel<float><<<grid, 256, 0, st>>>((const float*)X, c->D, c->V, (float*)out)),
                   (row_sums_kernel<double><<<grid, 256, 0, st>>>((const double*)X, c->D, c->V, (double*)out)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_corrnmf_sample_scalings(sal_ctx* c, const void* xsum, const void* a, const void* L, const void* U, int m, void* b,
                                       cudaStream_t st) {
    if (c->D == 0) return 0;
    const int grid = (int)((c->D + 255) / 256);
    SAL_DISPATCH_T(c, (sample_scalings_kernel<float><<<grid, 256, 0, st>>>((const float*)xsum, (const float*)a, (const float*)L, (const float*)U, c->D, c->k, m, (float*)b)),
                   (sample_scalings_kernel<double><<<grid, 256, 0, st>>>((const double*)xsum, (const double*)a, (const double*)L, (const double*)U, c->D, c->k, m, (double*)b)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

// counters of the "last block adds the partials" reductions: [0..3] norms, [4..4 + SAL_KMAX) signature scalings
static int ensure_counters(sal_ctx* c, cudaStream_t st) {
    if (!c->norm_counter) {
        SAL_CUDA(cudaMalloc((void**)&c->norm_counter, (4 + SAL_KMAX) * sizeof(unsigned int)));
        SAL_CUDA(cudaMemsetAsync(c->norm_counter, 0, (4 + SAL_KMAX) * sizeof(unsigned int), st));
    }
    return 0;
}

int sal_launch_corrnmf_signature_scalings_sums(sal_ctx* c, const void* auxT, const void* b, const void* L, const void* U, int m,
                                               double* sums, cudaStream_t st) {
    if (int e = ensure_counters(c, st)) return e;
    // scratch: 2 * k * SCAL_BLOCKS doubles at the start of the W-numerator partials (no pass is in flight on this stream)
    static_assert(2 * SAL_KMAX * SCAL_BLOCKS * sizeof(double) <= (size_t)148 * SAL_KMAX * SAL_VMAX * 4, "scratch too small");
    double* partial = (double*)c->partial_wnum;
    unsigned int* counter = c->norm_counter + 4;
    const dim3 grid(SCAL_BLOCKS, c->k);
    SAL_DISPATCH_T(c, (signature_scalings_kernel<float><<<grid, 256, 0, st>>>((const float*)auxT, (const float*)b, (const float*)L, (const float*)U, c->D, c->k, m, partial, counter, sums)),
                   (signature_scalings_kernel<double><<<grid, 256, 0, st>>>((const double*)auxT, (const double*)b, (const double*)L, (const double*)U, c->D, c->k, m, partial, counter, sums)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_corrnmf_signature_scalings_finish(sal_ctx* c, const double* sums, void* a, cudaStream_t st) {
    SAL_DISPATCH_T(c, (signature_scalings_finish_kernel<float><<<1, 32, 0, st>>>(sums, c->k, (float*)a)),
                   (signature_scalings_finish_kernel<double><<<1, 32, 0, st>>>(sums, c->k, (double*)a)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_corrnmf_sample_embeddings(sal_ctx* c, const void* auxT, const void* a, const void* b, int b_is_matrix, const void* L,
                                         void* U, int m, double variance, int maxiter, cudaStream_t st) {
    if (c->D == 0) return 0;
    const int grid = (int)((c->D + 127) / 128);
    // the embedding dimension as a template constant (2 .. 5: registers instead of local-memory arrays), 0 = run-time dimension
#define SAL_SAMPLE_EMB(TT_, M_)                                                                                                          \
    sample_embeddings_kernel<TT_, M_><<<grid, 128, 0, st>>>((const TT_*)auxT, (const TT_*)a, (const TT_*)b, b_is_matrix, (const TT_*)L, \
                                                            (TT_*)U, c->D, c->k, m, variance, maxiter)
    if (c->dtype == SAL_F32) {  // (float storage is the rare case for the correlated models: run-time dimension only)
        SAL_SAMPLE_EMB(float, 0);
    } else {
        switch (m) {
            case 2: SAL_SAMPLE_EMB(double, 2); break;
            case 3: SAL_SAMPLE_EMB(double, 3); break;
            case 4: SAL_SAMPLE_EMB(double, 4); break;
            case 5: SAL_SAMPLE_EMB(double, 5); break;
            default: SAL_SAMPLE_EMB(double, 0); break;
        }
    }
#undef SAL_SAMPLE_EMB
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_corrnmf_norms(sal_ctx* c, const void* L, const void* U, int m, const void* X_or_null, double* out, cudaStream_t st) {
    // scratch: the objective partials of the pass (>= 296 doubles) hold the 3 x 64 partials; the counters follow out[3]
    static_assert(3 * NORM_BLOCKS <= 148 * SAL_KMAX, "partial_hsum has n_sm * SAL_KMAX slots at least");
    if (int e = ensure_counters(c, st)) return e;
    const dim3 grid(NORM_BLOCKS, X_or_null ? 3 : 2);
    SAL_DISPATCH_T(c, (norms_kernel<float><<<grid, 256, 0, st>>>((const float*)L, (int64_t)c->k * m, (const float*)U, c->D * m, (const float*)X_or_null, c->D * c->V, c->partial_hsum, c->norm_counter, out)),
                   (norms_kernel<double><<<grid, 256, 0, st>>>((const double*)L, (int64_t)c->k * m, (const double*)U, c->D * m, (const double*)X_or_null, c->D * c->V, c->partial_hsum, c->norm_counter, out)));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_corrnmf_max_dim(void) { return MAXM; }
