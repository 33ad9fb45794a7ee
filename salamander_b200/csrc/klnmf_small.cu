// Small-problem KL-NMF: the whole fit state lives in ONE CTA and a launch runs many joint updates.
//
// BASELINE config 0 (KLNMF on 96 x 192 PCAWG counts) is launch-latency bound: one update is ~0.3 MFLOP, a kernel
// launch costs more than the arithmetic.  This kernel keeps X in registers (thread (d, q) owns 24 features of sample d),
// W and H in shared memory, and performs `n_iter` updates of reference update_WH (_utils_klnmf.py:281-361) back to
// back; the KL divergence of the INCOMING iterate (kl_divergence, :11-55) is produced on the way, which is what the
// period-wise fit driver needs (KLNMF._fit_loop).  Unweighted, no l-half penalty; given signatures supported.
// Arithmetic in the handle's dtype with fixed summation orders (deterministic).
#include <cooperative_groups.h>

#include <stdlib.h>

#include "sal_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NT = 1024;         // 4 threads per sample
constexpr int DMAX = NT / 4;     // 256 samples
constexpr int VQ = SAL_VMAX / 4; // 24 features per thread
constexpr int RP = SAL_VMAX + 1; // pitch of the quotient tile (bank spread for the column reads of phase 2)

template <typename T>
__device__ __forceinline__ T tlog(T x);
template <>
__device__ __forceinline__ float tlog(float x) { return logf(x); }
template <>
__device__ __forceinline__ double tlog(double x) { return log(x); }

// x / y.  fp64: reciprocal seed (>= 20 bits) + two Newton steps + one residual correction of the quotient -- a few
// DFMAs instead of the ~25-instruction IEEE division sequence; the result differs from the correctly rounded quotient
// by at most 1 ulp, far inside the 1e-9 trajectory tolerance.  y > 0 and normal here (WH >= k * eps^2).
__device__ __forceinline__ float tdiv(float x, float y) { return x / y; }
__device__ __forceinline__ double tdiv(double x, double y) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    double e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    const double q = x * r;
    return fma(fma(-y, q, x), r, q);
}

template <typename T, int KT>  // KT >= k: unrolled signature loops, per-thread arrays stay in registers
__global__ void __launch_bounds__(NT, 1)
klnmf_small_kernel(const T* X, const T* W_in, T* W_out, const T* H_in, T* H_out, int D, int V, int k, int n_given, int n_iter,
                   double* objective) {
    extern __shared__ __align__(16) unsigned char raw[];
    T* sR = reinterpret_cast<T*>(raw);    // [D][RP]   quotient X / (WH)
    T* sH = sR + (size_t)D * RP;          // [D][k]
    T* sW = sH + (size_t)D * k;           // [k][V]
    T* sWn = sW + (size_t)k * SAL_VMAX;   // [k][V]    W * numerator before normalisation
    __shared__ double s_red[NT / 32];
    __shared__ T s_col[SAL_KMAX];
    const int tid = threadIdx.x, d = tid >> 2, q = tid & 3, v0 = q * VQ;
    const T eps = (T)SAL_EPS_F32;
    const bool row = d < D;

    T x[VQ];
#pragma unroll
    for (int i = 0; i < VQ; ++i) x[i] = (row && v0 + i < V) ? X[(size_t)d * V + v0 + i] : (T)0;
    for (int i = tid; i < k * V; i += NT) sW[(i / V) * SAL_VMAX + i % V] = W_in[i];
    for (int i = tid; i < D * k; i += NT) sH[i] = H_in[i];
    __syncthreads();

    const int n_pass = n_iter > 0 ? n_iter : (objective ? 1 : 0);  // n_iter == 0: only the objective of the iterate
    for (int it = 0; it < n_pass; ++it) {
        // ---- phase 1: quotient, H numerator (old W), optional KL of the incoming iterate
        T hn[KT], hd[KT];
        double kl = 0.0;
#pragma unroll
        for (int j = 0; j < KT; ++j) hn[j] = (T)0, hd[j] = (row && j < k) ? sH[d * k + j] : (T)0;
        if (row) {
#pragma unroll(KT <= 8 ? VQ : 2)
            for (int i = 0; i < VQ; ++i) {
                const int v = v0 + i;
                if (v < V) {
                    T wv[KT];
                    T wh = (T)0;
#pragma unroll
                    for (int j = 0; j < KT; ++j) wv[j] = j < k ? sW[j * SAL_VMAX + v] : (T)0, wh += wv[j] * hd[j];
                    const T r = tdiv(x[i], wh);
                    sR[d * RP + v] = r;
#pragma unroll
                    for (int j = 0; j < KT; ++j) hn[j] += wv[j] * r;
                    if (it == 0 && objective) {
                        if (sizeof(T) == 4) {
                            // fp32 iterates, fp64 objective: x ln(x / wh) - x + wh cancels to ~(x - wh)^2 / (2 wh); evaluated in
                            // single precision its rounding noise (~4e-8 of the total on 96 x 192) is the size of the reference's
                            // convergence tolerance 1e-7 and ends fits early (tests/test_fp32_plateau.py)
                            double whd = 0.0;
#pragma unroll
                            for (int j = 0; j < KT; ++j) whd += (double)wv[j] * (double)hd[j];
                            const double xd = (double)x[i];
                            kl += xd != 0.0 ? xd * log(xd / whd) - xd + whd : whd;
                        } else {
                            kl += (double)(x[i] != (T)0 ? x[i] * tlog(r) - x[i] + wh : wh);
                        }
                    }
                }
            }
        }
        if (it == 0 && objective) {  // fixed-order block sum
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
            if ((tid & 31) == 0) s_red[tid >> 5] = kl;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w = 0; w < NT / 32; ++w) t += s_red[w];
                *objective = t;
            }
        }
        if (it >= n_iter) break;  // objective-only call
        // the four threads of a sample are adjacent lanes: add their partial H numerators
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            T t = hn[j];
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            hn[j] = t;
        }
        __syncthreads();  // sR complete, every read of the old sH / sW in phase 1 done
        // ---- phase 2: W numerator  Wn[j][v] = W[j][v] * sum_d R[d][v] H[d][j]  (old H), thread <-> (j, v)
        for (int i = tid; i < k * V; i += NT) {
            const int j = i / V, v = i - j * V;
            T a0 = (T)0, a1 = (T)0, a2 = (T)0, a3 = (T)0;  // four interleaved partial sums: short dependency chains
            int dd = 0;
            for (; dd + 3 < D; dd += 4) {
                a0 += sR[dd * RP + v] * sH[dd * k + j];
                a1 += sR[(dd + 1) * RP + v] * sH[(dd + 1) * k + j];
                a2 += sR[(dd + 2) * RP + v] * sH[(dd + 2) * k + j];
                a3 += sR[(dd + 3) * RP + v] * sH[(dd + 3) * k + j];
            }
            for (; dd < D; ++dd) a0 += sR[dd * RP + v] * sH[dd * k + j];
            sWn[j * SAL_VMAX + v] = sW[j * SAL_VMAX + v] * ((a0 + a1) + (a2 + a3));
        }
        __syncthreads();
        // ---- H update (one of the four threads of a sample writes), column sums of Wn (one warp per signature)
        if (row && q == 0) {
#pragma unroll
            for (int j = 0; j < KT; ++j)
                if (j < k) sH[d * k + j] = max(hd[j] * hn[j], eps);
        }
        {
            const int w = tid >> 5, lane = tid & 31;
            if (w < k) {
                T t = (T)0;
                for (int v = lane; v < V; v += 32) t += sWn[w * SAL_VMAX + v];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) s_col[w] = t;
            }
        }
        __syncthreads();
        // ---- W epilogue: normalise, keep the given signatures, clip ALL columns (update_WH, :338-341)
        if (n_given < k)
            for (int i = tid; i < k * V; i += NT) {
                const int j = i / V, v = i - j * V;
                T out = sWn[j * SAL_VMAX + v] / s_col[j];
                if (j < n_given) out = sW[j * SAL_VMAX + v];
                sW[j * SAL_VMAX + v] = max(out, eps);
            }
        __syncthreads();
    }
    for (int i = tid; i < k * V; i += NT) W_out[i] = sW[(i / V) * SAL_VMAX + i % V];
    for (int i = tid; i < D * k; i += NT) H_out[i] = sH[i];
}

template <typename T>
size_t small_smem(int D, int k) {
    return sizeof(T) * ((size_t)D * RP + (size_t)D * k + 2 * (size_t)k * SAL_VMAX);
}

// ---------------------------------------------------------------------------------------------------------
// The same updates on a thread-block CLUSTER: CTA r of C owns a contiguous run of ceil(D / C) samples (counts in registers,
// exposures and quotient rows in ITS shared memory, 8 threads per sample); W is replicated.  Per update the CTAs exchange their
// partial W numerators (and, for the objective, their KL sums) through distributed shared memory and add them in CTA order --
// every CTA then normalises the same W.  The partial buffers alternate with the parity of the update, so ONE cluster barrier per
// update is enough (a CTA can only overwrite a buffer two updates later, after everybody has passed the barrier in between).
// ---------------------------------------------------------------------------------------------------------
constexpr int CNT = 256;             // threads per CTA
constexpr int SPC_MAX = CNT / 8;     // samples per CTA with 8 threads per sample (sizes the shared-memory tiles)
constexpr int CMAX = 8;              // portable cluster size (8 threads per sample); 16 CTAs x 16 threads per sample where allowed

template <typename T, int KT, int TPS>
__global__ void __launch_bounds__(CNT, 1)
klnmf_cluster_kernel(const T* X, const T* W_in, T* W_out, const T* H_in, T* H_out, int D, int V, int k, int n_given, int n_iter,
                     double* objective) {
    constexpr int CVQ = SAL_VMAX / TPS, SPC = SPC_MAX;  // features per thread; rows of the shared-memory tiles
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
    const int Dl_max = (D + C - 1) / C;
    const int d_lo = cr * Dl_max, Dl = max(0, min(D - d_lo, Dl_max));
    extern __shared__ __align__(16) unsigned char raw[];
    T* sR = reinterpret_cast<T*>(raw);          // [SPC][RP]  quotient rows of this CTA's samples
    T* sH = sR + (size_t)SPC * RP;              // [SPC][k]
    T* sW = sH + (size_t)SPC * k;               // [k][V]
    T* sWn = sW + (size_t)k * SAL_VMAX;         // [k][V]     W * numerator before normalisation
    T* sWp = sWn + (size_t)k * SAL_VMAX;        // [2][k][V]  this CTA's partial numerator of the even / odd updates
    __shared__ double s_red[CNT / 32];
    __shared__ double s_kl;                     // this CTA's KL sum (read by CTA 0)
    __shared__ T s_col[SAL_KMAX];
    const int tid = threadIdx.x, dl = tid / TPS, q = tid % TPS, v0 = q * CVQ;
    const T eps = (T)SAL_EPS_F32;
    const bool row = dl < Dl;
    const int d = d_lo + dl;

    T x[CVQ];
#pragma unroll
    for (int i = 0; i < CVQ; ++i) x[i] = (row && v0 + i < V) ? X[(size_t)d * V + v0 + i] : (T)0;
    for (int i = tid; i < k * V; i += CNT) sW[(i / V) * SAL_VMAX + i % V] = W_in[i];
    for (int i = tid; i < Dl * k; i += CNT) sH[i] = H_in[(size_t)d_lo * k + i];
    __syncthreads();

    const int n_pass = n_iter > 0 ? n_iter : (objective ? 1 : 0);
    for (int it = 0; it < n_pass; ++it) {
        const bool want_kl = it == 0 && objective != nullptr;
        // ---- phase 1: quotient, H numerator (old W), optional KL of the incoming iterate
        T hn[KT], hd[KT];
        double kl = 0.0;
#pragma unroll
        for (int j = 0; j < KT; ++j) hn[j] = (T)0, hd[j] = (row && j < k) ? sH[dl * k + j] : (T)0;
        if (row) {
#pragma unroll(KT <= 8 ? CVQ : 2)
            for (int i = 0; i < CVQ; ++i) {
                const int v = v0 + i;
                if (v < V) {
                    T wv[KT];
                    T wh = (T)0;
#pragma unroll
                    for (int j = 0; j < KT; ++j) wv[j] = j < k ? sW[j * SAL_VMAX + v] : (T)0, wh += wv[j] * hd[j];
                    const T r = tdiv(x[i], wh);
                    sR[dl * RP + v] = r;
#pragma unroll
                    for (int j = 0; j < KT; ++j) hn[j] += wv[j] * r;
                    if (want_kl) {
                        if (sizeof(T) == 4) {  // fp32 iterates, fp64 objective (see klnmf_small_kernel)
                            double whd = 0.0;
#pragma unroll
                            for (int j = 0; j < KT; ++j) whd += (double)wv[j] * (double)hd[j];
                            const double xd = (double)x[i];
                            kl += xd != 0.0 ? xd * log(xd / whd) - xd + whd : whd;
                        } else {
                            kl += (double)(x[i] != (T)0 ? x[i] * tlog(r) - x[i] + wh : wh);
                        }
                    }
                }
            }
        }
        if (want_kl) {  // fixed-order sum over the CTA; the cluster-wide sum follows behind the barrier below
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
            if ((tid & 31) == 0) s_red[tid >> 5] = kl;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int w = 0; w < CNT / 32; ++w) t += s_red[w];
                s_kl = t;
            }
        }
        if (it >= n_iter) {  // objective-only call
            cluster.sync();
            if (cr == 0 && tid == 0) {
                double t = 0.0;
                for (int r = 0; r < C; ++r) t += *cluster.map_shared_rank(&s_kl, r);
                *objective = t;
            }
            break;
        }
        // the eight threads of a sample are adjacent lanes: add their partial H numerators
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            T t = hn[j];
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 4);
            if (TPS == 16) t += __shfl_xor_sync(0xffffffffu, t, 8);
            hn[j] = t;
        }
        __syncthreads();  // sR complete, every read of the old sH / sW in phase 1 done
        // ---- phase 2: this CTA's partial numerator sum_d R[d][v] H[d][j] over ITS samples (old H), thread <-> (j, v)
        T* part = sWp + (size_t)(it & 1) * k * SAL_VMAX;
        for (int i = tid; i < k * V; i += CNT) {
            const int j = i / V, v = i - j * V;
            T a0 = (T)0, a1 = (T)0;
            int dd = 0;
            for (; dd + 1 < Dl; dd += 2) {
                a0 += sR[dd * RP + v] * sH[dd * k + j];
                a1 += sR[(dd + 1) * RP + v] * sH[(dd + 1) * k + j];
            }
            for (; dd < Dl; ++dd) a0 += sR[dd * RP + v] * sH[dd * k + j];
            part[j * SAL_VMAX + v] = a0 + a1;
        }
        cluster.sync();  // every CTA's partial (and KL sum) is published; all reads of the old sH in this CTA are done
        if (want_kl && cr == 0 && tid == 0) {
            double t = 0.0;
            for (int r = 0; r < C; ++r) t += *cluster.map_shared_rank(&s_kl, r);
            *objective = t;
        }
        // ---- numerator over all samples (CTA order), times W; H update of this CTA's samples
        for (int i = tid; i < k * V; i += CNT) {
            const int at = (i / V) * SAL_VMAX + i % V;
            T t = (T)0;
            for (int r = 0; r < C; ++r) t += cluster.map_shared_rank(part, r)[at];
            sWn[at] = sW[at] * t;
        }
        if (row && q == 0) {
#pragma unroll
            for (int j = 0; j < KT; ++j)
                if (j < k) sH[dl * k + j] = max(hd[j] * hn[j], eps);
        }
        __syncthreads();
        {
            const int w = tid >> 5, lane = tid & 31;
            for (int j = w; j < k; j += CNT / 32) {
                T t = (T)0;
                for (int v = lane; v < V; v += 32) t += sWn[j * SAL_VMAX + v];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) s_col[j] = t;
            }
        }
        __syncthreads();
        // ---- W epilogue (replicated): normalise, keep the given signatures, clip ALL columns (update_WH, :338-341)
        if (n_given < k)
            for (int i = tid; i < k * V; i += CNT) {
                const int j = i / V, v = i - j * V;
                T out = sWn[j * SAL_VMAX + v] / s_col[j];
                if (j < n_given) out = sW[j * SAL_VMAX + v];
                sW[j * SAL_VMAX + v] = max(out, eps);
            }
        __syncthreads();
    }
    if (cr == 0)
        for (int i = tid; i < k * V; i += CNT) W_out[i] = sW[(i / V) * SAL_VMAX + i % V];
    for (int i = tid; i < Dl * k; i += CNT) H_out[(size_t)d_lo * k + i] = sH[i];
    cluster.sync();  // nobody leaves while a neighbour may still read its shared memory
}

template <typename T>
size_t cluster_smem(int k) {
    return sizeof(T) * ((size_t)SPC_MAX * RP + (size_t)SPC_MAX * k + 4 * (size_t)k * SAL_VMAX);
}

template <typename T, int KT, int TPS>
int launch_cluster_t(sal_ctx* c, int csize, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, int n_given,
                     int n_iter, double* objective, cudaStream_t st) {
    const size_t smem = cluster_smem<T>(c->k);
    SAL_CUDA(cudaFuncSetAttribute(klnmf_cluster_kernel<T, KT, TPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csize > 8) SAL_CUDA(cudaFuncSetAttribute(klnmf_cluster_kernel<T, KT, TPS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize), cfg.blockDim = dim3(CNT), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = csize, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    const T *Xp = (const T*)X, *Wi = (const T*)W_in, *Hi = (const T*)H_in;
    T *Wo = (T*)W_out, *Ho = (T*)H_out;
    const int D = (int)c->D, V = c->V, k = c->k;
    SAL_CUDA(cudaLaunchKernelEx(&cfg, klnmf_cluster_kernel<T, KT, TPS>, Xp, Wi, Wo, Hi, Ho, D, V, k, n_given, n_iter, objective));
    c->launches++;
    return 0;
}

template <typename T, int TPS>
int launch_cluster_k(sal_ctx* c, int csize, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, int n_given,
                     int n_iter, double* objective, cudaStream_t st) {
    if (c->k <= 4) return launch_cluster_t<T, 4, TPS>(c, csize, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
    if (c->k <= 8) return launch_cluster_t<T, 8, TPS>(c, csize, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
    if (c->k <= 16) return launch_cluster_t<T, 16, TPS>(c, csize, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
    return launch_cluster_t<T, 32, TPS>(c, csize, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
}

}  // namespace

static bool single_cta_fits(const sal_ctx* c) {
    const size_t need = c->dtype == SAL_F32 ? small_smem<float>((int)c->D, c->k) : small_smem<double>((int)c->D, c->k);
    return need <= 220 * 1024;
}
static bool cluster_fits(const sal_ctx* c) {  // (the cluster kernel's tiles do not grow with D: 32 samples per CTA at most)
    const size_t need = c->dtype == SAL_F32 ? cluster_smem<float>(c->k) : cluster_smem<double>(c->k);
    return c->D >= 64 && (c->D + CMAX - 1) / CMAX <= SPC_MAX && need <= 220 * 1024;
}

bool sal_small_supported(const sal_ctx* c) {
    if (c->D < 1 || c->D > DMAX || c->V > SAL_VMAX) return false;
    return single_cta_fits(c) || cluster_fits(c);
}

template <typename T, int KT>
int launch_small_t(sal_ctx* c, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, int n_given, int n_iter,
                   double* objective, cudaStream_t st) {
    const int D = (int)c->D;
    const size_t smem = small_smem<T>(D, c->k);
    SAL_CUDA(cudaFuncSetAttribute(klnmf_small_kernel<T, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    klnmf_small_kernel<T, KT><<<1, NT, smem, st>>>((const T*)X, (const T*)W_in, (T*)W_out, (const T*)H_in, (T*)H_out, D, c->V, c->k, n_given,
                                                  n_iter, objective);
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

template <typename T>
int launch_small_k(sal_ctx* c, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, int n_given, int n_iter,
                   double* objective, cudaStream_t st) {
    if (c->k <= 4) return launch_small_t<T, 4>(c, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
    if (c->k <= 8) return launch_small_t<T, 8>(c, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
    if (c->k <= 16) return launch_small_t<T, 16>(c, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
    return launch_small_t<T, 32>(c, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
}

int sal_launch_klnmf_small(sal_ctx* c, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, int n_given,
                           int n_iter, double* objective, cudaStream_t st) {
    // enough samples for several SMs: the cluster kernel, 8 CTAs x 8 threads per sample (16 CTAs x 16 threads per sample, the
    // non-portable cluster size, measured the same 121k - 123k iterations/s on PCAWG: the update is bound by its barriers, not by
    // arithmetic); SAL_B200_KLNMF_CLUSTER = 0 / 8 / 16 forces a choice
    int csize = c->D >= 64 ? CMAX : 1;
    if (const char* e = getenv("SAL_B200_KLNMF_CLUSTER")) {
        const int forced = atoi(e);
        if (forced <= 1) csize = 1;
        else if (forced == 8 || (forced == 16 && c->D <= 16 * (CNT / 16))) csize = forced;
    }
    if (csize == 8 && !cluster_fits(c)) csize = 1;
    if (csize == 1 && !single_cta_fits(c)) {
        if (cluster_fits(c)) {
            csize = 8;  // (the only kernel this problem fits)
        } else {
            sal_set_error("sal_klnmf_small_updates: problem does not fit the small-problem kernels");
            return SAL_EUNSUPPORTED;
        }
    }
    if (csize == 16) {  // (refused on devices / partitions without room for a 16-CTA cluster: fall back to the portable size)
        const int err = c->dtype == SAL_F32 ? launch_cluster_k<float, 16>(c, 16, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st)
                                            : launch_cluster_k<double, 16>(c, 16, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
        if (err == 0) return 0;
        (void)cudaGetLastError();
        csize = cluster_fits(c) ? 8 : 1;
    }
    if (csize == 8)
        return c->dtype == SAL_F32 ? launch_cluster_k<float, 8>(c, 8, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st)
                                   : launch_cluster_k<double, 8>(c, 8, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
    return c->dtype == SAL_F32 ? launch_small_k<float>(c, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st)
                               : launch_small_k<double>(c, X, W_in, W_out, H_in, H_out, n_given, n_iter, objective, st);
}
