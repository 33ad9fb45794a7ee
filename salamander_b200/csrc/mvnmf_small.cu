// Small-problem MvNMF: the whole fit state lives in ONE CTA and a launch runs many iterations, line search included.
//
// BASELINE config 1 (MvNMF n_signatures = 10 on 96 x 192 PCAWG counts) is launch-latency bound when every step is its own
// kernel (3 passes + 4 single-CTA kernels + a host decision per line-search trial: ~190 us per iteration).  Here one
// 1024-thread CTA keeps X in registers (thread (d, q) owns 24 features of sample d), W / H / the quotient in shared memory and
// performs `n_iter` iterations of reference MvNMF._update_parameters (models/mvnmf.py:190-210) back to back:
//
//   H step                  update_H                  _utils_klnmf.py:220-264
//   previous objective      kl_divergence_penalized   mvnmf.py:27-34 (KL of (W, H') + lam * volume_logdet(W), :19-24)
//   unconstrained W step    update_W_unconstrained    mvnmf.py:37-66  (numerator N = (X / (W H')) H'^T, row sums of H')
//   line search             line_search               mvnmf.py:69-92  (first trial without gamma, then gamma <- 0.8 gamma while
//                                                     the penalised objective went up; normalize_WH utils.py:155-158 + clip;
//                                                     gamma <- min(1, 1.2 gamma) persists across iterations, mvnmf.py:177-188)
//
// The penalised objective of the INCOMING iterate is produced on the way (what the period-wise fit driver needs).
// Arithmetic in the handle's dtype with fixed summation orders (deterministic); the k x k algebra, the logarithms and every
// objective sum are float64.  gamma lives in device memory (gamma_in -> gamma_out).
#include <cooperative_groups.h>

#include <stdlib.h>

#include "sal_common.cuh"
#include "mvnmf_kk.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NT = 1024;          // 4 threads per sample
constexpr int VQ = SAL_VMAX / 4;  // 24 features per thread
constexpr int RP = SAL_VMAX + 1;  // pitch of the quotient tile
constexpr int WP = SAL_VMAX;      // pitch of the [k][V] matrices

__device__ __forceinline__ float tdiv(float x, float y) { return x / y; }
__device__ __forceinline__ double tdiv(double x, double y) {  // reciprocal seed + two Newton steps + residual correction (<= 1 ulp)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
    double e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    e = fma(-y, r, 1.0);
    r = fma(r, e, r);
    const double q = x * r;
    return fma(fma(-y, q, x), r, q);
}

struct Shared {  // block-wide scalars and the k x k scratch (float64)
    double red[NT / 32];
    double colsum[SAL_KMAX];
    double col[SAL_KMAX];
    double hsum[SAL_KMAX];
    double bcast[4];
    int piv;
};

__device__ double block_sum(double v, Shared& sh) {  // fixed order; result on every thread
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();  // (sh.red may still be read from the previous call)
    if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < NT / 32; ++w) t += sh.red[w];
    return t;
}

// G = M M^T + delta I for M [k][WP] in shared memory
template <typename T, int NTH = NT>
__device__ void gram(const T* M, double* G, int GP, int V, int k, double delta) {
    for (int i = threadIdx.x; i < k * k; i += NTH) {
        const int a = i / k, b = i - a * k;
        double t = 0.0;
        for (int v = 0; v < V; ++v) t += (double)M[a * WP + v] * (double)M[b * WP + v];
        G[a * GP + b] = t + (a == b ? delta : 0.0);
    }
    __syncthreads();
}

template <typename T, int KT>
__global__ void __launch_bounds__(NT, 1)
mvnmf_small_kernel(const T* X, const T* W_in, T* W_out, const T* H_in, T* H_out, int D, int V, int k, double lam, double delta,
                   int n_given, int n_iter, const double* gamma_in, double* gamma_out, double* objective) {
    extern __shared__ __align__(16) unsigned char raw[];
    T* sR = reinterpret_cast<T*>(raw);     // [D][RP]  quotient X / (W H')
    T* sH = sR + (size_t)D * RP;           // [D][k]
    T* sW = sH + (size_t)D * k;            // [k][WP]
    T* sWu = sW + (size_t)k * WP;          // [k][WP]  numerator N, then W_unconstrained
    T* sWt = sWu + (size_t)k * WP;         // [k][WP]  line-search candidate
    const int GP = k + 1;
    const size_t t_bytes = sizeof(T) * ((size_t)D * RP + (size_t)D * k + 3 * (size_t)k * WP);
    double* G = reinterpret_cast<double*>(raw + ((t_bytes + 7) & ~(size_t)7));  // [k][GP] Gram / LU scratch
    double* Y = G + k * GP;                                                     // [k][GP] inverse
    __shared__ Shared sh;
    const int tid = threadIdx.x, d = tid >> 2, q = tid & 3, v0 = q * VQ;
    const T eps = (T)SAL_EPS_F32;
    const bool row = d < D;

    T x[VQ];
#pragma unroll
    for (int i = 0; i < VQ; ++i) x[i] = (row && v0 + i < V) ? X[(size_t)d * V + v0 + i] : (T)0;
    for (int i = tid; i < k * V; i += NT) sW[(i / V) * WP + i % V] = W_in[i];
    for (int i = tid; i < k * WP; i += NT) sWt[i] = (T)0, sWu[i] = (T)0;
    for (int i = tid; i < D * k; i += NT) sH[i] = H_in[i];
    double gamma = *gamma_in;
    __syncthreads();

    // KL(X || M h) over this thread's 24 features with float64 terms (x = 0 contributes m h)
    auto kl_row = [&](const T* M, const T (&hv)[KT]) {
        double kl = 0.0;
        if (row) {
#pragma unroll 2
            for (int i = 0; i < VQ; ++i) {
                const int v = v0 + i;
                if (v < V) {
                    double wh = 0.0;
#pragma unroll
                    for (int j = 0; j < KT; ++j)
                        if (j < k) wh += (double)M[j * WP + v] * (double)hv[j];
                    const double xd = (double)x[i];
                    kl += xd != 0.0 ? xd * log(xd / wh) - xd + wh : wh;
                }
            }
        }
        return kl;
    };
    auto logdet = [&](const T* M) {  // block-uniform result
        gram(M, G, GP, V, k, delta);
        if (tid < 32) {
            const double det = lu_det_warp(G, GP, k);
            if (tid == 0) sh.bcast[0] = log(det);
        }
        __syncthreads();
        const double r = sh.bcast[0];
        __syncthreads();
        return r;
    };

    double ld_W = 0.0;  // lam-free volume of the current W; carried from the accepted candidate of the previous iteration
    bool have_ld = false;
    const int n_pass = n_iter > 0 ? n_iter : (objective ? 1 : 0);
    for (int it = 0; it < n_pass; ++it) {
        T hd[KT];
#pragma unroll
        for (int j = 0; j < KT; ++j) hd[j] = (row && j < k) ? sH[d * k + j] : (T)0;
        if (it == 0 && objective) {  // penalised objective of the incoming iterate
            const double kl = block_sum(kl_row(sW, hd), sh);
            ld_W = logdet(sW), have_ld = true;
            if (tid == 0) *objective = kl + lam * ld_W;
        }
        if (it >= n_iter) break;  // objective-only call
        // ---- H step: hn = W^T (X / (W h)), h' = clip(h * hn) ----
        T hn[KT];
#pragma unroll
        for (int j = 0; j < KT; ++j) hn[j] = (T)0;
        if (row) {
#pragma unroll(KT <= 8 ? VQ : 2)
            for (int i = 0; i < VQ; ++i) {
                const int v = v0 + i;
                if (v < V) {
                    T wv[KT];
                    T wh = (T)0;
#pragma unroll
                    for (int j = 0; j < KT; ++j) wv[j] = j < k ? sW[j * WP + v] : (T)0, wh += wv[j] * hd[j];
                    const T r = tdiv(x[i], wh);
#pragma unroll
                    for (int j = 0; j < KT; ++j) hn[j] += wv[j] * r;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            T t = hn[j];
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            hd[j] = (row && j < k) ? max(hd[j] * t, eps) : (T)0;  // h' (identical on the four threads of a sample)
        }
        if (n_given >= k) {  // all signatures given: the iteration is the H step (reference mvnmf.py:190-196)
            __syncthreads();
            if (row && q == 0) {
#pragma unroll
                for (int j = 0; j < KT; ++j)
                    if (j < k) sH[d * k + j] = hd[j];
            }
            __syncthreads();
            continue;
        }
        // ---- quotient with the new exposures, previous KL, numerator N = (X / (W H')) H'^T, row sums of H' ----
        double kl_prev = 0.0;
        if (row) {
#pragma unroll(KT <= 8 ? VQ : 2)
            for (int i = 0; i < VQ; ++i) {
                const int v = v0 + i;
                if (v < V) {
                    T wh = (T)0;
                    double whd = 0.0;
#pragma unroll
                    for (int j = 0; j < KT; ++j)
                        if (j < k) {
                            const T w = sW[j * WP + v];
                            wh += w * hd[j];
                            if (sizeof(T) == 4) whd += (double)w * (double)hd[j];
                        }
                    sR[d * RP + v] = tdiv(x[i], wh);
                    const double xd = (double)x[i], wd = sizeof(T) == 4 ? whd : (double)wh;
                    kl_prev += xd != 0.0 ? xd * log(xd / wd) - xd + wd : wd;
                }
            }
        }
        __syncthreads();  // every read of the old sH is done
        if (row && q == 0) {
#pragma unroll
            for (int j = 0; j < KT; ++j)
                if (j < k) sH[d * k + j] = hd[j];
        }
        kl_prev = block_sum(kl_prev, sh);  // (its barriers also publish sR and the new sH)
        for (int i = tid; i < k * V; i += NT) {
            const int j = i / V, v = i - j * V;
            T a0 = (T)0, a1 = (T)0, a2 = (T)0, a3 = (T)0;
            int dd = 0;
            for (; dd + 3 < D; dd += 4) {
                a0 += sR[dd * RP + v] * sH[dd * k + j];
                a1 += sR[(dd + 1) * RP + v] * sH[(dd + 1) * k + j];
                a2 += sR[(dd + 2) * RP + v] * sH[(dd + 2) * k + j];
                a3 += sR[(dd + 3) * RP + v] * sH[(dd + 3) * k + j];
            }
            for (; dd < D; ++dd) a0 += sR[dd * RP + v] * sH[dd * k + j];
            sWu[j * WP + v] = (a0 + a1) + (a2 + a3);
        }
        {
            const int w = tid >> 5, lane = tid & 31;
            if (w < k) {
                double t = 0.0;
                for (int dd = lane; dd < D; dd += 32) t += (double)sH[dd * k + w];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) sh.hsum[w] = (double)(T)t;
            }
        }
        __syncthreads();
        if (!have_ld) ld_W = logdet(sW), have_ld = true;
        const double prev = kl_prev + lam * ld_W;
        // ---- unconstrained W step (mvnmf.py:37-66) ----
        gram(sW, G, GP, V, k, delta);
        if (tid < 32) invert_warp(G, Y, sh.col, GP, k);
        __syncthreads();
        for (int i = tid; i < k * V; i += NT) {
            const int j = i / V, v = i - j * V;
            const double w = (double)sW[j * WP + v];
            double out;
            if (j < n_given) {
                out = w;
            } else {
                double wym = 0.0, wya = 0.0;
                for (int a = 0; a < k; ++a) {
                    const double y = Y[a * GP + j];
                    const double wa = (double)sW[a * WP + v];
                    wym += wa * fmax(0.0, -y);
                    wya += wa * fabs(y);
                }
                const double r = sh.hsum[j];
                const double a1 = r - 4.0 * lam * wym;
                const double s2 = 8.0 * lam * wya * (double)sWu[j * WP + v];
                const double num = sqrt(a1 * a1 + s2) + (-r + 4.0 * lam * wym);
                out = fmax(w * num / (4.0 * lam * wya), (double)SAL_EPS_F32);
            }
            sWu[j * WP + v] = (T)out;
        }
        __syncthreads();
        // ---- line search (mvnmf.py:69-92): candidate = normalise + clip of W_u (first) or of the blend with W ----
        double g_blend = -1.0, ld_t = 0.0;
        while (true) {
            for (int i = tid; i < k * V; i += NT) {
                const int j = i / V, v = i - j * V;
                const double wu = (double)sWu[j * WP + v];
                sWt[j * WP + v] = (T)(g_blend < 0.0 ? wu : (1.0 - g_blend) * (double)sW[j * WP + v] + g_blend * wu);
            }
            __syncthreads();
            {  // column sums of the candidate: a warp per signature
                const int w = tid >> 5, lane = tid & 31;
                if (w < k) {
                    double t = 0.0;
                    for (int v = lane; v < V; v += 32) t += (double)sWt[w * WP + v];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    if (lane == 0) sh.colsum[w] = (double)(T)t;
                }
            }
            __syncthreads();
            for (int i = tid; i < k * V; i += NT) {
                const int j = i / V, v = i - j * V;
                sWt[j * WP + v] = (T)fmax((double)sWt[j * WP + v] / sh.colsum[j], (double)SAL_EPS_F32);
            }
            __syncthreads();
            ld_t = logdet(sWt);
            T ht[KT];
#pragma unroll
            for (int j = 0; j < KT; ++j) ht[j] = (row && j < k) ? max(hd[j] * (T)sh.colsum[j], eps) : (T)0;
            const double val = block_sum(kl_row(sWt, ht), sh) + lam * ld_t;
            if (!(val > prev && gamma > 1e-16)) break;
            gamma *= 0.8;
            g_blend = gamma;
            __syncthreads();  // sh.colsum is rewritten by the next candidate
        }
        gamma = fmin(1.0, 1.2 * gamma);
        ld_W = ld_t;
        // ---- accept: W <- candidate, H <- clip(H' * colsum) ----
        for (int i = tid; i < k * V; i += NT) sW[(i / V) * WP + i % V] = sWt[(i / V) * WP + i % V];
        if (row && q == 0) {
#pragma unroll
            for (int j = 0; j < KT; ++j)
                if (j < k) sH[d * k + j] = max(hd[j] * (T)sh.colsum[j], eps);
        }
        __syncthreads();
    }
    for (int i = tid; i < k * V; i += NT) W_out[i] = sW[(i / V) * WP + i % V];
    for (int i = tid; i < D * k; i += NT) H_out[i] = sH[i];
    if (tid == 0) *gamma_out = gamma;
}

// ---------------------------------------------------------------------------------------------------------
// The same iterations on a thread-block CLUSTER: CTA r of C owns the samples r, r + C, ... in blocks (a contiguous run of
// ceil(D / C) samples), keeps THEIR counts in registers and their exposures / quotient rows in its shared memory; W, the k x k
// algebra and every decision are replicated (same numbers in every CTA, so the CTAs take the same branches).  What needs all
// samples -- the numerator N, the row sums of H', the KL sums -- is exchanged through distributed shared memory and added in
// CTA order (deterministic).  One SM's worth of float64 work becomes C SMs' worth; the serial k x k part stays.
// ---------------------------------------------------------------------------------------------------------
constexpr int CNT = 256;         // threads per CTA
constexpr int SPC_MAX = CNT / 8;  // samples per CTA with 8 threads per sample (sizes the shared-memory tiles)
constexpr int CMAX = 8;           // portable cluster size (8 threads per sample); 16 CTAs x 16 threads per sample where allowed

struct CShared {
    double red[CNT / 32];
    double colsum[SAL_KMAX];
    double col[SAL_KMAX];
    double hsum[SAL_KMAX];
    double bcast[4];
    double xch[2][2 + SAL_KMAX];  // [parity][kl | spare | hsum partials]: this CTA's contribution to a cluster-wide sum
    int piv;
};

__device__ double block_sum_c(double v, CShared& sh) {  // fixed order; result on every thread of the CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < CNT / 32; ++w) t += sh.red[w];
    return t;
}

template <typename T, int KT, int TPS>
__global__ void __launch_bounds__(CNT, 1)
mvnmf_cluster_kernel(const T* X, const T* W_in, T* W_out, const T* H_in, T* H_out, int D, int V, int k, double lam, double delta,
                     int n_given, int n_iter, const double* gamma_in, double* gamma_out, double* objective) {
    constexpr int CVQ = SAL_VMAX / TPS, SPC = SPC_MAX;  // features per thread; rows of the shared-memory tiles
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), cr = (int)cluster.block_rank();
    const int Dl_max = (D + C - 1) / C;                       // samples per CTA (the last CTA may own fewer)
    const int d_lo = cr * Dl_max, Dl = max(0, min(D - d_lo, Dl_max));
    extern __shared__ __align__(16) unsigned char raw[];
    T* sR = reinterpret_cast<T*>(raw);     // [SPC][RP]  quotient rows of this CTA's samples
    T* sH = sR + (size_t)SPC * RP;         // [SPC][k]
    T* sW = sH + (size_t)SPC * k;          // [k][WP]
    T* sWu = sW + (size_t)k * WP;          // [k][WP]  numerator N, then W_unconstrained
    T* sWt = sWu + (size_t)k * WP;         // [k][WP]  line-search candidate
    T* sWp = sWt + (size_t)k * WP;         // [k][WP]  this CTA's partial numerator (read by the whole cluster)
    const int GP = k + 1;
    const size_t t_bytes = sizeof(T) * ((size_t)SPC_MAX * RP + (size_t)SPC_MAX * k + 4 * (size_t)k * WP);
    double* G = reinterpret_cast<double*>(raw + ((t_bytes + 7) & ~(size_t)7));
    double* Y = G + k * GP;
    __shared__ CShared sh;
    const int tid = threadIdx.x, dl = tid / TPS, q = tid % TPS, v0 = q * CVQ;
    const T eps = (T)SAL_EPS_F32;
    const bool row = dl < Dl;
    const int d = d_lo + dl;
    unsigned int n_xch = 0;  // exchanges so far (parity = buffer)

    T x[CVQ];
#pragma unroll
    for (int i = 0; i < CVQ; ++i) x[i] = (row && v0 + i < V) ? X[(size_t)d * V + v0 + i] : (T)0;
    for (int i = tid; i < k * V; i += CNT) sW[(i / V) * WP + i % V] = W_in[i];
    for (int i = tid; i < k * WP; i += CNT) sWt[i] = (T)0, sWu[i] = (T)0, sWp[i] = (T)0;
    for (int i = tid; i < Dl * k; i += CNT) sH[i] = H_in[(size_t)d_lo * k + i];
    double gamma = *gamma_in;
    __syncthreads();

    // cluster-wide sum of a CTA-uniform scalar (+ optionally the k row sums of H'), contributions added in CTA order.
    // The contribution buffers alternate with the parity of the exchange: one cluster barrier per exchange (a CTA can only
    // overwrite a buffer two exchanges later, after everybody has passed the barrier in between).
    auto exchange = [&](double mine, bool with_hsum) {
        const int b = (int)(n_xch & 1u);
        ++n_xch;
        __syncthreads();
        if (tid == 0) sh.xch[b][0] = mine;
        // (the hsum partials were written into sh.xch[b][2 + j] by the caller)
        cluster.sync();
        double total = 0.0;
        for (int r = 0; r < C; ++r) total += cluster.map_shared_rank(&sh.xch[b][0], r)[0];
        if (with_hsum && tid < k) {
            double t = 0.0;
            for (int r = 0; r < C; ++r) t += cluster.map_shared_rank(&sh.xch[b][0], r)[2 + tid];
            sh.hsum[tid] = (double)(T)t;
        }
        return total;
    };

    // KL(X || M h) over this thread's features with float64 terms (x = 0 contributes m h)
    auto kl_row = [&](const T* M, const T (&hv)[KT]) {
        double kl = 0.0;
        if (row) {
#pragma unroll 2
            for (int i = 0; i < CVQ; ++i) {
                const int v = v0 + i;
                if (v < V) {
                    double wh = 0.0;
#pragma unroll
                    for (int j = 0; j < KT; ++j)
                        if (j < k) wh += (double)M[j * WP + v] * (double)hv[j];
                    const double xd = (double)x[i];
                    kl += xd != 0.0 ? xd * log(xd / wh) - xd + wh : wh;
                }
            }
        }
        return kl;
    };
    auto logdet = [&](const T* M) {  // CTA-uniform (and, the inputs being replicated, cluster-uniform) result
        gram<T, CNT>(M, G, GP, V, k, delta);
        return log(lu_det_block<CNT>(G, &sh.piv, GP, k));  // (every thread computes the same determinant; ends with a barrier)
    };

    double ld_W = 0.0;
    bool have_ld = false;
    const int n_pass = n_iter > 0 ? n_iter : (objective ? 1 : 0);
    for (int it = 0; it < n_pass; ++it) {
        T hd[KT];
#pragma unroll
        for (int j = 0; j < KT; ++j) hd[j] = (row && j < k) ? sH[dl * k + j] : (T)0;
        if (it == 0 && objective) {  // penalised objective of the incoming iterate
            const double kl = exchange(block_sum_c(kl_row(sW, hd), sh), false);
            ld_W = logdet(sW), have_ld = true;
            if (tid == 0 && cr == 0) *objective = kl + lam * ld_W;
        }
        if (it >= n_iter) break;  // objective-only call
        // ---- H step ----
        T hn[KT];
#pragma unroll
        for (int j = 0; j < KT; ++j) hn[j] = (T)0;
        if (row) {
#pragma unroll(KT <= 8 ? CVQ : 2)
            for (int i = 0; i < CVQ; ++i) {
                const int v = v0 + i;
                if (v < V) {
                    T wv[KT];
                    T wh = (T)0;
#pragma unroll
                    for (int j = 0; j < KT; ++j) wv[j] = j < k ? sW[j * WP + v] : (T)0, wh += wv[j] * hd[j];
                    const T r = tdiv(x[i], wh);
#pragma unroll
                    for (int j = 0; j < KT; ++j) hn[j] += wv[j] * r;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            T t = hn[j];
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 4);
            if (TPS == 16) t += __shfl_xor_sync(0xffffffffu, t, 8);
            hd[j] = (row && j < k) ? max(hd[j] * t, eps) : (T)0;  // h' (identical on the eight threads of a sample)
        }
        if (n_given >= k) {  // all signatures given: the iteration is the H step
            __syncthreads();
            if (row && q == 0) {
#pragma unroll
                for (int j = 0; j < KT; ++j)
                    if (j < k) sH[dl * k + j] = hd[j];
            }
            __syncthreads();
            continue;
        }
        // ---- quotient with the new exposures, previous KL, partial numerator and partial row sums of H' ----
        double kl_prev = 0.0;
        if (row) {
#pragma unroll(KT <= 8 ? CVQ : 2)
            for (int i = 0; i < CVQ; ++i) {
                const int v = v0 + i;
                if (v < V) {
                    T wh = (T)0;
                    double whd = 0.0;
#pragma unroll
                    for (int j = 0; j < KT; ++j)
                        if (j < k) {
                            const T w = sW[j * WP + v];
                            wh += w * hd[j];
                            if (sizeof(T) == 4) whd += (double)w * (double)hd[j];
                        }
                    sR[dl * RP + v] = tdiv(x[i], wh);
                    const double xd = (double)x[i], wd = sizeof(T) == 4 ? whd : (double)wh;
                    kl_prev += xd != 0.0 ? xd * log(xd / wd) - xd + wd : wd;
                }
            }
        }
        __syncthreads();  // every read of the old sH is done
        if (row && q == 0) {
#pragma unroll
            for (int j = 0; j < KT; ++j)
                if (j < k) sH[dl * k + j] = hd[j];
        }
        kl_prev = block_sum_c(kl_prev, sh);  // (its barriers also publish sR and the new sH)
        for (int i = tid; i < k * V; i += CNT) {
            const int j = i / V, v = i - j * V;
            T a0 = (T)0, a1 = (T)0;
            int dd = 0;
            for (; dd + 1 < Dl; dd += 2) {
                a0 += sR[dd * RP + v] * sH[dd * k + j];
                a1 += sR[(dd + 1) * RP + v] * sH[(dd + 1) * k + j];
            }
            for (; dd < Dl; ++dd) a0 += sR[dd * RP + v] * sH[dd * k + j];
            sWp[j * WP + v] = a0 + a1;
        }
        {
            const int w = tid >> 5, lane = tid & 31, b = (int)(n_xch & 1u);
            for (int j = w; j < k; j += CNT / 32) {
                double t = 0.0;
                for (int dd = lane; dd < Dl; dd += 32) t += (double)sH[dd * k + j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) sh.xch[b][2 + j] = t;
            }
        }
        kl_prev = exchange(kl_prev, true);  // (its cluster barrier also publishes every CTA's partial numerator)
        for (int i = tid; i < k * V; i += CNT) {
            const int at = (i / V) * WP + i % V;
            T t = (T)0;
            for (int r = 0; r < C; ++r) t += cluster.map_shared_rank(sWp, r)[at];
            sWu[at] = t;
        }
        __syncthreads();
        if (!have_ld) ld_W = logdet(sW), have_ld = true;
        const double prev = kl_prev + lam * ld_W;
        // ---- unconstrained W step (replicated) ----
        gram<T, CNT>(sW, G, GP, V, k, delta);
        invert_block<CNT>(G, Y, sh.col, &sh.piv, GP, k);
        for (int i = tid; i < k * V; i += CNT) {
            const int j = i / V, v = i - j * V;
            const double w = (double)sW[j * WP + v];
            double out;
            if (j < n_given) {
                out = w;
            } else {
                double wym = 0.0, wya = 0.0;
                for (int a = 0; a < k; ++a) {
                    const double y = Y[a * GP + j];
                    const double wa = (double)sW[a * WP + v];
                    wym += wa * fmax(0.0, -y);
                    wya += wa * fabs(y);
                }
                const double r = sh.hsum[j];
                const double a1 = r - 4.0 * lam * wym;
                const double s2 = 8.0 * lam * wya * (double)sWu[j * WP + v];
                const double num = sqrt(a1 * a1 + s2) + (-r + 4.0 * lam * wym);
                out = fmax(w * num / (4.0 * lam * wya), (double)SAL_EPS_F32);
            }
            sWu[j * WP + v] = (T)out;
        }
        __syncthreads();
        // ---- line search (replicated candidates, sharded KL) ----
        double g_blend = -1.0, ld_t = 0.0;
        while (true) {
            for (int i = tid; i < k * V; i += CNT) {
                const int j = i / V, v = i - j * V;
                const double wu = (double)sWu[j * WP + v];
                sWt[j * WP + v] = (T)(g_blend < 0.0 ? wu : (1.0 - g_blend) * (double)sW[j * WP + v] + g_blend * wu);
            }
            __syncthreads();
            {
                const int w = tid >> 5, lane = tid & 31;
                for (int j = w; j < k; j += CNT / 32) {
                    double t = 0.0;
                    for (int v = lane; v < V; v += 32) t += (double)sWt[j * WP + v];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    if (lane == 0) sh.colsum[j] = (double)(T)t;
                }
            }
            __syncthreads();
            for (int i = tid; i < k * V; i += CNT) {
                const int j = i / V, v = i - j * V;
                sWt[j * WP + v] = (T)fmax((double)sWt[j * WP + v] / sh.colsum[j], (double)SAL_EPS_F32);
            }
            __syncthreads();
            ld_t = logdet(sWt);
            T ht[KT];
#pragma unroll
            for (int j = 0; j < KT; ++j) ht[j] = (row && j < k) ? max(hd[j] * (T)sh.colsum[j], eps) : (T)0;
            const double val = exchange(block_sum_c(kl_row(sWt, ht), sh), false) + lam * ld_t;
            if (!(val > prev && gamma > 1e-16)) break;
            gamma *= 0.8;
            g_blend = gamma;
            __syncthreads();
        }
        gamma = fmin(1.0, 1.2 * gamma);
        ld_W = ld_t;
        for (int i = tid; i < k * V; i += CNT) sW[(i / V) * WP + i % V] = sWt[(i / V) * WP + i % V];
        if (row && q == 0) {
#pragma unroll
            for (int j = 0; j < KT; ++j)
                if (j < k) sH[dl * k + j] = max(hd[j] * (T)sh.colsum[j], eps);
        }
        __syncthreads();
    }
    if (cr == 0) {
        for (int i = tid; i < k * V; i += CNT) W_out[i] = sW[(i / V) * WP + i % V];
        if (tid == 0) *gamma_out = gamma;
    }
    for (int i = tid; i < Dl * k; i += CNT) H_out[(size_t)d_lo * k + i] = sH[i];
    cluster.sync();  // nobody leaves while a neighbour may still read its shared memory
}

template <typename T>
size_t cluster_smem(int k) {
    return sizeof(T) * ((size_t)SPC_MAX * RP + (size_t)SPC_MAX * k + 4 * (size_t)k * WP) + 8 + sizeof(double) * 2 * (size_t)k * (k + 1);
}

template <typename T, int KT, int TPS>
int launch_cluster_t(sal_ctx* c, int csize, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, double lam,
                     double delta, int n_given, int n_iter, const double* gamma_in, double* gamma_out, double* objective, cudaStream_t st) {
    const size_t smem = cluster_smem<T>(c->k);
    SAL_CUDA(cudaFuncSetAttribute(mvnmf_cluster_kernel<T, KT, TPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csize > 8) SAL_CUDA(cudaFuncSetAttribute(mvnmf_cluster_kernel<T, KT, TPS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize), cfg.blockDim = dim3(CNT), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = csize, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    const T *Xp = (const T*)X, *Wi = (const T*)W_in, *Hi = (const T*)H_in;
    T *Wo = (T*)W_out, *Ho = (T*)H_out;
    const int D = (int)c->D, V = c->V, k = c->k;
    SAL_CUDA(cudaLaunchKernelEx(&cfg, mvnmf_cluster_kernel<T, KT, TPS>, Xp, Wi, Wo, Hi, Ho, D, V, k, lam, delta, n_given, n_iter, gamma_in,
                                gamma_out, objective));
    c->launches++;
    return 0;
}

template <typename T, int TPS>
int launch_cluster_k(sal_ctx* c, int csize, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, double lam,
                     double delta, int n_given, int n_iter, const double* gamma_in, double* gamma_out, double* objective, cudaStream_t st) {
#define SAL_MVC(KT_) \
    launch_cluster_t<T, KT_, TPS>(c, csize, X, W_in, W_out, H_in, H_out, lam, delta, n_given, n_iter, gamma_in, gamma_out, objective, st)
    if (c->k <= 4) return SAL_MVC(4);
    if (c->k <= 8) return SAL_MVC(8);
    if (c->k <= 12) return SAL_MVC(12);
    if (c->k <= 16) return SAL_MVC(16);
    return SAL_MVC(32);
#undef SAL_MVC
}

template <typename T>
size_t small_smem(int D, int k) {
    return sizeof(T) * ((size_t)D * RP + (size_t)D * k + 3 * (size_t)k * WP) + 8 + sizeof(double) * 2 * (size_t)k * (k + 1);
}

template <typename T, int KT>
int launch_t(sal_ctx* c, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, double lam, double delta,
             int n_given, int n_iter, const double* gamma_in, double* gamma_out, double* objective, cudaStream_t st) {
    const size_t smem = small_smem<T>((int)c->D, c->k);
    SAL_CUDA(cudaFuncSetAttribute(mvnmf_small_kernel<T, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mvnmf_small_kernel<T, KT><<<1, NT, smem, st>>>((const T*)X, (const T*)W_in, (T*)W_out, (const T*)H_in, (T*)H_out, (int)c->D, c->V,
                                                   c->k, lam, delta, n_given, n_iter, gamma_in, gamma_out, objective);
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

template <typename T>
int launch_k(sal_ctx* c, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, double lam, double delta,
             int n_given, int n_iter, const double* gamma_in, double* gamma_out, double* objective, cudaStream_t st) {
    if (c->k <= 4) return launch_t<T, 4>(c, X, W_in, W_out, H_in, H_out, lam, delta, n_given, n_iter, gamma_in, gamma_out, objective, st);
    if (c->k <= 8) return launch_t<T, 8>(c, X, W_in, W_out, H_in, H_out, lam, delta, n_given, n_iter, gamma_in, gamma_out, objective, st);
    if (c->k <= 12) return launch_t<T, 12>(c, X, W_in, W_out, H_in, H_out, lam, delta, n_given, n_iter, gamma_in, gamma_out, objective, st);
    if (c->k <= 16) return launch_t<T, 16>(c, X, W_in, W_out, H_in, H_out, lam, delta, n_given, n_iter, gamma_in, gamma_out, objective, st);
    return launch_t<T, 32>(c, X, W_in, W_out, H_in, H_out, lam, delta, n_given, n_iter, gamma_in, gamma_out, objective, st);
}

}  // namespace

static bool single_cta_fits(const sal_ctx* c) {
    const size_t smem = c->dtype == SAL_F32 ? small_smem<float>((int)c->D, c->k) : small_smem<double>((int)c->D, c->k);
    return smem <= 220 * 1024;
}
static bool cluster_fits(const sal_ctx* c) {  // (the cluster kernel's tiles do not grow with D: 32 samples per CTA at most)
    const size_t smem = c->dtype == SAL_F32 ? cluster_smem<float>(c->k) : cluster_smem<double>(c->k);
    return c->D >= 64 && (c->D + CMAX - 1) / CMAX <= SPC_MAX && smem <= 220 * 1024;
}

bool sal_mvnmf_small_ok(const sal_ctx* c) {
    if (c->D < 1 || c->D > NT / 4) return false;
    return single_cta_fits(c) || cluster_fits(c);
}

int sal_launch_mvnmf_small(sal_ctx* c, const void* X, const void* W_in, void* W_out, const void* H_in, void* H_out, double lam,
                           double delta, int n_given, int n_iter, const double* gamma_in, double* gamma_out, double* objective,
                           cudaStream_t st) {
    // enough samples for several SMs: the cluster kernel -- 16 CTAs x 16 threads per sample (non-portable cluster size) when the
    // samples fit 16 per CTA, else 8 CTAs x 8 threads per sample; SAL_B200_MVNMF_CLUSTER = 0 / 8 / 16 forces a choice
    int csize = c->D >= 64 ? (c->D <= 16 * (CNT / 16) ? 16 : CMAX) : 1;
    if (const char* e = getenv("SAL_B200_MVNMF_CLUSTER")) {
        const int forced = atoi(e);
        if (forced <= 1) csize = 1;
        else if (forced == 8 || (forced == 16 && c->D <= 16 * (CNT / 16))) csize = forced;
    }
    if (csize == 8 && !cluster_fits(c)) csize = 1;
    if (csize == 1 && !single_cta_fits(c)) {
        if (cluster_fits(c)) {
            csize = 8;  // (the only kernel this problem fits)
        } else {
            sal_set_error("sal_mvnmf_small_updates: problem does not fit the small-problem kernels");
            return SAL_EUNSUPPORTED;
        }
    }
#define SAL_MVCL(TT_, TPS_, CS_) \
    launch_cluster_k<TT_, TPS_>(c, CS_, X, W_in, W_out, H_in, H_out, lam, delta, n_given, n_iter, gamma_in, gamma_out, objective, st)
    if (csize == 16) {  // (refused where a 16-CTA cluster has no room: fall back to the portable size)
        const int err = c->dtype == SAL_F32 ? SAL_MVCL(float, 16, 16) : SAL_MVCL(double, 16, 16);
        if (err == 0) return 0;
        (void)cudaGetLastError();
        csize = cluster_fits(c) ? 8 : 1;
    }
    if (csize == 8) return c->dtype == SAL_F32 ? SAL_MVCL(float, 8, 8) : SAL_MVCL(double, 8, 8);
#undef SAL_MVCL
    return c->dtype == SAL_F32
               ? launch_k<float>(c, X, W_in, W_out, H_in, H_out, lam, delta, n_given, n_iter, gamma_in, gamma_out, objective, st)
               : launch_k<double>(c, X, W_in, W_out, H_in, H_out, lam, delta, n_given, n_iter, gamma_in, gamma_out, objective, st);
}
