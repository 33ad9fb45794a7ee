// Correlated NMF, signature embeddings (reference models/_utils_corrnmf.py:182-410 as called from corrnmf_det.py:88-113): one
// thread-block cluster per signature, every objective / gradient / Hessian evaluation of its Newton-CG is a fixed-order block
// (and cluster) reduction over the samples.  Own translation unit: the Newton-CG is instantiated per embedding dimension.
#include <cooperative_groups.h>

#include <stdlib.h>
#include <string.h>

#include "corrnmf_newton.cuh"

namespace cg = cooperative_groups;

namespace {

// ---------------------------------------------------------------------------------------------------------
// signature embeddings: one CTA per signature; evaluations are block reductions over the samples
// ---------------------------------------------------------------------------------------------------------
constexpr int SIG_THREADS = 256;

template <typename T, int M>
struct SignatureProblem {
    const T *U, *b, *auxT;  // U [D][m], b [D], auxT [D][k]
    double* red;            // shared [SIG_THREADS / 32][1 + MAXM + MAXM * MAXM]
    double* cl;             // shared [1 + MAXM + MAXM * MAXM]: this CTA's totals, read by the other CTAs of the cluster
    double s, inv_var;
    int64_t D;
    int k, m_rt, j;
    int rank, nrank;        // position in the thread-block cluster that shares signature j (samples are interleaved)
    static constexpr int MM = M > 0 ? M : MAXM;
    static constexpr int NV = 1 + MM + MM * MM;

    // fixed-order block reduction of n values per thread; result broadcast to all threads through shared memory
    __device__ void reduce(double* v, int n) const {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, stride = 1 + MAXM + MAXM * MAXM;
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (i < n) {
                double t = v[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) red[w * stride + i] = t;
            }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (i < n) {
                double t = 0.0;
                for (int ww = 0; ww < SIG_THREADS / 32; ++ww) t += red[ww * stride + i];
                v[i] = t;
            }
        __syncthreads();
        if (nrank > 1) {  // sum the CTAs' totals in rank order through distributed shared memory
            cg::cluster_group cluster = cg::this_cluster();
            if (threadIdx.x == 0) {
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (i < n) cl[i] = v[i];
            }
            cluster.sync();
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (i < n) {
                    double t = 0.0;
                    for (int r = 0; r < nrank; ++r) t += cluster.map_shared_rank(cl, r)[i];
                    v[i] = t;
                }
            cluster.sync();
        }
    }
    // (not inlined: the Newton-CG calls it from two places, and the sweep is the whole cost of this kernel anyway)
    __device__ __noinline__ double f_grad_hess(const double* x, double* g, double* A) const {
        const int m = M > 0 ? M : m_rt;
        double v[NV];  // [f | gradient | Hessian]: one collective reduction for all of it
        const int n = 1 + m + m * m;
#pragma unroll
        for (int q = 0; q < NV; ++q) v[q] = 0.0;
        for (int64_t d = (int64_t)rank * SIG_THREADS + threadIdx.x; d < D; d += (int64_t)nrank * SIG_THREADS) {
            double u[MM], sp = 0.0;
            _Pragma("unroll") for (int q = 0; q < m; ++q) u[q] = (double)U[d * m + q], sp += u[q] * x[q];
            const double e = exp(s + (double)b[d] + sp), ax = (double)auxT[d * k + j];
            v[0] += sp * ax - e;
            const double w = e - ax;
            _Pragma("unroll") for (int q = 0; q < m; ++q) v[1 + q] += w * u[q];
            _Pragma("unroll") for (int p = 0; p < m; ++p)
                _Pragma("unroll") for (int q = 0; q < m; ++q) v[1 + m + p * m + q] += e * u[p] * u[q];
        }
        reduce(v, n);
        double nrm = 0.0;
        _Pragma("unroll") for (int q = 0; q < m; ++q) nrm += x[q] * x[q], g[q] = v[1 + q] + x[q] * inv_var;
        _Pragma("unroll") for (int q = 0; q < m * m; ++q) A[q] = v[1 + m + q];
        _Pragma("unroll") for (int q = 0; q < m; ++q) A[q * m + q] += inv_var;
        return -(v[0] - 0.5 * nrm * inv_var);
    }
};

template <typename T, int M>
__global__ void __launch_bounds__(SIG_THREADS) signature_embeddings_kernel(const T* auxT, const T* a, const T* b, T* L, const T* U,
                                                                          int64_t D, int k, int m_rt, double variance, int sig_begin) {
    constexpr int MM = M > 0 ? M : MAXM;
    const int m = M > 0 ? M : m_rt;
    __shared__ double red[(SIG_THREADS / 32) * (1 + MAXM + MAXM * MAXM)];
    __shared__ double cl[1 + MAXM + MAXM * MAXM];
    cg::cluster_group cluster = cg::this_cluster();
    const int nrank = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int j = sig_begin + blockIdx.x / nrank;  // one cluster per signature; every CTA of it runs the same Newton-CG on the same numbers
    double x[MM];
#pragma unroll
    for (int q = 0; q < MM; ++q)
        if (q < m) x[q] = (double)L[j * m + q];
    SignatureProblem<T, M> p{U, b, auxT, red, cl, (double)a[j], 1.0 / variance, D, k, m, j, rank, nrank};
    if (nrank > 1) cluster.sync();  // nobody writes L[j] before everybody has read it
    newton_cg<M>(p, x, m, 200 * m);
    if (rank == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < MM; ++q)
            if (q < m) L[j * m + q] = (T)snap_eps(x[q]);
    }
}

}  // namespace

// the signature-embedding kernel for (dtype, embedding dimension): dimension 2 .. 5 as a template constant, 0 = run time
static const void* sig_kernel(int dtype, int m) {
#define SAL_SIG(M_) ((const void*)signature_embeddings_kernel<double, M_>)
    if (dtype == SAL_F32) return (const void*)signature_embeddings_kernel<float, 0>;  // float storage: run-time dimension only
    switch (m) {
        case 2: return SAL_SIG(2);
        case 3: return SAL_SIG(3);
        case 4: return SAL_SIG(4);
        case 5: return SAL_SIG(5);
        default: return SAL_SIG(0);
    }
#undef SAL_SIG
}

int sal_launch_corrnmf_signature_embeddings(sal_ctx* c, const void* auxT, const void* a, const void* b, void* L, const void* U, int m,
                                            double variance, int sig_begin, int sig_count, cudaStream_t st) {
    if (sig_count <= 0) return 0;
    // one thread-block cluster per signature: 8 CTAs share the sums over samples once there is enough work for them, 16
    // (non-portable size: one cluster per GPC, 8 GPCs) when there are few signatures and a lot of samples
    int csize = c->D >= 8 * 4 * SIG_THREADS ? 8 : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(SIG_THREADS), cfg.dynamicSmemBytes = 0, cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    if (sig_count <= 8 && c->D >= 16 * 8 * SIG_THREADS) {
        // (the occupancy query can take tens of milliseconds: asked once per device, dtype and signature count)
        static signed char cached[64][2][9][8];
        static bool cached_init = false;
        if (!cached_init) memset(cached, -1, sizeof(cached)), cached_init = true;
        signed char& ok16 = cached[c->device & 63][c->dtype == SAL_F32 ? 0 : 1][sig_count][(c->dtype != SAL_F32 && m >= 2 && m <= 5) ? m : 0];
        if (ok16 < 0) {
            const void* fn = sig_kernel(c->dtype, m);
            int n_active = 0;
            attr.val.clusterDim.x = 16;
            cfg.gridDim = dim3(sig_count * 16);
            ok16 = (cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
                    cudaOccupancyMaxActiveClusters(&n_active, fn, &cfg) == cudaSuccess && n_active >= sig_count)
                       ? 1
                       : 0;
            (void)cudaGetLastError();  // a refused query is not an error of this call: fall back to the portable size
        }
        if (ok16 == 1) csize = 16;
    }
    if (const char* e = getenv("SAL_B200_SIG_CLUSTER")) {  // diagnostics: force the cluster size (1, 8 or 16)
        const int forced = atoi(e);
        if (forced == 1 || forced == 8 || (forced == 16 && csize == 16)) csize = forced;
    }
    attr.val.clusterDim.x = csize;
    cfg.gridDim = dim3(sig_count * csize);
    const int64_t D = c->D;
    const int k = c->k;
    {
        const void* fn = sig_kernel(c->dtype, m);
        void* args[] = {(void*)&auxT, (void*)&a, (void*)&b, (void*)&L, (void*)&U, (void*)&D, (void*)&k, (void*)&m, (void*)&variance, (void*)&sig_begin};
        SAL_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    }
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

