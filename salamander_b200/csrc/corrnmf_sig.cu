// Correlated NMF, signature embeddings (reference models/_utils_corrnmf.py:182-410 as called from corrnmf_det.py:88-113): one
// thread-block cluster per signature, every objective / gradient / Hessian evaluation of its Newton-CG is a fixed-order block
// (and cluster) reduction over the samples.  Own translation unit: the Newton-CG is instantiated per embedding dimension.
#include <cooperative_groups.h>

#include <stdlib.h>
#include <string.h>

#include <vector>

#include "corrnmf_newton.cuh"

namespace cg = cooperative_groups;

namespace {

// ---------------------------------------------------------------------------------------------------------
// signature embeddings: one thread-block cluster per signature; evaluations are reductions over the samples
// ---------------------------------------------------------------------------------------------------------
constexpr int SIG_THREADS = 256;
constexpr int SIG_MAX_GPUS = 8, SIG_MAX_VIRTUAL = 2;
constexpr int NVS = 1 + MAXM + MAXM * (MAXM + 1) / 2;  // values of one evaluation: f | gradient | upper triangle of the Hessian
constexpr unsigned int SIG_TAG_SHIFT = 17;             // tag = launch number << 17 | evaluation (<= 200 m * 99 < 2^17)

__device__ __forceinline__ unsigned long long sig_global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// sequence-tagged 16-byte words {lo, tag, hi, tag}: data and flag travel together (same protocol as the KL-NMF period kernel)
__device__ __forceinline__ void sig_push(uint4* p, double v, unsigned int tag) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)bits), "r"(tag), "r"((unsigned int)(bits >> 32)),
                 "r"(tag)
                 : "memory");
}
__device__ __forceinline__ uint4 sig_peek(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <typename T, int M>
struct SignatureProblem {
    const T *U, *b, *auxT;  // U [D][m], b [D], auxT [D][k]: this GPU's samples
    double* red;            // shared [SIG_THREADS / 32][NVS]
    double* cl;             // shared [2][NVS]: this CTA's totals of the even / odd evaluations, read by the other CTAs of the cluster
    double* tot;            // shared [2][NVS]: the evaluation's totals, broadcast to the threads of this CTA
    double s, inv_var;
    int64_t D;
    int k, m_rt, j;
    int rank, nrank;        // position in the thread-block cluster that shares signature j (samples are interleaved)
    // several GPUs: every GPU sweeps its own samples and the totals of an evaluation are exchanged over NVLink as tagged words
    // (pushed into every peer's receive buffer by CTA 0 of the cluster, polled locally by every CTA, summed in GPU order: all
    // GPUs obtain the same bits and their Newton-CGs stay in lock step)
    uint4* const* peers;    // [n_gpus] receive buffers [2 slots][n_gpus][k][NVS], or null
    int n_gpus, gpu;
    unsigned int tag0, evals;
    static constexpr int MM = M > 0 ? M : MAXM;
    static constexpr int NV = 1 + MM + MM * (MM + 1) / 2;

    // n values per thread -> their sums over all threads of the cluster (and all GPUs), in a fixed order, in every thread
    __device__ void reduce(double* v, int n) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, tid = threadIdx.x;
        const int buf = (int)(evals & 1u);
        ++evals;
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (i < n) {
                double t = v[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if (lane == 0) red[w * NVS + i] = t;
            }
        __syncthreads();
        double t = 0.0;
        if (tid < n) {
            for (int ww = 0; ww < SIG_THREADS / 32; ++ww) t += red[ww * NVS + tid];
            cl[buf * NVS + tid] = t;
        }
        if (nrank > 1) {
            // the CTAs' totals in rank order through distributed shared memory.  `cl` is double-buffered by the parity of the
            // evaluation, so ONE cluster barrier per evaluation is enough: a CTA can only overwrite its buffer two evaluations later,
            // after everybody has passed the barrier in between -- i.e. has finished reading this one
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();
            if (tid < n) {
                t = 0.0;
                for (int r = 0; r < nrank; ++r) t += cluster.map_shared_rank(cl, r)[buf * NVS + tid];
            }
        }
        if (n_gpus > 1 && tid < n) {
            const unsigned int tag = tag0 + evals;
            const size_t per_src = (size_t)k * NVS, slot = (size_t)(tag & 1u) * n_gpus;
            const size_t at = (size_t)j * NVS + tid;
            if (rank == 0) {
                for (int p = 0; p < n_gpus; ++p)
                    if (p != gpu) sig_push(peers[p] + (slot + gpu) * per_src + at, t, tag);
            }
            // batched polling: every missing word is asked for again in the same round
            const uint4* mine = peers[gpu] + slot * per_src + at;
            double got[SIG_MAX_GPUS];
            unsigned int pending = ((1u << n_gpus) - 1u) & ~(1u << gpu), spins = 0;
            unsigned long long t0 = 0;
            while (pending) {
#pragma unroll
                for (int p = 0; p < SIG_MAX_GPUS; ++p)
                    if ((pending >> p) & 1u) {
                        const uint4 wd = sig_peek(mine + (size_t)p * per_src);
                        if (wd.y == tag && wd.w == tag) {
                            got[p] = __longlong_as_double((long long)(((unsigned long long)wd.z << 32) | wd.x));
                            pending &= ~(1u << p);
                        }
                    }
                if (pending && (++spins & 255u) == 0) {
                    const unsigned long long now = sig_global_ns();
                    if (t0 == 0) t0 = now;
                    if (now - t0 > 20ull * 1000 * 1000 * 1000) __trap();  // a lost peer traps instead of hanging the box
                }
            }
            double sum = 0.0;
#pragma unroll
            for (int p = 0; p < SIG_MAX_GPUS; ++p)
                if (p < n_gpus) sum += p == gpu ? t : got[p];
            t = sum;
        }
        if (tid < n) tot[buf * NVS + tid] = t;
        __syncthreads();  // (also: `red` may be rewritten; `tot` is double-buffered like `cl`)
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (i < n) v[i] = tot[buf * NVS + i];
    }

    // (not inlined: the Newton-CG calls it from two places, and the sweep is the whole cost of this kernel anyway)
    __device__ __noinline__ double f_grad_hess(const double* x, double* g, double* A) {
        const int m = M > 0 ? M : m_rt;
        const int n = 1 + m + m * (m + 1) / 2;
        double v[NV];  // [f | gradient | upper triangle of the Hessian]: one collective reduction for all of it
#pragma unroll
        for (int q = 0; q < NV; ++q) v[q] = 0.0;
        auto accumulate = [&](const double (&u)[MM], double sp, double e, double ax) {
            v[0] += sp * ax - e;
            const double w = e - ax;
            _Pragma("unroll") for (int q = 0; q < m; ++q) v[1 + q] += w * u[q];
            int at = 1 + m;
            _Pragma("unroll") for (int p = 0; p < m; ++p) {
                const double eu = e * u[p];
                _Pragma("unroll") for (int q = p; q < m; ++q) v[at++] += eu * u[q];
            }
        };
        const int64_t stride = (int64_t)nrank * SIG_THREADS;
        int64_t d = (int64_t)rank * SIG_THREADS + threadIdx.x;
        if (M > 0) {
            // four samples per trip: their loads and the four exp() chains overlap (with 8 warps per SM the sweep is otherwise
            // bound by the latency of one sample's load -> exp -> accumulate chain)
            constexpr int IL = 4;
            for (; d + (IL - 1) * stride < D; d += IL * stride) {
                double u[IL][MM], sp[IL], bb[IL], ax[IL], e[IL];
#pragma unroll
                for (int z = 0; z < IL; ++z) {
                    const int64_t dz = d + z * stride;
                    _Pragma("unroll") for (int q = 0; q < MM; ++q) u[z][q] = (double)U[dz * MM + q];
                    bb[z] = (double)b[dz], ax[z] = (double)auxT[dz * k + j];
                }
#pragma unroll
                for (int z = 0; z < IL; ++z) {
                    sp[z] = 0.0;
                    _Pragma("unroll") for (int q = 0; q < MM; ++q) sp[z] += u[z][q] * x[q];
                    e[z] = exp(s + bb[z] + sp[z]);
                }
#pragma unroll
                for (int z = 0; z < IL; ++z) accumulate(u[z], sp[z], e[z], ax[z]);
            }
        }
        for (; d < D; d += stride) {
            double u[MM], sp = 0.0;
            _Pragma("unroll") for (int q = 0; q < MM; ++q)
                if (q < m) u[q] = (double)U[d * m + q], sp += u[q] * x[q];
            accumulate(u, sp, exp(s + (double)b[d] + sp), (double)auxT[d * k + j]);
        }
        reduce(v, n);
        double nrm = 0.0;
        _Pragma("unroll") for (int q = 0; q < MM; ++q)
            if (q < m) nrm += x[q] * x[q], g[q] = v[1 + q] + x[q] * inv_var;
        int at = 1 + m;
        _Pragma("unroll") for (int p = 0; p < MM; ++p)
            _Pragma("unroll") for (int q = p; q < MM; ++q)
                if (q < m) {
                    const double h = v[at++];
                    A[p * m + q] = h, A[q * m + p] = h;
                }
        _Pragma("unroll") for (int q = 0; q < MM; ++q)
            if (q < m) A[q * m + q] += inv_var;
        return -(v[0] - 0.5 * nrm * inv_var);
    }
};

struct SigRank {
    const void *auxT, *a, *b, *U;
    void* L;
    const void* peers;  // device array [n_gpus] of receive buffers, or null
    int64_t D;
    int gpu;
};
struct SigParams {
    SigRank r[SIG_MAX_VIRTUAL];  // one entry; two when ranks are emulated on one GPU (tests)
    double variance;
    int k, m, sig_begin, sig_count, n_gpus;
    unsigned int tag0;
};

template <typename T, int M>
__global__ void __launch_bounds__(SIG_THREADS) signature_embeddings_kernel(const __grid_constant__ SigParams P) {
    constexpr int MM = M > 0 ? M : MAXM;
    const int m = M > 0 ? M : P.m;
    __shared__ double red[(SIG_THREADS / 32) * NVS];
    __shared__ double cl[2 * NVS];
    __shared__ double tot[2 * NVS];
    cg::cluster_group cluster = cg::this_cluster();
    const int nrank = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / nrank, vr = cid / P.sig_count;  // one cluster per (virtual rank, signature)
    const SigRank& R = P.r[vr];
    const int j = P.sig_begin + cid - vr * P.sig_count;  // every CTA of the cluster runs the same Newton-CG on the same numbers
    T* L = (T*)R.L;
    double x[MM];
#pragma unroll
    for (int q = 0; q < MM; ++q)
        if (q < m) x[q] = (double)L[j * m + q];
    SignatureProblem<T, M> p{(const T*)R.U, (const T*)R.b, (const T*)R.auxT, red, cl, tot, (double)((const T*)R.a)[j], 1.0 / P.variance, R.D, P.k, m, j,
                             rank, nrank, (uint4* const*)R.peers, R.peers ? P.n_gpus : 1, R.gpu, P.tag0, 0u};
    if (nrank > 1) cluster.sync();  // nobody writes L[j] before everybody has read it
    newton_cg<M>(p, x, m, 200 * m);
    if (rank == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < MM; ++q)
            if (q < m) L[j * m + q] = (T)snap_eps(x[q]);
    }
    if (nrank > 1) cluster.sync();  // nobody leaves while a neighbour may still read its totals
}

// ---------------------------------------------------------------------------------------------------------
// small all-reduce over peer memory: the few-KB sums of a CorrNMF iteration (W numerator, scaling sums, norms, likelihood)
// ---------------------------------------------------------------------------------------------------------
// Thread i pushes buf[i] as a tagged word into every peer's receive buffer [2 slots][n_gpus][P2P_AR_MAX], polls its own GPU's
// buffer and replaces buf[i] by the sum of the contributions in rank order (bit-identical on all ranks).  One tag per call
// (the callers' launch number, counted alike on every rank); two slots by tag parity suffice: a rank finishes call t only
// after it has read every peer's words of call t, and pushes the words of call t + 1 only afterwards.
constexpr int P2P_AR_MAX = 4096;
struct AllreduceParams {
    double* buf[SIG_MAX_VIRTUAL];
    const void* peers[SIG_MAX_VIRTUAL];
    int gpu[SIG_MAX_VIRTUAL];
    int n, n_gpus;
    unsigned int tag;
};
__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const __grid_constant__ AllreduceParams P) {
    const int vr = blockIdx.y, i = (int)blockIdx.x * 256 + (int)threadIdx.x;
    if (i >= P.n) return;
    uint4* const* peers = (uint4* const*)P.peers[vr];
    const int gpu = P.gpu[vr], n_gpus = P.n_gpus;
    const unsigned int tag = P.tag;
    const double mine = P.buf[vr][i];
    const size_t slot = (size_t)(tag & 1u) * n_gpus;
    for (int p = 0; p < n_gpus; ++p)
        if (p != gpu) sig_push(peers[p] + (slot + gpu) * P2P_AR_MAX + i, mine, tag);
    const uint4* own = peers[gpu] + slot * P2P_AR_MAX + i;
    double got[SIG_MAX_GPUS];
    unsigned int pending = ((1u << n_gpus) - 1u) & ~(1u << gpu), spins = 0;
    unsigned long long t0 = 0;
    while (pending) {
#pragma unroll
        for (int p = 0; p < SIG_MAX_GPUS; ++p)
            if ((pending >> p) & 1u) {
                const uint4 wd = sig_peek(own + (size_t)p * P2P_AR_MAX);
                if (wd.y == tag && wd.w == tag) {
                    got[p] = __longlong_as_double((long long)(((unsigned long long)wd.z << 32) | wd.x));
                    pending &= ~(1u << p);
                }
            }
        if (pending && (++spins & 255u) == 0) {
            const unsigned long long now = sig_global_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > 20ull * 1000 * 1000 * 1000) __trap();
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int p = 0; p < SIG_MAX_GPUS; ++p)
        if (p < n_gpus) sum += p == gpu ? mine : got[p];
    P.buf[vr][i] = sum;
}

}  // namespace

size_t sal_p2p_allreduce_words(int n_gpus) { return (size_t)2 * n_gpus * P2P_AR_MAX; }
int sal_p2p_allreduce_max(void) { return P2P_AR_MAX; }

int sal_launch_p2p_allreduce(int n_virtual, double* const* bufs, const void* const* peer_tables, const int* gpus, int n, int n_gpus,
                             unsigned int launch_id, cudaStream_t st) {
    if (n <= 0) return 0;
    if (n > P2P_AR_MAX || n_virtual < 1 || n_virtual > SIG_MAX_VIRTUAL || n_gpus < 2 || n_gpus > SIG_MAX_GPUS) {
        sal_set_error("p2p all-reduce: %d values (<= %d), %d emulated ranks (<= %d), %d GPUs (2 .. %d)", n, P2P_AR_MAX, n_virtual, SIG_MAX_VIRTUAL,
                      n_gpus, SIG_MAX_GPUS);
        return SAL_EINVAL;
    }
    AllreduceParams P;
    memset(&P, 0, sizeof(P));
    for (int v = 0; v < n_virtual; ++v) P.buf[v] = bufs[v], P.peers[v] = peer_tables[v], P.gpu[v] = gpus[v];
    P.n = n, P.n_gpus = n_gpus, P.tag = launch_id;
    p2p_allreduce_kernel<<<dim3((n + 255) / 256, n_virtual), 256, 0, st>>>(P);
    SAL_CUDA(cudaGetLastError());
    return 0;
}

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

// Largest cluster size of `sizes` (descending) for which `need` clusters of this kernel are resident at once; cached per device,
// kernel and count (the occupancy query can take tens of milliseconds).  0 when the query fails for every size.
static int resident_cluster_size(const void* fn, int device, int key, int need, const int* sizes, int n_sizes) {
    struct Entry { const void* fn; int device, need, key, answer; };
    static std::vector<Entry> cache;
    for (const Entry& e : cache)
        if (e.fn == fn && e.device == device && e.need == need && e.key == key) return e.answer;
    int answer = 0;
    for (int i = 0; i < n_sizes && !answer; ++i) {
        const int cs = sizes[i];
        if (cs == 1) {
            answer = 1;
            break;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(SIG_THREADS), cfg.gridDim = dim3(need * cs);
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = cs, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
        cfg.attrs = &attr, cfg.numAttrs = 1;
        int n_active = 0;
        const bool allowed = cs <= 8 || cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        if (allowed && cudaOccupancyMaxActiveClusters(&n_active, fn, &cfg) == cudaSuccess && n_active >= need) answer = cs;
        (void)cudaGetLastError();  // a refused query is not an error of this call: try the next size
    }
    cache.push_back({fn, device, need, key, answer});
    return answer;
}

int sal_launch_corrnmf_signature_embeddings_v(const SigLaunchRank* rs, int n_virtual, int m, double variance, int sig_begin, int sig_count,
                                              int n_gpus, unsigned int launch_id, cudaStream_t st) {
    if (sig_count <= 0) return 0;
    sal_ctx* c = rs[0].c;
    if (n_virtual < 1 || n_virtual > SIG_MAX_VIRTUAL || n_gpus < 1 || n_gpus > SIG_MAX_GPUS) {
        sal_set_error("signature embeddings: %d emulated ranks / %d GPUs not supported (<= %d / <= %d)", n_virtual, n_gpus, SIG_MAX_VIRTUAL, SIG_MAX_GPUS);
        return SAL_EINVAL;
    }
    const bool exchanging = n_gpus > 1;
    const void* fn = sig_kernel(c->dtype, m);
    int64_t D_min = rs[0].c->D;
    for (int v = 1; v < n_virtual; ++v) D_min = rs[v].c->D < D_min ? rs[v].c->D : D_min;
    // one thread-block cluster per signature: 8 CTAs share the sums over samples once there is enough work for them, 16
    // (non-portable size: one cluster per GPC) when there are few signatures and a lot of samples.  With an exchange every
    // cluster waits for its counterparts on the peers: all clusters of the launch must then be resident together.
    const int need = sig_count * n_virtual;
    int csize = D_min >= 8 * 4 * SIG_THREADS ? 8 : 1;
    if (need <= 8 && D_min >= 16 * 8 * SIG_THREADS) csize = 16;
    if (const char* e = getenv("SAL_B200_SIG_CLUSTER")) {  // diagnostics: cap the cluster size (1, 2, 4, 8 or 16)
        const int forced = atoi(e);
        if (forced >= 1 && forced <= csize) csize = forced;
    }
    if (csize == 16 || exchanging) {
        const int all[] = {16, 8, 4, 2, 1};
        int first = 0;
        while (all[first] > csize) ++first;
        csize = resident_cluster_size(fn, c->device, (c->dtype == SAL_F32 ? 0 : 64) + m, need, all + first, 5 - first);
        if (csize < 1) {
            sal_set_error("signature embeddings: %d clusters cannot be resident together on this device", need);
            return SAL_EUNSUPPORTED;
        }
    }
    SigParams P;
    memset(&P, 0, sizeof(P));
    for (int v = 0; v < n_virtual; ++v) {
        SigRank& R = P.r[v];
        R.auxT = rs[v].auxT, R.a = rs[v].a, R.b = rs[v].b, R.U = rs[v].U, R.L = rs[v].L;
        R.peers = exchanging ? rs[v].peers : nullptr, R.D = rs[v].c->D, R.gpu = rs[v].gpu;
    }
    P.variance = variance, P.k = c->k, P.m = m, P.sig_begin = sig_begin, P.sig_count = sig_count, P.n_gpus = n_gpus;
    P.tag0 = launch_id << SIG_TAG_SHIFT;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(SIG_THREADS), cfg.gridDim = dim3(need * csize), cfg.dynamicSmemBytes = 0, cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = csize, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    void* args[] = {(void*)&P};
    SAL_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    SAL_CUDA(cudaGetLastError());
    for (int v = 0; v < n_virtual; ++v) rs[v].c->launches++;
    return 0;
}

int sal_launch_corrnmf_signature_embeddings(sal_ctx* c, const void* auxT, const void* a, const void* b, void* L, const void* U, int m,
                                            double variance, int sig_begin, int sig_count, cudaStream_t st) {
    const SigLaunchRank r = {c, auxT, a, b, U, L, nullptr, 0};
    return sal_launch_corrnmf_signature_embeddings_v(&r, 1, m, variance, sig_begin, sig_count, 1, 0u, st);
}

size_t sal_corrnmf_sig_exchange_words(int k, int n_gpus) { return (size_t)2 * n_gpus * k * NVS; }
