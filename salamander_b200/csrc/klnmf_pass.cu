// Fused KL-NMF pass, CUDA-core FMA flavour (exact fp32 / fp64 arithmetic).
//
// One launch streams the resident count matrix X [D][V] once.  Per sample d it rebuilds
// (WH)[:,d] from W (pinned in shared memory) and H[d][:] (registers), forms the quotient
// A = X/(WH), and accumulates -- depending on flags --
//     H_out[d][:]  = clip(H * W^T A)            update_H / update_WH   _utils_klnmf.py:258-278,343-361
//     Wnum[k][v]  += wkl_d A[v,d] H[k,d]         update_W / update_WH   _utils_klnmf.py:207-212,328-338
//     objective   += wkl_d KL(x_d || W h_d)      kl_divergence          _utils_klnmf.py:37-55
// (WH) and A never touch HBM.  Algorithmic traffic: V*D + 2*k*D reals per launch.
//
// Work decomposition (192 threads = 6 warps per CTA, persistent grid):
//   stage : cooperative coalesced copy of a tile of TS samples (X rows, H rows) to shared memory
//   phase1: thread <-> sample.  h[KP], hn[KP] in registers; W rows are broadcast 128-bit LDS.
//           The quotient tile overwrites the X tile in shared memory.
//   phase2: thread <-> (feature v, half of the signatures).  Accumulates sum_s A[s][v] * H[s][k]
//           over the tile in registers that persist across all tiles of the CTA.
//   end   : per-CTA partial numerators / objective go to a scratch buffer; a second tiny kernel
//           sums them in a FIXED order (deterministic, rank-independent).
#include "sal_common.cuh"

namespace {

constexpr int NT = 192;
constexpr int VP = SAL_VMAX;

template <typename T>
struct Cfg;
template <>
struct Cfg<float> {
    static constexpr int TS = 192, XP = 100, HPAD = 4, VEC = 4, OCC = 2;
};
template <>
struct Cfg<double> {
    static constexpr int TS = 96, XP = 98, HPAD = 2, VEC = 2, OCC = 1;
};

// CTAs per SM: float64 used to run ONE 192-thread CTA per SM of which only 96 threads work in the sample phase -- three warps
// cannot keep the FP64 pipe busy.  Up to 16 signatures two CTAs fit the shared memory (2 x 101 KB).
template <typename T, int KP>
__host__ __device__ constexpr int pass_occ() {
    return sizeof(T) == 4 ? Cfg<T>::OCC : (KP <= 16 ? 2 : 1);
}

template <typename T, int N>
struct alignas(16) Vec {
    T v[N];
};

template <typename T>
struct PassParams {
    const T *X, *W, *H_in, *w_kl, *w_lhalf, *h_scale;
    T *H_out, *partial_wnum, *per_sample;
    double *partial_obj, *partial_hsum;
    int64_t D;
    int V, k, flags, vec_x, vec_h;
};

__device__ __forceinline__ float sal_div(float a, float b) { return __fdividef(a, b); }
// x / y in float64 through a reciprocal seed (>= 20 bits), two Newton steps and one residual correction of the quotient: a few
// DFMAs instead of the ~40-instruction IEEE division sequence, within 1 ulp of the correctly rounded quotient (far inside the
// 1e-9 trajectory tolerance; the single-CTA kernels do the same).  y > 0 and normal here (WH >= k * eps^2).
__device__ __forceinline__ double sal_div(double a, double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    const double q = a * r;
    return fma(fma(-b, q, a), r, q);
}
__device__ __forceinline__ float sal_log(float a) { return logf(a); }
__device__ __forceinline__ double sal_log(double a) { return log(a); }
__device__ __forceinline__ float sal_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ double sal_sqrt(double a) { return sqrt(a); }
__device__ __forceinline__ float sal_fma(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double sal_fma(double a, double b, double c) { return fma(a, b, c); }

__device__ __forceinline__ double block_sum_192(double x, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = x;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) t += s_red[w];
    return t;
}

template <typename T, int KP>
__global__ void __launch_bounds__(NT, pass_occ<T, KP>()) klnmf_pass_kernel(PassParams<T> p) {
    using C = Cfg<T>;
    constexpr int TS = C::TS, XP = C::XP, HP = KP + C::HPAD, VEC = C::VEC, KH = KP / 2;
    using V16 = Vec<T, VEC>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sW = reinterpret_cast<T*>(smem_raw);  // [VP][KP]   sW[v][j] = W[j][v]
    T* sH = sW + VP * KP;                    // [TS][HP]
    T* sX = sH + TS * HP;                    // [TS][XP]   counts, overwritten by the quotient
    __shared__ double s_red[NT / 32];

    const int tid = threadIdx.x;
    const int V = p.V, k = p.k;
    const T eps = (T)SAL_EPS_F32;
    const bool do_h = p.flags & SAL_PASS_UPDATE_H, do_w = p.flags & SAL_PASS_WNUM;
    const bool do_kl = p.flags & (SAL_PASS_OBJECTIVE | SAL_PASS_SAMPLEWISE);
    const bool do_pois = p.flags & SAL_PASS_POISSON;
    const bool do_hsum = p.flags & SAL_PASS_HSUM;

    for (int i = tid; i < VP * KP; i += NT) {
        const int v = i / KP, j = i - v * KP;
        sW[i] = (v < V && j < k) ? p.W[(size_t)j * V + v] : (T)0;
    }

    const int pv = tid % VP, kh = tid / VP;  // phase-2 identity
    T acc[KH];
    T hs_acc[KH];
#pragma unroll
    for (int j = 0; j < KH; ++j) acc[j] = (T)0, hs_acc[j] = (T)0;
    double obj_acc = 0.0;

    const int64_t n_tiles = (p.D + TS - 1) / TS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t d0 = tile * TS;
        const int n_valid = (int)min((int64_t)TS, p.D - d0);
        __syncthreads();  // previous tile's phase 2 done (and sW ready on the first trip)

        // ---- stage H tile: sH[s][j] = H_in[d0+s][j] (optionally scaled+clipped), zero padded
        {
            const T* Hg = p.H_in + (size_t)d0 * k;
            for (int i = tid; i < TS * KP; i += NT) {
                const int s = i / KP, j = i - s * KP;
                T val = (T)0;
                if (s < n_valid && j < k) {
                    val = Hg[s * k + j];
                    if (p.h_scale) val = max(val * p.h_scale[j], eps);
                }
                sH[s * HP + j] = val;
            }
        }
        // ---- stage X tile (coalesced; 128-bit when rows are 16-byte aligned)
        {
            const T* Xg = p.X + (size_t)d0 * V;
            if (p.vec_x) {
                const int vpr = V / VEC;  // vectors per row
                const V16* Xv = reinterpret_cast<const V16*>(Xg);
                for (int i = tid; i < n_valid * vpr; i += NT) {
                    const int s = i / vpr, c = i - s * vpr;
                    *reinterpret_cast<V16*>(sX + s * XP + c * VEC) = Xv[i];
                }
            } else {
                for (int i = tid; i < n_valid * V; i += NT) {
                    const int s = i / V, v = i - s * V;
                    sX[s * XP + v] = Xg[i];
                }
                const int vr = (V + VEC - 1) / VEC * VEC;
                if (vr != V)
                    for (int i = tid; i < n_valid * (vr - V); i += NT) {
                        const int s = i / (vr - V), v = V + (i - s * (vr - V));
                        sX[s * XP + v] = (T)0;
                    }
            }
        }
        __syncthreads();

        // ---- phase 1: one sample per thread (float) / per PAIR of adjacent threads (double: the tile has 96 samples for 192
        // threads; each thread of a pair takes half of the features and the pair adds its partial H numerators and KL terms
        // with one shuffle -- otherwise half of the CTA idles through the most expensive phase)
        constexpr int SPLIT = sizeof(T) == 8 ? 2 : 1;
        T hn[KP];
        if (tid < TS * SPLIT) {
            const int s = tid / SPLIT, half = tid % SPLIT;
            const int64_t d = d0 + s;
            T* xrow = sX + s * XP;
            if (s < n_valid) {
                T h[KP];
#pragma unroll
                for (int j = 0; j < KP; j += VEC) {
                    const V16 t = *reinterpret_cast<const V16*>(sH + s * HP + j);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) h[j + e] = t.v[e], hn[j + e] = (T)0;
                }
                const T wk = p.w_kl ? p.w_kl[d] : (T)1;
                // x ln(x/wh) - x + wh is evaluated as fma(x, ln r, wh - x): the cancelling pair (wh - x) is formed first
                // (exact within a factor of two), so the term keeps ~1e-7 relative accuracy instead of 1e-7 * x; terms are
                // summed in double.  The fit's convergence test (tol 1e-7, signature_nmf.py:376) sees this number.
                double kl = 0.0;
                const int v_half = (((V + VEC - 1) / VEC + SPLIT - 1) / SPLIT) * VEC;
                const int v_begin = half * v_half, v_end = v_begin + v_half < V ? v_begin + v_half : V;
                for (int v0 = v_begin; v0 < v_end; v0 += VEC) {
                    const V16 x = *reinterpret_cast<const V16*>(xrow + v0);
                    V16 r;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        const int v = v0 + e;
                        T w[KP];
#pragma unroll
                        for (int j = 0; j < KP; j += VEC) {
                            const V16 t = *reinterpret_cast<const V16*>(sW + v * KP + j);
#pragma unroll
                            for (int q = 0; q < VEC; ++q) w[j + q] = t.v[q];
                        }
                        T wh0 = (T)0, wh1 = (T)0;
#pragma unroll
                        for (int j = 0; j < KP; j += 2) {
                            wh0 = sal_fma(w[j], h[j], wh0);
                            wh1 = sal_fma(w[j + 1], h[j + 1], wh1);
                        }
                        const T wh = wh0 + wh1;
                        const T xv = x.v[e];
                        const bool in = v < V;
                        const T rr = in ? sal_div(xv, wh) : (T)0;
#pragma unroll
                        for (int j = 0; j < KP; ++j) hn[j] = sal_fma(w[j], rr, hn[j]);
                        r.v[e] = rr * wk;
                        if (do_kl && in) kl += (double)(xv != (T)0 ? sal_fma(xv, sal_log(rr), wh - xv) : wh);
                        if (do_pois && in) kl += (double)(wh != (T)0 ? sal_fma(xv, sal_log(wh), -wh) : -wh);
                    }
                    *reinterpret_cast<V16*>(xrow + v0) = r;
                }
                if (SPLIT == 2) {  // the pair's halves (both lanes of a pair are always active together)
                    const unsigned int pm = __activemask();
#pragma unroll
                    for (int j = 0; j < KP; ++j) hn[j] += __shfl_xor_sync(pm, hn[j], 1);
                    kl += __shfl_xor_sync(pm, kl, 1);
                }
                if ((p.flags & SAL_PASS_SAMPLEWISE) && half == 0) p.per_sample[d] = (T)kl;
                if ((p.flags & (SAL_PASS_OBJECTIVE | SAL_PASS_POISSON)) && half == 0) {
                    double o = kl * (double)wk;
                    if (p.w_lhalf && !do_pois) {
                        T sq = (T)0;
#pragma unroll
                        for (int j = 0; j < KP; ++j) sq += sal_sqrt(h[j]);
                        o += (double)p.w_lhalf[d] * (double)sq;
                    }
                    obj_acc += o;
                }
                if (do_h && half == 0) {
                    if (p.h_scale && !(p.flags & SAL_PASS_SCALED_UPDATE)) {
#pragma unroll
                        for (int j = 0; j < KP; ++j) hn[j] = h[j];
                    } else if (p.flags & SAL_PASS_NOCLIP) {
#pragma unroll
                        for (int j = 0; j < KP; ++j) hn[j] = h[j] * hn[j];
                    } else if (!p.w_lhalf) {
#pragma unroll
                        for (int j = 0; j < KP; ++j) hn[j] = max(h[j] * hn[j], eps);
                    } else {
                        const T lam = p.w_lhalf[d];
                        const T wsq = p.w_kl ? wk * wk : (T)1;
#pragma unroll
                        for (int j = 0; j < KP; ++j) {
                            T t = (T)4 * h[j] * hn[j];
                            if (p.w_kl) t *= wsq;
                            const T disc = (T)0.25 * lam * lam + t;
                            const T root = lam / (T)2 - sal_sqrt(disc);
                            T o = (T)0.25 * root * root;
                            if (p.w_kl) o = o / wsq;
                            hn[j] = max(o, eps);
                        }
                    }
                    T* og = p.H_out + (size_t)d * k;
                    if (p.vec_h) {
#pragma unroll
                        for (int j = 0; j < KP; j += VEC)
                            if (j < k) {
                                V16 t;
#pragma unroll
                                for (int e = 0; e < VEC; ++e) t.v[e] = hn[j + e];
                                *reinterpret_cast<V16*>(og + j) = t;
                            }
                    } else {
#pragma unroll
                        for (int j = 0; j < KP; ++j)
                            if (j < k) og[j] = hn[j];
                    }
                }
            } else if (half == 0) {
                const int vr = (V + VEC - 1) / VEC * VEC;
                for (int v0 = 0; v0 < vr; ++v0) xrow[v0] = (T)0;
            }
        }
        __syncthreads();

        // ---- phase 2: numerator tile  acc[j] += A[s][v] * H[s][kh*KH + j]
        if ((do_w || do_hsum) && pv < V) {
            const T* hbase = sH + kh * KH;
#pragma unroll 4
            for (int s = 0; s < TS; ++s) {
                const T r = sX[s * XP + pv];
                T hv[KH];
#pragma unroll
                for (int j = 0; j < KH; j += VEC) {
                    const V16 t = *reinterpret_cast<const V16*>(hbase + s * HP + j);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) hv[j + e] = t.v[e];
                }
#pragma unroll
                for (int j = 0; j < KH; ++j) acc[j] = sal_fma(r, hv[j], acc[j]);
                if (do_hsum && pv == 0) {
#pragma unroll
                    for (int j = 0; j < KH; ++j) hs_acc[j] += hv[j];
                }
            }
        }
    }

    // ---- per-CTA partials (every launched CTA owns >= 1 tile, so all slots are written)
    if (do_w && pv < V) {
#pragma unroll
        for (int j = 0; j < KH; ++j)
            p.partial_wnum[((size_t)blockIdx.x * KP + kh * KH + j) * VP + pv] = acc[j];
    }
    if (do_hsum && pv == 0) {
#pragma unroll
        for (int j = 0; j < KH; ++j) p.partial_hsum[(size_t)blockIdx.x * SAL_KMAX + kh * KH + j] = (double)hs_acc[j];
    }
    if (p.flags & (SAL_PASS_OBJECTIVE | SAL_PASS_POISSON)) {
        const double t = block_sum_192(obj_acc, s_red);
        if (tid == 0) p.partial_obj[blockIdx.x] = t;
    }
}

// Fixed-order reduction of the per-CTA partials (deterministic), optionally followed by the W epilogue.
//   blocks 0 .. k-1 : signature j.  Thread (part, v) sums the partials b = part, part + 8, ... of Wnum[j][v] in double;
//                     the eight parts are added in a fixed order.  With fuse_epilogue the block then applies
//                     W_out[j] = clip(colnorm(W[j] * Wnum[j])) exactly as w_epilogue_kernel does.
//   block k         : objective and row sums of H.
constexpr int FIN_PARTS = 8, FIN_THREADS = FIN_PARTS * VP;
template <typename T>
__global__ void __launch_bounds__(FIN_THREADS) klnmf_finish_kernel(const T* partial_wnum, const double* partial_obj,
                                                                  const double* partial_hsum, int n_part, int KP, int k,
                                                                  int V, int flags, T* Wnum, double* objective, T* hsum,
                                                                  int fuse_epilogue, const T* W_in, T* W_out, int n_given,
                                                                  int clip_given) {
    __shared__ double s_part[FIN_PARTS][VP];
    __shared__ double s_red[VP / 32];
    const int j = blockIdx.x, tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();  // the partials come from the pass kernel this launch may have overtaken
    if (j < k) {
        if (!(flags & SAL_PASS_WNUM)) return;
        const int part = tid / VP, v = tid - part * VP;
        double s = 0.0;
        if (v < V) {
            const T* src = partial_wnum + (size_t)j * VP + v;
            // (issuing all ~19 loads of a thread before the first add was measured SLOWER than this 8-way unrolled loop:
            // 24.5 vs 22.1 us per update at 125k samples)
#pragma unroll 8
            for (int b = part; b < n_part; b += FIN_PARTS) s += (double)src[(size_t)b * KP * VP];
        }
        s_part[part][v] = s;
        __syncthreads();
        if (part != 0) return;
        const bool in = v < V;
        const double num = ((s_part[0][v] + s_part[1][v]) + (s_part[2][v] + s_part[3][v])) +
                           ((s_part[4][v] + s_part[5][v]) + (s_part[6][v] + s_part[7][v]));
        if (in) Wnum[(size_t)j * V + v] = (T)num;
        if (!fuse_epilogue) return;
        // same arithmetic as w_epilogue_kernel on the value just stored (rounded to T like the unfused path)
        const double w = in ? (double)W_in[(size_t)j * V + v] : 0.0;
        const double val = in ? w * (double)(T)num : 0.0;
        double t = val;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((v & 31) == 0) s_red[v >> 5] = t;
        asm volatile("bar.sync 1, %0;" ::"r"(VP));  // the 96 threads of part 0
        t = s_red[0] + s_red[1] + s_red[2];
        if (!in) return;
        double out = val / t;
        if (j < n_given) out = w;
        if (clip_given || j >= n_given) out = fmax(out, (double)SAL_EPS_F32);
        W_out[(size_t)j * V + v] = (T)out;
        return;
    }
    // tail block: objective and row sums of H.  One load per thread where possible, fixed-order combination.
    if (flags & (SAL_PASS_OBJECTIVE | SAL_PASS_POISSON)) {
        double s = 0.0;
        for (int b = tid; b < n_part; b += FIN_THREADS) s += __ldcg(partial_obj + b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        double* s_w = &s_part[0][0];  // FIN_THREADS / 32 warp totals
        if ((tid & 31) == 0) s_w[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < FIN_THREADS / 32; ++w) t += s_w[w];
            *objective = t;
        }
        __syncthreads();
    }
    if (flags & SAL_PASS_HSUM) {
        // thread (part, jj): partials b = part, part + HS_PARTS, ... of signature jj
        constexpr int HS_PARTS = FIN_THREADS / SAL_KMAX;
        const int part = tid / SAL_KMAX, jj = tid - part * SAL_KMAX;
        double s = 0.0;
        if (jj < k)
            for (int b = part; b < n_part; b += HS_PARTS) s += __ldcg(partial_hsum + (size_t)b * SAL_KMAX + jj);
        double* s_h = &s_part[0][0];  // [HS_PARTS][SAL_KMAX]
        s_h[part * SAL_KMAX + jj] = s;
        __syncthreads();
        if (tid < k) {
            double t = 0.0;
            for (int q = 0; q < HS_PARTS; ++q) t += s_h[q * SAL_KMAX + tid];
            hsum[tid] = (T)t;
        }
    }
}

// ---- multi-GPU: reduction of the partials + one-shot all-reduce over NVLink peer memory + W epilogue, ONE kernel ----
// Every rank owns a receive buffer that all peers have mapped (symmetric memory):
//     uint2 recv[2][n_ranks][k + 1][VP][2];          slot = sequence number & 1; one (payload32, seq) word pair per double
// Thread v of block j reduces this rank's partials of Wnum[j][v] and PUSHES the double, as two 8-byte words each tagged
// with the sequence number (the "LL" idea: an 8-byte store is atomic, so data and flag arrive together and no fence is
// needed), into recv[slot][my rank][j][v] of every peer -- remote stores, one NVLink hop.  It then polls its OWN buffer
// until the words of every other rank carry the tag and sums the N contributions IN RANK ORDER, so every rank obtains
// bit-identical sums, and applies the W epilogue.  Block k does the same for the objective scalar.  Two slots suffice:
// a rank can be at most one update ahead of the slowest peer, because its next update needs that peer's words.
// seq lives in device memory and is advanced by the last block to finish, so the kernel can sit in a CUDA graph.
// The payload is 96 x k doubles: latency, not bandwidth.
struct P2PExchange {
    void* const* peers;   // device array [n_ranks] of the ranks' receive buffers (this process's mappings)
    unsigned int* seq;    // device scalar, starts at 1
    unsigned int* ticket; // device scalar, 0
    int n_ranks, rank;
};

// both tagged words of a double in one 16-byte access {lo, tag, hi, tag}: half the NVLink transactions and a warp writes
// 512 contiguous bytes; each 8-byte half still validates itself, so a torn 16-byte store cannot be mistaken for a whole one
__device__ __forceinline__ void st_ll4(uint4* p, unsigned int lo, unsigned int hi, unsigned int tag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ld_ll4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

template <typename T>
__global__ void __launch_bounds__(FIN_THREADS) klnmf_finish_p2p_kernel(const T* partial_wnum, const double* partial_obj, int n_part,
                                                                      int KP, int k, int V, int flags, T* Wnum, double* objective,
                                                                      const T* W_in, T* W_out, int n_given, int clip_given,
                                                                      P2PExchange x) {
    __shared__ double s_part[FIN_PARTS][VP];
    __shared__ double s_red[VP / 32];
    __shared__ double s_obj[FIN_THREADS / 32];
    const int j = blockIdx.x, tid = threadIdx.x;
    const int part = tid / VP, v = tid - part * VP;
    pdl_trigger();
    pdl_wait();  // partials, sequence number and W all come from kernels this launch may have overtaken
    const unsigned int seq = *(volatile unsigned int*)x.seq;
    const int slot = seq & 1;
    const bool is_obj = j == k;
    const bool active = is_obj ? (flags & SAL_PASS_OBJECTIVE) != 0 : (flags & SAL_PASS_WNUM) != 0;
    if (active) {
        // (1) this rank's contribution
        double s = 0.0;
        if (!is_obj) {
            if (v < V) {
                const T* src = partial_wnum + (size_t)j * VP + v;
#pragma unroll 8
                for (int b = part; b < n_part; b += FIN_PARTS) s += (double)src[(size_t)b * KP * VP];
            }
            s_part[part][v] = s;
            __syncthreads();
            s = ((s_part[0][v] + s_part[1][v]) + (s_part[2][v] + s_part[3][v])) + ((s_part[4][v] + s_part[5][v]) + (s_part[6][v] + s_part[7][v]));
        } else {
            for (int b = tid; b < n_part; b += FIN_THREADS) s += partial_obj[b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((tid & 31) == 0) s_obj[tid >> 5] = s;
            __syncthreads();
            s = 0.0;
            for (int w = 0; w < FIN_THREADS / 32; ++w) s += s_obj[w];
        }
        // (2) + (3): the exchange is spread over the eight thread groups ("parts") of the block -- group p pushes this
        // rank's value to peers p, p + 8, ... and collects the words of ranks p, p + 8, ..., one remote store and one
        // polled load per thread instead of a serial loop over the peers; group 0 then adds the contributions in rank order.
        const int nv = is_obj ? 1 : V;
        const size_t row = ((size_t)slot * x.n_ranks) * (k + 1) * VP;  // start of recv[slot]
        const unsigned long long bits = (unsigned long long)__double_as_longlong(s);
        if (v < nv)
            for (int r = part; r < x.n_ranks; r += FIN_PARTS)
                if (r != x.rank) {
                    uint4* dst = reinterpret_cast<uint4*>(x.peers[r]) + (row + ((size_t)x.rank * (k + 1) + j) * VP + v);
                    st_ll4(dst, (unsigned int)bits, (unsigned int)(bits >> 32), seq);
                }
        double total = 0.0;
        const uint4* mine = reinterpret_cast<const uint4*>(x.peers[x.rank]);
        __syncthreads();  // every thread has formed s from s_part / s_obj; s_part is reused for the received values
        for (int r0 = 0; r0 < x.n_ranks; r0 += FIN_PARTS) {
            const int r = r0 + part;
            double val = 0.0;
            if (v < nv && r < x.n_ranks) {
                if (r == x.rank) {
                    val = s;
                } else {
                    const uint4* src = mine + (row + ((size_t)r * (k + 1) + j) * VP + v);
                    uint4 w;
                    long long spins = 0;
                    do {
                        w = ld_ll4(src);
                        if (++spins > 2000000000LL) __trap();  // a peer died: fail loudly instead of hanging the box
                    } while (w.y != seq || w.w != seq);
                    val = __longlong_as_double((long long)(((unsigned long long)w.z << 32) | w.x));
                }
            }
            s_part[part][v] = val;
            __syncthreads();
            if (part == 0 && v < nv)
                for (int u = 0; u < FIN_PARTS && r0 + u < x.n_ranks; ++u) total += s_part[u][v];
            __syncthreads();
        }
        if (part == 0) {  // 96 threads carry on (named barrier 1 for the epilogue's block sum)
            if (is_obj) {
                if (v == 0) *objective = total;
            } else {
                const bool in = v < V;
                if (in) Wnum[(size_t)j * V + v] = (T)total;
                const double w = in ? (double)W_in[(size_t)j * V + v] : 0.0;
                const double val = in ? w * (double)(T)total : 0.0;
                double t = val;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                if ((v & 31) == 0) s_red[v >> 5] = t;
                asm volatile("bar.sync 1, %0;" ::"r"(VP));
                t = s_red[0] + s_red[1] + s_red[2];
                if (in) {
                    double out = val / t;
                    if (j < n_given) out = w;
                    if (clip_given || j >= n_given) out = fmax(out, (double)SAL_EPS_F32);
                    W_out[(size_t)j * V + v] = (T)out;
                }
            }
        }
    }
    // (4) the last block to finish advances the sequence number for the next launch
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(x.ticket, 1u) == gridDim.x - 1) {
            *x.ticket = 0;
            *(volatile unsigned int*)x.seq = seq + 1;
            __threadfence();
        }
    }
}

// W epilogue: one CTA per signature.
template <typename T>
__global__ void __launch_bounds__(128) w_epilogue_kernel(const T* W_in, const T* Wnum, int V, int n_given,
                                                         int clip_given, T* W_out) {
    __shared__ double s_red[4];
    const int j = blockIdx.x, v = threadIdx.x;
    const bool in = v < V;
    const double w = in ? (double)W_in[(size_t)j * V + v] : 0.0;
    double val = in ? w * (double)Wnum[(size_t)j * V + v] : 0.0;
    double s = val;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((v & 31) == 0) s_red[v >> 5] = s;
    __syncthreads();
    s = s_red[0] + s_red[1] + s_red[2] + s_red[3];
    if (!in) return;
    double out = val / s;
    if (j < n_given) out = w;
    if (clip_given || j >= n_given) out = fmax(out, (double)SAL_EPS_F32);
    W_out[(size_t)j * V + v] = (T)out;
}

// X <- max(X, eps) in place, counting the entries that changed (grid-stride, one atomic per CTA).
template <typename T>
__global__ void __launch_bounds__(256) clip_counts_kernel(T* X, int64_t n, unsigned long long* n_changed) {
    __shared__ unsigned int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const T eps = (T)SAL_EPS_F32;
    unsigned int cnt = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T x = X[i];
        if (x < eps) X[i] = eps, ++cnt;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0 && s_cnt) atomicAdd(n_changed, (unsigned long long)s_cnt);
}

template <typename T>
__global__ void __launch_bounds__(256) scale_clip_rows_kernel(T* H, const T* scale, int64_t n, int k) {
    const T eps = (T)SAL_EPS_F32;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        H[i] = max(H[i] * scale[i % k], eps);
}

template <typename T, int KP>
size_t pass_smem_bytes() {
    using C = Cfg<T>;
    return sizeof(T) * ((size_t)VP * KP + (size_t)C::TS * (KP + C::HPAD) + (size_t)C::TS * C::XP);
}

template <typename T, int KP>
int launch_pass_t(sal_ctx* c, const PassArgs& a, cudaStream_t st) {
    using C = Cfg<T>;
    static bool attr_set[16] = {false};
    const size_t smem = pass_smem_bytes<T, KP>();
    if (!attr_set[c->device & 15]) {
        SAL_CUDA(cudaFuncSetAttribute(klnmf_pass_kernel<T, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[c->device & 15] = true;
    }
    PassParams<T> p;
    p.X = (const T*)a.X, p.W = (const T*)a.W, p.H_in = (const T*)a.H_in;
    p.w_kl = (const T*)a.w_kl, p.w_lhalf = (const T*)a.w_lhalf, p.h_scale = (const T*)a.h_scale;
    p.H_out = (T*)a.H_out, p.partial_wnum = (T*)c->partial_wnum, p.per_sample = (T*)a.per_sample;
    p.partial_obj = c->partial_obj, p.partial_hsum = c->partial_hsum;
    p.D = c->D, p.V = c->V, p.k = c->k, p.flags = a.flags;
    p.vec_x = (c->V % C::VEC == 0) && (((uintptr_t)a.X) % 16 == 0);
    p.vec_h = (c->k % C::VEC == 0) && (((uintptr_t)a.H_out) % 16 == 0);
    const int64_t n_tiles = (c->D + C::TS - 1) / C::TS;
    const int64_t resident = (int64_t)c->n_sm * pass_occ<T, KP>();
    const int grid = (int)(n_tiles < resident ? n_tiles : resident);
    if (int e = sal_timing_begin(c, a.flags, st)) return e;
    klnmf_pass_kernel<T, KP><<<grid, NT, smem, st>>>(p);
    SAL_CUDA(cudaGetLastError());
    if (int e = sal_timing_end(c, a.flags, st)) return e;
    c->launches++;
    if (int e = sal_launch_pass_reduce(c, a, grid, st)) return e;
    return 0;
}

template <typename T>
int launch_pass_k(sal_ctx* c, const PassArgs& a, cudaStream_t st) {
    switch (c->KP) {
        case 8: return launch_pass_t<T, 8>(c, a, st);
        case 16: return launch_pass_t<T, 16>(c, a, st);
        case 24: return launch_pass_t<T, 24>(c, a, st);
        default: return launch_pass_t<T, 32>(c, a, st);
    }
}

}  // namespace

// Fixed-order sum of the per-CTA partials left by either flavour of the pass (n_part = its grid size).
int sal_launch_pass_reduce(sal_ctx* c, const PassArgs& a, int n_part, cudaStream_t st) {
    if (!(a.flags & (SAL_PASS_WNUM | SAL_PASS_OBJECTIVE | SAL_PASS_POISSON | SAL_PASS_HSUM)) || a.partials_only) return 0;
    if (a.p2p_peers) {
        P2PExchange x;
        x.peers = (void* const*)a.p2p_peers, x.seq = (unsigned int*)a.p2p_state, x.ticket = (unsigned int*)a.p2p_state + 1;
        x.n_ranks = a.p2p_n_ranks, x.rank = a.p2p_rank;
        const int grid = c->k + 1;
        if (c->dtype == SAL_F32)
            SAL_CUDA(sal_launch_pdl(klnmf_finish_p2p_kernel<float>, grid, FIN_THREADS, 0, st, (const float*)c->partial_wnum, c->partial_obj,
                                    n_part, c->KP, c->k, c->V, a.flags, (float*)a.Wnum, a.objective, (const float*)a.W, (float*)a.W_out,
                                    a.n_given, a.clip_given, x));
        else
            SAL_CUDA(sal_launch_pdl(klnmf_finish_p2p_kernel<double>, grid, FIN_THREADS, 0, st, (const double*)c->partial_wnum, c->partial_obj,
                                    n_part, c->KP, c->k, c->V, a.flags, (double*)a.Wnum, a.objective, (const double*)a.W, (double*)a.W_out,
                                    a.n_given, a.clip_given, x));
        SAL_CUDA(cudaGetLastError());
        c->launches++;
        return 0;
    }
    const bool need_tail = a.flags & (SAL_PASS_OBJECTIVE | SAL_PASS_POISSON | SAL_PASS_HSUM);
    const int grid = c->k + (need_tail ? 1 : 0);
    const int fuse = a.fuse_epilogue && (a.flags & SAL_PASS_WNUM);
    if (c->dtype == SAL_F32)
        SAL_CUDA(sal_launch_pdl(klnmf_finish_kernel<float>, grid, FIN_THREADS, 0, st, (const float*)c->partial_wnum, c->partial_obj,
                                c->partial_hsum, n_part, c->KP, c->k, c->V, a.flags, (float*)a.Wnum, a.objective, (float*)a.hsum, fuse,
                                (const float*)a.W, (float*)a.W_out, a.n_given, a.clip_given));
    else
        SAL_CUDA(sal_launch_pdl(klnmf_finish_kernel<double>, grid, FIN_THREADS, 0, st, (const double*)c->partial_wnum, c->partial_obj,
                                c->partial_hsum, n_part, c->KP, c->k, c->V, a.flags, (double*)a.Wnum, a.objective, (double*)a.hsum, fuse,
                                (const double*)a.W, (double*)a.W_out, a.n_given, a.clip_given));
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_pass_fma(sal_ctx* c, const PassArgs& a, cudaStream_t st) {
    return c->dtype == SAL_F32 ? launch_pass_k<float>(c, a, st) : launch_pass_k<double>(c, a, st);
}

int sal_pass_smem_bytes(int dtype, int KP) {
    if (dtype == SAL_F32) return (int)pass_smem_bytes<float, 32>() - (32 - KP) * 4 * (VP + Cfg<float>::TS);
    return (int)pass_smem_bytes<double, 32>() - (32 - KP) * 8 * (VP + Cfg<double>::TS);
}

int sal_launch_w_epilogue(sal_ctx* c, const void* W_in, const void* Wnum, int n_given, int clip_given,
                          void* W_out, cudaStream_t st) {
    const size_t bytes = (size_t)c->k * c->V * (c->dtype == SAL_F32 ? 4 : 8);
    if (n_given >= c->k) {
        if (W_out != W_in) SAL_CUDA(cudaMemcpyAsync(W_out, W_in, bytes, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    if (c->dtype == SAL_F32)
        w_epilogue_kernel<float><<<c->k, 128, 0, st>>>((const float*)W_in, (const float*)Wnum, c->V, n_given,
                                                      clip_given, (float*)W_out);
    else
        w_epilogue_kernel<double><<<c->k, 128, 0, st>>>((const double*)W_in, (const double*)Wnum, c->V, n_given,
                                                       clip_given, (double*)W_out);
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_clip_counts(sal_ctx* c, void* X, int64_t n, long long* n_changed, cudaStream_t st) {
    const int64_t want = (n + 255) / 256;
    const int grid = (int)(want < (int64_t)c->n_sm * 8 ? want : (int64_t)c->n_sm * 8);
    if (c->dtype == SAL_F32)
        clip_counts_kernel<float><<<grid, 256, 0, st>>>((float*)X, n, (unsigned long long*)n_changed);
    else
        clip_counts_kernel<double><<<grid, 256, 0, st>>>((double*)X, n, (unsigned long long*)n_changed);
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}

int sal_launch_scale_clip_rows(sal_ctx* c, void* H, const void* scale, cudaStream_t st) {
    const int64_t n = c->D * c->k, want = (n + 255) / 256;
    const int grid = (int)(want < (int64_t)c->n_sm * 8 ? want : (int64_t)c->n_sm * 8);
    if (c->dtype == SAL_F32)
        scale_clip_rows_kernel<float><<<grid, 256, 0, st>>>((float*)H, (const float*)scale, n, c->k);
    else
        scale_clip_rows_kernel<double><<<grid, 256, 0, st>>>((double*)H, (const double*)scale, n, c->k);
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return 0;
}
