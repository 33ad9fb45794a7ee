// Fused KL-NMF pass, tensor-core flavour (tcgen05 kind::tf32, fp32 accumulation in TMEM).
//
// Same contract as the FMA pass (klnmf_pass.cu; reference update_WH / update_H / update_W /
// kl_divergence, models/_utils_klnmf.py:11-55, 164-361) for fp32 handles, V = 96, k % 4 == 0, no
// per-sample weights.  The three thin contractions of one 128-sample tile run on the 5th-gen tensor
// cores; X is streamed from HBM exactly once by TMA and the quotient never leaves the SM:
//
//   G1  WH[s,f]   = sum_j H[s,j] W[j,f]        A = sH  (smem, K-major)      B = sW1 (smem, K-major)
//   E1  R[s,f]    = X[s,f] / WH[s,f]           thread s = TMEM lane s; R overwrites X in smem AND WH in TMEM
//   G2  Hn[s,j]   = sum_f R[s,f] W[j,f]        A = R   (TMEM, "TS" form)    B = sW2 (smem, K-major)
//   G3  Wn[f,j]  += sum_s R[s,f] H[s,j]        A = R   (smem, MN-major)     B = sHT (smem, K-major)
//   E2  H_out[s,j] = max(H[s,j] Hn[s,j], eps)
//
// Wn (96 x k) accumulates in TMEM over all tiles of the CTA and is written once as a per-CTA partial;
// the deterministic fixed-order reduction kernel of klnmf_pass.cu finishes the job.
//
// Shared-memory operand layouts
//   X / R stage : 3 TMA boxes [128 samples][32 features] fp32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  The very same
//                 bytes are the MN-major SWIZZLE_128B_BASE32B A operand of G3 (M = feature, K = sample) -- the only
//                 MN-major layout tcgen05 accepts for 32-bit operands -- so R is written in place over X.
//   sH, sW1, sW2, sHT : canonical no-swizzle K-major "core matrices" (8 rows x 16 B, 128 B contiguous),
//                 written by the threads with round-to-nearest tf32 conversion.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over tiles):  warp 0 TMA producer, warp 1 MMA issuer,
// warp 2 TMEM allocator, warps 4-7 and 8-11 two epilogue warpgroups that alternate tiles.
#include <cuda.h>  // CUtensorMap types; cuTensorMapEncodeTiled itself is fetched through the runtime (no -lcuda)

#include "sal_common.cuh"

namespace {

constexpr int TILE = 128;                       // samples per tile = UMMA M of G1 / G2
constexpr int VT = 96;                          // features
constexpr int NBOX = 3;                         // 32-feature TMA boxes per tile
constexpr int BOX_BYTES = TILE * 128;           // 16 KB
constexpr int XSTAGE_BYTES = NBOX * BOX_BYTES;  // 48 KB
constexpr int NTHREADS = 384;
constexpr int N2 = 32;                          // UMMA N of G2 / G3 (k zero-padded to 32)

// TMEM columns (512 allocated): two WH/R buffers, two Hn buffers, the persistent numerator
constexpr uint32_t TM_WH0 = 0, TM_WH1 = 96, TM_HN0 = 192, TM_HN1 = 224, TM_WN = 256, TM_COLS = 512;

// strides of the thread-written operands (bytes)
constexpr int SH_LBO = 2048, SH_SBO = 128;    // sH  [kc][sample/8][8][16B]   A of G1 (M = sample, K = signature)
constexpr int SW1_LBO = 1536, SW1_SBO = 128;  // sW1 [kc][feature/8][8][16B]  B of G1 (N = feature, K = signature)
constexpr int SW2_SBO = 128;                  // sW2 [fc][sig/8][8][16B]      B of G2 (N = signature, K = feature); LBO = Lay::SW2_LBO
constexpr int SHT_LBO = 528, SHT_SBO = 128;   // sHT [sc][sig/8][8][16B](+16) B of G3 (N = signature, K = sample)

template <int KP8>
struct Lay {
    static constexpr int S = KP8 <= 24 ? 3 : 2;  // X / H stages
    static constexpr int HRAW = TILE * KP8 * 4;  // raw H tile as TMA delivers it ([128][k] dense rows)
    // only ceil(k/8) signature groups of sW2 are stored; G2 runs with N = 32 and the groups beyond them read the
    // following bytes (finite garbage that only reaches output columns >= k, which nobody reads)
    static constexpr int SW2_LBO = (KP8 / 8) * 128;
    static constexpr int OFF_X = 0;
    static constexpr int OFF_HRAW = OFF_X + S * XSTAGE_BYTES;
    static constexpr int OFF_SW1 = OFF_HRAW + S * HRAW;
    static constexpr int OFF_SW2 = OFF_SW1 + (KP8 / 4) * SW1_LBO;
    static constexpr int OFF_SH = OFF_SW2 + 24 * SW2_LBO;
    static constexpr int OFF_SHT = OFF_SH + (KP8 / 4) * SH_LBO;
    static constexpr int OFF_BAR = OFF_SHT + 32 * SHT_LBO;
    static constexpr int N_BAR = 2 * S + 9;
    static constexpr int OFF_MISC = OFF_BAR + N_BAR * 8;
    static constexpr int TOTAL = OFF_MISC + 128;
    static constexpr int DYN_BYTES = TOTAL;
    // the G3 A operand spans 4 boxes (M = 128 features, 96 real): the bytes after the last stage must exist
    static_assert(S * HRAW + (KP8 / 4) * SW1_LBO + 24 * SW2_LBO + (KP8 / 4) * SH_LBO >= BOX_BYTES, "need 16 KB after the last X stage");
    static_assert(DYN_BYTES <= 232448, "shared memory budget");
};

struct TcParams {
    const float* W;
    float* H_out;
    float* partial_wnum;
    double* partial_obj;
    float* dbg;  // optional diagnostics buffer (see sal_set_debug_buffer)
    int64_t D;
    int k, flags, n_tiles;
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
        "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
        "r"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}

#define SAL_R8(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7])
#define SAL_W8(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7])

// thread i of warp q  <->  TMEM lane 32 q + i;  v[c] <-> column (addr & 0xffff) + c
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : SAL_R8(v, 0), SAL_R8(v, 8), SAL_R8(v, 16), SAL_R8(v, 24)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};" ::SAL_W8(v, 0),
        SAL_W8(v, 8), SAL_W8(v, 16), SAL_W8(v, 24), "r"(taddr)
        : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
// fp32 -> tf32, round to nearest (ties away from zero, like cvt.rna.tf32.f32) for finite values: two integer ops
// instead of the multi-instruction sequence cvt.rna expands to.
__device__ __forceinline__ uint32_t tf32_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ float tf32_rn(float x) { return __uint_as_float(tf32_bits(x)); }
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// One 32-feature box of the quotient tile for sample row s:  v[] holds WH[s][32 c .. 32 c + 31] (TMEM columns) on
// entry and tf32(R) on exit; X is read from / R written to the swizzled stage (32-byte chunk m of row s sits at
// chunk m ^ (s & 3); the two 16-byte halves are visited in lane-dependent order so that rows s and s + 4 never hit
// the same banks in one wavefront).
template <bool DO_R, bool DO_KL>
__device__ __forceinline__ void quotient_box(uint32_t (&v)[32], uint32_t rowbase, int s, uint32_t sw, float& kl) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const uint32_t a32 = rowbase + ((uint32_t)(m ^ (s & 3)) << 5);
        const float4 va = lds128(a32 + sw * 16), vb = lds128(a32 + (sw ^ 1) * 16);
        const float4 xlo = sw ? vb : va, xhi = sw ? va : vb;
        const float xv[8] = {xlo.x, xlo.y, xlo.z, xlo.w, xhi.x, xhi.y, xhi.z, xhi.w};
        float rr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float wh = __uint_as_float(v[8 * m + e]);
            const float r = xv[e] * rcp_approx(wh);
            if (DO_KL) {
                if (xv[e] != 0.f) kl += xv[e] * logf(r) - xv[e];
                kl += wh;
            }
            rr[e] = tf32_rn(r);
            v[8 * m + e] = __float_as_uint(rr[e]);
        }
        if (DO_R) {
            const float4 rlo = make_float4(rr[0], rr[1], rr[2], rr[3]), rhi = make_float4(rr[4], rr[5], rr[6], rr[7]);
            sts128(a32 + sw * 16, sw ? rhi : rlo);
            sts128(a32 + (sw ^ 1) * 16, sw ? rlo : rhi);
        }
    }
}

// shared-memory matrix descriptor (tcgen05): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | layout << 61
constexpr uint64_t LAYOUT_NONE = 0, LAYOUT_128B_BASE32B = 1;
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}
// instruction descriptor: fp32 accumulate, tf32 x tf32, majors (0 = K, 1 = MN), N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}

template <int KP8, bool DO_R, bool DO_KL>
__global__ void __launch_bounds__(NTHREADS, 1)
klnmf_pass_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapH, TcParams p) {
    using L = Lay<KP8>;
    constexpr int S = L::S;
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    const uint32_t base = smem_u32(smem_dyn);
    unsigned char* gbase = smem_dyn;
    if (base & 1023u) __trap();  // the swizzled TMA boxes need 1024-byte alignment
    const uint32_t sX = base + L::OFF_X, sHraw = base + L::OFF_HRAW, sW1 = base + L::OFF_SW1, sW2 = base + L::OFF_SW2;
    const uint32_t sH = base + L::OFF_SH, sHT = base + L::OFF_SHT, bars = base + L::OFF_BAR;
    // barriers
    const uint32_t bar_full = bars, bar_empty = bars + 8 * S, bar_hready = bars + 16 * S;
    const uint32_t bar_whfull = bar_hready + 8, bar_rready = bar_whfull + 16, bar_hnfull = bar_rready + 16;
    const uint32_t bar_shtfree = bar_hnfull + 16, bar_done = bar_shtfree + 8;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(gbase + L::OFF_MISC);
    double* s_red = reinterpret_cast<double*>(gbase + L::OFF_MISC + 16);  // [8]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = p.k;
    const bool do_h = p.flags & SAL_PASS_UPDATE_H, do_w = p.flags & SAL_PASS_WNUM;
    constexpr bool do_r = DO_R;  // quotient needed by G2 / G3 (UPDATE_H or WNUM requested)
    constexpr bool do_kl = DO_KL;
    const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA (>= 1)

    // ---- one-time setup --------------------------------------------------------------------------
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapH) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) mbar_init(bar_full + 8 * i, 1), mbar_init(bar_empty + 8 * i, 1);
        mbar_init(bar_hready, 128);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_whfull + 8 * i, 1);
            mbar_init(bar_rready + 8 * i, 128);
            mbar_init(bar_hnfull + 8 * i, 1);
        }
        mbar_init(bar_shtfree, 1);
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(TM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // W operands (tf32, round to nearest), zero padding for signatures >= k; sHT zero fill
    for (int i = tid; i < KP8 * VT; i += NTHREADS) {
        const int j = i / VT, f = i - j * VT;
        const float w = j < k ? tf32_rn(p.W[(size_t)j * VT + f]) : 0.f;
        sts32(sW1 + (j >> 2) * SW1_LBO + (f >> 3) * SW1_SBO + (f & 7) * 16 + (j & 3) * 4, w);
    }
    // Objective-only passes feed the convergence test (reference signature_nmf.py:373-378, tol 1e-7), so their WH is
    // formed with the error-compensated split  w = hi + lo, h = hi + lo,  WH ~ hi*hi + lo*hi + hi*lo  (3 x tf32, ~2^-22):
    // the lo parts live where sW2 / sHT would be (unused without G2 / G3).
    constexpr bool split = !DO_R;
    const uint32_t sW1lo = sW2, sHlo = sHT;
    for (int i = tid; i < KP8 * VT; i += NTHREADS) {
        const int j = i / VT, f = i - j * VT;
        const float wf = j < k ? p.W[(size_t)j * VT + f] : 0.f;
        if (split)
            sts32(sW1lo + (j >> 2) * SW1_LBO + (f >> 3) * SW1_SBO + (f & 7) * 16 + (j & 3) * 4, tf32_rn(wf - tf32_rn(wf)));
        else
            sts32(sW2 + (f >> 2) * L::SW2_LBO + (j >> 3) * SW2_SBO + (j & 7) * 16 + (f & 3) * 4, tf32_rn(wf));
    }
    if (!split)
        for (int i = tid; i < 32 * SHT_LBO / 4; i += NTHREADS) sts32(sHT + 4 * i, 0.f);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    double obj_acc = 0.0;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            const uint32_t tx = XSTAGE_BYTES + (uint32_t)(TILE * k * 4);
            for (int i = 0; i < n_my; ++i) {
                const int st = i % S;
                const int d0 = ((int)blockIdx.x + i * (int)gridDim.x) * TILE;
                mbar_wait(bar_empty + 8 * st, ((i / S) & 1) ^ 1);
                mbar_arrive_expect_tx(bar_full + 8 * st, tx);
                for (int c = 0; c < NBOX; ++c) tma_load_2d(sX + st * XSTAGE_BYTES + c * BOX_BYTES, &mapX, bar_full + 8 * st, c * 32, d0);
                tma_load_2d(sHraw + st * L::HRAW, &mapH, bar_full + 8 * st, 0, d0);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t ID1 = make_idesc(128, VT, 0, 0);
            constexpr uint32_t ID2 = make_idesc(128, N2, 0, 0);
            constexpr uint32_t ID3 = make_idesc(128, N2, 1, 0);
            auto issue_g1 = [&](int i) {
                mbar_wait(bar_hready, i & 1);
                tc_fence_after();
                const uint32_t d = tmem + ((i & 1) ? TM_WH1 : TM_WH0);
#pragma unroll
                for (int ks = 0; ks < KP8 / 8; ++ks)
                    mma_ss(d, make_desc(sH + ks * 2 * SH_LBO, SH_LBO, SH_SBO, LAYOUT_NONE),
                           make_desc(sW1 + ks * 2 * SW1_LBO, SW1_LBO, SW1_SBO, LAYOUT_NONE), ID1, ks > 0);
                if (split) {
#pragma unroll
                    for (int ks = 0; ks < KP8 / 8; ++ks) {
                        mma_ss(d, make_desc(sHlo + ks * 2 * SH_LBO, SH_LBO, SH_SBO, LAYOUT_NONE),
                               make_desc(sW1 + ks * 2 * SW1_LBO, SW1_LBO, SW1_SBO, LAYOUT_NONE), ID1, 1);
                        mma_ss(d, make_desc(sH + ks * 2 * SH_LBO, SH_LBO, SH_SBO, LAYOUT_NONE),
                               make_desc(sW1lo + ks * 2 * SW1_LBO, SW1_LBO, SW1_SBO, LAYOUT_NONE), ID1, 1);
                    }
                }
                tc_commit(bar_whfull + 8 * (i & 1));
            };
            issue_g1(0);
            for (int i = 0; i < n_my; ++i) {
                if (i + 1 < n_my) issue_g1(i + 1);
                const int st = i % S, b = i & 1;
                mbar_wait(bar_rready + 8 * b, (i >> 1) & 1);
                tc_fence_after();
                if (do_r) {
                    const uint32_t tR = tmem + (b ? TM_WH1 : TM_WH0), tHn = tmem + (b ? TM_HN1 : TM_HN0);
#pragma unroll
                    for (int ks = 0; ks < VT / 8; ++ks)
                        mma_ts(tHn, tR + ks * 8, make_desc(sW2 + ks * 2 * L::SW2_LBO, L::SW2_LBO, SW2_SBO, LAYOUT_NONE), ID2, ks > 0);
                    tc_commit(bar_hnfull + 8 * b);
                    const uint32_t xs = sX + st * XSTAGE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < TILE / 8; ++ks)
                        mma_ss(tmem + TM_WN, make_desc(xs + ks * 1024, BOX_BYTES, 512, LAYOUT_128B_BASE32B),
                               make_desc(sHT + ks * 2 * SHT_LBO, SHT_LBO, SHT_SBO, LAYOUT_NONE), ID3, (i > 0 || ks > 0));
                    tc_commit(bar_empty + 8 * st);
                    tc_commit(bar_shtfree);
                } else {  // objective only: nothing reads the stage after E1
                    mbar_arrive(bar_empty + 8 * st);
                    mbar_arrive(bar_shtfree);
                }
            }
            tc_commit(bar_done);
            mbar_wait(bar_done, 0);  // every MMA has retired before the CTA tears TMEM down
        }
    } else if (warp >= 4) {
        // ================= epilogue warpgroups =================
        const int g = (warp - 4) >> 2, q = warp & 3;
        const int s = q * 32 + lane;  // sample row of the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const uint32_t sw = (s >> 2) & 1;
        const float eps = (float)SAL_EPS_F32;
        for (int i = g; i < n_my; i += 2) {
            const int st = i % S, b = i & 1;
            const int64_t d0 = (int64_t)((int)blockIdx.x + i * (int)gridDim.x) * TILE;
            const bool valid = d0 + s < p.D;
            mbar_wait(bar_full + 8 * st, (i / S) & 1);
            // exposures of this sample (rows beyond D are zero-filled by TMA)
            float h[KP8];
            {
                const uint32_t hrow = sHraw + st * L::HRAW + s * (k * 4);
#pragma unroll
                for (int j = 0; j < KP8; j += 4) {
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < k) t = lds128(hrow + j * 4);
                    h[j] = t.x, h[j + 1] = t.y, h[j + 2] = t.z, h[j + 3] = t.w;
                }
            }
            if (i > 0) mbar_wait(bar_whfull + 8 * ((i - 1) & 1), ((i - 1) >> 1) & 1);  // G1(i-1) has read sH
#pragma unroll
            for (int kc = 0; kc < KP8 / 4; ++kc)
                sts128(sH + kc * SH_LBO + (s >> 3) * SH_SBO + (s & 7) * 16,
                       make_float4(tf32_rn(h[4 * kc]), tf32_rn(h[4 * kc + 1]), tf32_rn(h[4 * kc + 2]), tf32_rn(h[4 * kc + 3])));
            if (split) {
#pragma unroll
                for (int kc = 0; kc < KP8 / 4; ++kc) {
                    float4 lo;
                    lo.x = tf32_rn(h[4 * kc] - tf32_rn(h[4 * kc])), lo.y = tf32_rn(h[4 * kc + 1] - tf32_rn(h[4 * kc + 1]));
                    lo.z = tf32_rn(h[4 * kc + 2] - tf32_rn(h[4 * kc + 2])), lo.w = tf32_rn(h[4 * kc + 3] - tf32_rn(h[4 * kc + 3]));
                    sts128(sHlo + kc * SH_LBO + (s >> 3) * SH_SBO + (s & 7) * 16, lo);
                }
            }
            fence_proxy_async();
            mbar_arrive(bar_hready);

            mbar_wait(bar_whfull + 8 * b, (i >> 1) & 1);
            tc_fence_after();
            const uint32_t tWH = tmem + lane_off + (b ? TM_WH1 : TM_WH0);
            float kl = 0.f;
            {
                // software pipeline over the three boxes: the TMEM load of box c + 1 is in flight while box c is divided
                uint32_t v0[32], v1[32];
                const uint32_t rowbase = sX + st * XSTAGE_BYTES + s * 128;
                tmem_ld32(tWH, v0);
                tc_wait_ld();
                tmem_ld32(tWH + 32, v1);
                if (!valid) {  // rows past the end of X: x = 0 (TMA zero fill), make the quotient 0 * 1
#pragma unroll
                    for (int e = 0; e < 32; ++e) v0[e] = 0x3f800000u;
                }
                quotient_box<DO_R, DO_KL>(v0, rowbase, s, sw, kl);
                if (DO_R) tmem_st32(tWH, v0);
                if (p.dbg && blockIdx.x == 0 && i == 0) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) p.dbg[(size_t)s * VT + e] = __uint_as_float(v0[e]);
                }
                tc_wait_ld();
                tmem_ld32(tWH + 64, v0);
                if (!valid) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v1[e] = 0x3f800000u;
                }
                quotient_box<DO_R, DO_KL>(v1, rowbase + BOX_BYTES, s, sw, kl);
                if (DO_R) tmem_st32(tWH + 32, v1);
                if (p.dbg && blockIdx.x == 0 && i == 0) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) p.dbg[(size_t)s * VT + 32 + e] = __uint_as_float(v1[e]);
                }
                tc_wait_ld();
                if (!valid) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v0[e] = 0x3f800000u;
                }
                quotient_box<DO_R, DO_KL>(v0, rowbase + 2 * BOX_BYTES, s, sw, kl);
                if (DO_R) tmem_st32(tWH + 64, v0);
                if (p.dbg && blockIdx.x == 0 && i == 0) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) p.dbg[(size_t)s * VT + 64 + e] = __uint_as_float(v0[e]);
                }
            }
            if (do_kl && valid) obj_acc += (double)kl;
            if (do_r) {
                if (i > 0) mbar_wait(bar_shtfree, (i - 1) & 1);  // G3(i-1) has read sHT
                const uint32_t tbase = sHT + (s >> 2) * SHT_LBO + (s & 3) * 4;
#pragma unroll
                for (int j = 0; j < KP8; ++j)
                    if (j < k) sts32(tbase + (j >> 3) * SHT_SBO + (j & 7) * 16, tf32_rn(h[j]));
                tc_wait_st();
                fence_proxy_async();
            }
            tc_fence_before();
            mbar_arrive(bar_rready + 8 * b);

            if (do_h) {
                mbar_wait(bar_hnfull + 8 * b, (i >> 1) & 1);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(tmem + lane_off + (b ? TM_HN1 : TM_HN0), v);
                tc_wait_ld();
                if (p.dbg && blockIdx.x == 0 && i == 0) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) p.dbg[(size_t)TILE * VT + s * 32 + e] = __uint_as_float(v[e]);
                }
                if (valid) {
                    float* og = p.H_out + (size_t)(d0 + s) * k;
#pragma unroll
                    for (int j = 0; j < KP8; j += 4)
                        if (j < k) {
                            float4 o;
                            o.x = fmaxf(h[j] * __uint_as_float(v[j]), eps);
                            o.y = fmaxf(h[j + 1] * __uint_as_float(v[j + 1]), eps);
                            o.z = fmaxf(h[j + 2] * __uint_as_float(v[j + 2]), eps);
                            o.w = fmaxf(h[j + 3] * __uint_as_float(v[j + 3]), eps);
                            *reinterpret_cast<float4*>(og + j) = o;
                        }
                }
                tc_fence_before();
            }
        }
        // ---- per-CTA numerator partial: TMEM lane = feature, column = signature
        if (do_w && g == 0) {
            mbar_wait(bar_done, 0);
            tc_fence_after();
            if (q < 3) {
                uint32_t v[32];
                tmem_ld32(tmem + lane_off + TM_WN, v);
                tc_wait_ld();
                const int f = s;  // 0 .. 95
#pragma unroll
                for (int j = 0; j < KP8; ++j)
                    if (j < k) p.partial_wnum[((size_t)blockIdx.x * KP8 + j) * SAL_VMAX + f] = __uint_as_float(v[j]);
            }
            tc_fence_before();
        }
        if (do_kl) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) obj_acc += __shfl_xor_sync(0xffffffffu, obj_acc, o);
            if (lane == 0) s_red[warp - 4] = obj_acc;
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (do_kl && tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        p.partial_obj[blockIdx.x] = t;
    }
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)sym;
    return fn;
}

int encode_2d(CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint32_t box_inner, uint32_t box_rows,
              CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        sal_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SAL_EUNSUPPORTED;
    }
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {inner * 4};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        sal_set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner %llu rows %llu)", (int)r, (unsigned long long)inner,
                      (unsigned long long)rows);
        return SAL_EINVAL;
    }
    return 0;
}

template <int KP8, bool DO_R, bool DO_KL>
int launch_tc_v(sal_ctx* c, const PassArgs& a, cudaStream_t st) {
    using L = Lay<KP8>;
    static bool attr_set[16] = {false};
    if (!attr_set[c->device & 15]) {
        SAL_CUDA(cudaFuncSetAttribute(klnmf_pass_tc_kernel<KP8, DO_R, DO_KL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      L::DYN_BYTES));
        attr_set[c->device & 15] = true;
    }
    CUtensorMap mapX, mapH;
    if (int e = encode_2d(&mapX, a.X, VT, (uint64_t)c->D, 32, TILE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
    if (int e = encode_2d(&mapH, a.H_in, (uint64_t)c->k, (uint64_t)c->D, (uint32_t)c->k, TILE, CU_TENSOR_MAP_SWIZZLE_NONE)) return e;
    TcParams p;
    p.W = (const float*)a.W, p.H_out = (float*)a.H_out;
    p.partial_wnum = (float*)c->partial_wnum, p.partial_obj = c->partial_obj;
    p.dbg = (float*)c->dbg;
    p.D = c->D, p.k = c->k, p.flags = a.flags;
    p.n_tiles = (int)((c->D + TILE - 1) / TILE);
    const int grid = p.n_tiles < c->n_sm ? p.n_tiles : c->n_sm;
    klnmf_pass_tc_kernel<KP8, DO_R, DO_KL><<<grid, NTHREADS, L::DYN_BYTES, st>>>(mapX, mapH, p);
    SAL_CUDA(cudaGetLastError());
    c->launches++;
    return sal_launch_pass_reduce(c, a, grid, st);
}

template <int KP8>
int launch_tc(sal_ctx* c, const PassArgs& a, cudaStream_t st) {
    const bool r = a.flags & (SAL_PASS_UPDATE_H | SAL_PASS_WNUM), kl = a.flags & SAL_PASS_OBJECTIVE;
    if (r && !kl) return launch_tc_v<KP8, true, false>(c, a, st);
    if (!r && kl) return launch_tc_v<KP8, false, true>(c, a, st);
    return launch_tc_v<KP8, true, true>(c, a, st);
}

}  // namespace

bool sal_pass_tf32_supported(const sal_ctx* c, const PassArgs& a) {
    const int allowed = SAL_PASS_UPDATE_H | SAL_PASS_WNUM | SAL_PASS_OBJECTIVE;
    if (c->dtype != SAL_F32 || c->V != VT || c->k % 4 != 0 || c->k > 32) return false;
    if ((a.flags & ~allowed) || a.w_kl || a.w_lhalf || a.h_scale) return false;
    if (((uintptr_t)a.X | (uintptr_t)a.H_in | (uintptr_t)a.H_out) & 15) return false;
    if (c->D >= (int64_t)1 << 31) return false;
    return true;
}

int sal_launch_pass_tf32(sal_ctx* c, const PassArgs& a, cudaStream_t st) {
    switch (c->KP) {
        case 8: return launch_tc<8>(c, a, st);
        case 16: return launch_tc<16>(c, a, st);
        case 24: return launch_tc<24>(c, a, st);
        default: return launch_tc<32>(c, a, st);
    }
}
