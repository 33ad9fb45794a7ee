// tcgen05 / TMEM / TMA building blocks shared by the tensor-core kernels (klnmf_pass_tf32.cu: one fused pass per launch;
// klnmf_period_tf32.cu: a whole convergence-test period per launch).  Everything lives in an anonymous namespace: each
// translation unit gets its own copy.
#pragma once

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
constexpr int SMEM_LIMIT = 232448;              // 227 KB opt-in limit per CTA

// TMEM columns (512 allocated): WH / R x2, Hn x2, the persistent numerator, H operand (hi, lo) x2
constexpr uint32_t TM_WH0 = 0, TM_WH1 = 96, TM_HN0 = 192, TM_HN1 = 224, TM_WN = 256;
constexpr uint32_t TM_H0 = 288, TM_H1 = 352, TM_HLO = 32, TM_COLS = 512;

// strides of the thread-written operands (bytes)
constexpr int SW1_LBO = 1536, SW1_SBO = 128;  // sW1 [kc][feature/8][8][16B]  B of G1 (N = feature, K = signature)
constexpr int SW2_SBO = 128;                  // sW2 [fc][sig/8][8][16B]      B of G2 (N = signature, K = feature)
constexpr int SHT_SBO = 128;                  // sHT [sc][sig/8][8][16B](+16) B of G3 (N = signature, K = sample); LBO = Plan::sht_lbo
constexpr int NH = 4;                         // raw-H slots (load -> P0 -> output staging -> TMA store)

// Shared-memory plan for k signatures (KP8 = k rounded up to 8).  Only ceil(k/8) signature groups of sW2 are
// stored: G2 runs with N = 32 and the groups beyond them read the following bytes, finite garbage that only
// reaches output columns >= k which nobody reads.  The same holds for the 4th (non-existent) feature box of
// the G3 A operand: it reads the bytes after the stage, so at least 16 KB must follow the last stage.
struct Plan {
    int S;  // X / R stages
    int hraw, sw1, sw2, sw2_lbo, sht, sht_lbo;
    int off_hraw, off_w1hi, off_w1lo, off_w2, off_sht, off_bar, off_misc, total;
};
__host__ __device__ inline Plan make_plan(int k, int KP8, int extra = 0) {  // extra: bytes appended to the misc area
    Plan q;
    q.hraw = (TILE * k * 4 + 127) & ~127;
    q.sw1 = (KP8 / 4) * SW1_LBO;
    q.sw2_lbo = (KP8 / 8) * 128;
    q.sw2 = 24 * q.sw2_lbo;
    q.sht_lbo = (KP8 / 8) * 128 + 16;  // + 16: the per-sample scalar stores of a warp hit 32 different banks
    q.sht = 32 * q.sht_lbo;
    const int fixed = NH * q.hraw + 2 * q.sw1 + q.sw2 + q.sht + 32 * 8 + 128 + extra;
    q.S = (3 * XSTAGE_BYTES + fixed <= SMEM_LIMIT) ? 3 : 2;
    q.off_hraw = q.S * XSTAGE_BYTES;
    q.off_w1hi = q.off_hraw + NH * q.hraw;
    q.off_w1lo = q.off_w1hi + q.sw1;
    q.off_w2 = q.off_w1lo + q.sw1;
    q.off_sht = q.off_w2 + q.sw2;
    q.off_bar = q.off_sht + q.sht;
    q.off_misc = q.off_bar + 32 * 8;
    q.total = q.off_misc + 128 + extra;
    return q;
}

struct TcParams {
    const float* W;
    const float* H_in;
    const float* w_kl;     // per-sample weights of the KL term (weights_kl) or null
    const float* w_lhalf;  // per-sample l-half penalty weights (weights_lhalf) or null
    const float* h_scale;  // MvNMF line-search trial: exposures are read as clip(H * h_scale) and written back like that
    float* H_out;
    float* partial_wnum;
    double* partial_obj;
    float* dbg;  // optional diagnostics buffer (see sal_set_debug_buffer)
    int64_t D;
    int k, flags, n_tiles;
    int generic_k;  // k % 4 != 0: H rows are not 16-byte multiples, see the exposure loads below
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
// Bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.  The bound is wall-clock time
// (%globaltimer, looked at every 4096 polls), not a poll count: under a profiler's kernel replay or time slicing a wait can
// legitimately take many polls.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long SAL_WAIT_LIMIT_NS = 20ull * 1000 * 1000 * 1000;  // 20 s
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 4095u) == 0) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > SAL_WAIT_LIMIT_NS) __trap();
        }
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

// 3-D variants: with k % 4 != 0 a row of H (k floats) is not a multiple of 16 bytes, which a 2-D tensor map cannot
// describe.  A tile of 128 samples is still one contiguous run of 128 k floats, so H is viewed as [tile][k][128]
// (inner dimension 128 floats, then k with a 512-byte stride, then the tile) and box {128, k, 1} moves exactly that run.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c0),
                 "r"(c1), "r"(c2)
                 : "memory");
}

// one lane of a converged warp (the address arithmetic around it stays warp-uniform, i.e. in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred)::"memory");
    return pred != 0;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
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
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
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
__device__ __forceinline__ float lg2_approx(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
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
    // the loads of chunk m + 1 are issued before chunk m is divided and stored (LDS latency off the critical path)
    float4 va = lds128(rowbase + ((uint32_t)(0 ^ (s & 3)) << 5) + sw * 16);
    float4 vb = lds128(rowbase + ((uint32_t)(0 ^ (s & 3)) << 5) + (sw ^ 1) * 16);
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const uint32_t a32 = rowbase + ((uint32_t)(m ^ (s & 3)) << 5);
        const float4 xlo = sw ? vb : va, xhi = sw ? va : vb;
        if (m < 3) {
            const uint32_t n32 = rowbase + ((uint32_t)((m + 1) ^ (s & 3)) << 5);
            va = lds128(n32 + sw * 16), vb = lds128(n32 + (sw ^ 1) * 16);
        }
        const float xv[8] = {xlo.x, xlo.y, xlo.z, xlo.w, xhi.x, xhi.y, xhi.z, xhi.w};
        float rr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float wh = __uint_as_float(v[8 * m + e]);
            const float r = xv[e] * rcp_approx(wh);
            // KL term x ln(x/wh) - x + wh = x ln2 * lg2(r) + (wh - x), cancelling pair first.  lg2.approx (2^-22 absolute):
            // with D >= SAL_TF32_MIN_SAMPLES the per-term noise (~1e-7 x) averages to < 3e-8 of the objective.
            if (DO_KL) kl += xv[e] != 0.f ? fmaf(xv[e] * 0.693147180559945f, lg2_approx(r), wh - xv[e]) : wh;
            // round-to-nearest tf32 = add half an ulp; the MMA ignores the 13 low mantissa bits, no need to mask them
            rr[e] = __uint_as_float(__float_as_uint(r) + 0x1000u);
            v[8 * m + e] = __float_as_uint(rr[e]);
        }
        if (DO_R) {
            const float4 rlo = make_float4(rr[0], rr[1], rr[2], rr[3]), rhi = make_float4(rr[4], rr[5], rr[6], rr[7]);
            sts128(a32 + sw * 16, sw ? rhi : rlo);
            sts128(a32 + (sw ^ 1) * 16, sw ? rlo : rhi);
        }
    }
}

// The same box with the special-function work batched: the eight reciprocals of a chunk are issued back to back, then the
// eight logarithms (volatile asm keeps that order), so a warp has eight independent MUFU results in flight instead of one --
// ptxas otherwise schedules every MUFU right in front of its consumer (to save registers), which makes the KL variant
// latency bound with one or two epilogue warps per scheduler (profiles/r02_period_kl_ncu.md).  Two KL accumulators.
__device__ __forceinline__ float rcp_approx_v(float x) {
    float r;
    asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float lg2_approx_v(float x) {
    float r;
    asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
template <bool DO_R, bool DO_KL>
__device__ __forceinline__ void quotient_box_batched(uint32_t (&v)[32], uint32_t rowbase, int s, uint32_t sw, float& kl) {
    float4 va = lds128(rowbase + ((uint32_t)(0 ^ (s & 3)) << 5) + sw * 16);
    float4 vb = lds128(rowbase + ((uint32_t)(0 ^ (s & 3)) << 5) + (sw ^ 1) * 16);
    float kl0 = 0.f, kl1 = 0.f;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const uint32_t a32 = rowbase + ((uint32_t)(m ^ (s & 3)) << 5);
        const float4 xlo = sw ? vb : va, xhi = sw ? va : vb;
        if (m < 3) {
            const uint32_t n32 = rowbase + ((uint32_t)((m + 1) ^ (s & 3)) << 5);
            va = lds128(n32 + sw * 16), vb = lds128(n32 + (sw ^ 1) * 16);
        }
        const float xv[8] = {xlo.x, xlo.y, xlo.z, xlo.w, xhi.x, xhi.y, xhi.z, xhi.w};
        float rc[8], r[8], rr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) rc[e] = rcp_approx_v(__uint_as_float(v[8 * m + e]));
#pragma unroll
        for (int e = 0; e < 8; ++e) r[e] = xv[e] * rc[e];
        if (DO_KL) {
            // KL term x ln(x/wh) - x + wh = x ln2 * lg2(r) + (wh - x), cancelling pair first.  lg2.approx (2^-22 absolute):
            // with D >= SAL_TF32_MIN_SAMPLES the per-term noise (~1e-7 x) averages to < 3e-8 of the objective.
            // The chunk's contribution is ONE dependent FMA chain that starts with the logarithm issued LAST, so no consumer can
            // be scheduled before all eight logarithms are in flight.  x = 0 contributes wh: coefficient 0 (the logarithm of
            // the clamped quotient stays finite) plus wh - x.
            float lg[8], cf[8], dsum = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                cf[e] = xv[e] * 0.693147180559945f;
                dsum += __uint_as_float(v[8 * m + e]) - xv[e];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) lg[e] = lg2_approx_v(fmaxf(r[e], 1e-37f));
            float t = dsum;
#pragma unroll
            for (int e = 7; e >= 0; --e) t = fmaf(cf[e], lg[e], t);
            if (m & 1)
                kl1 += t;
            else
                kl0 += t;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            // round-to-nearest tf32 = add half an ulp; the MMA ignores the 13 low mantissa bits, no need to mask them
            rr[e] = __uint_as_float(__float_as_uint(r[e]) + 0x1000u);
            v[8 * m + e] = __float_as_uint(rr[e]);
        }
        if (DO_R) {
            const float4 rlo = make_float4(rr[0], rr[1], rr[2], rr[3]), rhi = make_float4(rr[4], rr[5], rr[6], rr[7]);
            sts128(a32 + sw * 16, sw ? rhi : rlo);
            sts128(a32 + (sw ^ 1) * 16, sw ? rlo : rhi);
        }
    }
    if (DO_KL) kl += kl0 + kl1;
}

// tcgen05.ld of 8 consecutive columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : SAL_R8(v, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8_(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" ::SAL_W8(v, 0), "r"(taddr) : "memory");
}

// The quotient of one sample row (96 features) as a LOOP over twelve chunks of eight features -- TMEM columns 8 c .. 8 c + 7,
// the 32-byte piece (c & 3) ^ (s & 3) of box c >> 2 -- two chunks per iteration, the next chunk's TMEM / shared-memory loads in
// flight while the current one is divided.  The unrolled form (three 32-column boxes, ~4,000 instructions per tile and variant)
// ran at the speed of the instruction fetch, not of the pipes (ncu: stall_no_inst; profiles/r02_period_kl_ncu.md): this body is
// ~150 instructions and stays in the instruction cache.  Rows past the end of X need no special case here: their exposures are
// read as 1 (see load_h), so WH > 0 and the zero-filled X gives a zero quotient.  Returns the row's KL term (DO_KL).
template <bool DO_R, bool DO_KL>
__device__ __forceinline__ float quotient_row_loop(uint32_t tWH, uint32_t rowbase, int s, uint32_t sw, float* dbg_row = nullptr) {
    float kl0 = 0.f, kl1 = 0.f;
    const uint32_t s3 = (uint32_t)(s & 3);
    const uint32_t off_a = sw * 16, off_b = (sw ^ 1u) * 16;
    auto chunk_addr = [&](int c) { return rowbase + (uint32_t)(c >> 2) * (uint32_t)BOX_BYTES + ((((uint32_t)c & 3u) ^ s3) << 5); };
    auto process = [&](int c, const uint32_t (&w)[8], const float4& va, const float4& vb, float& kl) {
        const float4 xlo = sw ? vb : va, xhi = sw ? va : vb;
        const float xv[8] = {xlo.x, xlo.y, xlo.z, xlo.w, xhi.x, xhi.y, xhi.z, xhi.w};
        float r[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) r[e] = xv[e] * rcp_approx(__uint_as_float(w[e]));
        if (DO_KL) {
            // KL term x ln(x/wh) - x + wh = ln2 * x lg2(r) + (wh - x), cancelling pair first.  lg2.approx (2^-22 absolute): with
            // D >= SAL_TF32_MIN_SAMPLES the per-term noise (~1e-7 x) averages to < 3e-8 of the objective.  x = 0 contributes wh:
            // the clamp keeps the logarithm finite and its coefficient x is 0.
            float lg[8], dsum = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                lg[e] = lg2_approx(fmaxf(r[e], 1e-37f));
                dsum += __uint_as_float(w[e]) - xv[e];
            }
            float t = 0.f;
#pragma unroll
            for (int e = 7; e >= 0; --e) t = fmaf(xv[e], lg[e], t);
            kl += fmaf(t, 0.693147180559945f, dsum);
        }
        if (DO_R) {
            uint32_t rb[8];
            // round-to-nearest tf32 = add half an ulp; the MMA ignores the 13 low mantissa bits, no need to mask them
#pragma unroll
            for (int e = 0; e < 8; ++e) rb[e] = __float_as_uint(r[e]) + 0x1000u;
            if (dbg_row) {  // diagnostics: the quotient as the MMAs will read it
#pragma unroll
                for (int e = 0; e < 8; ++e) dbg_row[8 * c + e] = __uint_as_float(rb[e]);
            }
            const float4 rlo = make_float4(__uint_as_float(rb[0]), __uint_as_float(rb[1]), __uint_as_float(rb[2]), __uint_as_float(rb[3]));
            const float4 rhi = make_float4(__uint_as_float(rb[4]), __uint_as_float(rb[5]), __uint_as_float(rb[6]), __uint_as_float(rb[7]));
            const uint32_t a = chunk_addr(c);
            sts128(a + off_a, sw ? rhi : rlo);
            sts128(a + off_b, sw ? rlo : rhi);
            tmem_st8_(tWH + 8 * c, rb);
        }
    };
    // two chunks (16 TMEM columns, one tcgen05.ld) per step; six steps, two per loop iteration
    uint32_t wa[16], wb[16];
    float4 xa[4], xb[4];
    auto fetch = [&](int c, uint32_t (&w)[16], float4 (&x)[4]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : SAL_R8(w, 0), SAL_R8(w, 8)
            : "r"(tWH + 8 * c)
            : "memory");
        const uint32_t a0 = chunk_addr(c), a1 = chunk_addr(c + 1);
        x[0] = lds128(a0 + off_a), x[1] = lds128(a0 + off_b), x[2] = lds128(a1 + off_a), x[3] = lds128(a1 + off_b);
    };
    auto process2 = [&](int c, const uint32_t (&w)[16], const float4 (&x)[4]) {
        const uint32_t w0[8] = {w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]};
        const uint32_t w1[8] = {w[8], w[9], w[10], w[11], w[12], w[13], w[14], w[15]};
        process(c, w0, x[0], x[1], kl0);
        process(c + 1, w1, x[2], x[3], kl1);
    };
    fetch(0, wa, xa);
#pragma unroll 1
    for (int c = 0; c < 12; c += 4) {
        tc_wait_ld();  // chunks c, c + 1 have arrived
        fetch(c + 2, wb, xb);
        process2(c, wa, xa);
        tc_wait_ld();  // chunks c + 2, c + 3 have arrived
        if (c + 4 < 12) fetch(c + 4, wa, xa);
        process2(c + 2, wb, xb);
    }
    return kl0 + kl1;
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

// Diagnostics timeline (only when a debug buffer is set): CTA 0 stamps clock64 per role / tile / phase.
constexpr int DBG_TL_OFF = TILE * VT + TILE * 32, DBG_TL_TILES = 48, DBG_TL_SLOTS = 8;
__device__ __forceinline__ void stamp(float* dbg, bool on, int role, int tile, int slot) {
    if (on && tile < DBG_TL_TILES)
        reinterpret_cast<unsigned int*>(dbg)[DBG_TL_OFF + (role * DBG_TL_TILES + tile) * DBG_TL_SLOTS + slot] = (unsigned int)clock64();
}

// tcgen05.st of 8 consecutive columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};" ::SAL_W8(v, 0), "r"(taddr) : "memory");
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

// H [D][k] with k % 4 != 0 as [full tiles][k][128]: see tma_load_3d.  Only whole tiles are described.
int encode_h3d(CUtensorMap* map, const void* ptr, int k, int64_t D) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        sal_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return SAL_EUNSUPPORTED;
    }
    const int64_t n_full = D / TILE;
    const cuuint64_t dims[3] = {(cuuint64_t)TILE, (cuuint64_t)k, (cuuint64_t)(n_full > 0 ? n_full : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)TILE * 4, (cuuint64_t)TILE * k * 4};
    const cuuint32_t box[3] = {(cuuint32_t)TILE, (cuuint32_t)k, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        sal_set_error("cuTensorMapEncodeTiled (3-D exposure view) failed with CUresult %d (k %d, D %lld)", (int)r, k, (long long)D);
        return SAL_EINVAL;
    }
    return 0;
}


}  // namespace
