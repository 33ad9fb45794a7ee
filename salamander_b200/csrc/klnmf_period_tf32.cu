// A whole convergence-test period of KL-NMF joint updates in ONE persistent launch (tcgen05 kind::tf32, fp32 handles, V = 96).
//
// Reference: the body of the fit loop, models/signature_nmf.py:361-380, around update_WH (models/_utils_klnmf.py:281-361) and
// kl_divergence (:11-55).  The per-tile arithmetic is that of klnmf_pass_tf32.cu (same tf32 recipe, same layouts); what is new is
// everything BETWEEN two updates, which used to be two kernel boundaries, a reduction kernel and -- on several GPUs -- an
// exchange, ~12-19 us of dependency chain per update (profiles/r01_scaling.md):
//
//   * the CTAs stay resident (cooperative launch, one CTA per SM) and loop over the updates of the period;
//   * the per-CTA numerator partial leaves TMEM as sequence-TAGGED words {float, tag}: data and flag travel together, so
//     consumers poll the data itself -- no counters, no fences, no grid barrier;
//   * the reduction of the G partials is spread over ALL CTAs: CTA c sums a slice of ~k*96/G values (fixed order, double) and,
//     on several GPUs, exchanges exactly that slice with the peers over NVLink (tagged 16-byte words pushed into every peer's
//     receive buffer, summed in rank order: bit-identical on all ranks), then publishes the totals, again as tagged words;
//   * every CTA polls the k*96 totals and applies the W epilogue itself (colnorm, given signatures, clip: reference
//     _utils_klnmf.py:338-341) straight into its shared-memory W operands -- W never goes through global memory between updates;
//   * the data-movement warps run free across updates: X tiles of update u + 1 are requested while update u drains (X never
//     changes), the exposures of update u + 1 as soon as this CTA's stores of update u have completed (a CTA always meets the
//     same tiles), both while the reduction is in flight;
//   * the KL objective of the incoming iterate rides on every `obj_every`-th update (fused-KL variant of the tile loop) and an
//     optional objective-only sweep after the last update gives the objective of the final iterate
//     (signature_nmf.py:365-380 evaluates it every conv_test_freq iterations).
//
// Emulated ranks (tests on one GPU): the grid may hold n_virtual independent "ranks", G CTAs each, every one with its own shard,
// buffers and receive buffer; they exchange through the same protocol.  One cooperative launch, so the ranks are co-resident by
// construction (separate launches that wait on one another must never share a GPU).
#include "tc_common.cuh"

#include <string.h>

#include <type_traits>

namespace {

// 12 warps = 3 warpgroups.  Warpgroup 0: warp 0 data movement (X loads, exposure loads and stores: one event loop), warp 1 MMA
// issuer (+ TMEM allocation), warps 2 - 3 idle; warpgroups 1 and 2: the two epilogue warpgroups.  The register file is split
// per SM sub-partition (3 warps each: 168 registers per thread at launch), which the epilogue's tile loop does not fit in --
// so after the set-up warpgroup 0 gives registers back (setmaxnreg.dec) and the epilogue warpgroups take them (setmaxnreg.inc).
constexpr int PER_THREADS = 384;
constexpr int EW0 = 4;              // first epilogue warp
constexpr int REGS_UTIL = 88, REGS_EPI = 208;  // 88 + 2 * 208 = 504 = 3 * 168: the pool is what the launch allocated
constexpr int EPI0 = EW0 * 32;      // first epilogue thread
constexpr int PER_MAX_RANKS = 8;    // one NVSwitch box (the exchange handles up to 32: one lane per peer)
constexpr int PER_MAX_VIRTUAL = 2;
constexpr int PER_EXTRA = 2432;  // appended to the misc area: s_part [8][32] doubles (2048), s_tot [32] doubles (256), slack
constexpr int PER_WREG = 12;        // W elements a thread owns: up to 4 signatures x 3 features
// The owned W elements (full fp32) survive the tile loops in otherwise unused TMEM columns -- the tile loop has no
// registers to spare (168 per thread at 384 threads): 16 columns per epilogue warpgroup, lane = the thread's own TMEM lane.
constexpr uint32_t TM_WST = 416;

struct PeriodRank {
    CUtensorMap mapX, mapH0, mapH1;  // H0: exposures read by the first sweep; H1: written by every update, read by the later ones
    const float* W_in;
    float* W_out;
    const float* H_in;
    float* H_out;
    uint2* partials;    // [2][G][k * 96] {float, tag}
    uint2* sums;        // [2][k * 96]    {float, tag}
    uint4* obj_part;    // [2][G]         {lo, tag, hi, tag}
    double* objective;  // [n_objectives]
    void* const* peers; // device array [n_ranks] of receive buffers (null: single rank)
    unsigned int* seq;  // device scalar: tag of the next update; advanced by the kernel
    int64_t D;
    int n_tiles, rank;
};

struct PeriodParams {
    PeriodRank r[PER_MAX_VIRTUAL];
    unsigned long long* tl;  // optional timeline (diagnostics)
    int n_virtual, G, k, n_updates, obj_every, final_obj, n_given, clip_given, n_ranks;
};

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {  // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bar_epi() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint2 ld_tag2(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_tag2(uint2* p, unsigned int val, unsigned int tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(val), "r"(tag) : "memory");
}
__device__ __forceinline__ void st_tag4(uint4* p, unsigned int lo, unsigned int hi, unsigned int tag) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(tag), "r"(hi), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ld_tag4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// Spin until the word carries the tag (a lost producer traps after SAL_WAIT_LIMIT_NS instead of hanging the box).
__device__ __forceinline__ float poll_f32(const uint2* p, unsigned int tag) {
    uint2 w = ld_tag2(p);
    if (w.y != tag) {
        unsigned int spins = 0;
        unsigned long long t0 = 0;
        do {
            w = ld_tag2(p);
            if ((++spins & 1023u) == 0) {
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > SAL_WAIT_LIMIT_NS) __trap();
            }
        } while (w.y != tag);
    }
    return __uint_as_float(w.x);
}
__device__ __forceinline__ double poll_f64(const uint4* p, unsigned int tag) {
    uint4 w = ld_tag4(p);
    if (w.y != tag || w.w != tag) {
        unsigned int spins = 0;
        unsigned long long t0 = 0;
        do {
            w = ld_tag4(p);
            if ((++spins & 1023u) == 0) {
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > SAL_WAIT_LIMIT_NS) __trap();
            }
        } while (w.y != tag || w.w != tag);
    }
    return __longlong_as_double((long long)(((unsigned long long)w.z << 32) | w.x));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : SAL_R8(v, 0), SAL_R8(v, 8)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};" ::SAL_W8(v, 0),
        SAL_W8(v, 8), "r"(taddr)
        : "memory");
}

// Batched polling: every word that does not carry the tag yet is re-requested in the SAME round, so a retry costs one memory
// round trip for the whole batch instead of one per word (a thread that arrives early -- the usual case -- would otherwise walk
// through its words one round trip at a time).  `pending` is a bit mask over the N slots; addr(z) gives slot z's address.
template <int N, class Addr>
__device__ __forceinline__ void poll_batch(uint2 (&w)[N], unsigned int pending, unsigned int tag, Addr addr) {
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (pending) {
#pragma unroll
        for (int z = 0; z < N; ++z)
            if ((pending >> z) & 1u) w[z] = ld_tag2(addr(z));
#pragma unroll
        for (int z = 0; z < N; ++z)
            if (((pending >> z) & 1u) && w[z].y == tag) pending &= ~(1u << z);
        if (pending && (++spins & 255u) == 0) {
            const unsigned long long now = global_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > SAL_WAIT_LIMIT_NS) __trap();
        }
    }
}

// timeline: slot s of update u, written by one thread of CTA 1 (or 0 when the grid has one CTA)
constexpr int TL_SLOTS = 8;
__device__ __forceinline__ void tl_stamp(unsigned long long* tl, bool on, int u, int s) {
    if (on) tl[u * TL_SLOTS + s] = global_ns();
}
// per-CTA stamps behind the timeline: [sweep][CTA][4] (0 own tiles done, 1 all MMAs retired, 2 slice published, 3 W epilogue done)
__device__ __forceinline__ void cta_stamp(unsigned long long* tl, bool on, int U, int G, int u, int c, int s) {
    if (on) tl[(U + 1) * TL_SLOTS + ((size_t)u * G + c) * 4 + s] = global_ns();
}

template <int KP8, bool GK>
__global__ void __launch_bounds__(PER_THREADS, 1) klnmf_period_tc_kernel(const __grid_constant__ PeriodParams P) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    const uint32_t base = smem_u32(smem_dyn);
    if (base & 1023u) __trap();  // the swizzled TMA boxes need 1024-byte alignment
    const int k = P.k;
    const Plan q = make_plan(k, KP8, PER_EXTRA);
    const int S = q.S;
    const uint32_t sX = base, sHraw = base + q.off_hraw, sW1hi = base + q.off_w1hi, sW1lo = base + q.off_w1lo;
    const uint32_t sW2 = base + q.off_w2, sHT = base + q.off_sht, bars = base + q.off_bar;
    const int SHT_LBO = q.sht_lbo;
    // barriers: full[3] empty[3] hready[2] whfull[2] rready[2] hnfull[2] shtfree done hfull[4] hempty[4] hout[4] wn
    const uint32_t bar_full = bars, bar_empty = bars + 24, bar_hready = bars + 48, bar_whfull = bars + 64;
    const uint32_t bar_rready = bars + 80, bar_hnfull = bars + 96, bar_shtfree = bars + 112, bar_done = bars + 120;
    const uint32_t bar_hfull = bars + 128, bar_hempty = bars + 160, bar_hout = bars + 192, bar_wn = bars + 224;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_dyn + q.off_misc);
    double* s_obj = reinterpret_cast<double*>(smem_dyn + q.off_misc + 16);   // [8]
    double* s_part = reinterpret_cast<double*>(smem_dyn + q.off_misc + 128);        // [8][32] warp partials / receive staging
    double* s_tot = reinterpret_cast<double*>(smem_dyn + q.off_misc + 128 + 2048);  // [32]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = P.G, vr = (int)blockIdx.x / G, c = (int)blockIdx.x - vr * G;
    const PeriodRank& R = P.r[vr];
    const int n_my = (R.n_tiles - c + G - 1) / G;  // tiles of this CTA per sweep (>= 1)
    const int L = P.n_updates, U = L + (P.final_obj ? 1 : 0);  // sweeps over X: L updates (+ one objective-only sweep)
    const int nvals = k * VT;
    const unsigned int seq0 = *(volatile unsigned int*)R.seq;
    const bool tl_on = P.tl != nullptr && c == (G > 1 ? 1 : 0) && vr == 0;

    // ---- one-time setup --------------------------------------------------------------------------
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&R.mapX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&R.mapH0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&R.mapH1) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(bar_full + 8 * i, 1), mbar_init(bar_empty + 8 * i, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_hready + 8 * i, 128);
            mbar_init(bar_whfull + 8 * i, 1);
            mbar_init(bar_rready + 8 * i, 128);
            mbar_init(bar_hnfull + 8 * i, 1);
        }
        mbar_init(bar_shtfree, 1);
        mbar_init(bar_done, 1);
        mbar_init(bar_wn, 1);
        for (int i = 0; i < NH; ++i) mbar_init(bar_hfull + 8 * i, 1), mbar_init(bar_hempty + 8 * i, 1), mbar_init(bar_hout + 8 * i, 128);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(TM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < q.sht / 4; i += PER_THREADS) sts32(sHT + 4 * i, 0.f);
    // W of the incoming iterate: epilogue warp ew owns the signatures j = ew, ew + 8, ... (three features per lane: lane,
    // lane + 32, lane + 64), so the row sums of the W epilogue never leave the warp.  Rows >= k of the operands are zero.
    uint32_t wst[16];  // the thread's W elements as bits (prologue only; stashed in TMEM afterwards)
    auto store_w_operands = [&](int j, int f, float w) {  // tf32 hi + lo (round to nearest)
        const float hi = tf32_rn(w), lo = tf32_rn(w - hi);
        const uint32_t o1 = (j >> 2) * SW1_LBO + (f >> 3) * SW1_SBO + (f & 7) * 16 + (j & 3) * 4;
        sts32(sW1hi + o1, hi), sts32(sW1lo + o1, lo);
        sts32(sW2 + (f >> 2) * q.sw2_lbo + (j >> 3) * SW2_SBO + (j & 7) * 16 + (f & 3) * 4, hi);
    };
    if (warp >= EW0) {
#pragma unroll
        for (int r = 0; r < PER_WREG / 3; ++r) {
            const int j = (warp - EW0) + 8 * r;
#pragma unroll
            for (int s3 = 0; s3 < 3; ++s3) {
                const int f = s3 * 32 + lane;
                float w = 0.f;
                if (j < KP8) {
                    if (j < k) w = __ldcg(R.W_in + (size_t)j * VT + f);
                    store_w_operands(j, f, w);
                }
                wst[3 * r + s3] = __float_as_uint(w);
            }
        }
#pragma unroll
        for (int r = PER_WREG; r < 16; ++r) wst[r] = 0u;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_wst = tmem + ((uint32_t)((warp & 3) * 32) << 16) + TM_WST + (uint32_t)(((warp - EW0) >> 2) & 1) * 16u;
    if (warp >= EW0) {
        tmem_st16(t_wst, wst);
        tc_wait_st();
    }

    // sweep u: update (u < L) or the trailing objective-only sweep; the KL term rides on every obj_every-th update
    auto sweep_kl = [&](int u) { return u >= L || (P.obj_every > 0 && u % P.obj_every == 0); };

    // (each setmaxnreg dominates exactly its own role code: the register budget of a region is what its dominating
    // setmaxnreg says)
    if (warp < EW0) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_UTIL));
      if (warp == 0) {
        // ================= data movement: one event loop =================
        // Three queues served by one warp with NON-blocking barrier tests (warp-uniform control flow, lane 0 issues):
        //   X tiles     : free-running, S tiles ahead, across sweeps (X never changes);
        //   exposures in: sweep 0 reads H_in, later sweeps read back what this CTA stored in the sweep before -- a CTA always
        //                 meets the same tiles -- so they wait until those bulk stores have completed (known locally: the same
        //                 thread issued them);
        //   exposures out: the epilogue leaves the updated rows in the tile's slot; one bulk store each.
        // The partial last tile of a generic-k problem cannot go through the 3-D map: the whole warp copies its rows with
        // plain loads / stores.
        const int T = U * n_my;
        int tx = 0, ix = 0, sx = 0, px = 1;           // X: running tile, tile in sweep, stage, parity to test on `empty`
        int th = 0, ih = 0, uh = 0, sh = 0, ph = 1;   // exposure loads: running tile, tile in sweep, sweep, slot, parity on `hempty`
        int ts = 0, is = 0, us = 0, ss = 0, ps = 0;   // exposure stores: ..., slot, parity on `hout`
        const bool has_ragged = GK && (int64_t)(c + (n_my - 1) * G) * TILE + TILE > R.D;  // this CTA owns the partial last tile
        while (ts < T) {  // (the last store is the last event)
            bool idle = true;
            if (tx < T && mbar_test(bar_empty + 8 * sx, (uint32_t)px)) {
                if (lane == 0) {
                    const int d0 = (c + ix * G) * TILE;
                    mbar_arrive_expect_tx(bar_full + 8 * sx, XSTAGE_BYTES);
                    for (int b = 0; b < NBOX; ++b) tma_load_2d(sX + sx * XSTAGE_BYTES + b * BOX_BYTES, &R.mapX, bar_full + 8 * sx, b * 32, d0);
                }
                idle = false;
                ++tx;
                if (++ix == n_my) ix = 0;
                if (++sx == S) sx = 0, px ^= 1;
            }
            // rows of sweep uh >= 1 were written by this CTA's store of running tile th - n_my: it must have been issued, and at
            // most the `newer` bulk groups committed after it may still be pending (groups complete in order)
            if (th < T && (uh == 0 || th - n_my < ts) && mbar_test(bar_hempty + 8 * sh, (uint32_t)ph)) {
                if (uh > 0 && lane == 0) {
                    const int newer = has_ragged ? 0 : ts - (th - n_my) - 1;
                    if (newer >= 3)
                        asm volatile("cp.async.bulk.wait_group 3;" ::: "memory");
                    else if (newer == 2)
                        asm volatile("cp.async.bulk.wait_group 2;" ::: "memory");
                    else if (newer == 1)
                        asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
                    else
                        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                    // (store and load both go through the async proxy, issued by this thread: no proxy fence needed)
                }
                __syncwarp();
                const CUtensorMap* map = uh == 0 ? &R.mapH0 : &R.mapH1;
                const int tile = c + ih * G, d0 = tile * TILE;
                const bool ragged = GK && (int64_t)d0 + TILE > R.D;
                if (!ragged) {
                    if (lane == 0) {
                        mbar_arrive_expect_tx(bar_hfull + 8 * sh, (uint32_t)(TILE * k * 4));
                        if (GK)
                            tma_load_3d(sHraw + sh * q.hraw, map, bar_hfull + 8 * sh, 0, 0, tile);
                        else
                            tma_load_2d(sHraw + sh * q.hraw, map, bar_hfull + 8 * sh, 0, d0);
                    }
                } else {
                    const float* src = (uh == 0 ? R.H_in : R.H_out) + (size_t)d0 * k;
                    const int n = (int)(R.D - d0) * k;
                    for (int x = lane; x < TILE * k; x += 32) sts32(sHraw + sh * q.hraw + x * 4, x < n ? __ldcg(src + x) : 0.f);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_hfull + 8 * sh);
                }
                idle = false;
                ++th;
                if (++ih == n_my) ih = 0, ++uh;
                if (++sh == NH) sh = 0, ph ^= 1;
            }
            if (mbar_test(bar_hout + 8 * ss, (uint32_t)ps)) {
                const int tile = c + is * G, d0 = tile * TILE;
                const bool ragged = GK && (int64_t)d0 + TILE > R.D;
                if (us < L) {
                    if (ragged) {
                        const int n = (int)(R.D - d0) * k;
                        float* dst = R.H_out + (size_t)d0 * k;
                        for (int x = lane; x < n; x += 32) dst[x] = lds32(sHraw + ss * q.hraw + x * 4);
                        __syncwarp();
                    } else if (lane == 0) {
                        if (GK)
                            tma_store_3d(&R.mapH1, sHraw + ss * q.hraw, 0, 0, tile);
                        else
                            tma_store_2d(&R.mapH1, sHraw + ss * q.hraw, 0, d0);
                        tma_store_commit_and_wait_read();
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_hempty + 8 * ss);
                ++ts;
                if (++ss == NH) ss = 0, ps ^= 1;
                if (++is == n_my) is = 0, ++us;
                idle = false;
            }
            if (idle) __nanosleep(64);  // leave the issue slots of this scheduler to the epilogue warps
        }
        if (lane == 0) tma_store_wait_all();
      } else if (warp == 1) {
        // ================= MMA issuer =================
        constexpr uint32_t ID1 = make_idesc(128, VT, 0, 0);
        constexpr uint32_t ID2 = make_idesc(128, N2, 0, 0);
        constexpr uint32_t ID3 = make_idesc(128, N2, 1, 0);
        const uint64_t dW1hi = make_desc(sW1hi, SW1_LBO, SW1_SBO, LAYOUT_NONE), dW1lo = make_desc(sW1lo, SW1_LBO, SW1_SBO, LAYOUT_NONE);
        const uint64_t dW2 = make_desc(sW2, q.sw2_lbo, SW2_SBO, LAYOUT_NONE);
        const uint64_t dHT = make_desc(sHT, SHT_LBO, SHT_SBO, LAYOUT_NONE);
        const uint32_t htstep = (uint32_t)(2 * SHT_LBO) >> 4;
        const uint64_t dX0 = make_desc(sX, BOX_BYTES, 512, LAYOUT_128B_BASE32B);
        const uint32_t w2step = (uint32_t)(2 * q.sw2_lbo) >> 4;
        auto issue_g1 = [&](int t) {
            const int b = t & 1;
            mbar_wait(bar_hready + 8 * b, (t >> 1) & 1);
            tc_fence_after();
            const uint32_t d = tmem + (b ? TM_WH1 : TM_WH0), th = tmem + (b ? TM_H1 : TM_H0);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < KP8 / 8; ++ks) {
                    const uint64_t bhi = dW1hi + (uint64_t)((ks * 2 * SW1_LBO) >> 4), blo = dW1lo + (uint64_t)((ks * 2 * SW1_LBO) >> 4);
                    mma_ts(d, th + ks * 8, bhi, ID1, ks > 0);      // hi * hi
                    mma_ts(d, th + TM_HLO + ks * 8, bhi, ID1, 1);  // lo * hi
                    mma_ts(d, th + ks * 8, blo, ID1, 1);           // hi * lo
                }
                tc_commit(bar_whfull + 8 * b);
            }
            __syncwarp();
        };
        for (int u = 0; u < U; ++u) {
            const bool do_r = u < L;
            // the first G1 of a sweep waits (through hready) for the W operands the epilogue warps rewrote after sweep u - 1
            issue_g1(u * n_my);
            for (int i = 0; i < n_my; ++i) {
                const int t = u * n_my + i;
                if (i + 1 < n_my) issue_g1(t + 1);
                const int st = t % S, b = t & 1;
                mbar_wait(bar_rready + 8 * b, (t >> 1) & 1);
                tc_fence_after();
                if (do_r) {
                    const uint32_t tR = tmem + (b ? TM_WH1 : TM_WH0), tHn = tmem + (b ? TM_HN1 : TM_HN0);
                    const uint64_t dX = dX0 + (uint64_t)((uint32_t)(st * XSTAGE_BYTES) >> 4);
                    if (elect_one()) {
                        // G3 first: it is the last reader of the X / R stage and of sHT; G2 (TMEM operand) follows
#pragma unroll
                        for (int ks = 0; ks < TILE / 8; ++ks)
                            mma_ss(tmem + TM_WN, dX + (uint64_t)(ks * (1024 >> 4)), dHT + (uint64_t)(ks * htstep), ID3, (i > 0 || ks > 0));
                        tc_commit(bar_empty + 8 * st);
                        tc_commit(bar_shtfree);
                        if (i == n_my - 1) tc_commit(bar_wn);  // phase u: the numerator of sweep u is complete (G2 of this tile is not)
#pragma unroll
                        for (int ks = 0; ks < VT / 8; ++ks) mma_ts(tHn, tR + ks * 8, dW2 + (uint64_t)(ks * w2step), ID2, ks > 0);
                        tc_commit(bar_hnfull + 8 * b);
                    }
                    __syncwarp();
                } else if (lane == 0) {  // objective only: nothing reads the stage after E1
                    mbar_arrive(bar_empty + 8 * st);
                    mbar_arrive(bar_shtfree);
                }
            }
            if (elect_one()) tc_commit(bar_done);  // phase u: every MMA of sweep u has retired
            __syncwarp();
        }
        mbar_wait(bar_done, (U - 1) & 1);  // before the CTA tears TMEM down
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
        // ================= epilogue warpgroups (+ reduction, exchange and W epilogue between the updates) =================
        const int qw = warp & 3;
        const int s = qw * 32 + lane;  // sample row of the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)(qw * 32) << 16;
        const uint32_t sw = (s >> 2) & 1;
        const float eps = (float)SAL_EPS_F32;
        const int g = (warp - EW0) >> 2;  // this warpgroup meets the tiles whose running number t is congruent g mod 2

        // exposures of sample s of running tile t -> registers and, as tf32 hi / lo, the TMEM A operand of G1
        // (stage / arrive: the first tile of a sweep is staged while the reduction of the previous sweep is still in flight --
        // the exposures do not depend on W -- and announced to the MMA warp once the W operands have been rewritten)
        auto load_h = [&](int t, int i, float (&h)[KP8], bool stage = true, bool arrive = true) {
            const int hs = t % NH, b = t & 1;
            if (!stage) {
                mbar_arrive(bar_hready + 8 * b);
                return;
            }
            mbar_wait(bar_hfull + 8 * hs, (t / NH) & 1);
            const uint32_t hrow = sHraw + hs * q.hraw + s * (k * 4);
            const uint32_t th = tmem + lane_off + (b ? TM_H1 : TM_H0);
            // rows past the end of X (zero-filled by TMA): exposures read as 1, so that WH > 0 and the zero counts give a zero
            // quotient (no contribution to anything; the rows are clipped again when the tile is stored)
            const bool row_valid = (int64_t)(c + i * G) * TILE + s < R.D;
#pragma unroll
            for (int j = 0; j < KP8; j += 8) {
                if (GK) {  // rows are not 16-byte aligned: scalar loads
#pragma unroll
                    for (int x = 0; x < 8; ++x) h[j + x] = j + x < k ? lds32(hrow + (j + x) * 4) : 0.f;
                } else {
                    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
                    if (j < k) t0 = lds128(hrow + j * 4);
                    if (j + 4 < k) t1 = lds128(hrow + j * 4 + 16);
                    h[j] = t0.x, h[j + 1] = t0.y, h[j + 2] = t0.z, h[j + 3] = t0.w;
                    h[j + 4] = t1.x, h[j + 5] = t1.y, h[j + 6] = t1.z, h[j + 7] = t1.w;
                }
                if (!row_valid) {
#pragma unroll
                    for (int x = 0; x < 8; ++x)
                        if (j + x < k) h[j + x] = 1.f;
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    hi[x] = tf32_bits(h[j + x]);
                    lo[x] = tf32_bits(h[j + x] - __uint_as_float(hi[x]));
                }
                tmem_st8(th + j, hi);
                tmem_st8(th + TM_HLO + j, lo);
            }
            tc_wait_st();
            tc_fence_before();
            if (arrive) mbar_arrive(bar_hready + 8 * b);
        };

        float h[KP8], hn[KP8];
        bool staged = false;  // the first tile of the coming sweep is already in TMEM

        // the tiles of one sweep; DO_R: update (quotient written back, G2 / G3 follow); DO_KL: KL term of the incoming iterate
        auto sweep_tiles = [&](auto do_r_c, auto do_kl_c, int u) -> double {
            constexpr bool DO_R = decltype(do_r_c)::value, DO_KL = decltype(do_kl_c)::value;
            double obj_acc = 0.0;
            const int t_begin = u * n_my;
            const int i0 = (g - t_begin) & 1;  // first tile of this warpgroup in the sweep
            if (i0 < n_my) load_h(t_begin + i0, i0, h, !staged, true);
            staged = false;
            for (int i = i0; i < n_my; i += 2) {
                const int t = t_begin + i, st = t % S, b = t & 1;
                const int64_t d0 = (int64_t)(c + i * G) * TILE;
                const bool valid = d0 + s < R.D;
                mbar_wait(bar_full + 8 * st, (t / S) & 1);
                mbar_wait(bar_whfull + 8 * b, (t >> 1) & 1);
                tc_fence_after();
                const uint32_t tWH = tmem + lane_off + (b ? TM_WH1 : TM_WH0);
                const float kl = quotient_row_loop<DO_R, DO_KL>(tWH, sX + st * XSTAGE_BYTES + s * 128, s, sw);
                if (DO_KL && valid) obj_acc += (double)kl;
                if (DO_R) {
                    if (t > 0) mbar_wait(bar_shtfree, (t - 1) & 1);  // G3 of the previous tile has read sHT
                    const uint32_t tbase = sHT + (s >> 2) * SHT_LBO + (s & 3) * 4;
#pragma unroll
                    for (int j = 0; j < KP8; ++j)
                        if (j < k) sts32(tbase + (j >> 3) * SHT_SBO + (j & 7) * 16, tf32_rn(h[j]));
                    tc_wait_st();
                    fence_proxy_async();
                }
                tc_fence_before();
                mbar_arrive(bar_rready + 8 * b);

                // the next tile's exposures go to the tensor core now, so that G1(t + 2) runs behind G2(t) / G3(t)
                if (i + 2 < n_my) load_h(t + 2, i + 2, hn);

                if (DO_R) {
                    mbar_wait(bar_hnfull + 8 * b, (t >> 1) & 1);
                    tc_fence_after();
                    uint32_t v[32];
                    tmem_ld32(tmem + lane_off + (b ? TM_HN1 : TM_HN0), v);
                    tc_wait_ld();
                    const uint32_t orow = sHraw + (t % NH) * q.hraw + s * (k * 4);
                    if (GK) {
#pragma unroll
                        for (int j = 0; j < KP8; ++j)
                            if (j < k) sts32(orow + j * 4, fmaxf(h[j] * __uint_as_float(v[j]), eps));
                    } else {
#pragma unroll
                        for (int j = 0; j < KP8; j += 4)
                            if (j < k) {
                                float4 o;
                                o.x = fmaxf(h[j] * __uint_as_float(v[j]), eps);
                                o.y = fmaxf(h[j + 1] * __uint_as_float(v[j + 1]), eps);
                                o.z = fmaxf(h[j + 2] * __uint_as_float(v[j + 2]), eps);
                                o.w = fmaxf(h[j + 3] * __uint_as_float(v[j + 3]), eps);
                                sts128(orow + j * 4, o);
                            }
                    }
                    fence_proxy_async();
                    tc_fence_before();
                }
                mbar_arrive(bar_hout + 8 * (t % NH));  // slot: staged output ready (or simply no longer needed)
#pragma unroll
                for (int j = 0; j < KP8; ++j) h[j] = hn[j];
            }
            return obj_acc;
        };

        for (int u = 0; u < U; ++u) {
            const bool do_r = u < L, do_kl = sweep_kl(u);
            tl_stamp(P.tl, tl_on && tid == EPI0, u, 0);
            double obj_acc = 0.0;
            if (do_r && do_kl)
                obj_acc = sweep_tiles(std::true_type{}, std::true_type{}, u);
            else if (do_r)
                obj_acc = sweep_tiles(std::true_type{}, std::false_type{}, u);
            else
                obj_acc = sweep_tiles(std::false_type{}, std::true_type{}, u);
            // (everything below is defined here, after the tile loop: the loop has no registers to spare)
            const int ew = warp - EW0;
            const unsigned int tag = seq0 + (unsigned int)u;
            const int slot = (int)(tag & 1u);
            const bool tl = tl_on && tid == EPI0;
            const bool cs = P.tl != nullptr && vr == 0 && tid == EPI0;
            tl_stamp(P.tl, tl, u, 1);
            cta_stamp(P.tl, cs, U, G, u, c, 0);

            // ---- the numerator leaves TMEM as tagged words as soon as the last G3 has retired (G2 / E2 / the store of the last
            // tile go on meanwhile); the warpgroup that did NOT have the last tile does it: it is free earlier ----
            if (do_r && g != ((u * n_my + n_my - 1) & 1)) {
                mbar_wait(bar_wn, u & 1);
                tc_fence_after();
                if (qw < 3) {
                    uint32_t v[32];
                    tmem_ld32(tmem + lane_off + TM_WN, v);
                    tc_wait_ld();
                    uint2* dst = R.partials + ((size_t)slot * G + c) * nvals + s;  // s = feature 0 .. 95
#pragma unroll
                    for (int j = 0; j < KP8; ++j)
                        if (j < k) st_tag2(dst + j * VT, v[j], tag);
                }
            }
            tl_stamp(P.tl, tl, u, 2);
            cta_stamp(P.tl, cs, U, G, u, c, 1);
            tc_fence_before();
            if (do_kl) {  // per-CTA objective partial: warp sums -> one tagged word
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) obj_acc += __shfl_xor_sync(0xffffffffu, obj_acc, o);
                if (lane == 0) s_obj[ew] = obj_acc;
                bar_epi();
                if (tid == EPI0) {
                    double t2 = 0.0;
                    for (int w8 = 0; w8 < 8; ++w8) t2 += s_obj[w8];
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(t2);
                    st_tag4(R.obj_part + (size_t)slot * G + c, (unsigned int)bits, (unsigned int)(bits >> 32), tag);
                }
            }
            tl_stamp(P.tl, tl, u, 3);

            const size_t xrow = (size_t)slot * P.n_ranks * (size_t)(nvals + VT);  // start of recv[slot]; (k + 1) * 96 words per rank
            if (do_r) {
                // ---- stage A: this CTA's slice of the k * 96 values, summed over the G partials in a fixed order.
                // Partial b holds the slice as one contiguous run, so a warp instruction reads 32 / VW whole runs (VW lanes each):
                // G requests of ~100 bytes per CTA instead of one request per word.  Thread (warp ew, run r, value vi) adds the
                // partials b = (8 it + ew) * (32 / VW) + r, it = 0, 1, ...; a shuffle tree over r, then the eight warps in order.
                const int VPC = (nvals + G - 1) / G;
                const int v_lo = c * VPC, v_hi = v_lo + VPC < nvals ? v_lo + VPC : nvals;
                const int VW = VPC <= 4 ? 4 : VPC <= 8 ? 8 : VPC <= 16 ? 16 : 32, BPW = 32 / VW;
                const int vi = lane & (VW - 1), bsub = lane / VW, e = tid - EPI0;
                const uint2* pbase = R.partials + (size_t)slot * G * nvals;
                for (int ch0 = v_lo; ch0 < v_hi; ch0 += 32) {
                    const int nv = v_hi - ch0 < 32 ? v_hi - ch0 : 32;
                    double acc = 0.0;
                    if (vi < nv) {
                        for (int b0 = ew * BPW + bsub; b0 < G; b0 += 40 * BPW) {
                            uint2 w[5];
                            unsigned int pending = 0;
#pragma unroll
                            for (int z = 0; z < 5; ++z)
                                if (b0 + z * 8 * BPW < G) pending |= 1u << z;
                            const unsigned int have = pending;
                            poll_batch<5>(w, pending, tag, [&](int z) { return pbase + (size_t)(b0 + z * 8 * BPW) * nvals + ch0 + vi; });
#pragma unroll
                            for (int z = 0; z < 5; ++z)
                                if ((have >> z) & 1u) acc += (double)__uint_as_float(w[z].x);
                        }
                    }
                    for (int o = VW; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane < VW) s_part[ew * 32 + lane] = acc;
                    bar_epi();
                    double total = 0.0;
                    if (e < nv) {
#pragma unroll
                        for (int w8 = 0; w8 < 8; ++w8) total += s_part[w8 * 32 + e];
                    }
                    if (P.n_ranks > 1) {
                        // ---- exchange over NVLink: thread (rank rk, value) pushes the sum to peer rk and polls peer rk's word;
                        // the contributions are added in rank order, so every rank obtains the same bits
                        if (e < nv) s_tot[e] = total;
                        bar_epi();  // (also: nobody reads s_part any more, it becomes the receive staging)
                        const int rk = e >> 5, rv = e & 31;
                        if (rk < P.n_ranks && rv < nv) {
                            double val = s_tot[rv];
                            if (rk != R.rank) {
                                const unsigned long long bits = (unsigned long long)__double_as_longlong(val);
                                st_tag4(reinterpret_cast<uint4*>(R.peers[rk]) + (xrow + (size_t)R.rank * (nvals + VT) + ch0 + rv), (unsigned int)bits,
                                        (unsigned int)(bits >> 32), tag);
                                val = poll_f64(reinterpret_cast<const uint4*>(R.peers[R.rank]) + (xrow + (size_t)rk * (nvals + VT) + ch0 + rv), tag);
                            }
                            s_part[rk * 32 + rv] = val;
                        }
                        bar_epi();
                        if (e < nv) {
                            total = 0.0;
                            for (int r2 = 0; r2 < P.n_ranks; ++r2) total += s_part[r2 * 32 + e];
                        }
                    }
                    if (e < nv) st_tag2(R.sums + (size_t)slot * nvals + ch0 + e, __float_as_uint((float)total), tag);
                    if (ch0 + 32 < v_hi) bar_epi();  // the staging is reused by the next chunk
                }
                tl_stamp(P.tl, tl, u, 4);
                cta_stamp(P.tl, cs, U, G, u, c, 2);
            }

            // ---- objective: the per-CTA partials are summed (and exchanged) by the last warp of the last CTA, which has the
            // fewest tiles and the fewest values of stage A ----
            if (do_kl && c == G - 1 && ew == 7) {
                double acc = 0.0;
                for (int b0 = lane; b0 < G; b0 += 32) acc += poll_f64(R.obj_part + (size_t)slot * G + b0, tag);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (P.n_ranks > 1) {
                    const int rk = lane;
                    double v0 = acc;
                    if (rk < P.n_ranks && rk != R.rank) {
                        const unsigned long long bits = (unsigned long long)__double_as_longlong(acc);
                        st_tag4(reinterpret_cast<uint4*>(R.peers[rk]) + (xrow + (size_t)R.rank * (nvals + VT) + nvals), (unsigned int)bits,
                                (unsigned int)(bits >> 32), tag);
                        v0 = poll_f64(reinterpret_cast<const uint4*>(R.peers[R.rank]) + (xrow + (size_t)rk * (nvals + VT) + nvals), tag);
                    }
                    acc = 0.0;
                    for (int r2 = 0; r2 < P.n_ranks; ++r2) acc += __shfl_sync(0xffffffffu, v0, r2);
                }
                if (lane == 0) {
                    const int idx = u >= L ? (P.obj_every > 0 ? (L + P.obj_every - 1) / P.obj_every : 0) : u / P.obj_every;
                    R.objective[idx] = acc;
                }
            }

            if (do_r && u + 1 < U) {
                // the coming sweep's first tile of this warpgroup: exposures to TMEM now, if they have arrived (never wait here)
                const int tb = (u + 1) * n_my, i0n = (g - tb) & 1, tn = tb + i0n;
                // (warp-uniform decision: the TMEM stores inside are .sync.aligned)
                if (i0n < n_my && __all_sync(0xffffffffu, mbar_test(bar_hfull + 8 * (tn % NH), (uint32_t)((tn / NH) & 1)))) {
                    load_h(tn, i0n, h, true, false);
                    staged = true;
                }
            }
            if (do_r) {
                // ---- stage B: every CTA polls the k * 96 totals and applies the W epilogue into its own operands ----
                // W[j] <- clip(W[j] * N[j] / sum_v(W[j][v] N[j][v])), given signatures restored (reference _utils_klnmf.py:338-341;
                // double arithmetic as in klnmf_finish_kernel / w_epilogue_kernel, one warp-wide tree per signature).  All polled loads
                // of a thread are in flight together.
                uint32_t wbits[16];
                tmem_ld16(t_wst, wbits);
                uint2 wd[PER_WREG];
                const uint2* sbase = R.sums + (size_t)slot * nvals + lane;
                {
                    unsigned int pending = 0;
#pragma unroll
                    for (int r = 0; r < PER_WREG / 3; ++r)
                        if (ew + 8 * r < k) pending |= 7u << (3 * r);
                    poll_batch<PER_WREG>(wd, pending, tag, [&](int z) { return sbase + (ew + 8 * (z / 3)) * VT + (z % 3) * 32; });
                }
                mbar_wait(bar_done, u & 1);  // G2 of the last tile reads the W operands that are rewritten below (long retired by now)
                tc_fence_after();
                tc_wait_ld();
#pragma unroll
                for (int r = 0; r < PER_WREG / 3; ++r) {
                    const int j = ew + 8 * r;
                    if (j < k) {
                        double val[3];
#pragma unroll
                        for (int s3 = 0; s3 < 3; ++s3)
                            val[s3] = (double)__uint_as_float(wbits[3 * r + s3]) * (double)__uint_as_float(wd[3 * r + s3].x);
                        double tot = (val[0] + val[1]) + val[2];
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                        // val / tot through one reciprocal per signature (seed + two Newton steps) and a residual correction of each
                        // quotient: within 1 ulp of the IEEE quotient, a handful of DFMAs instead of three division sequences
                        double rcp;
                        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(rcp) : "d"(tot));
                        double err = fma(-tot, rcp, 1.0);
                        rcp = fma(rcp, err, rcp);
                        err = fma(-tot, rcp, 1.0);
                        rcp = fma(rcp, err, rcp);
#pragma unroll
                        for (int s3 = 0; s3 < 3; ++s3) {
                            const double qv = val[s3] * rcp;
                            double out = fma(fma(-tot, qv, val[s3]), rcp, qv);
                            if (j < P.n_given) out = (double)__uint_as_float(wbits[3 * r + s3]);
                            if (P.clip_given || j >= P.n_given) out = fmax(out, (double)SAL_EPS_F32);
                            wbits[3 * r + s3] = __float_as_uint((float)out);
                            store_w_operands(j, s3 * 32 + lane, (float)out);
                        }
                    }
                }
                tl_stamp(P.tl, tl, u, 5);
                cta_stamp(P.tl, cs, U, G, u, c, 3);
                tmem_st16(t_wst, wbits);
                tc_wait_st();
                fence_proxy_async();
                bar_epi();  // all W operands rewritten: the next sweep's first G1 may be issued (through hready)
                tl_stamp(P.tl, tl, u, 6);
            }
            tl_stamp(P.tl, tl, u, 7);
        }
        // ---- the final signatures go to global memory once ----
        if (c == 0 && L > 0) {
            uint32_t wbits[16];
            tmem_ld16(t_wst, wbits);
            tc_wait_ld();
#pragma unroll
            for (int r = 0; r < PER_WREG / 3; ++r) {
                const int j = (warp - EW0) + 8 * r;
                if (j < k) {
#pragma unroll
                    for (int s3 = 0; s3 < 3; ++s3) R.W_out[(size_t)j * VT + s3 * 32 + lane] = __uint_as_float(wbits[3 * r + s3]);
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (c == 0 && tid == 0) atomicMax(R.seq, seq0 + (unsigned int)U);  // every CTA read seq0 long ago (it took part in sweep 0)
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
    }
}

// ---- host side --------------------------------------------------------------------------------------------------------
template <int KP8, bool GK>
int launch_period_v(sal_ctx* const* cs, int n_virtual, const PeriodArgs* as, cudaStream_t st) {
    sal_ctx* c0 = cs[0];
    const int k = c0->k;
    const Plan q = make_plan(k, KP8, PER_EXTRA);
    if (q.total > SMEM_LIMIT) {
        sal_set_error("period kernel: shared-memory plan of %d bytes exceeds the limit", q.total);
        return SAL_EUNSUPPORTED;
    }
    static bool attr_set[16] = {false};
    if (!attr_set[c0->device & 15]) {
        SAL_CUDA(cudaFuncSetAttribute(klnmf_period_tc_kernel<KP8, GK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_set[c0->device & 15] = true;
    }
    PeriodParams P;
    memset(&P, 0, sizeof(P));
    // one grid for all virtual ranks: G CTAs each, every CTA with at least one tile
    int G = c0->n_sm / n_virtual;
    for (int v = 0; v < n_virtual; ++v) {
        const int n_tiles = (int)((cs[v]->D + TILE - 1) / TILE);
        if (n_tiles < G) G = n_tiles;
    }
    const PeriodArgs& a0 = as[0];
    P.n_virtual = n_virtual, P.G = G, P.k = k, P.n_updates = a0.n_updates, P.obj_every = a0.obj_every, P.final_obj = a0.final_obj;
    P.n_given = a0.n_given, P.clip_given = a0.clip_given, P.n_ranks = a0.n_ranks;
    P.tl = (unsigned long long*)c0->dbg;
    for (int v = 0; v < n_virtual; ++v) {
        sal_ctx* c = cs[v];
        const PeriodArgs& a = as[v];
        PeriodRank& R = P.r[v];
        if (int err = sal_period_scratch(c, st)) return err;
        if (int err = sal_cached_map(c, &R.mapX, a.X, 0)) return err;
        if (int err = sal_cached_map(c, &R.mapH0, a.H_in, GK ? 2 : 1)) return err;
        if (int err = sal_cached_map(c, &R.mapH1, a.H_out, GK ? 2 : 1)) return err;
        R.W_in = (const float*)a.W_in, R.W_out = (float*)a.W_out, R.H_in = (const float*)a.H_in, R.H_out = (float*)a.H_out;
        R.partials = (uint2*)c->period_partials, R.sums = (uint2*)c->period_sums, R.obj_part = (uint4*)c->period_obj;
        R.objective = a.objectives;
        R.peers = (void* const*)a.peers, R.seq = a.state ? (unsigned int*)a.state : c->period_seq;
        R.D = c->D, R.n_tiles = (int)((c->D + TILE - 1) / TILE), R.rank = a.rank;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(G * n_virtual)), cfg.blockDim = dim3(PER_THREADS), cfg.dynamicSmemBytes = (size_t)q.total, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // co-residency of all CTAs is part of the protocol: they wait on one another
    attr[0].val.cooperative = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    SAL_CUDA(cudaLaunchKernelEx(&cfg, klnmf_period_tc_kernel<KP8, GK>, P));
    for (int v = 0; v < n_virtual; ++v) cs[v]->launches++;
    return 0;
}

}  // namespace

int sal_cached_map(sal_ctx* c, void* map_out, const void* ptr, int kind) {
    if (!c->map_cache) c->map_cache = new std::vector<sal_map_entry>();
    for (const sal_map_entry& m : *c->map_cache)
        if (m.ptr == ptr && m.kind == kind) {
            memcpy(map_out, m.map, sizeof(CUtensorMap));
            return 0;
        }
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    sal_map_entry m;
    m.ptr = ptr, m.kind = kind;
    CUtensorMap* t = reinterpret_cast<CUtensorMap*>(m.map);
    int err = 0;
    if (kind == 0)
        err = encode_2d(t, ptr, VT, (uint64_t)c->D, 32, TILE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    else if (kind == 1)
        err = encode_2d(t, ptr, (uint64_t)c->k, (uint64_t)c->D, (uint32_t)c->k, TILE, CU_TENSOR_MAP_SWIZZLE_NONE);
    else
        err = encode_h3d(t, ptr, c->k, c->D);
    if (err) return err;
    if (c->map_cache->size() >= 32) c->map_cache->clear();  // buffers come and go with the fits: start over
    c->map_cache->push_back(m);
    memcpy(map_out, m.map, sizeof(CUtensorMap));
    return 0;
}

bool sal_period_supported(const sal_ctx* c, const PeriodArgs& a) {
    if (c->dtype != SAL_F32 || c->V != VT || c->k > 32 || c->math == SAL_MATH_FMA) return false;
    if (((uintptr_t)a.X | (uintptr_t)a.H_in | (uintptr_t)a.H_out) & 15) return false;
    if (c->D < 1 || c->D >= (int64_t)1 << 31) return false;
    if (c->math != SAL_MATH_TF32_ALWAYS && c->D < SAL_TF32_MIN_SAMPLES) return false;
    if (a.n_given >= c->k && a.n_updates > 0) return false;  // nothing to reduce: the plain pass does that case
    if (a.n_ranks > PER_MAX_RANKS) return false;
    return true;
}

int sal_launch_period(sal_ctx* const* cs, int n_virtual, const PeriodArgs* as, cudaStream_t st) {
    if (n_virtual < 1 || n_virtual > PER_MAX_VIRTUAL) {
        sal_set_error("period kernel: %d virtual ranks (1 .. %d supported)", n_virtual, PER_MAX_VIRTUAL);
        return SAL_EINVAL;
    }
    const bool gk = (cs[0]->k & 3) != 0;
    switch (cs[0]->KP) {
        case 8: return gk ? launch_period_v<8, true>(cs, n_virtual, as, st) : launch_period_v<8, false>(cs, n_virtual, as, st);
        case 16: return gk ? launch_period_v<16, true>(cs, n_virtual, as, st) : launch_period_v<16, false>(cs, n_virtual, as, st);
        case 24: return gk ? launch_period_v<24, true>(cs, n_virtual, as, st) : launch_period_v<24, false>(cs, n_virtual, as, st);
        default: return gk ? launch_period_v<32, true>(cs, n_virtual, as, st) : launch_period_v<32, false>(cs, n_virtual, as, st);
    }
}
