// Fused KL-NMF pass, tensor-core flavour (tcgen05 kind::tf32, fp32 accumulation in TMEM).
//
// Same contract as the FMA pass (klnmf_pass.cu; reference update_WH / update_H / update_W /
// kl_divergence, models/_utils_klnmf.py:11-55, 164-361) for fp32 handles, V = 96, any k <= 32, no
// per-sample weights.  The three thin contractions of one 128-sample tile run on the 5th-gen tensor
// cores; X is streamed from HBM exactly once by TMA and the quotient never leaves the SM:
//
//   G1  WH[s,f]   = sum_j H[s,j] W[j,f]        A = H   (TMEM, "TS" form)    B = sW1 (smem, K-major)
//   E1  R[s,f]    = X[s,f] / WH[s,f]           thread s = TMEM lane s; R overwrites X in smem AND WH in TMEM
//   G2  Hn[s,j]   = sum_f R[s,f] W[j,f]        A = R   (TMEM, "TS" form)    B = sW2 (smem, K-major)
//   G3  Wn[f,j]  += sum_s R[s,f] H[s,j]        A = R   (smem, MN-major)     B = sHT (smem, K-major)
//   E2  H_out[s,j] = max(H[s,j] Hn[s,j], eps)
//
// tf32 recipe (scripts/tf32_precision_experiment.py): WH feeds a division and decides whether the
// fit's 1e-7 convergence test sees noise, so G1 is error compensated -- h = hi + lo, w = hi + lo,
// WH = hi*hi + lo*hi + hi*lo (3 x tf32, ~2^-22).  The quotient R and the operands of G2 / G3 are plain tf32
// (round to nearest): their rounding noise averages out over the 96 features / all samples.
//
// Wn (96 x k) accumulates in TMEM over all tiles of the CTA and is written once as a per-CTA partial;
// the deterministic fixed-order reduction kernel of klnmf_pass.cu finishes the job.
//
// Shared-memory operand layouts
//   X / R stage : 3 TMA boxes [128 samples][32 features] fp32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  The very same
//                 bytes are the MN-major SWIZZLE_128B_BASE32B A operand of G3 (M = feature, K = sample) -- the only
//                 MN-major layout tcgen05 accepts for 32-bit operands -- so R is written in place over X.
//   sW1, sW2, sHT : canonical no-swizzle K-major "core matrices" (8 rows x 16 B, 128 B contiguous),
//                 written by the threads with round-to-nearest tf32 conversion.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over tiles):  warp 0 TMA producer, warp 1 MMA issuer,
// warp 2 TMEM allocator, warp 3 TMA store of the updated exposures (staged in the tile's raw-H slot),
// warps 4-7 and 8-11 two epilogue warpgroups that alternate tiles.
#include "tc_common.cuh"

namespace {

// GK: k % 4 != 0 (3-D exposure view, scalar staging accesses); HS: MvNMF trial pass (exposures rescaled on the way in and
// written back).  Both are compile-time: as run-time branches in the per-tile loops they cost the common variant 10 %.
// WT: per-sample weights (weights_kl / weights_lhalf, reference _utils_klnmf.py:333-360, klnmf.py:75-79): the KL weight scales
// the sample's row of the H^T operand of G3 and its objective term, the l-half weight switches the H update to its
// closed form and adds lambda_d * sum_k sqrt(h_dk) to the objective.  Per-row work only; the quotient loop is untouched.
template <int KP8, bool DO_R, bool DO_KL, bool GK, bool HS, bool WT>
__global__ void __launch_bounds__(NTHREADS, 1)
klnmf_pass_tc_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapH,
                     const __grid_constant__ CUtensorMap mapHout, TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_dyn[];
    const uint32_t base = smem_u32(smem_dyn);
    if (base & 1023u) __trap();  // the swizzled TMA boxes need 1024-byte alignment
    const int k = p.k;
    const Plan q = make_plan(k, KP8);
    const int S = q.S;
    const uint32_t sX = base, sHraw = base + q.off_hraw, sW1hi = base + q.off_w1hi, sW1lo = base + q.off_w1lo;
    const uint32_t sW2 = base + q.off_w2, sHT = base + q.off_sht, bars = base + q.off_bar;
    const int SHT_LBO = q.sht_lbo;
    // barriers: full[3] empty[3] hready[2] whfull[2] rready[2] hnfull[2] shtfree done hfull[4] hempty[4] hout[4]
    const uint32_t bar_full = bars, bar_empty = bars + 24, bar_hready = bars + 48, bar_whfull = bars + 64;
    const uint32_t bar_rready = bars + 80, bar_hnfull = bars + 96, bar_shtfree = bars + 112, bar_done = bars + 120;
    const uint32_t bar_hfull = bars + 128, bar_hempty = bars + 160, bar_hout = bars + 192;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_dyn + q.off_misc);
    double* s_red = reinterpret_cast<double*>(smem_dyn + q.off_misc + 16);  // [8]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool do_h = p.flags & SAL_PASS_UPDATE_H, do_w = p.flags & SAL_PASS_WNUM;
    const int n_my = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles of this CTA (>= 1)

    // ---- one-time setup --------------------------------------------------------------------------
    // (programmatic dependent launch: everything down to pdl_wait() is independent of the previous kernels and
    // overlaps the reduction kernel of the previous update)
    pdl_trigger();
    stamp(p.dbg, p.dbg != nullptr && blockIdx.x == 1 && tid == 0, 3, 0, 1);
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapH) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapHout) : "memory");
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
        for (int i = 0; i < NH; ++i) mbar_init(bar_hfull + 8 * i, 1), mbar_init(bar_hempty + 8 * i, 1), mbar_init(bar_hout + 8 * i, 128);
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(TM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (DO_R)
        for (int i = tid; i < q.sht / 4; i += NTHREADS) sts32(sHT + 4 * i, 0.f);
    __syncthreads();  // barriers initialised
    // X never changes during a fit: the first S tiles of it are requested before the dependency wait, so the pipeline is
    // already full when the previous update's reduction kernel retires
    const int n_pre = n_my < S ? n_my : S;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < n_pre; ++i) {
            const int d0 = ((int)blockIdx.x + i * (int)gridDim.x) * TILE;
            mbar_arrive_expect_tx(bar_full + 8 * i, XSTAGE_BYTES);
            for (int c = 0; c < NBOX; ++c) tma_load_2d(sX + i * XSTAGE_BYTES + c * BOX_BYTES, &mapX, bar_full + 8 * i, c * 32, d0);
        }
    }
    // W operands as tf32 hi + lo (round to nearest), zero padding for signatures >= k.
    // KP8 * 96 = (KP8 / 4) * 384 elements: every thread first issues all its global loads, then converts and stores.
    pdl_wait();  // W (previous reduction kernel) and H (previous pass) are valid from here on
    {
        float wreg[KP8 / 4];
#pragma unroll
        for (int it = 0; it < KP8 / 4; ++it) {
            const int i = tid + it * NTHREADS, j = i / VT, f = i - j * VT;
            wreg[it] = j < k ? __ldcg(p.W + (size_t)j * VT + f) : 0.f;  // L2 only: W was written while this grid was resident
        }
#pragma unroll
        for (int it = 0; it < KP8 / 4; ++it) {
            const int i = tid + it * NTHREADS, j = i / VT, f = i - j * VT;
            const float w = wreg[it];
            const float hi = tf32_rn(w), lo = tf32_rn(w - hi);
            const uint32_t o1 = (j >> 2) * SW1_LBO + (f >> 3) * SW1_SBO + (f & 7) * 16 + (j & 3) * 4;
            sts32(sW1hi + o1, hi), sts32(sW1lo + o1, lo);
            if (DO_R) sts32(sW2 + (f >> 2) * q.sw2_lbo + (j >> 3) * SW2_SBO + (j & 7) * 16 + (f & 3) * 4, hi);
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    stamp(p.dbg, p.dbg != nullptr && blockIdx.x == 1 && tid == 0, 3, 0, 2);

    double obj_acc = 0.0;
    const bool tl = p.dbg != nullptr && blockIdx.x == 1;  // diagnostics timeline (CTA 1: CTA 0 is busy dumping its first tile)

    if (warp == 0) {
        // ================= TMA producer =================
        // Lane 0 issues the bulk copies.  The last, partial tile of a generic-k problem cannot go through the 3-D map
        // (it would describe memory past the end of H): the whole warp copies its rows with plain loads and zero-fills
        // the rest of the slot, then lane 0 completes the slot's barrier phase with a plain arrive.
        if (!GK) {
            if (lane == 0) {
                for (int i = 0; i < n_my; ++i) {
                    const int st = i % S, hs = i % NH;
                    const int d0 = ((int)blockIdx.x + i * (int)gridDim.x) * TILE;
                    mbar_wait(bar_hempty + 8 * hs, ((i / NH) & 1) ^ 1);
                    mbar_arrive_expect_tx(bar_hfull + 8 * hs, (uint32_t)(TILE * k * 4));
                    tma_load_2d(sHraw + hs * q.hraw, &mapH, bar_hfull + 8 * hs, 0, d0);
                    if (i < n_pre) continue;  // X tile requested in the prologue
                    mbar_wait(bar_empty + 8 * st, ((i / S) & 1) ^ 1);
                    stamp(p.dbg, tl, 3, i, 0);
                    mbar_arrive_expect_tx(bar_full + 8 * st, XSTAGE_BYTES);
                    for (int c = 0; c < NBOX; ++c) tma_load_2d(sX + st * XSTAGE_BYTES + c * BOX_BYTES, &mapX, bar_full + 8 * st, c * 32, d0);
                }
            }
        } else {
            constexpr bool gk = GK;
            for (int i = 0; i < n_my; ++i) {
                const int st = i % S, hs = i % NH;
                const int tile = (int)blockIdx.x + i * (int)gridDim.x;
                const int d0 = tile * TILE;
                const bool ragged = gk && (int64_t)d0 + TILE > p.D;  // warp-uniform
                if (lane == 0) {
                    mbar_wait(bar_hempty + 8 * hs, ((i / NH) & 1) ^ 1);
                    if (!ragged) {
                        mbar_arrive_expect_tx(bar_hfull + 8 * hs, (uint32_t)(TILE * k * 4));
                        if (gk)
                            tma_load_3d(sHraw + hs * q.hraw, &mapH, bar_hfull + 8 * hs, 0, 0, tile);
                        else
                            tma_load_2d(sHraw + hs * q.hraw, &mapH, bar_hfull + 8 * hs, 0, d0);
                    }
                }
                if (ragged) {
                    __syncwarp();
                    const int n = (int)(p.D - d0) * k;
                    const float* src = p.H_in + (size_t)d0 * k;
                    for (int e = lane; e < TILE * k; e += 32) sts32(sHraw + hs * q.hraw + e * 4, e < n ? __ldcg(src + e) : 0.f);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_hfull + 8 * hs);
                }
                if (lane == 0 && i >= n_pre) {  // the first n_pre X tiles were requested in the prologue
                    mbar_wait(bar_empty + 8 * st, ((i / S) & 1) ^ 1);
                    stamp(p.dbg, tl, 3, i, 0);
                    mbar_arrive_expect_tx(bar_full + 8 * st, XSTAGE_BYTES);
                    for (int c = 0; c < NBOX; ++c) tma_load_2d(sX + st * XSTAGE_BYTES + c * BOX_BYTES, &mapX, bar_full + 8 * st, c * 32, d0);
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        // The whole warp runs the loop (descriptor arithmetic stays in uniform registers); one elected lane issues.
        constexpr uint32_t ID1 = make_idesc(128, VT, 0, 0);
        constexpr uint32_t ID2 = make_idesc(128, N2, 0, 0);
        constexpr uint32_t ID3 = make_idesc(128, N2, 1, 0);
        const uint64_t dW1hi = make_desc(sW1hi, SW1_LBO, SW1_SBO, LAYOUT_NONE), dW1lo = make_desc(sW1lo, SW1_LBO, SW1_SBO, LAYOUT_NONE);
        const uint64_t dW2 = make_desc(sW2, q.sw2_lbo, SW2_SBO, LAYOUT_NONE);
        const uint64_t dHT = make_desc(sHT, SHT_LBO, SHT_SBO, LAYOUT_NONE);
        const uint32_t htstep = (uint32_t)(2 * SHT_LBO) >> 4;
        const uint64_t dX0 = make_desc(sX, BOX_BYTES, 512, LAYOUT_128B_BASE32B);
        const uint32_t w2step = (uint32_t)(2 * q.sw2_lbo) >> 4;
        auto issue_g1 = [&](int i) {
            const int b = i & 1;
            mbar_wait(bar_hready + 8 * b, (i >> 1) & 1);
            stamp(p.dbg, tl && lane == 0, 2, i, 0);
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
            stamp(p.dbg, tl && lane == 0, 2, i, 2);
        };
        issue_g1(0);
        for (int i = 0; i < n_my; ++i) {
            if (i + 1 < n_my) issue_g1(i + 1);
            const int st = i % S, b = i & 1;
            mbar_wait(bar_rready + 8 * b, (i >> 1) & 1);
            stamp(p.dbg, tl && lane == 0, 2, i, 1);
            tc_fence_after();
            if (DO_R) {
                const uint32_t tR = tmem + (b ? TM_WH1 : TM_WH0), tHn = tmem + (b ? TM_HN1 : TM_HN0);
                const uint64_t dX = dX0 + (uint64_t)((uint32_t)(st * XSTAGE_BYTES) >> 4);
                if (elect_one()) {
                    // G3 first: it is the last reader of the X / R stage and of sHT, so the stage goes back to the
                    // TMA producer as early as possible; G2 (TMEM operand) follows.
                    if (do_w) {
#pragma unroll
                        for (int ks = 0; ks < TILE / 8; ++ks)
                            mma_ss(tmem + TM_WN, dX + (uint64_t)(ks * (1024 >> 4)), dHT + (uint64_t)(ks * htstep), ID3,
                                   (i > 0 || ks > 0));
                    }
                    tc_commit(bar_empty + 8 * st);
                    tc_commit(bar_shtfree);
                    if (do_h) {
#pragma unroll
                        for (int ks = 0; ks < VT / 8; ++ks) mma_ts(tHn, tR + ks * 8, dW2 + (uint64_t)(ks * w2step), ID2, ks > 0);
                        tc_commit(bar_hnfull + 8 * b);
                    }
                }
                __syncwarp();
                stamp(p.dbg, tl && lane == 0, 2, i, 3);
            } else if (lane == 0) {  // objective only: nothing reads the stage after E1
                mbar_arrive(bar_empty + 8 * st);
                mbar_arrive(bar_shtfree);
            }
        }
        if (elect_one()) tc_commit(bar_done);
        __syncwarp();
        mbar_wait(bar_done, 0);  // every MMA has retired before the CTA tears TMEM down
    } else if (warp == 3) {
        // ================= exposure store =================
        // The epilogue leaves the updated exposures of tile i in raw-H slot i % NH; one thread streams them out with
        // a TMA store (rows beyond D are clipped by the 2-D tensor map; the partial last tile of a generic-k problem is
        // written by the whole warp with plain stores) and hands the slot back to the producer.
        if (!GK) {
            if (lane == 0) {
                for (int i = 0; i < n_my; ++i) {
                    const int hs = i % NH;
                    mbar_wait(bar_hout + 8 * hs, (i / NH) & 1);
                    if (do_h) {
                        tma_store_2d(&mapHout, sHraw + hs * q.hraw, 0, ((int)blockIdx.x + i * (int)gridDim.x) * TILE);
                        tma_store_commit_and_wait_read();
                    }
                    mbar_arrive(bar_hempty + 8 * hs);
                }
                tma_store_wait_all();
            }
        } else {
            constexpr bool gk = GK;
            for (int i = 0; i < n_my; ++i) {
                const int hs = i % NH;
                const int tile = (int)blockIdx.x + i * (int)gridDim.x;
                const int d0 = tile * TILE;
                const bool ragged = gk && (int64_t)d0 + TILE > p.D;  // warp-uniform
                if (lane == 0) mbar_wait(bar_hout + 8 * hs, (i / NH) & 1);
                if (do_h) {
                    if (ragged) {
                        __syncwarp();
                        const int n = (int)(p.D - d0) * k;
                        float* dst = p.H_out + (size_t)d0 * k;
                        for (int e = lane; e < n; e += 32) dst[e] = lds32(sHraw + hs * q.hraw + e * 4);
                        __syncwarp();
                    } else if (lane == 0) {
                        if (gk)
                            tma_store_3d(&mapHout, sHraw + hs * q.hraw, 0, 0, tile);
                        else
                            tma_store_2d(&mapHout, sHraw + hs * q.hraw, 0, d0);
                        tma_store_commit_and_wait_read();
                    }
                }
                if (lane == 0) mbar_arrive(bar_hempty + 8 * hs);
            }
            if (lane == 0) tma_store_wait_all();
        }
    } else if (warp >= 4) {
        // ================= epilogue warpgroups =================
        const int g = (warp - 4) >> 2, qw = warp & 3;
        const int s = qw * 32 + lane;  // sample row of the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)(qw * 32) << 16;
        const uint32_t sw = (s >> 2) & 1;
        const float eps = (float)SAL_EPS_F32;
        // P0: exposures of sample s of tile i (rows beyond D are zero-filled by TMA) -> registers and, as tf32
        // hi / lo, the TMEM A operand of G1.  TMEM buffer (i & 1) was last read by G1(i - 2), which retired before
        // this warpgroup began E1(i - 2).
        auto load_h = [&](int i, float (&h)[KP8]) {
            const int hs = i % NH, b = i & 1;
            mbar_wait(bar_hfull + 8 * hs, (i / NH) & 1);
            const uint32_t hrow = sHraw + hs * q.hraw + s * (k * 4);
            const uint32_t th = tmem + lane_off + (b ? TM_H1 : TM_H0);
            // rows past the end of X (zero-filled by TMA) are read as 1: WH > 0, the zero counts give a zero quotient
            const bool row_valid = (int64_t)((int)blockIdx.x + i * (int)gridDim.x) * TILE + s < p.D;
#pragma unroll
            for (int j = 0; j < KP8; j += 8) {
                if (GK) {  // rows are not 16-byte aligned: scalar loads
#pragma unroll
                    for (int e = 0; e < 8; ++e) h[j + e] = j + e < k ? lds32(hrow + (j + e) * 4) : 0.f;
                } else {
                    float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
                    if (j < k) t0 = lds128(hrow + j * 4);
                    if (j + 4 < k) t1 = lds128(hrow + j * 4 + 16);
                    h[j] = t0.x, h[j + 1] = t0.y, h[j + 2] = t0.z, h[j + 3] = t0.w;
                    h[j + 4] = t1.x, h[j + 5] = t1.y, h[j + 6] = t1.z, h[j + 7] = t1.w;
                }
                if (HS) {  // normalize_WH folded into the read (reference utils.py:155-158 as used by mvnmf.py:80-88)
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (j + e < k) h[j + e] = fmaxf(h[j + e] * __ldg(p.h_scale + j + e), eps);
                }
                if (!row_valid) {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        if (j + e < k) h[j + e] = 1.f;
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    hi[e] = tf32_bits(h[j + e]);
                    lo[e] = tf32_bits(h[j + e] - __uint_as_float(hi[e]));
                }
                tmem_st8(th + j, hi);
                tmem_st8(th + TM_HLO + j, lo);
            }
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(bar_hready + 8 * b);
        };

        float h[KP8], hn[KP8];
        if (g < n_my) load_h(g, h);
        for (int i = g; i < n_my; i += 2) {
            const int st = i % S, b = i & 1;
            const int64_t d0 = (int64_t)((int)blockIdx.x + i * (int)gridDim.x) * TILE;
            const bool valid = d0 + s < p.D;
            const bool tlw = tl && lane == 0 && qw == 0;
            stamp(p.dbg, tlw, g, i, 0);
            mbar_wait(bar_full + 8 * st, (i / S) & 1);
            mbar_wait(bar_whfull + 8 * b, (i >> 1) & 1);
            stamp(p.dbg, tlw, g, i, 3);
            tc_fence_after();
            const uint32_t tWH = tmem + lane_off + (b ? TM_WH1 : TM_WH0);
            // the quotient as a six-step loop (instruction-cache resident; see quotient_row_loop).  Rows past the end of X need
            // no special case: their exposures were read as 1 (load_h), so WH > 0 and the zero-filled X gives a zero quotient
            const float kl = quotient_row_loop<DO_R, DO_KL>(tWH, sX + st * XSTAGE_BYTES + s * 128, s, sw,
                                                            (p.dbg && blockIdx.x == 0 && i == 0) ? p.dbg + (size_t)s * VT : nullptr);
            float wk = 1.f, lam = 0.f;
            if (WT && valid) {
                if (p.w_kl) wk = __ldg(p.w_kl + d0 + s);
                if (p.w_lhalf) lam = __ldg(p.w_lhalf + d0 + s);
            }
            if (DO_KL && valid) {
                if (WT) {
                    float sq = 0.f;
                    if (p.w_lhalf) {
#pragma unroll
                        for (int j = 0; j < KP8; ++j)
                            if (j < k) sq += sqrtf(h[j]);
                    }
                    obj_acc += (double)kl * (double)wk + (double)lam * (double)sq;
                } else {
                    obj_acc += (double)kl;
                }
            }
            stamp(p.dbg, tlw, g, i, 4);
            if (DO_R) {
                if (i > 0) mbar_wait(bar_shtfree, (i - 1) & 1);  // G3(i-1) has read sHT
                stamp(p.dbg, tlw, g, i, 5);
                const uint32_t tbase = sHT + (s >> 2) * SHT_LBO + (s & 3) * 4;
#pragma unroll
                for (int j = 0; j < KP8; ++j)
                    if (j < k) sts32(tbase + (j >> 3) * SHT_SBO + (j & 7) * 16, tf32_rn(WT ? h[j] * wk : h[j]));
                tc_wait_st();
                stamp(p.dbg, tlw, g, i, 6);
                fence_proxy_async();
            }
            tc_fence_before();
            mbar_arrive(bar_rready + 8 * b);
            stamp(p.dbg, tlw, g, i, 1);

            // the next tile's exposures go to the tensor core now, so that G1(i + 2) runs behind G2(i) / G3(i)
            // while this warpgroup finishes tile i
            if (i + 2 < n_my) load_h(i + 2, hn);
            stamp(p.dbg, tlw, g, i, 2);

            if (DO_R && do_h) {
                mbar_wait(bar_hnfull + 8 * b, (i >> 1) & 1);
                stamp(p.dbg, tlw, g, i, 7);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(tmem + lane_off + (b ? TM_HN1 : TM_HN0), v);
                tc_wait_ld();
                if (p.dbg && blockIdx.x == 0 && i == 0) {
#pragma unroll
                    for (int e = 0; e < 32; ++e) p.dbg[(size_t)TILE * VT + s * 32 + e] = __uint_as_float(v[e]);
                }
                {
                    const uint32_t orow = sHraw + (i % NH) * q.hraw + s * (k * 4);
                    if (WT && p.w_lhalf) {
                        // closed form of the l-half penalised update (reference _utils_klnmf.py:349-360); the result takes the
                        // place of the plain product below by rewriting v[] so that h * v equals it
                        const float wsq = p.w_kl ? wk * wk : 1.f;
#pragma unroll
                        for (int j = 0; j < KP8; ++j)
                            if (j < k) {
                                float t = 4.f * h[j] * __uint_as_float(v[j]);
                                if (p.w_kl) t *= wsq;
                                const float root = 0.5f * lam - sqrtf(0.25f * lam * lam + t);
                                float o = 0.25f * root * root;
                                if (p.w_kl) o = o / wsq;
                                h[j] = fmaxf(o, eps);
                                v[j] = __float_as_uint(1.f);
                            }
                    }
                    if (GK) {
#pragma unroll
                        for (int j = 0; j < KP8; ++j)
                            if (j < k) sts32(orow + j * 4, fmaxf(h[j] * __uint_as_float(v[j]), eps));
                    } else
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
                    fence_proxy_async();
                }
                tc_fence_before();
            }
            if (HS && !DO_R && do_h) {  // rescale mode: the exposures as they were read (scaled, clipped) are the output
                const uint32_t orow = sHraw + (i % NH) * q.hraw + s * (k * 4);
#pragma unroll
                for (int j = 0; j < KP8; ++j)
                    if (j < k) sts32(orow + j * 4, h[j]);
                fence_proxy_async();
            }
            mbar_arrive(bar_hout + 8 * (i % NH));  // slot i % NH: staged output ready (or simply no longer needed)
#pragma unroll
            for (int j = 0; j < KP8; ++j) h[j] = hn[j];
        }
        // ---- per-CTA numerator partial: TMEM lane = feature, column = signature
        if (DO_R && do_w && g == 0) {
            mbar_wait(bar_done, 0);
            tc_fence_after();
            if (qw < 3) {
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
        if (DO_KL) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) obj_acc += __shfl_xor_sync(0xffffffffu, obj_acc, o);
            if (lane == 0) s_red[warp - 4] = obj_acc;
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (DO_KL && tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        p.partial_obj[blockIdx.x] = t;
    }
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
    }
    stamp(p.dbg, p.dbg != nullptr && blockIdx.x == 1 && tid == 0, 3, 0, 3);
}

// Column sums of H [D][k] (SAL_PASS_HSUM, MvNMF's rowsums_H): block b owns a contiguous run of rows and reads it as a flat
// array; HSUM_THREADS is rounded down to a multiple of k by masking, so a thread always meets the same column.  The
// partials land in partial_hsum[b][SAL_KMAX]; the reduction kernel adds the blocks in order.
constexpr int HSUM_THREADS = 1024;
__global__ void __launch_bounds__(HSUM_THREADS) hsum_partials_kernel(const float* H, int64_t D, int k, double* partial_hsum) {
    __shared__ double s_acc[HSUM_THREADS];
    const int tid = threadIdx.x;
    const int active = (HSUM_THREADS / k) * k;
    const int64_t rows_per = (D + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per, r1 = r0 + rows_per < D ? r0 + rows_per : D;
    double acc = 0.0;
    if (tid < active && r0 < r1) {
        const float* base = H + r0 * k;
        const int64_t n = (r1 - r0) * k, step = active;
        float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int64_t e = tid;
        int rounds = 0;
        for (; e + 7 * step < n; e += 8 * step) {  // eight independent loads in flight per thread
#pragma unroll
            for (int z = 0; z < 8; ++z) a[z] += __ldg(base + e + z * step);
            if (++rounds == 16) {  // keep the fp32 runs short
                acc += (double)(((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7])));
#pragma unroll
                for (int z = 0; z < 8; ++z) a[z] = 0.f;
                rounds = 0;
            }
        }
        for (; e < n; e += step) a[0] += __ldg(base + e);
        acc += (double)(((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7])));
    }
    s_acc[tid] = acc;
    __syncthreads();
    if (tid < SAL_KMAX) {
        double t = 0.0;
        if (tid < k)
            for (int u = tid; u < active; u += k) t += s_acc[u];
        partial_hsum[(size_t)blockIdx.x * SAL_KMAX + tid] = t;
    }
}

template <int KP8, bool DO_R, bool DO_KL, bool GK, bool HS, bool WT>
int launch_tc_v(sal_ctx* c, const PassArgs& a, cudaStream_t st) {
    const Plan q = make_plan(c->k, KP8);
    if (q.total > SMEM_LIMIT) {
        sal_set_error("tensor-core pass: shared-memory plan of %d bytes exceeds the limit", q.total);
        return SAL_EUNSUPPORTED;
    }
    static bool attr_set[16] = {false};  // (one flag array per instantiation)
    if (!attr_set[c->device & 15]) {
        SAL_CUDA(cudaFuncSetAttribute(klnmf_pass_tc_kernel<KP8, DO_R, DO_KL, GK, HS, WT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SMEM_LIMIT));
        attr_set[c->device & 15] = true;
    }
    // tensor maps come from the handle's cache: encoding three of them per launch was ~30 us of host time per update
    CUtensorMap mapX, mapH, mapHout;
    const void* hout = (a.flags & SAL_PASS_UPDATE_H) ? a.H_out : a.H_in;
    constexpr bool generic_k = GK;
    if (int e = sal_cached_map(c, &mapX, a.X, 0)) return e;
    if (int e = sal_cached_map(c, &mapH, a.H_in, generic_k ? 2 : 1)) return e;
    if (int e = sal_cached_map(c, &mapHout, hout, generic_k ? 2 : 1)) return e;
    TcParams p;
    p.W = (const float*)a.W, p.H_in = (const float*)a.H_in, p.H_out = (float*)hout, p.generic_k = generic_k ? 1 : 0;
    p.h_scale = (const float*)a.h_scale, p.w_kl = (const float*)a.w_kl, p.w_lhalf = (const float*)a.w_lhalf;
    p.partial_wnum = (float*)c->partial_wnum, p.partial_obj = c->partial_obj;
    p.dbg = (float*)c->dbg;
    p.D = c->D, p.k = c->k, p.flags = a.flags;
    p.n_tiles = (int)((c->D + TILE - 1) / TILE);
    const int grid = p.n_tiles < c->n_sm ? p.n_tiles : c->n_sm;
    if (int e = sal_timing_begin(c, a.flags, st)) return e;
    // The kernel requests its first X tiles BEFORE the dependency wait (X never changes during a fit).  If a kernel of this
    // handle has just written X (sal_clip_counts) that early read would race with it: this one launch is then made without the
    // programmatic-overlap attribute, i.e. in plain stream order.  (X written by anybody else right before a pass remains the
    // caller's business: synchronise the stream or call sal_mark_counts_written.)
    const bool overlap = !c->x_dirty;
    c->x_dirty = 0;
    SAL_CUDA(sal_launch_pdl_if(overlap, klnmf_pass_tc_kernel<KP8, DO_R, DO_KL, GK, HS, WT>, grid, NTHREADS, q.total, st, mapX, mapH, mapHout, p));
    if (int e = sal_timing_end(c, a.flags, st)) return e;
    c->launches++;
    if (a.flags & SAL_PASS_HSUM) {  // column sums of H_in: per-block partials in the layout the reduction kernel expects
        hsum_partials_kernel<<<grid, HSUM_THREADS, 0, st>>>((const float*)a.H_in, c->D, c->k, c->partial_hsum);
        SAL_CUDA(cudaGetLastError());
        c->launches++;
    }
    return sal_launch_pass_reduce(c, a, grid, st);
}

template <int KP8>
int launch_tc(sal_ctx* c, const PassArgs& a, cudaStream_t st) {
    // with h_scale, UPDATE_H means "write the rescaled exposures" (no multiplicative update): the objective-only pipeline
    const bool r = !a.h_scale && (a.flags & (SAL_PASS_UPDATE_H | SAL_PASS_WNUM)), kl = a.flags & SAL_PASS_OBJECTIVE;
    const bool gk = (c->k & 3) != 0, wt = a.w_kl || a.w_lhalf;
#define SAL_TC(R_, KL_, HS_)                                                                               \
    (gk ? (wt ? launch_tc_v<KP8, R_, KL_, true, HS_, true>(c, a, st) : launch_tc_v<KP8, R_, KL_, true, HS_, false>(c, a, st)) \
        : (wt ? launch_tc_v<KP8, R_, KL_, false, HS_, true>(c, a, st) : launch_tc_v<KP8, R_, KL_, false, HS_, false>(c, a, st)))
    if (a.h_scale && (a.flags & SAL_PASS_SCALED_UPDATE))  // the trial's objective + the next H step from the rescaled exposures
        return gk ? launch_tc_v<KP8, true, true, true, true, false>(c, a, st) : launch_tc_v<KP8, true, true, false, true, false>(c, a, st);
    if (a.h_scale) return gk ? launch_tc_v<KP8, false, true, true, true, false>(c, a, st) : launch_tc_v<KP8, false, true, false, true, false>(c, a, st);
    if (r && !kl) return SAL_TC(true, false, false);
    if (!r && kl) return SAL_TC(false, true, false);
    return SAL_TC(true, true, false);
#undef SAL_TC
}

}  // namespace

bool sal_pass_tf32_supported(const sal_ctx* c, const PassArgs& a) {
    const int allowed = SAL_PASS_UPDATE_H | SAL_PASS_WNUM | SAL_PASS_OBJECTIVE | SAL_PASS_HSUM | SAL_PASS_SCALED_UPDATE;
    if (c->dtype != SAL_F32 || c->V != VT || c->k > 32) return false;
    if (a.flags & ~allowed) return false;
    if ((a.w_kl || a.w_lhalf) && (a.h_scale || (a.flags & SAL_PASS_HSUM))) return false;  // weights are a KLNMF feature
    if ((a.flags & SAL_PASS_HSUM) && !(a.flags & ~SAL_PASS_HSUM)) return false;  // row sums alone: not worth this kernel
    if (a.h_scale && (a.flags & ~SAL_PASS_SCALED_UPDATE) != (SAL_PASS_UPDATE_H | SAL_PASS_OBJECTIVE)) return false;  // the MvNMF trial pass only
    if (((uintptr_t)a.X | (uintptr_t)a.H_in | (uintptr_t)a.H_out) & 15) return false;
    if (c->D >= (int64_t)1 << 31) return false;
    if (c->math != SAL_MATH_TF32_ALWAYS && c->D < SAL_TF32_MIN_SAMPLES) return false;
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
