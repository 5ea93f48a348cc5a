// kernels_tc2.cuh - CTA-pair (tcgen05 cta_group::2) variant of the SS H-update kernel, KP = 128.  EXPERIMENTAL, opt-in with
// PYMFB_TC2=1: bit-identical to k_h_update_tc and NOT faster - 2.48-2.56 vs 2.47-2.48 ms on a 16384 x 131072 shard at full
// clocks, 23.1-23.4 vs 23.5-23.6 ms (-1.3 %) on cfg3 under the power cap (same-box A/B).  One tcgen05.mma covers M = 256 = the 128-column tiles of BOTH CTAs of a cluster and takes half of its
// [W_hi|W_lo] operand from each CTA's shared memory, so per stage a CTA stages 24 KB instead of 32 KB of operand chunk
// (L2 -> SM traffic) and its tensor core reads 56 KB instead of 80 KB of operands from shared memory.  Derived from
// k_h_update_tc (kernels_tc.cuh: same rings, same barriers); differences:
//   * CTA `rank` of pair p works on tile 2 * sup + rank (a tile past the end reads zeros by TMA out-of-bounds fill and
//     stores nothing);
//   * chunk ring slot = region B (this CTA's KP columns of [W_hi|W_lo]: W_hi for rank 0, W_lo for rank 1 - the hardware
//     takes N/2 columns from the same offset in each CTA) + region A (this CTA's KP/2 columns of W_hi, for x_lo W_hi);
//   * only the leader's MMA warp issues MMAs; the peer's MMA warp FORWARDS its CTA's ready / tempty barriers to the leader
//     (one release.cluster arrive each - the producers, split and epilogue warps of both CTAs run the single-CTA
//     protocol unchanged); every commit is multicast to both CTAs' done / tfull barriers;
//   * TMEM alloc / dealloc with cta_group::2 between cluster barriers.
#pragma once
#include "kernels_ts2.cuh"

namespace pymfb {
namespace tc {

#ifndef PYMFB_TC2_NW
#define PYMFB_TC2_NW 4   // deeper rings than the single-CTA kernel: a slot also has to outlive the forwarding and multicast-commit latencies (3 / 2: 2.73 vs 2.45 ms, 4 / 3: 2.48-2.56 vs 2.47-2.48 ms)
#endif
#ifndef PYMFB_TC2_NL
#define PYMFB_TC2_NL 3
#endif
template <int KP>
struct H2Cfg {
    static constexpr int NCHB = KP / 32;                            // chunks of region B (KP columns)
    static constexpr int NCHA = KP / 64;                            // chunks of region A (KP/2 columns)
    static constexpr int BSTAGE_BYTES = (NCHB + NCHA) * R1 * 128;   // 24 KB at KP = 128
    static constexpr int NW = PYMFB_TC2_NW;
    static constexpr int NL = PYMFB_TC2_NL;
    static constexpr int NX_RAW = (SMEM_LIMIT - 2048 - NW * BSTAGE_BYTES - NL * XSTAGE_BYTES) / XSTAGE_BYTES;
    static constexpr int NX = NX_RAW > 8 ? 8 : NX_RAW;
    static constexpr int LO_OFF = NX * XSTAGE_BYTES;
    static constexpr int B_OFF = LO_OFF + NL * XSTAGE_BYTES;
    static constexpr int BAR_OFF = B_OFF + NW * BSTAGE_BYTES;
    static constexpr int NR = NW;                                   // the chunk slot index doubles as the ready-barrier index
    static constexpr int NBAR = 2 * NX + NR + 4 + NR + 2;           // fullx, done, ready, tfull[2], tempty[2], peer_ready, peer_tempty[2]
    static constexpr int SMEM_BYTES = BAR_OFF + 1024 + 512;
    static constexpr int SEG_COLS = 2 * KP;
    static constexpr int EPI_WARPS = 8;
    static constexpr int NJ = KP / 2;
    static constexpr int THREADS = 32 * (NPROD + 5 + EPI_WARPS);
    static_assert(KP == 128, "the CTA-pair SS kernel serves KP = 128");
    static_assert(NR >= NL && NR >= NW && NX >= NR, "barrier ring depths");
    static_assert(NBAR * 8 + 8 <= 512, "barrier area too small");
    static_assert(SMEM_BYTES <= SMEM_LIMIT, "rings do not fit");
};
template <class Cfg>
struct H2Bars {
    uint32_t base;
    __device__ __forceinline__ uint32_t fullx(int s) const { return base + 8u * s; }
    __device__ __forceinline__ uint32_t done(int s) const { return base + 8u * (Cfg::NX + s); }
    __device__ __forceinline__ uint32_t ready(int s) const { return base + 8u * (2 * Cfg::NX + s); }
    __device__ __forceinline__ uint32_t tfull(int a) const { return base + 8u * (2 * Cfg::NX + Cfg::NR + a); }
    __device__ __forceinline__ uint32_t tempty(int a) const { return base + 8u * (2 * Cfg::NX + Cfg::NR + 2 + a); }
    __device__ __forceinline__ uint32_t peer_ready(int s) const { return base + 8u * (2 * Cfg::NX + Cfg::NR + 4 + s); }
    __device__ __forceinline__ uint32_t peer_tempty(int a) const { return base + 8u * (2 * Cfg::NX + 2 * Cfg::NR + 4 + a); }
    __device__ __forceinline__ uint32_t tmem_slot() const { return base + 8u * Cfg::NBAR; }
};
// Forwarding arrive on the barrier at the same offset in CTA `cta`.  PYMFB_TC2_RELEASE: release at cluster scope - formally what
// hands the peer's shared-memory tiles over, but measured at ~1.1 us per arrive (the whole pass ran 3.9 vs 2.5 ms: one forward
// per stage became the pace).  Default: relaxed, as in kernels_ts2.cuh - what is handed over never leaves the peer SM: its tiles
// were written by TMA / by the split warps (fence.proxy.async, arrive.release.cta) and observed complete by the forwarding
// thread (acquire.cta) before it arrives here, and the reader is the PEER's own tensor core.
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t local_bar, uint32_t cta) {
#if defined(PYMFB_TC2_RELEASE)
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
#else
    mbar_arrive_cluster(local_bar, cta);
#endif
}
__device__ __forceinline__ void umma2_tf32_lh(uint32_t d_tmem, uint32_t ad_lo, uint32_t ad_hi, uint32_t bd_lo, uint32_t bd_hi,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 ad, bd;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 ad, {%1, %2};\n\t"
        "mov.b64 bd, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], ad, bd, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(ad_lo), "r"(ad_hi), "r"(bd_lo), "r"(bd_hi), "r"(idesc), "r"(accumulate) : "memory");
}

template <int KP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(H2Cfg<KP>::THREADS, 1)
k_h_update_tc2(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
               const __grid_constant__ CUtensorMap mapH, const __grid_constant__ CUtensorMap mapG,
               const DevState* __restrict__ st, const float* __restrict__ Hc, float* __restrict__ Hn,
               float* __restrict__ Hs, int64_t ldh, int d, int n_loc, int num_tiles, float* __restrict__ dbg,
               int kh_rows, int seg_c, float lam, const float* __restrict__ Dp, const float* __restrict__ Dn, int xsh) {
    using Cfg = H2Cfg<KP>;
    if (st->stop) return;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
    const int nsuper = (num_tiles + 1) / 2;      // super-tile = 2 x 128 columns; CTA `rank` owns tile 2 * sup + rank
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const H2Bars<Cfg> bar{smem_base + Cfg::BAR_OFF};
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::BAR_OFF + 8 * Cfg::NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapW); tma_prefetch_desc(&mapH); tma_prefetch_desc(&mapG);
        for (int s = 0; s < Cfg::NX; ++s) { mbar_init(bar.fullx(s), 1); mbar_init(bar.done(s), 1); }
        for (int s = 0; s < Cfg::NR; ++s) { mbar_init(bar.ready(s), 5); mbar_init(bar.peer_ready(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar.tfull(a), 1); mbar_init(bar.tempty(a), Cfg::EPI_WARPS); mbar_init(bar.peer_tempty(a), 1); }
        fence_barrier_init();
    }
    if (warp == NPROD) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar.tmem_slot()), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // both CTAs' barriers initialised and TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    const int nd = (d + R1 - 1) / R1;            // stages over the rows of X
    const int nit = nd + kh_rows / R1;           // + stages over the rows of H (for G H)
    auto xraw = [&](int s) { return smem_base + s * XSTAGE_BYTES; };
    auto xlo = [&](int s) { return smem_base + Cfg::LO_OFF + s * XSTAGE_BYTES; };
    auto wch = [&](int s) { return smem_base + Cfg::B_OFF + s * Cfg::BSTAGE_BYTES; };

    if (warp == 0) {
        // ===== TMA producer of the X / H boxes =====
        RingPos<Cfg::NX> rx;
        for (int sup = pair; sup < nsuper; sup += npairs) {
            const int col0 = (2 * sup + (int)rank) * TILE_COLS;
            for (int it = 0; it < nit; ++it) {
                mbar_wait_relaxed(bar.done(rx.s), rx.ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(bar.fullx(rx.s), XSTAGE_BYTES);
                    const bool xphase = it < nd;
                    const int r0 = (xphase ? it : it - nd) * R1;
                    const CUtensorMap* ma = xphase ? &mapX : &mapH;
#pragma unroll
                    for (int c = 0; c < 4; ++c) tma_load_x(xraw(rx.s) + c * (R1 * 128), ma, bar.fullx(rx.s), col0 + 32 * c, r0, xphase ? xsh : kNoPanel);
                }
                __syncwarp();
                rx.next();
            }
        }
    } else if (warp < NPROD) {
        // ===== TMA producers of this CTA's share of the [W_hi|W_lo] / [G_hi|G_lo] chunk =====
        RingPos<Cfg::NW> rw;
        DoneLag<Cfg::NX, Cfg::NW> lag;
        uint32_t cnt = 0;
        for (int sup = pair; sup < nsuper; sup += npairs) {
            for (int it = 0; it < nit; ++it) {
                const bool mine = cnt++ % (NPROD - 1) == (uint32_t)(warp - 1);
                lag.wait(bar, mine);
                if (mine) {
                    if (elect_one()) {
                        mbar_expect_tx(bar.ready(rw.s), Cfg::BSTAGE_BYTES);
                        const bool xphase = it < nd;
                        const int r0 = (xphase ? it : it - nd) * R1;
                        const CUtensorMap* mb = xphase ? &mapW : &mapG;
                        // region B (operand of x_hi x [b_hi|b_lo], N = 2KP over the pair): this CTA's KP columns = b_hi for rank 0,
                        // b_lo for rank 1; region A (operand of x_lo x b_hi, N = KP): this CTA's half of b_hi
#pragma unroll
                        for (int c = 0; c < Cfg::NCHB; ++c) tma_load_2d(wch(rw.s) + c * (R1 * 128), mb, bar.ready(rw.s), (int)rank * KP + 32 * c, r0);
#pragma unroll
                        for (int c = 0; c < Cfg::NCHA; ++c) tma_load_2d(wch(rw.s) + (Cfg::NCHB + c) * (R1 * 128), mb, bar.ready(rw.s), (int)rank * (KP / 2) + 32 * c, r0);
                    }
                    __syncwarp();
                }
                rw.next();
            }
        }
    } else if (warp == NPROD) {
        if (rank == 0) {
            // ===== leader: issues every tcgen05.mma.cta_group::2 of the pair (M = 256: 128 lanes in each CTA) =====
            constexpr uint32_t idesc_hl = make_idesc(256, 2 * KP, 1, 1);
            constexpr uint32_t idesc_h = make_idesc(256, KP, 1, 1);
            RingPos<Cfg::NX> rx; RingPos<Cfg::NL> rl; RingPos<Cfg::NW> rw;
            uint32_t g = 0;
            const uint64_t d_hi0 = make_desc(xraw(0), R1 * 128, 512, 1), d_lo0 = make_desc(xlo(0), R1 * 128, 512, 1);
            const uint64_t d_b0 = make_desc(wch(0), R1 * 128, 512, 1);
            const uint32_t dh = (uint32_t)(d_hi0 >> 32);
            const uint32_t ahi0 = (uint32_t)d_hi0, alo0 = (uint32_t)d_lo0, b0 = (uint32_t)d_b0;
            constexpr uint32_t REGION_A = (uint32_t)(Cfg::NCHB * (R1 * 128)) >> 4;
            for (int sup = pair; sup < nsuper; sup += npairs) {
                int it = 0;
                while (it < nit) {
                    const int seg_end = (it < nd) ? min(it + seg_c, nd) : nit;
                    const uint32_t b = g & 1u;
                    mbar_wait(bar.tempty(b), ((g >> 1) & 1u) ^ 1u);               // own epilogue
                    mbar_wait_cluster(bar.peer_tempty(b), ((g >> 1) & 1u) ^ 1u);  // the peer's, forwarded
                    tc_fence_after();
                    const uint32_t dcol = tmem_base + b * Cfg::SEG_COLS;
                    bool first = true;
                    for (; it < seg_end; ++it) {
                        mbar_wait(bar.ready(rw.s), rw.ph);
                        mbar_wait_cluster(bar.peer_ready(rw.s), rw.ph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t ox = ahi0 + rx.s * (XSTAGE_BYTES >> 4), ol = alo0 + rl.s * (XSTAGE_BYTES >> 4);
                            const uint32_t ob = b0 + rw.s * (Cfg::BSTAGE_BYTES >> 4);
#pragma unroll
                            for (int kg = 0; kg < R1 / 8; ++kg) {
                                const uint32_t o = kg * (1024 >> 4);
                                umma2_tf32_lh(dcol, ox + o, dh, ob + o, dh, idesc_hl, (first && kg == 0) ? 0u : 1u);
                                umma2_tf32_lh(dcol + KP, ol + o, dh, ob + REGION_A + o, dh, idesc_h, 1u);
                            }
                            umma2_commit(bar.done(rx.s));                          // both CTAs' barriers (same offset)
                            if (it + 1 == seg_end) umma2_commit(bar.tfull(b));
                        }
                        __syncwarp();
                        first = false;
                        rx.next(); rl.next(); rw.next();
                    }
                    ++g;
                }
            }
        } else {
            // ===== peer: forward this CTA's tempty / ready barriers to the leader, in the order the leader consumes them =====
            RingPos<Cfg::NW> rw;
            uint32_t g = 0;
            for (int sup = pair; sup < nsuper; sup += npairs) {
                int it = 0;
                while (it < nit) {
                    const int seg_end = (it < nd) ? min(it + seg_c, nd) : nit;
                    const uint32_t b = g & 1u;
                    // the first use of each accumulator needs no drain: both this wait and the leader's wait on peer_tempty pass on
                    // the fresh barriers, so arrivals are forwarded from the third segment on
                    mbar_wait(bar.tempty(b), ((g >> 1) & 1u) ^ 1u);
                    if (g >= 2 && lane == 0) mbar_arrive_cluster_release(bar.peer_tempty(b), 0);
                    __syncwarp();
                    for (; it < seg_end; ++it) {
                        mbar_wait(bar.ready(rw.s), rw.ph);
                        if (lane == 0) mbar_arrive_cluster_release(bar.peer_ready(rw.s), 0);
                        __syncwarp();
                        rw.next();
                    }
                    ++g;
                }
            }
        }
    } else if (warp < NPROD + 5) {
        // ===== split warps: lo tiles of the X / H operand (single-CTA protocol) =====
        const int tid_s = threadIdx.x - 32 * (NPROD + 1);
        RingPos<Cfg::NX> rx; RingPos<Cfg::NL> rl; RingPos<Cfg::NR> rr;
        DoneLag<Cfg::NX, Cfg::NL> lag;
        for (int sup = pair; sup < nsuper; sup += npairs) {
            for (int it = 0; it < nit; ++it) {
                mbar_wait(bar.fullx(rx.s), rx.ph);
                lag.wait(bar);
                float4* raw = reinterpret_cast<float4*>(smem_gen + rx.s * XSTAGE_BYTES);
                float4* lo = reinterpret_cast<float4*>(smem_gen + Cfg::LO_OFF + rl.s * XSTAGE_BYTES);
                split_buffer(raw, lo, XSTAGE_BYTES / 16, tid_s, 128);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar.ready(rr.s));
                rx.next(); rl.next(); rr.next();
            }
        }
    } else {
        // ===== epilogue warps: drain segments into registers, then the H update =====
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int jbase = ((warp - (NPROD + 5)) >> 2) * Cfg::NJ;   // basis columns [jbase, jbase + NJ) of this warp
        const int nsegC = (nd + seg_c - 1) / seg_c;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t g = 0;
        for (int sup = pair; sup < nsuper; sup += npairs) {
            const int tile = 2 * sup + (int)rank;
            float creg[Cfg::NJ];
#pragma unroll
            for (int j = 0; j < Cfg::NJ; ++j) creg[j] = 0.f;
            for (int seg = 0; seg < nsegC; ++seg, ++g) {
                const uint32_t b = g & 1u;
                mbar_wait_relaxed(bar.tfull(b), (g >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = lane_addr + b * Cfg::SEG_COLS + jbase;
#if !defined(PYMFB_EXP_SS_SKIP_DRAIN)
#pragma unroll
                for (int j0 = 0; j0 < Cfg::NJ; j0 += 16) {
                    float hi[16], sm[16];
                    tmem_ld16(taddr + j0, hi);
                    tmem_ld16(taddr + KP + j0, sm);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) creg[j0 + j] += hi[j] + sm[j];
                }
#else
                (void)taddr;
#endif
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar.tempty(b));
            }
            {   // D segment + H update
                const uint32_t b = g & 1u;
                mbar_wait_relaxed(bar.tfull(b), (g >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = lane_addr + b * Cfg::SEG_COLS + jbase;
                const int col = tile * TILE_COLS + q * 32 + lane;
#pragma unroll
                for (int j0 = 0; j0 < Cfg::NJ; j0 += 16) {
                    float dh[16], dl[16];
                    tmem_ld16(taddr + j0, dh);
                    tmem_ld16(taddr + KP + j0, dl);
                    tmem_ld_wait();
                    if (dbg != nullptr && tile == 0) {   // raw sums of tile 0 (tests/tc_probe.cu)
                        float* o = dbg + (size_t)(q * 32 + lane) * (2 * KP) + jbase + j0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) { o[j] = creg[j0 + j]; o[KP + j] = dh[j] + dl[j]; }
                    }
                    if (col < n_loc) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int64_t o = (int64_t)(jbase + j0 + j) * ldh + col;
                            const float h = Hc[o];
                            const float hn = (Dp != nullptr) ? snmf_ratio(h, creg[j0 + j], Dp[o], Dn[o]) : mu_ratio(h, creg[j0 + j], dh[j] + dl[j], lam);
                            const float hh = __uint_as_float(__float_as_uint(hn) & 0xFFFFE000u);
                            Hn[o] = hn;                                  // new H
                            Hs[hs_index((jbase + j0 + j), col, 2 * KP)] = hh;        // [H_hi ; H_lo] rows for the X.H^T pass
                            Hs[hs_index(KP + (jbase + j0 + j), col, 2 * KP)] = hn - hh;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar.tempty(b));
                ++g;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // the peer may still be reading our barriers / we its
    if (warp == NPROD) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

}  // namespace tc

// ---- host side: opt-in with PYMFB_TC2=1 on shapes the SS kernels serve with 128-wide blocks of bases ----
inline bool tc2_wanted(const TcPlan& p) {
    const char* e = getenv("PYMFB_TC2");
    return e && e[0] == '1' && p.ready && !p.use_ts && p.kpb == 128 && p.sm_count >= 2;
}
inline int tc2_prepare(TcPlan& p) {
    p.use_tc2 = false;
    if (!tc2_wanted(p)) return 0;
    if (cudaFuncSetAttribute(tc::k_h_update_tc2<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::H2Cfg<128>::SMEM_BYTES) != cudaSuccess) return 1;
    p.use_tc2 = true;
    return 0;
}
// same contract as the SS branch of tc_h_update (kernels_tc.cuh)
inline int tc2_h_update(TcPlan& p, const DevState* st, const float* Hc, float* Hn, cudaStream_t stream, int64_t* launches) {
    const int hsrc = (Hc == p.Hbuf[0]) ? 0 : 1;
    p.hs_valid[hsrc ^ 1] = true;
    const int nsuper = (p.h_tiles + 1) / 2;
    const int grid = 2 * std::min(nsuper, p.sm_count / 2);
    for (int b = 0; b < p.nblk; ++b) {
        const size_t hoff = (size_t)b * p.kpb * p.ldh;
        tc::k_h_update_tc2<128><<<grid, tc::H2Cfg<128>::THREADS, tc::H2Cfg<128>::SMEM_BYTES, stream>>>(
            p.mapX_h, p.mapW_b[b], p.mapH_h[hsrc], p.mapG_b[b], st, p.Hbuf[hsrc] + hoff, Hn + hoff, p.Hs[hsrc ^ 1] + 2 * hoff,
            p.ldh, (int)p.d, (int)p.n_loc, p.h_tiles, p.dbg, p.kp, p.seg_c, p.lam_h,
            p.Dp ? p.Dp + hoff : nullptr, p.Dn ? p.Dn + hoff : nullptr, p.xsh);
    }
    *launches += p.nblk;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace pymfb
