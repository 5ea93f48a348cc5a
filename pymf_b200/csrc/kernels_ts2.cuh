// kernels_ts2.cuh - CTA-pair (tcgen05 cta_group::2) variant of the TS H-update kernel, KP = 32 / 64.  EXPERIMENTAL, opt-in
// with PYMFB_TS2=1: one tcgen05.mma covers M = 256 = the 128-column tiles of BOTH CTAs of a
// cluster, so the pair issues half the instructions / commits per byte of X, and each CTA stages only 1.5 KP instead
// of 2 KP operand columns per stage.  Derived from k_h_update_ts (kernels_tc.cuh); differences:
//   * CTA `rank` of pair p works on tile 2 * sup + rank of super-tile sup (a tile past the end reads zeros by TMA
//     out-of-bounds fill and stores nothing);
//   * operand regions per stage: B = this CTA's KP columns of [W_hi|W_lo] (rank 0: W_hi, rank 1: W_lo), A = this
//     CTA's KP/2 columns of W_hi - the hardware takes N/2 columns from the same offset in each CTA;
//   * only the leader issues MMAs; its afull / tempty barriers collect the arrivals of both CTAs (remote arrives),
//     and every commit is multicast to both CTAs' empty / aempty / tfull barriers;
//   * TMEM alloc / dealloc with cta_group::2 between cluster barriers.
#pragma once
#include "kernels_tc.cuh"

namespace pymfb {
namespace tc {

template <int KP>
struct Ts2Cfg {
    static constexpr int NCHA = KP >= 64 ? KP / 64 : 1;                    // 32-column chunks of region A (KP/2 columns, padded to a chunk)
    static constexpr int BSTAGE_BYTES = (KP / 32 + NCHA) * R1 * 128;      // region B (KP columns) + region A
    static constexpr int STAGE_BYTES = XSTAGE_BYTES + BSTAGE_BYTES;
    static constexpr int STAGES_RAW = (SMEM_LIMIT - 2048) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int NCHAIN = 1;
    static constexpr int CHAIN_COLS = 2 * KP;
    static constexpr int SEG_COLS = CHAIN_COLS;
    static constexpr int A_COL0 = 2 * SEG_COLS;
    static constexpr int NT_RAW = (512 - A_COL0) / 64;
    static constexpr int NT = NT_RAW > 6 ? 6 : NT_RAW;
    static constexpr int EPI_WARPS = 4;
    static constexpr int NJ = KP;
    static constexpr int THREADS = 32 * (NPROD + 5 + EPI_WARPS);
    static constexpr int NBAR = 2 * STAGES + 2 * NT + 4;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 512;
    static_assert(KP == 32 || KP == 64, "the CTA-pair kernel serves KP = 32 and 64");
    static_assert(NBAR * 8 + 8 <= 512, "barrier area too small");
    static_assert(SMEM_BYTES <= SMEM_LIMIT, "stage ring does not fit");
};

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
    // relaxed: what is handed over lives in TENSOR memory and is ordered by tcgen05.wait / tcgen05.fence before this
    // arrive; a release at cluster scope made every arrive wait ~1 us for the thread's outstanding memory traffic
    // (first run: 2 190 cycles per stage, 7.9 ms vs 3.2 ms for the single-CTA kernel)
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait with cluster-scope acquire: the arrivals come from both CTAs of the pair
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    uint32_t spins = 0, ok = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (((++spins) & 0xFFu) == 0u) mbar_watchdog(t0);
    }
}
#if defined(PYMFB_TS2_WAIT_CTA)      // experiment: CTA-scope acquire on the leader's collecting barriers
#define TS2_WAIT(bar, par) mbar_wait(bar, par)
#else
#define TS2_WAIT(bar, par) mbar_wait_cluster(bar, par)
#endif
__device__ __forceinline__ void umma2_tf32_ts_lh(uint32_t d_tmem, uint32_t a_tmem, uint32_t bd_lo, uint32_t bd_hi, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 bd;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 bd, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], bd, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "r"(bd_lo), "r"(bd_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on `bar` in BOTH CTAs of the pair once all MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <int KP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Ts2Cfg<KP>::THREADS, 1)
k_h_update_ts2(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW,
              const __grid_constant__ CUtensorMap mapH, const __grid_constant__ CUtensorMap mapG,
              const DevState* __restrict__ st, const float* __restrict__ Hc, float* __restrict__ Hn,
              float* __restrict__ Hs, int64_t ldh, int d, int n_loc, int num_tiles, float* __restrict__ dbg,
              int seg_c, float lam, const float* __restrict__ Dp, const float* __restrict__ Dn,
              const float* __restrict__ wmean, const float* __restrict__ gmean, const float* __restrict__ xsum, int xsh) {
    // wmean != nullptr: the B operands are the CENTERED W / G (k_split_hilo_centered); the epilogue adds
    // wmean[j] * (column sum of X) to W^T X and gmean[j] * (column sum of the H tile) to G H
    // Dp != nullptr: Semi-NMF (pymf/snmf.py:72-90) - the epilogue takes G+ H and G- H from Dp / Dn (same layout
    // as H, written by k_gh_posneg_simt) instead of the G H accumulator
    __shared__ float s_mean[2 * KP];          // [w_mean | g_mean]: read on the critical tail of every tile, so not from global
    if (threadIdx.x < 2 * KP)
        s_mean[threadIdx.x] = (wmean == nullptr) ? 0.f : (threadIdx.x < KP ? wmean[threadIdx.x] : gmean[threadIdx.x - KP]);
    using Cfg = Ts2Cfg<KP>;
    if (st->stop) return;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
    const int nsuper = (num_tiles + 1) / 2;      // super-tile = 2 x 128 columns; CTA `rank` owns tile 2 * sup + rank
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    auto afull_bar = [&](int t) { return bar_base + 8u * (2 * Cfg::STAGES + t); };
    auto aempty_bar = [&](int t) { return bar_base + 8u * (2 * Cfg::STAGES + Cfg::NT + t); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 * Cfg::NT + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 * Cfg::NT + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * Cfg::NBAR;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * Cfg::NBAR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapX); tma_prefetch_desc(&mapW); tma_prefetch_desc(&mapH); tma_prefetch_desc(&mapG);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        // afull / tempty of the LEADER collect the arrivals of both CTAs (4 convert / EPI_WARPS epilogue warps each)
        for (int t = 0; t < Cfg::NT; ++t) { mbar_init(afull_bar(t), 8); mbar_init(aempty_bar(t), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * Cfg::EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == NPROD) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                           // both CTAs' barriers initialised and TMEM allocated
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    const int nd = (d + R1 - 1) / R1;
    const int nit = nd + KP / R1;
    auto xs_addr = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };                    // [32 rows][128 cols] plain
    auto wch = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + XSTAGE_BYTES; };        // MN-major chunks

    if (warp < NPROD) {
        {
            int s = 0; uint32_t ph = 0; uint32_t pcnt = 0;   // this warp issues stages pcnt % NPROD == warp
            for (int sup = pair; sup < nsuper; sup += npairs) {
                const int tile = 2 * sup + (int)rank; (void)tile;
                const int col0 = tile * TILE_COLS;
                for (int it = 0; it < nit; ++it) {
                    if (pcnt++ % NPROD == (uint32_t)warp) {
                    TRACE_AT(pcnt - 1, 0);
                    mbar_wait(empty_bar(s), ph ^ 1);
                    TRACE_AT(pcnt - 1, 1);
                    if (elect_one()) {
                        mbar_expect_tx(full_bar(s), XSTAGE_BYTES + Cfg::BSTAGE_BYTES);
                        const bool xphase = it < nd;
                        const int r0 = (xphase ? it : it - nd) * R1;
                        tma_load_x(xs_addr(s), xphase ? &mapX : &mapH, full_bar(s), col0, r0, xphase ? xsh : kNoPanel);
                        const CUtensorMap* mb = xphase ? &mapW : &mapG;
                        // region B (operand of a_hi x [b_hi|b_lo], N = 2KP over the pair): this CTA's KP columns = b_hi
                        // for rank 0, b_lo for rank 1; region A (operand of a_lo x b_hi, N = KP): this CTA's half of b_hi
#pragma unroll
                        for (int c = 0; c < KP / 32; ++c) tma_load_2d(wch(s) + c * (R1 * 128), mb, full_bar(s), (int)rank * KP + 32 * c, r0);
#pragma unroll
                        for (int c = 0; c < Cfg::NCHA; ++c) tma_load_2d(wch(s) + (KP / 32 + c) * (R1 * 128), mb, full_bar(s), (int)rank * (KP / 2) + 32 * c, r0);
                    }
                    __syncwarp();
                    }
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == NPROD) {
        if (rank == 0) {   // the leader CTA issues every tcgen05.mma.cta_group::2 of the pair (M = 256: 128 lanes in each CTA)
            constexpr uint32_t idesc_hl = make_idesc(256, 2 * KP, 0, 1);
            constexpr uint32_t idesc_h = make_idesc(256, KP, 0, 1);
            int s = 0; uint32_t ph = 0; int t = 0; uint32_t tph = 0; uint32_t g = 0; uint32_t mc = 0;
            // [W_hi|W_lo] / [G_hi|G_lo] operand descriptor of stage 0; stage s, k-group kg add (s * STAGE_BYTES + kg * 1024) >> 4
            // to the low word (shared-memory addresses are < 2^18, so the 14-bit address field never carries)
            const uint64_t bd0 = make_desc(wch(0), R1 * 128, 512, 1);
            const uint32_t bd_hi = (uint32_t)(bd0 >> 32), bd_lo0 = (uint32_t)bd0;
            constexpr uint32_t REGION_A = (uint32_t)((KP / 32) * (R1 * 128)) >> 4;   // region A follows the KP columns of region B
            uint32_t soff = 0;
            for (int sup = pair; sup < nsuper; sup += npairs) {
                const int tile = 2 * sup + (int)rank; (void)tile;
                int it = 0;
                while (it < nit) {
                    const int seg_end = (it < nd) ? min(it + seg_c, nd) : nit;
                    const uint32_t b = g & 1u;
                    TRACE_AT(mc, 8);
                    TS2_WAIT(tempty_bar(b), ((g >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t dcol = tmem_base + b * Cfg::SEG_COLS;
                    bool first = true;
                    for (; it < seg_end; ++it, ++mc) {
                        TRACE_AT(mc, 5);
                        // afull(t) implies full(s): the convert warps waited on full(s) - the barrier that also counts the
                        // [W_hi|W_lo] bytes of the stage - before they filled A slot t and arrived on afull(t)
                        TS2_WAIT(afull_bar(t), tph);
                        TRACE_AT(mc, 6);
                        tc_fence_after();
                        const uint32_t a_hi = tmem_base + Cfg::A_COL0 + t * 64;
                        if (elect_one()) {
                            const uint32_t bl = bd_lo0 + soff;
#pragma unroll
                            for (int kg = 0; kg < R1 / 8; ++kg) {
                                const uint32_t dc = dcol + (kg % Cfg::NCHAIN) * Cfg::CHAIN_COLS;
                                umma2_tf32_ts_lh(dc, a_hi + kg * 8, bl + kg * (1024 >> 4), bd_hi, idesc_hl, (first && kg < Cfg::NCHAIN) ? 0u : 1u);
                                umma2_tf32_ts_lh(dc + KP, a_hi + 32 + kg * 8, bl + REGION_A + kg * (1024 >> 4), bd_hi, idesc_h, 1u);
                            }
                            umma2_commit(empty_bar(s));       // both CTAs' barriers (same offset): stage and A slot are free
                            umma2_commit(aempty_bar(t));
                        }
                        __syncwarp();
                        TRACE_AT(mc, 7);
                        first = false;
                        soff += Cfg::STAGE_BYTES >> 4;
                        if (++s == Cfg::STAGES) { s = 0; ph ^= 1; soff = 0; }
                        if (++t == Cfg::NT) { t = 0; tph ^= 1; }
                    }
                    if (elect_one()) umma2_commit(tfull_bar(b));
                    __syncwarp();
                    ++g;
                }
            }
        }
    } else if (warp < NPROD + 5) {
        // ===== convert warps: smem X tile -> registers -> hi/lo -> TMEM A ring =====
        const int q = warp & 3;
        const int mylane = q * 32 + lane;                 // column of the tile = TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::A_COL0;
        int s = 0; uint32_t ph = 0; int t = 0; uint32_t tph = 0; uint32_t cc = 0;
        for (int sup = pair; sup < nsuper; sup += npairs) {
                const int tile = 2 * sup + (int)rank; (void)tile;
            for (int it = 0; it < nit; ++it, ++cc) {
                if (q == 0) TRACE_AT(cc, 2);
                mbar_wait(full_bar(s), ph);
                if (q == 0) TRACE_AT(cc, 9);
                mbar_wait(aempty_bar(t), tph ^ 1);
                if (q == 0) TRACE_AT(cc, 3);
                tc_fence_after();
#if !defined(PYMFB_EXP_SKIP_CONVERT)
                const float* xs = reinterpret_cast<const float*>(smem_gen + s * Cfg::STAGE_BYTES);
                float v[32];
#pragma unroll
                for (int r = 0; r < 32; ++r) v[r] = xs[r * TILE_COLS + mylane];
                park_hilo(lane_addr + t * 64, v);
                tmem_st_wait();
#endif
                tc_fence_before();
                __syncwarp();
                if (q == 0) TRACE_AT(cc, 4);
                if (lane == 0) mbar_arrive_cluster(afull_bar(t), 0);    // the leader's barrier counts both CTAs
                if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
                if (++t == Cfg::NT) { t = 0; tph ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;
        const int nsegC = (nd + seg_c - 1) / seg_c;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t g = 0;
        for (int sup = pair; sup < nsuper; sup += npairs) {
                const int tile = 2 * sup + (int)rank; (void)tile;
            float creg[KP];
#pragma unroll
            for (int j = 0; j < KP; ++j) creg[j] = 0.f;
            // Old H of this lane's column, fetched NOW: under a saturated memory system a dependent
            // global load takes ~3 us, and doing it after the last segment stalled every tile by ~10 us.
            const int col = tile * TILE_COLS + q * 32 + lane;
            float hreg[KP];
#pragma unroll
            for (int j = 0; j < KP; ++j) hreg[j] = (col < n_loc) ? __ldg(Hc + (int64_t)j * ldh + col) : 0.f;
            const float xs = (wmean != nullptr && col < n_loc) ? __ldg(xsum + col) : 0.f;
            for (int seg = 0; seg < nsegC; ++seg, ++g) {
                const uint32_t b = g & 1u;
                if (q == 0) TRACE_AT(g * seg_c, 10);
                mbar_wait(tfull_bar(b), (g >> 1) & 1u);
                if (q == 0) TRACE_AT(g * seg_c, 11);
                tc_fence_after();
                const uint32_t taddr = lane_addr + b * Cfg::SEG_COLS;
#pragma unroll
                for (int j0 = 0; j0 < KP; j0 += 16) {
#pragma unroll
                    for (int ch = 0; ch < Cfg::NCHAIN; ++ch) {
                        float hi[16], sm[16];
                        tmem_ld16(taddr + ch * Cfg::CHAIN_COLS + j0, hi);
                        tmem_ld16(taddr + ch * Cfg::CHAIN_COLS + KP + j0, sm);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) creg[j0 + j] += hi[j] + sm[j];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (q == 0) TRACE_AT(g * seg_c, 12);
                if (lane == 0) mbar_arrive_cluster(tempty_bar(b), 0);
            }
            {
                const uint32_t b = g & 1u;
                mbar_wait(tfull_bar(b), (g >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = lane_addr + b * Cfg::SEG_COLS;
                float hsum = 0.f;                     // column sum of the old H tile (short-lived: computed where it is used)
                if (wmean != nullptr) {
#pragma unroll
                    for (int j = 0; j < KP; ++j) hsum += hreg[j];
                }
#pragma unroll
                for (int j0 = 0; j0 < KP; j0 += 16) {
                    float dh[16], dl[16];
                    tmem_ld16(taddr + j0, dh);
                    tmem_ld16(taddr + KP + j0, dl);
                    tmem_ld_wait();
                    if (Cfg::NCHAIN > 1) {
                        float eh[16], el[16];
                        tmem_ld16(taddr + Cfg::CHAIN_COLS + j0, eh);
                        tmem_ld16(taddr + Cfg::CHAIN_COLS + KP + j0, el);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) { dh[j] += eh[j]; dl[j] += el[j]; }
                    }
                    if (dbg != nullptr && tile == 0) {
                        float* o = dbg + (size_t)(q * 32 + lane) * (2 * KP) + j0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) { o[j] = creg[j0 + j]; o[KP + j] = dh[j] + dl[j]; }
                    }
#if defined(PYMFB_EXP_SKIP_EPI_GLOBAL)
                    if (col < -1) {
#else
                    if (col < n_loc) {
#endif
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int64_t o = (int64_t)(j0 + j) * ldh + col;
                            const float h = hreg[j0 + j];
                            float cj = creg[j0 + j], dj = dh[j] + dl[j];
                            if (wmean != nullptr) {
                                cj = fmaf(s_mean[j0 + j], xs, cj);
                                dj = fmaf(s_mean[KP + j0 + j], hsum, dj);
                            }
                            const float hn = (Dp != nullptr) ? snmf_ratio(h, cj, Dp[o], Dn[o]) : mu_ratio(h, cj, dj, lam);
                            const float hh = __uint_as_float(__float_as_uint(hn) & 0xFFFFE000u);
                            Hn[o] = hn;                                  // new H
                            Hs[hs_index((j0 + j), col, 2 * KP)] = hh;        // [H_hi ; H_lo] rows for the X.H^T pass
                            Hs[hs_index(KP + (j0 + j), col, 2 * KP)] = hn - hh;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty_bar(b), 0);
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

// ---- host side: opt-in with PYMFB_TS2=1 on shapes the TS kernels serve with kp = 64 ----
inline bool ts2_wanted(const TcPlan& p) {
    const char* e = getenv("PYMFB_TS2");
    return e && e[0] == '1' && p.ready && p.use_ts && (p.kp == 64 || p.kp == 32) && p.sm_count >= 2;
}
inline int ts2_prepare(TcPlan& p) {
    p.use_ts2 = false;
    if (!ts2_wanted(p)) return 0;
    if (cudaFuncSetAttribute(tc::k_h_update_ts2<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Ts2Cfg<64>::SMEM_BYTES) != cudaSuccess ||
        cudaFuncSetAttribute(tc::k_h_update_ts2<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Ts2Cfg<32>::SMEM_BYTES) != cudaSuccess)
        return 1;
    p.use_ts2 = true;
    return 0;
}
// same contract as tc_h_update (kernels_tc.cuh)
inline int ts2_h_update(TcPlan& p, const DevState* st, const float* Hc, float* Hn, cudaStream_t stream, int64_t* launches) {
    const int hsrc = (Hc == p.Hbuf[0]) ? 0 : 1;
    p.hs_valid[hsrc ^ 1] = true;
    const int nsuper = (p.h_tiles + 1) / 2;
    const int grid = 2 * std::min(nsuper, grid_cap(p.sm_count) / 2);
    if (p.kp == 64)
        tc::k_h_update_ts2<64><<<grid, tc::Ts2Cfg<64>::THREADS, tc::Ts2Cfg<64>::SMEM_BYTES, stream>>>(
            p.mapX_p, p.mapW, p.mapH_p[hsrc], p.mapG, st, p.Hbuf[hsrc], Hn, p.Hs[hsrc ^ 1], p.ldh, (int)p.d, (int)p.n_loc, p.h_tiles, p.dbg,
            p.seg_c, p.lam_h, p.Dp, p.Dn, p.center ? p.wmean : nullptr, p.gmean, p.xsum, p.xsh);
    else
        tc::k_h_update_ts2<32><<<grid, tc::Ts2Cfg<32>::THREADS, tc::Ts2Cfg<32>::SMEM_BYTES, stream>>>(
            p.mapX_p, p.mapW, p.mapH_p[hsrc], p.mapG, st, p.Hbuf[hsrc], Hn, p.Hs[hsrc ^ 1], p.ldh, (int)p.d, (int)p.n_loc, p.h_tiles, p.dbg,
            p.seg_c, p.lam_h, p.Dp, p.Dn, p.center ? p.wmean : nullptr, p.gmean, p.xsum, p.xsh);
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace pymfb
