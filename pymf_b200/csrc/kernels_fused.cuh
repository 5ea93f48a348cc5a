// kernels_fused.cuh - ONE persistent kernel per iteration: H update + X H^T + H H^T with X read from
// HBM once (k = 32 path).
//
// The two passes of kernels_tc.cuh each stream X from HBM.  Here the column axis is cut into
// super-blocks of SBC = 1024 columns (16 MB of X at d = 4096) and the work of one iteration becomes a
// static list of tasks that the 148 persistent CTAs take round-robin:
//
//   A task (sb, tile, slab)  C_part[slab] = X[slab rows, 128 columns]^T [W_hi|W_lo][slab rows]
//                            (H-update contraction, split over NSLAB row slabs so that a whole
//                            super-block is in flight at once and stays in L2); partial sums go to an
//                            L2-resident ring, the LAST slab to arrive hands the tile to the CTA's
//                            update warps:  Hn = H * sum_slab C_part / (G H + 1e-9)   (G H in fp32 FMA)
//   B task (sb, row block)   P_A[128 rows] += X[128 rows, sb columns] [Hn_hi;Hn_lo]^T   (X again, from L2)
//                            (row block == d/128 is H itself -> P_B += Hn Hn^T)
//
// The list interleaves the A tasks of super-block s with the B tasks of super-block s-1; a B task
// waits (producer warps spin on a global counter) until every tile of its super-block has been
// updated.  Order in the list guarantees progress: a task only ever waits for tasks with a smaller
// index, every CTA runs its tasks in index order, and all CTAs are co-resident.
//
// Per-CTA machinery is that of the TS kernels: NPROD TMA producer warps -> smem ring -> convert
// warps (hi/lo split in registers, parked in TMEM) -> one MMA warp (A from TMEM) -> segment
// accumulators in TMEM -> epilogue warps (RN register sums), plus 4 update warps.
#pragma once
#include "kernels_tc.cuh"

namespace pymfb {
namespace tc {

constexpr int F_KP = 32;
// super-block columns (sbc), tiles per super-block and the C_part ring depth (nslot) are run-time
// parameters of the plan (FusedParams) so that the L2 window = (lag + 1) x d x sbc x 4 B can be tuned.
constexpr int F_CONV_GROUPS = 2;
constexpr int F_WARPS = NPROD + 1 + 4 * F_CONV_GROUPS + 4 + 4;
constexpr int F_THREADS = 32 * F_WARPS;
constexpr int F_CONV_WARP0 = NPROD + 1;
constexpr int F_EPI_WARP0 = F_CONV_WARP0 + 4 * F_CONV_GROUPS;
constexpr int F_UPD_WARP0 = F_EPI_WARP0 + 4;
constexpr int F_UQ = 4;                     // update queue depth

struct FusedParams {
    const DevState* st;
    const float* Hc;        // old H (kp x ldh)
    float* Hn;              // new H
    float* Hs;              // [Hn_hi ; Hn_lo] (2kp x ldh)
    const float* G;         // W^T W (kp x kp, fp32)
    float* PA;              // d x kp   (+= X Hn^T)
    float* PB;              // kp x kp  (+= Hn Hn^T)
    float* Cpart;           // [nslot][sbc][kp] fp32 accumulator of the slab partials (zero between uses)
    int* tile_cnt;          // [n_tiles]   slabs arrived per tile
    int* sb_cnt;            // [n_sb]      tiles updated per super-block
    int64_t ldh;
    int d, n_loc, n_tiles, n_sb, num_rb;
    int nslab, slab_rows;
    int sbc, tiles_per_sb, nslot;
    int bcols, nbsplit;     // B tasks cover bcols columns of a super-block (nbsplit = sbc / bcols parts)
    int* tile_done;         // [n_tiles] 1 once the tile's new H (and its hi/lo split) is visible
    int hint;               // 1: L2 cache hints on the X loads (first read evict_last, second read evict_first)
    int nA, nB;             // A / B tasks per super-block (nA = tiles_per_sb * nslab, nB = num_rb + 1)
    int lag;                // B tasks of super-block s run in step s + lag
    int num_tasks;
};

struct FTask {
    int type;      // 0 = A, 1 = B, -1 = empty
    int sb;
    int tile;      // A: global tile index
    int slab;      // A
    int rb;        // B: row block (== num_rb -> H H^T)
    int nst;       // stages
    int c0;        // B: first column
};

// global task index -> task.  Step s holds the A tasks of super-block s (s < n_sb) interleaved with the
// B tasks of super-block s - lag (s >= lag).  The lag keeps a B task ~2 CTA-waves behind the A tasks it
// depends on, so its wait is normally already satisfied; X stays in L2 for (lag + 1) super-blocks.
__device__ __forceinline__ FTask f_decode(const FusedParams& p, int idx) {
    FTask t;
    t.type = -1; t.sb = 0; t.tile = 0; t.slab = 0; t.rb = 0; t.nst = 0; t.c0 = 0;
    const int T = p.nA + p.nB;
    const int headN = p.lag * p.nA, midN = (p.n_sb - p.lag) * T;
    int step, pos;
    bool isB; int local;
    if (idx < headN) { step = idx / p.nA; pos = idx % p.nA; isB = false; local = pos; }
    else if (idx < headN + midN) {
        const int r = idx - headN;
        step = p.lag + r / T; pos = r % T;
        const int kb0 = (int)(((long long)pos * p.nB) / T), kb1 = (int)(((long long)(pos + 1) * p.nB) / T);
        isB = kb1 > kb0;
        local = isB ? kb0 : pos - kb0;
    } else {
        const int r = idx - headN - midN;
        step = p.n_sb + r / p.nB; pos = r % p.nB; isB = true; local = pos;
        if (step >= p.n_sb + p.lag) return t;
    }
    if (!isB) {
        t.sb = step;
        t.tile = step * p.tiles_per_sb + local / p.nslab;
        t.slab = local % p.nslab;
        if (t.tile >= p.n_tiles) return t;
        const int r0 = t.slab * p.slab_rows;
        const int r1 = min(p.d, r0 + p.slab_rows);
        t.nst = (r1 - r0 + R1 - 1) / R1;
        t.type = 0;
    } else {
        // part-major: the B tasks of the first columns of a super-block come first (their tiles are
        // the first to be updated)
        t.sb = step - p.lag;
        const int part = local / (p.num_rb + 1);
        t.rb = local % (p.num_rb + 1);
        const int c0 = t.sb * p.sbc + part * p.bcols;
        const int c1 = min(p.n_loc, min(c0 + p.bcols, (t.sb + 1) * p.sbc));
        if (c0 >= c1) return t;
        t.c0 = c0;
        t.nst = (c1 - c0 + 31) / 32;
        t.type = 1;
    }
    return t;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until_ge(const int* p, int want) {
    if (ld_acquire_gpu(p) >= want) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu(p) < want) {
        __nanosleep(200);
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}
// TMA load with an L2 eviction-priority hint (createpolicy.fractional.L2::evict_*)
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// 16 values -> hi at column c0.., lo at column 32 + c0.. of an A-ring slot
__device__ __forceinline__ void park_hilo16(uint32_t slot_addr, int c0, const float (&v)[16]) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        hi[i] = __float_as_uint(v[i]) & 0xFFFFE000u;
        lo[i] = __float_as_uint(v[i] - __uint_as_float(hi[i]));
    }
    tmem_st16(slot_addr + c0, hi);
    tmem_st16(slot_addr + 32 + c0, lo);
}

__global__ void __launch_bounds__(F_THREADS, 1)
k_fused_ts(const __grid_constant__ CUtensorMap mapXp,   // X, plain boxes 128 cols x 32 rows   (A tasks)
           const __grid_constant__ CUtensorMap mapW,    // [W_hi|W_lo], MN-major chunks        (A tasks)
           const __grid_constant__ CUtensorMap mapXx,   // X, SW128 boxes 32 cols x 128 rows   (B tasks)
           const __grid_constant__ CUtensorMap mapHs,   // [Hn_hi;Hn_lo], SW128 32 cols x 2kp  (B tasks)
           const __grid_constant__ CUtensorMap mapHa,   // Hn as the A operand of H H^T        (B tasks)
           const FusedParams p) {
    constexpr int KP = F_KP;
    using Cfg = TsCfg<KP>;
    if (p.st->stop) return;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    // smem: ring | G (4 KB) | epilogue transpose tiles (4 x 32 x 33 floats) | barriers | queue
    constexpr int RING = Cfg::STAGES * Cfg::STAGE_BYTES;
    constexpr int G_OFF = RING;
    constexpr int TR_OFF = G_OFF + KP * KP * 4;
    constexpr int BAR_OFF = TR_OFF + 4 * 32 * 33 * 4;
    const uint32_t bar_base = smem_base + BAR_OFF;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    auto afull_bar = [&](int t) { return bar_base + 8u * (2 * Cfg::STAGES + t); };
    auto aempty_bar = [&](int t) { return bar_base + 8u * (2 * Cfg::STAGES + Cfg::NT + t); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 * Cfg::NT + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * Cfg::STAGES + 2 * Cfg::NT + 2 + a); };
    constexpr int NB0 = 2 * Cfg::STAGES + 2 * Cfg::NT + 4;
    auto uqfull_bar = [&](int u) { return bar_base + 8u * (NB0 + u); };
    auto uqempty_bar = [&](int u) { return bar_base + 8u * (NB0 + F_UQ + u); };
    constexpr int NBAR = NB0 + 2 * F_UQ;
    const uint32_t tmem_slot = bar_base + 8u * NBAR;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + BAR_OFF + 8 * NBAR);
    volatile int* uq_tile = reinterpret_cast<volatile int*>(smem_gen + BAR_OFF + 8 * NBAR + 16);   // [F_UQ]
    volatile int* last_flag = reinterpret_cast<volatile int*>(smem_gen + BAR_OFF + 8 * NBAR + 16 + 4 * F_UQ);
    float* Gs = reinterpret_cast<float*>(smem_gen + G_OFF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapXp); tma_prefetch_desc(&mapW); tma_prefetch_desc(&mapXx); tma_prefetch_desc(&mapHs); tma_prefetch_desc(&mapHa);
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int t = 0; t < Cfg::NT; ++t) { mbar_init(afull_bar(t), 4); mbar_init(aempty_bar(t), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        for (int u = 0; u < F_UQ; ++u) { mbar_init(uqfull_bar(u), 1); mbar_init(uqempty_bar(u), 4); }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < KP * KP; i += F_THREADS) Gs[i] = p.G[i];
    if (warp == NPROD) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    auto xs_addr = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES; };
    auto bop = [&](int s) { return smem_base + s * Cfg::STAGE_BYTES + XSTAGE_BYTES; };

    if (warp < NPROD) {
        // ===== TMA producers (stage pcnt % NPROD == warp) =====
        uint32_t pcnt = 0;
        const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
        for (int idx = blockIdx.x; idx < p.num_tasks; idx += gridDim.x) {
            const FTask t = f_decode(p, idx);
            if (t.type < 0) continue;
            if (t.type == 1) {
                // every tile this task reads must have been updated (Hn / Hs final)
                const int tl0 = t.c0 / TILE_COLS, tl1 = min(p.n_tiles - 1, (t.c0 + 32 * t.nst - 1) / TILE_COLS);
                for (int tl = tl0; tl <= tl1; ++tl) spin_until_ge(p.tile_done + tl, 1);
                fence_proxy_async_all();        // generic-proxy stores of other CTAs -> our async-proxy (TMA) reads
            }
            for (int it = 0; it < t.nst; ++it, ++pcnt) {
                if (pcnt % NPROD != (uint32_t)warp) continue;
                const int s = (int)(pcnt % Cfg::STAGES);
                const uint32_t ph = (pcnt / Cfg::STAGES) & 1u;
                mbar_wait(empty_bar(s), ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(full_bar(s), XSTAGE_BYTES + Cfg::BSTAGE_BYTES);
                    if (t.type == 0) {
                        const int r0 = t.slab * p.slab_rows + it * R1;
                        if (p.hint) tma_load_2d_hint(xs_addr(s), &mapXp, full_bar(s), t.tile * TILE_COLS, r0, pol_keep);
                        else tma_load_2d(xs_addr(s), &mapXp, full_bar(s), t.tile * TILE_COLS, r0);
                        tma_load_2d(bop(s), &mapW, full_bar(s), 0, r0);
                        tma_load_2d(bop(s) + R1 * 128, &mapW, full_bar(s), 32, r0);
                    } else {
                        const int c0 = t.c0 + 32 * it;
                        if (t.rb < p.num_rb) {
                            if (p.hint) tma_load_2d_hint(xs_addr(s), &mapXx, full_bar(s), c0, t.rb * 128, pol_drop);
                            else tma_load_2d(xs_addr(s), &mapXx, full_bar(s), c0, t.rb * 128);
                        } else tma_load_2d(xs_addr(s), &mapHa, full_bar(s), c0, 0);
                        tma_load_2d(bop(s), &mapHs, full_bar(s), 0, (c0 >> 5) * (2 * KP));
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == NPROD) {
        // ===== MMA issuer =====
        constexpr uint32_t idA_hl = make_idesc(128, 2 * KP, 0, 1), idA_h = make_idesc(128, KP, 0, 1);
        constexpr uint32_t idB_hl = make_idesc(128, 2 * KP, 0, 0), idB_h = make_idesc(128, KP, 0, 0);
        uint32_t c = 0, g = 0;
        for (int idx = blockIdx.x; idx < p.num_tasks; idx += gridDim.x) {
            const FTask t = f_decode(p, idx);
            if (t.type < 0) continue;
            int it = 0;
            while (it < t.nst) {
                const int seg_end = min(it + SEG_STAGES, t.nst);
                const uint32_t b = g & 1u;
                mbar_wait(tempty_bar(b), ((g >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t dcol = tmem_base + b * Cfg::SEG_COLS;
                bool first = true;
                for (; it < seg_end; ++it, ++c) {
                    const int s = (int)(c % Cfg::STAGES), ts = (int)(c % Cfg::NT);
                    mbar_wait(full_bar(s), (c / Cfg::STAGES) & 1u);
                    mbar_wait(afull_bar(ts), (c / Cfg::NT) & 1u);
                    tc_fence_after();
                    const uint32_t a_hi = tmem_base + Cfg::A_COL0 + ts * 64;
                    if (elect_one()) {
                        if (t.type == 0) {
#pragma unroll
                            for (int kg = 0; kg < 4; ++kg) {
                                const uint64_t bd = make_desc(bop(s) + kg * 1024, R1 * 128, 512, 1);
                                const uint32_t dc = dcol + (kg % Cfg::NCHAIN) * Cfg::CHAIN_COLS;
                                umma_tf32_ts(dc, a_hi + kg * 8, bd, idA_hl, (first && kg < Cfg::NCHAIN) ? 0u : 1u);
                                umma_tf32_ts(dc + KP, a_hi + 32 + kg * 8, bd, idA_h, 1u);
                            }
                        } else {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint64_t bd = make_desc(bop(s) + ks * 32, 16, 1024);
                                const uint32_t dc = dcol + (ks % Cfg::NCHAIN) * Cfg::CHAIN_COLS;
                                umma_tf32_ts(dc, a_hi + ks * 8, bd, idB_hl, (first && ks < Cfg::NCHAIN) ? 0u : 1u);
                                umma_tf32_ts(dc + KP, a_hi + 32 + ks * 8, bd, idB_h, 1u);
                            }
                        }
                        umma_commit(empty_bar(s));
                        umma_commit(aempty_bar(ts));
                    }
                    __syncwarp();
                    first = false;
                }
                if (elect_one()) umma_commit(tfull_bar(b));
                __syncwarp();
                ++g;
            }
        }
    } else if (warp < F_EPI_WARP0) {
        // ===== convert warps: smem tile -> registers -> hi/lo -> TMEM A ring =====
        const int q = warp & 3;
        const int group = (warp - F_CONV_WARP0) >> 2;
        const int mylane = q * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::A_COL0;
        uint32_t c = 0;
        for (int idx = blockIdx.x; idx < p.num_tasks; idx += gridDim.x) {
            const FTask t = f_decode(p, idx);
            if (t.type < 0) continue;
            for (int it = 0; it < t.nst; ++it, ++c) {
                if ((int)(c % F_CONV_GROUPS) != group) continue;
                const int s = (int)(c % Cfg::STAGES), ts = (int)(c % Cfg::NT);
                mbar_wait(full_bar(s), (c / Cfg::STAGES) & 1u);
                mbar_wait(aempty_bar(ts), ((c / Cfg::NT) & 1u) ^ 1u);
                tc_fence_after();
                const uint8_t* stage = smem_gen + s * Cfg::STAGE_BYTES;
                const uint32_t slot = lane_addr + ts * 64;
                if (t.type == 0) {
                    // plain [32 rows][128 cols]: lane = column, TMEM column = row
                    const float* xs = reinterpret_cast<const float*>(stage);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float v[16];
#pragma unroll
                        for (int r = 0; r < 16; ++r) v[r] = xs[(h * 16 + r) * TILE_COLS + mylane];
                        park_hilo16(slot, h * 16, v);
                    }
                } else {
                    // SW128 [128 rows][128 B]: lane = row, 16 B chunk j at (j ^ (row & 7))
                    const uint8_t* rowp = stage + mylane * 128;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float v[16];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 f = *reinterpret_cast<const float4*>(rowp + (((h * 4 + j) ^ (mylane & 7)) << 4));
                            v[4 * j + 0] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
                        }
                        park_hilo16(slot, h * 16, v);
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(afull_bar(ts));
            }
        }
    } else if (warp < F_UPD_WARP0) {
        // ===== epilogue warps: drain segments; A -> partial C + arrival counter; B -> atomics =====
        const int q = warp & 3;
        const int et = threadIdx.x - 32 * F_EPI_WARP0;          // 0..127
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        float* tr = reinterpret_cast<float*>(smem_gen + TR_OFF) + (warp - F_EPI_WARP0) * (32 * 33);
        uint32_t g = 0, uq = 0;
        for (int idx = blockIdx.x; idx < p.num_tasks; idx += gridDim.x) {
            const FTask t = f_decode(p, idx);
            if (t.type < 0) continue;
            const int nseg = (t.nst + SEG_STAGES - 1) / SEG_STAGES;
            float acc[KP];
#pragma unroll
            for (int j = 0; j < KP; ++j) acc[j] = 0.f;
            for (int seg = 0; seg < nseg; ++seg, ++g) {
                const uint32_t b = g & 1u;
                mbar_wait(tfull_bar(b), (g >> 1) & 1u);
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
                        for (int j = 0; j < 16; ++j) acc[j0 + j] += hi[j] + sm[j];
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar(b));
            }
            if (t.type == 0) {
                // partial C of (tile, slab): [slot][slab][column in sb][KP]
                if (t.sb >= p.nslot && et == 0) spin_until_ge(p.sb_cnt + (t.sb - p.nslot), min(p.tiles_per_sb, p.n_tiles - (t.sb - p.nslot) * p.tiles_per_sb));
                named_bar_sync(1, 128);
                // fp32 REDs into the L2-resident accumulator [slot][column in sb][KP] (zero on entry: the
                // update warps clear every block they consume).  Transposed through shared memory so that
                // each warp instruction adds one 128-byte row.
#pragma unroll
                for (int j = 0; j < KP; ++j) tr[lane * 33 + j] = acc[j];
                __syncwarp();
                {
                    const int csb0 = (t.tile % p.tiles_per_sb) * TILE_COLS + q * 32;
                    float* dst = p.Cpart + ((size_t)(t.sb % p.nslot) * p.sbc + csb0) * KP + lane;
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) atomicAdd(dst + r * KP, tr[r * 33 + lane]);
                }
                __syncwarp();
                __threadfence();
                named_bar_sync(1, 128);
                if (et == 0) {
                    const int old = atomicAdd(p.tile_cnt + t.tile, 1);
                    *last_flag = (old == p.nslab - 1) ? 1 : 0;
                    if (old == p.nslab - 1) {
                        __threadfence();
                        const uint32_t u = uq % F_UQ;
                        mbar_wait(uqempty_bar(u), ((uq / F_UQ) & 1u) ^ 1u);
                        uq_tile[u] = t.tile;
                        mbar_arrive(uqfull_bar(u));       // release: the update warps take the tile
                    }
                }
                named_bar_sync(1, 128);
                if (*last_flag) ++uq;
                named_bar_sync(1, 128);                    // last_flag may be rewritten by the next task
            } else {
                // coalesced flush: transpose this warp's 32 rows x KP through shared memory
#pragma unroll
                for (int j = 0; j < KP; ++j) tr[lane * 33 + j] = acc[j];
                __syncwarp();
                const bool hh = t.rb >= p.num_rb;
                const int row_lim = hh ? KP : p.d;
                float* P = hh ? p.PB : p.PA;
                const int row0 = (hh ? 0 : t.rb * 128) + q * 32;
                for (int r = 0; r < 32; ++r)
                    if (row0 + r < row_lim) atomicAdd(P + (int64_t)(row0 + r) * KP + lane, tr[r * 33 + lane]);
                __syncwarp();
            }
        }
        if (et == 0) {       // tell the update warps to exit
            const uint32_t u = uq % F_UQ;
            mbar_wait(uqempty_bar(u), ((uq / F_UQ) & 1u) ^ 1u);
            uq_tile[u] = -1;
            mbar_arrive(uqfull_bar(u));
        }
    } else {
        // ===== update warps: Hn = H * C / (G H + 1e-9) for tiles whose last slab arrived here =====
        const int ut = threadIdx.x - 32 * F_UPD_WARP0;          // 0..127 = column of the tile
        uint32_t uq = 0;
        while (true) {
            const uint32_t u = uq % F_UQ;
            mbar_wait(uqfull_bar(u), (uq / F_UQ) & 1u);
            const int tile = uq_tile[u];
            __syncwarp();
            if (lane == 0) mbar_arrive(uqempty_bar(u));
            ++uq;
            if (tile < 0) break;
            const int sb = tile / p.tiles_per_sb;
            const int col = tile * TILE_COLS + ut;
            const int csb = (tile % p.tiles_per_sb) * TILE_COLS + ut;
            float cs[KP];
            __threadfence();
            {
                float4* src = reinterpret_cast<float4*>(p.Cpart + ((size_t)(sb % p.nslot) * p.sbc + csb) * KP);
#pragma unroll
                for (int j = 0; j < KP; j += 4) {
                    const float4 v = __ldcg(src + j / 4);
                    cs[j] = v.x; cs[j + 1] = v.y; cs[j + 2] = v.z; cs[j + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < KP; j += 4) __stcg(src + j / 4, make_float4(0.f, 0.f, 0.f, 0.f));   // ready for sb + nslot
            }
            if (col < p.n_loc) {
                float h[KP];
#pragma unroll
                for (int l = 0; l < KP; ++l) h[l] = __ldg(p.Hc + (int64_t)l * p.ldh + col);
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    float dj = 0.f;
#pragma unroll
                    for (int l = 0; l < KP; ++l) dj = fmaf(Gs[j * KP + l], h[l], dj);
                    const float hn = (h[j] * cs[j]) / (dj + kEpsDenom);
                    const float hh = __uint_as_float(__float_as_uint(hn) & 0xFFFFE000u);
                    const int64_t o = (int64_t)j * p.ldh + col;
                    p.Hn[o] = hn;
                    p.Hs[hs_index(j, col, 2 * KP)] = hh;
                    p.Hs[hs_index(KP + j, col, 2 * KP)] = hn - hh;
                }
            }
            __threadfence();
            fence_proxy_async_all();
            named_bar_sync(2, 128);
            if (ut == 0) { atomicExch(p.tile_done + tile, 1); atomicAdd(p.sb_cnt + sb, 1); }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == NPROD) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct FusedPlan {
    bool ready = false;
    float* Cpart = nullptr;
    int* cnt = nullptr;          // [n_tiles] tile arrivals, [n_sb] super-block completions, [n_tiles] tile done flags
    int n_tiles = 0, n_sb = 0, nslab = 0, slab_rows = 0, nA = 0, nB = 0, lag = 1, num_tasks = 0;
    int sbc = 1024, tiles_per_sb = 8, nslot = 4, hint = 0, bcols = 1024, nbsplit = 1;
    int smem = 0;
};

inline void fused_release(FusedPlan& f) {
    if (f.Cpart) cudaFree(f.Cpart);
    if (f.cnt) cudaFree(f.cnt);
    f.Cpart = nullptr; f.cnt = nullptr; f.ready = false;
}

// EXPERIMENTAL, off by default: PYMFB_FUSED=1 enables it (k = 32 shapes).  Measured on B200 (cfg2,
// 4096 x 262144): correct (tests/test_gpu_parity.py::test_fused_one_pass_kernel_matches_two_pass) but
// 2.07 ms/iteration at lag 4 against 1.38 ms for the two HBM passes.  The B tasks of a super-block can
// only start ~20 us after its A tasks were issued (task queueing on 148 deep-pipelined CTAs + the
// H update); at 6 TB/s that is ~120 MB of X in flight, i.e. all of L2, so the second read does not
// stay cached unless the lag is too short to hide the dependency.  See DESIGN.md section 5.4.
inline bool fused_wanted(const TcPlan& p) {
    if (!p.ready || !p.use_ts || p.kp != tc::F_KP) return false;
    const char* e = getenv("PYMFB_FUSED");
    return e && e[0] == '1';
}

inline int fused_plan(FusedPlan& f, const TcPlan& p) {
    fused_release(f);
    f.n_tiles = p.h_tiles;
    auto env_int = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
    f.sbc = env_int("PYMFB_FUSED_SBC", 1024);
    f.sbc = std::max(tc::TILE_COLS, f.sbc / tc::TILE_COLS * tc::TILE_COLS);
    f.tiles_per_sb = f.sbc / tc::TILE_COLS;
    f.hint = env_int("PYMFB_FUSED_HINT", 0);
    const int slab_want = std::max(tc::R1, env_int("PYMFB_FUSED_SLAB", 512) / tc::R1 * tc::R1);
    f.n_sb = (int)((p.n_loc + f.sbc - 1) / f.sbc);
    int nslab = (int)std::max<int64_t>(1, (p.d + slab_want / 2) / slab_want);
    f.slab_rows = (int)(((p.d + nslab - 1) / nslab + tc::R1 - 1) / tc::R1 * tc::R1);
    f.nslab = (int)((p.d + f.slab_rows - 1) / f.slab_rows);
    f.nA = f.tiles_per_sb * f.nslab;
    f.bcols = std::min(f.sbc, std::max(32, env_int("PYMFB_FUSED_BCOLS", f.sbc) / 32 * 32));
    f.nbsplit = (f.sbc + f.bcols - 1) / f.bcols;
    f.nB = (p.x_rb + 1) * f.nbsplit;
    {
        const char* e = getenv("PYMFB_FUSED_LAG");
        f.lag = e ? atoi(e) : 3;
        f.lag = std::max(1, std::min(f.lag, f.n_sb));
    }
    f.nslot = f.lag + 4;
    f.num_tasks = f.lag * f.nA + (f.n_sb - f.lag) * (f.nA + f.nB) + f.lag * f.nB;
    if (cudaMalloc(&f.Cpart, (size_t)f.nslot * f.sbc * tc::F_KP * sizeof(float)) != cudaSuccess) return 1;
    if (cudaMemset(f.Cpart, 0, (size_t)f.nslot * f.sbc * tc::F_KP * sizeof(float)) != cudaSuccess) return 1;
    if (cudaMalloc(&f.cnt, (size_t)(2 * f.n_tiles + f.n_sb) * sizeof(int)) != cudaSuccess) return 1;
    using Cfg = tc::TsCfg<tc::F_KP>;
    f.smem = Cfg::STAGES * Cfg::STAGE_BYTES + tc::F_KP * tc::F_KP * 4 + 4 * 32 * 33 * 4 + 1024 /*barriers, queue*/ + 1024 /*align*/;
    if (cudaFuncSetAttribute(tc::k_fused_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, f.smem) != cudaSuccess) return 1;
    f.ready = true;
    return 0;
}

// H[hsrc] -> H[hsrc^1] (and its [hi;lo] companion), P += [X Hn^T | Hn Hn^T].  P must be zero on entry.
inline int fused_launch(FusedPlan& f, TcPlan& p, const DevState* st, const float* Hc, float* Hn, const float* G, float* P,
                        cudaStream_t stream, int64_t* launches) {
    const int hsrc = (Hc == p.Hbuf[0]) ? 0 : 1, hdst = hsrc ^ 1;
    if (cudaMemsetAsync(f.cnt, 0, (size_t)(2 * f.n_tiles + f.n_sb) * sizeof(int), stream) != cudaSuccess) return 1;
    tc::FusedParams fp;
    fp.st = st; fp.Hc = Hc; fp.Hn = Hn; fp.Hs = p.Hs[hdst]; fp.G = G; fp.PA = P; fp.PB = P + p.d * p.kp;
    fp.Cpart = f.Cpart; fp.tile_cnt = f.cnt; fp.sb_cnt = f.cnt + f.n_tiles; fp.tile_done = f.cnt + f.n_tiles + f.n_sb;
    fp.ldh = p.ldh; fp.d = (int)p.d; fp.n_loc = (int)p.n_loc; fp.n_tiles = f.n_tiles; fp.n_sb = f.n_sb; fp.num_rb = p.x_rb;
    fp.bcols = f.bcols; fp.nbsplit = f.nbsplit;
    fp.sbc = f.sbc; fp.tiles_per_sb = f.tiles_per_sb; fp.nslot = f.nslot; fp.hint = f.hint;
    fp.nslab = f.nslab; fp.slab_rows = f.slab_rows; fp.nA = f.nA; fp.nB = f.nB; fp.lag = f.lag; fp.num_tasks = f.num_tasks;
    const int grid = std::min(p.sm_count, f.num_tasks);
    tc::k_fused_ts<<<grid, tc::F_THREADS, f.smem, stream>>>(p.mapX_p, p.mapW, p.mapX_x, p.mapH_x[hdst], p.mapH_a[hdst], fp);
    *launches += 1;
    p.hs_valid[hdst] = true;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace pymfb
