// mma_probe.cu - standalone tcgen05.mma issue/throughput probe (TEST TOOL, not product).
// One warp per CTA issues `count` MMAs (M = 128, K = 32 bytes) and waits for their completion; reports cycles per
// MMA for: operand source (SS = A from shared memory, TS = A from tensor memory), N, number of independent
// accumulators the sequence rotates over (1 = every MMA depends on the previous one), and kind (tf32 / bf16).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tests/mma_probe tests/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../pymf_b200/csrc/kernels_tc.cuh"

using namespace pymfb;
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void umma_bf16_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 = tf32 SS, 1 = tf32 TS, 2 = bf16 SS
template <int MODE>
__global__ void __launch_bounds__(128, 1) k_mma_probe(int N, int nacc, int count, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - tc::smem_u32(smem_raw));
    const uint32_t a_addr = base, b_addr = base + 16384, bar = base + 16384 + 32768, slot = bar + 16;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(gen)[i] = 0.f;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc(slot, 512);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 16384 + 32768 + 16);
    if (warp == 0) {
        // accumulators: nacc regions of N columns from column 0; TS A operand at column 448 (64 columns)
        const uint32_t idesc = (MODE == 2)
            ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
            : tc::make_idesc(128, N, 0, 0);
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {                 // rep 0 warms up
            t0 = clock64();
            if (tc::elect_one()) {
#pragma unroll 8
                for (int i = 0; i < count; ++i) {
                    const uint32_t d = tmem + (uint32_t)((i & (nacc - 1)) * N);   // nacc is a power of two (a runtime modulo here cost more than the MMA)
                    const uint64_t bd = tc::make_desc(b_addr + (i & 3) * 32, 16, 1024);
                    if (MODE == 1) tc::umma_tf32_ts(d, tmem + 448 + (i & 3) * 8, bd, idesc, 1u);
                    else if (MODE == 0) tc::umma_tf32(d, tc::make_desc(a_addr + (i & 3) * 32, 16, 1024), bd, idesc, 1u);
                    else umma_bf16_ss(d, tc::make_desc(a_addr + (i & 3) * 32, 16, 1024), bd, idesc, 1u);
                }
                tc::umma_commit(bar);
            }
            __syncwarp();
            tc::mbar_wait(bar, rep & 1);
            t1 = clock64();
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- CTA pair (cta_group::2): M = 256 over two SMs, each CTA holds its 128 rows of A and N/2 rows of B ----
__device__ __forceinline__ void umma2_tf32_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_tf32_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
// mode: 0 = tf32 SS, 1 = tf32 TS
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k_mma_probe2(int N, int count, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - tc::smem_u32(smem_raw));
    const uint32_t a_addr = base, b_addr = base + 16384, bar = base + 16384 + 32768, slot = bar + 16;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(gen)[i] = 0.f;
    const int warp = threadIdx.x >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::fence_barrier_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    tc::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 16384 + 32768 + 16);
    if (warp == 0 && rank == 0) {
        const uint32_t idesc = tc::make_idesc(256, N, 0, 0);
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {
            t0 = clock64();
            if (tc::elect_one()) {
#pragma unroll 8
                for (int i = 0; i < count; ++i) {
                    const uint64_t bd = tc::make_desc(b_addr + (i & 3) * 32, 16, 1024);
                    if (MODE == 1) umma2_tf32_ts(tmem, tmem + 448 + (i & 3) * 8, bd, idesc, 1u);
                    else umma2_tf32_ss(tmem, tc::make_desc(a_addr + (i & 3) * 32, 16, 1024), bd, idesc, 1u);
                }
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(bar), "h"((uint16_t)1) : "memory");
            }
            __syncwarp();
            tc::mbar_wait(bar, rep & 1);
            t1 = clock64();
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int MODE>
static void run2(const char* name, int grid, long long* dout) {
    const int smem = 16384 + 32768 + 2048;
    CHECK(cudaFuncSetAttribute(k_mma_probe2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int count = 4096;
    for (int N : {32, 64, 128, 256}) {
        k_mma_probe2<MODE><<<grid, 128, smem>>>(N, count, dout);
        CHECK(cudaDeviceSynchronize());
        long long cyc = 0;
        CHECK(cudaMemcpy(&cyc, dout, sizeof(cyc), cudaMemcpyDeviceToHost));
        printf("%-12s grid %3d  M 256  N %3d : %7.1f cycles / MMA   (per 128 rows of A: %.1f)\n", name, grid, N,
               (double)cyc / count, (double)cyc / count / 2);
    }
}

template <int MODE>
static void run(const char* name, int grid, long long* dout) {
    const int smem = 16384 + 32768 + 2048;
    CHECK(cudaFuncSetAttribute(k_mma_probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int count = 4096;
    for (int N : {32, 64, 128, 256}) {
        for (int nacc : {1, 2, 4}) {
            if (nacc * N > 448) continue;
            k_mma_probe<MODE><<<grid, 128, smem>>>(N, nacc, count, dout);
            CHECK(cudaDeviceSynchronize());
            long long cyc = 0;
            CHECK(cudaMemcpy(&cyc, dout, sizeof(cyc), cudaMemcpyDeviceToHost));
            printf("%-8s grid %3d  N %3d  accumulators %d : %7.1f cycles / MMA   (floor N/2 = %d)\n", name, grid, N, nacc,
                   (double)cyc / count, N / 2);
        }
    }
}

int main() {
    long long* dout; CHECK(cudaMalloc(&dout, 64));
    for (int grid : {1, 148}) {
        run<0>("tf32 SS", grid, dout);
        run<1>("tf32 TS", grid, dout);
        run<2>("bf16 SS", grid, dout);
    }
    for (int grid : {2, 148}) {
        run2<0>("tf32 SS 2cta", grid, dout);
        run2<1>("tf32 TS 2cta", grid, dout);
    }
    return 0;
}
