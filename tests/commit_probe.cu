// commit_probe.cu - what does a tcgen05.commit cost the issuing warp?  (TEST TOOL, not product.)
// One warp per CTA issues `groups` groups of n tf32 SS MMAs (M = 128, alternating N = 256 / N = 128 like the k = 128
// kernels) each followed by `ncommit` tcgen05.commit on barriers nobody waits on; reports cycles per group against the
// tensor time of the group.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tests/commit_probe tests/commit_probe.cu
#include <cstdio>
#include <cstdlib>
#include "../pymf_b200/csrc/kernels_tc.cuh"

using namespace pymfb;
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(128, 1) k_probe(int n, int ncommit, int groups, int waits, long long* out, int ts, int nbig, int pattern) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gen = smem_raw + (base - tc::smem_u32(smem_raw));
    const uint32_t a_addr = base, b_addr = base + 16384, bar = base + 16384 + 32768, slot = bar + 64;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(gen)[i] = 0.f;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) tc::mbar_init(bar + 8 * i, 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc(slot, 512);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 16384 + 32768 + 64);
    if (warp == 0) {
        const uint32_t id256 = tc::make_idesc(128, nbig, 0, 0), id128 = tc::make_idesc(128, nbig / 2, 0, 0);
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {
            t0 = clock64();
            uint32_t ph = 0;
            for (int g = 0; g < groups; ++g) {
                if (waits) {                                   // a wait that passes immediately (barrier 7 completed in phase 0 below)
                    for (int w = 0; w < waits; ++w) tc::mbar_wait(bar + 8 * 7, 1);
                }
                if (tc::elect_one()) {
#pragma unroll 8
                    for (int i = 0; i < n; ++i) {
                        const uint64_t ad = tc::make_desc(a_addr + (i & 3) * 32, 16, 1024), bd = tc::make_desc(b_addr + (i & 3) * 32, 16, 1024);
                        // pattern 0: wide / narrow MMAs alternate, the narrow one accumulates into the upper half of the wide one's columns
                        //         1: the same, but the narrow MMA has its own columns (no accumulator range shared by consecutive MMAs)
                        //         2: four wide MMAs, then four narrow ones (overlapping columns as in 0)
                        const bool narrow = pattern == 2 ? ((i >> 2) & 1) : (i & 1);
                        const uint32_t dn = pattern == 1 ? tmem + nbig : tmem + nbig / 2;
                        if (ts) {
                            if (narrow) tc::umma_tf32_ts(dn, tmem + 480 + (i & 3) * 8, bd, id128, 1u);
                            else tc::umma_tf32_ts(tmem, tmem + 448 + (i & 3) * 8, bd, id256, 1u);
                        } else {
                            if (narrow) tc::umma_tf32(dn, ad, bd, id128, 1u);
                            else tc::umma_tf32(tmem, ad, bd, id256, 1u);
                        }
                    }
                    for (int c = 0; c < ncommit; ++c) tc::umma_commit(bar + 8 * (1 + c));
                }
                __syncwarp();
            }
            if (tc::elect_one()) tc::umma_commit(bar);
            __syncwarp();
            tc::mbar_wait(bar, rep & 1);
            t1 = clock64();
            (void)ph;
        }
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
    long long* dout; CHECK(cudaMalloc(&dout, 64));
    const int smem = 16384 + 32768 + 2048;
    CHECK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int groups = 1024;
    for (int ts : {1, 0})
        for (int nbig : {256, 128, 64})
            for (int pattern : {0, 1, 2}) {
                const int n = 32;
                k_probe<<<148, 128, smem>>>(n, 0, groups, 0, dout, ts, nbig, pattern);
                CHECK(cudaDeviceSynchronize());
                long long cyc = 0;
                CHECK(cudaMemcpy(&cyc, dout, sizeof(cyc), cudaMemcpyDeviceToHost));
                printf("%s N %3d/%3d pattern %d : %6.1f cycles / MMA  (nominal %.1f)\n", ts ? "TS" : "SS", nbig, nbig / 2, pattern,
                       (double)cyc / groups / n, (nbig / 2 + nbig / 4) / 2.0);
            }
    return 0;
}
