// tc_probe.cu - standalone hardware probe of the tcgen05 kernels (TEST TOOL, not product).
// Runs k_h_update_tc / k_xht_tc on small integer-valued inputs whose exact results are known
// and dumps the raw TMEM accumulators, to validate descriptor layouts on a real B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tests/tc_probe tests/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <string>
#include <algorithm>
#include "../pymf_b200/csrc/kernels_tc.cuh"

using namespace pymfb;
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

template <typename F> static std::vector<float> fill(int64_t rows, int64_t cols, int64_t ld, F f) {
    std::vector<float> v((size_t)rows * ld, 0.f);
    for (int64_t r = 0; r < rows; ++r) for (int64_t c = 0; c < cols; ++c) v[r * ld + c] = f(r, c);
    return v;
}
static float* up(const std::vector<float>& v) { float* p; CHECK(cudaMalloc(&p, v.size() * 4)); CHECK(cudaMemcpy(p, v.data(), v.size() * 4, cudaMemcpyHostToDevice)); return p; }

// ---- raw TMA streaming microbenchmark: how fast can CTAs pull 16 KB boxes from L2 / DRAM? ----
// mode 0: one 2-D tensor box per stage (box_w floats x box_h rows = 16 KB); mode 1: 1-D bulk copy of 16 KB.
// NW producer warps per CTA, each with its own ring of S stages.
template <int S, int NW>
__global__ void __launch_bounds__(32 * NW, 1) k_tma_stream(const __grid_constant__ CUtensorMap map, const float* base_ptr,
                                                          int box_w, int box_h, int nbox_cols, int nbox_rows, int iters, int mode) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base0 = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const int w = threadIdx.x >> 5;
    const uint32_t base = base0 + w * (S * 16384);
    const uint32_t bar = base0 + NW * S * 16384 + w * (8 * S);
    if ((threadIdx.x & 31) == 0) { for (int s = 0; s < S; ++s) tc::mbar_init(bar + 8 * s, 1); tc::fence_barrier_init(); }
    __syncthreads();
    const int nbox = nbox_cols * nbox_rows;
    auto issue = [&](int i, int s) {
        const int b = (int)(((long long)(blockIdx.x * NW + w) + (long long)i * gridDim.x * NW) % nbox);
        tc::mbar_expect_tx(bar + 8 * s, 16384);
        if (mode == 0) tc::tma_load_2d(base + s * 16384, &map, bar + 8 * s, (b % nbox_cols) * box_w, (b / nbox_cols) * box_h);
        else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                          ::"r"(base + s * 16384), "l"(base_ptr + (size_t)b * 4096), "r"(16384), "r"(bar + 8 * s) : "memory");
    };
    if ((threadIdx.x & 31) == 0) {
        for (int i = 0; i < S && i < iters; ++i) issue(i, i);
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < iters; ++i) {
            tc::mbar_wait(bar + 8 * s, ph);
            if (i + S < iters) issue(i + S, s);
            if (++s == S) { s = 0; ph ^= 1; }
        }
    }
}
template <int S, int NW>
static void l2bw_run(const char* name, const CUtensorMap& m, const float* buf, int bw, int bh, int nbc, int nbr, int mode) {
    const int smem = NW * S * 16384 + 2048;
    CHECK(cudaFuncSetAttribute(k_tma_stream<S, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int ctas : {8, 148}) {
        k_tma_stream<S, NW><<<ctas, 32 * NW, smem>>>(m, buf, bw, bh, nbc, nbr, iters, mode);
        CHECK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        k_tma_stream<S, NW><<<ctas, 32 * NW, smem>>>(m, buf, bw, bh, nbc, nbr, iters, mode);
        cudaEventRecord(e1);
        CHECK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("L2BW %-34s depth %d x %d warps, %3d CTAs: %.2f TB/s (%.1f GB/s per CTA)\n", name, S, NW, ctas,
               (double)ctas * NW * iters * 16384 / ms / 1e9, (double)NW * iters * 16384 / ms / 1e6);
    }
}
static bool make_map_any(CUtensorMap* m, const float* base, int64_t rows, int64_t cols, int bw, int bh, CUtensorMapSwizzle sw) {
    EncodeTiledFn fn = get_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
    cuuint32_t estr[2] = {1u, 1u};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static int l2bw_mode() {
    const long long mb = 32;
    const int64_t cols = 8192, rows = mb * 1024 * 1024 / (cols * 4);
    float* buf; CHECK(cudaMalloc(&buf, rows * cols * 4)); CHECK(cudaMemset(buf, 0, rows * cols * 4));
    CUtensorMap m;
    make_map_any(&m, buf, rows, cols, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
    l2bw_run<4, 1>("plain box 128w x 32h", m, buf, 128, 32, (int)(cols / 128), (int)(rows / 32), 0);
    l2bw_run<4, 2>("plain box 128w x 32h", m, buf, 128, 32, (int)(cols / 128), (int)(rows / 32), 0);
    l2bw_run<3, 4>("plain box 128w x 32h", m, buf, 128, 32, (int)(cols / 128), (int)(rows / 32), 0);
    make_map_any(&m, buf, rows, cols, 256, 16, CU_TENSOR_MAP_SWIZZLE_NONE);
    l2bw_run<4, 1>("plain box 256w x 16h", m, buf, 256, 16, (int)(cols / 256), (int)(rows / 16), 0);
    make_map_any(&m, buf, rows, cols, 32, 128, CU_TENSOR_MAP_SWIZZLE_NONE);
    l2bw_run<4, 1>("plain box 32w x 128h", m, buf, 32, 128, (int)(cols / 32), (int)(rows / 128), 0);
    make_map_any(&m, buf, rows, cols, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    l2bw_run<4, 1>("SW128 box 32w x 128h", m, buf, 32, 128, (int)(cols / 32), (int)(rows / 128), 0);
    l2bw_run<4, 2>("SW128 box 32w x 128h", m, buf, 32, 128, (int)(cols / 32), (int)(rows / 128), 0);
    l2bw_run<4, 1>("1-D bulk 16 KB", m, buf, 0, 0, (int)(rows * cols / 4096), 1, 1);
    l2bw_run<4, 2>("1-D bulk 16 KB", m, buf, 0, 0, (int)(rows * cols / 4096), 1, 1);
    cudaFree(buf);
    return 0;
}

static int timing_mode(int64_t d, int64_t n, int KPv, int reps, int64_t pad) {
    DevState* st; CHECK(cudaMalloc(&st, sizeof(DevState))); CHECK(cudaMemset(st, 0, sizeof(DevState)));
    const int64_t ldx = n + pad, ldh = n + pad;
    float *dX, *dW, *dG, *dH0, *dH1, *dP;
    CHECK(cudaMalloc(&dX, d * ldx * 4)); CHECK(cudaMemset(dX, 0, d * ldx * 4));
    CHECK(cudaMalloc(&dW, d * KPv * 4)); CHECK(cudaMemset(dW, 0, d * KPv * 4));
    CHECK(cudaMalloc(&dG, KPv * KPv * 4)); CHECK(cudaMemset(dG, 0, KPv * KPv * 4));
    CHECK(cudaMalloc(&dH0, KPv * ldh * 4)); CHECK(cudaMemset(dH0, 0, KPv * ldh * 4));
    CHECK(cudaMalloc(&dH1, KPv * ldh * 4)); CHECK(cudaMemset(dH1, 0, KPv * ldh * 4));
    CHECK(cudaMalloc(&dP, (d * KPv + KPv * KPv) * 4)); CHECK(cudaMemset(dP, 0, (d * KPv + KPv * KPv) * 4));
    TcPlan p;
    if (tc_plan(p, 0, 148, d, n, KPv, KPv, dX, ldx, ldh, dH0, dH1, 0, pymfb::kNoPanelShift)) { printf("plan failed: %s\n", p.err.c_str()); return 1; }
    int64_t launches = 0;
    tc_after_gram(p, st, dW, dG, 0, &launches);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 2; ++which) {
        for (int w = 0; w < 3; ++w) { if (which == 0) tc_h_update(p, st, dH0, dH1, 0, &launches); else tc_xht(p, st, dH0, dP, 0, &launches); }
        CHECK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) { if (which == 0) tc_h_update(p, st, dH0, dH1, 0, &launches); else tc_xht(p, st, dH0, dP, 0, &launches); }
        cudaEventRecord(e1);
        CHECK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1e3 / reps;
        const double stages_per_cta = which == 0 ? (double)((n + 127) / 128) / 148.0 * ((d + 31) / 32 + KPv / 32)
                                                 : (double)d / 128.0 * (double)n / 32.0 / 148.0;
        printf("TIMING pad=%lld %s d=%lld n=%lld kp=%d: %.1f us/launch, X %.2f TB/s, ~%.0f ns per stage per CTA (%.0f stages/CTA)\n",
               (long long)pad, which == 0 ? "h_update" : "xht", (long long)d, (long long)n, KPv, us, 4.0 * d * n / us / 1e6, us * 1e3 / stages_per_cta, stages_per_cta);
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 2 && std::string(argv[1]) == "l2bw") return l2bw_mode();
    if (argc >= 4) return timing_mode(atoll(argv[1]), atoll(argv[2]), atoi(argv[3]), argc >= 5 ? atoi(argv[4]) : 20, argc >= 6 ? atoll(argv[5]) : 0);
    const int KP = 32;
    const int64_t d = 256, n = 512, ldx = 512, ldh = 512;
    DevState* st; CHECK(cudaMalloc(&st, sizeof(DevState))); CHECK(cudaMemset(st, 0, sizeof(DevState)));
    float* dbg; CHECK(cudaMalloc(&dbg, 128 * 4 * KP * 4));
    struct Case { const char* name; float (*x)(int64_t, int64_t); float (*w)(int64_t, int64_t); float (*h)(int64_t, int64_t); float (*g)(int64_t, int64_t); };
    Case cases[] = {
        {"ones", [](int64_t, int64_t) { return 1.f; }, [](int64_t, int64_t) { return 1.f; }, [](int64_t, int64_t) { return 1.f; }, [](int64_t, int64_t) { return 1.f; }},
        {"x=col%128", [](int64_t, int64_t c) { return (float)(c % 128); }, [](int64_t, int64_t) { return 1.f; }, [](int64_t, int64_t c) { return (float)(c % 128); }, [](int64_t, int64_t) { return 1.f; }},
        {"w=j", [](int64_t, int64_t) { return 1.f; }, [](int64_t, int64_t j) { return (float)j; }, [](int64_t, int64_t) { return 1.f; }, [](int64_t, int64_t j) { return (float)j; }},
        {"k-pair", [](int64_t r, int64_t) { return (float)(r % 8); }, [](int64_t r, int64_t) { return (r % 8 == 3) ? 1.f : 0.f; }, [](int64_t r, int64_t) { return (float)(r % 8); }, [](int64_t r, int64_t) { return (r % 8 == 3) ? 1.f : 0.f; }},
        {"lo-terms", [](int64_t, int64_t) { return 1.f + 1.f / 4096.f; }, [](int64_t, int64_t) { return 1.f + 1.f / 8192.f; }, [](int64_t, int64_t) { return 1.f; }, [](int64_t, int64_t) { return 1.f; }},
    };
    for (auto& cs : cases) {
        auto X = fill(d, n, ldx, cs.x);
        auto W = fill(d, KP, KP, cs.w);
        auto H = fill(KP, n, ldh, cs.h);
        auto G = fill(KP, KP, KP, cs.g);
        float *dX = up(X), *dW = up(W), *dH0 = up(H), *dH1 = up(H), *dG = up(G);
        TcPlan p;
        if (tc_plan(p, 0, 148, d, n, KP, KP, dX, ldx, ldh, dH0, dH1, 0, pymfb::kNoPanelShift)) { printf("plan failed: %s\n", p.err.c_str()); return 1; }
        p.dbg = dbg;
        int64_t launches = 0;
        tc_after_gram(p, st, dW, dG, 0, &launches);
        CHECK(cudaMemset(dbg, 0xFF, 128 * 4 * KP * 4));
        if (tc_h_update(p, st, dH0, dH1, 0, &launches)) { printf("launch failed\n"); return 1; }
        cudaError_t e = cudaDeviceSynchronize();
        printf("== case %s: h_update sync -> %s\n", cs.name, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<float> acc(128 * 4 * KP);
        CHECK(cudaMemcpy(acc.data(), dbg, acc.size() * 4, cudaMemcpyDeviceToHost));
        // expected: C[col][j] = sum_r x(r,col) w(r,j);  D[col][j] = sum_l h(l,col) g(l,j)
        double maxerr = 0;
        for (int col = 0; col < 128; ++col) for (int j = 0; j < KP; ++j) {
            double c = 0, dd = 0;
            for (int64_t r = 0; r < d; ++r) c += (double)cs.x(r, col) * cs.w(r, j);
            for (int64_t l = 0; l < KP; ++l) dd += (double)cs.h(l, col) * cs.g(l, j);
            double gc = (double)acc[col * 2 * KP + j];
            double gd = (double)acc[col * 2 * KP + KP + j];
            maxerr = std::max(maxerr, std::fabs(gc - c) / (std::fabs(c) + 1e-30));
            maxerr = std::max(maxerr, std::fabs(gd - dd) / (std::fabs(dd) + 1e-30));
        }
        printf(" h_update tile0 max rel err vs exact = %.3e\n", maxerr);

        // ---- X H^T: P[row][j] = sum_c x(row,c) h(j,c)
        float* dP; CHECK(cudaMalloc(&dP, (d * KP + KP * KP) * 4)); CHECK(cudaMemset(dP, 0, (d * KP + KP * KP) * 4));
        CHECK(cudaMemset(dbg, 0xFF, 128 * 4 * KP * 4));
        if (tc_xht(p, st, dH0, dP, 0, &launches)) { printf("launch failed\n"); return 1; }
        e = cudaDeviceSynchronize();
        printf("   xht sync -> %s (tasks %d, cols/task %d)\n", cudaGetErrorString(e), p.x_tasks, p.x_cols_per_task);
        if (e != cudaSuccess) return 1;
        std::vector<float> P(d * KP);
        CHECK(cudaMemcpy(P.data(), dP, P.size() * 4, cudaMemcpyDeviceToHost));
        CHECK(cudaMemcpy(acc.data(), dbg, 128 * 2 * KP * 4, cudaMemcpyDeviceToHost));
        printf("   task0 sums row0: [0,1,31]=%g %g %g ; row5 [0]=%g\n", acc[0], acc[1], acc[31], acc[5 * KP]);
        maxerr = 0;
        for (int64_t r = 0; r < d; ++r) for (int j = 0; j < KP; ++j) {
            double a = 0;
            for (int64_t c = 0; c < n; ++c) a += (double)cs.x(r, c) * cs.h(j, c);
            maxerr = std::max(maxerr, std::fabs(P[r * KP + j] - a) / (std::fabs(a) + 1e-30));
        }
        printf("   xht P max rel err vs exact = %.3e   (P[0][0]=%g P[5][3]=%g P[200][31]=%g)\n", maxerr, P[0], P[5 * KP + 3], P[200 * KP + 31]);
        {   // H H^T block (TS kernels only): B[i][j] = sum_c h(i,c) h(j,c)
            std::vector<float> Bm(KP * KP);
            CHECK(cudaMemcpy(Bm.data(), dP + d * KP, Bm.size() * 4, cudaMemcpyDeviceToHost));
            double be = 0;
            for (int i = 0; i < KP; ++i) for (int j = 0; j < KP; ++j) {
                double a = 0;
                for (int64_t c = 0; c < n; ++c) a += (double)cs.h(i, c) * cs.h(j, c);
                be = std::max(be, std::fabs(Bm[i * KP + j] - a) / (std::fabs(a) + 1e-30));
            }
            printf("   H H^T max rel err vs exact = %.3e (use_ts=%d)\n", be, (int)p.use_ts);
        }
        cudaFree(dX); cudaFree(dW); cudaFree(dH0); cudaFree(dH1); cudaFree(dG); cudaFree(dP);
        tc_release(p);
    }
    return 0;
}
