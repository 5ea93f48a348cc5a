// tc_probe.cu - standalone hardware probe of the tcgen05 kernels (TEST TOOL, not product).
// Runs k_h_update_tc / k_xht_tc on small integer-valued inputs whose exact results are known
// and dumps the raw TMEM accumulators, to validate descriptor layouts on a real B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tests/tc_probe tests/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
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

int main() {
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
        if (tc_plan(p, 0, 148, d, n, KP, KP, dX, ldx, ldh, dH0, dH1)) { printf("plan failed: %s\n", p.err.c_str()); return 1; }
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
