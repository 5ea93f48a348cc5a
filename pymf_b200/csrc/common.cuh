// common.cuh - shared definitions for libpymfb (NMF multiplicative updates on sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pymfb {

constexpr float kEpsDenom = 1e-9f;   // pymf/nmf.py:124,130  (+ 10**-9 on the denominator)
constexpr double kEpsConv = 1e-8;    // pymf/nmf.py:69,136   (NMF._EPS)

// Device-resident loop state.  The stop decision of factorize() (pymf/nmf.py:198-202)
// is taken on the device so that the host never synchronises inside the loop: once
// `stop` is set every later kernel of the run returns immediately.
// Layout of the data matrix X.  Row-major (leading dimension ldx) for borrowed matrices and small ones; context-owned
// copies of wide matrices are PANEL-MAJOR: column panels of 2^xsh columns, each a dense d x 2^xsh block (row stride
// ldx = 2^xsh), xps floats apart.  A TMA box of 128 rows then spans 128 x 16 KB = one 2 MB page instead of 128 pages
// (row stride 4 MB at n = 2^20): the X H^T pass of cfg3 ran 12 % faster at a quarter of the stride.
// Element (r, c) lives at X + xpanel_off(c, xps, xsh) + r * ldx + c;  row-major is xps = 0, xsh = kNoPanelShift.
constexpr int kNoPanelShift = 40;
__host__ __device__ __forceinline__ int64_t xpanel_off(int64_t col, int64_t xps, int xsh) {
    const int64_t p = col >> xsh;
    return p * xps - (p << xsh);
}

// xx - 2<W,A> + <G,B> carries an absolute error of ~2e-7 ||X||^2 (fp32 A, B and their accumulation), i.e. a relative
// error of ~2e-7 ||X||^2 / e^2 in ferr^2.  Below this ratio (ferr off by > ~1e-4) the identity is no longer trusted.
constexpr double kTraceCancel = 1e-3;

struct DevState {
    int stop;            // 1 after converged(i) fired
    int n_exec;          // iterations executed when stop fired (i + 1)
    unsigned ticket;     // last-block-done counter of the error reduction
    unsigned ticket2;    // same, for ||X||^2
    double xx;           // ||X||_F^2 over ALL ranks
    double xx_local;     // this rank's share (all-reduced into xx)
    double last_ferr;    // most recent error (pymfb_frobenius)
    double resid_local;  // direct-residual mode: this rank's sum (X - W H)^2
    double resid;        // ... summed over ranks
    unsigned ticket3;    // last-block-done counter of the direct residual
    int cancel;          // trace-identity error below kTraceCancel * ||X||^2 seen: the identity has lost too many digits
                         // (the host switches this context to the direct residual pass for its next calls)
    int it;              // iteration index inside the current run: k_err stores ferr[it] and advances it, so the
                         // per-iteration launch sequence carries no host-side index (replayable as a CUDA graph)
};

// The multiplicative ratio shared by every update kernel.  num = W^T X (or X H^T), den = G H (or W B).
//   lam == 0 : NMF   v * num / (den + 1e-9)                                        pymf/nmf.py:124-126, 130-132
//   lam != 0 : BNMF  v * ((num + 3 lam v^2) / (den + 2 lam v^3 + lam v + 1e-9))    pymf/bnmf.py:79-82, 87-90
// (lam is uniform over the launch, so the branch does not diverge; lam == 0 is exactly the NMF expression.)
__device__ __forceinline__ float mu_ratio(float v, float num, float den, float lam) {
    if (lam == 0.f) return (v * num) / (den + kEpsDenom);
    const float v2 = v * v;
    return v * ((num + 3.f * lam * v2) / (den + 2.f * lam * (v2 * v) + lam * v + kEpsDenom));
}

// Semi-NMF H ratio (pymf/snmf.py:72-90): c = (W^T X)[j, col] (signed), dp = (G+ H)[j, col], dn = (G- H)[j, col]
// with G+ = (|G| + G)/2, G- = (|G| - G)/2 (separate_positive / separate_negative, :73-77):
//   H <- H * sqrt((c+ + dn) / (c- + dp + 1e-9))
__device__ __forceinline__ float snmf_ratio(float h, float c, float dp, float dn) {
    const float cp = fmaxf(c, 0.f), cn = fmaxf(-c, 0.f);
    return h * sqrtf((cp + dn) / (cn + dp + kEpsDenom));
}

__device__ __forceinline__ uint64_t mix64(uint64_t seed, uint64_t idx) {
    // splitmix64 finaliser; identical to oracle/nmf_oracle.py:hash_uniform
    uint64_t z = idx + seed * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t idx) {
    return (float)(uint32_t)(mix64(seed, idx) >> 40) * 5.9604644775390625e-08f;  // 2^-24
}

}  // namespace pymfb
