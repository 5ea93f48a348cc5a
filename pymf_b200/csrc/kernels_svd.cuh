// kernels_svd.cuh - NNDSVD initialisation (pymf/nndsvd.py:79-108) on the device.
//
// The reference takes the exact SVD of the data (pymf/svd.py:111-158: eigh of X X^T) and, for every basis i >= 1,
// a second exact SVD of the d x n matrix max(0, s_i u_i v_i^T).  Here:
//   * the leading singular triplets come from SUBSPACE ITERATION with Rayleigh-Ritz extraction.  Its two
//     products per sweep, Q^T X and X Z^T, are the contractions of the H-update pass and of the X H^T pass
//     (k_ltr_partial_simt / k_xht_simt); everything else is b x b work in fp64 on one CTA;
//   * the second SVD has a closed form: max(0, s u v^T) = s (u+ v+^T + u- v-^T) with disjoint supports, so its
//     singular triplets are (s |u+| |v+|, u+/|u+|, v+/|v+|) and the same for the negative parts - the reference's
//     top triplet is the larger of the two (Boutsidis & Gallopoulos 2008, the paper the reference cites).
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace pymfb {
namespace svd {

// ---------------------------------------------------------------------------------------------
// fp64 Gram matrix of b vectors of length len held in a fp32 matrix:
//   by_rows = 0: vectors are the COLUMNS of A (len x b, leading dimension ld)      -> A^T A   (Y^T Y)
//   by_rows = 1: vectors are the ROWS of A (b x len, leading dimension ld)         -> A A^T   (Z Z^T)
// grid.x = splits of len, grid.y = 64 x 64 output blocks (bi * nb + bj, nb = b / 64 rounded up).
// part[split][b * b] is summed in split order by k_sum_f64 (deterministic).
// ---------------------------------------------------------------------------------------------
constexpr int GR_T = 32;
__global__ void __launch_bounds__(256)
k_gram_f64(const float* __restrict__ A, int64_t ld, int b, int64_t len, int by_rows, int64_t len_per_split,
           double* __restrict__ part) {
    __shared__ float ti[GR_T][65], tj[GR_T][65];
    const int nb = (b + 63) / 64;
    const int bi = blockIdx.y / nb, bj = blockIdx.y % nb;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t t0 = (int64_t)blockIdx.x * len_per_split, t1 = min(len, t0 + len_per_split);
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int64_t tc = t0; tc < t1; tc += GR_T) {
        for (int f = threadIdx.x; f < GR_T * 64; f += 256) {
            int t, v;
            if (by_rows) { t = f % GR_T; v = f / GR_T; } else { v = f % 64; t = f / 64; }
            const int64_t tt = tc + t;
            const int vi = bi * 64 + v, vj = bj * 64 + v;
            float a = 0.f, c = 0.f;
            if (tt < t1) {
                if (vi < b) a = by_rows ? A[(int64_t)vi * ld + tt] : A[tt * ld + vi];
                if (vj < b) c = by_rows ? A[(int64_t)vj * ld + tt] : A[tt * ld + vj];
            }
            ti[t][v] = a; tj[t][v] = c;
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < GR_T; ++t) {
            double x[4], y[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { x[i] = (double)ti[t][ty * 4 + i]; y[i] = (double)tj[t][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fma(x[i], y[j], acc[i][j]);
        }
        __syncthreads();
    }
    double* out = part + (int64_t)blockIdx.x * b * b;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = bi * 64 + ty * 4 + i, c = bj * 64 + tx * 4 + j;
            if (r < b && c < b) out[(int64_t)r * b + c] = acc[i][j];
        }
}

__global__ void k_sum_f64(const double* __restrict__ part, int nsplit, int64_t count, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double s = 0.0;
    for (int p = 0; p < nsplit; ++p) s += part[(int64_t)p * count + i];
    out[i] = s;
}

// ---------------------------------------------------------------------------------------------
// Symmetric eigen-decomposition of T (b x b, fp64) by cyclic Jacobi rotations, ONE CTA.
// Outputs: vals[i] descending; R[j][i] = component j of eigenvector i and Rt = R^T, both fp32 row-major b x b.
// work: 2 b^2 doubles (A, V).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_jacobi_eigh_f64(const double* __restrict__ T, int b, double* __restrict__ work, double* __restrict__ vals,
                  float* __restrict__ R, float* __restrict__ Rt) {
    double* A = work;
    double* V = work + (int64_t)b * b;
    __shared__ double s_off, s_diag;
    __shared__ int s_perm[256];
    const int tid = threadIdx.x;
    for (int i = tid; i < b * b; i += blockDim.x) { A[i] = T[i]; V[i] = (i / b == i % b) ? 1.0 : 0.0; }
    __syncthreads();
    for (int sweep = 0; sweep < 40; ++sweep) {
        // off-diagonal / diagonal norms (thread 0 .. b-1 each sum a row, then thread 0 combines)
        if (tid == 0) { s_off = 0.0; s_diag = 0.0; }
        __syncthreads();
        if (tid < b) {
            double o = 0.0;
            for (int j = 0; j < b; ++j) if (j != tid) o += A[tid * b + j] * A[tid * b + j];
            atomicAdd(&s_off, o);
            atomicAdd(&s_diag, A[tid * b + tid] * A[tid * b + tid]);
        }
        __syncthreads();
        if (s_off <= 1e-28 * s_diag || s_diag == 0.0) break;
        for (int p = 0; p < b - 1; ++p) {
            for (int q = p + 1; q < b; ++q) {
                const double apq = A[p * b + q];
                const double app = A[p * b + p], aqq = A[q * b + q];
                // every thread derives the same rotation from values no thread is writing right now
                double c = 1.0, s = 0.0;
                if (fabs(apq) > 1e-300 && fabs(apq) > 1e-17 * sqrt(fabs(app * aqq)) ) {
                    const double tau = (aqq - app) / (2.0 * apq);
                    const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = 1.0 / sqrt(1.0 + t * t);
                    s = t * c;
                }
                __syncthreads();                       // all threads have read app / aqq / apq
                if (s != 0.0) {
                    for (int k = tid; k < b; k += blockDim.x) {      // columns p, q:  A <- A J,  V <- V J
                        const double akp = A[k * b + p], akq = A[k * b + q];
                        A[k * b + p] = c * akp - s * akq;
                        A[k * b + q] = s * akp + c * akq;
                        const double vkp = V[k * b + p], vkq = V[k * b + q];
                        V[k * b + p] = c * vkp - s * vkq;
                        V[k * b + q] = s * vkp + c * vkq;
                    }
                }
                __syncthreads();
                if (s != 0.0) {
                    for (int k = tid; k < b; k += blockDim.x) {      // rows p, q:  A <- J^T A
                        const double apk = A[p * b + k], aqk = A[q * b + k];
                        A[p * b + k] = c * apk - s * aqk;
                        A[q * b + k] = s * apk + c * aqk;
                    }
                }
                __syncthreads();
            }
        }
    }
    __syncthreads();
    // eigenvalues descending: rank of each diagonal entry (ties broken by index)
    if (tid < b) {
        const double v = A[tid * b + tid];
        int rank = 0;
        for (int j = 0; j < b; ++j) {
            const double w = A[j * b + j];
            if (w > v || (w == v && j < tid)) ++rank;
        }
        s_perm[rank] = tid;
    }
    __syncthreads();
    if (tid < b) vals[tid] = A[s_perm[tid] * b + s_perm[tid]];
    for (int i = tid; i < b * b; i += blockDim.x) {
        const int j = i / b, e = i % b;                  // component j of eigenvector e
        const float v = (float)V[j * b + s_perm[e]];
        R[j * b + e] = v;
        Rt[e * b + j] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// Orthonormalisation factor of Y from its Gram matrix M = Y^T Y (fp64, b x b): Cholesky M = L L^T, then
// Ct = (L^-T)^T = L^-1 as fp32 row-major, so that Q = Y L^-T is  Q[r][i] = sum_j Y[r][j] Ct[i][j]
// (the X H^T form of k_xht_simt with H := Ct).  Columns whose pivot collapses (rank-deficient Y: zero padding,
// data of rank < b) are DROPPED: their row of Ct is zero, so Q gets a zero column.  ONE CTA; work: 2 b^2 doubles.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_chol_inv_f64(const double* __restrict__ M, int b, double* __restrict__ work, float* __restrict__ Ct) {
    double* L = work;
    double* Li = work + (int64_t)b * b;
    __shared__ double s_piv, s_maxd;
    const int tid = threadIdx.x;
    for (int i = tid; i < b * b; i += blockDim.x) { L[i] = M[i]; Li[i] = 0.0; }
    __syncthreads();
    if (tid == 0) { double m = 0.0; for (int j = 0; j < b; ++j) m = fmax(m, M[j * b + j]); s_maxd = m; }
    __syncthreads();
    for (int j = 0; j < b; ++j) {
        if (tid == 0) {
            double v = L[j * b + j];
            for (int t = 0; t < j; ++t) v -= L[j * b + t] * L[j * b + t];
            s_piv = (v > 1e-12 * s_maxd && v > 0.0) ? sqrt(v) : 0.0;
            L[j * b + j] = s_piv;
        }
        __syncthreads();
        const double piv = s_piv;
        for (int i = j + 1 + tid; i < b; i += blockDim.x) {
            double v = L[i * b + j];
            for (int t = 0; t < j; ++t) v -= L[i * b + t] * L[j * b + t];
            L[i * b + j] = (piv > 0.0) ? v / piv : 0.0;
        }
        __syncthreads();
    }
    // Li = L^-1 (lower triangular), one column per thread by forward substitution; dropped pivots give zero rows
    for (int c = tid; c < b; c += blockDim.x) {
        for (int i = c; i < b; ++i) {
            const double piv = L[i * b + i];
            if (piv <= 0.0) { Li[i * b + c] = 0.0; continue; }
            double v = (i == c) ? 1.0 : 0.0;
            for (int t = c; t < i; ++t) v -= L[i * b + t] * Li[t * b + c];
            Li[i * b + c] = v / piv;
        }
    }
    __syncthreads();
    for (int i = tid; i < b * b; i += blockDim.x) Ct[i] = (float)Li[i];
}

// ---------------------------------------------------------------------------------------------
// NNDSVD assembly.  U: d x b (fp32, leading dimension b), Vs: b x n (leading dimension ldv) holding
// R^T Q^T X = S V^T (rows NOT yet divided by the singular values), ev[i] = s_i^2.
// k_posneg_norms: norms[i][0..3] = |u+|^2, |u-|^2, |v+|^2, |v-|^2 of vector i (v = Vs[i] / s_i);
//   grid = (k, 2): blockIdx.y = 0 reduces u_i, 1 reduces v_i.
// k_nndsvd_fill: W[:, i] and H[i, :] per pymf/nndsvd.py:88-108.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_posneg_norms(const float* __restrict__ U, int b, int64_t d, const float* __restrict__ Vs, int64_t ldv, int64_t n,
               const double* __restrict__ ev, double* __restrict__ norms) {
    __shared__ double sp[256], sn[256];
    const int i = blockIdx.x;
    const bool isv = blockIdx.y == 1;
    const double s = sqrt(fmax(ev[i], 0.0));
    const double inv = (isv && s > 0.0) ? 1.0 / s : (isv ? 0.0 : 1.0);
    const int64_t len = isv ? n : d;
    double p = 0.0, q = 0.0;
    for (int64_t t = threadIdx.x; t < len; t += blockDim.x) {
        const double v = (double)(isv ? Vs[(int64_t)i * ldv + t] : U[t * b + i]) * inv;
        if (v > 0.0) p += v * v; else q += v * v;
    }
    sp[threadIdx.x] = p; sn[threadIdx.x] = q;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) { sp[threadIdx.x] += sp[threadIdx.x + k]; sn[threadIdx.x] += sn[threadIdx.x + k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { norms[i * 4 + (isv ? 2 : 0)] = sp[0]; norms[i * 4 + (isv ? 3 : 1)] = sn[0]; }
}

__global__ void __launch_bounds__(256)
k_nndsvd_fill(const float* __restrict__ U, int b, int64_t d, const float* __restrict__ Vs, int64_t ldv, int64_t n,
              const double* __restrict__ ev, const double* __restrict__ norms, int k,
              float* __restrict__ W, int kp, float* __restrict__ H, int64_t ldh) {
    const int i = blockIdx.y;
    const double s = sqrt(fmax(ev[i], 0.0));
    const double inv_s = s > 0.0 ? 1.0 / s : 0.0;
    double scale_u, scale_v;
    int mode;                         // 0: |.| (first triplet), +1: positive parts, -1: negative parts
    if (i == 0) {
        mode = 0; scale_u = sqrt(s); scale_v = sqrt(s) * inv_s;                           // nndsvd.py:88-91
    } else {
        const double up = sqrt(norms[i * 4 + 0]), un = sqrt(norms[i * 4 + 1]);
        const double vp = sqrt(norms[i * 4 + 2]), vn = sqrt(norms[i * 4 + 3]);
        const bool pos = up * vp >= un * vn;                                              // top triplet of max(0, s u v^T), :95-102
        mode = pos ? 1 : -1;
        const double nu = pos ? up : un, nv = pos ? vp : vn;
        const double s2 = s * nu * nv;                                                    // its singular value
        scale_u = nu > 0.0 ? sqrt(s2) / nu : 0.0;                                         // :105,108
        scale_v = nv > 0.0 ? sqrt(s2) / nv * inv_s : 0.0;
    }
    const int64_t total = d + n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const bool isv = t >= d;
        const int64_t idx = isv ? t - d : t;
        double v = isv ? (double)Vs[(int64_t)i * ldv + idx] : (double)U[idx * b + i];
        if (mode == 0) v = fabs(v);
        else if (mode == 1) v = fmax(v, 0.0);
        else v = fmax(-v, 0.0);
        v *= isv ? scale_v : scale_u;
        if (isv) H[(int64_t)i * ldh + idx] = (float)v; else W[idx * kp + i] = (float)v;
    }
}

// Q0: deterministic pseudo-random start block, entries in [-0.5, 0.5)
__global__ void k_svd_seed(float* __restrict__ Q, int64_t d, int b, int b_real, uint64_t seed) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d * b) return;
    const int c = (int)(i % b);
    uint64_t z = (uint64_t)i + seed * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    Q[i] = (c < b_real) ? (float)(z >> 40) * 5.9604644775390625e-8f - 0.5f : 0.f;
}

}  // namespace svd
}  // namespace pymfb
