// Shared device helpers for libbqo_sm100: hyper-sample block layout, Matern-5/2 math,
// warp reductions.  FP64 throughout (the reference is float64 numpy).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/bqo_sm100.h"

#define BQO_CHECK_ARG(cond, idx) \
    do {                         \
        if (!(cond)) return -(idx); \
    } while (0)

#define BQO_RETURN_LAST_ERROR()              \
    do {                                     \
        cudaError_t e__ = cudaGetLastError(); \
        return (int)e__;                     \
    } while (0)

// ---- per-hyper-sample block (doubles), written by bqo_hyper_prepare ------------------------
//  [0] var_noise  [1] mean  [2] sigma2 (1 when the kernel is not scaled)  [3] unused
//  [4 .. 4+MAX)      length scales ls[l]
//  [4+MAX .. 4+2MAX) reciprocal length scales
//  then (PRODUCT only) C[T*T] task covariance, Lt[T*T] its factor (exp'ed lower triangle, or C
//  itself under same_correlation, as chol_base_cov_matrix in sbo/kernels/tasks_kernel.py:224-239)
#define BQO_H_VARNOISE 0
#define BQO_H_MEAN 1
#define BQO_H_SIGMA2 2
#define BQO_H_LS 4
#define BQO_H_INVLS (4 + BQO_MAX_DIM)
#define BQO_H_TASKC (4 + 2 * BQO_MAX_DIM)

__host__ __device__ inline int64_t bqo_hyper_stride_impl(const bqo_kernel_desc& d) {
    int64_t s = BQO_H_TASKC;
    if (d.kind == BQO_KIND_PRODUCT_TASKS) s += 2 * (int64_t)d.n_tasks * d.n_tasks;
    return s;
}

#define BQO_SQRT5 2.23606797749978969640917366873128
#define BQO_FIVE_THIRDS (5.0 / 3.0)

// Matern-5/2 correlation as a function of the squared scaled distance
// (Matern52.cross_cov, sbo/kernels/matern52.py:180-183): abs() as in the reference.
__device__ __forceinline__ double bqo_matern52(double r2) {
    r2 = fabs(r2);
    double r = sqrt(r2);
    double s5r = BQO_SQRT5 * r;
    return (1.0 + s5r + BQO_FIVE_THIRDS * r2) * exp(-s5r);
}

// k(r) and g(r) = k'(r)/r = -(5/3)(1 + sqrt5 r) exp(-sqrt5 r).  The reference evaluates
// k'(r) = part_1 + part_2 (GradientLSMatern52.gradient_respect_distance_cross,
// sbo/kernels/matern52.py:417-424) and divides by r later; the closed form is the same
// function without the cancellation of the sqrt5 terms and without the 0/0 at r = 0
// (where the reference's nan_to_num yields a zero gradient, as g * 0 does here).
__device__ __forceinline__ void bqo_matern52_kg(double r2, double& k, double& g) {
    r2 = fabs(r2);
    double r = sqrt(r2);
    double s5r = BQO_SQRT5 * r;
    double e = exp(-s5r);
    k = (1.0 + s5r + BQO_FIVE_THIRDS * r2) * e;
    g = -BQO_FIVE_THIRDS * (1.0 + s5r) * e;
}

// k'(r) itself, in the reference's association (used where dK/d ls needs k'(r)/r * delta^2).
__device__ __forceinline__ double bqo_matern52_dr(double r, double r2) {
    double e = exp(-BQO_SQRT5 * r);
    double part_1 = (1.0 + BQO_SQRT5 * r + BQO_FIVE_THIRDS * r2) * e * (-BQO_SQRT5);
    double part_2 = e * (BQO_SQRT5 + (10.0 / 3.0) * r);
    return part_1 + part_2;
}

// ---- fast path of the streaming stage (stage C) ---------------------------------------------
// Every FP64 instruction counts there (64 DFMA lanes per SM per clock is the roof), so sqrt and
// exp are open-coded:
//   sqrt : MUFU.RSQ64H seed (SFU pipe), one Goldschmidt step and one Newton correction;
//   exp  : exp(-sqrt5 r) = 2^(n/32), n = round(-32 sqrt5 log2(e) r); 2^(n/32) = 2^k T[j] P(f)
//          with a 32-entry table T[j] = 2^(j/32) in shared memory, a degree-6 Taylor
//          polynomial P on |f| <= 1/2 (argument <= ln2/64, truncation error 3e-18) and the
//          power of two applied to the exponent field with integer adds (ALU pipe).
// Both are accurate to about 1.5 ulp (checked against libm in tests/test_gpu_parity.py).
#define BQO_EXP_TABLE 32

// Constants of the fast path live in constant memory so that DFMA/DMUL read them as c[bank][off]
// operands; as literals ptxas re-materialises each 64-bit constant with two UMOVs inside the loop
// (27% of all issued instructions in the first profile, profiles/r01_inner_maximize_v0.md).
static __constant__ double BQO_FC[10] = {
    2.23606797749978969640917366873128,   // 0 sqrt(5)
    -46.16624130844683,                   // 1 -32/ln2
    -0.02166084938653512,                 // 2 -(ln2/32) high 32 bits
    -5.9631716539705866e-12,              // 3 -(ln2/32) low part
    1.0 / 720.0,                          // 4
    1.0 / 120.0,                          // 5
    1.0 / 24.0,                           // 6
    1.0 / 6.0,                            // 7
    5.0 / 3.0,                            // 8
    -5.0 / 3.0,                           // 9
};

__device__ __forceinline__ void bqo_exp_table_init(double* tab) {
    for (int j = threadIdx.x; j < BQO_EXP_TABLE; j += blockDim.x)
        tab[j] = exp2((double)j / (double)BQO_EXP_TABLE);
}

// sqrt(x) for normal x > 0.
__device__ __forceinline__ double bqo_fast_sqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y;        // ~ sqrt(x)
    double h = 0.5 * y;      // ~ 1 / (2 sqrt(x))
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    double d = fma(-g, g, x);
    return fma(d, h, g);
}

// k(r) and g(r) = k'(r)/r from the squared scaled distance r2 (must be a normal number > 0:
// callers start the accumulation of r2 at 1e-300).
__device__ __forceinline__ void bqo_matern52_kg_fast(double r2, const double* __restrict__ tab,
                                                     double& k, double& g) {
    const double r = bqo_fast_sqrt(r2);
    const double s = BQO_FC[0] * r;  // the reference's np.sqrt(5) * r, same rounding
    // n = round(-s * 32 / ln2); Cody-Waite reduction fr = -s - n ln2/32 with a two-part constant
    const double magic = 6755399441055744.0;  // 1.5 * 2^52
    const double tk = fma(s, BQO_FC[1], magic);
    int n = __double2loint(tk);
    const double kf = tk - magic;
    double fr = fma(kf, BQO_FC[2], -s);   // ln2/32, high 32 bits
    fr = fma(kf, BQO_FC[3], fr);          // ln2/32, low part
    // exp(fr), |fr| <= ln2/64: Taylor degree 6 (truncation error < 3e-18)
    double p = fma(BQO_FC[4], fr, BQO_FC[5]);
    p = fma(p, fr, BQO_FC[6]);
    p = fma(p, fr, BQO_FC[7]);
    p = fma(p, fr, 0.5);
    p = fma(p, fr, 1.0);
    p = fma(p, fr, 1.0);
    n = max(n, -32000);  // exp underflow: keep the exponent arithmetic in the normal range
    double e = p * tab[n & (BQO_EXP_TABLE - 1)];
    e = __hiloint2double(__double2hiint(e) + ((n >> 5) << 20), __double2loint(e));
    // k = (1 + sqrt5 r + 5/3 r^2) e,  g = -(5/3)(1 + sqrt5 r) e
    const double t1 = 1.0 + s;
    k = fma(BQO_FC[8], r2, t1) * e;
    g = (t1 * e) * BQO_FC[9];
}

// NV independent evaluations in lock step.  A single evaluation is one long dependent chain of
// DFMAs (profiles/r01_inner_maximize_v1.md: 46 % of the warp stalls were fixed-latency waits),
// so the hot loop feeds NV of them through every stage together; ptxas then has NV independent
// instructions per stage to interleave.  Outputs: e = exp(-sqrt5 r), t1 = 1 + sqrt5 r and
// poly = 1 + sqrt5 r + 5/3 r^2, i.e. k = poly e and k'(r)/r = -(5/3) t1 e.
template <int NV>
__device__ __forceinline__ void bqo_matern52_fast_v(const double (&r2)[NV],
                                                    const double* __restrict__ tab,
                                                    double (&e)[NV], double (&t1)[NV],
                                                    double (&poly)[NV]) {
    double g[NV], h[NV], w[NV], s[NV], kf[NV], fr[NV], p[NV];
    int n[NV];
    const double magic = 6755399441055744.0;  // 1.5 * 2^52
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(r2[q]));
        g[q] = r2[q] * y;
        h[q] = 0.5 * y;
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) w[q] = fma(-h[q], g[q], 0.5);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        g[q] = fma(g[q], w[q], g[q]);
        h[q] = fma(h[q], w[q], h[q]);
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) w[q] = fma(-g[q], g[q], r2[q]);
#pragma unroll
    for (int q = 0; q < NV; ++q) g[q] = fma(w[q], h[q], g[q]);  // r
#pragma unroll
    for (int q = 0; q < NV; ++q) s[q] = BQO_FC[0] * g[q];
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double tk = fma(s[q], BQO_FC[1], magic);
        n[q] = __double2loint(tk);
        kf[q] = tk - magic;
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) fr[q] = fma(kf[q], BQO_FC[2], -s[q]);
#pragma unroll
    for (int q = 0; q < NV; ++q) fr[q] = fma(kf[q], BQO_FC[3], fr[q]);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(BQO_FC[4], fr[q], BQO_FC[5]);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], BQO_FC[6]);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], BQO_FC[7]);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 0.5);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 1.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 1.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        n[q] = max(n[q], -32000);
        double ev = p[q] * tab[n[q] & (BQO_EXP_TABLE - 1)];
        e[q] = __hiloint2double(__double2hiint(ev) + ((n[q] >> 5) << 20), __double2loint(ev));
        t1[q] = 1.0 + s[q];
        poly[q] = fma(BQO_FC[8], r2[q], t1[q]);
    }
}

// Leanest form, used by the stage-C evaluation loop.  The caller works in coordinates that are
// pre-multiplied by sqrt(5) (BQO_PRESCALE), i.e. it passes q2 = 5 r^2, so that
// s = sqrt5 r = sqrt(q2) needs no multiply.  Returns e = exp(-s) and t1 = 1 + s; with them
//   k = (t1 + q2/3) e            (1 + sqrt5 r + 5/3 r^2 = t1 + q2/3)
//   k'(r)/r = -(5/3) t1 e
// so the caller keeps just two running sums per training point, sum t1 e and sum q2 e.
// 23 FP64 instructions per evaluation including the distance and the two accumulations:
//   sqrt 6 (the 1/(2 sqrt) estimate h is not refreshed: the last correction d*h is already
//   2^-44 small, so h's 2^-22 accuracy suffices), exp 11, t1 1, distance 4, sums 2 - minus one.
template <int NV>
__device__ __forceinline__ void bqo_matern52_q_v(const double (&q2)[NV],
                                                 const double* __restrict__ tab, double (&e)[NV],
                                                 double (&t1)[NV]) {
    double g[NV], h[NV], w[NV], kf[NV], fr[NV], p[NV];
    int n[NV];
    const double magic = 6755399441055744.0;  // 1.5 * 2^52
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(q2[q]));
        g[q] = q2[q] * y;
        h[q] = 0.5 * y;
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) w[q] = fma(-h[q], g[q], 0.5);
#pragma unroll
    for (int q = 0; q < NV; ++q) g[q] = fma(g[q], w[q], g[q]);
#pragma unroll
    for (int q = 0; q < NV; ++q) w[q] = fma(-g[q], g[q], q2[q]);
#pragma unroll
    for (int q = 0; q < NV; ++q) g[q] = fma(w[q], h[q], g[q]);  // s = sqrt(q2)
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double tk = fma(g[q], BQO_FC[1], magic);
        n[q] = __double2loint(tk);
        kf[q] = tk - magic;
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) fr[q] = fma(kf[q], BQO_FC[2], -g[q]);
#pragma unroll
    for (int q = 0; q < NV; ++q) fr[q] = fma(kf[q], BQO_FC[3], fr[q]);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(BQO_FC[4], fr[q], BQO_FC[5]);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], BQO_FC[6]);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], BQO_FC[7]);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 0.5);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 1.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 1.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        n[q] = max(n[q], -32000);
        double ev = p[q] * tab[n[q] & (BQO_EXP_TABLE - 1)];
        e[q] = __hiloint2double(__double2hiint(ev) + ((n[q] >> 5) << 20), __double2loint(ev));
        t1[q] = 1.0 + g[q];
    }
}

__device__ __forceinline__ double bqo_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; `red` must hold >= 32 doubles of shared memory.  Result valid in all threads.
__device__ __forceinline__ double bqo_block_sum(double v, double* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = bqo_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    double t = (lane < nw) ? red[lane] : 0.0;
    t = bqo_warp_sum(t);
    return t;
}

// Task covariance lookup (TasksKernel.cross_cov, sbo/kernels/tasks_kernel.py:251-255).
__device__ __forceinline__ double bqo_task_cov(const double* __restrict__ h, int T, int a, int b) {
    return h[BQO_H_TASKC + a * T + b];
}
