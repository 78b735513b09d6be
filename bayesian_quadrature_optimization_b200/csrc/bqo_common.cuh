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

// Every kernel launch of the library is written `BQO_L, kernel<<<...>>>(...)`: the comma expression
// bumps a process-wide counter first, so bqo_launch_count() reports how many kernels this library
// has launched (bench.py's `gpu_launches`).
extern "C" long long bqo_launch_count_add(long long delta);
#define BQO_L ((void)bqo_launch_count_add(1))

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
// On B200 an FP64 instruction occupies the issue port of an SM sub-partition for two cycles and
// every other instruction for one (tests/microbench/fp64_issue.cu), and the evaluation loop is
// bound by exactly that port (profiles/r01_inner_maximize_v1.md).  So every instruction counts,
// FP64 or not, and sqrt and exp are open-coded:
//   sqrt : MUFU.RSQ64H seed y (SFU pipe, 2^-22), g = u2 y, then ONE third-order (Halley) step:
//          with e = 1 - g y the root is g (1 - e)^(-1/2) = g (1 + e/2 + 3 e^2/8) + (5/16) e^3 g,
//          i.e. e = fma(-g, y, 1); p = fma(e, 3/8, 1/2); u = fma(g, e p, g): 5 FP64 instructions
//          like the Goldschmidt + Newton pair it replaces, but no y/2 (one integer instruction
//          less) and no DFMA with three register sources (one issue cycle less); remainder
//          (5/16) e^3 < 5e-18, rounding about one ulp;
//   exp  : table-driven.  With N = 2^BQO_EXP_BITS table entries per octave,
//          exp(-s) = 2^-(m div N) * 2^-((m mod N)/N) * exp(fr),  m = round(s N / ln2),
//          |fr| <= ln2 / (2N).  The larger the table the shorter the polynomial for exp(fr):
//            N = 32   : degree 6, (ln2/64)^7/5040   = 1.2e-17
//            N = 256  : degree 4, (ln2/512)^5/120   = 3.8e-17
//            N = 2048 : degree 3, (ln2/4096)^4/24   = 3.4e-17   (16 KB of shared memory)
//          all below half an ulp.  The polynomial is kept in MONIC Horner form, e.g.
//          ((f+3)f+6)f+6 = 6 (1 + f + f^2/2 + f^3/6), whose coefficients are exact FP64 immediates
//          (no constant loads); the common factor (1/6) is folded into the table
//          T[j] = 2^(-j/N) / 6, and 2^-(m div N) goes into the exponent field with one IMAD.
// Accuracy: about 1.8 ulp against libm (tests/test_gpu_parity.py compares the whole a/b path
// with the reference at 1e-9).
#ifndef BQO_EXP_BITS
#define BQO_EXP_BITS 11
#endif
#define BQO_EXP_TABLE (1 << BQO_EXP_BITS)
#if BQO_EXP_BITS >= 11
#define BQO_EXP_DEGREE 3
#define BQO_EXP_FOLD 6.0
#elif BQO_EXP_BITS >= 8
#define BQO_EXP_DEGREE 4
#define BQO_EXP_FOLD 24.0
#else
#define BQO_EXP_DEGREE 6
#define BQO_EXP_FOLD 720.0
#endif

// Coordinates of the evaluation loop are pre-multiplied by K = sqrt5 * N / ln2, so that for
// u2 = sum (K dx)^2 the square root u = sqrt(u2) IS the exponent argument in units of ln2/N:
//   exp(-sqrt5 r) = exp(-u c),  c = ln2/N;   m = round(u);   fr = -(u - m) c  (u - m exact)
// i.e. the range reduction needs no multiplier constant and no two-part Cody-Waite step.
#define BQO_LN2 0.693147180559945309417232121458
#define BQO_PRESCALE (BQO_SQRT5 * (double)BQO_EXP_TABLE / BQO_LN2)  // K
#define BQO_LN2_N (BQO_LN2 / (double)BQO_EXP_TABLE)                 // c = ln2 / N
#define BQO_H_TO_K (BQO_LN2_N * BQO_LN2_N / 3.0)                    // sum k = G + this * H
#define BQO_GRAD_FAC (-BQO_FIVE_THIRDS / BQO_PRESCALE)              // d/dx = this * G * dx' / ls
// Bound of each of the two parts (x part, W part) of a squared distance on the unclamped path:
// their sum stays below 2^(2 (BITS + 9)), i.e. u < 2^(BITS + 9) (k < 1e-150 beyond that anyway).
#define BQO_U2_PART_HI ((1023 + 2 * (BQO_EXP_BITS + 9) - 1) << 20)   // high word of 2^(2(BITS+9)-1)

// The one constant that is not an FP64 immediate lives in constant memory so that DFMA/DMUL read
// it as a c[bank][off] operand; as literals ptxas re-materialised each 64-bit constant with two
// UMOVs inside the loop (27% of all issued instructions in the first profile).
static __constant__ double BQO_FC[6] = {
    -BQO_LN2_N,                                      // 0: -c = -ln2/N
    -3.0 / BQO_LN2_N,                                // 1: b2 }  6 exp(-c d) / (-c^3) =
    6.0 / (BQO_LN2_N * BQO_LN2_N),                   // 2: b1 }    ((d + b2) d + b1) d + b0
    -6.0 / (BQO_LN2_N * BQO_LN2_N * BQO_LN2_N),      // 3: b0 }  (degree 3, N = 2048)
    BQO_LN2_N,                                       // 4: c        } per-training-point constants of
    BQO_H_TO_K,                                      // 5: c^2 / 3  } the stage-C loop (same values)
};
// Degree-3 variant that skips the multiplication by -c: the polynomial is taken in d = u - m
// itself, monic, with its three coefficients as uniform-register operands (one LDCU each per
// training point), and the factor -c^3/6 goes into the table.
#ifndef BQO_POLY_UR
#define BQO_POLY_UR (BQO_EXP_DEGREE == 3)
#endif
#if BQO_POLY_UR
#define BQO_TAB_FOLD (-(BQO_LN2_N * BQO_LN2_N * BQO_LN2_N) / 6.0)
#else
#define BQO_TAB_FOLD (1.0 / BQO_EXP_FOLD)
#endif

// Table entry j: T[j] = 2^(-j/N) * fold, with j * 2^(20 - BITS) ADDED to its high word.  The
// evaluation turns the entry of m = q N + j into 2^-q T[j] with one IMAD on the high word,
// hi - m * 2^(20 - BITS) = hi(T[j]) - q * 2^20 (the exponent field), and needs no mask to split
// q from j: the j part of the product cancels the bias stored here (32-bit wrap-around arithmetic
// on both sides).  The biased entry is not a meaningful double until that IMAD has been applied.
__device__ __forceinline__ double bqo_exp_table_entry(int j) {
    const double v = exp2(-(double)j / (double)BQO_EXP_TABLE) * BQO_TAB_FOLD;
    return __hiloint2double(__double2hiint(v) + (j << (20 - BQO_EXP_BITS)), __double2loint(v));
}

__device__ __forceinline__ void bqo_exp_table_init(double* tab) {
    for (int j = threadIdx.x; j < BQO_EXP_TABLE; j += blockDim.x) tab[j] = bqo_exp_table_entry(j);
}

// NV independent evaluations in lock step (a single evaluation is one long dependent chain).
// Input u2 = K^2 r^2 (a normal number > 0: callers start its accumulation at 1e-300).
// Returns e = exp(-sqrt5 r) and u = K r (so sqrt5 r = u ln2/N); with them
//   k = (1 + u c + u2 c^2/3) e,      k'(r)/r = -(5/3) (1 + u c) e,      c = ln2/N
// so the caller keeps three running sums per training point, sum e, sum u e and sum u2 e, and
// applies the constants once per point (no constant operand inside the W-sample loop).
// `tab(m)` reads table entry m mod N (a functor so that the caller decides how the shared-memory
// address is formed).
// FP64 instructions per evaluation (N = 2048): sqrt 5, reduction 3, polynomial 3, table multiply 1.
// CLAMP = false: the caller guarantees u2 <= 2^(2 (BITS + 9)) (bqo_wdist_precompute bounds its table
// and the x part is clamped once per training point), so the per-evaluation clamp is dropped.
template <int NV, class Tab, bool CLAMP = true, bool LOSRC = false>
__device__ __forceinline__ void bqo_matern52_q_v(const double (&u2)[NV], Tab tab, double (&e)[NV],
                                                 double (&t1)[NV], const double* losrc = nullptr) {
    double g[NV], h[NV], w[NV], kf[NV], fr[NV], p[NV];
    unsigned n[NV];
    const double magic = 4503599627370496.0;  // 2^52: low word of (magic + u) = round(u) >= 0
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(u2[q]));
        // LOSRC: the seed's low word is taken from `losrc` (any value the caller has in a dying
        // register pair) instead of being zeroed -- one MOV less per evaluation.  The seed is
        // then off by up to 2^-20 instead of 2^-22: e <= 2.4e-6 and the Halley remainder
        // (5/16) e^3 < 5e-18, far below half an ulp.
        if (LOSRC) y = __hiloint2double(__double2hiint(y), __double2loint(losrc[q]));
        g[q] = u2[q] * y;
        h[q] = y;
    }
    // Halley step: with e = 1 - g y (= 1 - u2 y^2), sqrt(u2) = g (1 - e)^(-1/2) = g (1 + e/2 + 3 e^2/8) + O(e^3)
#pragma unroll
    for (int q = 0; q < NV; ++q) w[q] = fma(-g[q], h[q], 1.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) h[q] = fma(w[q], 0.375, 0.5);
#pragma unroll
    for (int q = 0; q < NV; ++q) w[q] = w[q] * h[q];
#pragma unroll
    for (int q = 0; q < NV; ++q) g[q] = fma(g[q], w[q], g[q]);  // u = sqrt(u2)
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        // One unsigned clamp of the HIGH word of u (u > 0, so the word is monotone in u) bounds
        // u below 2^(BITS+9) + 1: exp(-u c) stops at 2^-512, k <= 1e-150, instead of decaying
        // further, the low word of magic + u can never wrap, and the exponent arithmetic below
        // stays inside the normal range whatever the inputs (tiny length scales included).
        if (CLAMP) {
            unsigned uh = min((unsigned)__double2hiint(g[q]),
                              (unsigned)((1023 + BQO_EXP_BITS + 9) << 20));
            g[q] = __hiloint2double((int)uh, __double2loint(g[q]));
        }
        // (F2I.F64 / I2F.F64 instead of the two DADDs measured 2.3% slower: profiles/r01_tuning_sweeps.md)
        double tk = magic + g[q];
        n[q] = (unsigned)__double2loint(tk);
        kf[q] = tk - magic;              // round(u) as a double (exact)
    }
#if BQO_POLY_UR
#pragma unroll
    for (int q = 0; q < NV; ++q) fr[q] = g[q] - kf[q];  // d = u - m
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fr[q] + BQO_FC[1];
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], BQO_FC[2]);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], BQO_FC[3]);
#else
#pragma unroll
    for (int q = 0; q < NV; ++q) fr[q] = (g[q] - kf[q]) * BQO_FC[0];  // -(u - m) ln2/N
#endif
#if BQO_POLY_UR
#elif BQO_EXP_DEGREE == 3
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fr[q] + 3.0;
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 6.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 6.0);
#elif BQO_EXP_DEGREE == 4
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fr[q] + 4.0;
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 12.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 24.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 24.0);
#else
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fr[q] + 6.0;
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 30.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 120.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 360.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 720.0);
#pragma unroll
    for (int q = 0; q < NV; ++q) p[q] = fma(p[q], fr[q], 720.0);
#endif
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        const double tv = tab(n[q]);
        // 2^-(m div N) into the exponent field of the table entry (see bqo_exp_table_entry)
        int hi;
        asm("mad.lo.s32 %0, %1, %3, %2;"
            : "=r"(hi)
            : "r"((int)n[q]), "r"(__double2hiint(tv)), "n"(-(1 << (20 - BQO_EXP_BITS))));
        e[q] = p[q] * __hiloint2double(hi, __double2loint(tv));
        t1[q] = g[q];
    }
}

// Exp table of the fast evaluation in global memory, for kernels whose CTAs are too small to
// build their own copy in shared memory: 16 KB, L1/L2 resident.  One copy per translation unit
// that uses it, filled once per device by the first launch that needs it.
static __device__ double bqo_g_exp_tab[BQO_EXP_TABLE];

static __global__ void bqo_exp_table_global_init_kernel() {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < BQO_EXP_TABLE; j += gridDim.x * blockDim.x)
        bqo_g_exp_tab[j] = bqo_exp_table_entry(j);
}

struct GlobalExpTab {
    __device__ __forceinline__ double operator()(unsigned m) const {
        return __ldg(&bqo_g_exp_tab[m & (unsigned)(BQO_EXP_TABLE - 1)]);
    }
};

static inline void bqo_ensure_global_exp_table(cudaStream_t cs) {
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || done[dev]) return;
    bqo_exp_table_global_init_kernel<<<8, 256, 0, cs>>>();
    cudaStreamSynchronize(cs);  // once per device: later callers may be on other streams
    done[dev] = true;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of one kernel: set it once
// per (kernel, device) and again only when a larger size is asked for.
template <auto Kernel>
static void bqo_smem_opt_in(size_t bytes) {
    static int granted[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || granted[dev] < (int)bytes) {
        cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (dev >= 0 && dev < 64) granted[dev] = (int)bytes;
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
