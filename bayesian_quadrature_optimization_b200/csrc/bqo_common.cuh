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
