// Hyper-sample preparation and kernel-matrix assembly.
//   bqo_hyper_prepare, bqo_scale_points, bqo_kernel_assemble_batched,
//   bqo_kernel_grad_params_batched, bqo_cross_cov_batched, bqo_task_quadrature_weights
#include "bqo_common.cuh"

extern "C" int bqo_abi_version(void) { return BQO_ABI_VERSION; }

#include <atomic>
static std::atomic<long long> g_bqo_launches{0};
extern "C" long long bqo_launch_count_add(long long delta) { return g_bqo_launches.fetch_add(delta) + delta; }
extern "C" int64_t bqo_launch_count(void) { return (int64_t)g_bqo_launches.load(); }

extern "C" int bqo_set_device(int device) { return (int)cudaSetDevice(device); }

extern "C" int64_t bqo_hyper_stride(const bqo_kernel_desc* desc) {
    if (!desc) return -1;
    return bqo_hyper_stride_impl(*desc);
}

static int check_desc(const bqo_kernel_desc* d) {
    if (!d) return 0;
    if (d->kind < 0 || d->kind > 2) return 0;
    if (d->dim_m < 1 || d->dim_m > BQO_MAX_DIM || d->dim < d->dim_m) return 0;
    if (d->kind == BQO_KIND_PRODUCT_TASKS && (d->n_tasks < 1 || d->dim != d->dim_m + 1)) return 0;
    if (d->kind != BQO_KIND_PRODUCT_TASKS && d->dim != d->dim_m) return 0;
    return 1;
}

// ---------------------------------------------------------------------------------------------
// theta -> hyper block.  One CTA per hyper-sample.
// Task covariance: TasksKernel.compute_cov_matrix (sbo/kernels/tasks_kernel.py:198-239).
__global__ void hyper_prepare_kernel(bqo_kernel_desc d, const double* __restrict__ theta,
                                     double* __restrict__ hyp, int64_t stride) {
    int s = blockIdx.x;
    const double* th = theta + (int64_t)s * (2 + d.n_params);
    double* h = hyp + (int64_t)s * stride;
    int t = threadIdx.x;
    if (t == 0) {
        h[BQO_H_VARNOISE] = th[0];
        h[BQO_H_MEAN] = th[1];
        h[BQO_H_SIGMA2] = (d.kind == BQO_KIND_SCALED_MATERN52) ? th[2 + d.dim_m] : 1.0;
        h[3] = 0.0;
    }
    for (int l = t; l < BQO_MAX_DIM; l += blockDim.x) {
        double ls = (l < d.dim_m) ? th[2 + l] : 1.0;
        h[BQO_H_LS + l] = ls;
        h[BQO_H_INVLS + l] = 1.0 / ls;
    }
    if (d.kind != BQO_KIND_PRODUCT_TASKS) return;
    const int T = d.n_tasks;
    const double* p = th + 2 + d.dim_m;
    double* C = h + BQO_H_TASKC;
    double* Lt = C + T * T;
    if (d.same_corr) {
        double offd = (T > 1) ? exp(p[1]) : 0.0;
        double diag = (T > 1) ? (exp(p[0]) + offd * (double)(T - 1)) : exp(p[0]);
        for (int e = t; e < T * T; e += blockDim.x) {
            int a = e / T, b = e % T;
            double v = (a == b) ? diag : offd;
            C[e] = v;
            Lt[e] = v;
        }
        return;
    }
    for (int e = t; e < T * T; e += blockDim.x) {
        int a = e / T, b = e % T;
        Lt[e] = (b <= a) ? exp(p[a * (a + 1) / 2 + b]) : 0.0;
    }
    __syncthreads();
    for (int e = t; e < T * T; e += blockDim.x) {
        int a = e / T, b = e % T;
        int m = a < b ? a : b;
        double acc = 0.0;
        for (int j = 0; j <= m; ++j) acc += Lt[a * T + j] * Lt[b * T + j];
        C[e] = acc;
    }
}

extern "C" int bqo_hyper_prepare(const bqo_kernel_desc* desc, const double* theta, int32_t S,
                                 double* hyp, void* stream) {
    BQO_CHECK_ARG(check_desc(desc), 1);
    BQO_CHECK_ARG(theta != nullptr, 2);
    BQO_CHECK_ARG(S > 0, 3);
    BQO_CHECK_ARG(hyp != nullptr, 4);
    BQO_L, hyper_prepare_kernel<<<S, 128, 0, (cudaStream_t)stream>>>(*desc, theta, hyp,
                                                             bqo_hyper_stride_impl(*desc));
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
__global__ void scale_points_kernel(bqo_kernel_desc d, const double* __restrict__ hyp,
                                    int64_t stride, const double* __restrict__ X, int n, int ldx,
                                    double* __restrict__ Xs, int32_t* __restrict__ tid) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int s = blockIdx.y;
    if (i >= n) return;
    const double* h = hyp + (int64_t)s * stride;
    for (int l = 0; l < d.dim_m; ++l)
        Xs[((int64_t)s * d.dim_m + l) * n + i] = X[(int64_t)i * ldx + l] / h[BQO_H_LS + l];
    if (s == 0 && tid != nullptr && d.kind == BQO_KIND_PRODUCT_TASKS)
        tid[i] = (int32_t)X[(int64_t)i * ldx + d.dim - 1];
}

extern "C" int bqo_scale_points(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                                const double* X, int32_t n, int32_t ldx, double* Xs, int32_t* tid,
                                void* stream) {
    BQO_CHECK_ARG(check_desc(desc), 1);
    BQO_CHECK_ARG(hyp != nullptr, 2);
    BQO_CHECK_ARG(S > 0, 3);
    BQO_CHECK_ARG(X != nullptr, 4);
    BQO_CHECK_ARG(n > 0, 5);
    BQO_CHECK_ARG(ldx >= desc->dim, 6);
    BQO_CHECK_ARG(Xs != nullptr, 7);
    BQO_CHECK_ARG(desc->kind != BQO_KIND_PRODUCT_TASKS || tid != nullptr, 8);
    dim3 grid((n + 255) / 256, S);
    BQO_L, scale_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        *desc, hyp, bqo_hyper_stride_impl(*desc), X, n, ldx, Xs, tid);
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// K[s][i][j] in 64 x 64 tiles, one CTA of 8 warps per tile ON OR BELOW the diagonal (the matrix is
// symmetric bit for bit: (x_i - x_j)^2 and the task covariance do not depend on the order).
//   * the 64 row points sit in shared memory (broadcast loads), each lane keeps its two column
//     points in registers and pushes the two kernel values of a row through the fast Matern
//     evaluation together (independent chains);
//   * a warp writes two 256-byte row segments per row; for the full matrix (`lower_only` = 0) the
//     tile also goes to shared memory and is written once more transposed, as the mirror tile.
// The kernel is FP64-issue bound, not HBM bound: ~30 FP64 + ~15 other instructions per element
// (the 32 x 32 version it replaces executed 33 + 101: six global loads with 64-bit address
// arithmetic per element, and computed both halves of the matrix).  The diagonal is exactly 1
// like the reference's k(x, x).
#define ASM_T 64
template <int DM, bool TASKS>
__global__ void __launch_bounds__(256)
assemble_kernel(bqo_kernel_desc d, const double* __restrict__ hyp, int64_t stride,
                const double* __restrict__ Xs, const int32_t* __restrict__ tid, int n,
                const double* __restrict__ noise_obs, int add_noise,
                const double* __restrict__ extra_diag, double* __restrict__ K, int lower_only) {
    constexpr int DMS = (DM > 0) ? DM : BQO_MAX_DIM;
    __shared__ double s_row[ASM_T][DMS];
    __shared__ int s_tid[ASM_T];
    __shared__ double s_tile[ASM_T][ASM_T + 1];
    // packed enumeration of the lower-triangular tiles: blockIdx.x = ti (ti + 1) / 2 + tj
    const int t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    const int tj = t - ti * (ti + 1) / 2;
    const int s = blockIdx.y;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* h = hyp + (int64_t)s * stride;
    const double* xs = Xs + (int64_t)s * d.dim_m * n;
    const int dm = (DM > 0) ? DM : d.dim_m;
    constexpr bool tasks = TASKS;   // d.kind == BQO_KIND_PRODUCT_TASKS, resolved at compile time
    for (int e = threadIdx.x; e < ASM_T * dm; e += blockDim.x) {
        const int l = e / ASM_T, r = e - l * ASM_T;
        const int gi = ti * ASM_T + r;
        s_row[r][l] = (gi < n) ? xs[(int64_t)l * n + gi] : 0.0;
    }
    if (tasks && threadIdx.x < ASM_T) {
        const int gi = ti * ASM_T + threadIdx.x;
        s_tid[threadIdx.x] = (gi < n) ? tid[gi] : 0;
    }
    const int j0 = tj * ASM_T + lane, j1 = j0 + 32;
    double xj0[DMS], xj1[DMS];
#pragma unroll
    for (int l = 0; l < DMS; ++l) {
        xj0[l] = (l < dm && j0 < n) ? xs[(int64_t)l * n + j0] : 0.0;
        xj1[l] = (l < dm && j1 < n) ? xs[(int64_t)l * n + j1] : 0.0;
    }
    const int tj0 = (tasks && j0 < n) ? tid[j0] : 0, tj1 = (tasks && j1 < n) ? tid[j1] : 0;
    // ScaledKernel's sigma^2 as a plain factor (1 for the other kinds): no select per element
    const double kscale = (d.kind == BQO_KIND_SCALED_MATERN52) ? h[BQO_H_SIGMA2] : 1.0;
    const bool mirror = !lower_only && ti != tj;
    double* Ks = K + (int64_t)s * n * n;
    __syncthreads();
    const int r0 = w * (ASM_T / 8);
    double* out = Ks + (int64_t)(ti * ASM_T + r0) * n + j0;   // this warp's first row, column j0
    const bool in0 = j0 < n, in1 = j1 < n;
    for (int ii = 0; ii < ASM_T / 8; ++ii, out += n) {
        const int r = r0 + ii;
        const int i = ti * ASM_T + r;
        if (i >= n) break;  // warp-uniform
        double u2v[2] = {0.0, 0.0}, ev[2], uv[2], k[2];
#pragma unroll
        for (int l = 0; l < DMS; ++l) {
            if (l < dm) {
                const double xi = s_row[r][l];
                const double d0 = xi - xj0[l], d1 = xi - xj1[l];
                u2v[0] = fma(d0, d0, u2v[0]);
                u2v[1] = fma(d1, d1, u2v[1]);
            }
        }
        u2v[0] = fma(u2v[0], BQO_PRESCALE * BQO_PRESCALE, 1e-300);
        u2v[1] = fma(u2v[1], BQO_PRESCALE * BQO_PRESCALE, 1e-300);
        bqo_matern52_q_v<2>(u2v, GlobalExpTab(), ev, uv);
        // k = (1 + u c + u2 c^2/3) e with c = ln2/N
        k[0] = fma(u2v[0], BQO_H_TO_K, fma(uv[0], BQO_LN2_N, 1.0)) * ev[0] * kscale;
        k[1] = fma(u2v[1], BQO_H_TO_K, fma(uv[1], BQO_LN2_N, 1.0)) * ev[1] * kscale;
        if (tasks) {
            const int tr = s_tid[r];
            k[0] *= bqo_task_cov(h, d.n_tasks, tr, tj0);
            k[1] *= bqo_task_cov(h, d.n_tasks, tr, tj1);
        }
        if (ti == tj && (r == lane || r == lane + 32)) {
            // the diagonal entry of this row (one lane of a diagonal tile): k(x, x) = 1 exactly
            double kd = kscale;
            if (tasks) kd *= bqo_task_cov(h, d.n_tasks, s_tid[r], s_tid[r]);
            if (add_noise) {
                if (noise_obs) kd += noise_obs[i];
                kd += h[BQO_H_VARNOISE];
            }
            if (extra_diag) kd += extra_diag[s];
            if (r == lane) k[0] = kd; else k[1] = kd;
        }
        if (in0) out[0] = k[0];
        if (in1) out[32] = k[1];
        if (mirror) {
            s_tile[r][lane] = k[0];
            s_tile[r][lane + 32] = k[1];
        }
    }
    if (mirror) {
        __syncthreads();
        for (int cc = 0; cc < ASM_T / 8; ++cc) {
            const int c = w * (ASM_T / 8) + cc;
            const int gj = tj * ASM_T + c;
            if (gj >= n) break;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const int r = lane + 32 * hh;
                const int gi = ti * ASM_T + r;
                if (gi < n) Ks[(int64_t)gj * n + gi] = s_tile[r][c];
            }
        }
    }
}

int bqo_internal_assemble(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                          const double* Xs, const int32_t* tid, int32_t n, const double* noise_obs,
                          int32_t add_noise, const double* extra_diag, double* K, int lower_only,
                          void* stream);

extern "C" int bqo_kernel_assemble_batched(const bqo_kernel_desc* desc, const double* hyp,
                                           int32_t S, const double* Xs, const int32_t* tid,
                                           int32_t n, const double* noise_obs, int32_t add_noise,
                                           const double* extra_diag, double* K, void* stream) {
    return bqo_internal_assemble(desc, hyp, S, Xs, tid, n, noise_obs, add_noise, extra_diag, K, 0,
                                 stream);
}

int bqo_internal_assemble(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                          const double* Xs, const int32_t* tid, int32_t n, const double* noise_obs,
                          int32_t add_noise, const double* extra_diag, double* K, int lower_only,
                          void* stream) {
    BQO_CHECK_ARG(check_desc(desc), 1);
    BQO_CHECK_ARG(hyp != nullptr, 2);
    BQO_CHECK_ARG(S > 0, 3);
    BQO_CHECK_ARG(Xs != nullptr, 4);
    BQO_CHECK_ARG(desc->kind != BQO_KIND_PRODUCT_TASKS || tid != nullptr, 5);
    BQO_CHECK_ARG(n > 0, 6);
    BQO_CHECK_ARG(K != nullptr, 10);
    const int nt = (n + ASM_T - 1) / ASM_T;
    dim3 block(256);
    dim3 grid(nt * (nt + 1) / 2, S);
    int64_t st = bqo_hyper_stride_impl(*desc);
    cudaStream_t cs = (cudaStream_t)stream;
    bqo_ensure_global_exp_table(cs);
#define LAUNCH_ASM(DMV)                                                                          \
    do {                                                                                         \
        if (desc->kind == BQO_KIND_PRODUCT_TASKS)                                                \
            BQO_L, assemble_kernel<DMV, true><<<grid, block, 0, cs>>>(                           \
                *desc, hyp, st, Xs, tid, n, noise_obs, add_noise, extra_diag, K, lower_only);    \
        else                                                                                     \
            BQO_L, assemble_kernel<DMV, false><<<grid, block, 0, cs>>>(                          \
                *desc, hyp, st, Xs, tid, n, noise_obs, add_noise, extra_diag, K, lower_only);    \
    } while (0)
    switch (desc->dim_m) {
        case 1: LAUNCH_ASM(1); break;
        case 2: LAUNCH_ASM(2); break;
        case 3: LAUNCH_ASM(3); break;
        case 4: LAUNCH_ASM(4); break;
        case 6: LAUNCH_ASM(6); break;
        default: LAUNCH_ASM(0); break;
    }
#undef LAUNCH_ASM
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// Derivative of the task covariance with respect to task parameter m, entry (u, v)
// (GradientTasksKernel.gradient_respect_parameters, sbo/kernels/tasks_kernel.py:580-622).
__device__ __forceinline__ double bqo_task_dcov(const bqo_kernel_desc& d,
                                                const double* __restrict__ h, int m, int u, int v) {
    const int T = d.n_tasks;
    const double* C = h + BQO_H_TASKC;
    const double* Lt = C + T * T;
    if (d.same_corr) {
        if (T == 1) return (u == v) ? C[0] : 0.0;
        double value = C[1];  // chol_base_cov_matrix[0, 1]
        if (m == 0) return (u == v) ? (C[0] - value * (double)(T - 1)) : 0.0;
        return (u == v) ? value * (double)(T - 1) : C[u * T + v];
    }
    // m <-> (a, b), b <= a: tmp_der[a, b] = L[a, b]; M = tmp_der L^T; grad = M + M^T
    int a = (int)floor((sqrt(8.0 * (double)m + 1.0) - 1.0) * 0.5);
    while ((a + 1) * (a + 2) / 2 <= m) ++a;
    while (a * (a + 1) / 2 > m) --a;
    int b = m - a * (a + 1) / 2;
    double lab = Lt[a * T + b];
    double g = 0.0;
    if (u == a) g += lab * Lt[v * T + b];
    if (v == a) g += lab * Lt[u * T + b];
    return g;
}

// dK[s][p][i][j]
__global__ void grad_params_kernel(bqo_kernel_desc d, const double* __restrict__ hyp,
                                   int64_t stride, const double* __restrict__ Xs,
                                   const int32_t* __restrict__ tid, int n,
                                   double* __restrict__ dK) {
    int j = blockIdx.x * 32 + threadIdx.x;
    int i = blockIdx.y * 8 + threadIdx.y;
    int s = blockIdx.z;
    if (i >= n || j >= n) return;
    const double* h = hyp + (int64_t)s * stride;
    const double* xs = Xs + (int64_t)s * d.dim_m * n;
    double df[BQO_MAX_DIM];
    double r2 = 0.0;
    for (int l = 0; l < d.dim_m; ++l) {
        df[l] = xs[(int64_t)l * n + i] - xs[(int64_t)l * n + j];
        r2 = fma(df[l], df[l], r2);
    }
    double km, g;
    bqo_matern52_kg(r2, km, g);
    double other = 1.0;  // factor multiplying the Matern derivative
    if (d.kind == BQO_KIND_SCALED_MATERN52) other = h[BQO_H_SIGMA2];
    int ti = 0, tj = 0;
    if (d.kind == BQO_KIND_PRODUCT_TASKS) {
        ti = tid[i];
        tj = tid[j];
        other = bqo_task_cov(h, d.n_tasks, ti, tj);
    }
    double* base = dK + ((int64_t)s * d.n_params) * n * n + (int64_t)i * n + j;
    // d k_M / d ls_l = k'(r) (1/r) delta_l^2 (-1/ls_l^3) = -g(r) (delta_l/ls_l)^2 / ls_l,
    // diagonal forced to 0 (sbo/lib/distances.py:53-63, sbo/kernels/matern52.py:386-392).
    for (int l = 0; l < d.dim_m; ++l) {
        double v = (i == j) ? 0.0 : (-g * df[l] * df[l] * h[BQO_H_INVLS + l]) * other;
        base[(int64_t)l * n * n] = v;
    }
    if (d.kind == BQO_KIND_SCALED_MATERN52) {
        // d(sigma2 k)/d sigma2 = cov / sigma2 (sbo/kernels/scaled_kernel.py:212)
        base[(int64_t)d.dim_m * n * n] = (km * h[BQO_H_SIGMA2]) / h[BQO_H_SIGMA2];
    } else if (d.kind == BQO_KIND_PRODUCT_TASKS) {
        int pt = d.n_params - d.dim_m;
        for (int m = 0; m < pt; ++m)
            base[(int64_t)(d.dim_m + m) * n * n] = km * bqo_task_dcov(d, h, m, ti, tj);
    }
}

extern "C" int bqo_kernel_grad_params_batched(const bqo_kernel_desc* desc, const double* hyp,
                                              int32_t S, const double* Xs, const int32_t* tid,
                                              int32_t n, double* dK, void* stream) {
    BQO_CHECK_ARG(check_desc(desc), 1);
    BQO_CHECK_ARG(hyp != nullptr, 2);
    BQO_CHECK_ARG(S > 0, 3);
    BQO_CHECK_ARG(Xs != nullptr, 4);
    BQO_CHECK_ARG(desc->kind != BQO_KIND_PRODUCT_TASKS || tid != nullptr, 5);
    BQO_CHECK_ARG(n > 0, 6);
    BQO_CHECK_ARG(dK != nullptr, 7);
    dim3 block(32, 8);
    dim3 grid((n + 31) / 32, (n + 7) / 8, S);
    BQO_L, grad_params_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
        *desc, hyp, bqo_hyper_stride_impl(*desc), Xs, tid, n, dK);
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// Cross covariance of m raw points against the training set, and its gradient w.r.t. the point.
__global__ void cross_cov_kernel(bqo_kernel_desc d, const double* __restrict__ hyp, int64_t stride,
                                 const double* __restrict__ Xs, const int32_t* __restrict__ tid,
                                 int n, const double* __restrict__ P, int m, int interleave,
                                 double* __restrict__ KP, double* __restrict__ dKP) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int q = blockIdx.y;
    int s = blockIdx.z;
    if (i >= n) return;
    const double* h = hyp + (int64_t)s * stride;
    const double* xs = Xs + (int64_t)s * d.dim_m * n;
    const double* p = P + (int64_t)q * d.dim;
    double df[BQO_MAX_DIM];
    double r2 = 0.0;
    for (int l = 0; l < d.dim_m; ++l) {
        df[l] = p[l] / h[BQO_H_LS + l] - xs[(int64_t)l * n + i];
        r2 = fma(df[l], df[l], r2);
    }
    double k, g;
    bqo_matern52_kg(r2, k, g);
    double other = 1.0;
    if (d.kind == BQO_KIND_SCALED_MATERN52) other = h[BQO_H_SIGMA2];
    if (d.kind == BQO_KIND_PRODUCT_TASKS)
        other = bqo_task_cov(h, d.n_tasks, (int)p[d.dim - 1], tid[i]);
    const int rows = 1 + d.dim_m;
    if (interleave) {
        double* base = KP + (((int64_t)s * m + q) * rows) * n + i;
        base[0] = k * other;
        for (int l = 0; l < d.dim_m; ++l)
            base[(int64_t)(1 + l) * n] = g * df[l] * h[BQO_H_INVLS + l] * other;
    } else {
        KP[((int64_t)s * m + q) * n + i] = k * other;
        if (dKP) {
            double* base = dKP + (((int64_t)s * m + q) * d.dim_m) * n + i;
            for (int l = 0; l < d.dim_m; ++l)
                base[(int64_t)l * n] = g * df[l] * h[BQO_H_INVLS + l] * other;
        }
    }
}

extern "C" int bqo_cross_cov_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                                     const double* Xs, const int32_t* tid, int32_t n,
                                     const double* P, int32_t m, int32_t interleave, double* KP,
                                     double* dKP, void* stream) {
    BQO_CHECK_ARG(check_desc(desc), 1);
    BQO_CHECK_ARG(hyp != nullptr, 2);
    BQO_CHECK_ARG(S > 0, 3);
    BQO_CHECK_ARG(Xs != nullptr, 4);
    BQO_CHECK_ARG(desc->kind != BQO_KIND_PRODUCT_TASKS || tid != nullptr, 5);
    BQO_CHECK_ARG(n > 0, 6);
    BQO_CHECK_ARG(P != nullptr, 7);
    BQO_CHECK_ARG(m > 0 && m <= 65535, 8);
    BQO_CHECK_ARG(KP != nullptr, 10);
    BQO_CHECK_ARG(S <= 65535, 3);
    dim3 grid((n + 255) / 256, m, S);
    BQO_L, cross_cov_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        *desc, hyp, bqo_hyper_stride_impl(*desc), Xs, tid, n, P, m, interleave, KP, dKP);
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// qt[s][t'] = sum_t w_t C_s[t][t'] / sum_t w_t  (np.average over the task axis,
// sbo/lib/expectations.py:73-74).
__global__ void task_qweights_kernel(bqo_kernel_desc d, const double* __restrict__ hyp,
                                     int64_t stride, const double* __restrict__ weights,
                                     double* __restrict__ qt) {
    int s = blockIdx.x;
    const int T = d.n_tasks;
    const double* C = hyp + (int64_t)s * stride + BQO_H_TASKC;
    double wsum = 0.0;
    for (int t = 0; t < T; ++t) wsum += weights ? weights[t] : 1.0;
    for (int tp = threadIdx.x; tp < T; tp += blockDim.x) {
        double acc = 0.0;
        for (int t = 0; t < T; ++t) acc += (weights ? weights[t] : 1.0) * C[t * T + tp];
        qt[(int64_t)s * T + tp] = acc / wsum;
    }
}

extern "C" int bqo_task_quadrature_weights(const bqo_kernel_desc* desc, const double* hyp,
                                           int32_t S, const double* weights, double* qt,
                                           void* stream) {
    BQO_CHECK_ARG(check_desc(desc) && desc->kind == BQO_KIND_PRODUCT_TASKS, 1);
    BQO_CHECK_ARG(hyp != nullptr, 2);
    BQO_CHECK_ARG(S > 0, 3);
    BQO_CHECK_ARG(qt != nullptr, 5);
    BQO_L, task_qweights_kernel<<<S, 128, 0, (cudaStream_t)stream>>>(
        *desc, hyp, bqo_hyper_stride_impl(*desc), weights, qt);
    BQO_RETURN_LAST_ERROR();
}
