// Second derivatives with respect to the query point: used only by
// the non-default Newton / dogleg inner optimisers (`method_opt_mc`), so these are plain kernels
// with libm sqrt / exp -- no staging, no tables.
//
// Matern-5/2 in r^2 = sum_l u_l^2, u_l = (x_l - X_l) / ls_l:
//   k(r)              = (1 + sqrt5 r + 5/3 r^2) e^{-sqrt5 r}
//   k'(r) / r         = -5/3 (1 + sqrt5 r) e^{-sqrt5 r}                       =: g(r)
//   (k'(r)/r)' / r    = 25/3 e^{-sqrt5 r}                                      =: h(r)
//   d2k / dx_l dx_m   = g(r) delta_lm / ls_l^2 + h(r) u_l u_m / (ls_l ls_m)
// which is sbo/kernels/matern52.py:466-503 (hess(r) k'(r) + grad r grad r^T k''(r), with
// sbo/lib/distances.py:70-116) after the 1/r and 1/r^3 terms cancel -- finite at r = 0, where the
// reference returns NaN.
#include "bqo_common.cuh"

#define SQRT5 2.23606797749978969641

__device__ __forceinline__ void matern52_gh(double r2, double& g, double& h) {
    const double r = sqrt(r2);
    const double e = exp(-SQRT5 * r);
    g = -(5.0 / 3.0) * (1.0 + SQRT5 * r) * e;
    h = (25.0 / 3.0) * e;
}

// out[s][q][l][m][i] = d2 k(P_q, X_i) / dP_l dP_m, l, m < dim_m (the Matern coordinates; the task
// coordinate of a product kernel has no derivative, sbo/kernels/product_kernels.py:415-439).
__global__ void cross_cov_hessian_kernel(bqo_kernel_desc d, const double* __restrict__ hyp,
                                         int64_t stride, const double* __restrict__ Xs,
                                         const int32_t* __restrict__ tid, int n,
                                         const double* __restrict__ P, int m,
                                         double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y, s = blockIdx.z;
    if (i >= n) return;
    const double* hs = hyp + (int64_t)s * stride;
    const double* xs = Xs + (int64_t)s * d.dim_m * n;
    const double* p = P + (int64_t)q * d.dim;
    const int dm = d.dim_m;
    double u[BQO_MAX_DIM];
    double r2 = 0.0;
    for (int l = 0; l < dm; ++l) {
        u[l] = p[l] / hs[BQO_H_LS + l] - xs[(int64_t)l * n + i];
        r2 = fma(u[l], u[l], r2);
    }
    double g, h;
    matern52_gh(r2, g, h);
    double other = 1.0;
    if (d.kind == BQO_KIND_SCALED_MATERN52) other = hs[BQO_H_SIGMA2];
    if (d.kind == BQO_KIND_PRODUCT_TASKS)
        other = bqo_task_cov(hs, d.n_tasks, (int)p[d.dim - 1], tid[i]);
    double* base = out + (((int64_t)s * m + q) * dm * dm) * n + i;
    for (int l = 0; l < dm; ++l) {
        const double il = hs[BQO_H_INVLS + l];
        for (int c = 0; c < dm; ++c) {
            const double ic = hs[BQO_H_INVLS + c];
            double v = h * u[l] * u[c] * il * ic;
            if (l == c) v = fma(g, il * il, v);
            base[(int64_t)(l * dm + c) * n] = v * other;
        }
    }
}

extern "C" int bqo_cross_cov_hessian_batched(const bqo_kernel_desc* desc, const double* hyp,
                                             int32_t S, const double* Xs, const int32_t* tid,
                                             int32_t n, const double* P, int32_t m, double* out,
                                             void* stream) {
    BQO_CHECK_ARG(desc != nullptr && desc->dim_m > 0 && desc->dim_m <= BQO_MAX_DIM, 1);
    BQO_CHECK_ARG(hyp != nullptr, 2);
    BQO_CHECK_ARG(S > 0 && S <= 65535, 3);
    BQO_CHECK_ARG(Xs != nullptr, 4);
    BQO_CHECK_ARG(desc->kind != BQO_KIND_PRODUCT_TASKS || tid != nullptr, 5);
    BQO_CHECK_ARG(n > 0, 6);
    BQO_CHECK_ARG(P != nullptr, 7);
    BQO_CHECK_ARG(m > 0 && m <= 65535, 8);
    BQO_CHECK_ARG(out != nullptr, 9);
    dim3 grid((n + 255) / 256, m, S);
    BQO_L, cross_cov_hessian_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        *desc, hyp, bqo_hyper_stride_impl(*desc), Xs, tid, n, P, m, out);
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// Hessians of a(x) and b(x) of one unit at P points
// (compute_hessian_parameters_for_sample, sbo/numerical_tools/bayesian_quadrature.py:1241-1286):
//   hess a = sum_j hess_x B(x, X_j) alpha_j
//   hess b = (hess_x B(x, c) - sum_j hess_x B(x, X_j) (Sigma^-1 gamma)_j) / den     (0 if den = 0)
// with B the expectation over W of the kernel (finite W: the task sum is already folded into the
// q-weighted alpha and Sigma^-1 gamma rows; Monte-Carlo W: mean over the M samples of the set).
// One CTA per point; out[q] = [hess a (dx x dx), hess b (dx x dx)].
#define HESS_THREADS 256
#define HESS_MAX_PAIRS (BQO_MAX_DIM * (BQO_MAX_DIM + 1) / 2)

__global__ void __launch_bounds__(HESS_THREADS)
    sample_ab_hess_kernel(bqo_stage_c pb, const double* __restrict__ pts,
                          const int32_t* __restrict__ pt_unit, const int32_t* __restrict__ pt_wset,
                          double* __restrict__ out) {
    __shared__ double red[32];
    __shared__ double wsc[64 * 4];   // scaled W samples of this point's set: M <= 64, dw <= 4
    const int q = blockIdx.x;
    const int unit = pt_unit[q];
    const int k = pb.unit_k[unit], c = pb.unit_c[unit];
    const int dx = pb.dx, dw = pb.dw, dm = pb.desc.dim_m, n = pb.n;
    const int64_t stride = bqo_hyper_stride_impl(pb.desc);
    const double* hs = pb.hyp + (int64_t)k * stride;
    const double* xs = pb.Xs + (int64_t)k * dm * n;
    const double* wa = pb.alpha + (int64_t)k * n;
    const double* wb = pb.R + (((int64_t)k * pb.C + c) * (1 + dm)) * n;
    const int M = dw > 0 ? pb.M : 1;
    int wset = 0;
    if (dw > 0) wset = pt_wset ? pt_wset[q] : (c % pb.n_wsets);
    for (int e = threadIdx.x; e < M * dw; e += HESS_THREADS)
        wsc[e] = pb.wsets[(int64_t)wset * M * dw + e] / hs[BQO_H_LS + dx + (e % dw)];
    __syncthreads();
    double xsx[BQO_MAX_DIM];
    for (int l = 0; l < dx; ++l) xsx[l] = pts[(int64_t)q * dx + l] / hs[BQO_H_LS + l];
    const int npair = dx * (dx + 1) / 2;
    double Ha[HESS_MAX_PAIRS], Hb[HESS_MAX_PAIRS];
    for (int e = 0; e < npair; ++e) Ha[e] = Hb[e] = 0.0;
    // accumulate the Hessian of one kernel row sum_m k((x, w_m), z) into `acc` with weight w
    auto add_row = [&](const double* u, double d0, const double* zw, double w, double* acc, double w2,
                       double* acc2) {
        double G = 0.0, Hh = 0.0;
        for (int mm = 0; mm < M; ++mm) {
            double r2 = d0;
            for (int l = 0; l < dw; ++l) {
                const double dd = wsc[mm * dw + l] - zw[l];
                r2 = fma(dd, dd, r2);
            }
            double g, h;
            matern52_gh(r2, g, h);
            G += g;
            Hh += h;
        }
        int e = 0;
        for (int l = 0; l < dx; ++l)
            for (int cc = 0; cc <= l; ++cc, ++e) {
                double v = Hh * u[l] * u[cc];
                if (l == cc) v += G;
                acc[e] = fma(v, w, acc[e]);
                if (acc2) acc2[e] = fma(v, w2, acc2[e]);
            }
    };
    for (int j = threadIdx.x; j < n; j += HESS_THREADS) {
        double u[BQO_MAX_DIM], zw[4];
        double d0 = 0.0;
        for (int l = 0; l < dx; ++l) {
            u[l] = xsx[l] - xs[(int64_t)l * n + j];
            d0 = fma(u[l], u[l], d0);
        }
        for (int l = 0; l < dw; ++l) zw[l] = xs[(int64_t)(dx + l) * n + j];
        add_row(u, d0, zw, wa[j], Ha, wb[j], Hb);
    }
    // candidate row B(x, c): thread 0 adds it to a separate accumulator
    double Hc[HESS_MAX_PAIRS];
    for (int e = 0; e < npair; ++e) Hc[e] = 0.0;
    if (threadIdx.x == 0) {
        double u[BQO_MAX_DIM], zw[4];
        double d0 = 0.0;
        for (int l = 0; l < dx; ++l) {
            u[l] = xsx[l] - pb.cand[(int64_t)c * pb.desc.dim + l] / hs[BQO_H_LS + l];
            d0 = fma(u[l], u[l], d0);
        }
        for (int l = 0; l < dw; ++l)
            zw[l] = pb.cand[(int64_t)c * pb.desc.dim + dx + l] / hs[BQO_H_LS + dx + l];
        add_row(u, d0, zw, 1.0, Hc, 0.0, nullptr);
    }
    double scale = 1.0, wc = 1.0;
    if (pb.desc.kind == BQO_KIND_SCALED_MATERN52) scale = hs[BQO_H_SIGMA2];
    if (dw > 0) scale /= (double)M;
    if (pb.desc.kind == BQO_KIND_PRODUCT_TASKS) {
        const int tc = (int)pb.cand[(int64_t)c * pb.desc.dim + pb.desc.dim - 1];
        wc = pb.qt[(int64_t)k * pb.desc.n_tasks + tc];
    }
    const double den = pb.den[(int64_t)k * pb.C + c];
    const double inv_den = (den != 0.0) ? 1.0 / den : 0.0;
    double* o = out + (int64_t)q * 2 * dx * dx;
    int e = 0;
    for (int l = 0; l < dx; ++l)
        for (int cc = 0; cc <= l; ++cc, ++e) {
            const double sa = bqo_block_sum(Ha[e], red);
            const double sb = bqo_block_sum(Hb[e], red);
            if (threadIdx.x == 0) {
                const double il = hs[BQO_H_INVLS + l] * hs[BQO_H_INVLS + cc];
                const double ha = scale * sa * il;
                const double hb = scale * (wc * Hc[e] - sb) * il * inv_den;
                o[l * dx + cc] = o[cc * dx + l] = ha;
                o[dx * dx + l * dx + cc] = o[dx * dx + cc * dx + l] = hb;
            }
        }
}

extern "C" int bqo_sample_ab_hess_batched(const bqo_stage_c* pb, const double* pts,
                                          const int32_t* pt_unit, const int32_t* pt_wset, int32_t P,
                                          double* out, void* stream) {
    BQO_CHECK_ARG(pb != nullptr && pb->n > 0 && pb->dx > 0 && pb->dw >= 0 &&
                      pb->dx + pb->dw == pb->desc.dim_m && pb->desc.dim_m <= BQO_MAX_DIM, 1);
    BQO_CHECK_ARG(pb->dw <= 4 && (pb->dw == 0 || (pb->M > 0 && pb->M <= 64 && pb->wsets != nullptr)), 1);
    BQO_CHECK_ARG(pb->hyp && pb->Xs && pb->alpha && pb->R && pb->den && pb->cand && pb->unit_k &&
                      pb->unit_c, 1);
    BQO_CHECK_ARG(pb->desc.kind != BQO_KIND_PRODUCT_TASKS || (pb->qt && pb->dw == 0), 1);
    BQO_CHECK_ARG(pts != nullptr, 2);
    BQO_CHECK_ARG(pt_unit != nullptr, 3);
    BQO_CHECK_ARG(P > 0, 5);
    BQO_CHECK_ARG(out != nullptr, 6);
    BQO_L, sample_ab_hess_kernel<<<P, HESS_THREADS, 0, (cudaStream_t)stream>>>(*pb, pts, pt_unit,
                                                                         pt_wset, out);
    BQO_RETURN_LAST_ERROR();
}
