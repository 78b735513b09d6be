// Gradient of the log marginal likelihood for S hyper-samples at once.
//
// Reference: grad_log_likelihood_dict + GradientGPFittingGaussian
// (sbo/models/gp_fitting_gaussian.py:832-897, 1654-1697):
//     d llh / d theta_j = 0.5 trace(alpha alpha^T dK_j - Sigma^-1 dK_j)
// which the reference evaluates with one n x n dpotrs and one n x n x n product PER parameter.
// Here Sigma^-1 is formed once per hyper-sample (potrs on the identity) and every parameter
// derivative is the Frobenius product <alpha alpha^T - Sigma^-1, dK_j>, with dK_j generated on the
// fly in registers from the same pairwise distance (never stored), over the lower triangle of
// tiles only.  Partial sums are written per CTA and added in a fixed order, so the result is
// bit-reproducible.
#include "bqo_gemm.cuh"

int bqo_internal_inverse_from_chol(const double* L, int n, int S, double* Inv, double* Wbuf,
                                   cudaStream_t cs);
int bqo_internal_tri_inverse(const double* L, int n, int S, double* Wbuf, double* scratch,
                             cudaStream_t cs);
int bqo_internal_inv_syrk(const double* Wbuf, int n, int S, double* Inv, cudaStream_t cs);

#define LG_MAX_ACC (BQO_MAX_DIM + 4)
// accumulator slots: [0..dim_m) length scales, then
#define LG_SLOT_KM(dm) (dm)        // sum M_ij k_M(i,j) other-free  (sigma2 derivative)
#define LG_SLOT_SAME(dm) ((dm) + 1) // sum over pairs with equal tasks of M_ij k_M
#define LG_SLOT_DIFF(dm) ((dm) + 2) // sum over pairs with different tasks of M_ij k_M
#define LG_SLOT_TRACE(dm) ((dm) + 3) // sum_i M_ii

__global__ void __launch_bounds__(256)
    loglik_grad_partial_kernel(bqo_kernel_desc d, const double* __restrict__ hyp, int64_t stride,
                               const double* __restrict__ Xs, const int32_t* __restrict__ tid,
                               int n, const double* __restrict__ Inv,
                               const double* __restrict__ alpha, double* __restrict__ partial,
                               double* __restrict__ Wtask) {
    __shared__ double red[32];
    const int s = blockIdx.z;
    const int by = (int)gridDim.y - 1 - (int)blockIdx.y;  // longest rows of tiles first
    const int i0 = by * 32 + (threadIdx.x >> 5) * 4;
    const double* h = hyp + (int64_t)s * stride;
    const double* xs = Xs + (int64_t)s * d.dim_m * n;
    const double* inv = Inv + (int64_t)s * n * n;
    const double* al = alpha + (int64_t)s * n;
    const int dm = d.dim_m;
    const int nacc = dm + 4;
    double acc[LG_MAX_ACC];
    for (int q = 0; q < nacc; ++q) acc[q] = 0.0;
    const bool general_tasks = (d.kind == BQO_KIND_PRODUCT_TASKS) && !d.same_corr;
    // M and every dK_j are symmetric: only tiles on or below the diagonal are visited, the
    // strictly lower ones with weight 2 (the task-pair sums are used symmetrised downstream).
    // Inv is only defined on those tiles (bqo_internal_inverse_from_chol stores 64x64 lower
    // tiles; a 32x32 tile on or below the diagonal always lies inside one of them).
    // A CTA walks one row of tiles, so the eleven block reductions below are paid once per
    // row of tiles and not once per tile (they dominated the kernel).
    for (int bx = 0; bx <= by; ++bx) {
    const int j = bx * 32 + (threadIdx.x & 31);
    const double wsym = (bx < by) ? 2.0 : 1.0;
    if (j < n) {
        double xj[BQO_MAX_DIM];
        for (int l = 0; l < dm; ++l) xj[l] = xs[(int64_t)l * n + j];
        const double aj = al[j];
        const int tj = (d.kind == BQO_KIND_PRODUCT_TASKS) ? tid[j] : 0;
        for (int ii = 0; ii < 4; ++ii) {
            const int i = i0 + ii;
            if (i >= n) break;
            const double Mij = wsym * (al[i] * aj - inv[(int64_t)i * n + j]);
            double df[BQO_MAX_DIM];
            double r2 = 0.0;
            for (int l = 0; l < dm; ++l) {
                df[l] = xs[(int64_t)l * n + i] - xj[l];
                r2 = fma(df[l], df[l], r2);
            }
            // fast evaluation (bqo_common.cuh): e = exp(-sqrt5 r), u = K r
            double km = 1.0, g = -BQO_FIVE_THIRDS;
            if (i != j) {
                double u2v[1], ev[1], uv[1];
                u2v[0] = fma(r2, BQO_PRESCALE * BQO_PRESCALE, 1e-300);
                bqo_matern52_q_v<1>(u2v, GlobalExpTab(), ev, uv);
                const double lin = fma(uv[0], BQO_LN2_N, 1.0) * ev[0];  // (1 + sqrt5 r) e
                km = fma(u2v[0] * ev[0], BQO_H_TO_K, lin);
                g = -BQO_FIVE_THIRDS * lin;                             // k'(r) / r
            }
            double other = 1.0;
            int ti = 0;
            if (d.kind == BQO_KIND_SCALED_MATERN52) other = h[BQO_H_SIGMA2];
            if (d.kind == BQO_KIND_PRODUCT_TASKS) {
                ti = tid[i];
                other = bqo_task_cov(h, d.n_tasks, ti, tj);
            }
            if (i != j) {
                const double w = -Mij * g * other;
                for (int l = 0; l < dm; ++l)
                    acc[l] = fma(w * df[l] * df[l], h[BQO_H_INVLS + l], acc[l]);
            } else {
                acc[LG_SLOT_TRACE(dm)] += Mij;
            }
            const double mk = Mij * km;
            acc[LG_SLOT_KM(dm)] += mk;
            if (d.kind == BQO_KIND_PRODUCT_TASKS) {
                if (general_tasks) atomicAdd(Wtask + ((int64_t)s * d.n_tasks + ti) * d.n_tasks + tj, mk);
                else if (ti == tj) acc[LG_SLOT_SAME(dm)] += mk;
                else acc[LG_SLOT_DIFF(dm)] += mk;
            }
        }
    }
    }  // row of tiles
    const int nblk = gridDim.y;
    const int blk = by;
    for (int q = 0; q < nacc; ++q) {
        double tot = bqo_block_sum(acc[q], red);
        if (threadIdx.x == 0) partial[((int64_t)s * nacc + q) * nblk + blk] = tot;
    }
}



// Contraction of one 64 x 64 tile of Sigma^-1 (in the accumulator registers) with M = alpha alpha^T
// - Sigma^-1 against every dK_j.  DMT = dim_m at compile time (0: run-time value, local-memory
// arrays): with a run-time loop bound the per-pair arrays live in local memory and the epilogue
// costs as much as the tile product itself.
struct LgTile {
    const double *xi, *xj, *ai, *aj;
    const int *ti, *tj;
    int a0, b0, ra, rb, s;
    double wsym;
};

template <int DMT>
__device__ __forceinline__ void lg_contract_tile(BqoAcc& acc, const bqo_kernel_desc& d,
                                                 const double* __restrict__ h, const LgTile& tl,
                                                 double* __restrict__ Wtask, double* __restrict__ out,
                                                 int ntiles, double* red) {
    constexpr int DMA = (DMT > 0) ? DMT : BQO_MAX_DIM;
    const int dm = (DMT > 0) ? DMT : d.dim_m;
    const bool prod = d.kind == BQO_KIND_PRODUCT_TASKS;
    const bool general_tasks = prod && !d.same_corr;
    const double sigma2 = (d.kind == BQO_KIND_SCALED_MATERN52) ? h[BQO_H_SIGMA2] : 1.0;
    double invls[DMA], sums[DMA], s_km = 0.0, s_same = 0.0, s_diff = 0.0, s_trace = 0.0;
#pragma unroll
    for (int l = 0; l < DMA; ++l) {
        invls[l] = (l < dm) ? h[BQO_H_INVLS + l] : 0.0;
        sums[l] = 0.0;
    }
    bqo_acc_foreach(acc, [&](int r, int c, double& v) {
        if (r >= tl.ra || c >= tl.rb) return;
        const int i = tl.a0 + r, j = tl.b0 + c;
        const double Mij = tl.wsym * (tl.ai[r] * tl.aj[c] - v);
        double df[DMA];
        double r2 = 0.0;
#pragma unroll
        for (int l = 0; l < DMA; ++l) {
            df[l] = (l < dm) ? tl.xi[l * BQO_TILE + r] - tl.xj[l * BQO_TILE + c] : 0.0;
            r2 = fma(df[l], df[l], r2);
        }
        double km = 1.0, g = -BQO_FIVE_THIRDS;
        if (i != j) {
            double u2v[1], ev[1], uv[1];
            u2v[0] = fma(r2, BQO_PRESCALE * BQO_PRESCALE, 1e-300);
            bqo_matern52_q_v<1>(u2v, GlobalExpTab(), ev, uv);
            const double lin = fma(uv[0], BQO_LN2_N, 1.0) * ev[0];  // (1 + sqrt5 r) e
            km = fma(u2v[0] * ev[0], BQO_H_TO_K, lin);
            g = -BQO_FIVE_THIRDS * lin;                             // k'(r) / r
        }
        double other = sigma2;
        if (prod) other = bqo_task_cov(h, d.n_tasks, tl.ti[r], tl.tj[c]);
        if (i != j) {
            const double w = -Mij * g * other;
#pragma unroll
            for (int l = 0; l < DMA; ++l) sums[l] = fma(w * df[l] * df[l], invls[l], sums[l]);
        } else {
            s_trace += Mij;
        }
        const double mk = Mij * km;
        s_km += mk;
        if (prod) {
            if (general_tasks)
                atomicAdd(Wtask + ((int64_t)tl.s * d.n_tasks + tl.ti[r]) * d.n_tasks + tl.tj[c], mk);
            else if (tl.ti[r] == tl.tj[c]) s_same += mk;
            else s_diff += mk;
        }
    });
    // slots: [0..dm) length scales, then LG_SLOT_KM, LG_SLOT_SAME, LG_SLOT_DIFF, LG_SLOT_TRACE
#pragma unroll
    for (int l = 0; l < DMA; ++l) {
        if (l < dm) {
            const double tot = bqo_block_sum(sums[l], red);
            if (threadIdx.x == 0) out[(int64_t)l * ntiles] = tot;
        }
    }
    const double extra[4] = {s_km, s_same, s_diff, s_trace};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double tot = bqo_block_sum(extra[q], red);
        if (threadIdx.x == 0) out[(int64_t)(dm + q) * ntiles] = tot;
    }
}

// ---------------------------------------------------------------------------------------------
// Sigma^-1 = W^T W and the contraction in ONE kernel: a CTA forms one 64 x 64 lower tile of
// Sigma^-1 on DMMA (K range from the tile's first row: W = L^-1 vanishes above it) and contracts it,
// still in registers, with M = alpha alpha^T - Sigma^-1 against every dK_j generated on the fly.
// Sigma^-1 is never written to HBM (the separate contraction pass read 8 n^2 / 2 bytes per
// hyper-sample back and took 1.0 of the gradient's 5.3 ms at n = 2048, S = 20).
// The per-tile partial sums go to partial[s][slot][tile] and are added in a fixed order.
#define LGF_SMEM_BYTES ((size_t)(2 * BQO_OPER_BUF_DOUBLES(BQO_GEMM_STAGES)) * sizeof(double))
__global__ void __launch_bounds__(128)
    inv_syrk_contract_kernel(bqo_kernel_desc d, const double* __restrict__ hyp, int64_t stride,
                             const double* __restrict__ Xs, const int32_t* __restrict__ tid, int n,
                             const double* __restrict__ W, const double* __restrict__ alpha,
                             double* __restrict__ partial, double* __restrict__ Wtask, int ntiles) {
    extern __shared__ __align__(16) double dyn_smem[];
    double* Asm = dyn_smem;
    double* Bsm = Asm + BQO_OPER_BUF_DOUBLES(BQO_GEMM_STAGES);
    __shared__ double red[32];
    const int s = blockIdx.z;
    const int ba = blockIdx.y, bb = blockIdx.x;
    if (bb > ba) return;
    const double* Wm = W + (int64_t)s * n * n;
    const int a0 = ba * BQO_TILE, b0 = bb * BQO_TILE;
    const int ra = (n - a0 < BQO_TILE) ? n - a0 : BQO_TILE;
    const int rb = (n - b0 < BQO_TILE) ? n - b0 : BQO_TILE;
    BqoAcc acc;
    acc.zero();
    bqo_gemm_tile<false, false>(acc, Wm + (int64_t)a0 * n + a0, n, ra, Wm + (int64_t)a0 * n + b0, n,
                                rb, n - a0, Asm, Bsm);
    __syncthreads();
    // coordinates, alpha and task ids of the tile's 64 rows (i) and 64 columns (j) -> shared memory
    const int dm = d.dim_m;
    const double* h = hyp + (int64_t)s * stride;
    const double* xs = Xs + (int64_t)s * dm * n;
    const double* al = alpha + (int64_t)s * n;
    double* xi = Asm;                      // [dm][64]
    double* xj = xi + dm * BQO_TILE;       // [dm][64]
    double* ai = xj + dm * BQO_TILE;       // [64]
    double* aj = ai + BQO_TILE;            // [64]
    int* ti = (int*)(aj + BQO_TILE);       // [64]
    int* tj = ti + BQO_TILE;               // [64]
    const bool prod = d.kind == BQO_KIND_PRODUCT_TASKS;
    for (int e = threadIdx.x; e < dm * BQO_TILE; e += 128) {
        const int l = e / BQO_TILE, r = e % BQO_TILE;
        xi[e] = (r < ra) ? xs[(int64_t)l * n + a0 + r] : 0.0;
        xj[e] = (r < rb) ? xs[(int64_t)l * n + b0 + r] : 0.0;
    }
    for (int r = threadIdx.x; r < BQO_TILE; r += 128) {
        ai[r] = (r < ra) ? al[a0 + r] : 0.0;
        aj[r] = (r < rb) ? al[b0 + r] : 0.0;
        ti[r] = (prod && r < ra) ? tid[a0 + r] : 0;
        tj[r] = (prod && r < rb) ? tid[b0 + r] : 0;
    }
    __syncthreads();
    const double wsym = (bb < ba) ? 2.0 : 1.0;   // strictly lower tiles stand for their mirror too
    const int tile = ba * (ba + 1) / 2 + bb;
    LgTile tl;
    tl.xi = xi; tl.xj = xj; tl.ai = ai; tl.aj = aj; tl.ti = ti; tl.tj = tj;
    tl.a0 = a0; tl.b0 = b0; tl.ra = ra; tl.rb = rb; tl.wsym = wsym; tl.s = s;
    double* out = partial + (int64_t)s * (dm + 4) * ntiles + tile;
    switch (dm) {
        case 1: lg_contract_tile<1>(acc, d, h, tl, Wtask, out, ntiles, red); break;
        case 2: lg_contract_tile<2>(acc, d, h, tl, Wtask, out, ntiles, red); break;
        case 3: lg_contract_tile<3>(acc, d, h, tl, Wtask, out, ntiles, red); break;
        case 4: lg_contract_tile<4>(acc, d, h, tl, Wtask, out, ntiles, red); break;
        case 5: lg_contract_tile<5>(acc, d, h, tl, Wtask, out, ntiles, red); break;
        case 6: lg_contract_tile<6>(acc, d, h, tl, Wtask, out, ntiles, red); break;
        case 7: lg_contract_tile<7>(acc, d, h, tl, Wtask, out, ntiles, red); break;
        case 8: lg_contract_tile<8>(acc, d, h, tl, Wtask, out, ntiles, red); break;
        default: break;   // dim_m > 8 takes the two-kernel path (bqo_loglik_grad_batched)
    }
}

__global__ void loglik_grad_final_kernel(bqo_kernel_desc d, const double* __restrict__ hyp,
                                         int64_t stride, const double* __restrict__ partial,
                                         int nblk, const double* __restrict__ alpha, int n,
                                         const double* __restrict__ Wtask,
                                         double* __restrict__ grad) {
    __shared__ double red[32];
    __shared__ double tot[LG_MAX_ACC];
    const int s = blockIdx.x;
    const int dm = d.dim_m;
    const int nacc = dm + 4;
    const double* h = hyp + (int64_t)s * stride;
    for (int q = 0; q < nacc; ++q) {
        // fixed-order tree: each thread a strided serial sum, then the block tree
        double a = 0.0;
        for (int b = threadIdx.x; b < nblk; b += blockDim.x)
            a += partial[((int64_t)s * nacc + q) * nblk + b];
        a = bqo_block_sum(a, red);
        if (threadIdx.x == 0) tot[q] = a;
    }
    double sa = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sa += alpha[(int64_t)s * n + i];
    sa = bqo_block_sum(sa, red);
    __syncthreads();
    double* g = grad + (int64_t)s * (2 + d.n_params);
    if (threadIdx.x == 0) {
        g[0] = 0.5 * tot[LG_SLOT_TRACE(dm)];  // dK = I
        g[1] = sa;                            // 1^T Sigma^-1 (y - mean)
        for (int l = 0; l < dm; ++l) g[2 + l] = 0.5 * tot[l];
        if (d.kind == BQO_KIND_SCALED_MATERN52) g[2 + dm] = 0.5 * tot[LG_SLOT_KM(dm)];
        if (d.kind == BQO_KIND_PRODUCT_TASKS && d.same_corr) {
            const int T = d.n_tasks;
            const double* C = h + BQO_H_TASKC;
            if (T == 1) {
                g[2 + dm] = 0.5 * C[0] * tot[LG_SLOT_SAME(dm)];
            } else {
                double value = C[1];
                g[2 + dm] = 0.5 * (C[0] - value * (double)(T - 1)) * tot[LG_SLOT_SAME(dm)];
                g[2 + dm + 1] = 0.5 * (value * (double)(T - 1) * tot[LG_SLOT_SAME(dm)] +
                                       value * tot[LG_SLOT_DIFF(dm)]);
            }
        }
    }
    if (d.kind == BQO_KIND_PRODUCT_TASKS && !d.same_corr) {
        // grad_m = 0.5 sum_{u,v} dC_m[u][v] W[u][v],  dC_m = E + E^T, E[a][v] = L[a][b] L[v][b]
        const int T = d.n_tasks;
        const double* Lt = h + BQO_H_TASKC + T * T;
        const double* W = Wtask + (int64_t)s * T * T;
        const int pt = d.n_params - dm;
        for (int m = threadIdx.x; m < pt; m += blockDim.x) {
            int a = (int)floor((sqrt(8.0 * (double)m + 1.0) - 1.0) * 0.5);
            while ((a + 1) * (a + 2) / 2 <= m) ++a;
            while (a * (a + 1) / 2 > m) --a;
            int b = m - a * (a + 1) / 2;
            double lab = Lt[a * T + b];
            double accm = 0.0;
            for (int v = 0; v < T; ++v) accm += lab * Lt[v * T + b] * (W[a * T + v] + W[v * T + a]);
            g[2 + dm + m] = 0.5 * accm;
        }
    }
}

extern "C" int64_t bqo_loglik_grad_workspace_bytes(int32_t n, int32_t S) {
    int64_t nblk = (int64_t)((n + 31) / 32) * ((n + 31) / 32);
    int64_t d = 2 * (int64_t)S * n * n + (int64_t)S * LG_MAX_ACC * nblk + (int64_t)S * 128 * 128;
    return d * (int64_t)sizeof(double);
}

extern "C" int bqo_loglik_grad_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                                       const double* Xs, const int32_t* tid, int32_t n,
                                       const double* L, const double* alpha, const double* y,
                                       const double* Winv, double* grad, void* workspace,
                                       void* stream) {
    BQO_CHECK_ARG(desc != nullptr, 1);
    BQO_CHECK_ARG(hyp != nullptr, 2);
    BQO_CHECK_ARG(S > 0 && S <= 65535, 3);
    BQO_CHECK_ARG(Xs != nullptr, 4);
    BQO_CHECK_ARG(n > 0, 6);
    BQO_CHECK_ARG(L != nullptr, 7);
    BQO_CHECK_ARG(alpha != nullptr, 8);
    BQO_CHECK_ARG(grad != nullptr, 11);
    BQO_CHECK_ARG(workspace != nullptr, 12);
    BQO_CHECK_ARG(desc->n_tasks <= 128, 1);
    (void)y;
    cudaStream_t cs = (cudaStream_t)stream;
    const int nb = (n + 31) / 32;
    const int64_t nblk = (int64_t)nb * nb;
    double* Inv = (double*)workspace;
    double* Wbuf = Inv + (int64_t)S * n * n;
    double* partial = Wbuf + (int64_t)S * n * n;
    double* Wtask = partial + (int64_t)S * LG_MAX_ACC * nblk;
    // W = L^-1 either comes with the factorisation (bqo_chol_cov_batched's Winv) or is built here
    int rc = 0;
    const double* Wuse = Winv;
    if (!Wuse) {
        rc = bqo_internal_tri_inverse(L, n, S, Wbuf, Inv, cs);
        if (rc) return rc;
        Wuse = Wbuf;
    }
    if (desc->kind == BQO_KIND_PRODUCT_TASKS && !desc->same_corr)
        cudaMemsetAsync(Wtask, 0, sizeof(double) * S * desc->n_tasks * desc->n_tasks, cs);
    const int64_t st = bqo_hyper_stride_impl(*desc);
    bqo_ensure_global_exp_table(cs);
    if (desc->dim_m > 8) {
        // generic path: Sigma^-1 to memory, then the contraction over 32 x 32 tiles
        rc = bqo_internal_inv_syrk(Wuse, n, S, Inv, cs);
        if (rc) return rc;
        dim3 grid(1, nb, S);
        BQO_L, loglik_grad_partial_kernel<<<grid, 256, 0, cs>>>(*desc, hyp, st, Xs, tid, n, Inv, alpha,
                                                         partial, Wtask);
        BQO_L, loglik_grad_final_kernel<<<S, 256, 0, cs>>>(*desc, hyp, st, partial, nb, alpha, n, Wtask, grad);
        BQO_RETURN_LAST_ERROR();
    }
    const int nb64 = (n + BQO_TILE - 1) / BQO_TILE;
    const int ntiles = nb64 * (nb64 + 1) / 2;       // <= the (n/32)^2 slots of `partial`
    bqo_smem_opt_in<inv_syrk_contract_kernel>(LGF_SMEM_BYTES);
    dim3 grid(nb64, nb64, S);
    BQO_L, inv_syrk_contract_kernel<<<grid, 128, LGF_SMEM_BYTES, cs>>>(*desc, hyp, st, Xs, tid, n, Wuse,
                                                                alpha, partial, Wtask, ntiles);
    BQO_L, loglik_grad_final_kernel<<<S, 256, 0, cs>>>(*desc, hyp, st, partial, ntiles, alpha, n, Wtask, grad);
    BQO_RETURN_LAST_ERROR();
}
