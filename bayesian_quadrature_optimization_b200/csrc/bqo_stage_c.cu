// SBO / BQO stage B (per-unit setup) and stage C (a(x), b(x), inner maximisation over x,
// gradient of b w.r.t. the candidate, Monte-Carlo reduction).  SURVEY 8a rows a21-a34.
//
// Unit = one (candidate c, hyper-sample k) pair.  For a query point x the reference computes
//   a(x) = mean + B(x, X) alpha,   b(x) = (B(x, c) - B(x, X) s2) / den
// (compute_parameters_for_sample, sbo/numerical_tools/bayesian_quadrature.py:1148-1191) with
//   B(x, z_j) = E_W k((x, W), z_j)
// either a Monte-Carlo mean over M pre-drawn W samples (gamma_expect, sbo/lib/expectations.py:
// 76-118) or a finite weighted sum over tasks (uniform_finite, :10-74).  For the product kernel
// the task sum factorises, B(x, z_j) = k_M(x, z_j.x) q_j with q_j = sum_t w_t C[t, task_j], and the
// weights q_j are folded into alpha and into the solved rows R by the set-up entry points.
//
// Work decomposition: one WARP evaluates one point: lanes stride over the n training points,
// each lane loops over the M W samples of its point (the x part of the squared distance is
// shared by the M samples, and d/dx only needs sum_m k'(r_m)/r_m), and ten partial sums are
// combined with warp shuffles.  The inner maximiser keeps one warp per (Z sample, start) run;
// a CTA serves one unit so that X/ls, alpha and s2 are staged once in shared memory.  The W part
// of the squared distances does not depend on x and is read from a table built once per
// (hyper-sample, W set) (bqo_wdist_precompute).
#include "bqo_common.cuh"
#include <float.h>

#define SC_MAX_D BQO_MAX_DIM

// One dynamic shared-memory array for every kernel of this file.  Its first BQO_EXP_TABLE doubles
// hold the exp table of the fast path.
extern __shared__ __align__(16) double dyn_smem[];

// Table read T[m mod N] as LOP3 + IMAD + LDS: the address arithmetic is spelled in PTX because
// nvcc otherwise canonicalises (m & (N-1)) * 8 + base into shift, mask and add (one more
// instruction on the issue port per kernel evaluation).
struct SmemExpTab {
    unsigned base;  // shared-window address of dyn_smem[0]
    __device__ __forceinline__ SmemExpTab()
        : base((unsigned)__cvta_generic_to_shared(dyn_smem)) {}
    __device__ __forceinline__ double operator()(unsigned m) const {
        unsigned addr;
        double t;
        asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"(m & (unsigned)(BQO_EXP_TABLE - 1)), "r"(base));
        asm("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(addr));
        return t;
    }
};

// ---------------------------------------------------------------------------------------------
// alpha_q[s][j] = q_s[task_j] alpha_s[j]   (PRODUCT), plain copy otherwise.
__global__ void weight_alpha_kernel(const double* __restrict__ alpha, const double* __restrict__ qt,
                                    const int32_t* __restrict__ tid, int T, int n,
                                    double* __restrict__ out) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    int s = blockIdx.y;
    if (j >= n) return;
    double a = alpha[(int64_t)s * n + j];
    if (qt) a *= qt[(int64_t)s * T + tid[j]];
    out[(int64_t)s * n + j] = a;
}

extern "C" int bqo_weight_alpha(const bqo_kernel_desc* desc, const double* alpha, const double* qt,
                                const int32_t* tid, int32_t S, int32_t n, double* out,
                                void* stream) {
    BQO_CHECK_ARG(desc != nullptr, 1);
    BQO_CHECK_ARG(alpha != nullptr, 2);
    BQO_CHECK_ARG(S > 0 && n > 0, 5);
    BQO_CHECK_ARG(out != nullptr, 7);
    const bool prod = desc->kind == BQO_KIND_PRODUCT_TASKS;
    BQO_CHECK_ARG(!prod || (qt != nullptr && tid != nullptr), 3);
    dim3 grid((n + 255) / 256, S);
    BQO_L, weight_alpha_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(alpha, prod ? qt : nullptr, tid,
                                                               desc->n_tasks, n, out);
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// Stage B reductions.  One CTA per (candidate, hyper-sample).
//   G rows: gamma, grad_l gamma (unsolved);  R rows: s2 = Sigma^-1 gamma, Sigma^-1 grad_l gamma.
//   den = sqrt(clip(k(c,c) - gamma.s2 + noise, 0));  beta4_l = 2 grad_l gamma . s2
// then (PRODUCT) the solved rows are multiplied by q_j for stage C.
__global__ void candidate_setup_kernel(bqo_kernel_desc d, const double* __restrict__ hyp,
                                       int64_t stride, const double* __restrict__ G,
                                       double* __restrict__ R, int n, int C,
                                       const double* __restrict__ cand,
                                       const double* __restrict__ add_noise,
                                       const double* __restrict__ qt,
                                       const int32_t* __restrict__ tid, double* __restrict__ den,
                                       double* __restrict__ beta4) {
    __shared__ double red[32];
    const int c = blockIdx.x, s = blockIdx.y;
    const int rows = 1 + d.dim_m;
    const double* h = hyp + (int64_t)s * stride;
    const double* g0 = G + (((int64_t)s * C + c) * rows) * n;
    double* r0 = R + (((int64_t)s * C + c) * rows) * n;
    double acc[1 + SC_MAX_D];
    for (int l = 0; l < rows; ++l) acc[l] = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double s2 = r0[j];
        for (int l = 0; l < rows; ++l) acc[l] = fma(g0[(int64_t)l * n + j], s2, acc[l]);
    }
    for (int l = 0; l < rows; ++l) acc[l] = bqo_block_sum(acc[l], red);
    if (threadIdx.x == 0) {
        // k(c, c): Matern at distance 0 is 1; times sigma2 or the task variance
        double kcc = 1.0;
        if (d.kind == BQO_KIND_SCALED_MATERN52) kcc = h[BQO_H_SIGMA2];
        if (d.kind == BQO_KIND_PRODUCT_TASKS) {
            int tc = (int)cand[(int64_t)c * d.dim + d.dim - 1];
            kcc = bqo_task_cov(h, d.n_tasks, tc, tc);
        }
        double den2 = kcc - acc[0] + add_noise[s];
        den2 = den2 < 0.0 ? 0.0 : den2;  // np.clip(denominator, 0, None)
        den[(int64_t)s * C + c] = sqrt(den2);
        for (int l = 0; l < d.dim_m; ++l)
            beta4[((int64_t)s * C + c) * d.dim_m + l] = 2.0 * acc[1 + l];
    }
    if (qt != nullptr) {
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            double q = qt[(int64_t)s * d.n_tasks + tid[j]];
            for (int l = 0; l < rows; ++l) r0[(int64_t)l * n + j] *= q;
        }
    }
}

extern "C" int bqo_candidate_setup_batched(const bqo_kernel_desc* desc, const double* hyp,
                                           int32_t S, const double* G, double* R, int32_t n,
                                           int32_t C, const double* cand, const double* add_noise,
                                           const double* qt, const int32_t* tid, double* den,
                                           double* beta4, void* stream) {
    BQO_CHECK_ARG(desc != nullptr, 1);
    BQO_CHECK_ARG(hyp != nullptr, 2);
    BQO_CHECK_ARG(S > 0 && S <= 65535, 3);
    BQO_CHECK_ARG(G != nullptr, 4);
    BQO_CHECK_ARG(R != nullptr, 5);
    BQO_CHECK_ARG(n > 0, 6);
    BQO_CHECK_ARG(C > 0, 7);
    BQO_CHECK_ARG(cand != nullptr, 8);
    BQO_CHECK_ARG(add_noise != nullptr, 9);
    BQO_CHECK_ARG(den != nullptr, 12);
    BQO_CHECK_ARG(beta4 != nullptr, 13);
    const bool prod = desc->kind == BQO_KIND_PRODUCT_TASKS;
    BQO_CHECK_ARG(!prod || (qt != nullptr && tid != nullptr), 10);
    dim3 grid(C, S);
    BQO_L, candidate_setup_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        *desc, hyp, bqo_hyper_stride_impl(*desc), G, R, n, C, cand, add_noise,
        prod ? qt : nullptr, tid, den, beta4);
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// Per-unit constants resolved once per warp / CTA.
template <int DX, int DW>
struct UnitCtx {
    const double* xs;   // [(DX+DW)][ldx] scaled training coordinates (global or shared)
    const double* wa;   // [n] weights of a (alpha, q-weighted for PRODUCT)
    const double* wb;   // [n] weights of b (s2, q-weighted for PRODUCT)
    const double* tab;  // shared-memory exp table of the fast evaluation (bqo_common.cuh)
    int ldx;            // stride between coordinate rows of xs
    int n;
    int M;
    double ls[DX + DW];      // length scales (x then w)
    double inv_ls[DX + DW];
    double cs[DX + DW];      // scaled candidate coordinates
    double wc;               // 1, or q[task_c]
    double scale;            // sigma2 / M, or 1
    double mean;
    double inv_den;          // 1/den, or 0 when den == 0 (b := 0, bayesian_quadrature.py:1186-1189)
};

template <int DX, int DW>
__device__ __forceinline__ void unit_ctx_init(UnitCtx<DX, DW>& u, const bqo_stage_c& pb, int unit) {
    const int k = pb.unit_k[unit], c = pb.unit_c[unit];
    const int64_t stride = bqo_hyper_stride_impl(pb.desc);
    const double* h = pb.hyp + (int64_t)k * stride;
    const int dm = pb.desc.dim_m;
    u.xs = pb.Xs + (int64_t)k * dm * pb.n;
    u.ldx = pb.n;
    u.wa = pb.alpha + (int64_t)k * pb.n;
    u.wb = pb.R + (((int64_t)k * pb.C + c) * (1 + dm)) * pb.n;
    u.n = pb.n;
    u.M = (DW > 0) ? pb.M : 1;
#pragma unroll
    for (int l = 0; l < DX + DW; ++l) {
        u.ls[l] = h[BQO_H_LS + l];
        u.inv_ls[l] = h[BQO_H_INVLS + l];
        u.cs[l] = pb.cand[(int64_t)c * pb.desc.dim + l] / u.ls[l];
    }
    u.wc = 1.0;
    u.scale = 1.0;
    if (pb.desc.kind == BQO_KIND_SCALED_MATERN52) u.scale = h[BQO_H_SIGMA2];
    if (DW > 0) u.scale = u.scale / (double)u.M;
    if (pb.desc.kind == BQO_KIND_PRODUCT_TASKS) {
        int tc = (int)pb.cand[(int64_t)c * pb.desc.dim + pb.desc.dim - 1];
        u.wc = pb.qt[(int64_t)k * pb.desc.n_tasks + tc];
    }
    u.mean = h[BQO_H_MEAN];
    double den = pb.den[(int64_t)k * pb.C + c];
    u.inv_den = (den != 0.0) ? 1.0 / den : 0.0;
}

// Result of one evaluation (valid in every lane).
template <int DX>
struct AbOut {
    double a, b;
    double ga[DX], gb[DX];
};

// Operand pointers of one evaluation, kept in registers.  In the inner maximiser they are built
// straight from the dynamic shared-memory carve-up so that ptxas knows the address space (LDS);
// loaded back from a struct in shared memory they degrade to generic LD instructions.
struct EvalPtrs {
    const double* xs;   // [(DX+DW)][ldx]
    const double* wa;   // [n]
    const double* wb;   // [n]
    const double* tab;  // exp table (shared memory)
    int ldx, n, M;
};

template <int DX, int DW>
__device__ __forceinline__ EvalPtrs eval_ptrs_from(const UnitCtx<DX, DW>& u) {
    EvalPtrs p;
    p.xs = u.xs;
    p.wa = u.wa;
    p.wb = u.wb;
    p.tab = u.tab;
    p.ldx = u.ldx;
    p.n = u.n;
    p.M = u.M;
    return p;
}

// W samples pushed through the Matern stages together (independent DFMA chains per lane).
#ifndef EVAL_NV
#define EVAL_NV 2
#endif
#ifndef BQO_LOSRC
#define BQO_LOSRC true
#endif

// Warp-cooperative evaluation of a, b and their x-gradients at the raw point x.
// Works in coordinates multiplied by K = sqrt5 N / ln2 (u2 = K^2 r^2, see bqo_matern52_q_v):
//   `ws`: this evaluation's W samples times K / ls_w, [M][DW];
//   PRE = true: u.xs already holds K * X / ls (staged copy); false: X / ls, scaled on load.
// Per training point three sums over the W samples are kept, E0 = sum e, E1 = sum u e and
// H = sum u2 e (e = exp(-sqrt5 r)); then with c = ln2/N, G = E0 + c E1 = sum (1 + sqrt5 r) e,
// sum_m k = G + (c^2/3) H and sum_m k'(r)/r = -(5/3) G.
// CMB = true evaluates F = a + Z b and grad F only, with ONE weight per training point,
// w_j = wa_j - (Z/den) wb_j (`zd` = Z/den): half the accumulators, half the per-point FMAs and
// ten registers less in the line-search evaluations of the inner maximiser.  F is returned in
// o.a and grad F in o.ga (o.b = o.gb = 0).
// WD = true takes the W part of every squared distance, sum_l (K w_ml/ls_l - K X_jl/ls_l)^2, from
// the table `wd` built once per (hyper-sample, W set) by bqo_wdist_precompute: it does not depend
// on x, so the 60 evaluations per (unit, Z) need not recompute it.  `wd` points at this lane's
// column of the unit's table, laid out [j / 32][m][32] (coalesced, W-sample offsets immediate).
template <int DX, int DW, bool PRE, int MT = 0, bool CMB = false, bool WD = false>
__device__ __forceinline__ void eval_ab_warp(const EvalPtrs& p, const UnitCtx<DX, DW>& u,
                                             const double* x, const double* __restrict__ ws,
                                             AbOut<DX>& o, double zd = 0.0,
                                             const double* __restrict__ wd = nullptr) {
    const int lane = threadIdx.x & 31;
    double xsc[DX];
#pragma unroll
    for (int l = 0; l < DX; ++l) xsc[l] = (x[l] / u.ls[l]) * BQO_PRESCALE;
    double Sa = 0.0, Sb = 0.0, GA[DX], GB[DX];
#pragma unroll
    for (int l = 0; l < DX; ++l) GA[l] = GB[l] = 0.0;
    const int M = (MT > 0) ? MT : p.M;
    constexpr bool REC = PRE && WD && (DW > 0);
    constexpr int RECLD = (DX + 2) | 1;
    // finite W: one kernel evaluation per training point, so the loop bookkeeping is a quarter of
    // the non-FP64 work: two points per trip (the W-sample loop already fills a trip otherwise)
#pragma unroll(DW == 0 ? 2 : 1)
    for (int j = lane; j < p.n; j += 32) {
        double dxl[DX];
        // keeps q2 a normal positive number for the fast sqrt (k(0) = 1); with the W-distance
        // table the floor is part of the table entries
        double Dx = (DW > 0 && WD) ? 0.0 : 1e-300;
#pragma unroll
        for (int l = 0; l < DX; ++l) {
            // staged + W-distance table: one record [x_1..x_DX, wa, wb, pad] per training point
            // (odd stride in doubles: conflict-free 64-bit loads at immediate offsets)
            double xj = REC ? p.xs[(size_t)j * RECLD + l] : p.xs[l * p.ldx + j];
            if (!PRE) xj *= BQO_PRESCALE;
            dxl[l] = xsc[l] - xj;
            Dx = fma(dxl[l], dxl[l], Dx);
        }
        double E0 = 0.0, E1 = 0.0, H = 0.0;  // sum e, sum u e, sum u2 e over the W samples
        if (DW > 0 && WD) {
            const double* wj = wd + (size_t)(j >> 5) * (M * 32);
            // one clamp of the x part per training point (the table is bounded when it is built)
            // replaces the clamp inside every kernel evaluation
            Dx = __hiloint2double((int)min((unsigned)__double2hiint(Dx), (unsigned)BQO_U2_PART_HI),
                                  __double2loint(Dx));
            auto step_nv = [&](int m) {
                double q2v[EVAL_NV], ev[EVAL_NV], t1v[EVAL_NV], wv[EVAL_NV];
#pragma unroll
                for (int q = 0; q < EVAL_NV; ++q) {
                    wv[q] = __ldg(wj + (m + q) * 32);
                    q2v[q] = Dx + wv[q];
                }
                bqo_matern52_q_v<EVAL_NV, SmemExpTab, false, BQO_LOSRC>(q2v, SmemExpTab(), ev, t1v, wv);
#pragma unroll
                for (int q = 0; q < EVAL_NV; ++q) {
                    E0 += ev[q];
                    E1 = fma(t1v[q], ev[q], E1);
                    H = fma(q2v[q], ev[q], H);
                }
            };
            auto step_1 = [&](int m) {
                double q2v[1], ev[1], t1v[1];
                q2v[0] = Dx + __ldg(wj + m * 32);
                bqo_matern52_q_v<1, SmemExpTab, false>(q2v, SmemExpTab(), ev, t1v);
                E0 += ev[0];
                E1 = fma(t1v[0], ev[0], E1);
                H = fma(q2v[0], ev[0], H);
            };
            if (MT > 0) {
#pragma unroll
                for (int m = 0; m + EVAL_NV <= MT; m += EVAL_NV) step_nv(m);
#pragma unroll
                for (int m = MT - (MT % EVAL_NV); m < MT; ++m) step_1(m);
            } else {
                int m = 0;
                for (; m + EVAL_NV <= M; m += EVAL_NV) step_nv(m);
                for (; m < M; ++m) step_1(m);
            }
        } else if (DW > 0) {
            double xw[DW > 0 ? DW : 1];
#pragma unroll
            for (int l = 0; l < DW; ++l) {
                xw[l] = p.xs[(DX + l) * p.ldx + j];
                if (!PRE) xw[l] *= BQO_PRESCALE;
            }
            // The W samples do not depend on j: hide that from the optimiser, otherwise the
            // unrolled build hoists all 2M of them out of the j loop and spills them.
            const double* wsj = ws;
            asm volatile("" : "+l"(wsj));
            // NV W samples go through the Matern stages together; G and H are accumulated directly
            // (the loop is issue-bound, not latency-bound: no split accumulators).  With MT > 0
            // the trip count is a compile-time constant: full unroll, W-sample offsets become
            // immediates and the loop bookkeeping disappears.
            auto step_nv = [&](int m) {
                double q2v[EVAL_NV], ev[EVAL_NV], t1v[EVAL_NV];
#pragma unroll
                for (int q = 0; q < EVAL_NV; ++q) {
                    double q2 = Dx;
#pragma unroll
                    for (int l = 0; l < DW; ++l) {
                        double dw = wsj[(m + q) * DW + l] - xw[l];
                        q2 = fma(dw, dw, q2);
                    }
                    q2v[q] = q2;
                }
                bqo_matern52_q_v<EVAL_NV>(q2v, SmemExpTab(), ev, t1v);
#pragma unroll
                for (int q = 0; q < EVAL_NV; ++q) {
                    E0 += ev[q];
                    E1 = fma(t1v[q], ev[q], E1);
                    H = fma(q2v[q], ev[q], H);
                }
            };
            auto step_1 = [&](int m) {
                double q2v[1], ev[1], t1v[1];
                q2v[0] = Dx;
#pragma unroll
                for (int l = 0; l < DW; ++l) {
                    double dw = wsj[m * DW + l] - xw[l];
                    q2v[0] = fma(dw, dw, q2v[0]);
                }
                bqo_matern52_q_v<1>(q2v, SmemExpTab(), ev, t1v);
                E0 += ev[0];
                E1 = fma(t1v[0], ev[0], E1);
                H = fma(q2v[0], ev[0], H);
            };
            if (MT > 0) {
#pragma unroll
                for (int m = 0; m + EVAL_NV <= MT; m += EVAL_NV) step_nv(m);
#pragma unroll
                for (int m = MT - (MT % EVAL_NV); m < MT; ++m) step_1(m);
            } else {
                int m = 0;
                for (; m + EVAL_NV <= M; m += EVAL_NV) step_nv(m);
                for (; m < M; ++m) step_1(m);
            }
        } else {
            double q2v[1], ev[1], t1v[1];
            q2v[0] = Dx;
            bqo_matern52_q_v<1>(q2v, SmemExpTab(), ev, t1v);
            E0 = ev[0];
            E1 = t1v[0] * ev[0];
            H = q2v[0] * ev[0];
        }
        // c and c^2/3 as constant-bank operands: as literals they cost two UMOVs each per point
        const double G = fma(E1, BQO_FC[4], E0);  // sum_m (1 + sqrt5 r_m) e_m
        const double ksum = fma(H, BQO_FC[5], G);
        const double wa = REC ? p.xs[(size_t)j * RECLD + DX] : p.wa[j];
        const double wb = REC ? p.xs[(size_t)j * RECLD + DX + 1] : p.wb[j];
        if (CMB) {
            const double w = fma(-zd, wb, wa);
            Sa = fma(ksum, w, Sa);
            const double uw = G * w;
#pragma unroll
            for (int l = 0; l < DX; ++l) GA[l] = fma(uw, dxl[l], GA[l]);
        } else {
            Sa = fma(ksum, wa, Sa);
            Sb = fma(ksum, wb, Sb);
            const double ua = G * wa, ub = G * wb;
#pragma unroll
            for (int l = 0; l < DX; ++l) {
                GA[l] = fma(ua, dxl[l], GA[l]);
                GB[l] = fma(ub, dxl[l], GB[l]);
            }
        }
    }
    // candidate term B(x, c): lane m evaluates W sample m
    double Kc = 0.0, GC[DX];
#pragma unroll
    for (int l = 0; l < DX; ++l) GC[l] = 0.0;
    for (int m = lane; m < M; m += 32) {
        double dxl[DX];
        double q2v[1], ev[1], t1v[1];
        q2v[0] = 1e-300;
#pragma unroll
        for (int l = 0; l < DX; ++l) {
            dxl[l] = xsc[l] - u.cs[l] * BQO_PRESCALE;
            q2v[0] = fma(dxl[l], dxl[l], q2v[0]);
        }
#pragma unroll
        for (int l = 0; l < DW; ++l) {
            double dw = ws[m * DW + l] - u.cs[DX + l] * BQO_PRESCALE;
            q2v[0] = fma(dw, dw, q2v[0]);
        }
        bqo_matern52_q_v<1>(q2v, SmemExpTab(), ev, t1v);
        const double gm = fma(t1v[0], BQO_LN2_N, 1.0) * ev[0];
        Kc += fma(q2v[0] * ev[0], BQO_H_TO_K, gm);
#pragma unroll
        for (int l = 0; l < DX; ++l) GC[l] = fma(gm, dxl[l], GC[l]);
    }
    // d/dx: k'(r)/r * (x - X)/ls^2 = -(5/3) G * (dx'/K) / ls  with dx' the difference in units of K
    const double gfac = BQO_GRAD_FAC;
    if (CMB) {
        Sa = bqo_warp_sum(Sa);
        Kc = bqo_warp_sum(Kc);
        const double zc = zd * u.wc;
        o.a = u.mean + u.scale * fma(zc, Kc, Sa);
        o.b = 0.0;
#pragma unroll
        for (int l = 0; l < DX; ++l) {
            double gs = bqo_warp_sum(GA[l]);
            double gc = bqo_warp_sum(GC[l]);
            o.ga[l] = u.scale * gfac * fma(zc, gc, gs) * u.inv_ls[l];
            o.gb[l] = 0.0;
        }
        return;
    }
    Sa = bqo_warp_sum(Sa);
    Sb = bqo_warp_sum(Sb);
    Kc = bqo_warp_sum(Kc);
    o.a = u.mean + u.scale * Sa;
    o.b = u.scale * (u.wc * Kc - Sb) * u.inv_den;
#pragma unroll
    for (int l = 0; l < DX; ++l) {
        double ga = bqo_warp_sum(GA[l]) * gfac;
        double gb = bqo_warp_sum(GB[l]) * gfac;
        double gc = bqo_warp_sum(GC[l]) * gfac;
        o.ga[l] = u.scale * ga * u.inv_ls[l];
        o.gb[l] = u.scale * (u.wc * gc - gb) * u.inv_ls[l] * u.inv_den;
    }
}

// Scale one W sample set by the w length scales into `dst` ([M][DW]); executed by a warp.
template <int DX, int DW>
__device__ __forceinline__ void scale_wset(const UnitCtx<DX, DW>& u, const double* __restrict__ w,
                                           double* dst, double mult) {
    if (DW == 0) return;
    const int lane = threadIdx.x & 31;
    for (int e = lane; e < u.M * DW; e += 32)
        dst[e] = (w[e] / u.ls[DX + (e % (DW > 0 ? DW : 1))]) * mult;
}

// ---------------------------------------------------------------------------------------------
// a, b, grad a, grad b at P points; warp per point, 8 warps per CTA, operands read from global.
#define AB_WARPS 8
template <int DX, int DW>
__global__ void __launch_bounds__(AB_WARPS * 32) sample_ab_kernel(bqo_stage_c pb,
                                                                  const double* __restrict__ pts,
                                                                  const int32_t* __restrict__ pt_unit,
                                                                  const int32_t* __restrict__ pt_wset,
                                                                  int P, double* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = blockIdx.x * AB_WARPS + warp;
    bqo_exp_table_init(dyn_smem);
    __syncthreads();
    if (q >= P) return;
    double* ws = dyn_smem + BQO_EXP_TABLE + warp * (pb.M * (DW > 0 ? DW : 1));
    UnitCtx<DX, DW> u;
    unit_ctx_init<DX, DW>(u, pb, pt_unit[q]);
    u.tab = dyn_smem;
    if (DW > 0) {
        int wsi = pt_wset ? pt_wset[q] : 0;
        scale_wset<DX, DW>(u, pb.wsets + (int64_t)wsi * pb.M * DW, ws, BQO_PRESCALE);
        __syncwarp();
    }
    double x[DX];
#pragma unroll
    for (int l = 0; l < DX; ++l) x[l] = pts[(int64_t)q * DX + l];
    AbOut<DX> o;
    const EvalPtrs ep = eval_ptrs_from<DX, DW>(u);
    eval_ab_warp<DX, DW, false>(ep, u, x, ws, o);
    if (lane == 0) {
        double* op = out + (int64_t)q * (2 + 2 * DX);
        op[0] = o.a;
        op[1] = o.b;
#pragma unroll
        for (int l = 0; l < DX; ++l) {
            op[2 + l] = o.ga[l];
            op[2 + DX + l] = o.gb[l];
        }
    }
}

static int check_stage_c(const bqo_stage_c* pb) {
    if (!pb) return 0;
    if (pb->n <= 0 || pb->S <= 0 || pb->dx <= 0 || pb->dw < 0) return 0;
    if (pb->dx + pb->dw != pb->desc.dim_m) return 0;
    if (pb->desc.kind == BQO_KIND_PRODUCT_TASKS && (pb->dw != 0 || !pb->qt || !pb->tid)) return 0;
    if (pb->dw > 0 && (pb->M <= 0 || pb->n_wsets <= 0 || !pb->wsets)) return 0;
    if (!pb->hyp || !pb->Xs || !pb->alpha || !pb->R || !pb->den || !pb->cand || !pb->unit_k ||
        !pb->unit_c)
        return 0;
    return 1;
}

// (DX, DW) instantiations: x dims 1..6, w dims 0..2.
#define SC_DISPATCH(pb, MACRO)                                               \
    do {                                                                     \
        const int dx__ = (pb)->dx, dw__ = (pb)->dw;                          \
        if (dw__ == 0) {                                                     \
            switch (dx__) {                                                  \
                case 1: MACRO(1, 0); break;                                  \
                case 2: MACRO(2, 0); break;                                  \
                case 3: MACRO(3, 0); break;                                  \
                case 4: MACRO(4, 0); break;                                  \
                case 5: MACRO(5, 0); break;                                  \
                case 6: MACRO(6, 0); break;                                  \
                default: return -1;                                          \
            }                                                                \
        } else if (dw__ == 1) {                                              \
            switch (dx__) {                                                  \
                case 1: MACRO(1, 1); break;                                  \
                case 2: MACRO(2, 1); break;                                  \
                case 3: MACRO(3, 1); break;                                  \
                case 4: MACRO(4, 1); break;                                  \
                case 5: MACRO(5, 1); break;                                  \
                case 6: MACRO(6, 1); break;                                  \
                default: return -1;                                          \
            }                                                                \
        } else if (dw__ == 2) {                                              \
            switch (dx__) {                                                  \
                case 1: MACRO(1, 2); break;                                  \
                case 2: MACRO(2, 2); break;                                  \
                case 3: MACRO(3, 2); break;                                  \
                case 4: MACRO(4, 2); break;                                  \
                case 5: MACRO(5, 2); break;                                  \
                case 6: MACRO(6, 2); break;                                  \
                default: return -1;                                          \
            }                                                                \
        } else {                                                             \
            return -1;                                                       \
        }                                                                    \
    } while (0)

extern "C" int bqo_sample_ab_batched(const bqo_stage_c* pb, const double* pts,
                                     const int32_t* pt_unit, const int32_t* pt_wset, int32_t P,
                                     double* out, void* stream) {
    BQO_CHECK_ARG(check_stage_c(pb), 1);
    BQO_CHECK_ARG(pts != nullptr, 2);
    BQO_CHECK_ARG(pt_unit != nullptr, 3);
    BQO_CHECK_ARG(P > 0, 5);
    BQO_CHECK_ARG(out != nullptr, 6);
    cudaStream_t cs = (cudaStream_t)stream;
    int grid = (P + AB_WARPS - 1) / AB_WARPS;
    size_t smem = sizeof(double) * (BQO_EXP_TABLE + AB_WARPS * (size_t)(pb->M > 0 ? pb->M : 1) * (pb->dw > 0 ? pb->dw : 1));
#define LAUNCH_AB(DXV, DWV) \
    BQO_L, sample_ab_kernel<DXV, DWV><<<grid, AB_WARPS * 32, smem, cs>>>(*pb, pts, pt_unit, pt_wset, P, out)
    SC_DISPATCH(pb, LAUNCH_AB);
#undef LAUNCH_AB
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// W-distance table: out[s][ws][j / 32][m][j % 32] = sum_l (K w_ml / ls_l - K X_jl / ls_l)^2 over the
// w coordinates, in the units of the fast evaluation (K = BQO_PRESCALE), with exactly the
// operand expressions of the on-the-fly path.  It depends on the hyper-sample and the W set
// only, so it is built once per BQ engine and read by every evaluation of the inner maximiser.
__global__ void wdist_kernel(const double* __restrict__ hyp, int64_t stride,
                             const double* __restrict__ Xs, int n, int dim_m, int dx, int dw,
                             const double* __restrict__ wsets, int n_wsets, int M,
                             double* __restrict__ out) {
    const int nblk = (n + 31) / 32;
    const int64_t per = (int64_t)nblk * M * 32;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y, wsi = blockIdx.z;
    if (idx >= per) return;
    const int lane = (int)(idx & 31);
    const int m = (int)((idx >> 5) % M);
    const int jb = (int)((idx >> 5) / M);
    const int j = jb * 32 + lane;
    double acc = 0.0;
    if (j < n) {
        const double* h = hyp + (int64_t)s * stride;
        for (int l = 0; l < dw; ++l) {
            const double wv = (wsets[((int64_t)wsi * M + m) * dw + l] / h[BQO_H_LS + dx + l]) *
                              BQO_PRESCALE;
            const double xw = Xs[((int64_t)s * dim_m + dx + l) * n + j] * BQO_PRESCALE;
            const double d = wv - xw;
            acc = fma(d, d, acc);
        }
        // bounded (the evaluation then needs no clamp of its own) and strictly positive (the
        // fast square root needs a normal positive argument; + 1e-300 only changes acc == 0)
        acc = fmin(acc, __hiloint2double(BQO_U2_PART_HI, 0)) + 1e-300;
    }
    out[((int64_t)s * n_wsets + wsi) * per + idx] = acc;
}

extern "C" int bqo_wdist_precompute(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                                    const double* Xs, int32_t n, int32_t dx, int32_t dw,
                                    const double* wsets, int32_t n_wsets, int32_t M, double* out,
                                    void* stream) {
    BQO_CHECK_ARG(desc != nullptr, 1);
    BQO_CHECK_ARG(hyp != nullptr, 2);
    BQO_CHECK_ARG(S > 0 && S <= 65535, 3);
    BQO_CHECK_ARG(Xs != nullptr, 4);
    BQO_CHECK_ARG(n > 0, 5);
    BQO_CHECK_ARG(dx > 0 && dw > 0 && dx + dw == desc->dim_m, 6);
    BQO_CHECK_ARG(wsets != nullptr, 8);
    BQO_CHECK_ARG(n_wsets > 0 && n_wsets <= 65535, 9);
    BQO_CHECK_ARG(M > 0, 10);
    BQO_CHECK_ARG(out != nullptr, 11);
    const int64_t per = (int64_t)((n + 31) / 32) * M * 32;
    dim3 grid((unsigned)((per + 255) / 256), S, n_wsets);
    BQO_L, wdist_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(hyp, bqo_hyper_stride_impl(*desc), Xs, n,
                                                         desc->dim_m, dx, dw, wsets, n_wsets, M, out);
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// Inner maximisation.  CTA = (unit, chunk of Z samples); warp = one run (Z sample, start).
// The optimiser is a projected BFGS with Armijo backtracking on phi = -(a + Z b); its iteration
// cap, relative-decrease test (factr) and projected-gradient test (pgtol) are those of the
// L-BFGS-B call the reference makes (scipy fmin_l_bfgs_b via Optimization.optimize,
// sbo/lib/optimization.py:87-155, with maxiter=10, factr=1e12 from sbo/services/bgo.py:43-44).
// oracle/inner_opt.py restates the same iteration in numpy.
// 28 warps x 72 registers: measured best on B200 once the W-distance table freed the registers
// that held the W samples (sweeps of {12..32} warps x {1,2,5} interleaved W samples,
// profiles/r01_inner_maximize_v1.md and r01_tuning_sweeps.md; 24 x 80 before the table)
#ifndef IM_WARPS
#define IM_WARPS 28
#endif
#define IM_MAX_VERT 64  // 2^DX for DX <= 6

template <int DX>
struct RunState {  // per-warp optimiser state in shared memory (warp-uniform values)
    double x[DX], g[DX], xn[DX], gn[DX], d[DX];
    double H[DX * DX];
    double phi;
};

template <int DX, int DW, bool STAGE, int MT, bool WD = false>
__global__ void __launch_bounds__(IM_WARPS * 32, 1)
    inner_maximize_kernel(bqo_stage_c pb, bqo_inner_opt opt, const double* __restrict__ z,
                          const double* __restrict__ starts, double* __restrict__ out_max,
                          double* __restrict__ out_arg, unsigned long long* __restrict__ n_evals) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int zc_per_unit = (opt.n_z + opt.z_per_cta - 1) / opt.z_per_cta;
    // Without staging (n too large for shared memory) the operands of a unit stream from L2, and
    // those shared by a hyper-sample (scaled points, alpha, the W-distance table: 3 MB at
    // n = 8192) should be shared by the CTAs that run together: blocks walk the units
    // hyper-sample-major (unit = c * S_u + k: block b -> c = b % C, k = b / C), not candidate-major
    // (at S = 64 the candidate-major order touches 64 x 3 MB at once and falls out of the L2).
    int unit = blockIdx.x / zc_per_unit;
    if (!STAGE && pb.C > 0 && pb.n_units % pb.C == 0) {
        const int s_units = pb.n_units / pb.C;
        unit = (unit % pb.C) * s_units + unit / pb.C;
    }
    const int zc = blockIdx.x % zc_per_unit;
    const int z0 = zc * opt.z_per_cta;
    const int nz = (opt.n_z - z0 < opt.z_per_cta) ? opt.n_z - z0 : opt.z_per_cta;
    const int n = pb.n, M = (DW > 0) ? pb.M : 1;
    constexpr int D = DX + DW;
    constexpr int NV = 1 << DX;
    const int n_starts = opt.n_starts;
    const int n_pre = n_starts + (opt.use_vertices ? NV : 0);

    // ---- shared memory carve-up -------------------------------------------------------------
    double* s_tab = dyn_smem;                // [BQO_EXP_TABLE] exp table
    double* sm = dyn_smem + BQO_EXP_TABLE;
    constexpr int XROWS = WD ? DX : D;       // with the W-distance table the w rows stay in HBM
    constexpr int RECLD = (DX + 2) | 1;      // WD: records [x_1..x_DX, wa, wb, pad] per point
    double* s_xs = sm;                       // STAGE: [XROWS][n], or [n][RECLD] with WD
    double* s_wa = s_xs + (STAGE ? (WD ? (size_t)RECLD * n : (size_t)XROWS * n) : 0);
    double* s_wb = s_wa + ((STAGE && !WD) ? n : 0);
    double* s_ws = s_wb + ((STAGE && !WD) ? n : 0);   // [M][DW] the unit's W set, scaled
    double* s_pre = s_ws + (size_t)M * (DW > 0 ? DW : 0);  // [n_pre][2+2DX]
    double* s_prex = s_pre + (size_t)n_pre * (2 + 2 * DX);              // [n_pre][DX]
    double* s_run = s_prex + (size_t)n_pre * DX;  // [nz][n_starts][1+DX] run results
    RunState<DX>* s_state =
        reinterpret_cast<RunState<DX>*>(s_run + (size_t)opt.z_per_cta * n_starts * (1 + DX));
    int* s_counter = reinterpret_cast<int*>(s_state + IM_WARPS);

    // the unit's constants live in shared memory (read with broadcast loads where needed), not in
    // 40 registers of every thread: registers buy resident warps, and resident warps are what
    // hides the 8-cycle DFMA latency (tests/microbench/fp64_issue.cu)
    __shared__ UnitCtx<DX, DW> s_u;
    if (threadIdx.x == 0) {
        unit_ctx_init<DX, DW>(s_u, pb, unit);
        s_u.tab = s_tab;
    }
    bqo_exp_table_init(s_tab);
    __syncthreads();
    if (STAGE) {
        const double* gxs = s_u.xs;
        const double* gwa = s_u.wa;
        const double* gwb = s_u.wb;
        if (WD) {
            for (int e = threadIdx.x; e < DX * n; e += blockDim.x) {
                const int l = e / n, j = e - l * n;  // coalesced global reads
                s_xs[(size_t)j * RECLD + l] = gxs[e] * BQO_PRESCALE;
            }
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                s_xs[(size_t)j * RECLD + DX] = gwa[j];
                s_xs[(size_t)j * RECLD + DX + 1] = gwb[j];
            }
        } else {
            for (int e = threadIdx.x; e < XROWS * n; e += blockDim.x)
                s_xs[e] = gxs[e] * BQO_PRESCALE;
            for (int e = threadIdx.x; e < n; e += blockDim.x) {
                s_wa[e] = gwa[e];
                s_wb[e] = gwb[e];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s_u.xs = s_xs;
            s_u.wa = s_wa;
            s_u.wb = s_wb;
        }
    }
    const UnitCtx<DX, DW>& u = s_u;
    EvalPtrs ep;
    ep.xs = STAGE ? s_xs : pb.Xs + (int64_t)pb.unit_k[unit] * D * n;
    ep.wa = STAGE ? s_wa : pb.alpha + (int64_t)pb.unit_k[unit] * n;
    ep.wb = STAGE ? s_wb
                  : pb.R + (((int64_t)pb.unit_k[unit] * pb.C + pb.unit_c[unit]) * (1 + D)) * n;
    ep.tab = s_tab;
    ep.ldx = n;
    ep.n = n;
    ep.M = M;
    if (DW > 0) {
        // every evaluation of this unit uses the W set of its candidate (sample-average
        // approximation: the objective of a run is a fixed smooth function)
        const double* wsrc = pb.wsets + (size_t)(pb.unit_c[unit] % pb.n_wsets) * M * DW;
        for (int e = threadIdx.x; e < M * DW; e += blockDim.x)
            s_ws[e] = (wsrc[e] / u.ls[DX + (e % (DW > 0 ? DW : 1))]) * BQO_PRESCALE;
    }
    // this lane's column of the unit's W-distance table (hyper-sample k, W set c % n_wsets)
    const double* wd = nullptr;
    if (WD) {
        const size_t nblk = (size_t)(n + 31) / 32;
        wd = pb.wdist + (((size_t)pb.unit_k[unit] * pb.n_wsets + (pb.unit_c[unit] % pb.n_wsets)) *
                         nblk) * ((size_t)M * 32) + lane;
    }
    if (threadIdx.x == 0) *s_counter = 0;
    const double* st = starts + (opt.starts_per_hyper ? (int64_t)pb.unit_k[unit] * n_starts * DX : 0);
    __syncthreads();

    unsigned long long evals = 0;
    // ---- phase 1: a, b, grads at the start points and the box vertices (shared by all Z) ----
    for (int p = warp; p < n_pre; p += IM_WARPS) {
        double x[DX];
        if (p < n_starts) {
#pragma unroll
            for (int l = 0; l < DX; ++l) {
                double v = st[p * DX + l];
                v = fmin(fmax(v, pb.lower[l]), pb.upper[l]);
                x[l] = v;
            }
        } else {
            // itertools.product(*bounds_x) order: last coordinate varies fastest (sbo.py:266-269)
            int v = p - n_starts;
#pragma unroll
            for (int l = 0; l < DX; ++l) x[l] = ((v >> (DX - 1 - l)) & 1) ? pb.upper[l] : pb.lower[l];
        }
        AbOut<DX> o;
        eval_ab_warp<DX, DW, STAGE, MT, false, WD>(ep, u, x, s_ws, o, 0.0, wd);
        ++evals;
        if (lane == 0) {
            double* op = s_pre + (size_t)p * (2 + 2 * DX);
            op[0] = o.a;
            op[1] = o.b;
#pragma unroll
            for (int l = 0; l < DX; ++l) {
                op[2 + l] = o.ga[l];
                op[2 + DX + l] = o.gb[l];
                s_prex[p * DX + l] = x[l];
            }
        }
    }
    __syncthreads();

    // ---- phase 2: runs --------------------------------------------------------------------------
    RunState<DX>& S = s_state[warp];
    const int n_runs = nz * n_starts;
    const double c1 = 1e-4;
    for (;;) {
        int run = 0;
        if (lane == 0) run = atomicAdd(s_counter, 1);
        run = __shfl_sync(0xffffffffu, run, 0);
        if (run >= n_runs) break;
        const int zi = run / n_starts, sidx = run % n_starts;
        const double Z = z[z0 + zi];
        // phi = -(a + Z b)
        if (lane == 0) {
            const double* op = s_pre + (size_t)sidx * (2 + 2 * DX);
            S.phi = -(op[0] + Z * op[1]);
#pragma unroll
            for (int l = 0; l < DX; ++l) {
                S.x[l] = s_prex[sidx * DX + l];
                S.g[l] = -(op[2 + l] + Z * op[2 + DX + l]);
            }
#pragma unroll
            for (int e = 0; e < DX * DX; ++e) S.H[e] = ((e / DX) == (e % DX)) ? 1.0 : 0.0;
        }
        __syncwarp();
        bool first_update = true;
        for (int it = 0; it < opt.maxiter; ++it) {
            // projected-gradient convergence test
            double pgn = 0.0;
#pragma unroll
            for (int l = 0; l < DX; ++l) {
                double t = S.x[l] - S.g[l];
                t = fmin(fmax(t, pb.lower[l]), pb.upper[l]);
                pgn = fmax(pgn, fabs(S.x[l] - t));
            }
            if (pgn <= opt.pgtol) break;
            // search direction on the free variables
            double gf[DX], dd[DX];
#pragma unroll
            for (int l = 0; l < DX; ++l) {
                bool act = (S.x[l] <= pb.lower[l] && S.g[l] > 0.0) ||
                           (S.x[l] >= pb.upper[l] && S.g[l] < 0.0);
                gf[l] = act ? 0.0 : S.g[l];
            }
            double slope = 0.0, gnorm2 = 0.0;
#pragma unroll
            for (int l = 0; l < DX; ++l) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < DX; ++q) acc = fma(S.H[l * DX + q], gf[q], acc);
                dd[l] = (gf[l] == 0.0) ? 0.0 : -acc;
                slope = fma(dd[l], S.g[l], slope);
                gnorm2 = fma(gf[l], gf[l], gnorm2);
            }
            if (!(slope < 0.0)) {  // not a descent direction: restart from steepest descent
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int e = 0; e < DX * DX; ++e) S.H[e] = ((e / DX) == (e % DX)) ? 1.0 : 0.0;
                }
                __syncwarp();
                first_update = true;
                slope = -gnorm2;
#pragma unroll
                for (int l = 0; l < DX; ++l) dd[l] = -gf[l];
                if (gnorm2 == 0.0) break;
            }
            double dnorm = 0.0;
#pragma unroll
            for (int l = 0; l < DX; ++l) dnorm = fma(dd[l], dd[l], dnorm);
            dnorm = sqrt(dnorm);
            double step = (it == 0 || first_update) ? fmin(1.0, 1.0 / dnorm) : 1.0;
            // Armijo backtracking along the projected path
            bool accepted = false;
            double phin = 0.0;
            for (int ls = 0; ls < opt.max_ls; ++ls) {
                double xn[DX];
                double decr = 0.0;
#pragma unroll
                for (int l = 0; l < DX; ++l) {
                    double t = fma(step, dd[l], S.x[l]);
                    t = fmin(fmax(t, pb.lower[l]), pb.upper[l]);
                    xn[l] = t;
                    decr = fma(S.g[l], t - S.x[l], decr);
                }
                AbOut<DX> o;
                eval_ab_warp<DX, DW, STAGE, MT, true, WD>(ep, u, xn, s_ws, o, Z * u.inv_den, wd);
                ++evals;
                phin = -(o.a + Z * o.b);
                if (phin <= S.phi + c1 * decr) {
                    accepted = true;
                    __syncwarp();
                    if (lane == 0) {
#pragma unroll
                        for (int l = 0; l < DX; ++l) {
                            S.xn[l] = xn[l];
                            S.gn[l] = -(o.ga[l] + Z * o.gb[l]);
                        }
                    }
                    __syncwarp();
                    break;
                }
                // safeguarded quadratic interpolation through phi(0), phi'(0) = decr/step * step..., phi(step):
                // minimiser of q(t) = phi + decr_t t + c t^2 along the projected path, kept in
                // [0.1, 0.5] of the current step
                double denom_q = 2.0 * (phin - S.phi - decr);
                double shrink = (denom_q > 0.0) ? (-decr / denom_q) : 0.5;
                shrink = fmin(fmax(shrink, 0.1), 0.5);
                if (!(shrink == shrink)) shrink = 0.5;
                step *= shrink;
            }
            if (!accepted) break;
            // BFGS update of the inverse Hessian approximation
            double sv[DX], yv[DX];
            double sy = 0.0, yy = 0.0, ss = 0.0;
#pragma unroll
            for (int l = 0; l < DX; ++l) {
                sv[l] = S.xn[l] - S.x[l];
                yv[l] = S.gn[l] - S.g[l];
                sy = fma(sv[l], yv[l], sy);
                yy = fma(yv[l], yv[l], yy);
                ss = fma(sv[l], sv[l], ss);
            }
            const double phi_old = S.phi;
            double Hn[DX * DX];
            bool upd = sy > 1e-10 * sqrt(ss) * sqrt(yy);
            if (upd) {
                double Hc[DX * DX];
#pragma unroll
                for (int e = 0; e < DX * DX; ++e) Hc[e] = S.H[e];
                if (first_update) {
                    double sc = sy / yy;
#pragma unroll
                    for (int e = 0; e < DX * DX; ++e) Hc[e] = ((e / DX) == (e % DX)) ? sc : 0.0;
                }
                const double rho = 1.0 / sy;
                double Hy[DX];
                double yHy = 0.0;
#pragma unroll
                for (int l = 0; l < DX; ++l) {
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < DX; ++q) acc = fma(Hc[l * DX + q], yv[q], acc);
                    Hy[l] = acc;
                }
#pragma unroll
                for (int l = 0; l < DX; ++l) yHy = fma(yv[l], Hy[l], yHy);
                // H+ = H - rho (s Hy^T + Hy s^T) + rho (1 + rho yHy) s s^T   (H symmetric)
                const double cc = rho * (1.0 + rho * yHy);
#pragma unroll
                for (int l = 0; l < DX; ++l)
#pragma unroll
                    for (int q = 0; q < DX; ++q)
                        Hn[l * DX + q] = Hc[l * DX + q] - rho * (sv[l] * Hy[q] + Hy[l] * sv[q]) +
                                         cc * sv[l] * sv[q];
            }
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int l = 0; l < DX; ++l) {
                    S.x[l] = S.xn[l];
                    S.g[l] = S.gn[l];
                }
                S.phi = phin;
                if (upd) {
#pragma unroll
                    for (int e = 0; e < DX * DX; ++e) S.H[e] = Hn[e];
                }
            }
            __syncwarp();
            if (upd) first_update = false;
            // relative-decrease test of L-BFGS-B: (f_k - f_{k+1}) / max(|f_k|, |f_{k+1}|, 1)
            double denom = fmax(fmax(fabs(phi_old), fabs(phin)), 1.0);
            if ((phi_old - phin) <= opt.factr * DBL_EPSILON * denom) break;
        }
        if (lane == 0) {
            double* rp = s_run + ((size_t)zi * n_starts + sidx) * (1 + DX);
            rp[0] = -S.phi;
#pragma unroll
            for (int l = 0; l < DX; ++l) rp[1 + l] = S.x[l];
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- phase 3: best run per Z sample, then the vertices (sbo.py:343-364) ---------------------
    for (int zi = threadIdx.x; zi < nz; zi += blockDim.x) {
        double best = s_run[(size_t)zi * n_starts * (1 + DX)];
        int bi = 0;
        for (int sidx = 1; sidx < n_starts; ++sidx) {  // np.argmax: first maximum wins
            double v = s_run[((size_t)zi * n_starts + sidx) * (1 + DX)];
            if (v > best) {
                best = v;
                bi = sidx;
            }
        }
        double arg[DX];
#pragma unroll
        for (int l = 0; l < DX; ++l) arg[l] = s_run[((size_t)zi * n_starts + bi) * (1 + DX) + 1 + l];
        if (opt.use_vertices) {
            const double Z = z[z0 + zi];
            double vb = -DBL_MAX;
            int vi = 0;
            for (int v = 0; v < NV; ++v) {
                const double* op = s_pre + (size_t)(n_starts + v) * (2 + 2 * DX);
                double val = op[0] + Z * op[1];
                if (v == 0 || val > vb) {
                    vb = val;
                    vi = v;
                }
            }
            if (vb > best) {
                best = vb;
#pragma unroll
                for (int l = 0; l < DX; ++l) arg[l] = s_prex[(n_starts + vi) * DX + l];
            }
        }
        out_max[(int64_t)unit * opt.n_z + z0 + zi] = best;
#pragma unroll
        for (int l = 0; l < DX; ++l) out_arg[((int64_t)unit * opt.n_z + z0 + zi) * DX + l] = arg[l];
    }
    if (n_evals != nullptr && lane == 0 && evals) atomicAdd(n_evals + unit, evals);
}

static size_t inner_smem_bytes(const bqo_stage_c* pb, const bqo_inner_opt* opt, bool stage,
                               size_t state_bytes, bool wd = false) {
    // staged doubles per training point: D + 2 rows, or one record of odd length with the table
    const int D = wd ? (((pb->dx + 2) | 1) - 2) : pb->dx + pb->dw;
    const int M = pb->dw > 0 ? pb->M : 1;
    const int NV = 1 << pb->dx;
    const int n_pre = opt->n_starts + (opt->use_vertices ? NV : 0);
    size_t d = BQO_EXP_TABLE;
    if (stage) d += (size_t)(D + 2) * pb->n;
    d += (size_t)M * pb->dw;
    d += (size_t)n_pre * (2 + 2 * pb->dx) + (size_t)n_pre * pb->dx;
    d += (size_t)opt->z_per_cta * opt->n_starts * (1 + pb->dx);
    return d * sizeof(double) + state_bytes * IM_WARPS + 16;
}

extern "C" int bqo_inner_maximize_batched(const bqo_stage_c* pb, const bqo_inner_opt* opt,
                                          const double* z, const double* starts, double* out_max,
                                          double* out_arg, unsigned long long* n_evals,
                                          void* stream) {
    BQO_CHECK_ARG(check_stage_c(pb), 1);
    BQO_CHECK_ARG(opt != nullptr && opt->n_z > 0 && opt->n_starts > 0 && opt->z_per_cta > 0 &&
                      opt->maxiter >= 0 && opt->max_ls > 0,
                  2);
    BQO_CHECK_ARG(pb->lower != nullptr && pb->upper != nullptr, 1);
    BQO_CHECK_ARG(pb->n_units > 0, 1);
    BQO_CHECK_ARG(z != nullptr, 3);
    BQO_CHECK_ARG(starts != nullptr, 4);
    BQO_CHECK_ARG(out_max != nullptr, 5);
    BQO_CHECK_ARG(out_arg != nullptr, 6);
    BQO_CHECK_ARG(pb->dx <= 6, 1);
    cudaStream_t cs = (cudaStream_t)stream;
    const int zc_per_unit = (opt->n_z + opt->z_per_cta - 1) / opt->z_per_cta;
    const int64_t grid64 = (int64_t)pb->n_units * zc_per_unit;
    BQO_CHECK_ARG(grid64 < 2147483647LL, 1);
    const int grid = (int)grid64;
    if (n_evals) cudaMemsetAsync(n_evals, 0, sizeof(unsigned long long) * pb->n_units, cs);
#define LAUNCH_IM_K(DXV, DWV, STG, MTV, WDV, SB)                                                 \
    do {                                                                                         \
        bqo_smem_opt_in<inner_maximize_kernel<DXV, DWV, STG, MTV, WDV>>(SB);                     \
        BQO_L, inner_maximize_kernel<DXV, DWV, STG, MTV, WDV><<<grid, IM_WARPS * 32, (SB), cs>>>(       \
            *pb, *opt, z, starts, out_max, out_arg, n_evals);                                    \
    } while (0)
#define LAUNCH_IM(DXV, DWV)                                                                      \
    do {                                                                                         \
        const bool use_wd = (DWV > 0) && pb->wdist != nullptr && pb->M == 10;                    \
        size_t sb = inner_smem_bytes(pb, opt, true, sizeof(RunState<DXV>), use_wd);              \
        if (sb <= 227 * 1024) {                                                                  \
            /* M = 10 (N_SAMPLES of sbo/lib/expectations.py:7) gets the fully unrolled build */  \
            if (use_wd) LAUNCH_IM_K(DXV, DWV, true, (DWV > 0 ? 10 : 0), (DWV > 0), sb);          \
            else if (DWV > 0 && pb->M == 10)                                                     \
                LAUNCH_IM_K(DXV, DWV, true, (DWV > 0 ? 10 : 0), false, sb);                      \
            else LAUNCH_IM_K(DXV, DWV, true, 0, false, sb);                                      \
        } else {                                                                                 \
            /* training set too large for shared memory: operands stream from L2; the      */  \
            /* unrolled W loop and the distance table still apply                            */  \
            sb = inner_smem_bytes(pb, opt, false, sizeof(RunState<DXV>));                        \
            if (sb > 227 * 1024) return -2;                                                      \
            if (use_wd) LAUNCH_IM_K(DXV, DWV, false, (DWV > 0 ? 10 : 0), (DWV > 0), sb);         \
            else LAUNCH_IM_K(DXV, DWV, false, 0, false, sb);                                     \
        }                                                                                        \
    } while (0)
    SC_DISPATCH(pb, LAUNCH_IM);
#undef LAUNCH_IM
#undef LAUNCH_IM_K
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// gradient_vector_b: warp per (unit, point).
//   beta1 = 1/den, beta2 = B(p,c) - B(p,X) s2, beta3_l = grad_{c_l} B(p,c) - sum_j V[l][j] B(p,X_j),
//   out_l = beta1 beta3_l + 0.5 beta1^3 beta2 beta4_l      (bayesian_quadrature.py:1506-1517)
template <int DX, int DW>
__global__ void __launch_bounds__(AB_WARPS * 32)
    grad_vector_b_kernel(bqo_stage_c pb, const double* __restrict__ pts, int npts,
                         const int32_t* __restrict__ pt_wset, double* __restrict__ out,
                         int32_t* __restrict__ is_nan) {
    constexpr int D = DX + DW;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gp = (int64_t)blockIdx.x * AB_WARPS + warp;
    bqo_exp_table_init(dyn_smem);
    __syncthreads();
    if (gp >= (int64_t)pb.n_units * npts) return;
    const int unit = (int)(gp / npts);
    const int M = (DW > 0) ? pb.M : 1;
    double* ws = dyn_smem + BQO_EXP_TABLE + warp * (M * (DW > 0 ? DW : 1));
    UnitCtx<DX, DW> u;
    unit_ctx_init<DX, DW>(u, pb, unit);
    const int k = pb.unit_k[unit], c = pb.unit_c[unit];
    const double den = pb.den[(int64_t)k * pb.C + c];
    if (den * den == 0.0) {  // beta_1[0] == 0 -> np.nan (bayesian_quadrature.py:1483-1484)
        if (lane == 0) {
            if (gp % npts == 0) is_nan[unit] = 1;
            for (int l = 0; l < D; ++l) out[gp * D + l] = nan("");
        }
        return;
    }
    if (lane == 0 && gp % npts == 0) is_nan[unit] = 0;
    if (DW > 0) {
        // default: the W set of the unit's candidate, like the inner maximiser
        int wsi = pt_wset ? pt_wset[gp] : (c % pb.n_wsets);
        scale_wset<DX, DW>(u, pb.wsets + (int64_t)wsi * pb.M * DW, ws, BQO_PRESCALE);
        __syncwarp();
    }
    // same fast kernel evaluation as eval_ab_warp: coordinates in units of K (bqo_common.cuh)
    double xsc[DX];
#pragma unroll
    for (int l = 0; l < DX; ++l) xsc[l] = (pts[gp * DX + l] / u.ls[l]) * BQO_PRESCALE;
    const double* V = u.wb + pb.n;  // rows 1..D of the unit: Sigma^-1 grad_l gamma (q-weighted)
    double Sb = 0.0, SV[D];
#pragma unroll
    for (int l = 0; l < D; ++l) SV[l] = 0.0;
    for (int j = lane; j < u.n; j += 32) {
        double Dx = 1e-300;
#pragma unroll
        for (int l = 0; l < DX; ++l) {
            double df = xsc[l] - u.xs[l * u.ldx + j] * BQO_PRESCALE;
            Dx = fma(df, df, Dx);
        }
        double E0 = 0.0, E1 = 0.0, H = 0.0;
        if (DW > 0) {
            double xw[DW > 0 ? DW : 1];
#pragma unroll
            for (int l = 0; l < DW; ++l) xw[l] = u.xs[(DX + l) * u.ldx + j] * BQO_PRESCALE;
            for (int m = 0; m < M; ++m) {
                double q2v[1], ev[1], t1v[1];
                q2v[0] = Dx;
#pragma unroll
                for (int l = 0; l < DW; ++l) {
                    double dw = ws[m * DW + l] - xw[l];
                    q2v[0] = fma(dw, dw, q2v[0]);
                }
                bqo_matern52_q_v<1>(q2v, SmemExpTab(), ev, t1v);
                E0 += ev[0];
                E1 = fma(t1v[0], ev[0], E1);
                H = fma(q2v[0], ev[0], H);
            }
        } else {
            double q2v[1], ev[1], t1v[1];
            q2v[0] = Dx;
            bqo_matern52_q_v<1>(q2v, SmemExpTab(), ev, t1v);
            E0 = ev[0];
            E1 = t1v[0] * ev[0];
            H = q2v[0] * ev[0];
        }
        const double ksum = fma(H, BQO_H_TO_K, fma(E1, BQO_LN2_N, E0));
        Sb = fma(ksum, u.wb[j], Sb);
#pragma unroll
        for (int l = 0; l < D; ++l) SV[l] = fma(ksum, V[(int64_t)l * pb.n + j], SV[l]);
    }
    // candidate terms: k((p, w_m), c) and its gradient with respect to c
    double Kc = 0.0, GC[D];
#pragma unroll
    for (int l = 0; l < D; ++l) GC[l] = 0.0;
    for (int m = lane; m < M; m += 32) {
        double df[D];
        double q2v[1], ev[1], t1v[1];
        q2v[0] = 1e-300;
#pragma unroll
        for (int l = 0; l < DX; ++l) df[l] = u.cs[l] * BQO_PRESCALE - xsc[l];
#pragma unroll
        for (int l = 0; l < DW; ++l) df[DX + l] = u.cs[DX + l] * BQO_PRESCALE - ws[m * DW + l];
#pragma unroll
        for (int l = 0; l < D; ++l) q2v[0] = fma(df[l], df[l], q2v[0]);
        bqo_matern52_q_v<1>(q2v, SmemExpTab(), ev, t1v);
        const double gm = fma(t1v[0], BQO_LN2_N, 1.0) * ev[0];
        Kc += fma(q2v[0] * ev[0], BQO_H_TO_K, gm);
        // k'(r)/r (c - p)/ls = -(5/3) gm (df'/K): the remaining 1/ls is applied below
        const double g = BQO_GRAD_FAC * gm;
#pragma unroll
        for (int l = 0; l < D; ++l) GC[l] = fma(g, df[l], GC[l]);
    }
    Sb = bqo_warp_sum(Sb);
    Kc = bqo_warp_sum(Kc);
    const double beta1 = 1.0 / den;
    const double beta2 = u.scale * (u.wc * Kc - Sb);
#pragma unroll
    for (int l = 0; l < D; ++l) {
        double sv = bqo_warp_sum(SV[l]);
        double gc = bqo_warp_sum(GC[l]);
        double beta3 = u.scale * (u.wc * gc * u.inv_ls[l] - sv);
        double beta4 = pb.beta4[((int64_t)k * pb.C + c) * D + l];
        double g = beta1 * beta3 + 0.5 * (beta1 * beta1 * beta1) * beta2 * beta4;
        if (lane == 0) out[gp * D + l] = g;
    }
}

// Staged variant for the Monte-Carlo-W models: CTA = unit, warps loop over the unit's points.
// The x rows (in units of K), s2 and the exp table are staged once per unit as in the inner
// maximiser, the W part of the distances comes from the bqo_wdist_precompute table (default W
// set of the unit), the D rows Sigma^-1 grad_l gamma are streamed from L2.
template <int DX, int DW>
__global__ void __launch_bounds__(IM_WARPS * 32, 1)
    grad_vector_b_staged_kernel(bqo_stage_c pb, const double* __restrict__ pts, int npts,
                                double* __restrict__ out, int32_t* __restrict__ is_nan) {
    constexpr int D = DX + DW;
    constexpr int MT = 10;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int unit = blockIdx.x;
    const int n = pb.n;
    double* s_tab = dyn_smem;
    double* s_xs = dyn_smem + BQO_EXP_TABLE;  // [DX][n]
    double* s_wb = s_xs + (size_t)DX * n;     // [n]
    double* s_ws = s_wb + n;                  // [MT][DW]
    __shared__ UnitCtx<DX, DW> s_u;
    if (threadIdx.x == 0) unit_ctx_init<DX, DW>(s_u, pb, unit);
    bqo_exp_table_init(s_tab);
    __syncthreads();
    const UnitCtx<DX, DW>& u = s_u;
    const int k = pb.unit_k[unit], c = pb.unit_c[unit];
    const double den = pb.den[(int64_t)k * pb.C + c];
    if (den * den == 0.0) {  // beta_1[0] == 0 -> np.nan (bayesian_quadrature.py:1483-1484)
        if (threadIdx.x == 0) is_nan[unit] = 1;
        for (int e = threadIdx.x; e < npts * D; e += blockDim.x)
            out[(int64_t)unit * npts * D + e] = nan("");
        return;
    }
    if (threadIdx.x == 0) is_nan[unit] = 0;
    {
        const double* gxs = u.xs;
        const double* gwb = u.wb;
        for (int e = threadIdx.x; e < DX * n; e += blockDim.x) s_xs[e] = gxs[e] * BQO_PRESCALE;
        for (int e = threadIdx.x; e < n; e += blockDim.x) s_wb[e] = gwb[e];
        const double* wsrc = pb.wsets + (size_t)(c % pb.n_wsets) * MT * DW;
        for (int e = threadIdx.x; e < MT * DW; e += blockDim.x)
            s_ws[e] = (wsrc[e] / u.ls[DX + (e % (DW > 0 ? DW : 1))]) * BQO_PRESCALE;
    }
    __syncthreads();
    const size_t nblk = (size_t)(n + 31) / 32;
    const double* wd = pb.wdist + (((size_t)k * pb.n_wsets + (c % pb.n_wsets)) * nblk) *
                                      ((size_t)MT * 32) + lane;
    const double* V = u.wb + n;  // rows 1..D of the unit: Sigma^-1 grad_l gamma (q-weighted)
    const double beta1 = 1.0 / den;
    for (int p = warp; p < npts; p += IM_WARPS) {
        const int64_t gp = (int64_t)unit * npts + p;
        double xsc[DX];
#pragma unroll
        for (int l = 0; l < DX; ++l) xsc[l] = (pts[gp * DX + l] / u.ls[l]) * BQO_PRESCALE;
        double Sb = 0.0, SV[D];
#pragma unroll
        for (int l = 0; l < D; ++l) SV[l] = 0.0;
        for (int j = lane; j < n; j += 32) {
            double Dx = 0.0;
#pragma unroll
            for (int l = 0; l < DX; ++l) {
                const double df = xsc[l] - s_xs[l * n + j];
                Dx = fma(df, df, Dx);
            }
            Dx = __hiloint2double((int)min((unsigned)__double2hiint(Dx), (unsigned)BQO_U2_PART_HI),
                                  __double2loint(Dx));
            const double* wj = wd + (size_t)(j >> 5) * (MT * 32);
            double E0 = 0.0, E1 = 0.0, H = 0.0;
#pragma unroll
            for (int m = 0; m < MT; m += 2) {
                double q2v[2], ev[2], t1v[2];
                q2v[0] = Dx + __ldg(wj + m * 32);
                q2v[1] = Dx + __ldg(wj + (m + 1) * 32);
                bqo_matern52_q_v<2, SmemExpTab, false>(q2v, SmemExpTab(), ev, t1v);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    E0 += ev[q];
                    E1 = fma(t1v[q], ev[q], E1);
                    H = fma(q2v[q], ev[q], H);
                }
            }
            const double ksum = fma(H, BQO_H_TO_K, fma(E1, BQO_LN2_N, E0));
            Sb = fma(ksum, s_wb[j], Sb);
#pragma unroll
            for (int l = 0; l < D; ++l) SV[l] = fma(ksum, __ldg(V + (int64_t)l * n + j), SV[l]);
        }
        // candidate terms: k((p, w_m), c) and its gradient with respect to c
        double Kc = 0.0, GC[D];
#pragma unroll
        for (int l = 0; l < D; ++l) GC[l] = 0.0;
        for (int m = lane; m < MT; m += 32) {
            double df[D];
            double q2v[1], ev[1], t1v[1];
            q2v[0] = 1e-300;
#pragma unroll
            for (int l = 0; l < DX; ++l) df[l] = u.cs[l] * BQO_PRESCALE - xsc[l];
#pragma unroll
            for (int l = 0; l < DW; ++l) df[DX + l] = u.cs[DX + l] * BQO_PRESCALE - s_ws[m * DW + l];
#pragma unroll
            for (int l = 0; l < D; ++l) q2v[0] = fma(df[l], df[l], q2v[0]);
            bqo_matern52_q_v<1>(q2v, SmemExpTab(), ev, t1v);
            const double gm = fma(t1v[0], BQO_LN2_N, 1.0) * ev[0];
            Kc += fma(q2v[0] * ev[0], BQO_H_TO_K, gm);
            const double g = BQO_GRAD_FAC * gm;
#pragma unroll
            for (int l = 0; l < D; ++l) GC[l] = fma(g, df[l], GC[l]);
        }
        Sb = bqo_warp_sum(Sb);
        Kc = bqo_warp_sum(Kc);
        const double beta2 = u.scale * (u.wc * Kc - Sb);
#pragma unroll
        for (int l = 0; l < D; ++l) {
            const double sv = bqo_warp_sum(SV[l]);
            const double gc = bqo_warp_sum(GC[l]);
            const double beta3 = u.scale * (u.wc * gc * u.inv_ls[l] - sv);
            const double beta4 = pb.beta4[((int64_t)k * pb.C + c) * D + l];
            const double g = beta1 * beta3 + 0.5 * (beta1 * beta1 * beta1) * beta2 * beta4;
            if (lane == 0) out[gp * D + l] = g;
        }
    }
}

extern "C" int bqo_grad_vector_b_batched(const bqo_stage_c* pb, const double* pts, int32_t npts,
                                         const int32_t* pt_wset, double* out, int32_t* is_nan,
                                         void* stream) {
    BQO_CHECK_ARG(check_stage_c(pb), 1);
    BQO_CHECK_ARG(pb->beta4 != nullptr && pb->n_units > 0, 1);
    BQO_CHECK_ARG(pts != nullptr, 2);
    BQO_CHECK_ARG(npts > 0, 3);
    BQO_CHECK_ARG(out != nullptr, 5);
    BQO_CHECK_ARG(is_nan != nullptr, 6);
    cudaStream_t cs = (cudaStream_t)stream;
    // staged kernel: Monte-Carlo W with the distance table, the default W set, enough points per
    // unit to pay for the staging, and the unit's operands fit shared memory
    {
        const size_t sb = sizeof(double) * ((size_t)BQO_EXP_TABLE + (size_t)(pb->dx + 1) * pb->n +
                                            (size_t)10 * pb->dw) + 16;
        if (pb->dw > 0 && pb->wdist != nullptr && pb->M == 10 && pt_wset == nullptr &&
            npts >= IM_WARPS && sb <= 227 * 1024) {
#define LAUNCH_GVBS(DXV, DWV)                                                                    \
    do {                                                                                         \
        if (DWV > 0) {                                                                           \
            bqo_smem_opt_in<grad_vector_b_staged_kernel<DXV, (DWV > 0 ? DWV : 1)>>(sb);          \
            BQO_L, grad_vector_b_staged_kernel<DXV, (DWV > 0 ? DWV : 1)>                                \
                <<<pb->n_units, IM_WARPS * 32, sb, cs>>>(*pb, pts, npts, out, is_nan);           \
        }                                                                                        \
    } while (0)
            SC_DISPATCH(pb, LAUNCH_GVBS);
#undef LAUNCH_GVBS
            BQO_RETURN_LAST_ERROR();
        }
    }
    int64_t total = (int64_t)pb->n_units * npts;
    int grid = (int)((total + AB_WARPS - 1) / AB_WARPS);
    size_t smem = sizeof(double) * (BQO_EXP_TABLE + AB_WARPS * (size_t)(pb->M > 0 ? pb->M : 1) * (pb->dw > 0 ? pb->dw : 1));
#define LAUNCH_GVB(DXV, DWV) \
    BQO_L, grad_vector_b_kernel<DXV, DWV><<<grid, AB_WARPS * 32, smem, cs>>>(*pb, pts, npts, pt_wset, out, is_nan)
    SC_DISPATCH(pb, LAUNCH_GVB);
#undef LAUNCH_GVB
    BQO_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------------------------
// Monte-Carlo reduction over Z samples and hyper-samples; fixed order, one warp per output.
__global__ void acq_reduce_kernel(const double* __restrict__ out_max, const double* __restrict__ gvb,
                                  const int32_t* __restrict__ is_nan, const double* __restrict__ z,
                                  const double* __restrict__ max_mean, int C, int S, int n_z,
                                  int dim_g, double* __restrict__ value,
                                  double* __restrict__ grad) {
    // one warp per output; lanes stride over the Z samples, fixed shuffle tree (bit-reproducible)
    const int lane = threadIdx.x & 31;
    int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int per = 1 + dim_g;
    if (idx >= C * per) return;
    int c = idx / per, q = idx % per;
    if (q == 0) {
        // evaluations[l] = mean_k( mean_i max - max_mean_k )   (sbo.py:646,668)
        double acc = 0.0;
        for (int k = 0; k < S; ++k) {
            const double* mx = out_max + ((int64_t)c * S + k) * n_z;
            double s = 0.0;
            for (int i = lane; i < n_z; i += 32) s += mx[i];
            s = bqo_warp_sum(s);
            double mm = max_mean ? max_mean[k] : 0.0;
            acc += s / (double)n_z - mm;
        }
        if (lane == 0) value[c] = acc / (double)S;
    } else if (grad != nullptr && gvb != nullptr) {
        int l = q - 1;
        bool bad = false;
        double acc = 0.0;
        for (int k = 0; k < S; ++k) {
            int u = c * S + k;
            if (is_nan && is_nan[u]) {
                bad = true;
                break;
            }
            const double* gp = gvb + ((int64_t)u * n_z) * dim_g + l;
            double s = 0.0;
            for (int i = lane; i < n_z; i += 32) s = fma(gp[(int64_t)i * dim_g], z[i], s);
            s = bqo_warp_sum(s);
            acc += s / (double)n_z;
        }
        if (lane == 0) grad[(int64_t)c * dim_g + l] = bad ? nan("") : acc / (double)S;
    }
}

extern "C" int bqo_acq_reduce(const double* out_max, const double* gvb, const int32_t* is_nan,
                              const double* z, const double* max_mean, int32_t C, int32_t S,
                              int32_t n_z, int32_t dim_g, double* value, double* grad,
                              void* stream) {
    BQO_CHECK_ARG(out_max != nullptr, 1);
    BQO_CHECK_ARG(z != nullptr, 4);
    BQO_CHECK_ARG(C > 0 && S > 0 && n_z > 0 && dim_g >= 0, 6);
    BQO_CHECK_ARG(value != nullptr, 10);
    int total = C * (1 + dim_g);
    BQO_L, acq_reduce_kernel<<<(total + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
        out_max, gvb, is_nan, z, max_mean, C, S, n_z, dim_g, value, grad);
    BQO_RETURN_LAST_ERROR();
}
