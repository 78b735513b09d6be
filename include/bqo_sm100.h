/*
 * bqo_sm100.h -- C ABI of libbqo_sm100.so: the B200 (sm_100a) implementation of the
 * GP-inference + BQO/SBO/EI acquisition hot path of
 * toscanosaul/bayesian_quadrature_optimization (package `stratified_bayesian_optimization`,
 * abbreviated `sbo/` below; line numbers refer to the reference tree).
 *
 * The reference is pure Python, so it has no FFI of its own.  Every entry point below
 * replaces the numpy/scipy body of one or more reference methods; the Python host classes in
 * `bayesian_quadrature_optimization_b200/` keep the reference's signatures and call these
 * functions through ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - all array arguments are DEVICE pointers owned by the caller (torch tensors on the host
 *     side); FP64 unless stated; row-major; nothing is allocated internally except where a
 *     `*_workspace_bytes` query says how much scratch the caller must pass;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream
 *     unless documented otherwise (only the roofline probes synchronise);
 *   - return value: 0 = ok, <0 = bad argument (-(1-based argument index)), >0 = CUDA error
 *     code (cudaError_t) from the launch;
 *   - hyper-sample vector layout (sbo/models/gp_fitting_gaussian.py:637-655):
 *         theta = [var_noise, mean, kernel parameters...]
 *     kernel parameters: MATERN52 = [ls_1..ls_d]              (sbo/kernels/matern52.py:104-109)
 *                        SCALED   = [ls_1..ls_d, sigma2]       (sbo/kernels/scaled_kernel.py:109-115)
 *                        PRODUCT  = [ls_1..ls_dx, task params] (sbo/kernels/product_kernels.py:209-218)
 *     task params: T(T+1)/2 log-Cholesky entries, or min(T,2) when same_correlation
 *                        (sbo/kernels/tasks_kernel.py:198-239, sbo/lib/util.py:142-148).
 */
#ifndef BQO_SM100_H
#define BQO_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BQO_ABI_VERSION 1
#define BQO_MAX_DIM 16 /* maximum number of columns under the Matern-5/2 factor */

enum bqo_kernel_kind {
    BQO_KIND_MATERN52 = 0,        /* sbo/kernels/matern52.py */
    BQO_KIND_SCALED_MATERN52 = 1, /* sbo/kernels/scaled_kernel.py wrapping Matern52 */
    BQO_KIND_PRODUCT_TASKS = 2    /* sbo/kernels/product_kernels.py: Matern52(x) * Tasks(w) */
};

/* Static description of the covariance function of the GP (one per model). */
typedef struct bqo_kernel_desc {
    int32_t kind;      /* enum bqo_kernel_kind */
    int32_t dim;       /* columns of one input point (x columns, then w / task column) */
    int32_t dim_m;     /* leading columns under the Matern-5/2 factor (dim, or dim-1 for PRODUCT) */
    int32_t n_tasks;   /* T for PRODUCT, else 0 */
    int32_t same_corr; /* 1: same_correlation task parameterisation */
    int32_t n_params;  /* kernel parameters per hyper-sample (p); theta has 2+p entries */
} bqo_kernel_desc;

/* ---- library / device ------------------------------------------------------------------ */
int bqo_abi_version(void);
int bqo_set_device(int device);
/* Kernels launched by this library in this process so far (every launch site counts itself;
 * bench.py reports the difference over its timed region as `gpu_launches`). */
int64_t bqo_launch_count(void);
/* Number of doubles of the per-hyper-sample block written by bqo_hyper_prepare. */
int64_t bqo_hyper_stride(const bqo_kernel_desc* desc);

/* FP64 roofline probes (no reference counterpart; SURVEY 8d asks for a measured FP64 peak).
 * kind 0: dependent-chain-free DFMA loop, kind 1: DMMA m8n8k4 loop.  Synchronous.
 * Writes elapsed milliseconds and the FP64 flops executed.  `workspace`: device scratch of
 * bqo_fp64_peak_probe_workspace_bytes() bytes (the library allocates nothing itself). */
int64_t bqo_fp64_peak_probe_workspace_bytes(void);
int bqo_fp64_peak_probe(int kind, int iters, void* workspace, double* ms_out, double* flops_out);
/* HBM copy probe: copies `bytes` bytes device to device `reps` times, returns the best ms. */
int bqo_hbm_copy_probe(void* dst, const void* src, int64_t bytes, int reps, double* best_ms_out);

/* ---- hyper-sample preparation ---------------------------------------------------------- */
/* theta[S][2+p] -> hyp[S][stride]: var_noise, mean, sigma2, ls, 1/ls, task covariance C and its
 * factor (TasksKernel.compute_cov_matrix, sbo/kernels/tasks_kernel.py:198-239). */
int bqo_hyper_prepare(const bqo_kernel_desc* desc, const double* theta, int32_t S, double* hyp,
                      void* stream);
/* Xs[s][l][i] = X[i][l] / ls_s[l]  (the division of Distances.dist_square_length_scale,
 * sbo/lib/distances.py:20-25), l < dim_m, structure-of-arrays per hyper-sample;
 * tid[i] = (int) X[i][dim-1] for PRODUCT (astype(int), sbo/kernels/tasks_kernel.py:252-253).
 * `tid` may be NULL for the other kinds.  `ldx` = row stride of X in doubles. */
int bqo_scale_points(const bqo_kernel_desc* desc, const double* hyp, int32_t S, const double* X,
                     int32_t n, int32_t ldx, double* Xs, int32_t* tid, void* stream);

/* ---- kernel matrices (a1,a2,a6,a8,a9) --------------------------------------- */
/* K[s] = k_theta_s(X, X) (+ diag(noise_obs) + var_noise_s I when add_noise != 0):
 * GPFittingGaussian.evaluate_cov / the Sigma of _chol_cov_including_noise
 * (sbo/models/gp_fitting_gaussian.py:701-770).  `extra_diag[S]` (nullable) is added to the
 * diagonal too (the jitter of sbo/lib/la_functions.py:27-33). */
int bqo_kernel_assemble_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                                const double* Xs, const int32_t* tid, int32_t n,
                                const double* noise_obs, int32_t add_noise,
                                const double* extra_diag, double* K, void* stream);
/* dK[s][j] = d k_theta_s(X,X) / d kernel-parameter j, j < p: GPFittingGaussian.evaluate_grad_cov
 * (sbo/models/gp_fitting_gaussian.py:800-829; a3,a7,a8,a9). */
int bqo_kernel_grad_params_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                                   const double* Xs, const int32_t* tid, int32_t n, double* dK,
                                   void* stream);
/* KP[s][q][i] = k_theta_s(P_q, X_i), and when dKP != NULL
 * dKP[s][q][l][i] = d k_theta_s(P_q, X_i) / d P_q[l], l < dim_m
 * (evaluate_cross_cov / evaluate_grad_cross_cov_respect_point,
 * sbo/models/gp_fitting_gaussian.py:1107-1170; a2,a4).  P is [m][dim] raw (unscaled).
 * Rows are laid out so that (KP row, then its dim_m gradient rows) of a point are
 * consecutive when `interleave != 0`: row index q*(1+dim_m) + {0, 1+l}; otherwise KP and dKP
 * are separate arrays.  With interleave, pass KP = base pointer and dKP = NULL-ignored. */
int bqo_cross_cov_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                          const double* Xs, const int32_t* tid, int32_t n, const double* P,
                          int32_t m, int32_t interleave, double* KP, double* dKP, void* stream);

/* Second derivatives of k(P_q, X_i) with respect to the point P_q
 * (GradientLSMatern52.hessian_respect_point, sbo/kernels/matern52.py:466-503, with
 * Distances.gradient_distance_length_scale_respect_point(second=True), sbo/lib/distances.py:70-116;
 * ScaledKernel / ProductKernels wrappers scaled_kernel.py:229-270, product_kernels.py:327-481).
 * out[S][m][dim_m][dim_m][n]; only the Matern coordinates are differentiated.  Finite at r = 0
 * (the reference divides by r there). */
int bqo_cross_cov_hessian_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                                  const double* Xs, const int32_t* tid, int32_t n, const double* P,
                                  int32_t m, double* out, void* stream);

/* ---- Cholesky / solves / likelihood (a10-a13, a17) ------------------------------------- */
/* In-place lower Cholesky of S row-major n x n matrices (lower triangle read; strict upper
 * triangle zeroed like scipy dpotrf clean=1, sbo/lib/la_functions.py:17).
 * info[s] = 0 ok, k>0: leading minor k not positive definite. */
int bqo_potrf_batched(double* A, int32_t n, int32_t S, int32_t* info, void* workspace,
                      void* stream);
/* Scratch of bqo_potrf_batched / bqo_chol_cov_batched / bqo_chol_cov_retry for S matrices of
 * order n (inverses of the 64 x 64 diagonal blocks and the progress counters of the persistent
 * kernel).  The content need not be initialised and is not needed after the call. */
int64_t bqo_potrf_workspace_bytes(int32_t n, int32_t S);
/* 1: factor with the multi-launch right-looking path of round 1 instead of the persistent kernel
 * (tests compare the two; also selected by the environment variable BQO_POTRF_LEGACY=1, and always
 * used for odd n).  Returns the previous setting. */
int bqo_potrf_set_multi_launch(int on);
/* Tuning aid (no reference counterpart): bqo_potrf_batched that also writes %globaltimer stamps of
 * the critical-path tasks of matrix 0 to prof[ceil(n/64)][2][8] (uint64, device, zeroed by the
 * caller): [j][0][.] diagonal task of column j, [j][1][.] tile (j+1, j). */
int bqo_potrf_profile(double* A, int32_t n, int32_t S, int32_t* info, void* workspace,
                      unsigned long long* prof, void* stream);
/* Solve L L^T x = b for m right-hand sides per matrix; Bt[s][r][i] holds right-hand side r
 * contiguously (RHS-major), overwritten by the solution (cho_solve, sbo/lib/la_functions.py:41-50). */
int bqo_potrs_batched(const double* L, int32_t n, int32_t S, double* Bt, int32_t m,
                      void* stream);
/* Sigma_s = K_s + noise, L_s = chol(Sigma_s): the first attempt of
 * cholesky(cov, max_tries) (sbo/lib/la_functions.py:8-38, called from
 * sbo/models/gp_fitting_gaussian.py:764), for all S hyper-samples, WITHOUT synchronising.
 * Device outputs: info[s] = dpotrf's code (0 ok, k > 0: leading minor k not positive definite),
 * diag_mean[s] = mean(diag Sigma_s) and diag_nonpos[s] = 1 when a diagonal entry is <= 0 (the
 * two inputs of the jitter policy).  When info[s] != 0 the caller continues the policy with
 * bqo_chol_cov_retry: "not positive definite matrix" if diag_nonpos, else jitter =
 * diag_mean * 1e-6 * 10^t for t = 0 .. max_tries - 1 (la_functions.py:19-38).
 * Winv (optional, [S][n][n]): the same launch also leaves W_s = L_s^-1 on the lower 64 x 64 tiles
 * (what the log-likelihood gradient needs to form Sigma^-1 = W^T W; pass it on to
 * bqo_loglik_grad_batched).  Winv_scratch: S n^2 doubles, needed only for odd n. */
int bqo_chol_cov_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                         const double* Xs, const int32_t* tid, int32_t n, const double* noise_obs,
                         double* L, int32_t* info, double* diag_mean, int32_t* diag_nonpos,
                         double* Winv, double* Winv_scratch, void* workspace, void* stream);
/* One retry of the policy for hyper-sample s: L_s = chol(Sigma_s + jitter I); info_one (device,
 * one int) receives the code; jitter_dev: one double of device scratch. */
int bqo_chol_cov_retry(const bqo_kernel_desc* desc, const double* hyp, int32_t s, const double* Xs,
                       const int32_t* tid, int32_t n, const double* noise_obs, double jitter,
                       double* jitter_dev, double* L, int32_t* info_one, void* workspace,
                       void* stream);
/* Append one training point to S factors in O(n^2) each (the "updated cache in a smart way" TODO of
 * add_points_evaluations, sbo/models/gp_fitting_gaussian.py:492-511, which refactorises instead).
 * knew[s][0..n) = Sigma'_s(new, old points), knew[s][n] = Sigma'_s(new, new) including noise (and the
 * jitter of the old factor, if any).  L_new: [S][n+1][n+1], lower, strict upper triangle zero.
 * info[s] = 0, or n+1 when the extended matrix is not positive definite (device array).
 * workspace: S*n doubles. */
int bqo_chol_append_batched(const double* L_old, int32_t n, int32_t S, const double* knew,
                            double* L_new, int32_t* info, void* workspace, void* stream);
/* alpha[s] = Sigma_s^{-1} (y - mean_s)  (sbo/models/gp_fitting_gaussian.py:1241-1244). */
int bqo_alpha_batched(const double* hyp, int64_t hyp_stride, const double* L, int32_t n, int32_t S,
                      const double* y, double* alpha, void* stream);
/* The same alpha from W_s = L_s^-1 (the Winv of bqo_chol_cov_batched): alpha = W^T (W (y - mean)),
 * two bandwidth-bound passes instead of two triangular solves.  work: S * n doubles. */
int bqo_alpha_from_inverse_batched(const double* hyp, int64_t hyp_stride, const double* Winv,
                                   int32_t n, int32_t S, const double* y, double* alpha,
                                   double* work, void* stream);
/* llh[s] = -sum_i log L_ii - 0.5 (y-mean)^T alpha   (log_likelihood, :772-798; a12). */
int bqo_loglik_batched(const double* hyp, int64_t hyp_stride, const double* L, const double* alpha,
                       const double* y, int32_t n, int32_t S, double* llh, void* stream);
/* Workspace (bytes) needed by bqo_loglik_grad_batched. */
int64_t bqo_loglik_grad_workspace_bytes(int32_t n, int32_t S);
/* grad[s][0]=d/d var_noise, [1]=d/d mean, [2+j]=d/d kernel param j
 * (grad_log_likelihood, sbo/models/gp_fitting_gaussian.py:832-897, 1654-1697; a13).
 * Computed as 0.5 <alpha alpha^T - Sigma^{-1}, dK_j>_F with Sigma^{-1} formed once from L and
 * dK_j generated on the fly (never written to HBM).  Winv: L^-1 from bqo_chol_cov_batched, or NULL
 * (it is then built here from L). */
int bqo_loglik_grad_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                            const double* Xs, const int32_t* tid, int32_t n, const double* L,
                            const double* alpha, const double* y, const double* Winv, double* grad,
                            void* workspace, void* stream);

/* ---- GP posterior and EI (a18, a19, a39) ------------------------------------------------ */
/* For each hyper-sample s and query point q (P raw [m][dim]):
 * mean[s][q] = mean_s + k* alpha_s; var[s][q] = k(q,q) - k* Sigma^{-1} k*^T;
 * gmean[s][q][l], gvar[s][q][l] (l<dim_m) their gradients (nullable)
 * (compute_posterior_parameters / gradient_posterior_parameters,
 * sbo/models/gp_fitting_gaussian.py:1253-1353).  work: S*m*(1+dim_m)*n doubles. */
int bqo_posterior_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                          const double* Xs, const int32_t* tid, int32_t n, const double* L,
                          const double* alpha, const double* P, int32_t m, double* mean,
                          double* var, double* gmean, double* gvar, double* work, void* stream);
/* EI value and gradient from posterior moments (EI.evaluate / evaluate_gradient,
 * sbo/acquisition_functions/ei.py:81-112,135-176): count = S*m entries, best[s] per
 * hyper-sample; gei nullable. */
int bqo_ei_batched(const double* mean, const double* var, const double* gmean, const double* gvar,
                   const double* best, int32_t S, int32_t m, int32_t dim_g, double* ei,
                   double* gei, void* stream);

/* ---- Bayesian quadrature / SBO stage B: per (candidate, hyper-sample) unit setup (a28) --- */
/* Quadrature weights of the finite-W expectations (uniform_finite with optional weights,
 * sbo/lib/expectations.py:10-74): qt[s][t'] = sum_t w_t C_s[t][t'] / sum_t w_t.
 * `weights` NULL = uniform.  Only for PRODUCT. */
int bqo_task_quadrature_weights(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                                const double* weights, double* qt, void* stream);
/* After bqo_cross_cov_batched(interleave=1) + bqo_potrs_batched on R[s][c*(1+dim_m)+r][n]:
 * den[s][c] = sqrt(clip(k(c,c) - gamma^T s2 + noise, 0)), beta4[s][c][l] = 2 grad_gamma_l^T s2
 * (get_parameters_for_samples, sbo/numerical_tools/bayesian_quadrature.py:1111-1127; the
 * beta_4 of gradient_vector_b, :1510-1511).  G = the un-solved rows (gamma, grad gamma),
 * R = the solved rows (s2, Sigma^{-1} grad gamma).  add_noise[s] = noise term of :1122-1125.
 * For PRODUCT the solved rows are then multiplied in place by q_s[task_j] (qt, tid required) so
 * that stage C sees B(x, X_j) = k_M(x, X_j.x) q_j as plain weights. */
int bqo_candidate_setup_batched(const bqo_kernel_desc* desc, const double* hyp, int32_t S,
                                const double* G, double* R, int32_t n, int32_t C,
                                const double* cand, const double* add_noise, const double* qt,
                                const int32_t* tid, double* den, double* beta4, void* stream);
/* alpha_q[s][j] = q_s[task_j] alpha_s[j] for PRODUCT (the task sum of uniform_finite folded into
 * the weights), a plain copy for the other kinds. */
int bqo_weight_alpha(const bqo_kernel_desc* desc, const double* alpha, const double* qt,
                     const int32_t* tid, int32_t S, int32_t n, double* out, void* stream);

/* ---- SBO stage C: a(x), b(x), inner maximisation, gradient (a21,a23,a25,a29-a34) -------- */
/* Problem description shared by the stage-C entry points. */
typedef struct bqo_stage_c {
    bqo_kernel_desc desc;
    int32_t n;          /* training points */
    int32_t S;          /* hyper-samples resident */
    int32_t dx;         /* leading x columns (optimised over) */
    int32_t dw;         /* trailing continuous w columns integrated by Monte Carlo (0 for PRODUCT) */
    int32_t n_units;    /* (candidate, hyper-sample) units in this batch */
    int32_t C;          /* candidates in R/den/beta4 (unit u -> candidate unit_c[u]) */
    int32_t M;          /* W samples per quadrature evaluation (N_SAMPLES=10, expectations.py:7) */
    int32_t n_wsets;    /* pre-drawn W sample sets; the inner maximiser and (by default) the
                           candidate gradient use set (candidate index % n_wsets) for every
                           evaluation of a unit, so each run optimises a fixed smooth function */
    const double* hyp;      /* [S][stride] */
    const double* Xs;       /* [S][dim_m][n] */
    const int32_t* tid;     /* [n] or NULL */
    const double* alpha;    /* [S][n] (q-weighted by bqo_weight_alpha for PRODUCT) */
    const double* qt;       /* [S][T] quadrature weights (PRODUCT) or NULL */
    const double* R;        /* [S][C*(1+dim_m)][n] solved rows: s2 and Sigma^{-1} grad gamma */
    const double* den;      /* [S][C] */
    const double* beta4;    /* [S][C][dim_m] */
    const double* cand;     /* [C][dim] raw candidate points */
    const int32_t* unit_k;  /* [n_units] hyper-sample of the unit */
    const int32_t* unit_c;  /* [n_units] candidate of the unit */
    const double* wsets;    /* [n_wsets][M][dw] raw W samples (gamma.rvs of expectations.py:105) */
    const double* lower;    /* [dx] box bounds of x */
    const double* upper;    /* [dx] */
    const double* wdist;    /* [S][n_wsets][ceil(n/32)][M][32] from bqo_wdist_precompute, or NULL
                               (the W part of the squared distances is then recomputed in every
                               evaluation) */
} bqo_stage_c;

/* W part of the squared scaled distances between the W samples of every set and every training
 * point, per hyper-sample (it does not depend on the query point x): the term that gamma_expect
 * (sbo/lib/expectations.py:76-118) re-evaluates inside every kernel call through
 * Distances.dist_square_length_scale (sbo/lib/distances.py:10-29).
 * out: S * n_wsets * ceil(n/32) * M * 32 doubles. */
int bqo_wdist_precompute(const bqo_kernel_desc* desc, const double* hyp, int32_t S, const double* Xs,
                         int32_t n, int32_t dx, int32_t dw, const double* wsets, int32_t n_wsets,
                         int32_t M, double* out, void* stream);

/* a, b, grad a, grad b at P points: point q belongs to unit pt_unit[q], uses W set pt_wset[q]
 * (compute_parameters_for_sample / compute_gradient_parameters_for_sample,
 * sbo/numerical_tools/bayesian_quadrature.py:1148-1239).  out[q] = {a, b, ga[dx], gb[dx]}. */
int bqo_sample_ab_batched(const bqo_stage_c* pb, const double* pts, const int32_t* pt_unit,
                          const int32_t* pt_wset, int32_t P, double* out, void* stream);

/* Hessians of a and b with respect to x at P points (compute_hessian_parameters_for_sample,
 * sbo/numerical_tools/bayesian_quadrature.py:1241-1286, through evaluate_hessian_cross_cov :364-386
 * and hessian_uniform_finite / hessian_gamma, sbo/lib/expectations.py:260-324; only the Newton /
 * dogleg inner optimisers use them).  out[q] = {hess a [dx][dx], hess b [dx][dx]}; hess b = 0 when
 * den = 0. */
int bqo_sample_ab_hess_batched(const bqo_stage_c* pb, const double* pts, const int32_t* pt_unit,
                               const int32_t* pt_wset, int32_t P, double* out, void* stream);

typedef struct bqo_inner_opt {
    int32_t n_z;        /* Z samples (DEFAULT_N_SAMPLES = 100) */
    int32_t n_starts;   /* n_restarts + 1 start points per run set */
    int32_t maxiter;    /* maxiter_mc (10) */
    int32_t max_ls;     /* line-search trials per iteration (20, as in L-BFGS-B) */
    double factr;       /* factr_mc (1e12): stop when rel. decrease <= factr * DBL_EPSILON */
    double pgtol;       /* projected-gradient tolerance (1e-5, scipy default) */
    int32_t use_vertices; /* compare against the 2^dx box vertices (sbo.py:264-269,346-364) */
    int32_t z_per_cta;  /* Z samples handled per CTA (scheduling granularity) */
    int32_t starts_per_hyper; /* 1: starts is [S][n_starts][dx] (per hyper-sample), 0: shared */
} bqo_inner_opt;

/* For every unit u and Z sample i: max_x a(x) + Z_i b(x) from `n_starts` starts plus the box
 * vertices (SBO.evaluate_sbo_by_sample, sbo/acquisition_functions/sbo.py:237-366).
 * z[n_z]; starts[n_starts][dx]; out_max[u][i]; out_arg[u][i][dx]; n_evals[u] (nullable) counts
 * f/grad evaluations actually performed. */
int bqo_inner_maximize_batched(const bqo_stage_c* pb, const bqo_inner_opt* opt, const double* z,
                               const double* starts, double* out_max, double* out_arg,
                               unsigned long long* n_evals, void* stream);
/* grad_c b(x_p; c) for npts points per unit (BayesianQuadrature.gradient_vector_b,
 * sbo/numerical_tools/bayesian_quadrature.py:1447-1519).  pts[u][npts][dx];
 * out[u][npts][dim_m]; is_nan[u] = 1 where den^2 == 0 (the np.nan sentinel, :1483-1484). */
int bqo_grad_vector_b_batched(const bqo_stage_c* pb, const double* pts, int32_t npts,
                              const int32_t* pt_wset, double* out, int32_t* is_nan, void* stream);
/* value[c] = mean_k( mean_i max[u(c,k)][i] - max_mean[k] ),
 * grad[c][l] = mean_k mean_i z_i gvb[u][i][l]  (evaluate_mc_bayesian_candidate_points_no_restarts,
 * sbo/acquisition_functions/sbo.py:622-670).  Units must be ordered u = c*S + k.  NaN rows
 * propagate as in :655-666.  Fixed summation order (bit-reproducible). */
int bqo_acq_reduce(const double* out_max, const double* gvb, const int32_t* is_nan,
                   const double* z, const double* max_mean, int32_t C, int32_t S, int32_t n_z,
                   int32_t dim_g, double* value, double* grad, void* stream);

/* ---- discretised SBO / knowledge gradient (a41) ------------------------------------------ */
/* KG value from vectors a[m], b[m] per problem (AffineBreakPointsPrep/AffineBreakPoints +
 * hvoi, sbo/lib/affine_break_points.py:9-108, sbo/acquisition_functions/sbo.py:1394-1425,
 * 1952-1961).  a is shared ([m]); b[B][m]; work: bqo_kg_workspace_bytes(m, B) bytes.
 * kg[B]; nullable outputs for evaluate_gradient (sbo.py:1428-1474): keep_count[B],
 * keep_idx[B][m] = discretisation indices kept on the upper envelope (keep[keep1]) and
 * c_out[B][m] = the breakpoints c[keep1+1][0:M-1]. */
int64_t bqo_kg_workspace_bytes(int32_t m, int32_t B);
int bqo_kg_discretized_batched(const double* a, const double* b, int32_t m, int32_t B, double* kg,
                               int32_t* keep_count, int32_t* keep_idx, double* c_out, void* work,
                               void* stream);

/* ---- multi-start SGD driver (a36) --------------------------------------------------------- */
/* One Adam step for R restarts in place (SGD, sbo/lib/stochastic_gradient_descent.py:83-130):
 * point -= lr * mhat / (sqrt(vhat) + eps), box projection, and the relative-step stop rule;
 * grad is the gradient of the function being MINIMISED. active[r] is cleared when restart r
 * has converged (norm < 0.01). */
int bqo_adam_step_batched(double* point, const double* grad, double* m0, double* v0,
                          int32_t* active, const double* lower, const double* upper, int32_t R,
                          int32_t dim, int32_t t, double lr, double beta1, double beta2, double eps,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BQO_SM100_H */
