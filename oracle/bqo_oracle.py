"""CPU oracle: numpy/scipy restatement of the reference's GP + BQO/SBO/EI hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may.  It exists to check the CUDA
path, never to stand in for it.

Every function restates one reference function and cites it (paths relative to the reference
tree, `sbo/` = `stratified_bayesian_optimization/`).  The arithmetic keeps the reference's
association (same solves, same expectations over W, same order of means) so that timing it is a
fair CPU baseline; only hidden randomness is hoisted into explicit arguments (W samples, Z
samples, start points), as SURVEY.md section 7 requires for reproducible parity.

Parity pin: `tests/golden/make_golden.py` imports the untouched reference (py3 shim) in the build
container and stores its outputs for seeded inputs under `tests/golden/*.npz`;
`tests/test_oracle_golden.py` checks this module against those files and against the closed-form
known-answer values of the reference's own test-suite.

Third-party arithmetic the reference delegates to (not under /root/reference): scipy
`cdist(...,'sqeuclidean')`, LAPACK `dpotrf/dpotrs`, `fmin_l_bfgs_b`, `scipy.stats.norm/gamma`
(pinned by the reference at numpy==1.12.1, scipy==0.19.0, requirements.txt:12,17).
"""
from __future__ import annotations

import itertools

import numpy as np
from scipy import linalg
from scipy.linalg import lapack
from scipy.optimize import fmin_l_bfgs_b
from scipy.stats import norm

MATERN52 = 0
SCALED_MATERN52 = 1
PRODUCT_TASKS = 2

N_SAMPLES = 10  # sbo/lib/expectations.py:7


# --------------------------------------------------------------------------------------------
# distances / Matern 5/2
def dist_square_length_scale(ls, x1, x2=None):
    """sbo/lib/distances.py:10-29 (cdist 'sqeuclidean' on x/ls, direct difference form)."""
    x1 = np.asarray(x1, dtype=float) / ls
    x2 = x1 if x2 is None else np.asarray(x2, dtype=float) / ls
    d = x1[:, None, :] - x2[None, :, :]
    return np.einsum("ijk,ijk->ij", d, d)


def matern52_cross_cov(ls, x1, x2):
    """sbo/kernels/matern52.py:173-183."""
    r2 = np.abs(dist_square_length_scale(ls, x1, x2))
    r = np.sqrt(r2)
    return (1.0 + np.sqrt(5) * r + (5.0 / 3.0) * r2) * np.exp(-np.sqrt(5) * r)


def matern52_dr(ls, x1, x2):
    """k'(r): GradientLSMatern52.gradient_respect_distance_cross, sbo/kernels/matern52.py:407-440."""
    r2 = np.abs(dist_square_length_scale(ls, x1, x2))
    r = np.sqrt(r2)
    exp_r = np.exp(-np.sqrt(5) * r)
    part_1 = (1.0 + np.sqrt(5) * r + (5.0 / 3.0) * r2) * exp_r * (-np.sqrt(5))
    part_2 = exp_r * (np.sqrt(5) + (10.0 / 3.0) * r)
    return part_1 + part_2


def matern52_grad_ls(ls, x):
    """dK/d ls_l: sbo/kernels/matern52.py:374-394 with sbo/lib/distances.py:32-67.

    Returns a list of n x n matrices.  Off-diagonal duplicate points give 0/0 in the reference;
    that single case is mapped to 0 here (edge cases).
    """
    x = np.asarray(x, dtype=float)
    dr = matern52_dr(ls, x, x)
    r = np.sqrt(np.abs(dist_square_length_scale(ls, x, x)))
    out = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for l in range(len(ls)):
            x_dist = (x[:, l : l + 1] - x[:, l : l + 1].T) ** 2
            product_1 = (1.0 / r) * x_dist
            derivative = product_1 * (-1.0 / (ls[l] ** 3))
            derivative[~np.isfinite(derivative)] = 0.0
            np.fill_diagonal(derivative, 0.0)
            out.append(derivative * dr)
    return out


def matern52_grad_respect_point(ls, point, inputs):
    """grad_x k(point, inputs): sbo/kernels/matern52.py:443-462, sbo/lib/distances.py:70-94."""
    point = np.asarray(point, dtype=float).reshape(1, -1)
    inputs = np.asarray(inputs, dtype=float)
    dr = matern52_dr(ls, point, inputs)  # 1 x n
    r = np.sqrt(np.abs(dist_square_length_scale(ls, point, inputs)))
    grad = np.zeros((inputs.shape[0], len(ls)))
    with np.errstate(divide="ignore", invalid="ignore"):
        for l in range(len(ls)):
            differences = (point[0:1, l] - inputs[:, l]) / (ls[l] ** 2)
            grad[:, l] = (differences / r)[0, :]
        grad = grad * dr.T
    return np.nan_to_num(grad)


# --------------------------------------------------------------------------------------------
# tasks kernel
def matern52_hessian_respect_point(ls, point, inputs):
    """Hessians of k(point, inputs_i) with respect to point -> n x d x d:
    GradientLSMatern52.hessian_respect_point (sbo/kernels/matern52.py:466-503) =
    hess(r) k'(r) + grad r grad r^T k''(r) with r = dist(point, x)
    (Distances.gradient_distance_length_scale_respect_point(second=True), sbo/lib/distances.py:70-116).
    Written in the equivalent form without the 1/r and 1/r^3 factors:
    d2k/dx_l dx_m = g(r) delta_lm / ls_l^2 + h(r) u_l u_m / (ls_l ls_m), u = (point - x) / ls,
    g = k'(r)/r = -5/3 (1 + sqrt5 r) e^{-sqrt5 r}, h = g'(r)/r = 25/3 e^{-sqrt5 r}."""
    ls = np.asarray(ls, dtype=float)
    point = np.asarray(point, dtype=float).reshape(1, -1)
    inputs = np.asarray(inputs, dtype=float)
    u = (point - inputs) / ls                                   # n x d
    r = np.sqrt(np.sum(u * u, axis=1))
    e = np.exp(-np.sqrt(5.0) * r)
    g = -(5.0 / 3.0) * (1.0 + np.sqrt(5.0) * r) * e
    h = (25.0 / 3.0) * e
    v = u / ls
    hess = h[:, None, None] * v[:, :, None] * v[:, None, :]
    d = len(ls)
    hess[:, np.arange(d), np.arange(d)] += g[:, None] / (ls * ls)[None, :]
    return hess


def tasks_cov_matrix(params, n_tasks, same_correlation=False):
    """TasksKernel.compute_cov_matrix, sbo/kernels/tasks_kernel.py:198-239 -> (covM, L)."""
    params = np.asarray(params, dtype=float)
    if not same_correlation:
        L = np.zeros((n_tasks, n_tasks))
        count = 0
        for i in range(n_tasks):
            for j in range(i + 1):
                L[i, j] = np.exp(params[count + j])
            count += i + 1
        return np.dot(L, L.T), L
    covM = np.zeros((n_tasks, n_tasks))
    if n_tasks > 1:
        covM.fill(np.exp(params[1]))
        for i in range(n_tasks):
            covM[i, i] = np.exp(params[0]) + np.exp(params[1]) * (n_tasks - 1)
    else:
        covM[0, 0] = np.exp(params[0])
    return covM, covM


def tasks_grad_base(chol_base, n_tasks, same_correlation=False):
    """GradientTasksKernel.gradient_respect_parameters, sbo/kernels/tasks_kernel.py:580-622."""
    gradient = {}
    if not same_correlation:
        count = 0
        for i in range(n_tasks):
            for j in range(i + 1):
                tmp = np.zeros((n_tasks, n_tasks))
                tmp[i, j] = chol_base[i, j]
                mat = np.dot(tmp, chol_base.T)
                gradient[count + j] = mat + mat.T
            count += i + 1
        return gradient
    gradient[0] = chol_base[0, 0] * np.identity(n_tasks)
    if n_tasks == 1:
        return gradient
    gradient[1] = chol_base.copy()
    value = chol_base[0, 1]
    gradient[0] = (chol_base[0, 0] - value * (n_tasks - 1)) * np.identity(n_tasks)
    for i in range(n_tasks):
        gradient[1][i, i] = value * (n_tasks - 1)
    return gradient


# --------------------------------------------------------------------------------------------
class KernelSpec(object):
    """Which covariance function: mirrors (type_kernel, dimensions) of GPFittingGaussian."""

    def __init__(self, kind, dim, n_tasks=0, same_correlation=False):
        self.kind = kind
        self.dim = dim
        self.n_tasks = n_tasks
        self.same_correlation = same_correlation
        self.dim_m = dim - 1 if kind == PRODUCT_TASKS else dim
        if kind == MATERN52:
            self.n_params = dim
        elif kind == SCALED_MATERN52:
            self.n_params = dim + 1
        else:
            nt = min(n_tasks, 2) if same_correlation else n_tasks * (n_tasks + 1) // 2
            self.n_params = self.dim_m + nt

    # ---- kernel evaluations (the evaluate_*_defined_by_params classmethods) ----
    def cross_cov(self, params, x1, x2):
        """Matern52 / ScaledKernel / ProductKernels cross_cov:
        sbo/kernels/matern52.py:173-183, scaled_kernel.py:186-194, product_kernels.py:260-270,
        tasks_kernel.py:241-257."""
        params = np.asarray(params, dtype=float)
        x1 = np.asarray(x1, dtype=float)
        x2 = np.asarray(x2, dtype=float)
        dm = self.dim_m
        km = matern52_cross_cov(params[:dm], x1[:, :dm], x2[:, :dm])
        if self.kind == MATERN52:
            return km
        if self.kind == SCALED_MATERN52:
            return km * params[dm]
        covM, _ = tasks_cov_matrix(params[dm:], self.n_tasks, self.same_correlation)
        s = x1[:, dm].astype(int)
        t = x2[:, dm].astype(int)
        return km * covM[np.ix_(s, t)]

    def cov(self, params, x):
        return self.cross_cov(params, x, x)

    def grad_params(self, params, x):
        """evaluate_grad_defined_by_params_respect_params -> list of n x n matrices in parameter
        order (sbo/kernels/scaled_kernel.py:196-214, product_kernels.py:272-302,
        tasks_kernel.py:259-287)."""
        params = np.asarray(params, dtype=float)
        x = np.asarray(x, dtype=float)
        dm = self.dim_m
        gls = matern52_grad_ls(params[:dm], x[:, :dm])
        if self.kind == MATERN52:
            return gls
        if self.kind == SCALED_MATERN52:
            sigma2 = params[dm]
            out = [g * sigma2 for g in gls]
            out.append(self.cov(params, x) / sigma2)
            return out
        covM, L = tasks_cov_matrix(params[dm:], self.n_tasks, self.same_correlation)
        tid = x[:, dm].astype(int)
        kt = covM[np.ix_(tid, tid)]
        km = matern52_cross_cov(params[:dm], x[:, :dm], x[:, :dm])
        out = [g * kt for g in gls]
        gb = tasks_grad_base(L, self.n_tasks, self.same_correlation)
        for m in range(len(gb)):
            out.append(km * gb[m][np.ix_(tid, tid)])
        return out

    def grad_respect_point(self, params, point, inputs):
        """evaluate_grad_respect_point -> n x dim (sbo/kernels/scaled_kernel.py:216-229,
        product_kernels.py:304-325,415-439; the task column has zero gradient)."""
        params = np.asarray(params, dtype=float)
        point = np.asarray(point, dtype=float).reshape(1, -1)
        inputs = np.asarray(inputs, dtype=float)
        dm = self.dim_m
        g = matern52_grad_respect_point(params[:dm], point[:, :dm], inputs[:, :dm])
        if self.kind == MATERN52:
            return g
        if self.kind == SCALED_MATERN52:
            return g * params[dm]
        covM, _ = tasks_cov_matrix(params[dm:], self.n_tasks, self.same_correlation)
        kt = covM[np.ix_(point[:, dm].astype(int), inputs[:, dm].astype(int))].T  # n x 1
        return np.concatenate([g * kt, np.zeros((inputs.shape[0], 1))], axis=1)


    def hessian_respect_point(self, params, point, inputs):
        """evaluate_hessian_respect_point -> n x dim x dim (sbo/kernels/scaled_kernel.py:229-270,
        product_kernels.py:327-481: the Matern block times the other factor; rows and columns of
        the task coordinate are zero because its gradient and Hessian vanish)."""
        params = np.asarray(params, dtype=float)
        point = np.asarray(point, dtype=float).reshape(1, -1)
        inputs = np.asarray(inputs, dtype=float)
        dm = self.dim_m
        hm = matern52_hessian_respect_point(params[:dm], point[:, :dm], inputs[:, :dm])
        if self.kind == MATERN52:
            return hm
        if self.kind == SCALED_MATERN52:
            return hm * params[dm]
        covM, _ = tasks_cov_matrix(params[dm:], self.n_tasks, self.same_correlation)
        kt = covM[point[0, dm].astype(int), inputs[:, dm].astype(int)]
        out = np.zeros((inputs.shape[0], self.dim, self.dim))
        out[:, :dm, :dm] = hm * kt[:, None, None]
        return out


# --------------------------------------------------------------------------------------------
# linear algebra
def cholesky(cov, max_tries=5):
    """sbo/lib/la_functions.py:8-38 (jitter escalation included)."""
    cov = np.ascontiguousarray(cov)
    L, info = lapack.dpotrf(cov, lower=1)
    if info == 0:
        return L
    diag_cov = np.diag(cov)
    if np.any(diag_cov <= 0.0):
        raise linalg.LinAlgError("not positive definite matrix")
    jitter = diag_cov.mean() * 1e-6
    n_tries = 1
    while n_tries <= max_tries and np.isfinite(jitter):
        try:
            return linalg.cholesky(cov + np.eye(cov.shape[0]) * jitter, lower=True)
        except Exception:
            jitter *= 10
        finally:
            n_tries += 1
    raise linalg.LinAlgError("not positive definite, even with jitter.")


def cho_solve(chol, y):
    """sbo/lib/la_functions.py:41-50."""
    chol = np.asfortranarray(chol)
    return lapack.dpotrs(chol, y, lower=1)[0]


# --------------------------------------------------------------------------------------------
class OracleGP(object):
    """GPFittingGaussian restricted to the numerical hot path
    (sbo/models/gp_fitting_gaussian.py).  theta = [var_noise, mean, kernel params...]."""

    def __init__(self, spec, points, evaluations, var_noise_obs=None):
        self.spec = spec
        self.points = np.asarray(points, dtype=float)
        self.evaluations = np.asarray(evaluations, dtype=float)
        self.var_noise_obs = None if var_noise_obs is None else np.asarray(var_noise_obs, dtype=float)
        self._cache = {}

    def chol_cov_including_noise(self, var_noise, params):
        """:733-770."""
        key = (float(var_noise), tuple(np.asarray(params, dtype=float)))
        if key in self._cache:
            return self._cache[key]
        n = self.points.shape[0]
        cov = self.spec.cov(params, self.points)
        if self.var_noise_obs is not None:
            cov = cov + np.diag(self.var_noise_obs)
        cov = cov + np.diag(var_noise * np.ones(n))
        chol = cholesky(cov, max_tries=7)
        self._cache = {key: (chol, cov)}
        return chol, cov

    def clean_cache(self):
        self._cache = {}

    def log_likelihood(self, var_noise, mean, params):
        """:772-798 (no -n/2 log 2 pi term)."""
        chol, _ = self.chol_cov_including_noise(var_noise, params)
        y_unbiased = self.evaluations - mean
        solve = cho_solve(chol, y_unbiased)
        return -np.sum(np.log(np.diag(chol))) - 0.5 * np.dot(y_unbiased, solve)

    def grad_log_likelihood(self, var_noise, mean, params):
        """:832-897 with GradientGPFittingGaussian :1654-1697: one n x n dpotrs and one
        (alpha alpha^T) dK product per parameter, as the reference does."""
        grad_cov = self.spec.grad_params(params, self.points)
        chol, _ = self.chol_cov_including_noise(var_noise, params)
        y_unbiased = self.evaluations - mean
        solve = cho_solve(chol, y_unbiased)
        n = self.points.shape[0]

        def given_grad_cov(gc):
            s = solve.reshape((n, 1))
            solve_1 = cho_solve(chol, gc)
            product = np.dot(np.dot(s, s.T), gc)
            return 0.5 * np.trace(product - solve_1)

        g = np.zeros(2 + len(params))
        g[0] = given_grad_cov(np.identity(n))
        g[1] = -1.0 * np.dot(y_unbiased, cho_solve(chol, -1.0 * np.ones(n)))
        for i in range(len(params)):
            g[2 + i] = given_grad_cov(grad_cov[i])
        return g

    def chol_solve(self, var_noise, mean, params):
        """_cholesky_solve_vectors_for_posterior, :1202-1251."""
        chol, _ = self.chol_cov_including_noise(var_noise, params)
        solve = cho_solve(chol, self.evaluations - mean)
        return chol, solve

    def compute_posterior_parameters(self, points, var_noise, mean, params, only_mean=False):
        """:1253-1311."""
        chol, solve = self.chol_solve(var_noise, mean, params)
        vec_cov = self.spec.cross_cov(params, points, self.points)
        mu_n = mean + np.dot(vec_cov, solve)
        if only_mean:
            return {"mean": mu_n, "cov": None}
        solve_2 = cho_solve(chol, vec_cov.T)
        cov_n = self.spec.cov(params, points) - np.dot(vec_cov, solve_2)
        return {"mean": mu_n, "cov": cov_n}

    def gradient_posterior_parameters(self, point, var_noise, mean, params):
        """:1313-1353."""
        chol, solve = self.chol_solve(var_noise, mean, params)
        grad_cross_cov = self.spec.grad_respect_point(params, point, self.points)
        grad_mu = np.dot(grad_cross_cov.T, solve)
        vec_cov = self.spec.cross_cov(params, point, self.points)
        solve_2 = cho_solve(chol, grad_cross_cov)
        grad_cov = -2.0 * np.dot(vec_cov, solve_2)
        return {"mean": grad_mu, "cov": grad_cov}

    def get_historical_best_solution(self, var_noise, mean, params, noisy_evaluations=False):
        """:1590-1626."""
        if not noisy_evaluations:
            return np.max(self.evaluations)
        post = self.compute_posterior_parameters(self.points, var_noise, mean, params, only_mean=True)
        return np.max(post["mean"])


# --------------------------------------------------------------------------------------------
# EI on a GP (sbo/acquisition_functions/ei.py:81-112,135-176)
def ei_evaluate(gp, point, var_noise, mean, params, noisy_evaluations=False):
    post = gp.compute_posterior_parameters(point, var_noise, mean, params)
    mu = post["mean"]
    cov = np.clip(post["cov"], 0, None)
    best = gp.get_historical_best_solution(var_noise, mean, params, noisy_evaluations)
    with np.errstate(divide="ignore", invalid="ignore"):
        normalized_factor = (mu - best) / np.sqrt(cov)
        first_term = (mu - best) * norm.cdf(normalized_factor)
        second_term = np.sqrt(cov) * norm.pdf(normalized_factor)
    evaluation = first_term + second_term
    if len(evaluation.shape) == 2:
        evaluation = evaluation[0, :]
    return evaluation


def ei_evaluate_gradient(gp, point, var_noise, mean, params, noisy_evaluations=False):
    post = gp.compute_posterior_parameters(point, var_noise, mean, params)
    mu = post["mean"]
    cov = np.clip(post["cov"], 0, None)
    best = gp.get_historical_best_solution(var_noise, mean, params, noisy_evaluations)
    std = np.sqrt(cov)
    gradient = gp.gradient_posterior_parameters(point, var_noise, mean, params)
    grad_mu = gradient["mean"]
    grad_cov = gradient["cov"]
    grad_std = 0.5 * grad_cov / np.sqrt(cov)
    grad_factor = (grad_mu * std - grad_std * (mu - best)) / cov
    normalized_factor = (mu - best) / std
    first_term = grad_mu * norm.cdf(normalized_factor) + (mu - best) * grad_factor * norm.pdf(
        normalized_factor
    )
    second_term = grad_std * norm.pdf(normalized_factor) - std * norm.pdf(
        normalized_factor
    ) * grad_factor * normalized_factor
    return (first_term + second_term)[0, :]


# --------------------------------------------------------------------------------------------
# Expectations over W (sbo/lib/expectations.py)
UNIFORM_FINITE = "uniform_finite"
WEIGHTED_UNIFORM_FINITE = "weighted_uniform_finite"
GAMMA = "gamma"


class OracleBQ(object):
    """BayesianQuadrature (sbo/numerical_tools/bayesian_quadrature.py) with explicit W samples.

    For the gamma distribution every reference call draws fresh `gamma.rvs`; here the caller
    passes the samples (`w`, shape [M, d_w]) so both implementations see the same numbers.
    """

    def __init__(self, gp, x_domain, distribution, weights=None, domain_random=None):
        self.gp = gp
        self.spec = gp.spec
        self.x_domain = list(x_domain)
        self.w_domain = [i for i in range(self.spec.dim) if i not in self.x_domain]
        self.distribution = distribution
        if distribution == UNIFORM_FINITE:
            n_tasks = self.spec.n_tasks
            self.domain_random = np.arange(n_tasks).reshape((n_tasks, 1))
            self.weights = None
        elif distribution == WEIGHTED_UNIFORM_FINITE:
            self.domain_random = np.asarray(domain_random, dtype=float)
            self.weights = np.asarray(weights, dtype=float)
        else:
            self.domain_random = None
            self.weights = None
        self.var_noise = None
        if gp.var_noise_obs is not None:
            self.var_noise = np.mean(gp.var_noise_obs)  # bayesian_quadrature.py:153-155

    # new_points = (point, W-values) -- expectations.py:44-50 / :100-106
    def _new_points(self, point, w):
        point = np.asarray(point, dtype=float).reshape(1, -1)
        if self.distribution == GAMMA:
            random = np.asarray(w, dtype=float)
        else:
            random = self.domain_random
        m = random.shape[0]
        new_points = np.zeros((m, len(self.w_domain) + point.shape[1]))
        new_points[:, self.x_domain] = np.repeat(point, m, axis=0)
        new_points[:, self.w_domain] = random
        return new_points

    def evaluate_quadrature_cross_cov(self, point, points_2, params, w=None):
        """B(x, z_j): bayesian_quadrature.py:308-334 with uniform_finite (expectations.py:10-74)
        or gamma_expect (:76-118)."""
        values = self.spec.cross_cov(params, self._new_points(point, w), points_2)
        if self.distribution == GAMMA:
            return np.mean(values, axis=0)
        return np.average(values, axis=0, weights=self.weights)

    def evaluate_grad_quadrature_cross_cov(self, point, points_2, params, w=None):
        """grad_x B(x, z_j) -> (d_x, m): bayesian_quadrature.py:336-362, expectations.py:162-200,
        237-258."""
        new_points = self._new_points(point, w)
        gradients = np.zeros((new_points.shape[0], len(self.x_domain), points_2.shape[0]))
        for i in range(new_points.shape[0]):
            value = self.spec.grad_respect_point(params, new_points[i : i + 1, :], points_2)
            gradients[i, :, :] = value[:, self.x_domain].T
        if self.distribution == GAMMA:
            return np.mean(gradients, axis=0)
        return np.average(gradients, axis=0, weights=self.weights)

    def evaluate_grad_quadrature_cross_cov_resp_candidate(self, candidate_point, points, params,
                                                          w=None):
        """grad_c B(x_p, c) for each x_p -> (dim, m): bayesian_quadrature.py:388-415,
        expectations.py:327-385.  `w`: [m][M][d_w] (one set per point) or [M][d_w] (shared)."""
        candidate_point = np.asarray(candidate_point, dtype=float).reshape(1, -1)
        points = np.asarray(points, dtype=float)
        gradients = np.zeros((candidate_point.shape[1], points.shape[0]))
        for i in range(points.shape[0]):
            wi = None
            if w is not None:
                w_arr = np.asarray(w, dtype=float)
                wi = w_arr[i] if w_arr.ndim == 3 else w_arr
            new_points = self._new_points(points[i : i + 1, :], wi)
            values = self.spec.grad_respect_point(params, candidate_point, new_points)
            if self.distribution == GAMMA:
                gradients[:, i] = np.mean(values, axis=0)
            else:
                gradients[:, i] = np.average(values, axis=0, weights=self.weights)
        return gradients

    def evaluate_quadrate_cov(self, point, params, w=None, w2=None):
        """E_W E_W' k: bayesian_quadrature.py:282-306 (double=True paths of expectations.py)."""
        new_points = self._new_points(point, w)
        if self.distribution == GAMMA:
            new_points_2 = self._new_points(point, w2)
            return np.mean(self.spec.cross_cov(params, new_points, new_points_2))
        values = self.spec.cross_cov(params, new_points, new_points)
        weights = self.weights
        if weights is not None:
            weights = np.outer(weights, weights)
        return np.average(values, weights=weights)

    def compute_posterior_parameters(self, points, var_noise, mean, params, only_mean=False,
                                     w=None, w_cov=None):
        """Posterior of G(x) = E_W F(x, W): bayesian_quadrature.py:449-537."""
        chol, solve = self.gp.chol_solve(var_noise, mean, params)
        points = np.asarray(points, dtype=float)
        vec_covs = np.zeros((points.shape[0], self.gp.points.shape[0]))
        for i in range(points.shape[0]):
            vec_covs[i, :] = self.evaluate_quadrature_cross_cov(
                points[i : i + 1, :], self.gp.points, params, w=w)
        mu_n = mean + np.dot(vec_covs, solve)
        if only_mean:
            return {"mean": mu_n, "cov": None}
        solve_2 = cho_solve(chol, vec_covs.T)
        w1, w2 = (w_cov if w_cov is not None else (w, w))
        cov_n = self.evaluate_quadrate_cov(points, params, w=w1, w2=w2) - np.dot(vec_covs, solve_2)
        return {"mean": mu_n, "cov": cov_n[0, 0]}

    def gradient_posterior_mean(self, point, var_noise, mean, params, w=None):
        """:539-579."""
        gradient = self.evaluate_grad_quadrature_cross_cov(point, self.gp.points, params, w=w)
        _, solve = self.gp.chol_solve(var_noise, mean, params)
        return np.dot(gradient, solve)

    def gradient_posterior_parameters(self, point, var_noise, mean, params, w=None):
        """{'mean': grad mu_G, 'cov': grad var_G}: bayesian_quadrature.py:581-643 (the prior term
        E E k is taken as constant in x, as the reference does)."""
        point = np.asarray(point, dtype=float).reshape(1, -1)
        gradient = self.evaluate_grad_quadrature_cross_cov(point, self.gp.points, params, w=w)
        chol, solve = self.gp.chol_solve(var_noise, mean, params)
        vec_covs = self.evaluate_quadrature_cross_cov(point, self.gp.points, params, w=w)
        grad_mu = np.dot(gradient, solve)
        solve_3 = cho_solve(chol, gradient.transpose())
        grad_cov = -2.0 * np.dot(vec_covs.reshape(1, -1), solve_3)
        return {"mean": grad_mu, "cov": grad_cov.reshape(-1)}

    def get_parameters_for_samples(self, candidate_point, var_noise, mean, params):
        """:1074-1145."""
        chol, solve = self.gp.chol_solve(var_noise, mean, params)
        candidate_point = np.asarray(candidate_point, dtype=float).reshape(1, -1)
        cross_cov = self.spec.cross_cov(params, self.gp.points, candidate_point)
        solve_2 = cho_solve(chol, cross_cov)
        new_cross_cov = np.diag(self.spec.cross_cov(params, candidate_point, candidate_point))
        add_noise = self.var_noise if self.var_noise is not None else var_noise
        denominator = new_cross_cov - np.einsum("ij,ij->j", cross_cov, solve_2) + add_noise
        denominator = np.sqrt(np.clip(denominator, 0, None))
        return {"gamma": cross_cov, "solve_2": solve_2, "denominator": denominator,
                "chol": chol, "solve": solve}

    def compute_parameters_for_sample(self, point, candidate_point, var_noise, mean, params,
                                      w=None, w_new=None):
        """a(x), b(x): :1148-1191.  `w` feeds B(x, X), `w_new` feeds B(x, c) (default: same)."""
        ap = self.get_parameters_for_samples(candidate_point, var_noise, mean, params)
        candidate_point = np.asarray(candidate_point, dtype=float).reshape(1, -1)
        point = np.asarray(point, dtype=float).reshape(1, -1)
        vec_covs = self.evaluate_quadrature_cross_cov(point, self.gp.points, params, w=w)
        vec_covs = vec_covs.reshape(1, -1)
        b_new = self.evaluate_quadrature_cross_cov(
            point, candidate_point, params, w=w if w_new is None else w_new).reshape(1, 1)
        mu_n = mean + np.dot(vec_covs, ap["solve"])
        numerator = b_new - np.dot(vec_covs, ap["solve_2"])
        if ap["denominator"] != 0:
            b_value = numerator / ap["denominator"][None, :]
        else:
            b_value = np.array([[0]])
        return {"a": mu_n, "b": b_value}

    def compute_gradient_parameters_for_sample(self, point, candidate_point, var_noise, mean,
                                               params, w=None, w_new=None):
        """grad a, grad b: :1193-1239."""
        ap = self.get_parameters_for_samples(candidate_point, var_noise, mean, params)
        candidate_point = np.asarray(candidate_point, dtype=float).reshape(1, -1)
        point = np.asarray(point, dtype=float).reshape(1, -1)
        gradient = self.evaluate_grad_quadrature_cross_cov(point, self.gp.points, params, w=w)
        gradient_a = np.dot(gradient, ap["solve"])
        if ap["denominator"] != 0:
            gradient_new = self.evaluate_grad_quadrature_cross_cov(
                point, candidate_point, params, w=w if w_new is None else w_new)
            gradient_b = (gradient_new - np.dot(gradient, ap["solve_2"])) / ap["denominator"]
            gradient_b = gradient_b[:, 0]
        else:
            gradient_b = np.zeros(len(gradient_a))
        return {"a": gradient_a, "b": gradient_b}

    def evaluate_hessian_cross_cov(self, point, points_2, params, w=None):
        """hess_x B(x, z_j) -> (m, d_x, d_x): bayesian_quadrature.py:364-386 with
        hessian_uniform_finite / hessian_gamma (expectations.py:260-324)."""
        new_points = self._new_points(point, w)
        points_2 = np.asarray(points_2, dtype=float)
        dx = len(self.x_domain)
        hess = np.zeros((new_points.shape[0], points_2.shape[0], dx, dx))
        for i in range(new_points.shape[0]):
            value = self.spec.hessian_respect_point(params, new_points[i : i + 1, :], points_2)
            hess[i] = value[:, self.x_domain][:, :, self.x_domain]
        if self.distribution == GAMMA:
            return np.mean(hess, axis=0)
        return np.average(hess, axis=0, weights=self.weights)

    def compute_hessian_parameters_for_sample(self, point, candidate_point, var_noise, mean,
                                              params, w=None):
        """hess a, hess b: :1241-1286."""
        ap = self.get_parameters_for_samples(candidate_point, var_noise, mean, params)
        candidate_point = np.asarray(candidate_point, dtype=float).reshape(1, -1)
        point = np.asarray(point, dtype=float).reshape(1, -1)
        hessian = self.evaluate_hessian_cross_cov(point, self.gp.points, params, w=w)
        hessian_a = np.einsum("ijk,i->jk", hessian, ap["solve"])
        if ap["denominator"] != 0:
            hessian_new = self.evaluate_hessian_cross_cov(point, candidate_point, params, w=w)[0]
            hessian_b = (hessian_new - np.einsum("ijk,i->jk", hessian, ap["solve_2"][:, 0])) \
                / ap["denominator"]
        else:
            hessian_b = np.zeros(hessian_a.shape)
        return {"a": hessian_a, "b": hessian_b}

    def gradient_vector_b(self, candidate_point, points, var_noise, mean, params, w=None):
        """grad_c b(x_p; c) -> (m, dim) or np.nan: :1447-1519 (solves Sigma^-1 B^T with m
        right-hand sides, as the reference does).  `w`: [m][M][d_w] or [M][d_w]."""
        ap = self.get_parameters_for_samples(candidate_point, var_noise, mean, params)
        candidate_point = np.asarray(candidate_point, dtype=float).reshape(1, -1)
        points = np.asarray(points, dtype=float)
        solve_1 = ap["solve_2"][:, 0]
        beta_1 = ap["denominator"] ** 2
        if beta_1[0] == 0:
            return np.nan
        grad_gamma = self.spec.grad_respect_point(params, candidate_point, self.gp.points)
        grad_b_new = self.evaluate_grad_quadrature_cross_cov_resp_candidate(
            candidate_point, points, params, w=w)
        m = points.shape[0]
        vec_covs = np.zeros((m, self.gp.points.shape[0]))
        b_new = np.zeros((m, 1))
        for i in range(m):
            wi = None
            if w is not None:
                w_arr = np.asarray(w, dtype=float)
                wi = w_arr[i] if w_arr.ndim == 3 else w_arr
            vec_covs[i, :] = self.evaluate_quadrature_cross_cov(
                points[i : i + 1, :], self.gp.points, params, w=wi)
            b_new[i, :] = self.evaluate_quadrature_cross_cov(
                points[i : i + 1, :], candidate_point, params, w=wi)
        solve_2 = cho_solve(ap["chol"], vec_covs.T)
        beta_1 = beta_1 ** (-0.5)
        beta_2 = b_new[:, 0] - np.dot(vec_covs, solve_1)
        beta_3 = grad_b_new - np.dot(grad_gamma.T, solve_2)
        beta_4 = 2.0 * np.dot(grad_gamma.T, solve_1)
        beta_4 = beta_4.reshape((candidate_point.shape[1], 1))
        beta_5 = 0.0
        gradients = beta_1[0] * beta_3 - 0.5 * (beta_1[0] ** 3) * beta_2 * (beta_5 - beta_4)
        return gradients.T

    def compute_posterior_parameters_kg(self, points, candidate_point, var_noise, mean, params):
        """Vectors a, b over a discretisation: :1376-1445 (finite W only: no hidden samples)."""
        chol, solve = self.gp.chol_solve(var_noise, mean, params)
        candidate_point = np.asarray(candidate_point, dtype=float).reshape(1, -1)
        points = np.asarray(points, dtype=float)
        n = points.shape[0]
        vec_covs = np.zeros((n, self.gp.points.shape[0]))
        b_new = np.zeros((n, 1))
        for i in range(n):
            vec_covs[i, :] = self.evaluate_quadrature_cross_cov(
                points[i : i + 1, :], self.gp.points, params)
            b_new[i, :] = self.evaluate_quadrature_cross_cov(
                points[i : i + 1, :], candidate_point, params)
        mu_n = mean + np.dot(vec_covs, solve)
        cross_cov = self.spec.cross_cov(params, candidate_point, self.gp.points)
        solve_2 = cho_solve(chol, cross_cov[0, :])
        numerator = b_new[:, 0] - np.dot(vec_covs, solve_2)
        new_cross_cov = self.spec.cross_cov(params, candidate_point, candidate_point)
        add_noise = self.var_noise if self.var_noise is not None else var_noise
        denominator = new_cross_cov - np.dot(cross_cov, solve_2) + add_noise
        denominator = np.clip(denominator[0, 0], 0, None)
        with np.errstate(divide="ignore", invalid="ignore"):
            b_value = numerator / np.sqrt(denominator)
        return {"a": mu_n, "b": b_value}


# --------------------------------------------------------------------------------------------
# Frazier's upper-envelope routines (sbo/lib/affine_break_points.py:9-108)
def affine_break_points_prep(a, b):
    b1, i1 = min(b), np.argmin(b)
    a1 = a[i1]
    a2, i2 = max(a), np.argmax(a)
    b2 = b[i2]
    b3, i3 = max(b), np.argmax(b)
    a3 = a[i3]
    with np.errstate(divide="ignore", invalid="ignore"):
        cleft = (a.astype(float) - a1) / (b1 - b)
        cright = (a.astype(float) - a3) / (b3 - b)
        c2left = (a2 - a1) / (b1 - b2)
        c2right = (a2 - a3) / (b3 - b2)
    keep = [i for i in range(len(b)) if
            b[i] == b1 or b[i] == b3 or cleft[i] <= c2left or cright[i] >= c2right]
    a = a[keep]
    b = b[keep]
    keep1 = np.array(keep)
    ba = np.reshape(np.concatenate((b, a)), (len(a), 2), order="F")
    ind = np.lexsort((ba[:, 1], ba[:, 0]))
    ba = ba[ind, :]
    keep1 = keep1[ind]
    a = ba[:, 1]
    b = ba[:, 0]
    tmp = np.diff(b)
    keep = [i for i in range(len(tmp)) if tmp[i] != 0]
    keep.append(len(tmp))
    return a[keep], b[keep], keep1[keep]


def affine_break_points(a, b):
    M = len(a)
    c = np.zeros(M + 1)
    A = np.zeros(M + 1)
    c[0] = -float("Inf")
    c[1] = float("Inf")
    A[1] = 1
    Alen = 1
    for i in range(1, M):
        c[i + 1] = float("Inf")
        while True:
            j = int(A[Alen])
            c[j] = (a[j - 1] - a[i]) / (b[i] - b[j - 1])
            if len(A) > 1 and c[j] <= c[int(A[Alen - 1])]:
                Alen = Alen - 1
            else:
                break
        A[Alen + 1] = i + 1
        Alen = Alen + 1
    Alen = Alen + 1
    A = A[1:Alen] - 1
    return A, c


def hvoi(b, c, keep):
    """sbo/acquisition_functions/sbo.py:1952-1961."""
    M = len(keep)
    if M > 1:
        c = c[keep + 1]
        c2 = -np.abs(c[0 : M - 1])
        tmp = norm.pdf(c2) + c2 * norm.cdf(c2)
        return np.sum(np.diff(b[keep]) * tmp)
    return 0


# --------------------------------------------------------------------------------------------
class OracleSBO(object):
    """SBO (sbo/acquisition_functions/sbo.py): sample path, inner maximisation, MC estimate."""

    def __init__(self, bq, bounds_x, discretization=None):
        self.bq = bq
        self.bounds_x = [tuple(b) for b in bounds_x]
        self.discretization = discretization

    def evaluate_sample(self, point, candidate_point, sample, var_noise, mean, params, w=None):
        """:140-165."""
        v = self.bq.compute_parameters_for_sample(point, candidate_point, var_noise, mean, params, w=w)
        return (v["a"] + sample * v["b"])[0, 0]

    def evaluate_gradient_sample(self, point, candidate_point, sample, var_noise, mean, params,
                                 w=None):
        """:167-194."""
        g = self.bq.compute_gradient_parameters_for_sample(
            point, candidate_point, var_noise, mean, params, w=w)
        return g["a"] + sample * g["b"]

    def evaluate_sbo_by_sample(self, candidate_point, sample, starts, var_noise, mean, params,
                               w=None, factr=1e12, maxiter=10):
        """max_x a(x) + Z b(x): :237-366 -- L-BFGS-B from every start (scipy, as
        Optimization.optimize, sbo/lib/optimization.py:87-155), then the box vertices."""
        results = []
        for start in np.asarray(starts, dtype=float):
            f = lambda x: -1.0 * self.evaluate_sample(
                x, candidate_point, sample, var_noise, mean, params, w=w)
            g = lambda x: -1.0 * self.evaluate_gradient_sample(
                x, candidate_point, sample, var_noise, mean, params, w=w)
            opt = fmin_l_bfgs_b(f, start, fprime=g, bounds=self.bounds_x, factr=factr,
                                maxiter=maxiter)
            results.append((-1.0 * opt[1], opt[0]))
        values = [r[0] for r in results]
        max_ = np.max(values)
        arg_max = results[int(np.argmax(values))][1]
        if len(self.bounds_x) < 7:
            vertex = [np.array(p) for p in itertools.product(*self.bounds_x)]
            vals = [self.evaluate_sample(v, candidate_point, sample, var_noise, mean, params, w=w)
                    for v in vertex]
            if np.max(vals) > max_:
                max_ = np.max(vals)
                arg_max = vertex[int(np.argmax(vals))]
        return {"max": max_, "optimum": arg_max}

    def evaluate_mc_bayesian_candidate_points_no_restarts(
            self, candidate_points, parameters, samples, starts, w=None, max_mean=None,
            compute_gradient=True, factr=1e12, maxiter=10):
        """:523-670 with explicit hyper-samples `parameters`, Z `samples`, `starts`, W samples."""
        candidate_points = np.asarray(candidate_points, dtype=float)
        n_c = candidate_points.shape[0]
        n_samples = len(samples)
        evaluations = np.zeros(n_c)
        gradient = np.zeros((n_c, candidate_points.shape[1])) if compute_gradient else None
        for l in range(n_c):
            cand = candidate_points[l : l + 1, :]
            values_parameters = []
            gradients = []
            gradient_nan = False
            for k, param in enumerate(parameters):
                max_values = []
                optimum_values = np.zeros((n_samples, len(self.bq.x_domain)))
                for i in range(n_samples):
                    r = self.evaluate_sbo_by_sample(cand, samples[i], starts, param[0], param[1],
                                                    param[2:], w=w, factr=factr, maxiter=maxiter)
                    max_values.append(r["max"])
                    optimum_values[i, :] = r["optimum"]
                mm = 0.0 if max_mean is None else max_mean[k]
                values_parameters.append(np.mean(max_values) - mm)
                if compute_gradient and not gradient_nan:
                    gb = self.bq.gradient_vector_b(cand, optimum_values, param[0], param[1],
                                                   param[2:], w=w)
                    if gb is np.nan:
                        gradient_nan = True
                    else:
                        gradients.append(np.mean(gb * np.asarray(samples).reshape((n_samples, 1)),
                                                 axis=0))
            if compute_gradient:
                if gradient_nan:
                    gradient[l, :] = np.nan
                else:
                    gradient[l, :] = np.mean(gradients, axis=0)
            evaluations[l] = np.mean(values_parameters)
        return {"evaluations": evaluations, "gradient": gradient}

    def evaluate(self, point, var_noise, mean, params):
        """Discretised SBO / KG: :1394-1425."""
        vectors = self.bq.compute_posterior_parameters_kg(self.discretization, point, var_noise,
                                                          mean, params)
        a, b = vectors["a"], vectors["b"]
        if not np.all(np.isfinite(b)):
            return 0.0
        a, b, keep = affine_break_points_prep(a, b)
        keep1, c = affine_break_points(a, b)
        keep1 = keep1.astype(np.int64)
        return hvoi(b, c, keep1)

    def evaluate_gradient(self, point, var_noise, mean, params):
        """:1428-1474."""
        point = np.asarray(point, dtype=float).reshape(1, -1)
        vectors = self.bq.compute_posterior_parameters_kg(self.discretization, point, var_noise,
                                                          mean, params)
        a, b = vectors["a"], vectors["b"]
        a, b, keep = affine_break_points_prep(a, b)
        keep1, c = affine_break_points(a, b)
        keep1 = keep1.astype(np.int64)
        M = len(keep1)
        if M <= 1:
            return np.zeros(point.shape[1])
        keep = keep[keep1]
        c = c[keep1 + 1]
        c2 = np.abs(c[0 : M - 1])
        evalC = norm.pdf(c2)
        gradients = self.bq.gradient_vector_b(point, self.discretization[keep, :], var_noise, mean,
                                              params)
        gradient = np.zeros(point.shape[1])
        for i in range(point.shape[1]):
            gradient[i] = np.dot(np.diff(gradients[:, i]), evalC)
        return gradient


# --------------------------------------------------------------------------------------------
def sgd_adam(start, gradient, n, bounds=None, learning_rate=0.1, maxepoch=250, betas=(0.9, 0.999),
             eps=1e-8):
    """SGD(adam=True) of sbo/lib/stochastic_gradient_descent.py:12-132 for non-NaN gradients."""
    m0 = np.zeros(len(start))
    v0 = np.zeros(len(start))
    point = np.array(start, dtype=float)
    t_ = 0
    for _ in range(maxepoch):
        previous = point.copy()
        t_ += 1
        grad = [gradient(point) for _ in range(n)]
        gradient_ = np.mean(np.array(grad), axis=0)
        m0 = betas[0] * m0 + (1 - betas[0]) * gradient_
        v0 = betas[1] * v0 + (1 - betas[1]) * (gradient_ ** 2)
        m_1 = m0 / (1 - (betas[0]) ** (t_))
        v_1 = v0 / (1 - (betas[1]) ** (t_))
        point = point - learning_rate * m_1 / (np.sqrt(v_1) + eps)
        if bounds is not None:
            for dim, bound in enumerate(bounds):
                if bound[0] is not None:
                    point[dim] = max(bound[0], point[dim])
                if bound[1] is not None:
                    point[dim] = min(bound[1], point[dim])
        den_norm = np.sqrt(np.sum(previous ** 2))
        if den_norm == 0:
            nrm = np.sqrt(np.sum((previous - point) ** 2)) / 1e-2
        else:
            nrm = np.sqrt(np.sum((previous - point) ** 2)) / den_norm
        if nrm < 0.01:
            break
    return point
