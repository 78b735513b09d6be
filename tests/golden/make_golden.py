"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Each case builds the reference's own objects (GPFittingGaussian, BayesianQuadrature, SBO, EI)
through tests/golden/ref_shim.py, evaluates the hot-path methods on seeded inputs with explicit
hyper-parameter vectors, and stores inputs + outputs.  The oracle (oracle/bqo_oracle.py) and the
CUDA path are both checked against these files.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
warnings.filterwarnings("ignore")

from stratified_bayesian_optimization.models.gp_fitting_gaussian import GPFittingGaussian  # noqa: E402
from stratified_bayesian_optimization.numerical_tools.bayesian_quadrature import BayesianQuadrature  # noqa: E402
from stratified_bayesian_optimization.acquisition_functions.sbo import SBO  # noqa: E402
from stratified_bayesian_optimization.acquisition_functions.ei import EI  # noqa: E402
from stratified_bayesian_optimization.lib.constant import (  # noqa: E402
    PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME, SCALED_KERNEL, UNIFORM_FINITE,
    WEIGHTED_UNIFORM_FINITE, GAMMA, TASKS, SAME_CORRELATION)


def gp_outputs(gp, thetas, query, out):
    covs, gcovs, chols, llh, gllh = [], [], [], [], []
    pm, pc, gpm, gpc = [], [], [], []
    for th in thetas:
        vn, mu, pk = th[0], th[1], th[2:]
        covs.append(gp.evaluate_cov(gp.data["points"], pk))
        gc = gp.evaluate_grad_cov(pk, gp.data["points"])
        gcovs.append(np.stack([gc[i] for i in range(len(pk))]))
        gp.cache_chol_cov = {}
        gp.cache_sol_chol_y_unbiased = {}
        chol, _ = gp._chol_cov_including_noise(vn, pk)
        chols.append(chol)
        llh.append(gp.log_likelihood(vn, mu, pk))
        gllh.append(gp.grad_log_likelihood(vn, mu, pk))
        post = gp.compute_posterior_parameters(query, vn, mu, pk)
        pm.append(post["mean"])
        pc.append(post["cov"])
        gm, gv = [], []
        for q in range(query.shape[0]):
            gp.cache_cov_n = {}
            g = gp.gradient_posterior_parameters(query[q:q + 1, :], vn, mu, pk)
            gm.append(g["mean"])
            gv.append(np.asarray(g["cov"]).reshape(-1))
        gpm.append(np.stack(gm))
        gpc.append(np.stack(gv))
    out.update(cov=np.stack(covs), grad_cov=np.stack(gcovs), chol=np.stack(chols),
               llh=np.array(llh), grad_llh=np.stack(gllh), post_mean=np.stack(pm),
               post_cov=np.stack(pc), grad_post_mean=np.stack(gpm), grad_post_cov=np.stack(gpc))


def ei_outputs(gp, thetas, query, out):
    ei = EI(gp)
    vals, grads = [], []
    for th in thetas:
        vn, mu, pk = th[0], th[1], th[2:]
        v, g = [], []
        for q in range(query.shape[0]):
            gp.cache_cov_n = {}
            v.append(ei.evaluate(query[q:q + 1, :], vn, mu, pk)[0])
            gp.cache_cov_n = {}
            g.append(ei.evaluate_gradient(query[q:q + 1, :], vn, mu, pk))
        vals.append(v)
        grads.append(np.stack(g))
    out.update(ei=np.array(vals), grad_ei=np.stack(grads))


def clear_bq(bq):
    bq.cache_quadratures = {}
    bq.cache_posterior_mean = {}
    bq.cache_quadrature_with_candidate = {}
    bq.cache_sample = {}
    bq.gp.cache_chol_cov = {}
    bq.gp.cache_sol_chol_y_unbiased = {}


def bq_outputs(bq, sbo, thetas, xq, cands, zs, out):
    """xq: [P][dx] points; cands: [C][dim]."""
    S, P, C = len(thetas), xq.shape[0], cands.shape[0]
    n = bq.gp.data["points"].shape[0]
    dx = xq.shape[1]
    dim = cands.shape[1]
    B = np.zeros((S, P, n))
    gB = np.zeros((S, P, dx, n))
    gamma = np.zeros((S, C, n))
    s2 = np.zeros((S, C, n))
    den = np.zeros((S, C))
    ab = np.zeros((S, C, P, 2 + 2 * dx))
    sample_val = np.zeros((S, C, P, len(zs)))
    gvb = np.zeros((S, C, P, dim))
    qmean = np.zeros((S, P))
    qvar = np.zeros((S, P))
    gqmean = np.zeros((S, P, dx))
    for s, th in enumerate(thetas):
        vn, mu, pk = th[0], th[1], th[2:]
        for p in range(P):
            x = xq[p:p + 1, :]
            clear_bq(bq)
            B[s, p] = bq.evaluate_quadrature_cross_cov(x, bq.gp.data["points"], pk)
            gB[s, p] = bq.evaluate_grad_quadrature_cross_cov(x, bq.gp.data["points"], pk)
            clear_bq(bq)
            post = bq.compute_posterior_parameters(x, vn, mu, pk)
            qmean[s, p] = post["mean"][0]
            qvar[s, p] = post["cov"]
            gqmean[s, p] = bq.gradient_posterior_mean(x, vn, mu, pk)
        for c in range(C):
            cand = cands[c:c + 1, :]
            clear_bq(bq)
            par = bq.get_parameters_for_samples(True, cand, pk, vn, mu)
            gamma[s, c] = par["gamma"][:, 0]
            s2[s, c] = par["solve_2"][:, 0]
            den[s, c] = par["denominator"][0]
            for p in range(P):
                x = xq[p:p + 1, :]
                clear_bq(bq)
                v = bq.compute_parameters_for_sample(x, cand, vn, mu, pk)
                clear_bq(bq)
                g = bq.compute_gradient_parameters_for_sample(x, cand, vn, mu, pk)
                ab[s, c, p, 0] = v["a"][0]
                ab[s, c, p, 1] = np.asarray(v["b"]).reshape(-1)[0]
                ab[s, c, p, 2:2 + dx] = g["a"]
                ab[s, c, p, 2 + dx:] = g["b"]
                for iz, z in enumerate(zs):
                    clear_bq(bq)
                    sample_val[s, c, p, iz] = sbo.evaluate_sample(x, cand, z, vn, mu, pk)
            clear_bq(bq)
            g = bq.gradient_vector_b(cand, xq, vn, mu, pk, cache=False, parallel=False,
                                     monte_carlo=True)
            gvb[s, c] = g
    out.update(B=B, grad_B=gB, gamma=gamma, solve_2=s2, den=den, ab=ab, sample_val=sample_val,
               gvb=gvb, q_mean=qmean, q_var=qvar, grad_q_mean=gqmean)


def kg_outputs(bq, sbo, thetas, cands, out):
    vals, grads = [], []
    for th in thetas:
        vn, mu, pk = th[0], th[1], th[2:]
        v, g = [], []
        for c in range(cands.shape[0]):
            cand = cands[c:c + 1, :]
            clear_bq(bq)
            v.append(sbo.evaluate(cand, vn, mu, pk, cache=False))
            clear_bq(bq)
            g.append(sbo.evaluate_gradient(cand, vn, mu, pk, cache=False))
        vals.append(v)
        grads.append(np.stack(g))
    out.update(kg=np.array(vals), grad_kg=np.stack(grads))


def case_product_uniform():
    """Product(Matern52(dx=1), Tasks(T=2, general)), uniform_finite W -- the shape of the
    reference's test_problem_with_tasks (tests/acquisition_functions/test_sbo.py:57-78)."""
    rng = np.random.RandomState(5)
    n = 30
    x = np.linspace(0, 100, n)
    t = rng.randint(0, 2, n)
    pts = np.stack([x, t.astype(float)], 1)
    y = x / 10.0 + np.where(t == 0, 10.0, -10.0) + rng.normal(0, 0.1, n)
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}
    gp = GPFittingGaussian([PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME], td,
                           [2, 1, 2], bounds_domain=[[0, 100], [0, 1]], type_bounds=[0, 1])
    thetas = np.array([[1.0, 5.0, 50.0, 9.6, -3.0, -0.1],
                       [0.5, 4.0, 30.0, 1.2, 0.3, 0.8],
                       [2.0, 6.0, 80.0, 0.1, -0.5, 0.4]])
    query = np.array([[3.0, 0.0], [47.5, 1.0], [99.0, 0.0], [60.1, 1.0]])
    out = dict(points=pts, evaluations=y, thetas=thetas, query=query, kind=2, n_tasks=2,
               same_correlation=0, dx=1)
    gp_outputs(gp, thetas, query, out)
    ei_outputs(gp, thetas, query, out)
    bq = BayesianQuadrature(gp, [0], UNIFORM_FINITE, parameters_distribution={TASKS: 2})
    disc = np.linspace(0, 100, 21).reshape(-1, 1)
    sbo = SBO(bq, disc)
    xq = np.array([[3.0], [20.0], [77.7]])
    cands = np.array([[55.0, 1.0], [10.0, 0.0]])
    zs = np.array([0.7, -1.3])
    out.update(xq=xq, cands=cands, zs=zs, discretization=disc)
    bq_outputs(bq, sbo, thetas, xq, cands, zs, out)
    kg_outputs(bq, sbo, thetas, cands, out)
    return out


def case_product_weighted():
    """Product(Matern52(dx=2), Tasks(T=4, same_correlation)), weighted_uniform_finite W, with
    observation noise variances -- the branin_mt shape (problems/branin_mt/main.py)."""
    rng = np.random.RandomState(1)
    n, T = 40, 4
    x = rng.uniform(0, 1, (n, 2))
    t = rng.randint(0, T, n)
    pts = np.concatenate([x, t.reshape(-1, 1).astype(float)], 1)
    y = np.sin(3 * x[:, 0]) + np.cos(2 * x[:, 1]) + 0.3 * t + rng.normal(0, 0.05, n)
    vno = rng.uniform(0.01, 0.05, n)
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": list(vno)}
    gp = GPFittingGaussian([PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME], td,
                           [3, 2, T], bounds_domain=[[0, 1], [0, 1], list(range(T))],
                           type_bounds=[0, 0, 1], noise=True, **{SAME_CORRELATION: True})
    thetas = np.array([[0.0, 0.0, 0.3, 0.4, 0.0, -1.0],
                       [0.0, 0.0, 0.25, 0.5, 0.2, -0.6],
                       [0.0, 0.0, 0.5, 0.2, -0.3, -1.5]])
    query = np.concatenate([rng.uniform(0, 1, (4, 2)), np.array([[0.], [3.], [1.], [2.]])], 1)
    weights = np.array([0.1, 0.4, 0.3, 0.2])
    out = dict(points=pts, evaluations=y, var_noise_obs=vno, thetas=thetas, query=query, kind=2,
               n_tasks=T, same_correlation=1, dx=2, weights=weights)
    gp_outputs(gp, thetas, query, out)
    ei_outputs(gp, thetas, query, out)
    bq = BayesianQuadrature(gp, [0, 1], WEIGHTED_UNIFORM_FINITE,
                            parameters_distribution={"weights": weights,
                                                     "domain_random": np.arange(T).reshape(T, 1)})
    g = np.linspace(0, 1, 5)
    disc = np.array([[a, b] for a in g for b in g])
    sbo = SBO(bq, disc)
    xq = rng.uniform(0, 1, (3, 2))
    cands = np.array([[0.3, 0.8, 2.0], [0.9, 0.1, 0.0]])
    zs = np.array([0.4, -0.9])
    out.update(xq=xq, cands=cands, zs=zs, discretization=disc)
    bq_outputs(bq, sbo, thetas, xq, cands, zs, out)
    kg_outputs(bq, sbo, thetas, cands, out)
    return out


def case_scaled_gamma():
    """Scaled(Matern52) over (x in R^2, w in R^2), gamma W by Monte Carlo with M=10 replayed
    samples (every gamma.rvs call of sbo/lib/expectations.py returns the same pre-drawn set)."""
    rng = np.random.RandomState(0)
    n = 50
    x = rng.uniform(0, 1, (n, 2))
    w = rng.gamma(2.0, 1.0, (n, 2))
    pts = np.concatenate([x, w], 1)
    y = np.sin(pts.sum(1)) + 0.01 * rng.normal(0, 1, n)
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}
    gp = GPFittingGaussian([SCALED_KERNEL, MATERN52_NAME], td, [4],
                           bounds_domain=[[0, 1], [0, 1], [0, 10], [0, 10]], noise=True,
                           define_samplers=False)
    thetas = np.array([[1e-4, 0.0, 0.3, 0.35, 2.0, 2.2, 1.0],
                       [2e-4, 0.1, 0.4, 0.25, 1.5, 2.5, 1.3],
                       [5e-5, -0.1, 0.2, 0.5, 2.5, 1.8, 0.8]])
    query = np.concatenate([rng.uniform(0, 1, (4, 2)), rng.gamma(2.0, 1.0, (4, 2))], 1)
    wset = np.random.RandomState(4).gamma(2.0, 1.0, (10, 2))
    out = dict(points=pts, evaluations=y, thetas=thetas, query=query, kind=1, n_tasks=0,
               same_correlation=0, dx=2, wset=wset)
    gp_outputs(gp, thetas, query, out)
    ei_outputs(gp, thetas, query, out)
    bq = BayesianQuadrature(gp, [0, 1], GAMMA, parameters_distribution={"a": [2.0], "scale": [1.0]})
    sbo = SBO(bq, None)
    xq = rng.uniform(0, 1, (3, 2))
    cands = np.array([[0.3, 0.8, 1.5, 2.5], [0.9, 0.1, 0.7, 3.0]])
    zs = np.array([0.4, -0.9])
    out.update(xq=xq, cands=cands, zs=zs)
    ref_shim.patch_gamma([wset])
    # the double expectation of evaluate_quadrate_cov draws (100, 2) samples: replay two sets
    w100 = np.random.RandomState(7).gamma(2.0, 1.0, (2, 100, 2))
    out.update(w100=w100)

    class Mixed(object):
        def rvs(self, a, scale=1.0, size=None):
            if tuple(size) == (10, 2):
                return wset.copy()
            self.k = getattr(self, "k", 0)
            r = w100[self.k % 2].copy()
            self.k += 1
            return r

    import stratified_bayesian_optimization.lib.expectations as ex
    ex.gamma = Mixed()
    bq_outputs(bq, sbo, thetas, xq, cands, zs, out)
    return out


def case_matern52_plain():
    """Plain Matern52 GP (no scale, no tasks) in R^3 with a noise parameter: the GP-only rows
    (covariance, gradients, Cholesky, log-likelihood, posterior, EI) for kernel kind 0."""
    rng = np.random.RandomState(12)
    n = 45
    pts = rng.uniform(-1, 2, (n, 3))
    y = np.cos(pts[:, 0]) * pts[:, 1] + 0.3 * pts[:, 2] + 0.05 * rng.normal(0, 1, n)
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}
    gp = GPFittingGaussian([MATERN52_NAME], td, [3], bounds_domain=[[-1, 2]] * 3, noise=True,
                           define_samplers=False)
    thetas = np.array([[1e-3, 0.2, 0.8, 1.1, 0.6],
                       [5e-3, -0.1, 1.5, 0.7, 0.9],
                       [2e-4, 0.0, 0.5, 0.5, 2.0]])
    query = rng.uniform(-1, 2, (5, 3))
    out = dict(points=pts, evaluations=y, thetas=thetas, query=query, kind=0, n_tasks=0,
               same_correlation=0, dx=3)
    gp_outputs(gp, thetas, query, out)
    ei_outputs(gp, thetas, query, out)
    return out


def main():
    cases = [("product_uniform", case_product_uniform),
             ("product_weighted", case_product_weighted),
             ("scaled_gamma", case_scaled_gamma),
             ("matern52_plain", case_matern52_plain)]
    only = set(sys.argv[1:])  # optional: regenerate the named cases only
    for name, fn in cases:
        if only and name not in only:
            continue
        out = fn()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in out.items()})
        print(name, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
