"""BASELINE.json configs C1 and C2 as such (missing #5): the reference's own toy problem
(n = 100, theta = [1, 5, 50, 9.6, -3, -0.1] of tests/acquisition_functions/test_sbo.py:523-528) and
the branin_mt shape (T = 12 tasks, n = 200, weighted_uniform_finite, same_correlation), against
fixtures produced by the unmodified reference (tests/golden/make_golden_configs.py).

CPU: the oracle against the fixtures.  GPU: the drop-in classes (CUDA path) against the fixtures,
including the chosen task index of MultiTasks.choose_best_task_given_x (bit-exact).
Quantities that pass through Sigma^-1 use max(1e-9, 8 cond(Sigma) eps): C1's first theta is the
reference's own test vector, whose Sigma has cond ~ 1e9.
"""
import numpy as np
import pytest

from oracle import bqo_oracle as orc
from tests.helpers import assert_close, oracle_from_golden

CONFIGS = ["config_c1", "config_c2"]
EPS = np.finfo(float).eps


def _rtol(g, spec, s, base=1e-9):
    th = g["thetas"][s]
    cov = spec.cov(th[2:], g["points"]) + th[0] * np.eye(len(g["evaluations"]))
    return max(base, 8.0 * np.linalg.cond(cov) * EPS)


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CONFIGS)
def test_oracle_against_reference_at_config(golden, name):
    g = golden(name)
    spec, gp, bq = oracle_from_golden(g)
    xq, cands = g["xq"], g["cands"]
    for s, th in enumerate(g["thetas"]):
        vn, mu, pk = th[0], th[1], th[2:]
        tol = _rtol(g, spec, s)
        gp.clean_cache()
        chol, _ = gp.chol_cov_including_noise(vn, pk)
        assert_close(chol, g["chol"][s], rtol=max(1e-11, tol), atol=1e-14, what="chol")
        assert_close(gp.log_likelihood(vn, mu, pk), g["llh"][s], rtol=1e-11, what="llh")
        assert_close(gp.grad_log_likelihood(vn, mu, pk), g["grad_llh"][s], rtol=max(1e-8, tol),
                     atol=1e-8, what="grad_llh")
        post = gp.compute_posterior_parameters(g["query"], vn, mu, pk)
        assert_close(post["mean"], g["post_mean"][s], rtol=tol, what="post mean")
        for q in range(g["query"].shape[0]):
            pt = g["query"][q:q + 1, :]
            assert_close(orc.ei_evaluate(gp, pt, vn, mu, pk)[0], g["ei"][s, q], rtol=max(1e-8, tol),
                         atol=1e-300, what="ei")
        for p in range(xq.shape[0]):
            x = xq[p:p + 1, :]
            assert_close(bq.evaluate_quadrature_cross_cov(x, g["points"], pk), g["B"][s, p],
                         rtol=1e-12, what="B")
            post = bq.compute_posterior_parameters(x, vn, mu, pk)
            assert_close(post["mean"][0], g["q_mean"][s, p], rtol=tol, what="q mean")
        for c in range(cands.shape[0]):
            cand = cands[c:c + 1, :]
            par = bq.get_parameters_for_samples(cand, vn, mu, pk)
            assert_close(par["gamma"][:, 0], g["gamma"][s, c], rtol=1e-13)
            assert_close(par["denominator"][0], g["den"][s, c], rtol=max(1e-9, tol))
            for p in range(xq.shape[0]):
                x = xq[p:p + 1, :]
                v = bq.compute_parameters_for_sample(x, cand, vn, mu, pk)
                gr = bq.compute_gradient_parameters_for_sample(x, cand, vn, mu, pk)
                got = np.concatenate([[v["a"][0]], [np.asarray(v["b"]).reshape(-1)[0]], gr["a"],
                                      gr["b"]])
                assert_close(got, g["ab"][s, c, p], rtol=tol, atol=1e-11, what="ab")
            gvb = bq.gradient_vector_b(cand, xq, vn, mu, pk)
            assert_close(gvb, g["gvb"][s, c], rtol=10 * tol, atol=1e-11, what="gvb")
    # chosen task: argmax over tasks of the EI averaged over the hyper-samples
    for i, x in enumerate(g["task_x"]):
        vals = []
        for t in range(int(g["n_tasks"])):
            pt = np.concatenate([x, [t]]).reshape(1, -1)
            vals.append(np.mean([orc.ei_evaluate(gp, pt, th[0], th[1], th[2:])[0]
                                 for th in g["thetas"]]))
        assert int(np.argmax(vals)) == int(g["chosen_task"][i])


# ---------------------------------------------------------------------------------------------------
def _models(g):
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import \
        BayesianQuadrature
    from bayesian_quadrature_optimization_b200.lib import constant as K
    dim = g["points"].shape[1]
    dx = int(g["dx"])
    T = int(g["n_tasks"])
    hi = 100.0 if g["points"][:, 0].max() > 2 else 1.0
    td = {"points": g["points"].tolist(), "evaluations": list(g["evaluations"]), "var_noise": []}
    gp = GPFittingGaussian(
        [K.PRODUCT_KERNELS_SEPARABLE, K.MATERN52_NAME, K.TASKS_KERNEL_NAME], td, [dim, dx, T],
        bounds_domain=[[0, hi]] * dx + [list(range(T))], type_bounds=[0] * dx + [1],
        define_samplers=False, **{K.SAME_CORRELATION: bool(int(g["same_correlation"]))})
    if "weights" in g:
        params = {"weights": g["weights"], "domain_random": np.arange(T).reshape(T, 1)}
        dist = K.WEIGHTED_UNIFORM_FINITE
    else:
        params, dist = {K.TASKS: T}, K.UNIFORM_FINITE
    bq = BayesianQuadrature(gp, list(range(dx)), dist, parameters_distribution=params)
    bq_x = BayesianQuadrature(gp, list(range(dx)), dist, parameters_distribution=params,
                              model_only_x=True)
    return gp, bq, bq_x


@pytest.mark.gpu
@pytest.mark.parametrize("name", CONFIGS)
def test_gpu_classes_against_reference_at_config(golden, name):
    from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
    from bayesian_quadrature_optimization_b200.acquisition_functions.ei import EI
    g = golden(name)
    spec, _, _ = oracle_from_golden(g)
    gp, bq, _ = _models(g)
    sbo = SBO(bq, g["discretization"])
    ei = EI(gp)
    xq, cands, zs = g["xq"], g["cands"], g["zs"]
    for s, th in enumerate(g["thetas"]):
        vn, mu, pk = th[0], th[1], th[2:]
        tol = _rtol(g, spec, s)
        chol, _ = gp._chol_cov_including_noise(vn, pk)
        assert_close(chol, g["chol"][s], rtol=max(1e-10, tol), atol=1e-13, what="chol")
        # the quadratic term of the log-likelihood passes through Sigma^-1: cond(Sigma) eps
        # (1e-11 at the well-conditioned vector, 1.3e-9 observed at cond = 1e9 with the blocked
        # factorisation, whose panel solves multiply by the inverted 64 x 64 diagonal blocks)
        assert_close(gp.log_likelihood(vn, mu, pk), g["llh"][s], rtol=max(1e-11, tol), what="llh")
        assert_close(gp.grad_log_likelihood(vn, mu, pk), g["grad_llh"][s], rtol=max(1e-8, tol),
                     atol=1e-8, what="grad llh")
        post = gp.compute_posterior_parameters(g["query"], vn, mu, pk)
        assert_close(post["mean"], g["post_mean"][s], rtol=tol, what="post mean")
        # k(P,P) - k* Sigma^-1 k*^T cancels to ~1e-7 of its terms at the ill-conditioned theta
        assert_close(post["cov"], g["post_cov"][s], rtol=max(1e-8, 10 * tol),
                     atol=max(1e-10, tol * np.max(np.abs(np.diag(g["post_cov"][s])))), what="post cov")
        for q in range(g["query"].shape[0]):
            pt = g["query"][q:q + 1, :]
            assert_close(ei.evaluate(pt, vn, mu, pk)[0], g["ei"][s, q], rtol=max(1e-8, tol),
                         atol=1e-300, what="ei")
        for p in range(xq.shape[0]):
            x = xq[p:p + 1, :]
            assert_close(bq.evaluate_quadrature_cross_cov(x, g["points"], pk), g["B"][s, p],
                         rtol=1e-12, what="B")
            assert_close(bq.evaluate_grad_quadrature_cross_cov(x, g["points"], pk),
                         g["grad_B"][s, p], rtol=1e-10, atol=1e-16, what="grad B")
            full = bq.compute_posterior_parameters(x, vn, mu, pk)
            assert_close(full["mean"][0], g["q_mean"][s, p], rtol=tol, what="q mean")
            assert_close(full["cov"], g["q_var"][s, p], rtol=max(1e-8, tol), atol=1e-10, what="q var")
            assert_close(bq.gradient_posterior_mean(x, vn, mu, pk), g["grad_q_mean"][s, p],
                         rtol=tol, atol=1e-12, what="grad q mean")
        for c in range(cands.shape[0]):
            cand = cands[c:c + 1, :]
            par = bq.get_parameters_for_samples(True, cand, pk, vn, mu)
            assert_close(par["gamma"][:, 0], g["gamma"][s, c], rtol=1e-13, what="gamma")
            assert_close(par["solve_2"][:, 0], g["solve_2"][s, c], rtol=max(1e-8, tol),
                         atol=max(1e-11, tol * np.max(np.abs(g["solve_2"][s, c]))), what="solve_2")
            assert_close(par["denominator"][0], g["den"][s, c], rtol=max(1e-9, tol), what="den")
            for p in range(xq.shape[0]):
                x = xq[p:p + 1, :]
                v = bq.compute_parameters_for_sample(x, cand, vn, mu, pk)
                gr = bq.compute_gradient_parameters_for_sample(x, cand, vn, mu, pk)
                got = np.concatenate([v["a"], v["b"].reshape(-1), gr["a"], gr["b"]])
                # entries of one vector share the rounding error of the largest (cancellation)
                assert_close(got, g["ab"][s, c, p], rtol=tol,
                             atol=max(1e-11, tol * np.max(np.abs(g["ab"][s, c, p]))), what="ab")
                for iz, z in enumerate(zs):
                    assert_close(sbo.evaluate_sample(x, cand, z, vn, mu, pk),
                                 g["sample_val"][s, c, p, iz], rtol=tol, atol=1e-11, what="a + Z b")
            gvb = bq.gradient_vector_b(cand, xq, vn, mu, pk)
            assert_close(gvb, g["gvb"][s, c], rtol=10 * tol,
                         atol=max(1e-11, 10 * tol * np.max(np.abs(g["gvb"][s, c]))), what="gvb")
            assert_close(sbo.evaluate(cand, vn, mu, pk), g["kg"][s, c], rtol=max(1e-8, tol),
                         atol=1e-12, what="kg")
            assert_close(sbo.evaluate_gradient(cand, vn, mu, pk), g["grad_kg"][s, c],
                         rtol=max(1e-7, 10 * tol), atol=1e-11, what="grad kg")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CONFIGS)
def test_gpu_chosen_task_index_is_the_reference_one(golden, name):
    """a40: MultiTasks.choose_best_task_given_x -- exact index, value within 1e-9 / cond."""
    from bayesian_quadrature_optimization_b200.acquisition_functions.multi_task import MultiTasks
    g = golden(name)
    spec, _, _ = oracle_from_golden(g)
    gp, _, bq_x = _models(g)
    gp.samples_parameters = [t.copy() for t in g["thetas"]]
    mk = MultiTasks(bq_x, int(g["n_tasks"]))
    tol = max(_rtol(g, spec, s) for s in range(len(g["thetas"])))
    for i, x in enumerate(g["task_x"]):
        task, value = mk.choose_best_task_given_x(np.asarray(x, dtype=float))
        assert int(task) == int(g["chosen_task"][i])
        assert_close(value, g["chosen_task_value"][i], rtol=max(1e-8, tol), atol=1e-300,
                     what="chosen task value")


@pytest.mark.gpu
def test_gpu_refined_solves_at_the_ill_conditioned_vector(golden):
    """C1's first theta (the reference's own test vector) has cond(Sigma) = 7e9.  The engine's
    blocked solves multiply by inverted diagonal blocks, so there it refines against the
    re-assembled Sigma (HyperBatch.solve): the result must be as close to a longdouble-refined
    solution as LAPACK's backward-stable solve is (the reference's path), not cond * 1e-12 away."""
    import scipy.linalg as sl
    import torch
    g = golden("config_c1")
    spec, _, _ = oracle_from_golden(g)
    gp, bq, _ = _models(g)
    th = g["thetas"][0]
    cov = spec.cov(th[2:], g["points"]) + th[0] * np.eye(len(g["evaluations"]))
    assert np.linalg.cond(cov) > 1e9
    hb = gp.hyper_batch(np.asarray([th]))
    assert list(hb.ill_conditioned()) == [0]
    rng = np.random.RandomState(5)
    rhs = np.vstack([g["gamma"][0, 0], rng.randn(3, cov.shape[0])])
    x = hb.solve(hb.eng.to_device(rhs[None]).clone())[0].cpu().numpy()
    L = np.linalg.cholesky(cov)
    covl = cov.astype(np.longdouble)
    for i in range(rhs.shape[0]):
        ref64 = sl.cho_solve((L, True), rhs[i])
        xl = ref64.astype(np.longdouble)
        for _ in range(6):
            r = rhs[i].astype(np.longdouble) - covl @ xl
            xl = xl + sl.cho_solve((L, True), np.asarray(r, dtype=float)).astype(np.longdouble)
        scale = float(np.max(np.abs(xl)))
        err_dev = float(np.max(np.abs(x[i] - xl))) / scale
        err_lapack = float(np.max(np.abs(ref64 - xl))) / scale
        assert err_dev <= max(4.0 * err_lapack, 1e-9), (i, err_dev, err_lapack)
    # the well-conditioned second vector is not touched by the refinement
    hb2 = gp.hyper_batch(np.asarray([g["thetas"][1]]))
    assert len(hb2.ill_conditioned()) == 0
