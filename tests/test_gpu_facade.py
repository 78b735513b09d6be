"""GPU tests of the reference-facing Python API (GPFittingGaussian, BayesianQuadrature, SBO, EI,
MultiTasks, bgo): same calls as the reference's own tests make, checked against the fixtures
generated from the unmodified reference (tests/golden) and against the oracle."""
import numpy as np
import pytest

from oracle import bqo_oracle as orc
from tests.helpers import assert_close, cond_rtol, oracle_from_golden

pytestmark = pytest.mark.gpu

CASES = ["product_uniform", "product_weighted", "scaled_gamma"]


def _models(g):
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import \
        BayesianQuadrature
    from bayesian_quadrature_optimization_b200.lib import constant as K
    kind = int(g["kind"])
    dim = g["points"].shape[1]
    dx = int(g["dx"])
    vno = list(g["var_noise_obs"]) if "var_noise_obs" in g else []
    td = {"points": g["points"].tolist(), "evaluations": list(g["evaluations"]), "var_noise": vno}
    hi = 100.0 if g["points"][:, 0].max() > 2 else 1.0
    th0 = g["thetas"][0]
    if kind == 2:
        T = int(g["n_tasks"])
        gp = GPFittingGaussian(
            [K.PRODUCT_KERNELS_SEPARABLE, K.MATERN52_NAME, K.TASKS_KERNEL_NAME], td, [dim, dx, T],
            bounds_domain=[[0, hi]] * dx + [list(range(T))], type_bounds=[0] * dx + [1],
            noise=len(vno) > 0, kernel_values=list(th0[2:]), define_samplers=False,
            **{K.SAME_CORRELATION: bool(int(g["same_correlation"]))})
        if "weights" in g:
            bq = BayesianQuadrature(gp, list(range(dx)), K.WEIGHTED_UNIFORM_FINITE,
                                    parameters_distribution={
                                        "weights": g["weights"],
                                        "domain_random": np.arange(T).reshape(T, 1)})
        else:
            bq = BayesianQuadrature(gp, list(range(dx)), K.UNIFORM_FINITE,
                                    parameters_distribution={K.TASKS: T})
    else:
        gp = GPFittingGaussian([K.SCALED_KERNEL, K.MATERN52_NAME], td, [dim],
                               bounds_domain=[[0, 1]] * dx + [[0, 10]] * (dim - dx), noise=True,
                               kernel_values=list(th0[2:]), mean_value=[th0[1]],
                               var_noise_value=[th0[0]], define_samplers=False)
        bq = BayesianQuadrature(gp, list(range(dx)), K.GAMMA,
                                parameters_distribution={"a": [2.0], "scale": [1.0]})
        bq.set_w_sets(g["wset"])
        bq.set_w_double_sets(g["w100"])
    return gp, bq


@pytest.mark.parametrize("name", CASES)
def test_gp_facade_against_reference(golden, name):
    g = golden(name)
    gp, _ = _models(g)
    for s, th in enumerate(g["thetas"]):
        vn, mu, pk = th[0], th[1], th[2:]
        assert_close(gp.evaluate_cov(g["points"], pk), g["cov"][s], rtol=1e-13)
        gc = gp.evaluate_grad_cov(pk, g["points"])
        assert_close(np.stack([gc[i] for i in range(len(pk))]), g["grad_cov"][s], rtol=1e-10,
                     atol=1e-15)
        chol, cov = gp._chol_cov_including_noise(vn, pk)
        assert_close(chol, g["chol"][s], rtol=1e-10, atol=1e-14)
        assert isinstance(gp.log_likelihood(vn, mu, pk), float)
        assert_close(gp.log_likelihood(vn, mu, pk), g["llh"][s], rtol=1e-11)
        assert_close(gp.grad_log_likelihood(vn, mu, pk), g["grad_llh"][s], rtol=1e-8, atol=1e-9)
        post = gp.compute_posterior_parameters(g["query"], vn, mu, pk)
        assert_close(post["mean"], g["post_mean"][s], rtol=1e-10)
        assert_close(post["cov"], g["post_cov"][s], rtol=max(1e-8, 10 * cond_rtol(g, s)), atol=1e-11)
        one = gp.compute_posterior_parameters(g["query"][0:1], vn, mu, pk)
        assert one["cov"].shape == (1, 1)
        gr = gp.gradient_posterior_parameters(g["query"][1:2], vn, mu, pk)
        assert_close(gr["mean"], g["grad_post_mean"][s, 1], rtol=1e-9, atol=1e-13)
        assert_close(gr["cov"].reshape(-1), g["grad_post_cov"][s, 1], rtol=1e-8, atol=1e-12)
    sol = gp._cholesky_solve_vectors_for_posterior(*[g["thetas"][0][0], g["thetas"][0][1],
                                                    g["thetas"][0][2:]])
    assert sol["chol"].shape == g["chol"][0].shape and sol["solve"].shape == (len(g["points"]),)
    llb = gp.log_likelihood_batch(g["thetas"])
    assert_close(llb, g["llh"], rtol=1e-11)


@pytest.mark.parametrize("name", CASES)
def test_quadrature_and_sbo_facade_against_reference(golden, name):
    from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
    g = golden(name)
    gp, bq = _models(g)
    disc = g["discretization"] if "discretization" in g else None
    sbo = SBO(bq, disc)
    xq, cands, zs = g["xq"], g["cands"], g["zs"]
    for s, th in enumerate(g["thetas"]):
        vn, mu, pk = th[0], th[1], th[2:]
        tol = cond_rtol(g, s)
        for p in range(xq.shape[0]):
            x = xq[p:p + 1, :]
            assert_close(bq.evaluate_quadrature_cross_cov(x, g["points"], pk), g["B"][s, p],
                         rtol=1e-12)
            assert_close(bq.evaluate_grad_quadrature_cross_cov(x, g["points"], pk),
                         g["grad_B"][s, p], rtol=1e-10, atol=1e-16)
            post = bq.compute_posterior_parameters(x, vn, mu, pk, only_mean=True)
            assert_close(post["mean"][0], g["q_mean"][s, p], rtol=1e-10)
            assert_close(bq.gradient_posterior_mean(x, vn, mu, pk), g["grad_q_mean"][s, p],
                         rtol=1e-9, atol=1e-13)
            # prior variance: exact double sum (finite W) or the 100 x 100 grid of the replayed
            # gamma draws (a22, sbo/lib/expectations.py:93-112)
            full = bq.compute_posterior_parameters(x, vn, mu, pk)
            assert_close(full["cov"], g["q_var"][s, p], rtol=max(1e-8, tol), atol=1e-11)
        for c in range(cands.shape[0]):
            cand = cands[c:c + 1, :]
            par = bq.get_parameters_for_samples(True, cand, pk, vn, mu)
            assert_close(par["gamma"][:, 0], g["gamma"][s, c], rtol=1e-13)
            assert_close(par["solve_2"][:, 0], g["solve_2"][s, c], rtol=max(1e-8, tol), atol=1e-12)
            assert_close(par["denominator"][0], g["den"][s, c], rtol=1e-9)
            for p in range(xq.shape[0]):
                x = xq[p:p + 1, :]
                v = bq.compute_parameters_for_sample(x, cand, vn, mu, pk)
                gr = bq.compute_gradient_parameters_for_sample(x, cand, vn, mu, pk)
                got = np.concatenate([v["a"], v["b"].reshape(-1), gr["a"], gr["b"]])
                assert_close(got, g["ab"][s, c, p], rtol=tol, atol=1e-11)
                for iz, z in enumerate(zs):
                    val = sbo.evaluate_sample(x, cand, z, vn, mu, pk)
                    assert_close(val, g["sample_val"][s, c, p, iz], rtol=tol, atol=1e-11)
            gvb = bq.gradient_vector_b(cand, xq, vn, mu, pk)
            assert gvb.shape == g["gvb"][s, c].shape
            assert_close(gvb, g["gvb"][s, c], rtol=10 * tol, atol=1e-11)
            if disc is not None:
                assert_close(sbo.evaluate(cand, vn, mu, pk), g["kg"][s, c], rtol=1e-8, atol=1e-13)
                assert_close(sbo.evaluate_gradient(cand, vn, mu, pk), g["grad_kg"][s, c],
                             rtol=1e-7, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_ei_facade_and_bayesian_average(golden, name):
    from bayesian_quadrature_optimization_b200.acquisition_functions.ei import EI
    g = golden(name)
    gp, _ = _models(g)
    ei = EI(gp)
    vals = []
    for s, th in enumerate(g["thetas"]):
        vn, mu, pk = th[0], th[1], th[2:]
        for q in range(g["query"].shape[0]):
            pt = g["query"][q:q + 1, :]
            assert_close(ei.evaluate(pt, vn, mu, pk)[0], g["ei"][s, q], rtol=1e-8, atol=1e-300)
            assert_close(ei.evaluate_gradient(pt, vn, mu, pk), g["grad_ei"][s, q], rtol=1e-7,
                         atol=1e-300)
    gp.samples_parameters = [t.copy() for t in g["thetas"]]
    avg, _ = ei.evaluate_bayesian(g["query"], n_samples_parameters=3)
    assert_close(avg, g["ei"].mean(axis=0), rtol=1e-8, atol=1e-300)


@pytest.mark.parametrize("name", ["product_uniform", "product_weighted"])
def test_chosen_task_index_is_exact(golden, name):
    """MultiTasks.choose_best_task_given_x: argmax over tasks of the Bayesian-averaged EI of F at
    (x, t) -- the index must equal the oracle's (bit-exact contract of the north star)."""
    from bayesian_quadrature_optimization_b200.acquisition_functions.multi_task import MultiTasks
    g = golden(name)
    gp, bq = _models(g)
    gp.samples_parameters = [t.copy() for t in g["thetas"]]
    spec, ogp, obq = oracle_from_golden(g)
    T = int(g["n_tasks"])
    mt = MultiTasks(bq, T)
    rng = np.random.RandomState(0)
    dx = int(g["dx"])
    hi = 100.0 if g["points"][:, 0].max() > 2 else 1.0
    for _ in range(6):
        x = rng.uniform(0, hi, dx)
        task, value = mt.choose_best_task_given_x(x)
        want = []
        for t in range(T):
            pt = np.concatenate([x, [float(t)]]).reshape(1, -1)
            want.append(np.mean([orc.ei_evaluate(ogp, pt, th[0], th[1], th[2:])[0]
                                 for th in g["thetas"]]))
        assert task == int(np.argmax(want))
        assert_close(value, np.max(want), rtol=1e-8, atol=1e-300)
    sol = mt.optimize(random_seed=3, n_restarts=20, n_samples_parameters=0, start_new_chain=False)
    assert sol["solution"].shape == (dx + 1,) and 0 <= int(sol["solution"][-1]) < T


def test_mc_acquisition_through_public_api(golden):
    """evaluate_mc_bayesian_candidate_points_no_restarts and evaluate_sbo_by_sample through the
    reference-facing classes, against the oracle evaluated at the device's own maximisers."""
    from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
    g = golden("product_weighted")
    gp, bq = _models(g)
    gp.samples_parameters = [t.copy() for t in g["thetas"]]
    sbo = SBO(bq)
    spec, ogp, obq = oracle_from_golden(g)
    np.random.seed(4)
    out = sbo.evaluate_mc_bayesian_candidate_points_no_restarts(
        g["cands"], 3, 6, n_restarts=3, compute_max_mean=True, compute_gradient=True,
        factr=1e12, maxiter=10)
    assert out["evaluations"].shape == (2,) and out["gradient"].shape == g["cands"].shape
    assert np.all(np.isfinite(out["evaluations"])) and np.all(out["gradient"][:, -1] == 0.0)
    th = g["thetas"][1]
    r = sbo.evaluate_sbo_by_sample(g["cands"][0:1], 0.6, var_noise=th[0], mean=th[1],
                                   parameters_kernel=th[2:], n_restarts=4, factr=1e12, maxiter=10)
    v = obq.compute_parameters_for_sample(r["optimum"], g["cands"][0:1], th[0], th[1], th[2:])
    assert_close(r["max"], v["a"][0] + 0.6 * np.asarray(v["b"]).reshape(-1)[0], rtol=1e-9)
    # the maximum dominates every box vertex and random point
    osbo = orc.OracleSBO(obq, [(0.0, 1.0)] * 2)
    for x in np.random.RandomState(1).uniform(0, 1, (20, 2)):
        assert r["max"] >= osbo.evaluate_sample(x, g["cands"][0:1], 0.6, th[0], th[1], th[2:]) - 1e-3


def test_slice_sampling_chain_and_sbo_optimize():
    """Hyper-posterior chain on the GPU likelihood, then SBO.optimize (multi-start SGD)."""
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import \
        BayesianQuadrature
    from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
    from bayesian_quadrature_optimization_b200.lib import constant as K
    rng = np.random.RandomState(0)
    n, T = 24, 2
    x = rng.uniform(0, 1, (n, 1))
    t = rng.randint(0, T, n)
    pts = np.concatenate([x, t.reshape(-1, 1).astype(float)], 1)
    y = np.sin(4 * x[:, 0]) + 0.5 * t
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}

    def make(seed):
        return GPFittingGaussian(
            [K.PRODUCT_KERNELS_SEPARABLE, K.MATERN52_NAME, K.TASKS_KERNEL_NAME], td, [2, 1, T],
            bounds_domain=[[0, 1], [0, 1]], type_bounds=[0, 1], thinning=1, n_burning=4,
            max_steps_out=20, random_seed=seed, **{K.SAME_CORRELATION: True})

    gp = make(1)
    s1 = gp.sample_parameters(3, random_seed=7)
    gp2 = make(1)
    s2 = gp2.sample_parameters(3, random_seed=7)
    assert len(s1) == 3 and all(np.all(np.isfinite(v)) for v in s1)
    assert_close(np.array(s1), np.array(s2), rtol=1e-9)            # seeded chain reproduces
    for v in s1:
        assert v[0] == 0.0 and v[1] == 0.0 and v[2] > 0.0          # noise-free model, ls > 0
        assert np.isfinite(gp.log_prob_parameters(v))
    bq = BayesianQuadrature(gp, [0], K.UNIFORM_FINITE, parameters_distribution={K.TASKS: T})
    sbo = SBO(bq)
    res = sbo.optimize(random_seed=2, monte_carlo=True, n_samples=2, n_restarts_mc=2, n_restarts=2,
                       n_samples_parameters=2, start_new_chain=False, maxepoch=3,
                       default_n_samples=4, default_n_samples_parameters=3, factr=1e12, maxiter=10)
    assert res["solution"].shape == (2,) and 0.0 <= res["solution"][0] <= 1.0
    assert int(res["solution"][1]) in (0, 1) and np.isfinite(res["optimal_value"])


def test_bgo_end_to_end_small():
    from bayesian_quadrature_optimization_b200.services.bgo import bgo

    def g(x):
        return [-(x[0] - 0.3) ** 2]

    def f(xw):
        return [-(xw[0] - 0.3) ** 2 + (0.2 if int(xw[1]) == 0 else -0.2)]

    sol = bgo(g, [(0.0, 1.0)], integrand_function=f, bounds_domain_w=[[0, 1]], type_bounds=[0, 1],
              name_method='bqo', n_iterations=1, random_seed=3, n_training=6, thinning=1,
              n_burning=2, max_steps_out=10, n_restarts=2, n_samples_parameters=2, maxepoch=2,
              n_samples_mc=2, n_restarts_mc=2, default_n_samples=4, default_n_samples_parameters=3,
              n_restarts_mean=10)
    assert set(sol) == {"optimal_solution", "optimal_value"}
    assert sol["optimal_solution"].shape == (1,) and 0.0 <= sol["optimal_solution"][0] <= 1.0
    sol = bgo(g, [(0.0, 1.0)], name_method='ei', n_iterations=1, random_seed=3, n_training=5,
              thinning=1, n_burning=2, max_steps_out=10, n_restarts=3, n_samples_parameters=2,
              maxepoch=2, n_restarts_mean=10)
    assert 0.0 <= sol["optimal_solution"][0] <= 1.0 and np.isfinite(sol["optimal_value"])


@pytest.mark.parametrize("name", CASES)
def test_optimize_posterior_mean_quadrature_and_gp(golden, name):
    """a38: the maximisers return the reference's dict, stay in the box, dominate every start
    point / vertex, and their value is the oracle's posterior mean at the returned solution."""
    import itertools
    g = golden(name)
    gp, bq = _models(g)
    spec, ogp, obq = oracle_from_golden(g)
    th = g["thetas"][0]
    w = g["wset"] if "wset" in g else None
    dx = int(g["dx"])
    np.random.seed(3)
    res = bq.optimize_posterior_mean(n_restarts=8, n_best_restarts=4, var_noise=th[0], mean=th[1],
                                     parameters_kernel=th[2:])
    sol, val = res["solution"], res["optimal_value"]
    assert sol.shape == (dx,) and np.asarray(val).shape == (1,)
    bx = bq.bounds_x()
    assert all(bx[l][0] - 1e-12 <= sol[l] <= bx[l][-1] + 1e-12 for l in range(dx))
    want = obq.compute_posterior_parameters(sol.reshape(1, -1), th[0], th[1], th[2:],
                                            only_mean=True, w=w)["mean"]
    assert_close(val, np.asarray(want).reshape(-1), rtol=cond_rtol(g, 0), atol=1e-10,
                 what="posterior mean at the optimum")
    verts = np.array(list(itertools.product(*bx)), dtype=float)
    vmean = bq.compute_posterior_parameters(verts, th[0], th[1], th[2:], only_mean=True)["mean"]
    assert val[0] >= np.max(vmean) - 1e-10
    key = (th[0], th[1], tuple(th[2:]))
    assert bq.optimal_solutions[key][-1] is res and key in bq.max_mean
    # Bayesian average over the hyper-samples
    gp.samples_parameters = [t.copy() for t in g["thetas"]]
    res2 = bq.optimize_posterior_mean(n_restarts=6, n_best_restarts=3, n_samples_parameters=5,
                                      maxepoch=15)
    sol2 = res2["solution"]
    want2 = np.mean([obq.compute_posterior_parameters(sol2.reshape(1, -1), t[0], t[1], t[2:],
                                                      only_mean=True, w=w)["mean"][0]
                     for t in g["thetas"]])
    assert_close(res2["optimal_value"][0], want2, rtol=max(cond_rtol(g, s) for s in range(len(g["thetas"]))),
                 atol=1e-10, what="averaged posterior mean at the optimum")
    assert "mc_mean" in bq.optimal_solutions
    # GP version: full-dimensional points, task coordinate kept fixed
    np.random.seed(4)
    res3 = gp.optimize_posterior_mean(n_restarts=12, n_best_restarts=4, var_noise=th[0], mean=th[1],
                                      parameters_kernel=th[2:])
    s3 = res3["solution"]
    assert s3.shape == (g["points"].shape[1],)
    want3 = ogp.compute_posterior_parameters(s3.reshape(1, -1), th[0], th[1], th[2:],
                                             only_mean=True)["mean"]
    assert_close(res3["optimal_value"], np.asarray(want3).reshape(-1), rtol=cond_rtol(g, 0),
                 atol=1e-10, what="GP posterior mean at the optimum")
    if int(g["kind"]) == 2:
        assert float(s3[-1]) == float(int(s3[-1]))  # still a task id


def test_mle_parameters_and_serialisation_roundtrip(golden):
    """a16 + the JSON wire format: L-BFGS-B on the GPU log-likelihood reaches the optimum that
    the same optimiser finds on the oracle; serialize() -> json -> deserialize() reproduces the
    model."""
    import json
    from scipy.optimize import fmin_l_bfgs_b
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    g = golden("scaled_gamma")
    gp, _ = _models(g)
    spec, ogp, _ = oracle_from_golden(g)
    start = g["thetas"][1].copy()
    res = gp.mle_parameters(start=start)
    assert set(["solution", "optimal_value", "gradient", "warnflag", "task"]) <= set(res)
    assert res["optimal_value"] >= gp.log_likelihood(start[0], start[1], start[2:]) - 1e-9

    def f(x):
        try:
            return -ogp.log_likelihood(x[0], x[1], x[2:])
        except Exception:
            return np.inf

    ref = fmin_l_bfgs_b(f, start, fprime=lambda x: -np.clip(
        ogp.grad_log_likelihood(x[0], x[1], x[2:]), -1e300, 1e300), bounds=gp.get_bounds_parameters)
    assert_close(res["optimal_value"], -ref[1], rtol=1e-6, atol=1e-6, what="mle optimum")
    gp.fit_gp_regression(start=start)
    assert_close(gp.get_value_parameters_model, res["solution"], rtol=0, atol=0)
    # serialisation
    gp.samples_parameters = [t.copy() for t in g["thetas"]]
    wire = json.loads(json.dumps(gp.serialize()))
    assert set(wire) >= {"type_kernel", "training_data", "dimensions", "kernel_values", "mean_value",
                         "var_noise_value", "thinning", "n_burning", "max_steps_out", "data",
                         "bounds_domain", "type_bounds", "name_model", "training_name",
                         "problem_name", "same_correlation", "start_point_sampler",
                         "samples_parameters"}
    gp2 = GPFittingGaussian.deserialize(wire, noise=True, define_samplers=False)
    th = g["thetas"][0]
    assert gp2.log_likelihood(th[0], th[1], th[2:]) == gp.log_likelihood(th[0], th[1], th[2:])
    assert_close(np.array(gp2.samples_parameters), g["thetas"], rtol=0, atol=0)
    assert_close(gp2.get_value_parameters_model, gp.get_value_parameters_model, rtol=0, atol=0)


def test_add_points_extends_cached_factors(golden):
    """add_points_evaluations extends the factors that are asked for again after the update -- the
    model's own parameter values and the state of the slice-sampling chain -- in O(n^2) and drops
    every other cached batch; results equal the invalidate-and-refactorise path of the reference."""
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    g = golden("scaled_gamma")
    gp_inc, _ = _models(g)
    gp_ref, _ = _models(g)
    gp_ref.incremental_updates = False
    thetas = g["thetas"]
    rng = np.random.RandomState(9)
    for it in range(3):
        for gp in (gp_inc, gp_ref):
            gp.samples_parameters = [thetas[1].copy()]
            gp.log_likelihood_batch(thetas)          # a probe batch: not reused after the update
            for th in thetas:                        # single-theta batches (model values = thetas[0])
                gp.log_likelihood(th[0], th[1], th[2:])
        assert len(gp_inc._hb_cache) == 4
        p = np.concatenate([rng.uniform(0, 1, int(g["dx"])),
                            rng.gamma(2.0, 1.0, g["points"].shape[1] - int(g["dx"]))]).reshape(1, -1)
        v = np.array([np.sin(p.sum())])
        gp_inc.add_points_evaluations(p, v)
        gp_ref.add_points_evaluations(p, v)
        assert len(gp_inc._hb_cache) == 2 and len(gp_ref._hb_cache) == 0
        for hb in gp_inc._hb_cache.values():
            assert hb.S == 1 and hb.L is not None and hb.eng.n == len(gp_inc.data['evaluations'])
        kept = [hb.theta.cpu().numpy()[0] for hb in gp_inc._hb_cache.values()]
        assert sorted(map(tuple, kept)) == sorted(map(tuple, thetas[:2]))
        for th in thetas[:2]:
            n_cached = len(gp_inc._hb_cache)
            a = gp_inc.log_likelihood(th[0], th[1], th[2:])
            assert len(gp_inc._hb_cache) == n_cached  # served by the extended factor
            b = gp_ref.log_likelihood(th[0], th[1], th[2:])
            assert_close(a, b, rtol=1e-11, what="llh after %d appended points" % (it + 1))
        q = g["query"][0:2]
        th = thetas[0]
        pa = gp_inc.compute_posterior_parameters(q, th[0], th[1], th[2:])
        pb = gp_ref.compute_posterior_parameters(q, th[0], th[1], th[2:])
        assert_close(pa["mean"], pb["mean"], rtol=1e-9, atol=1e-12)
        assert_close(pa["cov"], pb["cov"], rtol=1e-7, atol=1e-11)


def test_acquisition_clean_cache_keeps_the_extended_factors(golden):
    """The BO loop (services/bgo.py) calls acquisition_function.clean_cache() right after
    add_points_evaluations: the quadrature caches go, the GP's device factors stay."""
    from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
    g = golden("scaled_gamma")
    gp, bq = _models(g)
    sbo = SBO(bq)
    th = gp.get_value_parameters_model
    gp.log_likelihood(th[0], th[1], th[2:])
    bq.max_mean[(th[0], th[1], tuple(th[2:]))] = 1.0
    p = np.concatenate([[0.3, 0.6], [1.0, 2.0]]).reshape(1, -1)
    gp.add_points_evaluations(p, np.array([0.5]))
    sbo.clean_cache()
    assert len(gp._hb_cache) == 1 and bq.max_mean == {} and len(bq._bq_cache) == 0
    hb = next(iter(gp._hb_cache.values()))
    assert hb.L is not None and hb.eng.n == len(gp.data['evaluations'])


def test_batched_slice_probes_give_the_sequential_chain():
    """speculative / paired probes evaluated in one batched factorisation
    leave the Markov chain (and the numpy random stream) exactly as the one-probe-at-a-time
    algorithm of the reference."""
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    from bayesian_quadrature_optimization_b200.lib import constant as K
    rng = np.random.RandomState(5)
    n = 40
    pts = rng.uniform(0, 1, (n, 3))
    y = np.sin(3 * pts.sum(1)) + 0.1 * rng.normal(0, 1, n)
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}

    def make(batched):
        GPFittingGaussian.batched_slice_probes = batched
        try:
            return GPFittingGaussian([K.SCALED_KERNEL, K.MATERN52_NAME], td, [3],
                                     bounds_domain=[[0, 1]] * 3, thinning=1, n_burning=2,
                                     max_steps_out=10, random_seed=3, noise=True)
        finally:
            GPFittingGaussian.batched_slice_probes = True

    calls = {"batch": 0, "single": 0}
    gp_b, gp_s = make(True), make(False)
    assert gp_b.slice_samplers[0].log_prob_batch is not None
    assert gp_s.slice_samplers[0].log_prob_batch is None
    orig_batch = gp_b.log_likelihood_batch
    orig_single = gp_s.log_likelihood

    def count_batch(th):
        calls["batch"] += 1
        return orig_batch(th)

    def count_single(*a):
        calls["single"] += 1
        return orig_single(*a)

    gp_b.log_likelihood_batch = count_batch
    gp_s.log_likelihood = count_single
    sb = gp_b.sample_parameters(4, random_seed=11)
    after_b = np.random.rand()
    ss = gp_s.sample_parameters(4, random_seed=11)
    after_s = np.random.rand()
    assert_close(np.array(sb), np.array(ss), rtol=0, atol=0, what="batched vs sequential chain")
    assert after_b == after_s  # the random stream is in the same state
    assert calls["batch"] > 0 and calls["single"] > 0


@pytest.mark.parametrize("name", CASES)
def test_quadrature_posterior_gradients_and_ei_on_g(golden, name):
    """a27 / a39 on the quadrature model: gradient of the posterior mean and variance of G against
    the oracle, EI(G) gradient against central differences of EI(G), and EI(bq).optimize /
    MultiTasks.optimize_first end to end."""
    from bayesian_quadrature_optimization_b200.acquisition_functions.ei import EI
    from bayesian_quadrature_optimization_b200.acquisition_functions.multi_task import MultiTasks
    from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import \
        BayesianQuadrature
    g = golden(name)
    gp, bq = _models(g)
    spec, ogp, obq = oracle_from_golden(g)
    w = g["wset"] if "wset" in g else None
    dx = int(g["dx"])
    th = g["thetas"][-1]
    x = g["xq"][0:1]
    got = bq.gradient_posterior_parameters(x, th[0], th[1], th[2:])
    want = obq.gradient_posterior_parameters(x, th[0], th[1], th[2:], w=w)
    rt = 10 * cond_rtol(g, len(g["thetas"]) - 1)
    assert_close(got["mean"], np.asarray(want["mean"]).reshape(-1), rtol=rt, atol=1e-11)
    assert_close(got["cov"], want["cov"], rtol=10 * rt, atol=1e-10)
    # EI on G: analytic gradient against central differences of the value.  With the reference's
    # two independent W draws the Monte-Carlo prior variance can fall below B Sigma^-1 B^T (the
    # reference's own q_var fixture is negative at this point), where EI clips the variance to 0
    # and its gradient is 0/0 in the reference as well; the formula is checked with the W set
    # against itself (a positive semi-definite estimate)
    if bq.w_sets is not None:
        bq.set_w_double_sets(np.stack([bq.w_sets[0], bq.w_sets[0]]))
    ei = EI(bq)
    gr = ei.evaluate_gradient(x, th[0], th[1], th[2:])
    assert gr.shape[0] >= dx
    scale = 100.0 if g["points"][:, 0].max() > 2 else 1.0
    for l in range(dx):
        h = 1e-5 * scale
        xp, xm = x.copy(), x.copy()
        xp[0, l] += h
        xm[0, l] -= h
        fd = (ei.evaluate(xp, th[0], th[1], th[2:])[0] - ei.evaluate(xm, th[0], th[1], th[2:])[0]) / (2 * h)
        assert_close(gr[l], fd, rtol=2e-4, atol=1e-9 + 1e-6 * abs(fd), what="dEI(G)/dx[%d]" % l)
    # Bayesian average: batched path equals the mean of the single-hyper-sample calls
    gp.samples_parameters = [t.copy() for t in g["thetas"]]
    vals, grads = ei.evaluate_bayesian(g["xq"], len(g["thetas"]), gradient=True)
    one = np.mean([[ei.evaluate(p.reshape(1, -1), t[0], t[1], t[2:])[0] for p in g["xq"]]
                   for t in g["thetas"]], axis=0)
    rt_all = max(cond_rtol(g, s) for s in range(len(g["thetas"])))
    assert_close(vals, one, rtol=100 * rt_all, atol=1e-13, what="Bayesian EI(G)")
    assert grads.shape == (len(g["xq"]), dx)
    if int(g["kind"]) == 2:
        # the x-only quadrature model that MultiTasks drives (sbo.py:1655-1661)
        bq2 = BayesianQuadrature(gp, list(range(dx)), bq.distribution,
                                 parameters_distribution=bq.parameters_distribution,
                                 model_only_x=True)
        mk = MultiTasks(bq2, int(g["n_tasks"]))
        res = mk.optimize(random_seed=1, n_restarts=6, n_best_restarts=3, n_samples_parameters=3,
                          start_new_chain=False, maxepoch=5)
        sol = res["solution"]
        assert sol.shape == (dx + 1,) and 0 <= int(sol[-1]) < int(g["n_tasks"])
        assert all(bq2.bounds[l][0] <= sol[l] <= bq2.bounds[l][-1] for l in range(dx))
        # the Adam climb does not end below the best of its own starting values
        assert np.isfinite(res["optimal_value"])


def test_one_sample_stochastic_gradient(golden):
    """a35: grad_voi_sgd / evaluate_gradient_given_sample_given_parameters return Z * grad_c b at the
    device's maximiser; checked against the oracle's gradient_vector_b at that maximiser."""
    from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
    g = golden("scaled_gamma")
    gp, bq = _models(g)
    spec, ogp, obq = oracle_from_golden(g)
    sbo = SBO(bq)
    th = g["thetas"][0]
    cand = g["cands"][0:1]
    np.random.seed(5)
    grad = sbo.evaluate_gradient_given_sample_given_parameters(cand, 0.7, th, n_restarts=3,
                                                               factr=1e12, maxiter=10)
    assert grad.shape == (1, g["points"].shape[1])
    xstar = _argmax_of(sbo, bq, cand, 0.7, th)
    want = obq.gradient_vector_b(cand, xstar, th[0], th[1], th[2:], w=g["wset"]) * 0.7
    assert_close(grad, want, rtol=1e-6, atol=1e-9, what="one-sample gradient")
    gp.samples_parameters = [t.copy() for t in g["thetas"]]
    np.random.seed(6)
    gv = sbo.grad_voi_sgd(cand[0], True, False, n_restarts=3, factr=1e12, maxiter=10)
    assert gv.shape == (g["points"].shape[1],) and np.all(np.isfinite(gv))


def _argmax_of(sbo, bq, cand, z, th):
    """The maximiser the device finds for (cand, z, theta) from sbo.starting_points_sbo."""
    eng = bq.engine_for(bq._theta(th[0], th[1], th[2:]))
    ub = eng.setup(np.asarray(cand, dtype=float).reshape(1, -1))
    _, arg, _ = eng.inner_maximize(ub, np.array([z]), np.asarray(sbo.starting_points_sbo),
                                   maxiter=10, factr=1e12, pgtol=sbo.inner_pgtol,
                                   max_ls=sbo.inner_max_ls, use_vertices=True)
    return arg.cpu().numpy()[0, 0].reshape(1, -1)


def test_independent_chains_batched_equal_sequential_chains():
    """sample_parameters_chains: chains advanced side by side with merged probes draw exactly what
    the same chains draw one after the other (each with its own RandomState)."""
    import copy
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    from bayesian_quadrature_optimization_b200.lib import constant as K
    rng = np.random.RandomState(5)
    n = 40
    pts = rng.uniform(0, 1, (n, 3))
    y = np.sin(3 * pts.sum(1)) + 0.1 * rng.normal(0, 1, n)
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}
    gp = GPFittingGaussian([K.SCALED_KERNEL, K.MATERN52_NAME], td, [3], bounds_domain=[[0, 1]] * 3,
                           thinning=1, n_burning=0, max_steps_out=10, random_seed=3, noise=True)
    calls = {"n": 0, "vectors": 0}
    orig = gp.log_likelihood_batch

    def counting(th):
        calls["n"] += 1
        calls["vectors"] += len(th)
        return orig(th)

    gp.log_likelihood_batch = counting
    start = gp.samples_parameters[-1].copy()
    n_before = len(gp.samples_parameters)
    got = gp.sample_parameters_chains(2, n_chains=3, random_seed=100, burn_in_sweeps=1)
    assert len(got) == 6 and len(gp.samples_parameters) == n_before + 6
    assert calls["vectors"] > calls["n"]  # probes of different chains shared factorisations
    # the same three chains, one after the other
    want = []
    for cid in range(3):
        chain_rng = np.random.RandomState(100 + cid)
        samplers = []
        for sp in gp.slice_samplers:
            sc = copy.copy(sp)
            sc.rng = chain_rng
            samplers.append(sc)
        point = gp._slice_sweeps(start, 1, samplers)[-1]
        want += gp._slice_sweeps(point, 2 * (gp.thinning + 1), samplers)[::gp.thinning + 1]
    assert_close(np.array(got), np.array(want), rtol=0, atol=0, what="batched independent chains")


def test_model_with_several_chains_samples_and_runs_bgo():
    """n_chains > 1 routes sample_parameters through the side-by-side chains: right count, valid
    parameters, reproducible under np.random.seed, and bgo() accepts the option."""
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    from bayesian_quadrature_optimization_b200.services.bgo import bgo
    from bayesian_quadrature_optimization_b200.lib import constant as K
    rng = np.random.RandomState(8)
    pts = rng.uniform(0, 1, (30, 2))
    y = np.cos(4 * pts[:, 0]) + pts[:, 1]
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}

    def draw():
        gp = GPFittingGaussian([K.SCALED_KERNEL, K.MATERN52_NAME], td, [2],
                               bounds_domain=[[0, 1]] * 2, thinning=1, n_burning=2, max_steps_out=10,
                               random_seed=3, noise=True, n_chains=3)
        np.random.seed(12)
        return gp, gp.sample_parameters(5)

    gp, s1 = draw()
    _, s2 = draw()
    assert len(s1) == 5 and len(gp.samples_parameters) >= 5
    assert_close(np.array(s1), np.array(s2), rtol=0, atol=0, what="seeded multi-chain draw")
    for v in s1:
        assert v[0] > 0 and np.all(v[2:] > 0) and np.isfinite(gp.log_prob_parameters(v))

    def f(x):
        return [-(x[0] - 0.3) ** 2]

    def g(x):
        return [-(x[0] - 0.3) ** 2 + 0.1 * (x[1] - 0.5)]

    out = bgo(f, [(0.0, 1.0)], integrand_function=g, bounds_domain_w=[[0, 1]], type_bounds=[0, 1],
              n_iterations=1, n_training=4, random_seed=2, n_samples_parameters=2, n_restarts=2,
              n_restarts_mc=1, n_samples_mc=2, maxepoch=2, thinning=1, n_burning=2, max_steps_out=5,
              n_restarts_mean=4, n_best_restarts_mean=2, n_samples_parameters_mean=2, maxepoch_mean=3,
              default_n_samples_parameters=3, default_n_samples=3, n_chains=2)
    assert out["optimal_solution"].shape == (1,) and 0.0 <= out["optimal_solution"][0] <= 1.0


@pytest.mark.parametrize("name", ["product_uniform", "product_weighted"])
def test_sbo_optimize_discretised_branch(golden, name):
    """SBO.optimize(monte_carlo=False, n_samples_parameters=0) (sbo/acquisition_functions/sbo.py:
    1800-1821): L-BFGS-B (maxiter 10) on the discretised knowledge gradient from every start.  The
    same scipy call on the CPU oracle from the same start points gives the same optimum."""
    from scipy.optimize import fmin_l_bfgs_b
    from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
    g = golden(name)
    gp, bq = _models(g)
    spec, ogp, obq = oracle_from_golden(g)
    disc = g["discretization"]
    sbo = SBO(bq, disc)
    th = gp.get_value_parameters_model
    np.random.seed(3)
    starts = sbo._restart_points(2)
    np.random.seed(3)
    res = sbo.optimize(monte_carlo=False, n_samples_parameters=0, n_restarts=2, start_ei=False)
    assert set(res) >= {"solution", "optimal_value", "gradient"}
    bounds_x = [(disc[:, l].min(), disc[:, l].max()) for l in range(disc.shape[1])]
    osbo = orc.OracleSBO(obq, bounds_x, discretization=disc)
    bounds = [tuple(b) for b in sbo.bounds_opt]
    best = -np.inf
    for x0 in starts:
        opt = fmin_l_bfgs_b(lambda x: -osbo.evaluate(x.reshape(1, -1), th[0], th[1], th[2:]), x0,
                            fprime=lambda x: -osbo.evaluate_gradient(x.reshape(1, -1), th[0], th[1], th[2:]),
                            bounds=bounds, maxiter=10)
        best = max(best, -opt[1])
    assert_close(res["optimal_value"], best, rtol=1e-5, atol=1e-9, what="discretised SBO optimum")
    # the returned value is the discretised SBO at the returned solution
    assert_close(sbo.evaluate(res["solution"].reshape(1, -1)), res["optimal_value"], rtol=1e-12)


def test_sbo_optimize_resident_samples_against_advancing_chain():
    """VERDICT r1 weak point 12: the SGD of SBO.optimize draws its hyper-samples from the resident
    posterior samples; advance_chain=True is the reference's source (one new slice sample per
    gradient, sbo.py:1129).  Same seeds and starts, one common scorer: the default must not be worse
    (profiles/r02_sgd_hyper_source.md has the 32-seed run)."""
    import sys
    sys.path.insert(0, "tests/microbench")
    import sgd_hyper_source as m
    td, T = m.problem()
    gp, bq, sbo = m.model(td, T, 11)
    gp.sample_parameters(8, random_seed=5)
    calls = []
    orig = gp.sample_parameters

    def counting(n_samples, *a, **k):
        calls.append(n_samples)
        return orig(n_samples, *a, **k)

    gp.sample_parameters = counting
    res = sbo.optimize(random_seed=3, monte_carlo=True, n_samples=2, n_restarts_mc=3, n_restarts=4,
                       n_samples_parameters=2, start_new_chain=False, maxepoch=4,
                       default_n_samples=6, default_n_samples_parameters=8, start_ei=False,
                       factr=1e12, maxiter=10, advance_chain=True)
    assert 1 <= len(calls) <= 4 and all(c == 2 for c in calls)   # two new samples per epoch run
    x = res["solution"]
    assert 0.0 <= x[0] <= 1.0 and 0.0 <= x[1] <= 1.0 and int(x[2]) in range(T)
    out = m.main(seeds=tuple(range(1, 9)), maxepoch=10)
    a = np.array([r["resident"]["score"] for r in out["rows"]])
    b = np.array([r["chain"]["score"] for r in out["rows"]])
    for r in out["rows"]:
        for k in ("resident", "chain"):
            xs = r[k]["x"]
            assert 0.0 <= xs[0] <= 1.0 and 0.0 <= xs[1] <= 1.0 and int(xs[2]) in range(T)
    se = (a - b).std(ddof=1) / np.sqrt(len(a))
    assert a.mean() >= b.mean() - 2.0 * se, (a.mean(), b.mean(), se)


def test_bgo_loop_converges_on_a_toy_problem():
    """Four iterations of the whole BO loop (data append, factor extension, chain advance,
    acquisition maximisation, posterior-mean optimum) for BQO and for EI: the reported optimum lands
    near the true one (tests/microbench/bgo_loop_check.py prints the same runs)."""
    from bayesian_quadrature_optimization_b200.services.bgo import bgo

    def g(x):
        return [-(x[0] - 0.3) ** 2 - 0.5 * (x[1] - 0.6) ** 2]

    def f(xw):
        return [-(xw[0] - 0.3) ** 2 - 0.5 * (xw[1] - 0.6) ** 2 + (0.2 if int(xw[2]) == 0 else -0.2)]

    sol = bgo(g, [(0.0, 1.0), (0.0, 1.0)], integrand_function=f, bounds_domain_w=[[0, 1]],
              type_bounds=[0, 0, 1], name_method='bqo', n_iterations=4, random_seed=5, n_training=12,
              thinning=2, n_burning=20, max_steps_out=20, n_restarts=4, n_samples_parameters=3,
              maxepoch=5, n_samples_mc=3, n_restarts_mc=3, default_n_samples=6,
              default_n_samples_parameters=5, n_restarts_mean=20)
    assert np.max(np.abs(sol["optimal_solution"] - np.array([0.3, 0.6]))) < 0.1
    assert abs(sol["optimal_value"]) < 0.02
    sol = bgo(g, [(0.0, 1.0), (0.0, 1.0)], name_method='ei', n_iterations=4, random_seed=5,
              n_training=8, thinning=2, n_burning=20, max_steps_out=20, n_restarts=4,
              n_samples_parameters=3, maxepoch=5, n_restarts_mean=20)
    assert np.max(np.abs(sol["optimal_solution"] - np.array([0.3, 0.6]))) < 0.1
