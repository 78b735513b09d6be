"""CPU: the oracle restatement against (1) the closed-form known answers of the reference's own
test-suite and (2) the fixtures generated from the unmodified reference
(tests/golden/make_golden.py).  Tolerance 1e-12 relative unless stated: both sides are
float64 numpy and differ only in summation order."""
import numpy as np
import pytest

from oracle import bqo_oracle as orc
from tests.helpers import assert_close, oracle_from_golden

CASES = ["product_uniform", "product_weighted", "scaled_gamma"]
GP_CASES = CASES + ["matern52_plain"]  # the last one has no quadrature part


def test_matern_known_answers():
    # reference tests/kernels/test_matern52.py:60-94 (ScaledKernel(Matern52(ls=[2,3]), sigma2=4))
    spec = orc.KernelSpec(orc.SCALED_MATERN52, 2)
    p = np.array([2.0, 3.0, 4.0])
    assert spec.cross_cov(p, [[2.0, 4.0]], [[3.0, 5.0]])[0, 0] == 3.0737065834936015
    got = spec.cross_cov(p, [[2.0, 4.0], [3.0, 5.0]], [[1.5, 9.0], [-3.0, 8.0]])
    want = np.array([[0.87752659905500319, 0.14684671522649542],
                     [1.0880320585678382, 0.084041575076539962]])
    assert_close(got, want, rtol=1e-15, what="matern golden")


def test_tasks_known_answers():
    # reference tests/kernels/test_tasks_kernel.py:52-71
    covM, L = orc.tasks_cov_matrix(np.array([3.0]), 1)
    assert L[0, 0] == np.exp(3.0) and covM[0, 0] == np.exp(3.0) * np.exp(3.0)
    covM, L = orc.tasks_cov_matrix(np.array([0.0, 0.0]), 2, same_correlation=True)
    assert np.all(covM == np.array([[2.0, 1.0], [1.0, 2.0]]))
    covM, _ = orc.tasks_cov_matrix(np.array([0.0]), 1, same_correlation=True)
    assert covM[0, 0] == 1.0
    # tests/kernels/test_tasks_kernel.py:128-131
    grad = orc.tasks_grad_base(np.array([[np.exp(1.0)]]), 1)
    assert_close(grad[0], np.array([[2.0 * np.exp(2.0)]]), rtol=1e-15)


def test_loglik_known_answer():
    # reference tests/models/test_gp_fitting_gaussian.py:261-276: one point, Matern52, var 1
    spec = orc.KernelSpec(orc.MATERN52, 1)
    gp = orc.OracleGP(spec, [[5.0]], [5.0])
    chol, cov = gp.chol_cov_including_noise(1.0, np.array([1.0]))
    assert cov[0, 0] == 2.0 and chol[0, 0] == np.sqrt(2.0)
    # -log(sqrt(2)) - 0.5 * 1 * 1/2 ... with mean 4: (y - mean) = 1
    llh = gp.log_likelihood(1.0, 4.0, np.array([1.0]))
    assert abs(llh - (-0.5 * np.log(2.0) - 0.25)) < 1e-15
    # :274-276 complex_gp: Product(Matern52, Tasks(T=1)), one point, observation noise 0.5
    spec = orc.KernelSpec(orc.PRODUCT_TASKS, 2, n_tasks=1)
    gp = orc.OracleGP(spec, [[42.2851784656, 0]], [1.0], [0.5])
    chol, cov = gp.chol_cov_including_noise(1.0, np.array([1.0, 0.0]))
    assert cov[0, 0] == 2.5 and chol[0, 0] == np.sqrt(2.5)
    assert gp.log_likelihood(1.0, 1.0, np.array([1.0, 0.0])) == -0.45814536593707761


def test_product_cross_cov_known_answer():
    # reference tests/models/test_gp_fitting_gaussian.py:680-684
    spec = orc.KernelSpec(orc.PRODUCT_TASKS, 2, n_tasks=1)
    got = spec.cross_cov(np.array([1.0, 0.0]), np.array([[2.0, 0.0]]), np.array([[1.0, 0.0]]))
    assert got[0, 0] == 0.52399410883182029


def test_hvoi_degenerate():
    # reference tests/acquisition_functions/test_sbo.py:219-224
    assert orc.hvoi(np.array([1.0]), np.array([0.0, 1.0]), np.array([0])) == 0


@pytest.mark.parametrize("name", GP_CASES)
def test_gp_against_reference(golden, name):
    g = golden(name)
    spec, gp, _ = oracle_from_golden(g)
    for s, th in enumerate(g["thetas"]):
        vn, mu, pk = th[0], th[1], th[2:]
        assert_close(spec.cov(pk, g["points"]), g["cov"][s], rtol=1e-13, what="cov")
        gc = np.stack(spec.grad_params(pk, g["points"]))
        assert_close(gc, g["grad_cov"][s], rtol=1e-11, atol=1e-15, what="grad_cov")
        gp.clean_cache()
        chol, _ = gp.chol_cov_including_noise(vn, pk)
        assert_close(chol, g["chol"][s], rtol=1e-12, atol=1e-15, what="chol")
        assert_close(gp.log_likelihood(vn, mu, pk), g["llh"][s], rtol=1e-12, what="llh")
        assert_close(gp.grad_log_likelihood(vn, mu, pk), g["grad_llh"][s], rtol=1e-9, atol=1e-10,
                     what="grad_llh")
        post = gp.compute_posterior_parameters(g["query"], vn, mu, pk)
        assert_close(post["mean"], g["post_mean"][s], rtol=1e-11, what="post mean")
        assert_close(post["cov"], g["post_cov"][s], rtol=1e-9, atol=1e-12, what="post cov")
        for q in range(g["query"].shape[0]):
            pt = g["query"][q:q + 1, :]
            gr = gp.gradient_posterior_parameters(pt, vn, mu, pk)
            assert_close(gr["mean"], g["grad_post_mean"][s, q], rtol=1e-10, atol=1e-13)
            assert_close(np.asarray(gr["cov"]).reshape(-1), g["grad_post_cov"][s, q], rtol=1e-9,
                         atol=1e-12)
            assert_close(orc.ei_evaluate(gp, pt, vn, mu, pk)[0], g["ei"][s, q], rtol=1e-9,
                         atol=1e-300)
            assert_close(orc.ei_evaluate_gradient(gp, pt, vn, mu, pk), g["grad_ei"][s, q],
                         rtol=1e-8, atol=1e-300)


@pytest.mark.parametrize("name", CASES)
def test_bq_against_reference(golden, name):
    g = golden(name)
    spec, gp, bq = oracle_from_golden(g)
    w = g["wset"] if "wset" in g else None
    xq, cands = g["xq"], g["cands"]
    for s, th in enumerate(g["thetas"]):
        vn, mu, pk = th[0], th[1], th[2:]
        for p in range(xq.shape[0]):
            x = xq[p:p + 1, :]
            assert_close(bq.evaluate_quadrature_cross_cov(x, g["points"], pk, w=w), g["B"][s, p],
                         rtol=1e-12, what="B")
            assert_close(bq.evaluate_grad_quadrature_cross_cov(x, g["points"], pk, w=w),
                         g["grad_B"][s, p], rtol=1e-11, atol=1e-16, what="grad B")
            wc = (g["w100"][0], g["w100"][1]) if "w100" in g else None
            post = bq.compute_posterior_parameters(x, vn, mu, pk, w=w, w_cov=wc)
            assert_close(post["mean"][0], g["q_mean"][s, p], rtol=1e-11, what="q mean")
            assert_close(post["cov"], g["q_var"][s, p], rtol=1e-8, atol=1e-12, what="q var")
            assert_close(bq.gradient_posterior_mean(x, vn, mu, pk, w=w), g["grad_q_mean"][s, p],
                         rtol=1e-10, atol=1e-13)
        for c in range(cands.shape[0]):
            cand = cands[c:c + 1, :]
            par = bq.get_parameters_for_samples(cand, vn, mu, pk)
            assert_close(par["gamma"][:, 0], g["gamma"][s, c], rtol=1e-13)
            assert_close(par["solve_2"][:, 0], g["solve_2"][s, c], rtol=1e-9, atol=1e-13)
            assert_close(par["denominator"][0], g["den"][s, c], rtol=1e-10)
            for p in range(xq.shape[0]):
                x = xq[p:p + 1, :]
                v = bq.compute_parameters_for_sample(x, cand, vn, mu, pk, w=w)
                gr = bq.compute_gradient_parameters_for_sample(x, cand, vn, mu, pk, w=w)
                got = np.concatenate([[v["a"][0]], [np.asarray(v["b"]).reshape(-1)[0]], gr["a"],
                                      gr["b"]])
                assert_close(got, g["ab"][s, c, p], rtol=1e-9, atol=1e-12, what="ab")
            gvb = bq.gradient_vector_b(cand, xq, vn, mu, pk, w=w)
            assert_close(gvb, g["gvb"][s, c], rtol=1e-8, atol=1e-12, what="gvb")


@pytest.mark.parametrize("name", ["product_uniform", "product_weighted"])
def test_discretised_sbo_against_reference(golden, name):
    g = golden(name)
    spec, gp, bq = oracle_from_golden(g)
    bounds = [(g["discretization"][:, l].min(), g["discretization"][:, l].max())
              for l in range(g["discretization"].shape[1])]
    sbo = orc.OracleSBO(bq, bounds, discretization=g["discretization"])
    for s, th in enumerate(g["thetas"]):
        vn, mu, pk = th[0], th[1], th[2:]
        for c in range(g["cands"].shape[0]):
            cand = g["cands"][c:c + 1, :]
            assert_close(sbo.evaluate(cand, vn, mu, pk), g["kg"][s, c], rtol=1e-9, atol=1e-14,
                         what="kg")
            assert_close(sbo.evaluate_gradient(cand, vn, mu, pk), g["grad_kg"][s, c], rtol=1e-8,
                         atol=1e-13, what="grad kg")
