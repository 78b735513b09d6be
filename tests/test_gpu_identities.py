"""Exact identities the reference's tests check by Monte-Carlo simulation (style 3:
tests/acquisition_functions/test_sbo.py:152-188 simulated fantasies against the KG value,
:267-292 sample moments of a + Z b), checked here deterministically on the CUDA path:

* observing y_c = mu(c) + Z * den at the candidate c moves the posterior mean of G(x) to
  a(x) + Z b(x): the a, b of stage C against the posterior of a SECOND model that really holds the
  extra point (a different code path: new factorisation, new alpha, plain quadrature posterior);
* the closed-form discretised knowledge gradient equals E_Z max_i (a_i + Z b_i) - max_i a_i, with
  the expectation taken by quadrature on the host.
"""
import numpy as np
import pytest

from tests.helpers import assert_close, cond_rtol
from tests.test_gpu_facade import CASES, _models

pytestmark = pytest.mark.gpu


def _best_conditioned(g):
    tols = [cond_rtol(g, s) for s in range(len(g["thetas"]))]
    return int(np.argmin(tols)), float(np.min(tols))


@pytest.mark.parametrize("name", CASES)
def test_fantasy_observation_moves_the_mean_of_g_along_b(golden, name):
    from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
    g = {k: golden(name)[k] for k in golden(name).keys()}
    gp, bq = _models(g)
    sbo = SBO(bq)
    s, tol = _best_conditioned(g)
    th = g["thetas"][s]
    vn, mu, pk = th[0], th[1], th[2:]
    for c in range(min(2, g["cands"].shape[0])):
        cand = g["cands"][c:c + 1, :]
        par = bq.get_parameters_for_samples(True, cand, pk, vn, mu)
        den = float(par["denominator"][0])
        assert den > 0.0
        mu_c = float(gp.compute_posterior_parameters(cand, vn, mu, pk, only_mean=True)["mean"][0])
        for z in (-1.3, 0.45):
            g2 = dict(g)
            g2["points"] = np.vstack([g["points"], cand])
            g2["evaluations"] = np.append(g["evaluations"], mu_c + z * den)
            if "var_noise_obs" in g:
                # the denominator's noise term is the mean observed noise (bayesian_quadrature.py:1122-1127)
                g2["var_noise_obs"] = np.append(g["var_noise_obs"], np.mean(g["var_noise_obs"]))
            _, bq2 = _models(g2)
            for p in range(g["xq"].shape[0]):
                x = g["xq"][p:p + 1, :]
                want = bq2.compute_posterior_parameters(x, vn, mu, pk, only_mean=True)["mean"][0]
                got = sbo.evaluate_sample(x, cand, z, vn, mu, pk)
                assert_close(got, want, rtol=max(1e-8, 10 * tol), atol=1e-10,
                             what="a + Z b against the posterior mean with the fantasised point")


@pytest.mark.parametrize("name", ["product_uniform", "product_weighted"])
def test_discretised_kg_is_the_expected_maximum_over_the_grid(golden, name):
    from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
    g = golden(name)
    gp, bq = _models(g)
    disc = g["discretization"]
    sbo = SBO(bq, disc)
    s, tol = _best_conditioned(g)
    th = g["thetas"][s]
    vn, mu, pk = th[0], th[1], th[2:]
    z = np.linspace(-9.0, 9.0, 180001)
    w = np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi)
    for c in range(g["cands"].shape[0]):
        cand = g["cands"][c:c + 1, :]
        vec = bq.compute_posterior_parameters_kg(disc, cand, var_noise=vn, mean=mu,
                                                 parameters_kernel=pk)
        a, b = np.asarray(vec["a"]).reshape(-1), np.asarray(vec["b"]).reshape(-1)
        best = np.full(z.shape, -np.inf)
        for ai, bi in zip(a, b):
            np.maximum(best, ai + bi * z, out=best)
        f = w * best
        expect = float(np.sum(0.5 * (f[1:] + f[:-1]) * np.diff(z))) - float(np.max(a))
        got = sbo.evaluate(cand, vn, mu, pk)
        assert_close(got, expect, rtol=1e-6, atol=1e-9 * max(1.0, float(np.max(np.abs(a)))),
                     what="closed-form KG against quadrature over Z")


@pytest.mark.parametrize("name", CASES)
def test_ei_is_the_expected_improvement_over_the_historical_best(golden, name):
    """tests/acquisition_functions/test_ei.py:88-110 samples the improvement; here
    E max(mu + sigma Z - best, 0) by quadrature against the device EI (GP model)."""
    from bayesian_quadrature_optimization_b200.acquisition_functions.ei import EI
    g = golden(name)
    gp, _ = _models(g)
    ei = EI(gp)
    s, tol = _best_conditioned(g)
    th = g["thetas"][s]
    vn, mu, pk = th[0], th[1], th[2:]
    best = gp.get_historical_best_solution(vn, mu, pk, False)
    z = np.linspace(-9.0, 9.0, 180001)
    w = np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi)
    for q in range(g["query"].shape[0]):
        pt = g["query"][q:q + 1, :]
        post = gp.compute_posterior_parameters(pt, vn, mu, pk)
        m, sd = float(post["mean"][0]), float(np.sqrt(max(post["cov"][0, 0], 0.0)))
        f = w * np.maximum(m + sd * z - best, 0.0)
        expect = float(np.sum(0.5 * (f[1:] + f[:-1]) * np.diff(z)))
        assert_close(ei.evaluate(pt, vn, mu, pk)[0], expect, rtol=max(1e-6, 10 * tol), atol=1e-12,
                     what="EI against quadrature over Z")
