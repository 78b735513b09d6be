"""Hessian rows a5 / a24 / a31 (sbo/kernels/matern52.py:466-503,
sbo/numerical_tools/bayesian_quadrature.py:364-386,1241-1286, sbo/acquisition_functions/sbo.py:196-235)
against fixtures produced by the unmodified reference (tests/golden/make_golden_hessian.py).

The reference evaluates hess(r) k'(r) + grad r grad r^T k''(r), whose two terms are O(1/r) and cancel;
its own rounding error is therefore about eps / r relative to the entries.  Tolerances: 1e-9 of the
largest entry of each Hessian (1e-8 through Sigma^-1)."""
import numpy as np
import pytest

from oracle import bqo_oracle as orc
from tests.helpers import assert_close, oracle_from_golden, cond_rtol

CASES = ["product_weighted", "scaled_gamma"]


def _close(got, want, rtol, what):
    scale = np.max(np.abs(want))
    assert_close(got, want, rtol=rtol, atol=rtol * scale, what=what)


@pytest.mark.parametrize("name", CASES)
def test_oracle_hessians_against_reference(golden, name):
    g, h = golden(name), golden("hessian_" + name)
    spec, gp, bq = oracle_from_golden(g)
    w = g["wset"] if "wset" in g else None
    for s, th in enumerate(g["thetas"]):
        vn, mu, pk = th[0], th[1], th[2:]
        tol = max(1e-8, cond_rtol(g, s))
        for q in range(g["query"].shape[0]):
            _close(spec.hessian_respect_point(pk, g["query"][q:q + 1], g["points"]), h["hess_k"][s, q],
                   1e-9, "hess k")
        for p in range(g["xq"].shape[0]):
            _close(bq.evaluate_hessian_cross_cov(g["xq"][p:p + 1], g["points"], pk, w=w),
                   h["hess_B"][s, p], 1e-9, "hess B")
        for c in range(g["cands"].shape[0]):
            for p in range(g["xq"].shape[0]):
                hh = bq.compute_hessian_parameters_for_sample(g["xq"][p:p + 1], g["cands"][c:c + 1],
                                                              vn, mu, pk, w=w)
                _close(hh["a"], h["hess_a"][s, c, p], tol, "hess a")
                _close(hh["b"], h["hess_b"][s, c, p], tol, "hess b")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_hessians_against_reference(golden, name):
    from tests.test_gpu_facade import _models
    from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
    g, h = golden(name), golden("hessian_" + name)
    gp, bq = _models(g)
    sbo = SBO(bq, g["discretization"] if "discretization" in g else None)
    for s, th in enumerate(g["thetas"]):
        vn, mu, pk = th[0], th[1], th[2:]
        tol = max(1e-8, cond_rtol(g, s))
        for q in range(g["query"].shape[0]):
            got = gp.evaluate_hessian_cross_cov_respect_point(g["query"][q:q + 1], g["points"], pk)
            _close(got, h["hess_k"][s, q], 1e-9, "hess k")
        for p in range(g["xq"].shape[0]):
            got = bq.evaluate_hessian_cross_cov(g["xq"][p:p + 1], g["points"], pk)
            _close(got, h["hess_B"][s, p], 1e-9, "hess B")
        for c in range(g["cands"].shape[0]):
            for p in range(g["xq"].shape[0]):
                hh = bq.compute_hessian_parameters_for_sample(g["xq"][p:p + 1], g["cands"][c:c + 1],
                                                              vn, mu, pk)
                _close(hh["a"], h["hess_a"][s, c, p], tol, "hess a")
                _close(hh["b"], h["hess_b"][s, c, p], tol, "hess b")
                for iz, z in enumerate(g["zs"]):
                    got = sbo.evaluate_hessian_sample(g["xq"][p:p + 1], g["cands"][c:c + 1], z, vn, mu, pk)
                    _close(got, h["hess_sample"][s, c, p, iz], tol, "hess a + Z b")


@pytest.mark.gpu
def test_gpu_hessian_is_the_derivative_of_the_gradient():
    """Central differences of grad a, grad b (tests/numerical_tools/test_bayesian_quadrature.py:436-476
    checks the reference's Hessians the same way) and the value at r = 0, where the reference's
    1/r form is NaN and the closed form is g(0) / ls^2 on the diagonal."""
    from tests.test_gpu_facade import _models
    from tests.conftest import GOLDEN_DIR
    import os
    with np.load(os.path.join(GOLDEN_DIR, "scaled_gamma.npz")) as f:
        g = {k: f[k] for k in f.files}
    gp, bq = _models(g)
    th = g["thetas"][0]
    x, cand = g["xq"][0:1].copy(), g["cands"][0:1]
    hh = bq.compute_hessian_parameters_for_sample(x, cand, th[0], th[1], th[2:])
    for l in range(x.shape[1]):
        e = 1e-5
        xp, xm = x.copy(), x.copy()
        xp[0, l] += e
        xm[0, l] -= e
        gp_ = bq.compute_gradient_parameters_for_sample(xp, cand, th[0], th[1], th[2:])
        gm_ = bq.compute_gradient_parameters_for_sample(xm, cand, th[0], th[1], th[2:])
        assert_close(hh["a"][l], (gp_["a"] - gm_["a"]) / (2 * e), rtol=1e-6, atol=1e-6, what="hess a fd")
        assert_close(hh["b"][l], (gp_["b"] - gm_["b"]) / (2 * e), rtol=1e-6, atol=1e-6, what="hess b fd")
    pk = th[2:]
    H0 = gp.evaluate_hessian_cross_cov_respect_point(g["points"][3:4], g["points"][3:5], pk)
    want = np.diag(-(5.0 / 3.0) * pk[-1] / pk[:-1] ** 2)
    assert_close(H0[0], want, rtol=1e-13, atol=0, what="Hessian at r = 0")
    assert np.all(np.isfinite(H0))
