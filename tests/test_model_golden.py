"""Model-level parity against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden_model.py): data-driven defaults, priors, log_prob_parameters, the seeded
slice-sampling chain, historical best, gamma prior variance and serialize().

CPU tests (`not gpu`) pin the host logic: defaults, bounds, priors, serialisation, and the slice
sampler driven by the oracle's log-likelihood (the reference's RNG call order and accept/reject
decisions).  GPU tests repeat log_prob_parameters and the chain with the CUDA likelihood and pin
the device-side rows (a14, a15, a20, a22, f1, f4 of SURVEY 8).
"""
import json
import os

import numpy as np
import pytest

from tests.helpers import assert_close, spec_from_golden
from tests.conftest import GOLDEN_DIR
from oracle import bqo_oracle as orc

from bayesian_quadrature_optimization_b200.lib.constant import (
    MATERN52_NAME, TASKS_KERNEL_NAME, PRODUCT_KERNELS_SEPARABLE, SCALED_KERNEL, SAME_CORRELATION,
    UNIFORM_FINITE, WEIGHTED_UNIFORM_FINITE, GAMMA, TASKS)
from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian

CASES = ["product_uniform", "product_weighted", "product_noise", "scaled_gamma", "matern52_plain"]


def load_model(name):
    with np.load(os.path.join(GOLDEN_DIR, "model_%s.npz" % name)) as f:
        g = {k: f[k] for k in f.files}
    with open(os.path.join(GOLDEN_DIR, "model_%s.json" % name)) as f:
        ser = json.load(f)
    return g, ser


def build_model(name, g, ser, **extra):
    """The constructor call of make_golden_model.py, on this repo's class."""
    td = ser["training_data"]
    kw = dict(thinning=int(g["thinning"]), max_steps_out=int(g["max_steps_out"]),
              noise=bool(int(g["noise"])))
    kw.update(extra)
    if name == "product_uniform":
        return GPFittingGaussian([PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME], td,
                                 [2, 1, 2], bounds_domain=[[0, 100], [0, 1]], type_bounds=[0, 1], **kw)
    if name == "product_weighted":
        return GPFittingGaussian([PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME], td,
                                 [3, 2, 4], bounds_domain=[[0, 1], [0, 1], list(range(4))],
                                 type_bounds=[0, 0, 1], **kw, **{SAME_CORRELATION: True})
    if name == "product_noise":
        return GPFittingGaussian([PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME], td,
                                 [3, 2, 3], bounds_domain=[[-1, 1], [-1, 1], list(range(3))],
                                 type_bounds=[0, 0, 1], **kw, **{SAME_CORRELATION: True})
    if name == "scaled_gamma":
        return GPFittingGaussian([SCALED_KERNEL, MATERN52_NAME], td, [4],
                                 bounds_domain=[[0, 1], [0, 1], [0, 10], [0, 10]], **kw)
    if name == "matern52_plain":
        return GPFittingGaussian([MATERN52_NAME], td, [3], bounds_domain=[[-1, 2]] * 3, **kw)
    raise KeyError(name)


def lp_rtol(g, theta, base=1e-9):
    """log-probabilities pass through the Cholesky factor of Sigma(theta): two correct
    factorisations differ by about cond(Sigma) * eps (tests/helpers.py:cond_rtol); the
    data-driven default length scale of the n = 30 toy (106) makes Sigma nearly singular."""
    spec = spec_from_golden(g)
    theta = np.asarray(theta, dtype=float)
    cov = spec.cov(theta[2:], g["points"]) + theta[0] * np.eye(len(g["evaluations"]))
    if "var_noise_obs" in g:
        cov = cov + np.diag(g["var_noise_obs"])
    return max(base, 8.0 * np.linalg.cond(cov) * np.finfo(float).eps)


def enc_bounds(bounds):
    out = np.full((len(bounds), 2), np.nan)
    for i, b in enumerate(bounds):
        for j in (0, 1):
            if b[j] is not None:
                out[i, j] = b[j]
    return out


def use_oracle_likelihood(gp, g):
    """Replace the two likelihood entry points of the model by the CPU oracle (host-logic tests
    only: the product itself has no CPU path)."""
    spec = spec_from_golden(g)
    vno = g["var_noise_obs"] if "var_noise_obs" in g else None
    ogp = orc.OracleGP(spec, g["points"], g["evaluations"], vno)

    def llh(var_noise, mean, params):
        ogp.clean_cache()
        return ogp.log_likelihood(var_noise, mean, np.asarray(params, dtype=float))

    def llh_batch(thetas):
        out = []
        for t in np.asarray(thetas, dtype=float):
            try:
                out.append(llh(t[0], t[1], t[2:]))
            except Exception:
                out.append(-np.inf)
        return np.array(out)

    gp.log_likelihood = llh
    gp.log_likelihood_batch = llh_batch
    return ogp


# ---------------------------------------------------------------------------------------------------
# CPU: host logic
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_defaults_bounds_and_indexes_match_the_reference(name):
    g, ser = load_model(name)
    gp = build_model(name, g, ser)
    assert_close(gp.kernel_values, g["kernel_values"], rtol=1e-12, what="kernel_values")
    assert_close(gp.mean_value, g["mean_value"], rtol=1e-13, what="mean_value")
    assert_close(gp.var_noise_value, g["var_noise_value"], rtol=1e-13, what="var_noise_value")
    assert_close(gp.get_value_parameters_model, g["value_parameters_model"], rtol=1e-12)
    assert gp.dimension_parameters == int(g["dimension_parameters"])
    assert list(gp.length_scale_indexes) == list(g["length_scale_indexes"])
    assert len(gp.slice_samplers) == int(g["n_slice_samplers"])
    assert_close(enc_bounds(gp.get_bounds_parameters), g["bounds_parameters"], rtol=1e-13,
                 what="bounds")
    assert_close(gp.start_point_sampler, g["start_point_sampler"], rtol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_log_priors_match_the_reference(name):
    g, ser = load_model(name)
    gp = build_model(name, g, ser)
    got = np.array([gp._log_prior_parameters(th) for th in g["thetas_lp"]], dtype=float)
    assert np.array_equal(np.isneginf(got), np.isneginf(g["log_prior"]))
    fin = np.isfinite(g["log_prior"])
    assert_close(got[fin], g["log_prior"][fin], rtol=1e-12, atol=1e-13, what="log prior")


@pytest.mark.parametrize("name", CASES)
def test_log_prob_parameters_with_oracle_likelihood(name):
    g, ser = load_model(name)
    gp = build_model(name, g, ser)
    use_oracle_likelihood(gp, g)
    got = np.array([gp.log_prob_parameters(th) for th in g["thetas_lp"]], dtype=float)
    assert np.array_equal(np.isneginf(got), np.isneginf(g["log_prob"]))
    fin = np.isfinite(g["log_prob"])
    assert_close(got[fin], g["log_prob"][fin], rtol=1e-9, what="log_prob_parameters")
    got_b = np.array(gp.log_prob_parameters_batch(list(g["thetas_lp"])), dtype=float)
    assert np.array_equal(got, got_b)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("batched", [False, True])
def test_slice_chain_is_the_reference_chain_given_the_same_log_prob(name, batched):
    """Seeded chain: same numpy.random call order, same interval doubling, acceptance test and
    shrinkage as sbo/samplers/slice_sampling.py:49-278 -- the kept samples agree to 1e-9 and the
    global random stream is left in the same state."""
    g, ser = load_model(name)
    gp = build_model(name, g, ser)
    old = GPFittingGaussian.batched_slice_probes
    GPFittingGaussian.batched_slice_probes = batched
    try:
        gp.set_samplers()
        use_oracle_likelihood(gp, g)
        np.random.seed(int(g["chain_seed"]))
        n_before = len(gp.samples_parameters)
        chain = np.stack(gp.sample_parameters(int(g["chain_n"])))
        after = np.random.rand()
    finally:
        GPFittingGaussian.batched_slice_probes = old
    assert_close(chain, g["chain"], rtol=1e-9, atol=1e-12, what="chain")
    assert after == float(g["chain_rand_after"])
    # the one-sampler branch of the reference returns the samples without storing them (:319-336)
    expect = n_before + (0 if int(g["n_slice_samplers"]) == 1 else len(chain))
    assert len(gp.samples_parameters) == expect


@pytest.mark.parametrize("name", CASES)
def test_serialize_matches_the_reference_dict(name):
    g, ser = load_model(name)
    gp = build_model(name, g, ser)
    use_oracle_likelihood(gp, g)
    np.random.seed(int(g["chain_seed"]))
    gp.sample_parameters(int(g["chain_n"]))
    mine = json.loads(json.dumps(gp.serialize(), default=lambda o: np.asarray(o).tolist()))
    assert set(mine) == set(ser)
    for k in ser:
        a, b = mine[k], ser[k]
        if k in ("kernel_values", "mean_value", "var_noise_value", "start_point_sampler",
                 "samples_parameters"):
            assert_close(np.array(a, dtype=float), np.array(b, dtype=float), rtol=1e-9, atol=1e-12,
                         what=k)
        elif k in ("data", "training_data"):
            for kk in b:
                assert_close(np.array(a[kk], dtype=float), np.array(b[kk], dtype=float), rtol=0,
                             what=k + "." + kk)
        else:
            assert a == b, (k, a, b)


@pytest.mark.parametrize("name", CASES)
def test_deserialize_a_reference_dict(name):
    """A dict written by the reference's serialize() rebuilds the model (f4): same parameter
    values, samples and data; serialising again gives the same dict."""
    g, ser = load_model(name)
    gp = GPFittingGaussian.deserialize(json.loads(json.dumps(ser)), noise=bool(int(g["noise"])))
    assert_close(gp.get_value_parameters_model,
                 np.concatenate([ser["var_noise_value"], ser["mean_value"], ser["kernel_values"]]),
                 rtol=0)
    assert_close(np.array(gp.samples_parameters), np.array(ser["samples_parameters"]), rtol=0)
    assert_close(gp.data["points"], g["points"], rtol=0)
    again = json.loads(json.dumps(gp.serialize(), default=lambda o: np.asarray(o).tolist()))
    for k in ser:
        if k in ("data", "training_data"):
            continue
        if k in ("kernel_values", "mean_value", "var_noise_value", "start_point_sampler",
                 "samples_parameters"):
            assert_close(np.array(again[k], dtype=float), np.array(ser[k], dtype=float), rtol=0, what=k)
        else:
            assert again[k] == ser[k], k


# ---------------------------------------------------------------------------------------------------
# GPU: the same rows with the CUDA likelihood, plus the device-side rows
# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_log_prob_parameters(name):
    g, ser = load_model(name)
    gp = build_model(name, g, ser)
    got = np.array([gp.log_prob_parameters(th) for th in g["thetas_lp"]], dtype=float).reshape(-1)
    assert np.array_equal(np.isneginf(got), np.isneginf(g["log_prob"]))
    gp.clean_cache()
    got_b = np.array(gp.log_prob_parameters_batch(list(g["thetas_lp"])), dtype=float).reshape(-1)
    assert np.array_equal(np.isneginf(got_b), np.isneginf(g["log_prob"]))
    for i, th in enumerate(g["thetas_lp"]):
        if np.isfinite(g["log_prob"][i]):
            rt = lp_rtol(g, th)
            assert_close(got[i], g["log_prob"][i], rtol=rt, what="log_prob_parameters[%d]" % i)
            assert_close(got_b[i], g["log_prob"][i], rtol=rt, what="log_prob_parameters_batch[%d]" % i)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("batched", [False, True])
def test_gpu_slice_chain_is_the_reference_chain(name, batched):
    g, ser = load_model(name)
    old = GPFittingGaussian.batched_slice_probes
    GPFittingGaussian.batched_slice_probes = batched
    try:
        gp = build_model(name, g, ser)
        np.random.seed(int(g["chain_seed"]))
        chain = np.stack(gp.sample_parameters(int(g["chain_n"])))
        after = np.random.rand()
    finally:
        GPFittingGaussian.batched_slice_probes = old
    assert_close(chain, g["chain"], rtol=1e-9, atol=1e-12, what="chain")
    assert after == float(g["chain_rand_after"])
    lp = np.array([gp.log_prob_parameters(c) for c in chain], dtype=float).reshape(-1)
    want = np.asarray(g["chain_log_prob"], dtype=float).reshape(-1)
    for i, c in enumerate(chain):
        assert_close(lp[i], want[i], rtol=lp_rtol(g, g["chain"][i]), what="chain log prob[%d]" % i)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_historical_best(name):
    g, ser = load_model(name)
    gp = build_model(name, g, ser)
    assert gp.get_historical_best_solution() == float(g["hist_best"])
    gp.best_solution = {}
    assert_close(gp.get_historical_best_solution(noisy_evaluations=True), g["hist_best_noisy"],
                 rtol=1e-9, what="hist best (posterior mean)")


def build_bq(name, g, gp):
    from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import BayesianQuadrature
    if name == "product_uniform":
        return BayesianQuadrature(gp, [0], UNIFORM_FINITE, parameters_distribution={TASKS: 2})
    if name == "product_weighted":
        return BayesianQuadrature(gp, [0, 1], WEIGHTED_UNIFORM_FINITE, parameters_distribution={
            "weights": g["weights"], "domain_random": np.arange(4).reshape(4, 1)})
    if name == "scaled_gamma":
        bq = BayesianQuadrature(gp, [0, 1], GAMMA, parameters_distribution={"a": [2.0], "scale": [1.0]})
        bq.set_w_sets(g["wset"])
        bq.set_w_double_sets(g["w100"])
        return bq
    raise KeyError(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["product_uniform", "product_weighted", "scaled_gamma"])
def test_gpu_bq_historical_best(name):
    g, ser = load_model(name)
    gp = build_model(name, g, ser)
    bq = build_bq(name, g, gp)
    assert_close(bq.get_historical_best_solution(), g["bq_hist_best"], rtol=1e-9, what="bq hist best")
    if "bq_hist_best_at_theta" in g:
        th = g["bq_hist_best_theta"]
        # Sigma of the reference's own test vector has cond = 2.3e9
        assert_close(bq.get_historical_best_solution(th[0], th[1], th[2:]), g["bq_hist_best_at_theta"],
                     rtol=2e-6, what="bq hist best at theta")


@pytest.mark.gpu
def test_gpu_gamma_prior_variance_is_the_double_expectation():
    """a22 for gamma W: E_W E_W' k((x,W),(x,W')) over the full 100 x 100 grid of two independent
    draws (sbo/lib/expectations.py:93-112 with double=True), W samples replayed from the fixture."""
    g, ser = load_model("scaled_gamma")
    gp = build_model("scaled_gamma", g, ser)
    bq = build_bq("scaled_gamma", g, gp)
    th = g["quadrate_cov_theta"]
    got = np.array([bq.evaluate_quadrate_cov(g["xq"][p:p + 1], th[2:]) for p in range(3)])
    assert_close(got, g["quadrate_cov"], rtol=1e-12, what="evaluate_quadrate_cov (gamma)")
