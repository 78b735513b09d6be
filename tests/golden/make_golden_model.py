"""Model-level fixtures from the UNMODIFIED reference (run in the build container only):

    python tests/golden/make_golden_model.py

For each seeded model this stores what the reference's own GPFittingGaussian / BayesianQuadrature
objects produce for the rows that round 1 left unpinned:

* the data-driven defaults (kernel_values, mean_value, var_noise_value), parameter bounds and
  length-scale indexes of `set_parameters_kernel` (sbo/models/gp_fitting_gaussian.py:355-422);
* every prior's log-density and `log_prob_parameters` (:1543-1567) on a list of parameter vectors,
  some of them outside the support of the priors;
* a seeded `sample_parameters` chain (:289-353, sbo/samplers/slice_sampling.py:49-278) together
  with the log-probability of every kept sample;
* `get_historical_best_solution` of the GP (:1590-1626) and of the quadrature model
  (sbo/numerical_tools/bayesian_quadrature.py:1542-1577);
* `serialize()` (:578-620) as JSON.

Output: tests/golden/model_<case>.npz and model_<case>.json.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
warnings.filterwarnings("ignore")

import stratified_bayesian_optimization.samplers.slice_sampling as ref_slice  # noqa: E402
from stratified_bayesian_optimization.models.gp_fitting_gaussian import GPFittingGaussian  # noqa: E402
from stratified_bayesian_optimization.numerical_tools.bayesian_quadrature import BayesianQuadrature  # noqa: E402
from stratified_bayesian_optimization.lib.constant import (  # noqa: E402
    PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME, SCALED_KERNEL, UNIFORM_FINITE,
    WEIGHTED_UNIFORM_FINITE, GAMMA, TASKS, SAME_CORRELATION)

# python 3: `npr.shuffle(range(n))` (slice_sampling.py:62-63) needs a mutable sequence; the module
# namespace gets a list-returning `range` (the reference file itself is untouched)
ref_slice.range = lambda *a: list(range(*a))


def enc_bounds(bounds):
    out = np.full((len(bounds), 2), np.nan)
    for i, b in enumerate(bounds):
        for j in (0, 1):
            if b[j] is not None:
                out[i, j] = b[j]
    return out


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return jsonable(o.tolist())
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.bool_,)):
        return bool(o)
    return o


def wire_samplers_noise(gp):
    """noise=True without observed noise variances: the reference's set_samplers raises TypeError
    (`len(ignore_index)` with ignore_index None, gp_fitting_gaussian.py:248), so such models can
    only be built with define_samplers=False.  To pin the chain for them as well, the reference's
    own SliceSampling objects are wired here the way set_samplers does for the other models, with
    nothing ignored, and start_point_sampler set as at :262-269."""
    from stratified_bayesian_optimization.lib.util_gp_fitting import wrapper_log_prob
    indexes = [i for i in range(gp.dimension_parameters) if i not in gp.length_scale_indexes]
    gp.slice_samplers = [
        ref_slice.SliceSampling(wrapper_log_prob, indexes, ignore_index=None,
                                max_steps_out=gp.max_steps_out, component_wise=False),
        ref_slice.SliceSampling(wrapper_log_prob, gp.length_scale_indexes,
                                max_steps_out=gp.max_steps_out, component_wise=True)]
    gp.samples_parameters = [gp.get_value_parameters_model]
    gp.start_point_sampler = gp.get_value_parameters_model


def model_outputs(gp, thetas_lp, seed, n_samples, out):
    out["kernel_values"] = np.array(gp.kernel_values, dtype=float)
    out["mean_value"] = np.array(gp.mean_value, dtype=float)
    out["var_noise_value"] = np.array(gp.var_noise_value, dtype=float)
    out["value_parameters_model"] = gp.get_value_parameters_model
    out["dimension_parameters"] = gp.dimension_parameters
    out["length_scale_indexes"] = np.array(gp.length_scale_indexes)
    out["bounds_parameters"] = enc_bounds(gp.get_bounds_parameters)
    out["n_slice_samplers"] = len(gp.slice_samplers)
    out["start_point_sampler"] = np.array(gp.start_point_sampler, dtype=float)
    base = gp.get_value_parameters_model
    thetas = [base] + [np.asarray(t, dtype=float) for t in thetas_lp]
    lprior, lprob, llh = [], [], []
    for th in thetas:
        lp = 0.0
        index = 0
        for parameter in gp.get_parameters_model:
            lp += parameter.log_prior(th[index:index + parameter.dimension])
            index += parameter.dimension
        lprior.append(float(lp))
        gp.cache_chol_cov = {}
        gp.cache_sol_chol_y_unbiased = {}
        lprob.append(float(gp.log_prob_parameters(th)))
        llh.append(float(gp.log_likelihood(th[0], th[1], th[2:])) if np.isfinite(lp) else np.nan)
    out.update(thetas_lp=np.stack(thetas), log_prior=np.array(lprior), log_prob=np.array(lprob),
               llh_lp=np.array(llh))
    # historical best (plain and through the posterior mean)
    gp.best_solution = {}
    out["hist_best"] = gp.get_historical_best_solution()
    gp.best_solution = {}
    out["hist_best_noisy"] = gp.get_historical_best_solution(noisy_evaluations=True)
    gp.best_solution = {}
    # seeded chain
    np.random.seed(seed)
    start = np.array(gp.samples_parameters[-1], dtype=float)
    chain = gp.sample_parameters(n_samples)
    chain = np.stack(chain)
    out.update(chain_seed=seed, chain_n=n_samples, chain_start=start, chain=chain,
               chain_log_prob=np.array([gp.log_prob_parameters(c) for c in chain]),
               chain_rand_after=np.random.rand())
    return gp.serialize()


def model_product_uniform():
    rng = np.random.RandomState(5)
    n = 30
    x = np.linspace(0, 100, n)
    t = rng.randint(0, 2, n)
    pts = np.stack([x, t.astype(float)], 1)
    y = x / 10.0 + np.where(t == 0, 10.0, -10.0) + rng.normal(0, 0.1, n)
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}
    gp = GPFittingGaussian([PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME], td,
                           [2, 1, 2], bounds_domain=[[0, 100], [0, 1]], type_bounds=[0, 1],
                           thinning=1, max_steps_out=3)
    out = dict(points=pts, evaluations=y, kind=2, n_tasks=2, same_correlation=0, dx=1, noise=0,
               thinning=1, max_steps_out=3)
    base = gp.get_value_parameters_model
    thetas = [base * np.array([1, 1, 1.2, 1, 1, 1]) + np.array([0, 0, 0, 0.1, -0.2, 0.05]),
              np.array([0.0, 0.0, 50.0, 9.6, -3.0, -0.1]),
              np.array([0.0, 0.0, -1.0, 0.2, 0.1, 0.1]),     # negative length scale
              np.array([0.5, 0.0, 30.0, 0.2, 0.1, 0.1]),     # var_noise off its constant
              np.array([0.0, 0.0, 2e5, 0.2, 0.1, 0.1])]      # length scale above its bound
    ser = model_outputs(gp, thetas, 11, 4, out)
    bq = BayesianQuadrature(gp, [0], UNIFORM_FINITE, parameters_distribution={TASKS: 2})
    out["bq_hist_best"] = bq.get_historical_best_solution()
    th = thetas[1]
    out["bq_hist_best_theta"] = th
    out["bq_hist_best_at_theta"] = bq.get_historical_best_solution(th[0], th[1], th[2:])
    return out, ser


def model_product_weighted():
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
                           type_bounds=[0, 0, 1], noise=True, thinning=0, max_steps_out=2,
                           **{SAME_CORRELATION: True})
    out = dict(points=pts, evaluations=y, var_noise_obs=vno, kind=2, n_tasks=T, same_correlation=1,
               dx=2, noise=1, thinning=0, max_steps_out=2,
               weights=np.array([0.1, 0.4, 0.3, 0.2]))
    base = gp.get_value_parameters_model
    thetas = [base + np.array([0, 0, 0.05, -0.03, 0.2, 0.1]),
              np.array([0.0, 0.0, 0.3, 0.4, 0.0, -1.0]),
              np.array([0.0, 0.1, 0.3, 0.4, 0.0, -1.0]),     # mean off its constant
              np.array([0.0, 0.0, 0.3, 0.0, 0.0, -1.0])]     # zero length scale
    ser = model_outputs(gp, thetas, 12, 5, out)
    bq = BayesianQuadrature(gp, [0, 1], WEIGHTED_UNIFORM_FINITE,
                            parameters_distribution={"weights": out["weights"],
                                                     "domain_random": np.arange(T).reshape(T, 1)})
    out["bq_hist_best"] = bq.get_historical_best_solution()
    return out, ser


def model_product_noise():
    """Product kernel, same_correlation, T=3, noise=True and no observed noise variances: every
    prior is active (Gaussian mean, horseshoe var_noise, uniform length scales, task prior)."""
    rng = np.random.RandomState(21)
    n, T = 36, 3
    x = rng.uniform(-1, 1, (n, 2))
    t = rng.randint(0, T, n)
    pts = np.concatenate([x, t.reshape(-1, 1).astype(float)], 1)
    y = np.sin(2 * x[:, 0]) * (1 + 0.2 * t) + x[:, 1] ** 2 + rng.normal(0, 0.1, n)
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}
    gp = GPFittingGaussian([PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME], td,
                           [3, 2, T], bounds_domain=[[-1, 1], [-1, 1], list(range(T))],
                           type_bounds=[0, 0, 1], noise=True, thinning=1, max_steps_out=2,
                           define_samplers=False, **{SAME_CORRELATION: True})
    wire_samplers_noise(gp)
    out = dict(points=pts, evaluations=y, kind=2, n_tasks=T, same_correlation=1, dx=2, noise=1,
               thinning=1, max_steps_out=2)
    base = gp.get_value_parameters_model
    thetas = [base * 1.1,
              base + np.array([0.01, 0.2, 0.1, 0.1, 0.3, -0.2]),
              np.array([-0.1, 0.0, 0.5, 0.5, 0.0, 0.0]),     # negative var_noise
              np.array([0.05, 0.3, 0.5, 0.5, 2.5, -2.0])]
    ser = model_outputs(gp, thetas, 13, 4, out)
    return out, ser


def model_scaled_gamma():
    rng = np.random.RandomState(0)
    n = 50
    x = rng.uniform(0, 1, (n, 2))
    w = rng.gamma(2.0, 1.0, (n, 2))
    pts = np.concatenate([x, w], 1)
    y = np.sin(pts.sum(1)) + 0.01 * rng.normal(0, 1, n)
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}
    gp = GPFittingGaussian([SCALED_KERNEL, MATERN52_NAME], td, [4],
                           bounds_domain=[[0, 1], [0, 1], [0, 10], [0, 10]], noise=True,
                           thinning=1, max_steps_out=2, define_samplers=False)
    wire_samplers_noise(gp)
    out = dict(points=pts, evaluations=y, kind=1, n_tasks=0, same_correlation=0, dx=2, noise=1,
               thinning=1, max_steps_out=2)
    base = gp.get_value_parameters_model
    thetas = [base * 0.9,
              np.array([1e-4, 0.0, 0.3, 0.35, 2.0, 2.2, 1.0]),
              np.array([1e-4, 0.0, 0.3, 0.35, 2.0, 2.2, -1.0]),   # negative sigma2
              np.array([0.0, 0.0, 0.3, 0.35, 2.0, 2.2, 1.0])]     # var_noise = 0
    ser = model_outputs(gp, thetas, 14, 4, out)
    # quadrature model: historical best and the prior variance with replayed W samples
    wset = np.random.RandomState(4).gamma(2.0, 1.0, (10, 2))
    w100 = np.random.RandomState(7).gamma(2.0, 1.0, (2, 100, 2))
    out.update(wset=wset, w100=w100)

    class Mixed(object):
        k = 0

        def rvs(self, a, scale=1.0, size=None):
            if tuple(size) == (10, 2):
                return wset.copy()
            r = w100[self.k % 2].copy()
            self.k += 1
            return r

    import stratified_bayesian_optimization.lib.expectations as ex
    ex.gamma = Mixed()
    bq = BayesianQuadrature(gp, [0, 1], GAMMA, parameters_distribution={"a": [2.0], "scale": [1.0]})
    out["bq_hist_best"] = bq.get_historical_best_solution()
    xq = rng.uniform(0, 1, (3, 2))
    th = thetas[1]
    out["xq"] = xq
    out["quadrate_cov_theta"] = th
    out["quadrate_cov"] = np.array([bq.evaluate_quadrate_cov(xq[p:p + 1], th[2:]) for p in range(3)])
    return out, ser


def model_matern52_plain():
    rng = np.random.RandomState(12)
    n = 45
    pts = rng.uniform(-1, 2, (n, 3))
    y = np.cos(pts[:, 0]) * pts[:, 1] + 0.3 * pts[:, 2] + 0.05 * rng.normal(0, 1, n)
    td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}
    gp = GPFittingGaussian([MATERN52_NAME], td, [3], bounds_domain=[[-1, 2]] * 3, noise=False,
                           thinning=0, max_steps_out=1)
    out = dict(points=pts, evaluations=y, kind=0, n_tasks=0, same_correlation=0, dx=3, noise=0,
               thinning=0, max_steps_out=1)
    base = gp.get_value_parameters_model
    thetas = [base * 1.3, np.array([0.0, 0.0, 0.8, 1.1, 0.6])]
    ser = model_outputs(gp, thetas, 15, 5, out)
    return out, ser


def main():
    cases = [("product_uniform", model_product_uniform),
             ("product_weighted", model_product_weighted),
             ("product_noise", model_product_noise),
             ("scaled_gamma", model_scaled_gamma),
             ("matern52_plain", model_matern52_plain)]
    only = set(sys.argv[1:])
    for name, fn in cases:
        if only and name not in only:
            continue
        out, ser = fn()
        np.savez_compressed(os.path.join(HERE, "model_%s.npz" % name),
                            **{k: np.asarray(v) for k, v in out.items()})
        with open(os.path.join(HERE, "model_%s.json" % name), "w") as f:
            json.dump(jsonable(ser), f)
        print(name, {k: np.asarray(v).shape for k, v in out.items()})
        print("   defaults", out["kernel_values"], out["mean_value"], out["var_noise_value"])
        print("   log_prob", out["log_prob"])
        print("   chain", out["chain"][-1])


if __name__ == "__main__":
    main()
