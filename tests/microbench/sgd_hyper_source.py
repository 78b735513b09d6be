"""SBO.optimize: where the SGD's hyper-samples come from.  Default: drawn from the resident posterior
samples; advance_chain=True: the reference's way (sbo.py:1129), the chain advanced every epoch.
Both runs start from the same restarts and are scored by ONE common estimator (an independent, longer
chain of posterior samples, many Z).  Writes gpurun_out/sgd_hyper_source.json."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian  # noqa: E402
from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import BayesianQuadrature  # noqa: E402
from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO  # noqa: E402
from bayesian_quadrature_optimization_b200.lib import constant as K  # noqa: E402


def problem(n=40, T=3, seed=0):
    rng = np.random.RandomState(seed)
    x = rng.uniform(0, 1, (n, 2))
    t = rng.randint(0, T, n)
    pts = np.concatenate([x, t.reshape(-1, 1).astype(float)], 1)
    y = np.sin(4 * x[:, 0]) * np.cos(3 * x[:, 1]) + 0.4 * t + 0.05 * rng.normal(0, 1, n)
    return {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}, T


def model(td, T, seed):
    gp = GPFittingGaussian(
        [K.PRODUCT_KERNELS_SEPARABLE, K.MATERN52_NAME, K.TASKS_KERNEL_NAME], td, [3, 2, T],
        bounds_domain=[[0, 1], [0, 1], list(range(T))], type_bounds=[0, 0, 1], thinning=2,
        n_burning=30, max_steps_out=50, random_seed=seed, **{K.SAME_CORRELATION: True})
    bq = BayesianQuadrature(gp, [0, 1], K.UNIFORM_FINITE, parameters_distribution={K.TASKS: T})
    return gp, bq, SBO(bq)


def run(advance, td, T, seed, maxepoch):
    gp, bq, sbo = model(td, T, 11)
    gp.sample_parameters(8, random_seed=5)
    t0 = time.perf_counter()
    res = sbo.optimize(random_seed=seed, monte_carlo=True, n_samples=2, n_restarts_mc=3, n_restarts=6,
                       n_samples_parameters=2, start_new_chain=False, maxepoch=maxepoch,
                       default_n_samples=10, default_n_samples_parameters=8, start_ei=False,
                       factr=1e12, maxiter=10, advance_chain=advance)
    return res["solution"], float(np.asarray(res["optimal_value"]).reshape(-1)[0]), time.perf_counter() - t0


def main(seeds=(1, 2, 3, 4, 5, 6), maxepoch=20):
    td, T = problem()
    gps, bqs, sbos = model(td, T, 23)                 # the common scorer: its own, longer chain
    gps.sample_parameters(24, random_seed=17)
    rows = []
    for seed in seeds:
        xa, va, ta = run(False, td, T, seed, maxepoch)
        xb, vb, tb = run(True, td, T, seed, maxepoch)
        np.random.seed(99)
        sc = sbos.evaluate_mc_bayesian_candidate_points_no_restarts(
            np.array([xa, xb]), 24, 60, n_restarts=3, compute_max_mean=True, compute_gradient=False,
            factr=1e12, maxiter=10)["evaluations"]
        rows.append({"seed": int(seed), "resident": {"x": xa.tolist(), "own_value": va, "score": float(sc[0]), "s": ta},
                     "chain": {"x": xb.tolist(), "own_value": vb, "score": float(sc[1]), "s": tb}})
        print("seed %d: resident %.6g (%.2f s) chain %.6g (%.2f s)  x_res %s x_chain %s"
              % (seed, sc[0], ta, sc[1], tb, np.round(xa, 3), np.round(xb, 3)))
    a = np.array([r["resident"]["score"] for r in rows])
    b = np.array([r["chain"]["score"] for r in rows])
    out = {"rows": rows, "mean_resident": float(a.mean()), "mean_chain": float(b.mean()),
           "mean_diff": float((a - b).mean()), "sd_diff": float((a - b).std(ddof=1)),
           "seconds_resident": float(np.mean([r["resident"]["s"] for r in rows])),
           "seconds_chain": float(np.mean([r["chain"]["s"] for r in rows]))}
    print({k: v for k, v in out.items() if k != "rows"})
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/sgd_hyper_source.json", "w"), indent=1)
    return out


if __name__ == "__main__":
    main()
