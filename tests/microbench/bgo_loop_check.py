"""A few iterations of bgo() (BQO and EI) on a toy problem: the BO loop end to end on the device
(data append, factor extension, chain advance, acquisition maximisation, posterior-mean optimum)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from bayesian_quadrature_optimization_b200.services.bgo import bgo  # noqa: E402


def g(x):
    return [-(x[0] - 0.3) ** 2 - 0.5 * (x[1] - 0.6) ** 2]


def f(xw):
    return [-(xw[0] - 0.3) ** 2 - 0.5 * (xw[1] - 0.6) ** 2 + (0.2 if int(xw[2]) == 0 else -0.2)]


t0 = time.perf_counter()
sol = bgo(g, [(0.0, 1.0), (0.0, 1.0)], integrand_function=f, bounds_domain_w=[[0, 1]],
          type_bounds=[0, 0, 1], name_method='bqo', n_iterations=4, random_seed=5, n_training=12,
          thinning=2, n_burning=20, max_steps_out=20, n_restarts=4, n_samples_parameters=3, maxepoch=5,
          n_samples_mc=3, n_restarts_mc=3, default_n_samples=6, default_n_samples_parameters=5,
          n_restarts_mean=20)
print("bqo: %s value %.5f (true optimum [0.3, 0.6], 0) in %.1f s" % (
    np.round(sol["optimal_solution"], 3), sol["optimal_value"], time.perf_counter() - t0))
assert sol["optimal_solution"].shape == (2,)
assert np.all(sol["optimal_solution"] >= 0) and np.all(sol["optimal_solution"] <= 1)
t0 = time.perf_counter()
sol = bgo(g, [(0.0, 1.0), (0.0, 1.0)], name_method='ei', n_iterations=4, random_seed=5, n_training=8,
          thinning=2, n_burning=20, max_steps_out=20, n_restarts=4, n_samples_parameters=3, maxepoch=5,
          n_restarts_mean=20)
print("ei:  %s value %.5f in %.1f s" % (np.round(sol["optimal_solution"], 3), sol["optimal_value"],
                                       time.perf_counter() - t0))
assert np.all(sol["optimal_solution"] >= 0) and np.all(sol["optimal_solution"] <= 1)
print("OK")
