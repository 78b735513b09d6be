"""Time of one batched log-likelihood (assemble + Cholesky + alpha + reduction) against the number
of hyper-samples in the batch, n = 2048, through the Python facade (fresh parameters per call)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian  # noqa: E402
from bayesian_quadrature_optimization_b200.lib.constant import SCALED_KERNEL, MATERN52_NAME  # noqa: E402

wl = bench.make_workload()
td = {"points": wl["pts"], "evaluations": wl["y"], "var_noise": []}
bounds = [[0, 1]] * bench.DX + [[0, 20]] * bench.DW
gp = GPFittingGaussian([SCALED_KERNEL, MATERN52_NAME], td, [bench.DX + bench.DW], bounds_domain=bounds,
                       noise=True, kernel_values=list(wl["thetas"][0][2:]), mean_value=[0.0],
                       var_noise_value=[1e-4], define_samplers=False)
rng = np.random.RandomState(0)
base = wl["thetas"][0]
for S in (1, 2, 4, 8, 20):
    for inv in (False, True):
        # inv: alpha from L^-1 produced by extra tasks of the factorisation launch (two mat-vecs)
        # instead of two triangular vector solves
        ts = []
        for rep in range(6):
            th = base[None, :] * (1.0 + 0.02 * rng.normal(0, 1, (S, base.size)))
            th[:, 1] = 0.0
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            hb = gp.hyper_batch(th)
            hb.factorize(max_tries=7, with_inverse=inv)
            hb.log_likelihood_host()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print("S=%2d, alpha from %s: %.2f ms per call (min of 5 after warm-up; all: %s)"
              % (S, "L^-1 " if inv else "solves", 1e3 * min(ts[1:]),
                 " ".join("%.1f" % (1e3 * t) for t in ts)))
