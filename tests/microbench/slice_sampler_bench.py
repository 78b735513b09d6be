"""Hyper-posterior slice sampling at the bench shape (n = 2048, d = 6): sweeps per second with
batched probes against one probe at a time (same chain, see tests/test_gpu_facade.py)."""
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
n_sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for batched in (True, False):
    GPFittingGaussian.batched_slice_probes = batched
    gp = GPFittingGaussian([SCALED_KERNEL, MATERN52_NAME], td, [bench.DX + bench.DW],
                           bounds_domain=bounds, noise=True, kernel_values=list(wl["thetas"][0][2:]),
                           mean_value=[0.0], var_noise_value=[1e-4], max_steps_out=1000,
                           random_seed=1)
    probes = {"n": 0, "calls": 0}
    ob, os_ = gp.log_likelihood_batch, gp.log_likelihood

    hist = {}

    def cb(th, ob=ob):
        probes["n"] += len(th)
        probes["calls"] += 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = ob(th)
        torch.cuda.synchronize()
        h = hist.setdefault(len(th), [0, 0.0, 0])
        h[0] += 1
        h[1] += time.perf_counter() - t0
        h[2] += int(np.sum(~np.isfinite(r)))
        return r

    def cs(*a, os_=os_):
        probes["n"] += 1
        probes["calls"] += 1
        return os_(*a)

    gp.log_likelihood_batch, gp.log_likelihood = cb, cs
    gp.sample_parameters(1, random_seed=3)  # warm-up
    torch.cuda.synchronize()
    probes["n"] = probes["calls"] = 0
    t0 = time.perf_counter()
    s = gp.sample_parameters(n_sweeps, random_seed=4)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for k in sorted(hist):
        print("   S=%d: %d calls, %.2f ms each, %d failed factorisations" % (k, hist[k][0], 1e3 * hist[k][1] / hist[k][0], hist[k][2]))
    print("batched=%s: %d sweeps in %.3f s = %.2f sweeps/s; %d likelihood probes in %d calls "
          "(%.2f ms per call); last sample %s"
          % (batched, n_sweeps, dt, n_sweeps / dt, probes["n"], probes["calls"],
             1e3 * dt / max(1, probes["calls"]), np.array2string(s[-1], precision=4)))

# independent chains side by side (sample_parameters_chains): sweeps per second over all chains
GPFittingGaussian.batched_slice_probes = True
for n_chains in (4, 8, 16):
    gp = GPFittingGaussian([SCALED_KERNEL, MATERN52_NAME], td, [bench.DX + bench.DW],
                           bounds_domain=bounds, noise=True, kernel_values=list(wl["thetas"][0][2:]),
                           mean_value=[0.0], var_noise_value=[1e-4], max_steps_out=1000,
                           random_seed=1)
    gp.sample_parameters_chains(1, n_chains=n_chains, random_seed=3, burn_in_sweeps=0)  # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gp.sample_parameters_chains(2, n_chains=n_chains, random_seed=4, burn_in_sweeps=0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("chains=%d: %d sweeps in %.3f s = %.2f sweeps/s" % (n_chains, 2 * n_chains, dt, 2 * n_chains / dt))
