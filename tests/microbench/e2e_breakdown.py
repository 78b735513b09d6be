"""Where the end-to-end (host buffers, Python API) step spends its time against the device-resident
step of bench.py: phases of SBO.evaluate_mc_bayesian_candidate_points_no_restarts, synchronised."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian  # noqa: E402
from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import BayesianQuadrature  # noqa: E402
from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO  # noqa: E402
from bayesian_quadrature_optimization_b200.lib.constant import SCALED_KERNEL, MATERN52_NAME, GAMMA  # noqa: E402

wl = bench.make_workload()
td = {"points": wl["pts"], "evaluations": wl["y"], "var_noise": []}
bounds = [[0, 1]] * bench.DX + [[0, 20]] * bench.DW
gp = GPFittingGaussian([SCALED_KERNEL, MATERN52_NAME], td, [bench.DX + bench.DW], bounds_domain=bounds,
                       noise=True, kernel_values=list(wl["thetas"][0][2:]), mean_value=[0.0],
                       var_noise_value=[1e-4], define_samplers=False)
gp.samples_parameters = [t.copy() for t in wl["thetas"]]
quad = BayesianQuadrature(gp, list(range(bench.DX)), GAMMA, parameters_distribution={"a": [2.0], "scale": [1.0]})
quad.set_w_sets(wl["wsets"])
sbo = SBO(quad)
C = bench.CAND_PER_STEP


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


np.random.seed(1)
for rep in range(3):
    cands = wl["cands"][rep * C:(rep + 1) * C]
    t0 = sync()
    samples, start = sbo.generate_samples_starting_points_evaluate_mc(bench.N_Z, bench.N_RESTARTS_MC, cache=False)
    t1 = sync()
    bq = quad.engine_for(np.array(gp.samples_parameters[-bench.N_HYPER:]))
    t2 = sync()
    ub = bq.setup(cands)
    t3 = sync()
    mx, arg, ne = bq.inner_maximize(ub, samples, start, maxiter=10, factr=1e12, count_evals=True)
    t4 = sync()
    gvb, is_nan = bq.grad_vector_b(ub, arg)
    t5 = sync()
    val, grad = bq.acq_reduce(ub, mx, gvb, is_nan, samples)
    v, g = val.cpu().numpy(), grad.cpu().numpy()
    t6 = sync()
    print("rep %d: samples/starts %.2f ms, engine_for %.2f, stage B %.2f, maximiser %.2f (%.2f evals per (unit,Z)), "
          "grad_vector_b %.2f, reduce + D2H %.2f; total %.2f ms"
          % (rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3),
             float(ne.sum().item()) / (C * bench.N_HYPER * bench.N_Z), 1e3 * (t5 - t4), 1e3 * (t6 - t5), 1e3 * (t6 - t0)))
