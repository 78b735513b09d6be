"""Fixtures at BASELINE.json's configs C1 and C2 from the UNMODIFIED reference (build container):

    python tests/golden/make_golden_configs.py

C1 = the reference's own toy, problems/test_problem_with_tasks/main.py:4-17 with the data of
tests/acquisition_functions/test_sbo.py:54-78 (n = 100 points linspace(0, 100) x random task,
seed 5, a sampled GP path +-10 by task) and the parameter vector of test_sbo.py:523-528,
theta = [1, 5, 50, 9.6, -3, -0.1], plus a second, better conditioned one.
C2 = branin_mt-shaped (problems/branin_mt/main.py:20-50): d_x = 2, T = 12, same_correlation,
weighted_uniform_finite with the flattened `probabilities`, n = 200 (the seeded workload of
tests/golden/make_lbfgs_fixture.py:workload('c2')).

Stored per config (tests/golden/config_<c1|c2>.npz): inputs, log-likelihood and gradient, Cholesky
factor, GP posterior at a few query points, EI, every stage-B / stage-C quantity of
make_golden.py:bq_outputs, the discretised KG value and gradient (C1), and the chosen task
`MultiTasks.choose_best_task_given_x` (argmax over tasks of the Bayesian-averaged EI,
sbo/acquisition_functions/multi_task.py:82-95) with the hyper-samples fixed to `thetas`.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import make_golden as mg  # noqa: E402  (installs the import shim)

warnings.filterwarnings("ignore")

from stratified_bayesian_optimization.models.gp_fitting_gaussian import GPFittingGaussian  # noqa: E402
from stratified_bayesian_optimization.numerical_tools.bayesian_quadrature import BayesianQuadrature  # noqa: E402
from stratified_bayesian_optimization.acquisition_functions.sbo import SBO  # noqa: E402
from stratified_bayesian_optimization.acquisition_functions.multi_task import MultiTasks  # noqa: E402
from stratified_bayesian_optimization.kernels.matern52 import Matern52  # noqa: E402
from stratified_bayesian_optimization.lib.sample_functions import SampleFunctions  # noqa: E402
from stratified_bayesian_optimization.lib.constant import (  # noqa: E402
    PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME, UNIFORM_FINITE,
    WEIGHTED_UNIFORM_FINITE, TASKS, SAME_CORRELATION)


def gp_outputs_light(gp, thetas, query, out):
    """make_golden.gp_outputs without the n x n covariance and gradient matrices (they are pinned
    at small n); keeps what passes through the factorisation."""
    chols, llh, gllh, pm, pc = [], [], [], [], []
    for th in thetas:
        vn, mu, pk = th[0], th[1], th[2:]
        gp.cache_chol_cov = {}
        gp.cache_sol_chol_y_unbiased = {}
        chol, _ = gp._chol_cov_including_noise(vn, pk)
        chols.append(chol)
        llh.append(gp.log_likelihood(vn, mu, pk))
        gllh.append(gp.grad_log_likelihood(vn, mu, pk))
        post = gp.compute_posterior_parameters(query, vn, mu, pk)
        pm.append(post["mean"])
        pc.append(post["cov"])
    out.update(chol=np.stack(chols), llh=np.array(llh), grad_llh=np.stack(gllh),
               post_mean=np.stack(pm), post_cov=np.stack(pc))


def chosen_tasks(bq_x, n_tasks, gp, thetas, xs, out):
    """choose_best_task_given_x with samples_parameters = thetas (multi_task.py:82-95)."""
    gp.samples_parameters = [np.array(t) for t in thetas]
    mk = MultiTasks(bq_x, n_tasks)
    tasks, values, all_values = [], [], []
    for x in xs:
        gp.cache_chol_cov = {}
        gp.cache_sol_chol_y_unbiased = {}
        gp.cache_cov_n = {}
        gp.best_solution = {}
        t, v = mk.choose_best_task_given_x(np.asarray(x, dtype=float))
        tasks.append(int(t))
        values.append(float(v))
        row = []
        for i in range(n_tasks):
            point = np.concatenate((x, [i])).reshape(1, -1)
            vals = []
            for th in thetas:
                gp.cache_cov_n = {}
                vals.append(mk.ei_tasks.evaluate(point, th[0], th[1], th[2:])[0])
            row.append(np.mean(vals))
        all_values.append(row)
    out.update(task_x=np.array(xs, dtype=float), chosen_task=np.array(tasks),
               chosen_task_value=np.array(values), task_ei_values=np.array(all_values))


def config_c1():
    np.random.seed(5)
    n_points = 100
    points = np.linspace(0, 100, n_points).reshape([n_points, 1])
    tasks = np.random.randint(2, size=(n_points, 1))
    add = [10, -10]
    kernel = Matern52.define_kernel_from_array(1, np.array([100.0, 1.0]))
    function = SampleFunctions.sample_from_gp(points, kernel)
    for i in range(n_points):
        function[0, i] += add[tasks[i, 0]]
    points = np.concatenate((points, tasks), axis=1)
    function = function[0, :]
    td = {"evaluations": list(function), "points": points, "var_noise": []}
    gp = GPFittingGaussian([PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME], td,
                           [2, 1, 2], bounds_domain=[[0, 100], [0, 1]], type_bounds=[0, 1])
    thetas = np.array([[1.0, 5.0, 50.0, 9.6, -3.0, -0.1],
                       [0.8, 4.0, 35.0, 2.0, 0.5, 1.2]])
    query = np.array([[3.0, 0.0], [47.5, 1.0], [99.0, 0.0], [52.5, 0.0]])
    out = dict(points=np.asarray(points, dtype=float), evaluations=function, thetas=thetas,
               query=query, kind=2, n_tasks=2, same_correlation=0, dx=1)
    gp_outputs_light(gp, thetas, query, out)
    mg.ei_outputs(gp, thetas, query, out)
    bq = BayesianQuadrature(gp, [0], UNIFORM_FINITE, {TASKS: 2})
    disc = np.linspace(0, 100, 100).reshape(-1, 1)  # number_points_each_dimension = [100]
    sbo = SBO(bq, disc)
    xq = np.array([[3.0], [52.5], [77.7]])
    cands = np.array([[52.5, 0.0], [10.0, 1.0], [80.0, 1.0]])
    zs = np.array([0.7, -1.3])
    out.update(xq=xq, cands=cands, zs=zs, discretization=disc)
    mg.bq_outputs(bq, sbo, thetas, xq, cands, zs, out)
    mg.kg_outputs(bq, sbo, thetas, cands, out)
    bq_x = BayesianQuadrature(gp, [0], UNIFORM_FINITE, {TASKS: 2}, model_only_x=True)
    chosen_tasks(bq_x, 2, gp, thetas, [[3.0], [40.0], [52.5], [90.0]], out)
    return out


def config_c2():
    from make_lbfgs_fixture import workload
    wl = workload("c2")
    T = wl["n_tasks"]
    td = {"points": wl["pts"].tolist(), "evaluations": list(wl["y"]), "var_noise": []}
    gp = GPFittingGaussian([PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME], td,
                           [3, 2, T], bounds_domain=[[0, 1], [0, 1], list(range(T))],
                           type_bounds=[0, 0, 1], **{SAME_CORRELATION: True})
    thetas = wl["thetas"][:2]
    rng = np.random.RandomState(17)
    query = np.concatenate([rng.uniform(0, 1, (4, 2)), np.array([[0.], [11.], [5.], [7.]])], 1)
    out = dict(points=wl["pts"], evaluations=wl["y"], thetas=thetas, query=query, kind=2, n_tasks=T,
               same_correlation=1, dx=2, weights=wl["weights"])
    gp_outputs_light(gp, thetas, query, out)
    mg.ei_outputs(gp, thetas, query, out)
    params = {"weights": wl["weights"], "domain_random": np.arange(T).reshape(T, 1)}
    bq = BayesianQuadrature(gp, [0, 1], WEIGHTED_UNIFORM_FINITE, parameters_distribution=params)
    g = np.linspace(0, 1, 6)
    disc = np.array([[a, b] for a in g for b in g])
    sbo = SBO(bq, disc)
    xq = rng.uniform(0, 1, (3, 2))
    cands = wl["cands"][[3, 40, 119]]
    zs = np.array([0.4, -0.9])
    out.update(xq=xq, cands=cands, zs=zs, discretization=disc)
    mg.bq_outputs(bq, sbo, thetas, xq, cands, zs, out)
    mg.kg_outputs(bq, sbo, thetas, cands, out)
    bq_x = BayesianQuadrature(gp, [0, 1], WEIGHTED_UNIFORM_FINITE, parameters_distribution=params,
                              model_only_x=True)
    chosen_tasks(bq_x, T, gp, thetas, rng.uniform(0, 1, (4, 2)).tolist(), out)
    return out


def main():
    only = set(sys.argv[1:])
    for name, fn in (("c1", config_c1), ("c2", config_c2)):
        if only and name not in only:
            continue
        out = fn()
        np.savez_compressed(os.path.join(HERE, "config_%s.npz" % name),
                            **{k: np.asarray(v) for k, v in out.items()})
        print(name, {k: np.asarray(v).shape for k, v in out.items()})
        print("   llh", out["llh"], "chosen tasks", out["chosen_task"], out["chosen_task_value"])


if __name__ == "__main__":
    main()
