"""bgo(): the public entry point of the reference (sbo/services/bgo.py:37-275) over the CUDA path.

Same keyword signature and return value ({'optimal_solution', 'optimal_value'}).  The BO loop
(sbo/services/bayesian_global_optimization.py:175-391) is kept on the host: per iteration the
acquisition function is maximised on the GPU, the user's black box is evaluated, the point is
appended to the model (all device-resident factors are invalidated, like the reference's caches,
gp_fitting_gaussian.py:492-511), and after the last iteration the posterior mean of the objective
is maximised.  JSON caching of models/training data, spec entities and logging are out of scope
(SURVEY.md section 2).
"""
from __future__ import annotations

import numpy as np

from ..acquisition_functions.ei import EI
from ..acquisition_functions.sbo import SBO
from ..lib.constant import (
    SBO_METHOD, EI_METHOD, PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME,
    SCALED_KERNEL, UNIFORM_FINITE, GAMMA, SAME_CORRELATION, LBFGS_NAME, DEFAULT_N_PARAMETERS,
)
from ..models.gp_fitting_gaussian import GPFittingGaussian
from ..numerical_tools.bayesian_quadrature import BayesianQuadrature
from .domain import DomainService

_possible_optimization_methods = [SBO_METHOD, EI_METHOD]


def _evaluate(function, point, noise, n_samples_noise):
    """TrainingDataService evaluation convention (sbo/services/training_data.py:229-242):
    f(point) -> [value] or, when noisy, f(point, n_samples) -> [value, variance]."""
    if noise:
        out = function(list(point), n_samples_noise)
        return float(out[0]), float(out[1])
    out = function(list(point))
    return float(out[0] if isinstance(out, (list, tuple, np.ndarray)) else out), None


def bgo(objective_function, bounds_domain_x, integrand_function=None, simplex_domain=None,
        noise=False, n_samples_noise=0, bounds_domain_w=None, type_bounds=None, distribution=None,
        parameters_distribution=None, name_method='bqo', n_iterations=50, type_kernel=None,
        dimensions_kernel=None, n_restarts=10, n_best_restarts=0, problem_name=None,
        n_training=None, random_seed=1, mle=False, n_samples_parameters=5, maxepoch=50, thinning=50,
        n_burning=500, max_steps_out=1000, parallel=True, same_correlation=True,
        monte_carlo_sbo=True, n_samples_mc=5, n_restarts_mc=5, n_best_restarts_mc=0, factr_mc=1e12,
        maxiter_mc=10, method_opt_mc=LBFGS_NAME, n_restarts_mean=100, n_best_restarts_mean=10,
        n_samples_parameters_mean=5, maxepoch_mean=50, parallel_training=False,
        default_n_samples_parameters=None, default_n_samples=None, device=None, n_chains=1,
        distributed=False):
    """Maximises the objective function (sbo/services/bgo.py:37-121 documents every argument).
    Three arguments are additions: `device` (CUDA device), `n_chains` (> 1: the hyper-parameter
    samples of every iteration come from that many independent slice-sampling chains advanced side
    by side on the GPU instead of one sequential chain) and `distributed` (replaces the
    `parallel` process pool of sbo/lib/parallel.py:16-76: run the same program on every GPU of the
    node with `torchrun --nproc-per-node N`; the process group is created from the environment
    when it does not exist yet, this process uses cuda:LOCAL_RANK, the hyper-samples of every
    acquisition evaluation are sharded over the ranks and combined with one all_gather +
    rank-ordered sum, so every rank takes bit-identical decisions)."""
    if distributed:
        import os
        import torch
        import torch.distributed as dist
        if device is None:
            device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(torch.device(device))
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device(device))
    np.random.seed(random_seed)
    dim_x = len(bounds_domain_x)
    x_domain = list(range(dim_x))
    if name_method == 'bqo':
        name_method = SBO_METHOD
    if name_method not in _possible_optimization_methods:
        raise Exception("Incorrect BGO method")
    dim_w = 0
    if name_method == SBO_METHOD:
        if type_bounds is None:
            dim_w = len(bounds_domain_w)
        elif type_bounds[-1] == 1:
            dim_w = 1
        else:
            dim_w = len(type_bounds[dim_x:])
    total_dim = dim_w + dim_x
    if type_bounds is None:
        type_bounds = total_dim * [0]
    if bounds_domain_w is None:
        bounds_domain_w = []
    bounds_domain = [list(bound) for bound in list(bounds_domain_x) + list(bounds_domain_w)]
    if type_kernel is None:
        if name_method == SBO_METHOD and type_bounds[-1] == 1:
            type_kernel = [PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME]
            dimensions_kernel = [total_dim, dim_x, len(bounds_domain[-1])]
        else:
            type_kernel = [SCALED_KERNEL, MATERN52_NAME]
            dimensions_kernel = [total_dim]
    if dimensions_kernel is None:
        raise Exception("Not enough inputs to run the BGO algorithm")
    if n_training is None:
        n_training = len(bounds_domain[-1]) if type_bounds[-1] == 1 else 5
    if distribution is None:
        distribution = UNIFORM_FINITE if type_bounds[-1] == 1 else GAMMA
    training_function = integrand_function if name_method == SBO_METHOD else objective_function

    # ---- training data: random points of the domain (TrainingDataService.get_training_data) ----
    points = DomainService.get_points_domain(n_training, bounds_domain, type_bounds=type_bounds,
                                             random_seed=random_seed)
    evaluations, var_noise = [], []
    for p in points:
        v, s2 = _evaluate(training_function, p, noise, n_samples_noise)
        evaluations.append(v)
        if s2 is not None:
            var_noise.append(s2)
    training_data = {'points': points, 'evaluations': evaluations, 'var_noise': var_noise}

    kernel_parameters = {}
    if TASKS_KERNEL_NAME in type_kernel:
        kernel_parameters[SAME_CORRELATION] = same_correlation
    gp_model = GPFittingGaussian(
        type_kernel, training_data, dimensions_kernel, bounds_domain=bounds_domain,
        type_bounds=type_bounds, thinning=thinning, n_burning=n_burning,
        max_steps_out=max_steps_out, random_seed=random_seed, noise=noise,
        simplex_domain=simplex_domain, problem_name=problem_name or 'user_problem', device=device,
        n_chains=n_chains, **kernel_parameters)

    if mle:  # GPFittingService.get_gp(..., mle=True): maximum-likelihood parameter values
        gp_model.fit_gp_regression(random_seed=random_seed)

    quadrature = None
    if name_method == SBO_METHOD:
        quadrature = BayesianQuadrature(gp_model, x_domain, distribution,
                                        parameters_distribution=parameters_distribution)
        acquisition_function = SBO(quadrature)
        acquisition_function.distributed = bool(distributed)
    else:
        acquisition_function = EI(gp_model, noisy_evaluations=noise)

    opt_params_mc = {}
    if factr_mc is not None:
        opt_params_mc['factr'] = factr_mc
    if maxiter_mc is not None:
        opt_params_mc['maxiter'] = maxiter_mc

    for iteration in range(n_iterations):
        if name_method == SBO_METHOD:
            new_point_sol = acquisition_function.optimize(
                parallel=parallel, monte_carlo=monte_carlo_sbo, n_samples=n_samples_mc,
                n_restarts_mc=n_restarts_mc, n_best_restarts_mc=n_best_restarts_mc,
                n_restarts=n_restarts, n_best_restarts=n_best_restarts,
                n_samples_parameters=n_samples_parameters, start_new_chain=True, maxepoch=maxepoch,
                default_n_samples_parameters=default_n_samples_parameters,
                default_n_samples=default_n_samples, method_opt_mc=method_opt_mc, **opt_params_mc)
        else:
            gp_model.start_new_chain()
            gp_model.sample_parameters(DEFAULT_N_PARAMETERS)
            new_point_sol = acquisition_function.optimize(
                parallel=parallel, n_restarts=n_restarts, n_best_restarts=n_best_restarts,
                n_samples_parameters=n_samples_parameters, maxepoch=maxepoch)
        new_point = np.asarray(new_point_sol['solution'], dtype=float)
        v, s2 = _evaluate(training_function, new_point, noise, n_samples_noise)
        gp_model.add_points_evaluations(new_point.reshape((1, len(new_point))), np.array([v]),
                                        var_noise_eval=None if s2 is None else np.array([s2]))
        acquisition_function.clean_cache()

    # maximise the posterior mean of the objective (bayesian_global_optimization.py:365-372)
    model = quadrature if quadrature is not None else gp_model
    optimize_mean = model.optimize_posterior_mean(
        minimize=False, n_restarts=n_restarts_mean, n_best_restarts=n_best_restarts_mean,
        n_samples_parameters=n_samples_parameters_mean, start_new_chain=True,
        maxepoch=maxepoch_mean)
    optimal_value = _evaluate(objective_function, optimize_mean['solution'][0:dim_x], noise,
                              n_samples_noise)[0]
    return {'optimal_solution': np.asarray(optimize_mean['solution'])[0:dim_x],
            'optimal_value': optimal_value}
