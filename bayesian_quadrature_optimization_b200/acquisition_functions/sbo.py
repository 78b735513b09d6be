"""SBO: host-side mirror of sbo/acquisition_functions/sbo.py (the BQO acquisition function).

Every (candidate, hyper-sample, Z sample) tuple is evaluated on the GPU: stage B (gamma,
Sigma^-1 gamma, denominator), the inner maximisation of a(x) + Z b(x) over x, the gradient of b
with respect to the candidate at the maximisers, and the Monte-Carlo means.
"""
from __future__ import annotations

import itertools
from copy import deepcopy

import numpy as np
import torch
from scipy.stats import norm

from ..engine import kg_discretized
from ..lib.constant import UNIFORM_FINITE, LBFGS_NAME, DEFAULT_N_PARAMETERS, DEFAULT_N_SAMPLES
from ..services.domain import DomainService


class SBO(object):

    def __init__(self, bayesian_quadrature, discretization_domain=None):
        self.bq = bayesian_quadrature
        self.discretization = discretization_domain
        self.bounds_opt = deepcopy(self.bq.bounds)
        if self.bq.separate_tasks and not self.bq.task_continue:
            self.bounds_opt.append([None, None])
        elif self.bq.separate_tasks:
            for i in self.bq.w_domain:
                bound = self.bounds_opt[i]
                self.bounds_opt[i] = [bound[0], bound[-1]]
        self.opt_separing_domain = self.bq.distribution == UNIFORM_FINITE
        self.domain_w = [self.bq.gp.bounds[i] for i in self.bq.w_domain]
        self.optimization_results = []
        self.samples = None
        self.starting_points_sbo = None
        self.optimal_samples = {}
        self.mc_bayesian = {}
        self.args_handler = ()
        # device-side inner maximiser settings (bgo defaults: sbo/services/bgo.py:43-44)
        self.inner_pgtol = 1e-5
        self.inner_max_ls = 20
        self.last_n_evals = None
        # multi-GPU (replaces the fork/pickle pool of sbo/lib/parallel.py:16-76): with
        # torch.distributed initialised (one process per GPU, every process running the same
        # program with the same seeds) and `distributed = True`, the hyper-samples are sharded
        # over the ranks -- each GPU factorises and scores only its slice -- and the per-candidate
        # partial sums meet in one all_gather + rank-ordered sum per acquisition evaluation
        self.distributed = False

    # -- one sample path ------------------------------------------------------------------------------
    def evaluate_sample(self, point, candidate_point, sample, var_noise=None, mean=None,
                        parameters_kernel=None, cache=True, n_threads=0, clear_cache=True, w_set=0):
        """:140-165 -> float: a_{n+1}(point) for the standard normal draw `sample`."""
        point = np.asarray(point, dtype=float)
        if len(point.shape) == 1:
            point = point.reshape((1, len(point)))
        vectors = self.bq.compute_parameters_for_sample(
            point, candidate_point, var_noise=var_noise, mean=mean,
            parameters_kernel=parameters_kernel, w_set=w_set)
        value = vectors['a'] + sample * vectors['b']
        return value[0, 0]

    def evaluate_gradient_sample(self, point, candidate_point, sample, var_noise=None, mean=None,
                                 parameters_kernel=None, cache=True, w_set=0):
        """:167-194 -> np.array(n)."""
        point = np.asarray(point, dtype=float)
        if len(point.shape) == 1:
            point = point.reshape((1, len(point)))
        g = self.bq.compute_gradient_parameters_for_sample(
            point, candidate_point, var_noise=var_noise, mean=mean,
            parameters_kernel=parameters_kernel, w_set=w_set)
        return g['a'] + sample * g['b']

    def evaluate_hessian_sample(self, point, candidate_point, sample, var_noise=None, mean=None,
                                parameters_kernel=None, cache=True, w_set=0):
        """:196-235 -> np.array(n x n): Hessian of a + Z b, with the reference's jitter on a
        slightly negative diagonal."""
        point = np.asarray(point, dtype=float)
        if len(point.shape) == 1:
            point = point.reshape((1, len(point)))
        h = self.bq.compute_hessian_parameters_for_sample(
            point, candidate_point, var_noise=var_noise, mean=mean,
            parameters_kernel=parameters_kernel, w_set=w_set)
        hessian = h['a'] + sample * h['b']
        diag_cov = np.diag(hessian)
        if np.any(diag_cov < 0.) and np.min(diag_cov) > -1e-6:
            max_tries, n_tries = 6, 0
            jitter = min(np.mean(diag_cov[np.where(diag_cov < 0)]) * 1e-6, 1e-6)
            while np.any(diag_cov < 0.) and n_tries < max_tries and np.isfinite(jitter):
                hessian += np.eye(hessian.shape[0]) * jitter
                diag_cov = np.diag(hessian)
                n_tries += 1
                jitter *= 10
        return hessian

    def _bounds_x(self):
        return [self.bounds_opt[i] for i in range(len(self.bounds_opt)) if i in self.bq.x_domain]

    def generate_starting_points_evaluate_mc(self, n_restarts, cache=True):
        """:378-399."""
        bounds_x = self._bounds_x()
        start = np.array(DomainService.get_points_domain(n_restarts + 1, bounds_x,
                                                         type_bounds=len(bounds_x) * [0]))
        if cache:
            self.starting_points_sbo = start
        return start

    def generate_samples_starting_points_evaluate_mc(self, n_samples, n_restarts, cache=True):
        """:368-376: Z samples first, then the start points (RNG call order of the reference)."""
        samples = np.random.normal(0, 1, n_samples)
        if cache:
            self.samples = samples
        start = self.generate_starting_points_evaluate_mc(n_restarts, cache=cache)
        return samples, start

    def evaluate_sbo_by_sample(self, candidate_point, sample, start=None, var_noise=None, mean=None,
                               parameters_kernel=None, n_restarts=5, parallel=True, n_threads=0,
                               method_opt=None, tol=None, **opt_params_mc):
        """:237-366 -> {'max': float, 'optimum': np.array(n)}: max_x a(x) + Z b(x) from
        n_restarts + 1 starts (the previous posterior-mean optimum first, when cached) and the box
        vertices, all on the device."""
        theta = self.bq._theta(var_noise, mean, parameters_kernel)
        bounds_x = self._bounds_x()
        index_cache = 'mc_mean' if var_noise is None else (theta[0], theta[1], tuple(theta[2:]))
        if start is None:
            start_points = DomainService.get_points_domain(n_restarts + 1, bounds_x,
                                                           type_bounds=len(bounds_x) * [0])
            prev = self.bq.optimal_solutions.get(index_cache)
            if prev is not None and len(prev) > 0:
                start = np.array([prev[-1]['solution']] + start_points[0:-1])
            else:
                start = np.array(start_points)
        else:
            start = np.asarray(start, dtype=float).reshape(1, -1)
        bq = self.bq.engine_for(theta)
        ub = bq.setup(np.asarray(candidate_point, dtype=float).reshape(1, -1))
        mx, arg, _ = bq.inner_maximize(
            ub, np.array([sample], dtype=float), start, maxiter=opt_params_mc.get('maxiter', 15000),
            factr=opt_params_mc.get('factr', 1e7), pgtol=self.inner_pgtol, max_ls=self.inner_max_ls,
            use_vertices=len(bounds_x) < 7)
        return {'max': float(mx.cpu().numpy()[0, 0]), 'optimum': arg.cpu().numpy()[0, 0]}

    # -- one-sample stochastic gradient (a35) ------------------------------------------------------------
    def evaluate_gradient_given_sample_given_parameters(
            self, candidate_point, sample, parameters=None, n_restarts=10, n_best_restarts=0,
            n_threads=0, parallel=True, method_opt=None, **opt_params_mc):
        """:964-1045 -> np.array(1 x d) or the object np.nan: Z * grad_c b(x*; c) at the maximiser
        x* of a + Z b for this Z and these hyper-parameters (device maximiser from the cached
        starting points, then gradient_vector_b at x*)."""
        gp = self.bq.gp
        if parameters is None:
            var_noise, mean = gp.var_noise.value[0], gp.mean.value[0]
            parameters_kernel = gp.kernel.hypers_values_as_array
        else:
            var_noise, mean, parameters_kernel = parameters[0], parameters[1], parameters[2:]
        if self.starting_points_sbo is None:
            self.generate_starting_points_evaluate_mc(n_restarts)
        theta = self.bq._theta(var_noise, mean, parameters_kernel)
        bq = self.bq.engine_for(theta)
        candidate_point = np.asarray(candidate_point, dtype=float).reshape(1, -1)
        ub = bq.setup(candidate_point)
        _, arg, _ = bq.inner_maximize(
            ub, np.array([sample], dtype=float), np.asarray(self.starting_points_sbo, dtype=float),
            maxiter=opt_params_mc.get('maxiter', 15000), factr=opt_params_mc.get('factr', 1e7),
            pgtol=self.inner_pgtol, max_ls=self.inner_max_ls,
            use_vertices=len(self._bounds_x()) < 7)
        maximum = arg.cpu().numpy()[0, 0].reshape(1, -1)
        gradient_b = self.bq.gradient_vector_b(candidate_point, maximum, var_noise=var_noise,
                                               mean=mean, parameters_kernel=parameters_kernel)
        if gradient_b is np.nan:
            return np.nan
        return gradient_b * sample

    def grad_voi_sgd(self, point, monte_carlo, bayesian, n_restarts=10, n_best_restarts=0,
                     n_threads=0, parallel=True, method_opt=None, random_seed=None,
                     **opt_params_mc):
        """:1103-1142 -> np.array(d) or the object np.nan: one slice sample of the
        hyper-parameters (if bayesian), one Z ~ N(0, 1) (if monte_carlo), then the gradient above.
        SBO.optimize averages many of these per epoch on the device instead."""
        if random_seed is not None:
            np.random.seed(random_seed)
        point = np.asarray(point, dtype=float).reshape((1, -1))
        params = self.bq.gp.sample_parameters(1)[0] if bayesian else None
        sample = np.random.normal(0, 1, 1)[0] if monte_carlo else None
        if sample is None:
            raise NotImplementedError("the discretised (non Monte-Carlo) SGD gradient is "
                                      "SBO.evaluate_gradient")
        grad = self.evaluate_gradient_given_sample_given_parameters(
            point, sample, params, n_restarts, n_best_restarts, n_threads, parallel,
            method_opt=method_opt, **opt_params_mc)
        if grad is np.nan:
            return grad
        return grad[0, :]

    # -- Monte-Carlo Bayesian estimate over many candidates -----------------------------------------------
    def _max_posterior_mean(self, bq, parameters, n_restarts=100, always_draw=False):
        """max_x E[G(x)] per hyper-sample (optimize_posterior_mean, bayesian_quadrature.py:693-910,
        called from sbo.py:597-620): the same device maximiser with Z = 0, i.e. the objective a(x).
        `bq` holds exactly the hyper-samples of `parameters`."""
        missing = [k for k, p in enumerate(parameters)
                   if (p[0], p[1], tuple(p[2:])) not in self.bq.max_mean]
        if missing or always_draw:
            bounds_x = self._bounds_x()
            # (always_draw: ranks with nothing missing still consume the same random numbers)
            starts = np.array(DomainService.get_points_domain(
                n_restarts + 1, bounds_x, type_bounds=len(bounds_x) * [0]))
        if missing:
            ub = bq.setup(self.bq._dummy_candidate())
            mx, arg, _ = bq.inner_maximize(ub, np.zeros(1), starts, maxiter=15000, factr=1e7,
                                           pgtol=self.inner_pgtol, max_ls=self.inner_max_ls,
                                           use_vertices=len(bounds_x) < 7)
            mx = mx.cpu().numpy()[:, 0]
            arg = arg.cpu().numpy()[:, 0]
            for k in missing:
                p = parameters[k]
                index_cache = (p[0], p[1], tuple(p[2:]))
                self.bq.max_mean[index_cache] = mx[k]
                self.bq.optimal_solutions.setdefault(index_cache, []).append(
                    {'solution': arg[k], 'optimal_value': mx[k]})
        return np.array([self.bq.max_mean[(p[0], p[1], tuple(p[2:]))] for p in parameters])

    def _world(self):
        from .. import distributed as D
        rank, ws = D.world()
        return (rank, ws) if self.distributed else (0, 1)

    def _evaluate_candidates(self, parameters, candidate_points, samples, start, compute_max_mean,
                             compute_gradient, hyper_indices=None, local_only=False, **opt):
        """value [C] and gradient [C][dim_m] (device tensors) of the Monte-Carlo acquisition for
        the hyper-samples `parameters` (all of them, or those at `hyper_indices`).  Single GPU: one
        pass over every hyper-sample.  Distributed: this rank's slice of `parameters` only, then
        one all_gather + rank-ordered sum of the partial sums (bit-identical on all ranks)."""
        from .. import distributed as D
        rank, ws = (0, 1) if local_only else self._world()
        S = len(parameters)
        lo, hi = D.shard_range(S, rank, ws) if ws > 1 else (0, S)
        if ws > 1 and hi == lo:
            raise ValueError("distributed SBO needs at least one hyper-sample per rank")
        local = parameters[lo:hi]
        bq = self.bq.engine_for(np.array(local))
        max_mean = None
        if compute_max_mean:
            max_mean = self._max_posterior_mean(bq, local, always_draw=ws > 1)
        else:
            mm = [self.bq.max_mean.get((p[0], p[1], tuple(p[2:])), 0) for p in local]
            if any(m != 0 for m in mm):
                max_mean = np.array(mm, dtype=float)
        n_used = S
        idx = None
        if hyper_indices is not None:
            n_used = len(hyper_indices)
            idx = [int(k) - lo for k in hyper_indices if lo <= int(k) < hi]
            if max_mean is not None:
                max_mean = max_mean[idx]
        dim_m = bq.eng.desc.dim_m
        Cn = int(np.asarray(candidate_points.shape[0]))
        if idx is not None and len(idx) == 0:      # none of the drawn hyper-samples lives here
            value = torch.zeros(Cn, dtype=torch.float64, device=bq.eng.device)
            grad = torch.zeros(Cn, dim_m, dtype=torch.float64, device=bq.eng.device)
            n_loc = 0
        else:
            res = bq.evaluate_candidates(
                candidate_points, samples, start, max_mean=max_mean,
                compute_gradient=compute_gradient, hyper_indices=idx,
                maxiter=opt.get('maxiter', 15000), factr=opt.get('factr', 1e7),
                pgtol=self.inner_pgtol, max_ls=self.inner_max_ls,
                use_vertices=len(self._bounds_x()) < 7)
            value, grad = res['evaluations'], res['gradient']
            n_loc = (hi - lo) if idx is None else len(idx)
        if ws == 1:
            return value, grad
        parts = [value.reshape(-1, 1) * float(n_loc)]
        if compute_gradient:
            parts.append(grad * float(n_loc))
        tot = D.allgather_ordered_sum(torch.cat(parts, dim=1).contiguous()) / float(n_used)
        return tot[:, 0].contiguous(), (tot[:, 1:].contiguous() if compute_gradient else None)

    def evaluate_mc_bayesian_candidate_points_no_restarts(
            self, candidate_points, n_samples_parameters, n_samples, n_restarts=10, n_threads=0,
            compute_max_mean=False, compute_gradient=False, method_opt=None, **opt_params_mc):
        """:523-670 -> {'evaluations': np.array(n), 'gradient': np.array(n x k)}.

        Same argument meaning as the reference; the C x S x Z inner maximisations the reference
        farms out to a process pool run as one kernel launch (per GPU when `distributed`)."""
        gp = self.bq.gp
        if n_samples_parameters == 0:
            parameters = [gp.get_value_parameters_model]
        else:
            parameters = gp.samples_parameters[-n_samples_parameters:]
        samples, start = self.generate_samples_starting_points_evaluate_mc(
            n_samples, n_restarts, cache=False)
        candidate_points = np.asarray(candidate_points, dtype=float)
        value, grad = self._evaluate_candidates(parameters, candidate_points, samples, start,
                                                compute_max_mean, compute_gradient, **opt_params_mc)
        evaluations = value.cpu().numpy()
        gradient = None
        if compute_gradient:
            g = grad.cpu().numpy()
            pad = candidate_points.shape[1] - g.shape[1]
            if pad > 0:  # task coordinate: zero gradient (NaN rows stay NaN)
                extra = np.where(np.isnan(g[:, :1]), np.nan, 0.0)
                g = np.concatenate([g] + pad * [extra], axis=1)
            gradient = g
        return {'evaluations': evaluations, 'gradient': gradient}

    # -- discretised SBO (closed-form knowledge gradient) -------------------------------------------------
    def evaluate(self, point, var_noise=None, mean=None, parameters_kernel=None, cache=True,
                 n_threads=0):
        """:1394-1425 -> float."""
        vectors = self.bq.compute_posterior_parameters_kg(
            self.discretization, point, var_noise=var_noise, mean=mean,
            parameters_kernel=parameters_kernel)
        a, b = vectors['a'], vectors['b']
        if not np.all(np.isfinite(b)):
            return 0.0
        kg, _, _, _ = kg_discretized(self.bq.gp._engine(), a, b[None, :])
        return float(kg.cpu().numpy()[0])

    def evaluate_gradient(self, point, var_noise=None, mean=None, parameters_kernel=None,
                          cache=True, n_threads=0):
        """:1428-1474 -> np.array(n)."""
        point = np.asarray(point, dtype=float).reshape(1, -1)
        vectors = self.bq.compute_posterior_parameters_kg(
            self.discretization, point, var_noise=var_noise, mean=mean,
            parameters_kernel=parameters_kernel)
        a, b = vectors['a'], vectors['b']
        _, cnt, idx, c_out = kg_discretized(self.bq.gp._engine(), a, b[None, :])
        M = int(cnt.cpu().numpy()[0])
        if M <= 1:
            return np.zeros(point.shape[1])
        keep = idx.cpu().numpy()[0, :M]
        c2 = np.abs(c_out.cpu().numpy()[0, :M - 1])
        evalC = norm.pdf(c2)
        gradients = self.bq.gradient_vector_b(point, self.discretization[keep, :],
                                              var_noise=var_noise, mean=mean,
                                              parameters_kernel=parameters_kernel)
        gradient = np.zeros(point.shape[1])
        for i in range(point.shape[1]):
            gradient[i] = np.dot(np.diff(gradients[:, i]), evalC)
        return gradient

    @staticmethod
    def hvoi(b, c, keep):
        """:1952-1961 (host helper kept for API parity)."""
        M = len(keep)
        if M > 1:
            c = c[keep + 1]
            c2 = -np.abs(c[0:M - 1])
            tmp = norm.pdf(c2) + c2 * norm.cdf(c2)
            return np.sum(np.diff(b[keep]) * tmp)
        return 0

    # -- maximisation of the acquisition function: multi-start SGD on the device -------------------------
    def _restart_points(self, n_restarts):
        """Start points of SBO.optimize (:1688-1765): 100 random points, squared distance to the
        data, draws from the (1/8,1/4) and (1/4,1/2) distance-quantile bands, replicated over the
        tasks when W is finite.  Same numpy RNG calls as the reference."""
        bq = self.bq
        bounds = bq.bounds
        finite = bq.separate_tasks and not bq.task_continue
        # :1737-1740: the simplex constraint (citi_bike_mt) applies to the non-finite branch only
        new_points = np.array(DomainService.get_points_domain(
            100, bounds, type_bounds=bq.type_bounds,
            simplex_domain=None if finite else bq.simplex_domain))
        current = np.array(bq.gp.data['points'])
        if finite:
            current = current[:, 0:-1]
        d = new_points[:, None, :] - current[None, :, :new_points.shape[1]]
        max_distances = np.min(np.einsum('ijk,ijk->ij', d, d), axis=1)
        sort_dist_ind = sorted(range(len(max_distances)), key=lambda k: max_distances[k])
        md = int(np.ceil(len(max_distances) / 2.0))
        uq = int(np.ceil(len(max_distances) / 4.0))
        lw = int(np.ceil(len(max_distances) / 8.0))
        index_1 = n_restarts // 2
        index_2 = n_restarts - index_1
        index_1 = np.random.choice(range(uq, md), index_1, replace=False)
        index_2 = np.random.choice(range(lw, uq), index_2, replace=False)
        index = [sort_dist_ind[t] for t in index_1] + [sort_dist_ind[t] for t in index_2]
        if finite:
            n_tasks = len(bq.tasks)
            return np.array([np.concatenate((new_points[i], [j])) for i in index
                             for j in range(n_tasks)])
        return np.array([new_points[i] for i in index])

    def optimize(self, start=None, random_seed=None, parallel=True, monte_carlo=False, n_samples=1,
                 n_restarts_mc=1, n_best_restarts_mc=0, n_restarts=1, n_best_restarts=0,
                 start_ei=True, n_samples_parameters=0, start_new_chain=True,
                 compute_max_mean_bayesian=False, maxepoch=10, default_n_samples=None,
                 default_n_samples_parameters=None, default_restarts_mc=None, method_opt_mc=None,
                 advance_chain=False, **opt_params_mc):
        """:1571-1950 -> {'solution', 'optimal_value', 'gradient'} of the best restart.

        Monte-Carlo / Bayesian branch on the device.  Every SGD epoch is ONE pass of the hot path:
        the stochastic gradient of all restarts is the mean over n_samples_parameters resident
        hyper-samples (drawn uniformly from the last n_parameters slice samples; the reference
        advances the chain by one slice step per gradient, sbo.py:1129) times n_samples fresh
        Z ~ N(0,1), i.e. N = n_samples * n_samples_parameters single-sample gradients
        (sbo.py:1845) averaged, followed by the reference's Adam step, box projection and stop
        rule (sbo/lib/stochastic_gradient_descent.py:83-130) for every restart at once (a restart
        that has met its stop rule is frozen; the loop ends when all have).

        advance_chain=True restores the reference's source of hyper-samples: every epoch advances
        the slice-sampling chain by n_samples_parameters new samples (gp.sample_parameters, as
        grad_voi_sgd does at sbo.py:1129) and takes the gradient at those, at the price of the
        sampler's sweeps and one factorisation per new sample and epoch.  The default draws from
        the resident posterior samples -- the set the final scoring of the end points averages
        over anyway (:1884-1900); tests/test_gpu_facade.py compares the two."""
        import torch
        from .. import _cabi
        from .ei import EI
        from .multi_task import MultiTasks
        from ..numerical_tools.bayesian_quadrature import BayesianQuadrature
        from ..lib.constant import TASKS
        if random_seed is not None:
            np.random.seed(random_seed)
        bq, gp = self.bq, self.bq.gp
        n_parameters = 0 if n_samples_parameters == 0 else (
            default_n_samples_parameters or DEFAULT_N_PARAMETERS)
        if default_n_samples is None:
            default_n_samples = DEFAULT_N_SAMPLES
        if default_restarts_mc is None:
            default_restarts_mc = n_restarts_mc
        if n_samples_parameters > 0 and start_new_chain:
            gp.start_new_chain()
            gp.sample_parameters(n_parameters)
        elif n_samples_parameters > 0 and len(gp.samples_parameters) < n_parameters:
            gp.sample_parameters(n_parameters - len(gp.samples_parameters))

        st_ei = None
        if start_ei and not bq.separate_tasks:
            opt_ei = EI(gp).optimize(n_restarts=100, n_samples_parameters=5, parallel=parallel,
                                     n_best_restarts=10, maxepoch=50)
            st_ei = opt_ei['solution'].reshape((1, -1))
            n_restarts -= 1
            n_best_restarts -= 1
        elif start_ei:
            quadrature_2 = BayesianQuadrature(gp, bq.x_domain, bq.distribution,
                                              parameters_distribution=bq.parameters_distribution,
                                              model_only_x=True)
            mk = MultiTasks(quadrature_2, bq.n_tasks)
            st_ei = mk.optimize(parallel=True, n_restarts=100, n_best_restarts=10,
                                n_samples_parameters=5, maxepoch=50,
                                start_new_chain=False)['solution'].reshape((1, -1))
        if start is None:
            if n_restarts > 0:
                start = self._restart_points(n_restarts)
                if 0 < n_best_restarts < start.shape[0]:
                    out = self.evaluate_mc_bayesian_candidate_points_no_restarts(
                        start, n_parameters, default_n_samples, default_restarts_mc,
                        compute_max_mean=True, compute_gradient=False, **opt_params_mc)
                    keep = np.argsort(out['evaluations'], kind='stable')[-n_best_restarts:]
                    start = start[keep]
                if st_ei is not None:
                    start = np.concatenate((start, st_ei), axis=0)
            else:
                start = st_ei
        start = np.asarray(start, dtype=float).reshape(-1, len(self.bounds_opt))

        if n_samples_parameters == 0 and not monte_carlo:
            return self._optimize_discretised(start)

        # ---- batched Adam over all restarts --------------------------------------------------------
        parameters = gp.samples_parameters[-n_parameters:] if n_parameters > 0 else \
            [gp.get_value_parameters_model]
        from .. import distributed as D
        rank, ws = self._world()
        lo_s, hi_s = D.shard_range(len(parameters), rank, ws) if ws > 1 else (0, len(parameters))
        bqe = bq.engine_for(np.array(parameters[lo_s:hi_s]))
        eng = bqe.eng
        dim = start.shape[1]
        dim_m = eng.desc.dim_m
        lo = np.array([(-np.inf if b[0] is None else b[0]) for b in self.bounds_opt], dtype=float)
        up = np.array([(np.inf if b[-1] is None else b[-1]) for b in self.bounds_opt], dtype=float)
        pts = eng.to_device(start.copy())
        R = int(pts.shape[0])
        m0, v0 = torch.zeros_like(pts), torch.zeros_like(pts)
        active = torch.ones(R, dtype=torch.int32, device=eng.device)
        lo_d, up_d = eng.to_device(lo), eng.to_device(up)
        n_k = max(1, min(n_samples_parameters if n_samples_parameters > 0 else 1, len(parameters)))
        n_zs = max(1, n_samples)
        bx = self._bounds_x()
        for t in range(1, maxepoch + 1):
            fresh = None
            if advance_chain and n_samples_parameters > 0:
                fresh = [np.asarray(p, dtype=float) for p in gp.sample_parameters(n_k)[-n_k:]]
            else:
                ks = np.random.choice(len(parameters), n_k, replace=False)
            zs = np.random.normal(0, 1, n_zs)
            starts_x = np.array(DomainService.get_points_domain(
                default_restarts_mc + 1, bx, type_bounds=len(bx) * [0]))
            if fresh is not None:
                # every rank advances the same chain (same seeds) and scores the new samples itself
                _, grad_t = self._evaluate_candidates(fresh, pts, zs, starts_x, False, True,
                                                      local_only=True, **opt_params_mc)
            else:
                _, grad_t = self._evaluate_candidates(parameters, pts, zs, starts_x, False, True,
                                                      hyper_indices=ks, **opt_params_mc)
            g = torch.zeros(R, dim, dtype=torch.float64, device=eng.device)
            g[:, :dim_m] = -grad_t                        # maximise: descend on -acquisition
            bad = torch.isnan(g).any(dim=1)
            if bool(bad.any()):
                # gradient unavailable (den = 0): uniform perturbation of relative size 1e-6 inside
                # the box (stochastic_gradient_descent.py:51-75), no Adam step for those restarts.
                # Only the continuous coordinates move -- a task id must stay an integer -- and
                # the draws come from numpy so that random_seed fixes the run.
                rows = np.nonzero(bad.cpu().numpy())[0]
                host = pts[bad].cpu().numpy()
                for r in range(host.shape[0]):
                    size = np.sqrt(np.sum(host[r] ** 2)) * 1e-6
                    for i in range(dim_m):
                        lb = min(size, host[r, i] - lo[i])
                        ub = min(size, up[i] - host[r, i])
                        host[r, i] += np.random.uniform(-lb, ub)
                pts[torch.as_tensor(rows, device=eng.device)] = eng.to_device(host)
                g = torch.where(bad.unsqueeze(1), torch.zeros_like(g), g)
            _cabi.check(eng.lib.bqo_adam_step_batched(
                _cabi.ptr(pts), _cabi.ptr(g.contiguous()), _cabi.ptr(m0), _cabi.ptr(v0),
                _cabi.ptr(active), _cabi.ptr(lo_d), _cabi.ptr(up_d), R, dim, t, 0.1, 0.9, 0.999, 1e-8,
                torch.cuda.current_stream().cuda_stream), "bqo_adam_step_batched")
            if bq.simplex_domain is not None and dim > 2:
                # stochastic_gradient_descent.py:112-114: points whose coordinates from the third
                # on sum to more than the simplex bound are scaled back onto it
                tail = pts[:, 2:]
                tot = tail.sum(dim=1, keepdim=True)
                over = tot > float(bq.simplex_domain)
                pts[:, 2:] = torch.where(over, tail * (float(bq.simplex_domain) / tot), tail)
            if int(active.sum().item()) == 0:
                break
        candidate_points = pts.cpu().numpy()

        # ---- final scoring of the end points (:1884-1900) ---------------------------------------------
        output = self.evaluate_mc_bayesian_candidate_points_no_restarts(
            candidate_points, n_parameters, default_n_samples, default_restarts_mc,
            compute_max_mean=True, compute_gradient=True, **opt_params_mc)
        ind_max = int(np.argmax(output['evaluations']))
        solution = candidate_points[ind_max].copy()
        gradient = output['gradient'][ind_max]
        if np.any(np.isnan(gradient)):
            # :1915-1946: perturb the solution when its gradient is unavailable
            npt = solution[:-1] if (not bq.task_continue and bq.separate_tasks) else solution
            pert = np.sqrt(np.sum(npt ** 2)) * 1e-6
            for i in range(len(npt)):
                lb = min(pert, npt[i] - lo[i])
                ub = min(pert, up[i] - npt[i])
                npt[i] += np.random.uniform(-lb, ub)
        result = {'solution': solution, 'optimal_value': float(output['evaluations'][ind_max]),
                  'gradient': gradient}
        self.optimization_results.append(result)
        return result

    def _optimize_discretised(self, start):
        """The branch `n_samples_parameters == 0 and not monte_carlo` of SBO.optimize (:1800-1821):
        the discretised SBO (closed-form knowledge gradient over `discretization_domain`, model
        parameters at their current values) is maximised from every start point with L-BFGS-B,
        maxiter = 10 (Optimization(LBFGS_NAME, wrapper_objective_voi, bounds, wrapper_gradient_voi,
        minimize=False, maxiter=10), sbo/lib/optimization.py:87-155); scipy drives, every value and
        gradient comes from the device (kg_kernel, gradient_vector_b).  Returns the best restart
        with the reference's result keys."""
        from scipy.optimize import fmin_l_bfgs_b
        if self.discretization is None:
            raise ValueError("the discretised SBO needs a discretization_domain")
        bounds = [tuple(bound) for bound in self.bounds_opt]
        results = []
        for x0 in start:
            opt = fmin_l_bfgs_b(
                lambda x: -1.0 * self.evaluate(x.reshape((1, len(x)))), np.array(x0, dtype=float),
                fprime=lambda x: -1.0 * self.evaluate_gradient(x.reshape((1, len(x)))),
                bounds=bounds, maxiter=10)
            results.append({'solution': opt[0], 'optimal_value': -1.0 * opt[1],
                            'gradient': -1.0 * opt[2]['grad'], 'warnflag': opt[2]['warnflag'],
                            'task': opt[2]['task'], 'nit': opt[2]['nit'],
                            'funcalls': opt[2]['funcalls']})
        ind_max = int(np.argmax([r['optimal_value'] for r in results]))
        result = results[ind_max]
        if np.any(np.isnan(result['gradient'])):
            # :1915-1946: perturb the solution when its gradient is unavailable
            bq = self.bq
            finite = not bq.task_continue and bq.separate_tasks
            point = result['solution'][:-1] if finite else result['solution']
            size = np.sqrt(np.sum(point ** 2)) * 1e-6
            for i in range(len(point)):
                lb = min(size, point[i] - bounds[i][0])
                ub = min(size, bounds[i][1] - point[i])
                point[i] += np.random.uniform(-lb, ub)
            result['gradient'] = 'unavailable'
        self.optimization_results.append(result)
        return result

    def clean_cache(self):
        self.samples = None
        self.starting_points_sbo = None
        self.optimal_samples = {}
        self.mc_bayesian = {}
        self.bq.clean_cache()
