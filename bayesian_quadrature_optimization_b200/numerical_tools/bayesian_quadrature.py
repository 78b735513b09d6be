"""BayesianQuadrature: host-side mirror of sbo/numerical_tools/bayesian_quadrature.py.

GP on G(x) = E_W[F(x, W)].  The expectation over W is a finite (optionally weighted) sum over
tasks or a Monte-Carlo mean over M = 10 W samples.  The reference draws fresh `gamma.rvs` inside
every call (sbo/lib/expectations.py:105); here the samples are drawn ONCE per object with the
same scipy call (`resample_w`), kept as `w_sets[n_wsets][M][d_w]`, uploaded to the GPU, and every
evaluation names the set it uses -- the explicit-randomness contract of SURVEY.md section 7.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
from scipy.stats import gamma

from .. import _cabi
from ..engine import BQEngine
from ..lib.constant import (
    UNIFORM_FINITE, WEIGHTED_UNIFORM_FINITE, GAMMA, EXPONENTIAL, TASKS, BAYESIAN_QUADRATURE,
    N_SAMPLES_W, DEFAULT_N_PARAMETERS,
)
from ..services.domain import DomainService


class BayesianQuadrature(object):

    def __init__(self, gp_model, x_domain, distribution, parameters_distribution=None,
                 model_only_x=False, tasks=None, n_w_sets=1):
        self.gp = gp_model
        self.simplex_domain = self.gp.simplex_domain
        self.name_model = BAYESIAN_QUADRATURE
        if parameters_distribution == {}:
            parameters_distribution = None
        if parameters_distribution is None and distribution == UNIFORM_FINITE:
            parameters_distribution = {TASKS: self.gp._n_tasks}
        self.parameters_distribution = parameters_distribution
        self.dimension_domain = self.gp.dimension_domain
        self.x_domain = list(x_domain)
        if self.x_domain != list(range(len(self.x_domain))):
            raise ValueError("the accelerated path needs x_domain = [0, ..., d_x - 1]")
        self.w_domain = [i for i in range(self.gp.dimension_domain) if i not in self.x_domain]
        self.distribution = distribution
        self.weights = None
        self.task_continue = False
        if distribution == UNIFORM_FINITE:
            self.n_tasks = self.parameters_distribution.get(TASKS)
        elif distribution == WEIGHTED_UNIFORM_FINITE:
            self.weights = np.array(self.parameters_distribution['weights'], dtype=float)
            self.n_tasks = len(self.weights)
            self.task_continue = True
        elif distribution not in (GAMMA, EXPONENTIAL):
            raise NameError("distribution %s is not on the accelerated path" % distribution)
        finite = distribution in (UNIFORM_FINITE, WEIGHTED_UNIFORM_FINITE)
        self.full_weights = self.weights
        if finite and self.parameters_distribution.get('n_samples'):
            self.resample_tasks(self.parameters_distribution.get('n_samples'))
        if finite != (self.gp._kind == _cabi.KIND_PRODUCT_TASKS):
            raise ValueError("finite W needs the Matern52 x Tasks product kernel and vice versa")

        self.var_noise = None
        if self.gp.noise and self.gp.data.get('var_noise') is not None:
            self.var_noise = np.mean(self.gp.data.get('var_noise'))  # :153-155

        self.separate_tasks = False
        if self.gp.type_bounds != [] and self.gp.type_bounds[-1] == 1 and not model_only_x:
            self.separate_tasks = True
        self.model_only_x = model_only_x
        if model_only_x or self.separate_tasks:
            self.type_bounds = [self.gp.type_bounds[i] for i in self.x_domain]
            self.bounds = [self.gp.bounds[i] for i in self.x_domain]
        else:
            self.type_bounds = list(self.gp.type_bounds)
            self.bounds = [list(b) for b in self.gp.bounds]
        if not model_only_x and self.task_continue:
            self.type_bounds = list(self.gp.type_bounds)
            self.bounds = [list(b) for b in self.gp.bounds]
        self.tasks = self.gp.bounds[-1] if (self.gp.bounds != [] and self.separate_tasks) else []
        self.type_kernel = self.gp.type_kernel

        self.w_sets = None
        self.w_double_sets = None
        if not finite:
            self.resample_w(n_w_sets)
            self.resample_w_double()
        self._bq_cache = OrderedDict()
        self.optimal_solutions = {}
        self.max_mean = {}
        self.best_solution = {}
        self.args_handler = ()

    # -- explicit W samples ------------------------------------------------------------------------
    def resample_w(self, n_w_sets=1, n_samples=N_SAMPLES_W):
        """Draw the W sample sets with the reference's own sampler call
        (gamma.rvs(a, scale=scale, size=(n_samples, dim_w)), sbo/lib/expectations.py:93-105)."""
        a = self.parameters_distribution['a'][0]
        scale = self.parameters_distribution['scale'][0]
        dim_w = len(self.w_domain)
        self.w_sets = np.stack([gamma.rvs(a, scale=scale, size=(n_samples, dim_w))
                                for _ in range(n_w_sets)])
        self._bq_cache = OrderedDict()

    def resample_tasks(self, n_samples):
        """Sub-sampled finite expectation (`n_samples` of parameters_distribution; uniform_finite,
        sbo/lib/expectations.py:37-42): n_samples tasks are drawn with replacement with the task
        probabilities, np.random.choice(T, n_samples, replace=True, p=weights), and the
        expectation becomes their plain average -- i.e. quadrature weights count / n_samples.  The
        reference redraws inside every call; here the draw is made once per object (and again on
        demand), like the W samples of the continuous case."""
        p = None if self.full_weights is None else self.full_weights / np.sum(self.full_weights)
        index = np.random.choice(self.n_tasks, int(n_samples), replace=True, p=p)
        self.task_sample = index
        self.weights = np.bincount(index, minlength=self.n_tasks).astype(float) / float(n_samples)
        self._bq_cache = OrderedDict()

    def set_w_sets(self, w_sets):
        w_sets = np.asarray(w_sets, dtype=float)
        self.w_sets = w_sets[None] if w_sets.ndim == 2 else w_sets
        self._bq_cache = OrderedDict()

    def resample_w_double(self, n_samples=100):
        """The two independent draws of the double expectation (gamma_expect with double=True,
        sbo/lib/expectations.py:93-112: n_samples = 100, two gamma.rvs calls in this order)."""
        a = self.parameters_distribution['a'][0]
        scale = self.parameters_distribution['scale'][0]
        dim_w = len(self.w_domain)
        first = gamma.rvs(a, scale=scale, size=(n_samples, dim_w))
        second = gamma.rvs(a, scale=scale, size=(n_samples, dim_w))
        self.w_double_sets = np.stack([first, second])

    def set_w_double_sets(self, w_double_sets):
        w = np.asarray(w_double_sets, dtype=float)
        if w.ndim != 3 or w.shape[0] != 2:
            raise ValueError("w_double_sets must be [2][n_samples][d_w]")
        self.w_double_sets = w
        self._bq_cache = OrderedDict()

    # -- device objects ------------------------------------------------------------------------------
    def bounds_x(self):
        return [self.gp.bounds[i] for i in self.x_domain]

    def engine_for(self, thetas) -> BQEngine:
        hb = self.gp.hyper_batch(thetas)
        key = id(hb)
        bq = self._bq_cache.get(key)
        if bq is None or bq.hb is not hb:
            hb.factorize(max_tries=7)
            for s in range(hb.S):
                self.gp._raise_info(int(hb.info[s]))
            bx = self.bounds_x()
            lower = np.array([b[0] for b in bx], dtype=float)
            upper = np.array([b[-1] for b in bx], dtype=float)
            bq = BQEngine(hb, len(self.x_domain), weights=self.weights, wsets=self.w_sets,
                          lower=lower, upper=upper, bq_var_noise=self.var_noise,
                          w_double=self.w_double_sets)
            self._bq_cache[key] = bq
            while len(self._bq_cache) > 8:
                self._bq_cache.popitem(last=False)
        # a batch the GP has evicted from its factor cache must not stay alive through this one
        alive = set(id(h) for h in self.gp._hb_cache.values())
        for k in [k for k in self._bq_cache if k not in alive]:
            del self._bq_cache[k]
        return bq

    def _theta(self, var_noise, mean, parameters_kernel):
        var_noise, mean, parameters_kernel = self.gp._defaults(var_noise, mean, parameters_kernel)
        return np.concatenate([[var_noise, mean], parameters_kernel])

    def clean_cache(self):
        """:1579-1588: quadrature caches; the GP keeps its device factors (GPFittingGaussian.clean_cache)."""
        self._bq_cache = OrderedDict()
        self.optimal_solutions = {}
        self.max_mean = {}
        self.best_solution = {}
        self.var_noise = None
        if self.gp.noise and self.gp.data.get('var_noise') is not None:
            self.var_noise = np.mean(self.gp.data.get('var_noise'))
        self.gp.clean_cache(keep_factors=True)

    # -- expectations of the kernel ------------------------------------------------------------------
    def _new_points(self, point, w_set=0):
        """(x, W values): uniform_finite / gamma_expect build the same array
        (sbo/lib/expectations.py:44-50, 100-106)."""
        point = np.asarray(point, dtype=float).reshape(1, -1)
        if self.w_sets is not None:
            random = self.w_sets[w_set]
        else:
            random = np.arange(self.n_tasks, dtype=float).reshape(-1, 1)
        m = random.shape[0]
        new_points = np.zeros((m, len(self.w_domain) + point.shape[1]))
        new_points[:, self.x_domain] = np.repeat(point, m, axis=0)
        new_points[:, self.w_domain] = random
        return new_points

    def _average(self, values):
        """np.average(values, axis=0, weights) / np.mean(values, axis=0) on the device."""
        if self.weights is None:
            return values.mean(dim=0)
        w = torch.as_tensor(self.weights, dtype=values.dtype, device=values.device)
        shape = [-1] + [1] * (values.dim() - 1)
        return (values * w.reshape(shape)).sum(dim=0) / w.sum()

    def evaluate_quadrature_cross_cov(self, point, points_2, parameters_kernel, w_set=0):
        """B(x, z_j): :308-334 -> np.array(m)."""
        eng = self.gp._temp_engine(points_2)
        hb = eng.hypers(np.concatenate([[0.0, 0.0], np.asarray(parameters_kernel, dtype=float)]))
        vals = hb.cross_cov(self._new_points(point, w_set))[0]
        return self._average(vals).cpu().numpy()

    def evaluate_grad_quadrature_cross_cov(self, point, points_2, parameters_kernel, w_set=0):
        """grad_x B(x, z_j): :336-362 -> np.array(k x m)."""
        eng = self.gp._temp_engine(points_2)
        hb = eng.hypers(np.concatenate([[0.0, 0.0], np.asarray(parameters_kernel, dtype=float)]))
        _, dKP = hb.cross_cov(self._new_points(point, w_set), grad=True)
        return self._average(dKP[0])[: len(self.x_domain)].cpu().numpy()

    def evaluate_hessian_cross_cov(self, point, points_2, parameters_kernel, w_set=0):
        """hess_x B(x, z_j): :364-386 with hessian_uniform_finite / hessian_gamma
        (sbo/lib/expectations.py:260-324) -> np.array(m x k x k)."""
        eng = self.gp._temp_engine(points_2)
        hb = eng.hypers(np.concatenate([[0.0, 0.0], np.asarray(parameters_kernel, dtype=float)]))
        H = hb.cross_cov_hessian(self._new_points(point, w_set))[0]   # [M][dm][dm][m]
        dx = len(self.x_domain)
        H = self._average(H)[:dx, :dx]                               # [dx][dx][m]
        return H.permute(2, 0, 1).contiguous().cpu().numpy()

    def evaluate_grad_quadrature_cross_cov_resp_candidate(self, candidate_point, points,
                                                          parameters_kernel, w_set=0):
        """grad_c B(x_p, c): :388-415 -> np.array(k x m)."""
        candidate_point = np.asarray(candidate_point, dtype=float).reshape(1, -1)
        points = np.asarray(points, dtype=float)
        out = np.zeros((candidate_point.shape[1], points.shape[0]))
        for i in range(points.shape[0]):
            new_points = self._new_points(points[i:i + 1, :], w_set)
            eng = self.gp._temp_engine(new_points)
            hb = eng.hypers(np.concatenate([[0.0, 0.0], np.asarray(parameters_kernel, dtype=float)]))
            _, dKP = hb.cross_cov(candidate_point, grad=True)  # [1][1][dim_m][M]
            g = self._average(dKP[0, 0].T).cpu().numpy()
            out[: g.shape[0], i] = g
        return out

    def evaluate_quadrate_cov(self, point, parameters_kernel):
        """E_W E_W' k((x, W), (x, W')): :282-306 -> float.  Finite W: the exact (weighted) double
        sum over tasks.  Gamma / exponential W: the mean of k over the full grid of two independent
        sets of 100 draws (`w_double_sets`, drawn like the reference on first use unless given by
        set_w_double_sets -- the reference redraws them at every call)."""
        point = np.asarray(point, dtype=float).reshape(1, -1)
        if self.w_sets is not None:
            if self.w_double_sets is None:
                self.resample_w_double()
            both = []
            for w in self.w_double_sets:
                pts = np.zeros((w.shape[0], len(self.w_domain) + point.shape[1]))
                pts[:, self.x_domain] = np.repeat(point, w.shape[0], axis=0)
                pts[:, self.w_domain] = w
                both.append(pts)
            vals = torch.as_tensor(self.gp.evaluate_cross_cov(both[0], both[1], parameters_kernel))
            return float(vals.mean())
        p1 = self._new_points(point)
        vals = torch.as_tensor(self.gp.evaluate_cross_cov(p1, p1, parameters_kernel))
        if self.weights is None:
            return float(vals.mean())
        w = torch.as_tensor(self.weights, dtype=vals.dtype)
        return float((vals * torch.outer(w, w)).sum() / (w.sum() ** 2))

    # -- posterior of G ------------------------------------------------------------------------------
    def _dummy_candidate(self):
        return self.gp.data['points'][0:1, :]

    def compute_posterior_parameters(self, points, var_noise=None, mean=None,
                                     parameters_kernel=None, historical_points=None,
                                     historical_evaluations=None, only_mean=False, cache=True,
                                     parallel=False, w_set=0):
        """:449-537 -> {'mean': np.array(t), 'cov': float}."""
        theta = self._theta(var_noise, mean, parameters_kernel)
        points = np.asarray(points, dtype=float).reshape(-1, len(self.x_domain))
        bq = self.engine_for(theta)
        ub = bq.setup(self._dummy_candidate())
        t = points.shape[0]
        out = bq.sample_ab(ub, points, np.zeros(t, dtype=np.int32),
                           np.full(t, w_set, dtype=np.int32) if self.w_sets is not None else None)
        mu_n = out[:, 0].cpu().numpy()
        if only_mean:
            return {'mean': mu_n, 'cov': None}
        B = torch.as_tensor(self.evaluate_quadrature_cross_cov(
            points[0:1, :], self.gp.data['points'], theta[2:], w_set), device=bq.eng.device)
        sol = bq.hb.solve(B.reshape(1, 1, -1).clone())
        cov = self.evaluate_quadrate_cov(points[0:1, :], theta[2:]) - \
            float((B * sol[0, 0]).sum())
        return {'mean': mu_n, 'cov': cov}

    def gradient_posterior_mean(self, point, var_noise=None, mean=None, parameters_kernel=None,
                                historical_points=None, historical_evaluations=None, cache=True,
                                w_set=0):
        """:539-579 -> np.array(k)."""
        theta = self._theta(var_noise, mean, parameters_kernel)
        bq = self.engine_for(theta)
        ub = bq.setup(self._dummy_candidate())
        dx = len(self.x_domain)
        out = bq.sample_ab(ub, np.asarray(point, dtype=float).reshape(1, dx),
                           np.zeros(1, dtype=np.int32),
                           np.full(1, w_set, dtype=np.int32) if self.w_sets is not None else None)
        return out[0, 2:2 + dx].cpu().numpy()

    def gradient_posterior_parameters(self, point, var_noise=None, mean=None,
                                      parameters_kernel=None, cache=True, parallel=True, w_set=0):
        """:581-643 -> {'mean': np.array(d_x), 'cov': np.array(d_x)}."""
        theta = self._theta(var_noise, mean, parameters_kernel)
        bq = self.engine_for(theta)
        point = np.asarray(point, dtype=float).reshape(1, len(self.x_domain))
        _, _, gm, gv = bq.quadrature_posterior_grad(point, w_set)
        return {'mean': gm[0, 0].cpu().numpy(), 'cov': gv[0, 0].cpu().numpy()}

    def objective_posterior_mean(self, point, var_noise=None, mean=None, parameters_kernel=None):
        point = np.asarray(point, dtype=float).reshape((1, -1))
        return self.compute_posterior_parameters(point, var_noise=var_noise, mean=mean,
                                                 parameters_kernel=parameters_kernel,
                                                 only_mean=True)['mean']

    def grad_posterior_mean(self, point, var_noise=None, mean=None, parameters_kernel=None):
        point = np.asarray(point, dtype=float).reshape((1, -1))
        return self.gradient_posterior_mean(point, var_noise=var_noise, mean=mean,
                                            parameters_kernel=parameters_kernel)

    # -- maximisation of the posterior mean of G ----------------------------------------------------------
    def _averaged_mean_and_grad(self, bq, ub, pts):
        """mean over the hyper-samples of (mu_k(x), grad mu_k(x)) for every row of `pts`
        (device tensors): one sample_ab launch over R x S (point, hyper-sample) pairs."""
        S = bq.hb.S
        dx = len(self.x_domain)
        pts = bq.eng.to_device(pts).reshape(-1, dx)
        R = int(pts.shape[0])
        rep = pts.repeat_interleave(S, dim=0).contiguous()
        units = np.tile(np.arange(S, dtype=np.int32), R)
        wset = np.zeros(R * S, dtype=np.int32) if self.w_sets is not None else None
        out = bq.sample_ab(ub, rep, units, wset).reshape(R, S, -1)
        return out[:, :, 0].mean(dim=1), out[:, :, 2:2 + dx].mean(dim=1)

    def optimize_posterior_mean(self, start=None, random_seed=None, minimize=False, n_restarts=1000,
                                n_best_restarts=100, parallel=True, n_treads=0, var_noise=None,
                                mean=None, parameters_kernel=None, n_samples_parameters=0,
                                start_new_chain=False, method_opt=None, maxepoch=10,
                                candidate_solutions=None, candidate_values=None):
        """:693-910 -> {'solution': np.array(d_x), 'optimal_value': np.array([value])} (the
        reference's callers read optimal_value[0], bayesian_global_optimization.py:281)

        Same arguments and bookkeeping (`optimal_solutions`, `max_mean`) as the reference.  With
.

        With fixed hyper-parameters the device quasi-Newton maximiser climbs the posterior mean of G
        (the stage-C objective a(x) with Z = 0) from every start in one launch; with
        n_samples_parameters > 0 the objective is the mean over the last DEFAULT_N_PARAMETERS
        hyper-samples (BayesianEvaluations, sbo/bayesian/bayesian_evaluations.py:9-44) and all
        restarts take `maxepoch` Adam steps together on its exact gradient (the reference feeds
        Adam one slice sample per gradient, :662-676)."""
        import itertools
        from ..lib.batched_ascent import batched_adam_ascent
        if candidate_solutions is not None and len(candidate_solutions) == 0:
            candidate_solutions = None
        if random_seed is not None:
            np.random.seed(random_seed)
        if start_new_chain and n_samples_parameters > 0:
            self.gp.start_new_chain()
            self.gp.sample_parameters(DEFAULT_N_PARAMETERS)
        bounds_x = self.bounds_x()
        dim_x = len(bounds_x)
        vertex = None
        if dim_x < 7 and self.simplex_domain is None:
            vertex = [list(p) for p in itertools.product(*bounds_x)]
        if n_samples_parameters == 0:
            theta = self._theta(var_noise, mean, parameters_kernel)
            thetas = theta.reshape(1, -1)
            index_cache = (theta[0], theta[1], tuple(theta[2:]))
        else:
            thetas = np.array(self.gp.samples_parameters[-DEFAULT_N_PARAMETERS:])
            index_cache = 'mc_mean'
        bq = self.engine_for(thetas)
        ub = bq.setup(self._dummy_candidate())
        sign = -1.0 if minimize else 1.0

        def averaged(p):
            return self._averaged_mean_and_grad(bq, ub, p)

        if start is None:
            start_points = DomainService.get_points_domain(
                n_restarts + 1, bounds_x, type_bounds=len(self.x_domain) * [0], simplex_domain=None)
            prev = self.optimal_solutions.get(index_cache)
            if prev is not None and len(prev) > 0:
                start = np.array([prev[-1]['solution']] + start_points[0:-1])
            else:
                start = np.array(start_points)
            if start.shape[0] > n_best_restarts > 0:
                values = averaged(start)[0].cpu().numpy()
                keep = np.argsort(values, kind='stable')[-n_best_restarts:]
                start = start[keep]
        else:
            start = np.asarray(start, dtype=float).reshape(1, -1)
        lower = np.array([b[0] for b in bounds_x], dtype=float)
        upper = np.array([b[-1] for b in bounds_x], dtype=float)
        if n_samples_parameters == 0 and not minimize:
            mx, arg, _ = bq.inner_maximize(ub, np.zeros(1), start, maxiter=100, factr=1e7,
                                           use_vertices=False)
            # the kernel already keeps the best run over the starts (np.argmax over restarts)
            ends = arg.cpu().numpy()[0, 0:1]
            values = mx.cpu().numpy()[0, 0:1]
        else:
            ends, values = batched_adam_ascent(bq.eng, averaged, start, lower, upper,
                                               maxepoch=maxepoch if n_samples_parameters > 0 else 100,
                                               sign=sign)
        score = sign * values
        ind = int(np.argmax(score))
        solution, value = ends[ind], float(values[ind])
        max_ = float(np.max(score))
        if candidate_solutions is not None or vertex is not None:
            cand = [list(c) for c in (candidate_solutions or [])] + (vertex or [])
            cvals = averaged(np.array(cand, dtype=float))[0].cpu().numpy()
            j = int(np.argmax(cvals))
            if cvals[j] > max_:  # (the reference compares the raw values, also when minimising)
                solution, value = np.array(cand[j], dtype=float), float(cvals[j])
        result = {'solution': np.asarray(solution, dtype=float), 'optimal_value': np.array([value])}
        self.optimal_solutions.setdefault(index_cache, []).append(result)
        self.max_mean[index_cache] = max_
        return result

    # -- per-candidate quantities --------------------------------------------------------------------
    def get_parameters_for_samples(self, cache, candidate_point, parameters_kernel, var_noise, mean,
                                   clear_cache=True):
        """:1074-1145."""
        theta = self._theta(var_noise, mean, parameters_kernel)
        bq = self.engine_for(theta)
        ub = bq.setup(np.asarray(candidate_point, dtype=float).reshape(1, -1))
        n = self.gp.data['points'].shape[0]
        rows = 1 + bq.eng.desc.dim_m
        gamma_v = ub.G.reshape(1, 1, rows, n)[0, 0, 0].cpu().numpy().reshape(n, 1)
        s2 = ub.R.reshape(1, 1, rows, n)[0, 0, 0]
        if bq.is_product:
            s2 = s2 / bq.qt[0][bq.hb.tid.long()]
        return {
            'gamma': gamma_v, 'solve_2': s2.cpu().numpy().reshape(n, 1),
            'denominator': ub.den.cpu().numpy().reshape(1),
            'chol': bq.hb.L.cpu().numpy()[0], 'solve': bq.hb.alpha.cpu().numpy()[0],
            'parameters_kernel': theta[2:], 'var_noise': theta[0], 'mean': theta[1],
        }

    def _ab(self, point, candidate_point, theta, w_set=0):
        bq = self.engine_for(theta)
        ub = bq.setup(np.asarray(candidate_point, dtype=float).reshape(1, -1))
        dx = len(self.x_domain)
        pts = np.asarray(point, dtype=float).reshape(-1, dx)
        t = pts.shape[0]
        out = bq.sample_ab(ub, pts, np.zeros(t, dtype=np.int32),
                           np.full(t, w_set, dtype=np.int32) if self.w_sets is not None else None)
        return out.cpu().numpy()

    def compute_parameters_for_sample(self, point, candidate_point, var_noise=None, mean=None,
                                      parameters_kernel=None, cache=True, n_threads=0,
                                      clear_cache=True, w_set=0):
        """:1148-1191 -> {'a': np.array(1), 'b': np.array(1x1)}."""
        out = self._ab(point, candidate_point, self._theta(var_noise, mean, parameters_kernel), w_set)
        return {'a': out[0:1, 0], 'b': out[0:1, 1:2]}

    def compute_gradient_parameters_for_sample(self, point, candidate_point, var_noise=None,
                                               mean=None, parameters_kernel=None, cache=True,
                                               w_set=0):
        """:1193-1239 -> {'a': np.array(n), 'b': np.array(n)}."""
        out = self._ab(point, candidate_point, self._theta(var_noise, mean, parameters_kernel), w_set)
        dx = len(self.x_domain)
        return {'a': out[0, 2:2 + dx], 'b': out[0, 2 + dx:2 + 2 * dx]}

    def compute_hessian_parameters_for_sample(self, point, candidate_point, var_noise=None,
                                              mean=None, parameters_kernel=None, cache=True,
                                              w_set=0):
        """:1241-1286 -> {'a': np.array(n x n), 'b': np.array(n x n)} (b = 0 when den = 0)."""
        theta = self._theta(var_noise, mean, parameters_kernel)
        bq = self.engine_for(theta)
        ub = bq.setup(np.asarray(candidate_point, dtype=float).reshape(1, -1))
        dx = len(self.x_domain)
        pts = np.asarray(point, dtype=float).reshape(1, dx)
        out = bq.sample_ab_hess(ub, pts, np.zeros(1, dtype=np.int32),
                                np.full(1, w_set, dtype=np.int32) if self.w_sets is not None else None)
        out = out.cpu().numpy()[0]
        return {'a': out[0], 'b': out[1]}

    def gradient_vector_b(self, candidate_point, points, var_noise=None, mean=None,
                          parameters_kernel=None, cache=True, keep_indexes=None, parallel=True,
                          monte_carlo=False, n_threads=0, w_set=0):
        """:1447-1519 -> np.array(n x m), or the object np.nan when den^2 == 0."""
        theta = self._theta(var_noise, mean, parameters_kernel)
        bq = self.engine_for(theta)
        cand = np.asarray(candidate_point, dtype=float).reshape(1, -1)
        ub = bq.setup(cand)
        pts = np.asarray(points, dtype=float).reshape(1, -1, len(self.x_domain))
        ws = None
        if self.w_sets is not None:
            ws = np.full(pts.shape[1], w_set, dtype=np.int32)
        out, is_nan = bq.grad_vector_b(ub, pts, ws)
        if int(is_nan.cpu().numpy()[0]):
            return np.nan
        g = out[0].cpu().numpy()
        pad = cand.shape[1] - g.shape[1]
        if pad > 0:
            g = np.concatenate([g, np.zeros((g.shape[0], pad))], axis=1)
        return g

    def compute_posterior_parameters_kg(self, points, candidate_point, var_noise=None, mean=None,
                                        parameters_kernel=None, cache=True, parallel=True,
                                        n_threads=0, w_set=0):
        """:1376-1445 -> {'a': np.array(n), 'b': np.array(n)}."""
        out = self._ab(points, candidate_point, self._theta(var_noise, mean, parameters_kernel), w_set)
        return {'a': out[:, 0], 'b': out[:, 1]}

    def get_historical_best_solution(self, var_noise=None, mean=None, parameters_kernel=None,
                                     noisy_evaluations=False):
        """:1542-1577: max posterior mean of G over the x of the training points."""
        theta = self._theta(var_noise, mean, parameters_kernel)
        index = (theta[0], theta[1], tuple(theta[2:]))
        if index in self.best_solution:
            return self.best_solution[index]
        pts = self.gp.data['points'][:, self.x_domain]
        best = np.max(self.compute_posterior_parameters(pts, theta[0], theta[1], theta[2:],
                                                        only_mean=True)['mean'])
        self.best_solution[index] = best
        return best
