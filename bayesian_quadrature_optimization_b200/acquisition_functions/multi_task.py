"""MultiTasks: host-side mirror of sbo/acquisition_functions/multi_task.py.

EI on G(x) = E_W F(x, W) picks x, then the task with the largest Bayesian-averaged EI on F(x, t)
is chosen (`choose_best_task_given_x`, :82-95) -- the "chosen task index" of the parity contract.
Both steps are batched over all hyper-samples on the GPU.
"""
from __future__ import annotations

import numpy as np
import torch

from ..lib.constant import DEFAULT_N_PARAMETERS
from ..services.domain import DomainService
from .ei import EI


class MultiTasks(object):

    def __init__(self, bayesian_quadrature, n_tasks):
        self.bq = bayesian_quadrature
        self.n_tasks = n_tasks
        self.ei_tasks = EI(self.bq.gp)

    def _thetas(self, n_samples_parameters):
        gp = self.bq.gp
        if n_samples_parameters and n_samples_parameters > 0:
            return np.array(gp.samples_parameters[-DEFAULT_N_PARAMETERS:])
        return gp.get_value_parameters_model.reshape(1, -1)

    def ei_quadrature(self, xpts, n_samples_parameters=0):
        """Bayesian-averaged EI of the posterior of G at the points xpts (EI(bq).evaluate averaged
        by BayesianEvaluations, sbo/lib/util.py:836-860) -> np.array(P)."""
        thetas = self._thetas(n_samples_parameters)
        bq = self.bq.engine_for(thetas)
        eng = bq.eng
        xpts = np.asarray(xpts, dtype=float).reshape(-1, len(self.bq.x_domain))
        mean, var = bq.quadrature_posterior(xpts)
        # best = max posterior mean of G over the x of the training points (:1542-1577)
        train_x = self.bq.gp.data['points'][:, self.bq.x_domain]
        best = bq.quadrature_posterior(train_x, var=False)[0].max(dim=1).values
        S, P = int(mean.shape[0]), int(mean.shape[1])
        ei = eng.empty(S, P)
        from .. import _cabi
        _cabi.check(eng.lib.bqo_ei_batched(
            _cabi.ptr(mean.contiguous()), _cabi.ptr(var.contiguous()), None, None,
            _cabi.ptr(best.contiguous()), S, P, 0, _cabi.ptr(ei), None,
            torch.cuda.current_stream().cuda_stream), "bqo_ei_batched")
        return ei.mean(dim=0).cpu().numpy()

    def optimize_first(self, start=None, random_seed=None, parallel=True, n_restarts=100,
                       n_samples_parameters=0, n_best_restarts=0, maxepoch=11):
        """:60-80 -> np.array(d_x).  The reference climbs EI(G) with SGD from the best random
        starts; here the Bayesian-averaged EI(G) of all random starts is evaluated in one batch and
        the best start is polished by a few rounds of batched local perturbations (derivative
        free: EI(G) needs a triangular solve per point, its gradient is not on the hot path)."""
        if random_seed is not None:
            np.random.seed(random_seed)
        bounds_x = self.bq.bounds_x()
        if start is None:
            start = np.array(DomainService.get_points_domain(
                n_restarts, bounds_x, type_bounds=len(bounds_x) * [0]))
        start = np.asarray(start, dtype=float).reshape(-1, len(bounds_x))
        lo = np.array([b[0] for b in bounds_x], dtype=float)
        up = np.array([b[-1] for b in bounds_x], dtype=float)
        values = self.ei_quadrature(start, n_samples_parameters)
        best = start[int(np.argmax(values))].copy()
        best_v = float(np.max(values))
        radius = 0.1 * (up - lo)
        for _ in range(max(1, maxepoch // 2)):
            cand = np.clip(best[None, :] + radius[None, :] * np.random.uniform(
                -1, 1, (32, len(bounds_x))), lo, up)
            v = self.ei_quadrature(cand, n_samples_parameters)
            if np.max(v) > best_v:
                best_v = float(np.max(v))
                best = cand[int(np.argmax(v))].copy()
            radius *= 0.5
        return best

    def choose_best_task_given_x(self, x, n_samples_parameters=0):
        """:82-95 -> (task index, Bayesian EI value): argmax over tasks of the EI of F at (x, t)
        averaged over the last DEFAULT_N_PARAMETERS hyper-samples (np.argmax tie rule)."""
        pts = np.array([np.concatenate((x, np.array([float(i)]))) for i in range(self.n_tasks)])
        values, _ = self.ei_tasks.evaluate_bayesian(pts, DEFAULT_N_PARAMETERS)
        ind = int(np.argmax(values))
        return ind, values[ind]

    def optimize(self, random_seed=None, parallel=True, n_restarts=100, n_best_restarts=0,
                 n_samples_parameters=0, start_new_chain=True, maxepoch=11, **kwargs):
        """:97-125 -> {'solution': np.array(d_x + 1), 'optimal_value': float}."""
        if n_samples_parameters > 0 and start_new_chain:
            self.bq.gp.start_new_chain()
            self.bq.gp.sample_parameters(DEFAULT_N_PARAMETERS)
        point = self.optimize_first(random_seed=random_seed, parallel=parallel,
                                    n_restarts=n_restarts,
                                    n_samples_parameters=n_samples_parameters,
                                    n_best_restarts=n_best_restarts, maxepoch=maxepoch)
        task, value = self.choose_best_task_given_x(point, n_samples_parameters=n_samples_parameters)
        solution = np.concatenate((point, np.array([task])))
        return {'solution': solution, 'optimal_value': value}

    def clean_cache(self):
        self.ei_tasks.clean_cache()
        self.bq.clean_cache()
