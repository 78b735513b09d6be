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
        self.ei = EI(self.bq)           # EI on G(x) = E_W F(x, W)   (multi_task.py:53)
        self.ei_tasks = EI(self.bq.gp)  # EI on F(x, t)              (multi_task.py:54)

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
        """:60-80 -> np.array(d_x): maximise EI on the posterior of G exactly as the reference
        does, through EI(bq).optimize (all restarts advance together on the device)."""
        solution = self.ei.optimize(start, random_seed, parallel, n_restarts,
                                    n_best_restarts=n_best_restarts,
                                    n_samples_parameters=n_samples_parameters, maxepoch=maxepoch)
        return np.asarray(solution['solution'], dtype=float)[:len(self.bq.x_domain)]

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
