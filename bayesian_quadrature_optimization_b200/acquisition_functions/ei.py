"""EI: host-side mirror of sbo/acquisition_functions/ei.py over the CUDA engine.

`evaluate` / `evaluate_gradient` keep the reference's single-hyper-sample signature; the
Bayesian average over the slice-sampled hyper-samples (BayesianEvaluations.evaluate,
sbo/bayesian/bayesian_evaluations.py:9-44) is ONE batched launch over all hyper-samples
(`evaluate_bayesian`).
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np
import torch
from scipy.optimize import fmin_l_bfgs_b
from scipy.stats import norm

from ..lib.constant import BAYESIAN_QUADRATURE, DEFAULT_N_PARAMETERS
from ..services.domain import DomainService


class EI(object):

    def __init__(self, gp, noisy_evaluations=False):
        """:param gp: GPFittingGaussian or BayesianQuadrature instance."""
        self.gp = gp
        self.noisy_evaluations = noisy_evaluations
        self.optimization_results = []
        self.bounds_opt = deepcopy(self.gp.bounds)
        self.simplex_domain = self.gp.simplex_domain
        if self.gp.separate_tasks and not self.gp.model_only_x and \
                self.gp.name_model == BAYESIAN_QUADRATURE:
            self.bounds_opt.append([None, None])
        elif self.gp.separate_tasks and self.gp.name_model != BAYESIAN_QUADRATURE:
            self.bounds_opt = self.bounds_opt[0:-1]
            self.bounds_opt.append([None, None])

    def _is_bq(self):
        return self.gp.name_model == BAYESIAN_QUADRATURE

    def _gp_model(self):
        return self.gp.gp if self._is_bq() else self.gp

    # -- single hyper-sample (reference signatures) ---------------------------------------------------
    def evaluate(self, point, var_noise=None, mean=None, parameters_kernel=None):
        """:81-112 -> np.array(k)."""
        point = np.asarray(point, dtype=float)
        if not self._is_bq():
            gp = self.gp
            var_noise, mean, parameters_kernel = gp._defaults(var_noise, mean, parameters_kernel)
            hb = gp._hb1(var_noise, mean, parameters_kernel)
            hb.factorize(max_tries=7)
            gp._raise_info(int(hb.info[0]))
            best = gp.get_historical_best_solution(var_noise, mean, parameters_kernel,
                                                   self.noisy_evaluations)
            ei, _ = hb.ei(point, np.array([best]), grad=False)
            return ei.cpu().numpy()[0]
        # the quadrature model lives on the x coordinates; extra (task) columns are ignored
        point = point.reshape(-1, point.shape[-1])[:, :len(self.gp.x_domain)]
        post = self.gp.compute_posterior_parameters(point, var_noise, mean, parameters_kernel)
        mu = post['mean']
        cov = np.clip(post['cov'], 0, None)
        best = self.gp.get_historical_best_solution(var_noise, mean, parameters_kernel,
                                                    self.noisy_evaluations)
        with np.errstate(divide='ignore', invalid='ignore'):
            normalized_factor = (mu - best) / np.sqrt(cov)
            evaluation = (mu - best) * norm.cdf(normalized_factor) + \
                np.sqrt(cov) * norm.pdf(normalized_factor)
        return np.atleast_1d(evaluation)

    def _bq_ei_batched(self, thetas, xpts, gradient):
        """EI of the posterior of G at xpts for every hyper-sample in `thetas` -> (ei [S][P],
        grad [S][P][d_x] or None), device tensors.  best_s = max posterior mean of G over the x of
        the training points (BayesianQuadrature.get_historical_best_solution, :1542-1577)."""
        from .. import _cabi
        quad = self.gp
        bq = quad.engine_for(np.asarray(thetas, dtype=float))
        eng = bq.eng
        dx = len(quad.x_domain)
        xpts = eng.to_device(xpts).reshape(-1, dx)
        if gradient:
            mean, var, gm, gv = bq.quadrature_posterior_grad(xpts)
        else:
            mean, var = bq.quadrature_posterior(xpts)
            gm = gv = None
        var = torch.clamp(var, min=0.0).contiguous()  # np.clip(cov, 0, None), ei.py:99
        train_x = quad.gp.data['points'][:, quad.x_domain]
        best = bq.quadrature_posterior(train_x, var=False)[0].max(dim=1).values.contiguous()
        S, P = int(mean.shape[0]), int(mean.shape[1])
        ei = eng.empty(S, P)
        gei = eng.empty(S, P, dx) if gradient else None
        _cabi.check(eng.lib.bqo_ei_batched(
            _cabi.ptr(mean.contiguous()), _cabi.ptr(var), _cabi.ptr(gm), _cabi.ptr(gv),
            _cabi.ptr(best), S, P, dx if gradient else 0, _cabi.ptr(ei), _cabi.ptr(gei),
            torch.cuda.current_stream().cuda_stream), "bqo_ei_batched")
        return ei, gei

    def evaluate_gradient(self, point, var_noise=None, mean=None, parameters_kernel=None):
        """:135-176 -> np.array(n)."""
        point = np.asarray(point, dtype=float).reshape(1, -1)
        if self._is_bq():
            theta = self.gp._theta(var_noise, mean, parameters_kernel)
            dx = len(self.gp.x_domain)
            _, gei = self._bq_ei_batched(theta.reshape(1, -1), point[:, :dx], True)
            g = gei.cpu().numpy()[0, 0]
            pad = point.shape[1] - dx
            return np.concatenate([g, np.zeros(pad)]) if pad > 0 else g
        gp = self.gp
        var_noise, mean, parameters_kernel = gp._defaults(var_noise, mean, parameters_kernel)
        hb = gp._hb1(var_noise, mean, parameters_kernel)
        hb.factorize(max_tries=7)
        gp._raise_info(int(hb.info[0]))
        best = gp.get_historical_best_solution(var_noise, mean, parameters_kernel,
                                               self.noisy_evaluations)
        _, gei = hb.ei(point, np.array([best]), grad=True)
        g = gei.cpu().numpy()[0, 0]
        pad = point.shape[1] - g.shape[0]
        return np.concatenate([g, np.zeros(pad)]) if pad > 0 else g

    # -- Bayesian average over hyper-samples (batched) -----------------------------------------------
    def evaluate_bayesian(self, points, n_samples_parameters=DEFAULT_N_PARAMETERS, gradient=False):
        """mean over the last n hyper-samples of EI(points; theta_k) -> (np.array(k), grad or None).
        The reference loops BayesianEvaluations.evaluate over theta with a cold Cholesky cache."""
        gp = self._gp_model()
        if self._is_bq():
            thetas = np.array(gp.samples_parameters[-n_samples_parameters:])
            dx = len(self.gp.x_domain)
            pts_in = points if isinstance(points, torch.Tensor) else np.asarray(points, dtype=float)
            pts_in = pts_in.reshape(-1, pts_in.shape[-1])[:, :dx]
            ei, gei = self._bq_ei_batched(thetas, pts_in, gradient)
            val = ei.mean(dim=0).cpu().numpy()
            return val, (gei.mean(dim=0).cpu().numpy() if gradient else None)
        thetas = np.array(gp.samples_parameters[-n_samples_parameters:])
        hb = gp.hyper_batch(thetas)
        hb.factorize(max_tries=7)
        for s in range(hb.S):
            gp._raise_info(int(hb.info[s]))
        if self.noisy_evaluations:
            best = hb.posterior(gp.data['points'], var=False)['mean'].max(dim=1).values
        else:
            best = torch.full((hb.S,), float(np.max(gp.data['evaluations'])), dtype=torch.float64,
                              device=hb.eng.device)
        pts_in = points if isinstance(points, torch.Tensor) else np.asarray(points, dtype=float)
        ei, gei = hb.ei(pts_in, best, grad=gradient)
        val = ei.mean(dim=0).cpu().numpy()
        grad = gei.mean(dim=0).cpu().numpy() if gradient else None
        return val, grad

    def optimize(self, start=None, random_seed=None, parallel=True, n_restarts=10,
                 n_best_restarts=0, n_samples_parameters=0, start_new_chain=False, maxepoch=11,
                 **kwargs):
        """:179-317 -> {'solution', 'optimal_value', 'gradient'}: multi-start maximisation of EI.
        With n_samples_parameters == 0: L-BFGS-B per start on the GPU objective.  Otherwise the
        Bayesian-averaged EI (all hyper-samples in one launch per step) is climbed from every
        start at once with the reference's Adam rule (sbo/lib/stochastic_gradient_descent.py)."""
        if random_seed is not None:
            np.random.seed(random_seed)
        gp = self.gp
        bounds = gp.bounds  # the quadrature model exposes the x bounds here (model_only_x)
        if start is None:
            start = np.array(DomainService.get_points_domain(
                n_restarts, bounds, type_bounds=gp.type_bounds, simplex_domain=self.simplex_domain))
        start = np.asarray(start, dtype=float).reshape(-1, len(bounds))
        n_par = n_samples_parameters if n_samples_parameters > 0 else 0
        if 0 < n_best_restarts < start.shape[0]:
            if n_par:
                values, _ = self.evaluate_bayesian(start, DEFAULT_N_PARAMETERS)
            else:
                values = np.array([self.evaluate(s.reshape(1, -1))[0] for s in start])
            start = start[np.argsort(values, kind='stable')[-n_best_restarts:]]
        lo = np.array([b[0] for b in bounds], dtype=float)
        up = np.array([b[-1] for b in bounds], dtype=float)
        results = []
        if n_par == 0:
            for s in start:
                f = lambda x: -1.0 * self.evaluate(x.reshape(1, -1))[0]
                g = lambda x: -1.0 * self.evaluate_gradient(x.reshape(1, -1))
                opt = fmin_l_bfgs_b(f, s, fprime=g, bounds=[(a, b) for a, b in zip(lo, up)])
                results.append({'solution': opt[0], 'optimal_value': -opt[1],
                                'gradient': -opt[2]['grad']})
        else:
            import ctypes as C
            from .. import _cabi
            eng = self._gp_model()._engine()
            pts = eng.to_device(start.copy())
            R, dim = int(pts.shape[0]), int(pts.shape[1])
            m0, v0 = torch.zeros_like(pts), torch.zeros_like(pts)
            active = torch.ones(R, dtype=torch.int32, device=eng.device)
            lo_d, up_d = eng.to_device(lo), eng.to_device(up)
            for t in range(1, maxepoch + 1):
                _, grad = self.evaluate_bayesian(pts, DEFAULT_N_PARAMETERS, gradient=True)
                g = np.zeros((R, dim))
                g[:, :grad.shape[1]] = -grad  # minimise -EI
                gd = eng.to_device(g)
                _cabi.check(eng.lib.bqo_adam_step_batched(
                    _cabi.ptr(pts), _cabi.ptr(gd), _cabi.ptr(m0), _cabi.ptr(v0), _cabi.ptr(active),
                    _cabi.ptr(lo_d), _cabi.ptr(up_d), R, dim, t, 0.1, 0.9, 0.999, 1e-8,
                    torch.cuda.current_stream().cuda_stream), "bqo_adam_step_batched")
                if int(active.sum().item()) == 0:
                    break
            sol = pts.cpu().numpy()
            values, grad = self.evaluate_bayesian(sol, DEFAULT_N_PARAMETERS, gradient=True)
            for r in range(R):
                results.append({'solution': sol[r], 'optimal_value': values[r], 'gradient': grad[r]})
        ind_max = int(np.argmax([r['optimal_value'] for r in results]))
        self.optimization_results.append(results[ind_max])
        return results[ind_max]

    def clean_cache(self):
        """ei.py:319-323; the GP keeps its device factors (GPFittingGaussian.clean_cache)."""
        if hasattr(self.gp, 'gp'):  # EI on a BayesianQuadrature model
            self.gp.clean_cache()
        else:
            self.gp.clean_cache(keep_factors=True)
