"""SGD: host-side mirror of sbo/lib/stochastic_gradient_descent.py:12-132 for ONE start point.

Minimises the mean of n stochastic objectives with Adam (default) or momentum SGD, box / simplex
projection and the reference's stop rule (relative step below 1 %).  A gradient that comes back as
the object `np.nan` (zero predictive variance at the point, bayesian_quadrature.py:1483-1484) makes
the point jump by a random perturbation of relative size 1e-6 until a gradient is available.
The multi-start drivers of this package (SBO.optimize, EI.optimize, optimize_posterior_mean) run
the same update for all restarts at once on the device (`bqo_adam_step_batched`,
lib/batched_ascent.py); this function keeps the reference's single-chain entry point.
"""
from __future__ import annotations

import numpy as np


def _perturb(point, bounds, project):
    """Random move of relative size 1e-6 that stays inside the box."""
    radius = np.sqrt(np.sum(point ** 2)) * 1e-6
    moved = point.copy()
    for i in range(len(point)):
        lo = hi = radius
        if project:
            lo = min(radius, point[i] - bounds[i][0])
            hi = min(radius, bounds[i][1] - point[i])
        moved[i] += np.random.uniform(-lo, hi)
    return moved


def _project(point, bounds, simplex_domain):
    for dim, (lo, hi) in enumerate(bounds):
        if lo is not None:
            point[dim] = max(lo, point[dim])
        if hi is not None:
            point[dim] = min(hi, point[dim])
    if simplex_domain is not None and np.sum(point[2:]) > simplex_domain:
        point[2:] = simplex_domain * (point[2:] / np.sum(point[2:]))
    return point


def _outside(point, bounds, simplex_domain):
    for dim, (lo, hi) in enumerate(bounds):
        if (lo is not None and point[dim] < lo) or (hi is not None and point[dim] > hi):
            return True
    return simplex_domain is not None and np.sum(point[2:]) > simplex_domain


def SGD(start, gradient, n, args=(), kwargs={}, bounds=None, learning_rate=0.1, momentum=0.9,
        maxepoch=250, adam=True, betas=None, eps=1e-8, simplex_domain=None):
    """:param start: np.array(d); :param gradient: callable(point, *args, **kwargs) -> np.array(d)
    or the object np.nan; :param n: gradients averaged per epoch; :return: np.array(d)."""
    project = bounds is not None or simplex_domain is not None
    if betas is None:
        betas = (0.9, 0.999)
    point = np.array(start, dtype=float)
    first = np.zeros(len(point))
    second = np.zeros(len(point))
    velocity = np.zeros(len(point))
    for epoch in range(1, maxepoch + 1):
        previous = point.copy()
        draws = []
        for _ in range(n):
            g = gradient(point, *args, **kwargs)
            while g is np.nan:
                point = _perturb(point, bounds, project)
                g = gradient(point, *args, **kwargs)
            draws.append(g)
        g = np.mean(np.array(draws), axis=0)
        before_step = point.copy()
        if adam:
            first = betas[0] * first + (1 - betas[0]) * g
            second = betas[1] * second + (1 - betas[1]) * (g ** 2)
            m_hat = first / (1 - betas[0] ** epoch)
            v_hat = second / (1 - betas[1] ** epoch)
            point = point - learning_rate * m_hat / (np.sqrt(v_hat) + eps)
        else:
            velocity = momentum * velocity + g
            point = point - learning_rate * velocity
        if project and _outside(point, bounds, simplex_domain):
            point = _project(point, bounds, simplex_domain)
            if not adam:
                velocity = (point - before_step) / learning_rate
        scale = np.sqrt(np.sum(previous ** 2))
        step = np.sqrt(np.sum((previous - point) ** 2)) / (scale if scale != 0 else 1e-2)
        if step < 0.01:
            break
    return point
