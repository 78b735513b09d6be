"""BayesianEvaluations: host-side mirror of sbo/bayesian/bayesian_evaluations.py:9-44.

Generic Bayesian average of f(point; theta) over the last `number_samples` hyper-samples of a GP
model.  It calls `function` once per hyper-sample like the reference (each call runs on the GPU and
finds its factor in the model's device-resident cache); the acquisition functions of this package
use their batched equivalents instead (EI.evaluate_bayesian, SBO.evaluate_mc_bayesian_*), which
evaluate all hyper-samples in one launch.
"""
from __future__ import annotations

import numpy as np


class BayesianEvaluations(object):

    @classmethod
    def evaluate(cls, function, point, gp_model, number_samples=20, random_seed=None, *args,
                 **kwargs):
        """-> (mean, std) of function(point, *args, var_noise, mean, parameters_kernel, **kwargs)
        over the hyper-samples; floats for scalar functions, arrays (axis 0) otherwise."""
        if random_seed is not None:
            np.random.seed(random_seed)
        values = []
        for theta in gp_model.samples_parameters[-number_samples:]:
            theta = np.asarray(theta, dtype=float)
            values.append(function(point, *(args + (theta[0], theta[1], theta[2:])), **kwargs))
        if isinstance(values[0], float):
            return np.mean(values), np.std(values)
        return np.mean(values, axis=0), np.std(values, axis=0)
