"""ParameterEntity: sbo/entities/parameter.py (name, value, prior, bounds of one parameter block)."""
import numpy as np


class ParameterEntity(object):
    """sbo/entities/parameter.py: name, value (np.array), prior, bounds."""

    def __init__(self, name, value, prior=None, bounds=None):
        from ..lib.constant import SMALLEST_NUMBER, LARGEST_NUMBER
        self.name = name
        self.value = np.asarray(value, dtype=float)
        self.prior = prior
        self.dimension = len(self.value)
        if bounds is None:  # entities/parameter.py:29-30,48-58
            bounds = self.dimension * [(SMALLEST_NUMBER, LARGEST_NUMBER)]
        self.bounds = bounds

    def set_value(self, value):
        self.value = np.asarray(value, dtype=float)

    def log_prior(self, value=None):
        if self.prior is None:
            return 0.0
        return self.prior.logprob(self.value if value is None else value)

    def sample_from_prior(self, n_samples, random_seed=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        return self.prior.sample(n_samples)
