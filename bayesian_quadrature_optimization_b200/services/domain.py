"""DomainService.get_points_domain (sbo/services/domain.py:93-200): random start points.
The order of the numpy RNG calls defines the start points for a seed, so it is kept."""
import numpy as np


class DomainService(object):

    @classmethod
    def get_points_domain(cls, n_samples, bounds_domain, type_bounds=None, random_seed=None,
                          simplex_domain=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        if simplex_domain is not None:
            raise NotImplementedError("simplex domains are outside the accelerated path")
        if type_bounds is None:
            type_bounds = [0] * len(bounds_domain)
        points = []
        for j in range(len(bounds_domain)):
            points.append(cls.get_point_one_dimension_domain(n_samples, bounds_domain[j],
                                                             type_bounds=type_bounds[j]))
        return [[point[j] for point in points] for j in range(n_samples)]

    @classmethod
    def get_point_one_dimension_domain(cls, n_samples, bounds, type_bounds=0):
        if type_bounds == 1:
            samples = bounds
            n_samples -= len(bounds)
            if n_samples <= 0:
                return samples
        if type_bounds == 0:
            return list(np.random.uniform(bounds[0], bounds[1], n_samples))
        return list(np.random.choice(bounds, n_samples)) + samples
