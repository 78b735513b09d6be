"""DomainService.get_points_domain (sbo/services/domain.py:93-200): random start points.
The order of the numpy RNG calls defines the start points for a seed, so it is kept."""
import numpy as np


class DomainService(object):

    @classmethod
    def get_points_domain(cls, n_samples, bounds_domain, type_bounds=None, random_seed=None,
                          simplex_domain=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        if type_bounds is None:
            type_bounds = [0] * len(bounds_domain)
        if simplex_domain is not None:
            return cls._points_simplex(n_samples, bounds_domain, type_bounds, simplex_domain)
        points = []
        for j in range(len(bounds_domain)):
            points.append(cls.get_point_one_dimension_domain(n_samples, bounds_domain[j],
                                                             type_bounds=type_bounds[j]))
        return [[point[j] for point in points] for j in range(n_samples)]

    @classmethod
    def _points_simplex(cls, n_samples, bounds_domain, type_bounds, simplex_domain):
        """:115-170, the citi_bike_mt domain: two free coordinates, then four integer coordinates
        drawn in order, each capped so that their running sum stays <= simplex_domain.  Same
        numpy calls in the same order as the reference."""
        if type_bounds[-1] == 1:
            raise NotImplementedError(
                "simplex_domain together with a finite last coordinate: the reference's own "
                "branch (services/domain.py:160-164) returns a malformed list")
        first = [cls.get_point_one_dimension_domain(n_samples, bounds_domain[j],
                                                    type_bounds=type_bounds[j]) for j in range(2)]
        s = np.random.uniform(0, 1, (n_samples, 4))
        s[:, 0] = np.floor(s[:, 0] * bounds_domain[2][1] + (1 - s[:, 0]) * bounds_domain[2][0])
        for j in range(n_samples):
            for k in (1, 2, 3):
                b = bounds_domain[2 + k]
                tmp = s[j, k] * b[1] + (1 - s[j, k]) * b[0]
                s[j, k] = min(np.floor(tmp), simplex_domain - np.sum(s[j, 0:k]))
        return [[first[0][i], first[1][i]] + list(s[i, :]) for i in range(n_samples)]

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
