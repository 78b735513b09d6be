"""Host-logic fixtures from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_host.py

* DomainService.get_points_domain with a simplex domain (sbo/services/domain.py:115-170), seeded;
* uniform_finite with sub-sampling (`n_samples`, sbo/lib/expectations.py:10-74): the tasks drawn
  for a seed and the resulting expectation of a known function.
Output: tests/golden/host_misc.npz.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()

from stratified_bayesian_optimization.services.domain import DomainService  # noqa: E402
from stratified_bayesian_optimization.lib.expectations import uniform_finite  # noqa: E402

bounds = [[0.0, 1.0], [2.0, 5.0], [0, 40], [0, 30], [0, 20], [0, 50]]
np.random.seed(17)
pts = np.array(DomainService.get_points_domain(7, bounds, type_bounds=6 * [0], simplex_domain=60))

T = 6
weights = np.array([0.05, 0.3, 0.1, 0.25, 0.2, 0.1])
domain_random = np.arange(T).reshape(T, 1)
table = np.array([1.0, -2.0, 0.5, 4.0, 3.0, -1.5])
f = lambda z: table[z[:, 1].astype(int)].reshape(-1, 1) * z[:, 0:1]
point = np.array([[2.0]])
np.random.seed(23)
sub_w = uniform_finite(f, point, [0], domain_random, [1], weights=weights, n_samples=11)
np.random.seed(23)
idx_w = np.random.choice(T, 11, replace=True, p=weights)
np.random.seed(29)
sub_u = uniform_finite(f, point, [0], domain_random, [1], weights=None, n_samples=9)
np.random.seed(29)
idx_u = np.random.choice(T, 9, replace=True, p=None)
np.savez_compressed(os.path.join(HERE, "host_misc.npz"), simplex_bounds=np.array(bounds, dtype=float),
                    simplex_points=pts, simplex_domain=60, weights=weights, table=table,
                    sub_weighted=np.asarray(sub_w), sub_weighted_index=idx_w,
                    sub_uniform=np.asarray(sub_u), sub_uniform_index=idx_u)
print(pts, sub_w, idx_w, sub_u, idx_u)
