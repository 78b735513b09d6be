"""Stage C at n = 8192 (training set does not fit shared memory): units/s of one pass with a few
candidates, with and without the W-distance table."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from bayesian_quadrature_optimization_b200 import engine as E, _cabi  # noqa: E402

n, S, C, NZ = 8192, 4, 37, 20
rng = np.random.RandomState(0)
pts = np.concatenate([rng.uniform(0, 1, (n, 4)), rng.gamma(2.0, 1.0, (n, 2))], 1)
y = np.sin(pts.sum(1)) + 0.01 * rng.normal(0, 1, n)
wl = bench.make_workload()
eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, 6)
eng.set_data(pts, y)
hb = eng.hypers(wl["thetas"][:S]).factorize()
bq = E.BQEngine(hb, 4, wsets=wl["wsets"], lower=wl["lower"], upper=wl["upper"])
ub = bq.setup(wl["cands"][:C])
for use in (True, False):
    E.BQEngine.use_wdist = use
    ts = []
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mx, arg, ne = bq.inner_maximize(ub, wl["zs"][:NZ], wl["starts"], maxiter=10, factr=1e12,
                                        count_evals=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print("n=%d, table=%s: %.1f ms for %d units x %d Z (%d evaluations): %.0f evaluations/ms"
          % (n, use, 1e3 * min(ts), C * S, NZ, int(ne.sum().item()), ne.sum().item() / (1e3 * min(ts))))
E.BQEngine.use_wdist = True
