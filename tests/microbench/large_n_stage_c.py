"""Stage C beyond the shared-memory limit (n = 8192) next to the staged case (n = 2048), same
hyper-samples, candidates and Z: rate of training-point evaluations of the inner maximiser
  * staged (n = 2048: the training set in shared memory),
  * streamed from L2 (n = 8192), with and without the W-distance table.
444 units (three waves of CTAs) so that the slowest unit does not set the time."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from bayesian_quadrature_optimization_b200 import engine as E, _cabi  # noqa: E402

S, C, NZ = 4, 111, 100
wl = bench.make_workload()
for n in (2048, 8192):
    rng = np.random.RandomState(0)
    pts = np.concatenate([rng.uniform(0, 1, (n, 4)), rng.gamma(2.0, 1.0, (n, 2))], 1)
    y = np.sin(pts.sum(1)) + 0.01 * rng.normal(0, 1, n)
    eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, 6)
    eng.set_data(pts, y)
    hb = eng.hypers(wl["thetas"][:S]).factorize()
    bq = E.BQEngine(hb, 4, wsets=wl["wsets"], lower=wl["lower"], upper=wl["upper"])
    ub = bq.setup(wl["cands"][:C])
    for use in ((True,) if n == 2048 else (True, False)):
        E.BQEngine.use_wdist = use
        ts = []
        for rep in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mx, arg, ne = bq.inner_maximize(ub, wl["zs"][:NZ], wl["starts"], maxiter=10, factr=1e12,
                                            count_evals=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ev = int(ne.sum().item())
        print("n=%d, table=%s, %s: %.1f ms for %d units x %d Z (%d evaluations): "
              "%.1f G training-point evaluations/s"
              % (n, use, "staged" if n == 2048 else "streamed",
                 1e3 * min(ts), C * S, NZ, ev, ev * n / min(ts) / 1e9))
    del bq, ub, hb, eng
E.BQEngine.use_wdist = True
