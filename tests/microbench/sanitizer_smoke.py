"""One small pass through every kernel family of libbqo_sm100.so for compute-sanitizer
(memcheck / racecheck / synccheck): n = 200 (four 64-wide tiles, so the persistent Cholesky's
cross-CTA hand-off is exercised), gamma-W and finite-W models.  Run under gpurun:

    compute-sanitizer --tool racecheck python tests/microbench/sanitizer_smoke.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bayesian_quadrature_optimization_b200 import engine as E, _cabi  # noqa: E402

rng = np.random.RandomState(0)


def run(kind, n, dx, dw, n_tasks=0):
    dim = dx + (dw if kind != _cabi.KIND_PRODUCT_TASKS else 1)
    if kind == _cabi.KIND_PRODUCT_TASKS:
        pts = np.concatenate([rng.uniform(0, 1, (n, dx)), rng.randint(0, n_tasks, (n, 1)).astype(float)], 1)
        thetas = np.array([[1e-3, 0.0] + dx * [0.4] + [0.0, -1.0], [2e-3, 0.1] + dx * [0.3] + [0.1, -0.8]])
        wsets, weights = None, np.full(n_tasks, 1.0 / n_tasks)
    else:
        pts = np.concatenate([rng.uniform(0, 1, (n, dx)), rng.gamma(2.0, 1.0, (n, dw))], 1)
        thetas = np.array([[1e-3, 0.0] + dx * [0.4] + dw * [2.0] + [1.0],
                           [2e-3, 0.1] + dx * [0.3] + dw * [2.5] + [1.2]])
        wsets, weights = rng.gamma(2.0, 1.0, (2, 10, dw)), None
    y = np.sin(pts.sum(1)) + 0.01 * rng.normal(0, 1, n)
    eng = E.GPEngine(kind, dim, n_tasks, kind == _cabi.KIND_PRODUCT_TASKS)
    eng.set_data(pts, y)
    hb = eng.hypers(thetas)
    hb.cov(add_noise=True)
    hb.grad_cov()
    hb.factorize(with_inverse=True)
    assert np.all(hb.info == 0)
    hb.log_likelihood()
    hb.grad_log_likelihood()
    hb2 = eng.hypers(thetas).factorize()
    hb2.grad_log_likelihood()
    q = pts[:3].copy()
    hb.posterior(q, var=True, grad=True)
    hb.ei(q, np.full(2, y.max()), grad=True)
    hb.cross_cov_hessian(q)
    bq = E.BQEngine(hb, dx, weights=weights, wsets=wsets, lower=np.zeros(dx), upper=np.ones(dx))
    ub = bq.setup(pts[5:8])
    zs = rng.normal(0, 1, 4)
    starts = rng.uniform(0, 1, (3, dx))
    res = bq.evaluate_candidates(pts[5:8], zs, starts, max_mean=np.zeros(2))
    n_units = int(ub.unit_k.shape[0])
    xq = rng.uniform(0, 1, (4, dx))
    bq.sample_ab(ub, xq, np.zeros(4, dtype=np.int32), np.zeros(4, dtype=np.int32) if dw else None)
    bq.sample_ab_hess(ub, xq, np.zeros(4, dtype=np.int32), np.zeros(4, dtype=np.int32) if dw else None)
    bq.grad_vector_b(ub, np.tile(rng.uniform(0, 1, (1, 40, dx)), (n_units, 1, 1)))
    E.kg_discretized(eng, rng.normal(0, 1, 30), rng.normal(0, 1, (2, 30)))
    torch.cuda.synchronize()
    return float(res["evaluations"].sum().item())


print("scaled/gamma", run(_cabi.KIND_SCALED_MATERN52, 200, 2, 2))
print("product/tasks", run(_cabi.KIND_PRODUCT_TASKS, 200, 2, 0, n_tasks=4))
print("odd n", run(_cabi.KIND_SCALED_MATERN52, 131, 2, 1))
print("one tile", run(_cabi.KIND_SCALED_MATERN52, 64, 1, 1))
print("n = 9", run(_cabi.KIND_PRODUCT_TASKS, 9, 1, 0, n_tasks=2))
print("sanitizer smoke done")
