"""BASELINE.json configs[3] (citi_bike_mt-shaped: 4-D x, 100 finite tasks with Poisson weights,
n = 4096) through the same step as bench.py's C3 line: stage B + inner maximiser + gradient of b +
reduction for 111 candidates x 20 hyper-samples x 100 Z.  A data point for the finite-W path, not a
bench line (CUDA events, 3 warm-up + 5 timed steps, L2 flushed between steps)."""
import sys

import numpy as np
import torch
from scipy.stats import poisson

sys.path.insert(0, ".")
from bayesian_quadrature_optimization_b200 import engine as E, _cabi  # noqa: E402

n, dx, T, S, C, Z, R = 4096, 4, 100, 20, 111, 100, 6
rng = np.random.RandomState(1)
x = np.floor(rng.uniform(0, 10, (n, dx)) * 4) / 4.0
task = rng.randint(0, T, n).astype(float)
pts = np.concatenate([x, task[:, None]], 1)
y = np.sin(x.sum(1) / 3.0) + 0.02 * (task - 50.0) / 50.0 + 0.05 * rng.normal(0, 1, n)
weights = poisson.pmf(np.arange(T), 50)
weights = weights / weights.sum()
base = np.array([0.05, 0.0, 2.0, 2.5, 3.0, 2.0, 0.0, -1.0])
thetas = base[None, :] + 0.05 * rng.normal(0, 1, (S, base.size))
thetas[:, 0] = np.abs(thetas[:, 0])
eng = E.GPEngine(_cabi.KIND_PRODUCT_TASKS, dx + 1, T, True)
eng.set_data(pts, y)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
hb = eng.hypers(thetas)
torch.cuda.synchronize()
ev0.record()
hb.factorize()
ev1.record()
torch.cuda.synchronize()
assert np.all(hb.info == 0)
print("factorise 20 x 4096^2: %.1f ms" % ev0.elapsed_time(ev1))
bq = E.BQEngine(hb, dx, weights=weights, lower=np.zeros(dx), upper=10.0 * np.ones(dx))
zs = eng.to_device(np.random.RandomState(3).normal(0, 1, Z))
starts = eng.to_device(np.random.RandomState(6).uniform(0, 10, (R, dx)))
cands = np.concatenate([rng.uniform(0, 10, (8 * C, dx)), rng.randint(0, T, (8 * C, 1)).astype(float)], 1)
cands = eng.to_device(cands)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)


def step(i):
    ub = bq.setup(cands[i * C:(i + 1) * C])
    mx, arg, ne = bq.inner_maximize(ub, zs, starts, maxiter=10, factr=1e12, count_evals=True)
    gvb, is_nan = bq.grad_vector_b(ub, arg)
    val, grad = bq.acq_reduce(ub, mx, gvb, is_nan, zs)
    flush.fill_(i & 0xFF)
    return ne


for i in range(3):
    step(i)
torch.cuda.synchronize()
ev0.record()
cnt = [step(3 + i) for i in range(5)]
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / 5
evals = float(sum(int(c.sum().item()) for c in cnt)) / 5
print("C4 shape: %.1f ms per step of %d units = %.0f units/s; %.1f evaluations per (unit, Z); "
      "%.2f G training-point kernel evaluations/s" % (ms, C * S, C * S / (ms * 1e-3), evals / (C * S * Z),
                                                       evals * n / (ms * 1e-3) / 1e9))
