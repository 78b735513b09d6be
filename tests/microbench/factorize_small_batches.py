"""Factorisation (+ alpha) of small batches, n = 2048: alpha by two triangular vector solves against
alpha from L^-1 produced by extra tasks of the same launch (CUDA events, min of 5)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from bayesian_quadrature_optimization_b200 import engine as E, _cabi  # noqa: E402

wl = bench.make_workload()
eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, bench.DX + bench.DW)
eng.set_data(wl["pts"], wl["y"])


def timed(f, reps=6):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        f()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for S in (1, 2, 4, 8):
    th = wl["thetas"][:S]
    for inv in (False, True):
        def run():
            hb = eng.hypers(th)
            hb.factorize(with_inverse=inv)
            return hb.log_likelihood()
        print("S=%d alpha from %s: hypers + factorize + alpha + log-likelihood %.3f ms"
              % (S, "L^-1 " if inv else "solves", timed(run)))
