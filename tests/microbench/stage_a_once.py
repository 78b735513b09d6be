"""Two passes of stage A (assemble + persistent Cholesky with L^-1 + alpha + log-likelihood +
gradient) at the bench shape, for the ncu launch list.  Tuning aid; run under gpurun."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from bayesian_quadrature_optimization_b200 import engine as E, _cabi  # noqa: E402

wl = bench.make_workload()
eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, bench.DX + bench.DW)
eng.set_data(wl["pts"], wl["y"])
for rep in range(2):
    hb = eng.hypers(wl["thetas"])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    hb.factorize(with_inverse=True)
    llh = hb.log_likelihood()
    g = hb.grad_log_likelihood()
    b.record()
    torch.cuda.synchronize()
    print("pass %d: %.3f ms" % (rep, a.elapsed_time(b)))
    # second flavour: inverse built after the factorisation (separate launches)
    hb2 = eng.hypers(wl["thetas"])
    a.record()
    hb2.factorize(with_inverse=False)
    g2 = hb2.grad_log_likelihood()
    b.record()
    torch.cuda.synchronize()
    print("pass %d, separate inverse: %.3f ms; max |grad diff| %.3e" % (rep, a.elapsed_time(b), (g - g2).abs().max().item()))
