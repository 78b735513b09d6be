"""Per-phase CUDA-event timing of stage A (assemble + Cholesky + alpha + log-likelihood + gradient)
at the bench shape (n = 2048, d = 6, 20 hyper-samples).  Tuning aid; run under gpurun."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from bayesian_quadrature_optimization_b200 import engine as E, _cabi  # noqa: E402
from bayesian_quadrature_optimization_b200._cabi import ptr, check  # noqa: E402

wl = bench.make_workload()
eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, bench.DX + bench.DW)
eng.set_data(wl["pts"], wl["y"])
n, S = eng.n, len(wl["thetas"])
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


hb = eng.hypers(wl["thetas"])
print("cov (assemble)      %.3f ms" % timed(lambda: hb.cov(add_noise=True)))
Sigma = hb.cov(add_noise=True)
info = torch.zeros(S, dtype=torch.int32, device=eng.device)
A = Sigma.clone()
work = torch.empty(int(eng.lib.bqo_potrf_workspace_bytes(n, S)) // 8 + 2, dtype=torch.float64, device=eng.device)


def potrf():
    A.copy_(Sigma)
    check(eng.lib.bqo_potrf_batched(ptr(A), n, S, ptr(info), ptr(work), st), "potrf")


tcopy = timed(lambda: A.copy_(Sigma))
print("potrf               %.3f ms (copy %.3f subtracted)" % (timed(potrf) - tcopy, tcopy))
eng.lib.bqo_potrf_set_multi_launch(1)
print("potrf multi-launch  %.3f ms (copy %.3f subtracted)" % (timed(potrf) - tcopy, tcopy))
eng.lib.bqo_potrf_set_multi_launch(0)
for S1 in (1, 4, 8):
    A1 = Sigma[:S1].clone()
    def potrf1():
        A1.copy_(Sigma[:S1])
        check(eng.lib.bqo_potrf_batched(ptr(A1), n, S1, ptr(info), ptr(work), st), "potrf")
    tc = timed(lambda: A1.copy_(Sigma[:S1]))
    print("potrf S=%d            %.3f ms" % (S1, timed(potrf1) - tc))
hb.factorize()
B = torch.randn(S, 1, n, dtype=torch.float64, device=eng.device)
print("potrs m=1           %.3f ms" % timed(lambda: check(eng.lib.bqo_potrs_batched(ptr(hb.L), n, S, ptr(B), 1, st), "potrs")))
B7 = torch.randn(S, 777, n, dtype=torch.float64, device=eng.device)
print("potrs m=777         %.3f ms" % timed(lambda: check(eng.lib.bqo_potrs_batched(ptr(hb.L), n, S, ptr(B7), 777, st), "potrs")))
print("log_likelihood      %.3f ms" % timed(lambda: hb.log_likelihood()))
print("grad_log_likelihood %.3f ms" % timed(lambda: hb.grad_log_likelihood()))


def full():
    h2 = eng.hypers(wl["thetas"])
    h2.factorize(with_inverse=True)
    h2.log_likelihood()
    h2.grad_log_likelihood()


print("full loglik+grad    %.3f ms" % timed(full, reps=3))


def fact(inv):
    h2 = eng.hypers(wl["thetas"])
    h2.factorize(with_inverse=inv)


print("hypers+factorize (assemble, potrf, alpha)              %.3f ms" % timed(lambda: fact(False), reps=3))
print("hypers+factorize with L^-1 (same launch)               %.3f ms" % timed(lambda: fact(True), reps=3))
