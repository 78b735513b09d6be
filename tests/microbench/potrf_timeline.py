"""Critical-path timeline of the persistent Cholesky (bqo_potrf_profile): per column j the stamps
of matrix 0's diagonal task and of tile (j+1, j).  Tuning aid; run under gpurun."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from bayesian_quadrature_optimization_b200 import engine as E, _cabi  # noqa: E402
from bayesian_quadrature_optimization_b200._cabi import ptr, check  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 20
wl = bench.make_workload()
eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, bench.DX + bench.DW)
eng.set_data(wl["pts"], wl["y"])
n = eng.n
hb = eng.hypers(wl["thetas"][:S])
Sigma = hb.cov(add_noise=True)
nb = (n + 63) // 64
st = torch.cuda.current_stream().cuda_stream
info = torch.zeros(S, dtype=torch.int32, device=eng.device)
work = torch.empty(int(eng.lib.bqo_potrf_workspace_bytes(n, S)) // 8 + 2, dtype=torch.float64, device=eng.device)
for rep in range(3):
    A = Sigma.clone()
    prof = torch.zeros(nb * 2 * 8, dtype=torch.int64, device=eng.device)
    torch.cuda.synchronize()
    check(eng.lib.bqo_potrf_profile(ptr(A), n, S, ptr(info), ptr(work), ptr(prof), st), "profile")
    torch.cuda.synchronize()
p = prof.cpu().numpy().reshape(nb, 2, 8).astype(np.float64)
t0 = p[0, 0, 0]
p = np.where(p > 0, (p - t0) / 1e3, np.nan)  # microseconds since the first task started
print("S=%d n=%d: total %.1f us" % (S, n, np.nanmax(p)))
print("diag task (us): start | gemm done | T ready | factor+inverse done | released   || tile (j+1,j): start | gemm done | T ready | W flag seen | W loaded | product | released")
for j in range(nb):
    d, o = p[j, 0], p[j, 1]
    print("j=%2d  %8.1f %8.1f %8.1f %8.1f %8.1f   ||  %8.1f %8.1f %8.1f %8.1f %8.1f %8.1f %8.1f" % (
        j, d[0], d[1], d[2], d[3], d[4], o[0], o[1], o[2], o[3], o[4], o[5], o[6]))
d = p[:, 0]
print("mean factor+inverse %.1f us; mean T-ready -> release %.1f us; mean release(j) -> release(j+1) %.1f us" % (
    np.nanmean(d[:, 3] - d[:, 2]), np.nanmean(d[:, 4] - d[:, 2]), np.nanmean(np.diff(d[:, 4]))))
o = p[:-1, 1]
print("tile (j+1,j): W flag seen -> released %.1f us (W load %.1f, product %.1f, write+fence %.1f); diag release -> flag seen %.1f us" % (
    np.nanmean(o[:, 6] - o[:, 3]), np.nanmean(o[:, 4] - o[:, 3]), np.nanmean(o[:, 5] - o[:, 4]),
    np.nanmean(o[:, 6] - o[:, 5]), np.nanmean(o[:, 3] - d[:-1, 4])))
print("next diag: tile released -> diag gemm done %.1f us -> T ready %.1f us" % (
    np.nanmean(d[1:, 1] - o[:, 6]), np.nanmean(d[1:, 2] - d[1:, 1])))
