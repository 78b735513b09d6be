"""Multi-GPU check of the public API (run under torchrun on N >= 2 GPUs, not part of pytest):
SBO.evaluate_mc_bayesian_candidate_points_no_restarts and SBO.optimize with `distributed = True`
(hyper-samples sharded over the ranks, one all_gather + rank-ordered sum per evaluation) against
the single-GPU path, and bit-identical results on all ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29534 tests/multigpu/run_sbo_distributed.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian  # noqa: E402
from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import BayesianQuadrature  # noqa: E402
from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO  # noqa: E402
from bayesian_quadrature_optimization_b200.lib import constant as K  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

rng = np.random.RandomState(1)
n, T = 300, 4
x = rng.uniform(0, 1, (n, 2))
t = rng.randint(0, T, n)
pts = np.concatenate([x, t.reshape(-1, 1).astype(float)], 1)
y = np.sin(3 * x[:, 0]) + np.cos(2 * x[:, 1]) + 0.3 * t + rng.normal(0, 0.05, n)
td = {"points": pts.tolist(), "evaluations": list(y), "var_noise": []}
thetas = np.array([[1e-4, 0.0, 0.3, 0.4, 0.0, -1.0]]) * np.ones((8, 1)) + \
    np.concatenate([np.zeros((8, 2)), 0.03 * rng.normal(0, 1, (8, 4))], 1)


def model(distributed):
    gp = GPFittingGaussian([K.PRODUCT_KERNELS_SEPARABLE, K.MATERN52_NAME, K.TASKS_KERNEL_NAME], td,
                           [3, 2, T], bounds_domain=[[0, 1], [0, 1], list(range(T))],
                           type_bounds=[0, 0, 1], kernel_values=list(thetas[0][2:]),
                           define_samplers=False, device="cuda:%d" % local,
                           **{K.SAME_CORRELATION: True})
    gp.samples_parameters = [th.copy() for th in thetas]
    bq = BayesianQuadrature(gp, [0, 1], K.UNIFORM_FINITE, parameters_distribution={K.TASKS: T})
    sbo = SBO(bq)
    sbo.distributed = distributed
    return sbo


cands = np.concatenate([rng.uniform(0, 1, (6, 2)), rng.randint(0, T, (6, 1)).astype(float)], 1)
ok = True
for compute_max_mean in (False, True):
    np.random.seed(5)
    d = model(True).evaluate_mc_bayesian_candidate_points_no_restarts(
        cands, 8, 6, n_restarts=3, compute_max_mean=compute_max_mean, compute_gradient=True,
        factr=1e12, maxiter=10)
    np.random.seed(5)
    s = model(False).evaluate_mc_bayesian_candidate_points_no_restarts(
        cands, 8, 6, n_restarts=3, compute_max_mean=compute_max_mean, compute_gradient=True,
        factr=1e12, maxiter=10)
    ev = np.max(np.abs(d["evaluations"] - s["evaluations"]))
    gv = np.max(np.abs(d["gradient"] - s["gradient"]))
    mine = torch.tensor(np.concatenate([d["evaluations"], d["gradient"].reshape(-1)]), device="cuda")
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine)
    same = all(torch.equal(bufs[0], b) for b in bufs)
    good = same and ev < 1e-11 and gv < 1e-10 and \
        int(np.argmax(d["evaluations"])) == int(np.argmax(s["evaluations"]))
    ok = ok and good
    if rank == 0:
        print("evaluate (max_mean=%s) over %d GPUs: identical bits %s, max |dv| %.2e, max |dg| %.2e: %s"
              % (compute_max_mean, world, same, ev, gv, "OK" if good else "FAILED"))
# the optimiser end to end: every rank must return the same solution bits
np.random.seed(9)
sol = model(True).optimize(monte_carlo=True, n_samples=4, n_restarts_mc=3, n_restarts=4,
                           n_samples_parameters=4, start_new_chain=False, maxepoch=3, start_ei=False,
                           default_n_samples=6, default_n_samples_parameters=8, factr=1e12, maxiter=10)
mine = torch.tensor(np.concatenate([sol["solution"], [sol["optimal_value"]]]), device="cuda")
bufs = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(bufs, mine)
same = all(torch.equal(bufs[0], b) for b in bufs)
ok = ok and same and np.all(np.isfinite(sol["solution"]))
if rank == 0:
    print("optimize over %d GPUs: identical solution on all ranks %s (%s, value %.6g): %s"
          % (world, same, sol["solution"], sol["optimal_value"], "OK" if ok else "FAILED"))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
