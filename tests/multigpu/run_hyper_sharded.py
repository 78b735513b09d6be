"""Multi-GPU check (run under torchrun on N >= 2 GPUs, not part of pytest):
hyper-samples sharded across ranks + all_gather / rank-ordered sum reproduce the single-GPU
acquisition values, gradients and the argmax candidate; every rank ends with identical bits.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/multigpu/run_hyper_sharded.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import bench  # noqa: E402
from bayesian_quadrature_optimization_b200 import engine as E, _cabi, distributed as D  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl")
wl = bench.make_workload()
n_sub = 512  # a smaller training set keeps the check short
eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, bench.DX + bench.DW)
eng.set_data(wl["pts"][:n_sub], wl["y"][:n_sub])
thetas = wl["thetas"][:8]
lo, up = wl["lower"], wl["upper"]


def make_bq(th):
    return E.BQEngine(eng.hypers(th), bench.DX, wsets=wl["wsets"], lower=lo, upper=up)


cands, zs, starts = wl["cands"][:12], wl["zs"][:10], wl["starts"]
res = D.evaluate_candidates_hyper_sharded(make_bq, thetas, cands, zs, starts)
# every rank holds the same bits
mine = torch.cat([res["evaluations"], res["gradient"].reshape(-1)])
bufs = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(bufs, mine)
same = all(torch.equal(bufs[0], b) for b in bufs)
ok = same
if rank == 0:
    full = make_bq(thetas).evaluate_candidates(cands, zs, starts)
    ev = (res["evaluations"] - full["evaluations"]).abs().max().item()
    gv = (res["gradient"] - full["gradient"]).abs().max().item()
    am = int(torch.argmax(full["evaluations"]).item())
    ok = same and ev < 1e-10 and gv < 1e-10 and am == res["argmax"]  # re-association of the means
    print("hyper-sharded over %d GPUs: identical bits on all ranks: %s; max |dv| %.2e, max |dg| %.2e "
          "against one GPU; argmax %d == %d: %s" % (world, same, ev, gv, res["argmax"], am,
                                                     "OK" if ok else "FAILED"))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
