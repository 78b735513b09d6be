"""scipy L-BFGS-B maxima of a(x) + Z b(x) on the CPU oracle at the C2 and C3 shapes of SURVEY 8d,
the reference's inner maximisation (sbo/acquisition_functions/sbo.py:237-366 through
Optimization.optimize, sbo/lib/optimization.py:87-155: fmin_l_bfgs_b(maxiter=10, factr=1e12) from
n_restarts_mc + 1 starts, then the 2^{d_x} box vertices).

    python tests/golden/make_lbfgs_fixture.py        # ~2-4 minutes on 8 cores

Stores, per workload, the maximum and argmax of every (candidate, hyper-sample, Z) tuple in
tests/golden/lbfgs_<cfg>.npz.  The inputs are not stored: they are regenerated from the seeds by
`workload()` below, which tests/test_gpu_inner_vs_lbfgs.py imports.  The objective is the
oracle's own a, b, grad a, grad b (OracleBQ.evaluate_quadrature_cross_cov /
evaluate_grad_quadrature_cross_cov), with the per-unit vectors of get_parameters_for_samples
computed once per unit instead of once per evaluation (identical values, 100x less time).
"""
import itertools
import multiprocessing as mp
import os
import sys

import numpy as np
from scipy.optimize import fmin_l_bfgs_b

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import bqo_oracle as orc  # noqa: E402

N_RESTARTS_MC, MAXITER_MC, FACTR_MC = 5, 10, 1e12  # sbo/services/bgo.py:43-44
# `--converged`: scipy's own defaults (factr=1e7, maxiter=15000; sbo/lib/optimization.py:87-155 passes
# them through when bgo's factr_mc / maxiter_mc are None) -- both optimisers then reach the same
# local maxima and the comparison isolates the objective from the termination rule
CONVERGED = "--converged" in sys.argv
if CONVERGED:
    MAXITER_MC, FACTR_MC = 15000, 1e7


X_2 = [0.25, 0.5, 0.75]
X_3 = [0.2, 0.4, 0.6, 0.8]
# problems/branin_mt/main.py:20-21, flattened in task order (task = 4 * index(x_2) + index(x_3))
PROBABILITIES = np.array([[0.0375, 0.0875, 0.0875, 0.0375], [0.0750, 0.1750, 0.1750, 0.0750],
                          [0.0375, 0.0875, 0.0875, 0.0375]]).reshape(-1)


def branin(u, v):
    """problems/branin_mt/main.py:24-27."""
    result = 10 + 10.0 * np.cos(u) * (1.0 - (1.0 / (8.0 * np.pi)))
    result = result + (v - 5.1 * (u ** 2) / (4.0 * np.pi * np.pi) + 5.0 * u / np.pi - 6) ** 2
    return result


def toy_example(x, task):
    """problems/branin_mt/main.py:29-50, vectorised: f(x1, x4, task) =
    branin(15 x1 - 5, 15 x2) * branin(15 x3 - 5, 15 x4) with (x2, x3) fixed by the task."""
    task = np.asarray(task, dtype=int)
    x2 = np.array(X_2)[task // 4]
    x3 = np.array(X_3)[task % 4]
    return branin(15 * x[:, 0] - 5, 15 * x2) * branin(15 * x3 - 5, 15 * x[:, 1])


def workload(cfg):
    """Seeded inputs of SURVEY 8d.  C2: Product(Matern52(d_x=2), Tasks(T=12, same_correlation)),
    weighted_uniform_finite, n=200.  C3: Scaled(Matern52) d=6 (d_x=4, d_w=2, gamma W, M_W=10),
    n=2048."""
    if cfg == "c2":
        T, n, dx = 12, 200, 2
        rng = np.random.RandomState(1)
        x = rng.uniform(0, 1, (n, dx))
        t = rng.randint(0, T, n)
        pts = np.concatenate([x, t.reshape(-1, 1).astype(float)], 1)
        y = toy_example(x, t)
        y = (y - np.mean(y)) / np.std(y)  # objective values are O(1e4): standardised evaluations
        r2 = np.random.RandomState(2)
        base = np.array([0.0, 0.0, 0.3, 0.3, 0.0, -1.0])
        thetas = base[None, :] + np.concatenate(
            [np.zeros((4, 2)), 0.05 * base[None, 2:4] * r2.normal(0, 1, (4, 2)),
             0.05 * r2.normal(0, 1, (4, 2))], 1)
        thetas[:, 0] = 1e-6
        thetas[:, 1] = np.mean(y)
        weights = PROBABILITIES.copy()
        r5 = np.random.RandomState(5)
        cx = r5.uniform(0, 1, (10, dx))
        cands = np.array([np.concatenate([cx[i], [j]]) for i in range(10) for j in range(T)])
        zs = np.random.RandomState(3).normal(0, 1, 100)
        starts = np.random.RandomState(6).uniform(0, 1, (N_RESTARTS_MC + 1, dx))
        return dict(kind=orc.PRODUCT_TASKS, pts=pts, y=y, thetas=thetas, weights=weights, n_tasks=T,
                    same_correlation=True, dx=dx, cands=cands, zs=zs, starts=starts, wset=None,
                    lower=np.zeros(dx), upper=np.ones(dx),
                    # the tuples that are compared: candidates x hyper-samples x Z
                    cand_idx=np.arange(0, 120, 5), hyper_idx=np.arange(2), z_idx=np.arange(0, 100, 4))
    if cfg == "c3":
        n, dx, dw = 2048, 4, 2
        rng = np.random.RandomState(0)
        pts = np.concatenate([rng.uniform(0, 1, (n, dx)), rng.gamma(2.0, 1.0, (n, dw))], 1)
        y = np.sin(pts.sum(1)) + 0.01 * rng.normal(0, 1, n)
        r2 = np.random.RandomState(2)
        base = np.array([1e-4, 0.0] + dx * [0.3] + dw * [2.0] + [1.0])
        thetas = base[None, :] * (1.0 + 0.05 * r2.normal(0, 1, (20, base.size)))
        thetas[:, 1] = 0.05 * r2.normal(0, 1, 20)
        zs = np.random.RandomState(3).normal(0, 1, 100)
        wset = np.random.RandomState(4).gamma(2.0, 1.0, (8, 10, dw))[0]
        r5 = np.random.RandomState(5)
        cands = np.concatenate([r5.uniform(0, 1, (10000, dx)), r5.gamma(2.0, 1.0, (10000, dw))], 1)[:32]
        starts = np.random.RandomState(6).uniform(0, 1, (N_RESTARTS_MC + 1, dx))
        return dict(kind=orc.SCALED_MATERN52, pts=pts, y=y, thetas=thetas, weights=None, n_tasks=0,
                    same_correlation=False, dx=dx, cands=cands, zs=zs, starts=starts, wset=wset,
                    lower=np.zeros(dx), upper=np.ones(dx),
                    cand_idx=np.arange(0, 10), hyper_idx=np.arange(2),
                    z_idx=np.arange(0, 100, 10 if CONVERGED else 2))
    raise KeyError(cfg)


def oracle_objects(wl):
    spec = orc.KernelSpec(wl["kind"], wl["pts"].shape[1], n_tasks=wl["n_tasks"],
                          same_correlation=wl["same_correlation"])
    gp = orc.OracleGP(spec, wl["pts"], wl["y"])
    dx = wl["dx"]
    if wl["kind"] == orc.PRODUCT_TASKS:
        T = wl["n_tasks"]
        bq = orc.OracleBQ(gp, list(range(dx)), orc.WEIGHTED_UNIFORM_FINITE, weights=wl["weights"],
                          domain_random=np.arange(T).reshape(T, 1))
    else:
        bq = orc.OracleBQ(gp, list(range(dx)), orc.GAMMA)
    return spec, gp, bq


class UnitObjective(object):
    """a, b, grad a, grad b of one (candidate, hyper-sample) unit, the arithmetic of
    OracleBQ.compute_parameters_for_sample / compute_gradient_parameters_for_sample with the unit's
    vectors (alpha, Sigma^-1 gamma, denominator) computed once."""

    def __init__(self, bq, cand, th, w):
        self.bq, self.cand, self.th, self.w = bq, cand.reshape(1, -1), th, w
        ap = bq.get_parameters_for_samples(self.cand, th[0], th[1], th[2:])
        self.alpha = ap["solve"]
        self.s2 = ap["solve_2"][:, 0]
        self.den = ap["denominator"][0]
        self.X = bq.gp.points

    def ab(self, x):
        bq, th = self.bq, self.th
        x = np.asarray(x, dtype=float).reshape(1, -1)
        B = bq.evaluate_quadrature_cross_cov(x, self.X, th[2:], w=self.w)
        gB = bq.evaluate_grad_quadrature_cross_cov(x, self.X, th[2:], w=self.w)
        a = th[1] + np.dot(B, self.alpha)
        ga = np.dot(gB, self.alpha)
        if self.den != 0:
            bn = bq.evaluate_quadrature_cross_cov(x, self.cand, th[2:], w=self.w)[0]
            gbn = bq.evaluate_grad_quadrature_cross_cov(x, self.cand, th[2:], w=self.w)[:, 0]
            b = (bn - np.dot(B, self.s2)) / self.den
            gb = (gbn - np.dot(gB, self.s2)) / self.den
        else:
            b, gb = 0.0, np.zeros_like(ga)
        return a, b, ga, gb


def lbfgs_unit(job):
    cfg, c, k = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    wl = _WL[cfg]
    _, gp, bq = _OBJ[cfg]
    th = wl["thetas"][k]
    obj = UnitObjective(bq, wl["cands"][c], th, wl["wset"])
    bounds = list(zip(wl["lower"], wl["upper"]))
    vertex = [np.array(p) for p in itertools.product(*bounds)]
    # a, b at the vertices do not depend on Z
    vab = [obj.ab(v)[:2] for v in vertex]
    out_max, out_arg, nfev = [], [], []
    for iz in wl["z_idx"]:
        z = wl["zs"][iz]
        count = [0]

        def f(x):
            count[0] += 1
            a, b, _, _ = obj.ab(x)
            return -(a + z * b)

        def g(x):
            _, _, ga, gb = obj.ab(x)
            return -(ga + z * gb)

        best, arg = -np.inf, None
        for start in wl["starts"]:
            opt = fmin_l_bfgs_b(f, start, fprime=g, bounds=bounds, factr=FACTR_MC, maxiter=MAXITER_MC)
            if -opt[1] > best:
                best, arg = -opt[1], opt[0]
        vals = [a + z * b for a, b in vab]
        if np.max(vals) > best:
            best, arg = np.max(vals), vertex[int(np.argmax(vals))]
        out_max.append(best)
        out_arg.append(arg)
        nfev.append(count[0])
    return c, k, np.array(out_max), np.array(out_arg), np.array(nfev)


_WL, _OBJ = {}, {}


def main():
    only = set(a for a in sys.argv[1:] if not a.startswith("--"))
    for cfg in ("c2", "c3"):
        if only and cfg not in only:
            continue
        wl = workload(cfg)
        _WL[cfg] = wl
        _OBJ[cfg] = oracle_objects(wl)
        for k in wl["hyper_idx"]:  # factor once per hyper-sample in the parent (inherited by fork)
            th = wl["thetas"][k]
            _OBJ[cfg][1].chol_solve(th[0], th[1], th[2:])
        jobs = [(cfg, int(c), int(k)) for c in wl["cand_idx"] for k in wl["hyper_idx"]]
        with mp.get_context("fork").Pool(processes=min(8, os.cpu_count() or 1)) as pool:
            res = pool.map(lbfgs_unit, jobs)
        C, K, Z = len(wl["cand_idx"]), len(wl["hyper_idx"]), len(wl["z_idx"])
        mx = np.zeros((C, K, Z))
        arg = np.zeros((C, K, Z, wl["dx"]))
        nf = np.zeros((C, K, Z), dtype=int)
        ci = {int(c): i for i, c in enumerate(wl["cand_idx"])}
        ki = {int(k): i for i, k in enumerate(wl["hyper_idx"])}
        for c, k, m, a, f in res:
            mx[ci[c], ki[k]], arg[ci[c], ki[k]], nf[ci[c], ki[k]] = m, a, f
        name = "lbfgs_%s_converged.npz" if CONVERGED else "lbfgs_%s.npz"
        np.savez_compressed(os.path.join(HERE, name % cfg), scipy_max=mx, scipy_arg=arg,
                            scipy_nfev=nf, cand_idx=wl["cand_idx"], hyper_idx=wl["hyper_idx"],
                            z_idx=wl["z_idx"])
        print(cfg, "tuples", mx.size, "mean f evaluations per tuple (6 runs)", nf.mean())


if __name__ == "__main__":
    main()
