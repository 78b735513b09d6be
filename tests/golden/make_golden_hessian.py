"""Hessian fixtures from the UNMODIFIED reference (build container only):

    python tests/golden/make_golden_hessian.py

* `hess_k`   gp.evaluate_hessian_cross_cov_respect_point (sbo/models/gp_fitting_gaussian.py:1174-1200
             -> kernels' hessian_respect_point, sbo/kernels/matern52.py:466-503)
* `hess_B`   bq.evaluate_hessian_cross_cov (sbo/numerical_tools/bayesian_quadrature.py:364-386)
* `hess_a`, `hess_b`  bq.compute_hessian_parameters_for_sample (:1241-1286)
* `hess_sample`       sbo.evaluate_hessian_sample (sbo/acquisition_functions/sbo.py:196-235)
on the seeded models of make_golden.py (same data, same parameter vectors).
Output: tests/golden/hessian_<case>.npz.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the import shim)
import ref_shim  # noqa: E402

warnings.filterwarnings("ignore")

from stratified_bayesian_optimization.models.gp_fitting_gaussian import GPFittingGaussian  # noqa: E402
from stratified_bayesian_optimization.numerical_tools.bayesian_quadrature import BayesianQuadrature  # noqa: E402
from stratified_bayesian_optimization.acquisition_functions.sbo import SBO  # noqa: E402
from stratified_bayesian_optimization.lib.constant import (  # noqa: E402
    PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME, SCALED_KERNEL,
    WEIGHTED_UNIFORM_FINITE, GAMMA, SAME_CORRELATION)


def hessians(gp, bq, sbo, thetas, query, xq, cands, zs, out):
    S = len(thetas)
    n = gp.data["points"].shape[0]
    dim = query.shape[1]
    dx = xq.shape[1]
    hk = np.zeros((S, query.shape[0], n, dim, dim))
    hB = np.zeros((S, xq.shape[0], n, dx, dx))
    ha = np.zeros((S, cands.shape[0], xq.shape[0], dx, dx))
    hb = np.zeros_like(ha)
    hs = np.zeros((S, cands.shape[0], xq.shape[0], len(zs), dx, dx))
    for s, th in enumerate(thetas):
        vn, mu, pk = th[0], th[1], th[2:]
        for q in range(query.shape[0]):
            hk[s, q] = gp.evaluate_hessian_cross_cov_respect_point(query[q:q + 1], gp.data["points"], pk)
        for p in range(xq.shape[0]):
            hB[s, p] = bq.evaluate_hessian_cross_cov(xq[p:p + 1], gp.data["points"], pk)
        for c in range(cands.shape[0]):
            for p in range(xq.shape[0]):
                mg.clear_bq(bq)
                h = bq.compute_hessian_parameters_for_sample(xq[p:p + 1], cands[c:c + 1], vn, mu, pk)
                ha[s, c, p], hb[s, c, p] = h["a"], h["b"]
                for iz, z in enumerate(zs):
                    mg.clear_bq(bq)
                    hs[s, c, p, iz] = sbo.evaluate_hessian_sample(xq[p:p + 1], cands[c:c + 1], z, vn, mu, pk)
    out.update(hess_k=hk, hess_B=hB, hess_a=ha, hess_b=hb, hess_sample=hs)


def case_product_weighted():
    g = dict(np.load(os.path.join(HERE, "product_weighted.npz")))
    T = int(g["n_tasks"])
    td = {"points": g["points"].tolist(), "evaluations": list(g["evaluations"]),
          "var_noise": list(g["var_noise_obs"])}
    gp = GPFittingGaussian([PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME], td,
                           [3, 2, T], bounds_domain=[[0, 1], [0, 1], list(range(T))],
                           type_bounds=[0, 0, 1], noise=True, **{SAME_CORRELATION: True})
    bq = BayesianQuadrature(gp, [0, 1], WEIGHTED_UNIFORM_FINITE,
                            parameters_distribution={"weights": g["weights"],
                                                     "domain_random": np.arange(T).reshape(T, 1)})
    sbo = SBO(bq, g["discretization"])
    out = {}
    hessians(gp, bq, sbo, g["thetas"], g["query"], g["xq"], g["cands"], g["zs"], out)
    return out


def case_scaled_gamma():
    g = dict(np.load(os.path.join(HERE, "scaled_gamma.npz")))
    td = {"points": g["points"].tolist(), "evaluations": list(g["evaluations"]), "var_noise": []}
    gp = GPFittingGaussian([SCALED_KERNEL, MATERN52_NAME], td, [4],
                           bounds_domain=[[0, 1], [0, 1], [0, 10], [0, 10]], noise=True,
                           define_samplers=False)
    bq = BayesianQuadrature(gp, [0, 1], GAMMA, parameters_distribution={"a": [2.0], "scale": [1.0]})
    sbo = SBO(bq, None)
    ref_shim.patch_gamma([g["wset"]])
    out = {}
    hessians(gp, bq, sbo, g["thetas"], g["query"], g["xq"], g["cands"], g["zs"], out)
    return out


def main():
    for name, fn in (("product_weighted", case_product_weighted), ("scaled_gamma", case_scaled_gamma)):
        out = fn()
        np.savez_compressed(os.path.join(HERE, "hessian_%s.npz" % name), **out)
        print(name, {k: v.shape for k, v in out.items()}, float(np.abs(out["hess_a"]).max()))


if __name__ == "__main__":
    main()
