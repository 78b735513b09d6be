"""The device's inner maximiser against scipy L-BFGS-B on the real objective.

The reference maximises a(x) + Z b(x) with fmin_l_bfgs_b(maxiter=10, factr=1e12) from
n_restarts_mc + 1 starts and then compares with the box vertices (sbo/acquisition_functions/
sbo.py:237-366, sbo/lib/optimization.py:87-155).  tests/golden/make_lbfgs_fixture.py ran exactly
that on the CPU oracle for 1200 (C2 shape) and 1000 (C3 shape) (candidate, hyper-sample, Z)
tuples; here the device maximiser runs on the same tuples from the same starts.

Two settings:

* bgo's defaults maxiter=10, factr=1e12 (sbo/services/bgo.py:43-44).  Both optimisers are cut
  off long before convergence (factr * eps = 2.2e-4 relative decrease, or the 10-iteration cap), so
  their maxima differ by the progress of the last iteration.  Measured on the B200
  (profiles/r02_inner_vs_lbfgs.md): device - scipy between -5.7e-5 and +2.4e-5 relative, median
  -9e-7 at C3; the acquisition value (mean over hyper-samples and Z of the maxima,
  sbo.py:640-650) differs by at most 2.8e-6 relative.  Asserted: device >= scipy - 1e-5 |scipy| in
  >= 99 % of the tuples, >= scipy - 1e-4 |scipy| in all of them (half the termination tolerance),
  acquisition value within 1e-5 relative.
* scipy's own defaults (maxiter=15000, factr=1e7), where both reach the local maxima: the maxima
  agree to 1e-7 relative in >= 99 % of the tuples -- the objective is the same function and the
  difference above is the termination rule, not the arithmetic.

The measured distributions are written to gpurun_out/inner_vs_lbfgs_<cfg>[_converged].json.
"""
import json
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN_DIR, ROOT
from tests.golden.make_lbfgs_fixture import workload, oracle_objects, UnitObjective

pytestmark = pytest.mark.gpu


def _device_maxima(wl, fx, maxiter=10, factr=1e12):
    from bayesian_quadrature_optimization_b200 import engine as E, _cabi
    from oracle import bqo_oracle as orc
    kind = _cabi.KIND_PRODUCT_TASKS if wl["kind"] == orc.PRODUCT_TASKS else _cabi.KIND_SCALED_MATERN52
    eng = E.GPEngine(kind, wl["pts"].shape[1], wl["n_tasks"], wl["same_correlation"])
    eng.set_data(wl["pts"], wl["y"])
    thetas = wl["thetas"][fx["hyper_idx"]]
    hb = eng.hypers(thetas)
    bq = E.BQEngine(hb, wl["dx"], weights=wl["weights"],
                    wsets=None if wl["wset"] is None else wl["wset"][None],
                    lower=wl["lower"], upper=wl["upper"])
    cands = wl["cands"][fx["cand_idx"]]
    ub = bq.setup(cands)
    zs = wl["zs"][fx["z_idx"]]
    mx, arg, n_evals = bq.inner_maximize(ub, zs, wl["starts"], maxiter=maxiter, factr=factr,
                                         count_evals=True)
    S = len(fx["hyper_idx"])
    C = len(fx["cand_idx"])
    mx = mx.cpu().numpy().reshape(C, S, len(zs))
    arg = arg.cpu().numpy().reshape(C, S, len(zs), wl["dx"])
    return mx, arg, n_evals.cpu().numpy().reshape(C, S)


@pytest.mark.parametrize("cfg", ["c2", "c3"])
def test_device_maxima_are_as_good_as_scipy_lbfgsb(cfg):
    with np.load(os.path.join(GOLDEN_DIR, "lbfgs_%s.npz" % cfg)) as f:
        fx = {k: f[k] for k in f.files}
    wl = workload(cfg)
    dev_max, dev_arg, n_evals = _device_maxima(wl, fx)
    sc = fx["scipy_max"]
    assert dev_max.shape == sc.shape and sc.size >= 1000
    diff = dev_max - sc
    rel = diff / np.abs(sc)
    ok = dev_max >= sc - 1e-6 * np.abs(sc)
    frac_ok = float(np.mean(ok))
    # acquisition value per candidate: mean over hyper-samples and Z of the maxima
    acq_dev, acq_sc = dev_max.mean(axis=(1, 2)), sc.mean(axis=(1, 2))
    acq_rel = (acq_dev - acq_sc) / np.abs(acq_sc)
    # the device's value is a true value of the objective: re-evaluate it on the CPU oracle at
    # the device's argmax for a sample of the tuples
    _, gp, bq = oracle_objects(wl)
    checked = 0
    for ci in range(0, sc.shape[0], max(1, sc.shape[0] // 3)):
        c, k = int(fx["cand_idx"][ci]), int(fx["hyper_idx"][0])
        obj = UnitObjective(bq, wl["cands"][c], wl["thetas"][k], wl["wset"])
        for zi in range(0, sc.shape[2], max(1, sc.shape[2] // 4)):
            z = wl["zs"][fx["z_idx"][zi]]
            a, b, _, _ = obj.ab(dev_arg[ci, 0, zi])
            assert abs((a + z * b) - dev_max[ci, 0, zi]) <= 1e-8 * max(1.0, abs(dev_max[ci, 0, zi]))
            checked += 1
    pct = lambda q: float(np.percentile(rel, q))
    summary = {
        "cfg": cfg, "tuples": int(sc.size), "checked_on_oracle": checked,
        "frac_device_ge_scipy_minus_1e-6_rel": frac_ok,
        "frac_device_strictly_better_by_1e-9_rel": float(np.mean(rel > 1e-9)),
        "frac_device_worse_by_1e-9_rel": float(np.mean(rel < -1e-9)),
        "rel_diff_percentiles": {"min": float(rel.min()), "p1": pct(1), "p5": pct(5), "p50": pct(50),
                                 "p95": pct(95), "p99": pct(99), "max": float(rel.max())},
        "abs_diff_min_max": [float(diff.min()), float(diff.max())],
        "acq_value_rel_diff_min_max": [float(acq_rel.min()), float(acq_rel.max())],
        "acq_value_abs_diff_min_max": [float((acq_dev - acq_sc).min()), float((acq_dev - acq_sc).max())],
        "device_evals_per_tuple": float(n_evals.sum() / sc.size),
        "scipy_f_evals_per_tuple": float(fx["scipy_nfev"].mean()) + 2 ** wl["dx"],
    }
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "inner_vs_lbfgs_%s.json" % cfg), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary))
    summary["frac_device_ge_scipy_minus_1e-5_rel"] = float(np.mean(rel >= -1e-5))
    with open(os.path.join(out_dir, "inner_vs_lbfgs_%s.json" % cfg), "w") as f:
        json.dump(summary, f, indent=1)
    assert np.mean(rel >= -1e-5) >= 0.99, summary
    assert rel.min() >= -1e-4, summary           # factr * eps = 2.2e-4 is where both stop trying
    # the acquisition value the user sees
    assert np.max(np.abs(acq_rel)) <= 1e-5, summary


@pytest.mark.parametrize("cfg", ["c2", "c3"])
def test_device_maxima_equal_scipy_when_both_converge(cfg):
    with np.load(os.path.join(GOLDEN_DIR, "lbfgs_%s_converged.npz" % cfg)) as f:
        fx = {k: f[k] for k in f.files}
    import tests.golden.make_lbfgs_fixture as gen
    old = gen.CONVERGED
    gen.CONVERGED = True     # the converged fixture of C3 uses a sparser Z grid
    try:
        wl = workload(cfg)
    finally:
        gen.CONVERGED = old
    assert np.array_equal(wl["z_idx"], fx["z_idx"])
    dev_max, _, n_evals = _device_maxima(wl, fx, maxiter=15000, factr=1e7)
    sc = fx["scipy_max"]
    rel = (dev_max - sc) / np.abs(sc)
    summary = {
        "cfg": cfg, "tuples": int(sc.size), "setting": "maxiter=15000, factr=1e7 (scipy defaults)",
        "frac_abs_rel_diff_le_1e-7": float(np.mean(np.abs(rel) <= 1e-7)),
        "frac_abs_rel_diff_le_1e-9": float(np.mean(np.abs(rel) <= 1e-9)),
        "rel_diff_percentiles": {"min": float(rel.min()), "p1": float(np.percentile(rel, 1)),
                                 "p50": float(np.percentile(rel, 50)),
                                 "p99": float(np.percentile(rel, 99)), "max": float(rel.max())},
        "acq_value_rel_diff_min_max": [float(v) for v in (
            ((dev_max.mean(axis=(1, 2)) - sc.mean(axis=(1, 2))) / np.abs(sc.mean(axis=(1, 2)))).min(),
            ((dev_max.mean(axis=(1, 2)) - sc.mean(axis=(1, 2))) / np.abs(sc.mean(axis=(1, 2)))).max())],
        "device_evals_per_tuple": float(n_evals.sum() / sc.size),
        "scipy_f_evals_per_tuple": float(fx["scipy_nfev"].mean()) + 2 ** wl["dx"],
    }
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "inner_vs_lbfgs_%s_converged.json" % cfg), "w") as f:
        json.dump(summary, f, indent=1)
    print(json.dumps(summary))
    assert np.mean(np.abs(rel) <= 1e-7) >= 0.99, summary
