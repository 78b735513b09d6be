"""Shared helpers for the parity tests: build oracle objects from a golden fixture."""
import numpy as np

from oracle import bqo_oracle as orc


def rel_err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    scale = np.maximum(np.abs(b), 1e-300)
    return np.max(np.abs(a - b) / scale) if a.size else 0.0


def assert_close(a, b, rtol=1e-9, atol=0.0, what=""):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = ~(np.abs(a - b) <= atol + rtol * np.abs(b))
    bad &= ~(np.isnan(a) & np.isnan(b))
    if np.any(bad):
        idx = np.argwhere(bad)[0]
        raise AssertionError("%s: mismatch at %s: %r vs %r (rtol %g, atol %g, %d bad)" % (
            what, tuple(idx), a[tuple(idx)], b[tuple(idx)], rtol, atol, int(bad.sum())))


def spec_from_golden(g):
    kind = int(g["kind"])
    dim = g["points"].shape[1]
    return orc.KernelSpec(kind, dim, n_tasks=int(g["n_tasks"]),
                          same_correlation=bool(int(g["same_correlation"])))


def oracle_from_golden(g):
    spec = spec_from_golden(g)
    vno = g["var_noise_obs"] if "var_noise_obs" in g else None
    gp = orc.OracleGP(spec, g["points"], g["evaluations"], vno)
    dx = int(g["dx"])
    if spec.kind == orc.PRODUCT_TASKS:
        if "weights" in g:
            T = spec.n_tasks
            bq = orc.OracleBQ(gp, list(range(dx)), orc.WEIGHTED_UNIFORM_FINITE,
                              weights=g["weights"], domain_random=np.arange(T).reshape(T, 1))
        else:
            bq = orc.OracleBQ(gp, list(range(dx)), orc.UNIFORM_FINITE)
    else:
        bq = orc.OracleBQ(gp, list(range(dx)), orc.GAMMA)
    return spec, gp, bq


def cond_rtol(g, s, base=1e-9):
    """Tolerance for quantities that pass through Sigma^-1: two correct factorisations of the
    same matrix differ by about cond(Sigma) * eps, so the 1e-9 bar applies while
    cond(Sigma) * eps stays below it (true for every fixture except the reference's own
    test vector theta = [1, 5, 50, 9.6, -3, -0.1], whose Sigma has cond = 2.3e9)."""
    th = g["thetas"][s]
    cov = g["cov"][s].copy()
    if "var_noise_obs" in g:
        cov += np.diag(g["var_noise_obs"])
    cov += th[0] * np.eye(cov.shape[0])
    return max(base, 8.0 * np.linalg.cond(cov) * np.finfo(float).eps)
