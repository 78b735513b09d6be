"""CPU restatement of the device inner maximiser (TEST INFRASTRUCTURE ONLY).

The reference maximises x -> a(x) + Z b(x) with scipy's L-BFGS-B (Fortran code of a third-party
dependency: scipy==0.19.0, `fmin_l_bfgs_b`, called from Optimization.optimize,
sbo/lib/optimization.py:87-155, with maxiter=10 and factr=1e12 from sbo/services/bgo.py:43-44),
started from n_restarts+1 points and compared with the 2^d box vertices
(SBO.evaluate_sbo_by_sample, sbo/acquisition_functions/sbo.py:237-366).

The CUDA kernel `inner_maximize_kernel` (csrc/bqo_stage_c.cu) runs a projected BFGS with Armijo
backtracking under the same iteration cap and the same two L-BFGS-B stopping rules
(relative decrease <= factr*eps, projected gradient <= pgtol).  This file restates THAT
iteration in numpy, statement for statement, so the device trajectory can be checked on the CPU.
The value the optimiser returns is then checked against the oracle's own evaluation of
a + Z b at the returned point, and against scipy's L-BFGS-B maxima (tests/test_inner_opt.py).
"""
from __future__ import annotations

import itertools

import numpy as np

DBL_EPSILON = np.finfo(float).eps


def pbfgs_maximize(fg, x0, lower, upper, maxiter=10, max_ls=20, factr=1e12, pgtol=1e-5,
                   f0g0=None):
    """Maximise F from x0 inside the box.  fg(x, eval_index) -> (F, grad F).

    Returns (F_best, x_best, n_evals).  eval_index counts evaluations of this run (0 = start),
    so that the caller can pick the W sample set the device kernel would use.
    """
    lower = np.asarray(lower, dtype=float)
    upper = np.asarray(upper, dtype=float)
    d = len(lower)
    x = np.minimum(np.maximum(np.asarray(x0, dtype=float), lower), upper)
    n_evals = 0
    if f0g0 is None:
        F, gF = fg(x, 0)
        n_evals += 1
    else:
        F, gF = f0g0
    eval_idx = 1
    phi = -F
    g = -np.asarray(gF, dtype=float)
    H = np.eye(d)
    first_update = True
    c1 = 1e-4
    for it in range(maxiter):
        t = np.minimum(np.maximum(x - g, lower), upper)
        pgn = np.max(np.abs(x - t))
        if pgn <= pgtol:
            break
        act = ((x <= lower) & (g > 0.0)) | ((x >= upper) & (g < 0.0))
        gf = np.where(act, 0.0, g)
        dd = np.where(gf == 0.0, 0.0, -(H @ gf))
        slope = float(dd @ g)
        gnorm2 = float(gf @ gf)
        if not (slope < 0.0):
            H = np.eye(d)
            first_update = True
            slope = -gnorm2
            dd = -gf
            if gnorm2 == 0.0:
                break
        dnorm = np.sqrt(float(dd @ dd))
        step = min(1.0, 1.0 / dnorm) if (it == 0 or first_update) else 1.0
        accepted = False
        for _ in range(max_ls):
            xn = np.minimum(np.maximum(x + step * dd, lower), upper)
            decr = float(g @ (xn - x))
            Fn, gFn = fg(xn, eval_idx)
            eval_idx += 1
            n_evals += 1
            phin = -Fn
            if phin <= phi + c1 * decr:
                accepted = True
                gn = -np.asarray(gFn, dtype=float)
                break
            denom_q = 2.0 * (phin - phi - decr)
            shrink = (-decr / denom_q) if denom_q > 0.0 else 0.5
            shrink = min(max(shrink, 0.1), 0.5)
            if shrink != shrink:
                shrink = 0.5
            step *= shrink
        if not accepted:
            break
        sv = xn - x
        yv = gn - g
        sy = float(sv @ yv)
        yy = float(yv @ yv)
        ss = float(sv @ sv)
        phi_old = phi
        upd = sy > 1e-10 * np.sqrt(ss) * np.sqrt(yy)
        if upd:
            Hc = H
            if first_update:
                Hc = (sy / yy) * np.eye(d)
            rho = 1.0 / sy
            Hy = Hc @ yv
            yHy = float(yv @ Hy)
            cc = rho * (1.0 + rho * yHy)
            H = Hc - rho * (np.outer(sv, Hy) + np.outer(Hy, sv)) + cc * np.outer(sv, sv)
            first_update = False
        x, g, phi = xn, gn, phin
        denom = max(abs(phi_old), abs(phin), 1.0)
        if (phi_old - phin) <= factr * DBL_EPSILON * denom:
            break
    return -phi, x, n_evals


def box_vertices(lower, upper):
    """itertools.product(*bounds_x) order (sbo.py:266-269)."""
    return [np.array(p, dtype=float) for p in itertools.product(*zip(lower, upper))]


def inner_maximize(ab, z, starts, lower, upper, wset=0, use_vertices=True, **opt):
    """max over starts (+ vertices) of a(x) + z b(x).

    ab(x, wset) -> (a, b, grad_a, grad_b).  Mirrors the three phases of the CUDA kernel; every
    evaluation of a unit uses the W set of its candidate (`wset` = candidate index % n_wsets).
    Returns (max, argmax, n_evals).
    """
    best = None
    arg = None
    n_evals = 0
    for x0 in np.asarray(starts, dtype=float):
        x0c = np.minimum(np.maximum(x0, lower), upper)
        a0, b0, ga0, gb0 = ab(x0c, wset)

        def fg(x, e):
            a, b, ga, gb = ab(x, wset)
            return a + z * b, ga + z * gb

        F, x, ne = pbfgs_maximize(fg, x0c, lower, upper, f0g0=(a0 + z * b0, ga0 + z * gb0), **opt)
        n_evals += ne
        if best is None or F > best:
            best, arg = F, x
    if use_vertices:
        vb, vx = None, None
        for v in box_vertices(lower, upper):
            a, b, _, _ = ab(v, wset)
            val = a + z * b
            if vb is None or val > vb:
                vb, vx = val, v
        if vb > best:
            best, arg = vb, vx
    return best, arg, n_evals
