"""GPU parity tests: the CUDA path (through the C ABI) against the fixtures generated from the
unmodified reference and against the numpy oracle on seeded inputs.

Tolerance: north star = 1e-9 relative in FP64 for posterior mean/variance, log-likelihood and
acquisition values; each assertion states its own (tighter where the arithmetic allows).
"""
import numpy as np
import pytest

from oracle import bqo_oracle as orc
from oracle import inner_opt
from tests.helpers import assert_close, cond_rtol, oracle_from_golden

pytestmark = pytest.mark.gpu

CASES = ["product_uniform", "product_weighted", "scaled_gamma"]
GP_CASES = CASES + ["matern52_plain"]  # the last one has no quadrature part


def _engine(g):
    from bayesian_quadrature_optimization_b200 import engine as E
    eng = E.GPEngine(int(g["kind"]), g["points"].shape[1], int(g["n_tasks"]),
                     bool(int(g["same_correlation"])))
    eng.set_data(g["points"], g["evaluations"], g["var_noise_obs"] if "var_noise_obs" in g else None)
    return eng


def _bq(g, hb):
    from bayesian_quadrature_optimization_b200 import engine as E
    dx = int(g["dx"])
    lo = np.zeros(dx)
    up = np.ones(dx) * (100.0 if g["points"][:, 0].max() > 2 else 1.0)
    bqv = float(np.mean(g["var_noise_obs"])) if "var_noise_obs" in g else None
    return E.BQEngine(hb, dx, weights=g["weights"] if "weights" in g else None,
                      wsets=g["wset"] if "wset" in g else None, lower=lo, upper=up,
                      bq_var_noise=bqv), lo, up


@pytest.mark.parametrize("name", GP_CASES)
def test_cov_and_param_gradients(golden, name):
    g = golden(name)
    eng = _engine(g)
    hb = eng.hypers(g["thetas"])
    assert_close(hb.cov().cpu().numpy(), g["cov"], rtol=1e-13, what="cov")
    assert_close(hb.grad_cov().cpu().numpy(), g["grad_cov"], rtol=1e-10, atol=1e-15, what="grad_cov")


@pytest.mark.parametrize("name", GP_CASES)
def test_cholesky_loglik_and_gradient(golden, name):
    g = golden(name)
    eng = _engine(g)
    hb = eng.hypers(g["thetas"]).factorize()
    assert np.all(hb.info == 0) and np.all(hb.jitter == 0.0)
    assert_close(hb.L.cpu().numpy(), g["chol"], rtol=1e-10, atol=1e-14, what="chol")
    assert_close(hb.log_likelihood().cpu().numpy(), g["llh"], rtol=1e-11, what="llh")
    assert_close(hb.grad_log_likelihood().cpu().numpy(), g["grad_llh"], rtol=1e-8, atol=1e-9,
                 what="grad llh")


@pytest.mark.parametrize("name", GP_CASES)
def test_loglik_and_gradient_with_inverse_from_the_factorisation(golden, name):
    """factorize(with_inverse=True): L^-1 from extra tasks of the persistent kernel (even n) or the
    level kernels (odd n), alpha = W^T W r, Sigma^-1 formed and contracted in one kernel -- against
    the reference fixtures, and against the path that solves with L."""
    g = golden(name)
    eng = _engine(g)
    hb = eng.hypers(g["thetas"]).factorize(with_inverse=True)
    assert hb.Winv is not None and np.all(hb.info == 0)
    assert_close(hb.log_likelihood().cpu().numpy(), g["llh"], rtol=1e-10, what="llh (alpha from W)")
    assert_close(hb.grad_log_likelihood().cpu().numpy(), g["grad_llh"], rtol=1e-8, atol=1e-9,
                 what="grad llh (fused)")
    hb2 = eng.hypers(g["thetas"]).factorize()
    assert_close(hb.alpha.cpu().numpy(), hb2.alpha.cpu().numpy(), rtol=1e-7,
                 atol=1e-9 * float(hb2.alpha.abs().max()), what="alpha from W vs solves")


@pytest.mark.parametrize("n", [1, 63, 64, 65, 130, 257, 300, 517])
def test_loglik_gradient_ragged_sizes(n):
    """Sizes that are not multiples of the 64-wide tiles (and of the doubling blocks of the
    triangular inverse) against the oracle's grad_log_likelihood."""
    from bayesian_quadrature_optimization_b200 import engine as E, _cabi
    rng = np.random.RandomState(n)
    pts = np.concatenate([rng.uniform(0, 1, (n, 2)), rng.randint(0, 3, (n, 1)).astype(float)], 1)
    y = np.sin(3 * pts[:, 0]) + pts[:, 2] + 0.1 * rng.normal(0, 1, n)
    thetas = np.array([[0.05, 0.1, 0.4, 0.6, 0.1, -0.3, 0.2, 0.0, -0.5, 0.3],
                       [0.02, -0.1, 0.5, 0.3, 0.0, 0.1, -0.2, 0.3, 0.2, -0.1]])
    eng = E.GPEngine(_cabi.KIND_PRODUCT_TASKS, 3, 3, False)
    eng.set_data(pts, y)
    hb = eng.hypers(thetas).factorize()
    assert np.all(hb.info == 0)
    spec = orc.KernelSpec(orc.PRODUCT_TASKS, 3, n_tasks=3, same_correlation=False)
    ogp = orc.OracleGP(spec, pts, y)
    want = np.stack([ogp.grad_log_likelihood(t[0], t[1], t[2:]) for t in thetas])
    assert_close(hb.grad_log_likelihood().cpu().numpy(), want, rtol=1e-8, atol=1e-9,
                 what="grad llh n=%d" % n)
    # the same with L^-1 produced by the factorisation launch
    hb = eng.hypers(thetas).factorize(with_inverse=True)
    assert_close(hb.grad_log_likelihood().cpu().numpy(), want, rtol=1e-8, atol=1e-9,
                 what="grad llh (inverse from the factorisation) n=%d" % n)
    want_llh = [ogp.log_likelihood(t[0], t[1], t[2:]) for t in thetas]
    assert_close(hb.log_likelihood().cpu().numpy(), want_llh, rtol=1e-10, what="llh n=%d" % n)


@pytest.mark.parametrize("name", GP_CASES)
def test_posterior_and_ei(golden, name):
    g = golden(name)
    eng = _engine(g)
    hb = eng.hypers(g["thetas"])
    post = hb.posterior(g["query"], var=True, grad=True)
    dm = eng.desc.dim_m
    assert_close(post["mean"].cpu().numpy(), g["post_mean"], rtol=1e-10, what="post mean")
    want_var = np.stack([np.diag(c) for c in g["post_cov"]])
    assert_close(post["var"].cpu().numpy(), want_var, rtol=1e-9, atol=1e-12, what="post var")
    assert_close(post["grad_mean"].cpu().numpy(), g["grad_post_mean"][:, :, :dm], rtol=1e-9,
                 atol=1e-13, what="grad mean")
    assert_close(post["grad_var"].cpu().numpy(), g["grad_post_cov"][:, :, :dm], rtol=1e-8,
                 atol=1e-12, what="grad var")
    best = np.full(len(g["thetas"]), np.max(g["evaluations"]))
    ei, gei = hb.ei(g["query"], best, grad=True)
    assert_close(ei.cpu().numpy(), g["ei"], rtol=1e-8, atol=1e-300, what="ei")
    assert_close(gei.cpu().numpy(), g["grad_ei"][:, :, :dm], rtol=1e-7, atol=1e-300, what="grad ei")


@pytest.mark.parametrize("name", CASES)
def test_stage_b_and_sample_ab(golden, name):
    g = golden(name)
    eng = _engine(g)
    hb = eng.hypers(g["thetas"])
    bq, lo, up = _bq(g, hb)
    ub = bq.setup(g["cands"])
    S, Cn, n = hb.S, ub.C, eng.n
    rows = 1 + eng.desc.dim_m
    G = ub.G.cpu().numpy().reshape(S, Cn, rows, n)
    R = ub.R.cpu().numpy().reshape(S, Cn, rows, n)
    assert_close(G[:, :, 0, :], g["gamma"], rtol=1e-13, what="gamma")
    s2 = R[:, :, 0, :]
    if bq.is_product:
        q = bq.qt.cpu().numpy()[:, g["points"][:, -1].astype(int)]  # [S][n]
        s2 = s2 / q[:, None, :]
    for s in range(S):  # Sigma^-1 gamma: two correct factorisations differ by cond(Sigma) * eps
        assert_close(s2[s], g["solve_2"][s], rtol=max(1e-8, cond_rtol(g, s)), atol=1e-12,
                     what="solve_2[%d]" % s)
    assert_close(ub.den.cpu().numpy(), g["den"], rtol=1e-9, what="den")
    # a, b, grad a, grad b at every (hyper, candidate, point)
    P = g["xq"].shape[0]
    pts, units = [], []
    for s in range(S):
        for c in range(Cn):
            for p in range(P):
                pts.append(g["xq"][p])
                units.append(c * S + s)
    out = bq.sample_ab(ub, np.array(pts), np.array(units, dtype=np.int32))
    got = out.cpu().numpy().reshape(S, Cn, P, -1)
    for s in range(S):
        assert_close(got[s], g["ab"][s], rtol=cond_rtol(g, s), atol=1e-11, what="ab[%d]" % s)
    # gradient_vector_b at the same points
    upts = np.zeros((Cn * S, P, int(g["dx"])))
    upts[:] = g["xq"][None, :, :]
    gvb, is_nan = bq.grad_vector_b(ub, upts)
    gvb = gvb.cpu().numpy().reshape(Cn, S, P, -1).transpose(1, 0, 2, 3)
    assert not is_nan.cpu().numpy().any()
    for s in range(S):
        assert_close(gvb[s], g["gvb"][s][:, :, :eng.desc.dim_m], rtol=10 * cond_rtol(g, s),
                     atol=1e-11, what="gvb[%d]" % s)


@pytest.mark.parametrize("name", ["product_uniform", "product_weighted"])
def test_discretised_kg(golden, name):
    from bayesian_quadrature_optimization_b200 import engine as E
    from scipy.stats import norm
    g = golden(name)
    eng = _engine(g)
    hb = eng.hypers(g["thetas"])
    bq, lo, up = _bq(g, hb)
    ub = bq.setup(g["cands"])
    S, Cn = hb.S, ub.C
    disc = g["discretization"]
    m = disc.shape[0]
    for s in range(S):
        for c in range(Cn):
            u = c * S + s
            ab = bq.sample_ab(ub, disc, np.full(m, u, dtype=np.int32)).cpu().numpy()
            kg, cnt, idx, c_out = E.kg_discretized(eng, ab[:, 0], ab[None, :, 1])
            assert_close(kg.cpu().numpy()[0], g["kg"][s, c], rtol=1e-8, atol=1e-13, what="kg")
            M = int(cnt.cpu().numpy()[0])
            if M <= 1:
                assert_close(np.zeros(eng.desc.dim_m), g["grad_kg"][s, c][:eng.desc.dim_m], atol=0)
                continue
            keep = idx.cpu().numpy()[0, :M]
            cc = np.abs(c_out.cpu().numpy()[0, :M - 1])
            upts = np.zeros((Cn * S, M, int(g["dx"])))
            upts[u] = disc[keep]
            gv, _ = bq.grad_vector_b(ub, upts)
            gv = gv.cpu().numpy()[u]
            grad = np.array([np.dot(np.diff(gv[:, l]), norm.pdf(cc)) for l in range(gv.shape[1])])
            assert_close(grad, g["grad_kg"][s, c][:eng.desc.dim_m], rtol=1e-7, atol=1e-12,
                         what="grad kg")


@pytest.mark.parametrize("name", CASES)
def test_inner_maximize_matches_cpu_restatement(golden, name):
    g = golden(name)
    eng = _engine(g)
    hb = eng.hypers(g["thetas"])
    bq, lo, up = _bq(g, hb)
    ub = bq.setup(g["cands"])
    spec, ogp, obq = oracle_from_golden(g)
    rng = np.random.RandomState(11)
    dx = int(g["dx"])
    zs = rng.normal(0, 1, 5)
    starts = rng.uniform(lo, up, (3, dx))
    out_max, out_arg, n_evals = bq.inner_maximize(ub, zs, starts, count_evals=True)
    out_max = out_max.cpu().numpy()
    out_arg = out_arg.cpu().numpy()
    w = g["wset"] if "wset" in g else None
    S = hb.S
    total = 0
    for c in range(ub.C):
        cand = g["cands"][c:c + 1, :]
        for s in (0, S - 1):
            th = g["thetas"][s]
            u = c * S + s

            def ab(x, wset):
                v = obq.compute_parameters_for_sample(x, cand, th[0], th[1], th[2:], w=w)
                gr = obq.compute_gradient_parameters_for_sample(x, cand, th[0], th[1], th[2:], w=w)
                return v["a"][0], np.asarray(v["b"]).reshape(-1)[0], gr["a"], gr["b"]

            for i, z in enumerate(zs):
                best, arg, ne = inner_opt.inner_maximize(ab, z, starts, lo, up)
                total += ne
                # the value is the acquisition ingredient: 1e-9 relative
                assert_close(out_max[u, i], best, rtol=1e-9, atol=1e-12, what="inner max")
                # the device argmax evaluated by the oracle reproduces the device value
                a, b, _, _ = ab(out_arg[u, i], 0)
                assert_close(a + z * b, out_max[u, i], rtol=1e-9, atol=1e-12, what="value at argmax")
                assert np.all(out_arg[u, i] >= lo - 1e-15) and np.all(out_arg[u, i] <= up + 1e-15)


@pytest.mark.parametrize("name", CASES)
def test_full_candidate_evaluation_against_oracle_pipeline(golden, name):
    """evaluate_mc_bayesian_candidate_points_no_restarts: device pipeline vs the same
    ingredients assembled on the CPU (inner maxima from the CPU restatement of the optimiser,
    gradient_vector_b and the means from the oracle)."""
    g = golden(name)
    eng = _engine(g)
    hb = eng.hypers(g["thetas"])
    bq, lo, up = _bq(g, hb)
    spec, ogp, obq = oracle_from_golden(g)
    rng = np.random.RandomState(3)
    dx = int(g["dx"])
    zs = rng.normal(0, 1, 4)
    starts = rng.uniform(lo, up, (2, dx))
    max_mean = rng.normal(0, 1, hb.S)
    res = bq.evaluate_candidates(g["cands"], zs, starts, max_mean=max_mean)
    w = g["wset"] if "wset" in g else None
    vals = np.zeros(len(g["cands"]))
    grads = np.zeros((len(g["cands"]), eng.desc.dim_m))
    for c in range(len(g["cands"])):
        cand = g["cands"][c:c + 1, :]
        vk, gk = [], []
        for s, th in enumerate(g["thetas"]):
            def ab(x, wset):
                v = obq.compute_parameters_for_sample(x, cand, th[0], th[1], th[2:], w=w)
                gr = obq.compute_gradient_parameters_for_sample(x, cand, th[0], th[1], th[2:], w=w)
                return v["a"][0], np.asarray(v["b"]).reshape(-1)[0], gr["a"], gr["b"]
            mx, args = [], []
            for z in zs:
                best, arg, _ = inner_opt.inner_maximize(ab, z, starts, lo, up)
                mx.append(best)
                args.append(arg)
            vk.append(np.mean(mx) - max_mean[s])
            gb = obq.gradient_vector_b(cand, np.array(args), th[0], th[1], th[2:], w=w)
            gk.append(np.mean(gb * zs.reshape(-1, 1), axis=0)[:eng.desc.dim_m])
        vals[c] = np.mean(vk)
        grads[c] = np.mean(gk, axis=0)
    assert_close(res["evaluations"].cpu().numpy(), vals, rtol=1e-9, atol=1e-12, what="acq value")
    assert_close(res["gradient"].cpu().numpy(), grads, rtol=1e-6, atol=1e-9, what="acq gradient")
    # argmax candidate index is exact
    assert int(np.argmax(res["evaluations"].cpu().numpy())) == int(np.argmax(vals))


def test_cholesky_solve_random_sizes():
    """Non-multiple-of-64 sizes, several right-hand sides, against LAPACK."""
    import torch
    from bayesian_quadrature_optimization_b200 import engine as E, _cabi
    import ctypes as C
    rng = np.random.RandomState(0)
    eng = E.GPEngine(_cabi.KIND_MATERN52, 1)
    # m * S <= 512 takes the vector-solve kernels, above that the DMMA tile solve: both at ragged
    # sizes, odd n (8-byte cp.async staging) and even n (16-byte staging)
    # even n: the persistent kernel (one launch); odd n: the multi-launch path
    for n, S, m in [(1, 2, 1), (63, 2, 3), (64, 1, 65), (65, 3, 7), (300, 2, 130), (517, 2, 5),
                    (65, 3, 200), (300, 2, 300), (517, 2, 261), (130, 9, 64), (257, 1, 513),
                    (64, 2, 300), (1, 3, 200), (2, 1, 1), (66, 5, 2), (128, 3, 4), (190, 1, 1),
                    (448, 7, 3), (1024, 2, 2), (1090, 3, 1)]:
        A = rng.normal(0, 1, (S, n, n + 5))
        A = A @ A.transpose(0, 2, 1) + 1e-3 * np.eye(n)
        Ad = eng.to_device(A.copy())
        info = E.potrf_batched(eng, Ad)
        assert not info.cpu().numpy().any()
        L = np.stack([np.linalg.cholesky(a) for a in A])
        assert_close(Ad.cpu().numpy(), L, rtol=1e-9, atol=1e-11, what="potrf n=%d" % n)
        B = rng.normal(0, 1, (S, m, n))
        Bd = eng.to_device(B.copy())
        _cabi.check(eng.lib.bqo_potrs_batched(_cabi.ptr(Ad), n, S, _cabi.ptr(Bd), m,
                                              torch.cuda.current_stream().cuda_stream), "potrs")
        X = np.stack([np.linalg.solve(A[s], B[s].T).T for s in range(S)])
        assert_close(Bd.cpu().numpy(), X, rtol=1e-7, atol=1e-9, what="potrs n=%d" % n)


def test_cholesky_not_positive_definite_and_jitter():
    """sbo/lib/la_functions.py:19-38: diag <= 0 -> error; otherwise jitter escalation."""
    import torch
    from bayesian_quadrature_optimization_b200 import engine as E, _cabi
    # duplicate training points and zero noise make Sigma singular: jitter must rescue it
    pts = np.array([[0.1], [0.1], [0.5], [0.9], [0.9]])
    y = np.array([1.0, 1.0, 2.0, 3.0, 3.0])
    eng = E.GPEngine(_cabi.KIND_MATERN52, 1)
    eng.set_data(pts, y)
    thetas = np.array([[0.0, 0.0, 0.3], [1e-3, 0.0, 0.3]])
    hb = eng.hypers(thetas).factorize()
    assert hb.info[1] == 0 and hb.jitter[1] == 0.0
    assert hb.info[0] == 0 and hb.jitter[0] > 0.0
    spec = orc.KernelSpec(orc.MATERN52, 1)
    cov = spec.cov(thetas[0][2:], pts)
    want = orc.cholesky(cov, max_tries=7)
    # same jitter level as the reference picked (mean(diag) * 1e-6 * 10^t)
    t = np.log10(hb.jitter[0] / (np.mean(np.diag(cov)) * 1e-6))
    assert abs(t - round(t)) < 1e-9
    ref_j = None
    for tt in range(8):
        j = np.mean(np.diag(cov)) * 1e-6 * 10 ** tt
        try:
            np.linalg.cholesky(cov + j * np.eye(5))
            ref_j = j
            break
        except np.linalg.LinAlgError:
            pass
    assert ref_j is not None
    if abs(ref_j - hb.jitter[0]) <= 1e-12 * ref_j:
        assert_close(hb.L.cpu().numpy()[0], want, rtol=1e-6, atol=1e-9, what="jittered chol")
    # a matrix with a non-positive diagonal entry reports the reference's first error
    A = eng.to_device(np.array([[[1.0, 2.0], [2.0, -1.0]]]))
    assert E.potrf_batched(eng, A).cpu().numpy()[0] == 2


def test_persistent_cholesky_against_multi_launch_and_failures():
    """The persistent kernel against the multi-launch factorisation (same matrices), dpotrf's
    failure code in the middle of a large matrix, and a failing matrix next to healthy ones in
    the same batch (its tasks stop, the others finish)."""
    import torch
    from bayesian_quadrature_optimization_b200 import engine as E, _cabi
    rng = np.random.RandomState(3)
    eng = E.GPEngine(_cabi.KIND_MATERN52, 1)
    for n, S in [(640, 4), (2048, 3), (130, 20)]:
        A = rng.normal(0, 1, (S, n, n + 8))
        A = A @ A.transpose(0, 2, 1) + 1e-2 * np.eye(n)
        a1, a2 = eng.to_device(A.copy()), eng.to_device(A.copy())
        assert not E.potrf_batched(eng, a1).cpu().numpy().any()
        old = eng.lib.bqo_potrf_set_multi_launch(1)
        try:
            assert not E.potrf_batched(eng, a2).cpu().numpy().any()
        finally:
            eng.lib.bqo_potrf_set_multi_launch(old)
        ref = torch.linalg.cholesky(eng.to_device(A))
        scale = ref.abs().max().item()
        assert (a1 - ref).abs().max().item() < 1e-11 * scale
        assert (a2 - ref).abs().max().item() < 1e-11 * scale
        assert a1.triu(1).abs().max().item() == 0.0
    # leading minor 301 of matrix 1 is not positive definite; matrices 0 and 2 are fine
    n, S = 512, 3
    A = rng.normal(0, 1, (S, n, n + 8))
    A = A @ A.transpose(0, 2, 1) + 1e-2 * np.eye(n)
    Lref = np.linalg.cholesky(A[1])
    # make the pivot at index 300 negative: subtract more than its Schur complement
    A[1, 300, 300] -= 2.0 * Lref[300, 300] ** 2
    Ad = eng.to_device(A.copy())
    info = E.potrf_batched(eng, Ad).cpu().numpy()
    assert info[0] == 0 and info[2] == 0 and info[1] == 301
    for s in (0, 2):
        assert_close(Ad[s].cpu().numpy(), np.linalg.cholesky(A[s]), rtol=1e-9, atol=1e-10,
                     what="healthy matrix next to a failing one")


def test_large_size_properties():
    """n = 2048 (BASELINE size): L L^T reproduces Sigma, potrs residual, log-likelihood vs LAPACK."""
    from bayesian_quadrature_optimization_b200 import engine as E, _cabi
    rng = np.random.RandomState(0)
    n, d = 2048, 6
    pts = np.concatenate([rng.uniform(0, 1, (n, 4)), rng.gamma(2.0, 1.0, (n, 2))], 1)
    y = np.sin(pts.sum(1)) + 0.01 * rng.normal(0, 1, n)
    thetas = np.array([[1e-4, 0.0, 0.3, 0.3, 0.3, 0.3, 2.0, 2.0, 1.0],
                       [2e-4, 0.05, 0.35, 0.25, 0.3, 0.4, 1.8, 2.2, 1.2]])
    eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, d)
    eng.set_data(pts, y)
    hb = eng.hypers(thetas)
    Sigma = hb.cov(add_noise=True)
    hb.factorize()
    assert np.all(hb.info == 0)
    L = hb.L
    rec = L @ L.transpose(1, 2)
    err = ((rec - Sigma).norm() / Sigma.norm()).item()
    assert err < 1e-13, err
    spec = orc.KernelSpec(orc.SCALED_MATERN52, d)
    ogp = orc.OracleGP(spec, pts, y)
    llh = hb.log_likelihood().cpu().numpy()
    want = [ogp.log_likelihood(t[0], t[1], t[2:]) for t in thetas]
    assert_close(llh, want, rtol=1e-9, what="llh n=2048")
    res = (Sigma @ hb.alpha.unsqueeze(2)).squeeze(2) - (eng.y.unsqueeze(0) - hb.theta[:, 1:2])
    assert (res.norm() / eng.y.norm()).item() < 1e-9
    # many right-hand sides (DMMA tile solve, the stage-B shape of the benchmark)
    import torch
    B = torch.randn(2, 600, n, dtype=torch.float64, device=eng.device,
                    generator=torch.Generator(device=eng.device).manual_seed(1))
    X = hb.solve(B.clone())
    resid = (X @ Sigma.transpose(1, 2)) - B  # rows are right-hand sides: X Sigma^T = B
    assert (resid.norm() / B.norm()).item() < 1e-9


def test_adam_step_matches_reference_sgd():
    import torch
    from bayesian_quadrature_optimization_b200 import engine as E, _cabi
    eng = E.GPEngine(_cabi.KIND_MATERN52, 1)
    rng = np.random.RandomState(2)
    R, dim = 5, 3
    target = rng.uniform(0.2, 0.8, (R, dim))
    start = rng.uniform(0, 1, (R, dim))
    bounds = [(0.0, 1.0)] * dim
    want = np.stack([orc.sgd_adam(start[r], lambda x, r=r: 2 * (x - target[r]), 1, bounds=bounds,
                                  maxepoch=25) for r in range(R)])
    pt = eng.to_device(start.copy())
    m0 = torch.zeros_like(pt)
    v0 = torch.zeros_like(pt)
    active = torch.ones(R, dtype=torch.int32, device=eng.device)
    lo = eng.to_device(np.zeros(dim))
    up = eng.to_device(np.ones(dim))
    tgt = eng.to_device(target)
    for t in range(1, 26):
        grad = (2 * (pt - tgt)).contiguous()
        _cabi.check(eng.lib.bqo_adam_step_batched(
            _cabi.ptr(pt), _cabi.ptr(grad), _cabi.ptr(m0), _cabi.ptr(v0), _cabi.ptr(active),
            _cabi.ptr(lo), _cabi.ptr(up), R, dim, t, 0.1, 0.9, 0.999, 1e-8,
            torch.cuda.current_stream().cuda_stream), "adam")
    assert_close(pt.cpu().numpy(), want, rtol=1e-12, atol=1e-14, what="adam")


@pytest.mark.parametrize("n", [1, 5, 63, 64, 65, 200, 300])
def test_factor_append_matches_refactorisation(n):
    """extending S factors by one point in O(n^2) gives the factor, alpha and
    log-likelihood of a fresh O(n^3) factorisation of the n + 1 points."""
    from bayesian_quadrature_optimization_b200 import engine as E, _cabi
    rng = np.random.RandomState(100 + n)
    pts = np.concatenate([rng.uniform(0, 1, (n + 1, 2)), rng.randint(0, 3, (n + 1, 1)).astype(float)], 1)
    y = np.sin(3 * pts[:, 0]) + pts[:, 2] + 0.1 * rng.normal(0, 1, n + 1)
    vno = 0.01 + 0.01 * rng.uniform(0, 1, n + 1)
    thetas = np.array([[0.05, 0.1, 0.4, 0.6, 0.3, -0.2],
                       [0.02, -0.1, 0.5, 0.3, 0.1, -0.4]])
    for noise_obs in (None, vno):
        eng = E.GPEngine(_cabi.KIND_PRODUCT_TASKS, 3, 3, True)
        eng.set_data(pts[:n], y[:n], None if noise_obs is None else noise_obs[:n])
        hb = eng.hypers(thetas).factorize()
        eng2 = E.GPEngine(_cabi.KIND_PRODUCT_TASKS, 3, 3, True)
        eng2.set_data(pts, y, noise_obs)
        got = hb.appended(eng2)
        want = eng2.hypers(thetas).factorize()
        assert np.all(got.info == 0)
        assert_close(got.L.cpu().numpy(), want.L.cpu().numpy(), rtol=1e-10, atol=1e-13,
                     what="appended factor n=%d" % n)
        assert_close(got.alpha.cpu().numpy(), want.alpha.cpu().numpy(), rtol=1e-8, atol=1e-10,
                     what="alpha after append")
        assert_close(got.log_likelihood().cpu().numpy(), want.log_likelihood().cpu().numpy(),
                     rtol=1e-11, what="llh after append")
        spec = orc.KernelSpec(orc.PRODUCT_TASKS, 3, n_tasks=3, same_correlation=True)
        ogp = orc.OracleGP(spec, pts, y, noise_obs)
        ref = [ogp.log_likelihood(t[0], t[1], t[2:]) for t in thetas]
        assert_close(got.log_likelihood().cpu().numpy(), ref, rtol=1e-10, what="llh vs oracle")


def test_hot_path_is_bit_reproducible_and_independent_of_the_z_chunking(golden):
    """Runs are handed to warps by an atomic counter, so the schedule differs from launch to
    launch; the results must not: two passes give identical bits, and so do different numbers of
    Z samples per CTA and the W-distance table against the on-the-fly distances of the argmax."""
    import torch
    from bayesian_quadrature_optimization_b200 import engine as E
    g = golden("scaled_gamma")
    eng = _engine(g)
    hb = eng.hypers(g["thetas"])
    bq, lo, up = _bq(g, hb)
    rng = np.random.RandomState(21)
    zs = rng.normal(0, 1, 12)
    starts = rng.uniform(lo, up, (4, int(g["dx"])))
    a = bq.evaluate_candidates(g["cands"], zs, starts)
    b = bq.evaluate_candidates(g["cands"], zs, starts)
    for key in ("evaluations", "gradient", "max", "optimum"):
        assert torch.equal(a[key], b[key]), key
    for zc in (1, 5, 12):
        c = bq.evaluate_candidates(g["cands"], zs, starts, z_per_cta=zc)
        assert torch.equal(a["max"], c["max"]) and torch.equal(a["optimum"], c["optimum"]), zc
        assert torch.equal(a["evaluations"], c["evaluations"])


def test_inner_maximize_other_w_sample_counts(golden):
    """M != 10 W samples (the rolled W loop without the distance table) and several W sets: the
    device maxima against the CPU restatement of the optimiser on the oracle's a, b."""
    from bayesian_quadrature_optimization_b200 import engine as E
    g = golden("scaled_gamma")
    eng = _engine(g)
    hb = eng.hypers(g["thetas"])
    dx = int(g["dx"])
    lo, up = np.zeros(dx), np.ones(dx)
    wsets = np.random.RandomState(31).gamma(2.0, 1.0, (2, 7, g["points"].shape[1] - dx))
    bq = E.BQEngine(hb, dx, wsets=wsets, lower=lo, upper=up)
    ub = bq.setup(g["cands"])
    spec, ogp, obq = oracle_from_golden(g)
    rng = np.random.RandomState(13)
    zs = rng.normal(0, 1, 3)
    starts = rng.uniform(lo, up, (2, dx))
    out_max, out_arg, _ = bq.inner_maximize(ub, zs, starts, count_evals=True)
    out_max, out_arg = out_max.cpu().numpy(), out_arg.cpu().numpy()
    S = hb.S
    for c in range(ub.C):
        cand = g["cands"][c:c + 1, :]
        w = wsets[c % 2]  # the unit's W set is candidate index % n_wsets
        th = g["thetas"][0]
        u = c * S

        def ab(x, wset):
            v = obq.compute_parameters_for_sample(x, cand, th[0], th[1], th[2:], w=w)
            gr = obq.compute_gradient_parameters_for_sample(x, cand, th[0], th[1], th[2:], w=w)
            return v["a"][0], np.asarray(v["b"]).reshape(-1)[0], gr["a"], gr["b"]

        for i, z in enumerate(zs):
            best, arg, _ = inner_opt.inner_maximize(ab, z, starts, lo, up)
            assert_close(out_max[u, i], best, rtol=1e-9, atol=1e-12, what="inner max, M = 7")
            a, b, _, _ = ab(out_arg[u, i], 0)
            assert_close(a + z * b, out_max[u, i], rtol=1e-9, atol=1e-12, what="value at argmax")


@pytest.mark.parametrize("dx,dw", [(3, 1), (5, 2), (1, 2), (6, 1), (6, 0), (3, 0)])
def test_stage_c_other_dimensions(dx, dw):
    """Template instantiations the fixtures do not reach (x dimensions 1..6, w dimensions 0..2):
    a, b and their gradients against the oracle, the maximiser's value reproduced at its argument."""
    from bayesian_quadrature_optimization_b200 import engine as E, _cabi
    rng = np.random.RandomState(10 * dx + dw)
    n = 70
    x = rng.uniform(0, 1, (n, dx))
    lo, up = np.zeros(dx), np.ones(dx)
    if dw > 0:
        w = rng.gamma(2.0, 1.0, (n, dw))
        pts = np.concatenate([x, w], 1)
        y = np.sin(pts.sum(1)) + 0.05 * rng.normal(0, 1, n)
        thetas = np.array([[1e-3, 0.1] + dx * [0.6] + dw * [2.0] + [1.2],
                           [2e-3, 0.0] + dx * [0.9] + dw * [1.5] + [0.8]])
        eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, dx + dw)
        eng.set_data(pts, y)
        wset = rng.gamma(2.0, 1.0, (10, dw))
        bq = E.BQEngine(eng.hypers(thetas), dx, wsets=wset, lower=lo, upper=up)
        spec = orc.KernelSpec(orc.SCALED_MATERN52, dx + dw)
        ogp = orc.OracleGP(spec, pts, y)
        obq = orc.OracleBQ(ogp, list(range(dx)), orc.GAMMA)
        cands = np.concatenate([rng.uniform(0, 1, (2, dx)), rng.gamma(2.0, 1.0, (2, dw))], 1)
    else:
        T = 3
        task = rng.randint(0, T, n).astype(float)
        pts = np.concatenate([x, task[:, None]], 1)
        y = np.sin(x.sum(1)) + 0.3 * task + 0.05 * rng.normal(0, 1, n)
        thetas = np.array([[1e-3, 0.1] + dx * [0.6] + [0.2, -0.1, 0.3, 0.0, 0.1, -0.2],
                           [2e-3, 0.0] + dx * [0.9] + [0.0, 0.1, -0.2, 0.3, -0.1, 0.2]])
        eng = E.GPEngine(_cabi.KIND_PRODUCT_TASKS, dx + 1, T, False)
        eng.set_data(pts, y)
        wset = None
        bq = E.BQEngine(eng.hypers(thetas), dx, lower=lo, upper=up)
        spec = orc.KernelSpec(orc.PRODUCT_TASKS, dx + 1, n_tasks=T, same_correlation=False)
        ogp = orc.OracleGP(spec, pts, y)
        obq = orc.OracleBQ(ogp, list(range(dx)), orc.UNIFORM_FINITE)
        cands = np.concatenate([rng.uniform(0, 1, (2, dx)), np.array([[0.0], [2.0]])], 1)
    ub = bq.setup(cands)
    S = 2
    xq = rng.uniform(0, 1, (3, dx))
    zs = np.array([-0.8, 1.1])
    starts = rng.uniform(0, 1, (2, dx))
    out_max, out_arg, _ = bq.inner_maximize(ub, zs, starts, count_evals=True)
    out_max, out_arg = out_max.cpu().numpy(), out_arg.cpu().numpy()
    for s in range(S):
        th = thetas[s]
        for c in range(2):
            u = c * S + s
            got = bq.sample_ab(ub, xq, np.full(len(xq), u, dtype=np.int32)).cpu().numpy()
            for p in range(len(xq)):
                v = obq.compute_parameters_for_sample(xq[p], cands[c:c + 1], th[0], th[1], th[2:], w=wset)
                gr = obq.compute_gradient_parameters_for_sample(xq[p], cands[c:c + 1], th[0], th[1],
                                                                th[2:], w=wset)
                want = np.concatenate([[v["a"][0], np.asarray(v["b"]).reshape(-1)[0]],
                                       np.asarray(gr["a"]).reshape(-1), np.asarray(gr["b"]).reshape(-1)])
                assert_close(got[p], want, rtol=1e-8, atol=1e-10, what="ab dx=%d dw=%d" % (dx, dw))
            for i, z in enumerate(zs):
                v = obq.compute_parameters_for_sample(out_arg[u, i], cands[c:c + 1], th[0], th[1],
                                                      th[2:], w=wset)
                assert_close(v["a"][0] + z * np.asarray(v["b"]).reshape(-1)[0], out_max[u, i],
                             rtol=1e-8, atol=1e-10, what="value at argmax")
                assert np.all(out_arg[u, i] >= 0) and np.all(out_arg[u, i] <= 1)
    gvb, is_nan = bq.grad_vector_b(ub, np.repeat(xq[None], 2 * S, axis=0))
    assert not is_nan.cpu().numpy().any()
    th = thetas[1]
    gb = obq.gradient_vector_b(cands[1:2], xq, th[0], th[1], th[2:], w=wset)
    assert_close(gvb.cpu().numpy()[1 * S + 1][:, :gb.shape[1]], gb[:, :gvb.shape[2]], rtol=1e-6,
                 atol=1e-9, what="gvb dx=%d dw=%d" % (dx, dw))


def test_no_kernel_writes_outside_its_buffers():
    """compute-sanitizer is closed on the B200 pool, so out-of-bounds WRITES are looked for with
    guard bands: every FP64 buffer the engine hands to the library sits between two bands of a
    known pattern (GPEngine.guard_doubles); one pass through every kernel family at ragged sizes
    (n = 200, 131, 64, 1; gamma-W and finite-W models) must leave all bands intact."""
    import runpy
    import os
    from bayesian_quadrature_optimization_b200 import engine as E
    from tests.conftest import ROOT
    E.GPEngine.guard_doubles = 512
    E.GPEngine._guards = []
    try:
        runpy.run_path(os.path.join(ROOT, "tests", "microbench", "sanitizer_smoke.py"), run_name="smoke")
        n_buffers = len(E.GPEngine._guards)
        bad = E.GPEngine.check_guards()
    finally:
        E.GPEngine.guard_doubles = 0
        E.GPEngine._guards = []
    assert n_buffers > 100
    assert bad == [], "out-of-bounds writes next to buffers of shape %s" % (bad,)
