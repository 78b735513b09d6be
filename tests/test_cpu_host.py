"""CPU tests: C ABI symbol table, host-side logic, the CPU restatement of the device optimiser,
and the world_size-2 gloo path.  No compute call is made into the CUDA library here."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "bqo_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bqo_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bayesian_quadrature_optimization_b200 import _build, _cabi
    path = _build.build()
    lib = ctypes.CDLL(str(path))
    declared = _declared_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name
    # the ctypes table binds exactly the declared entry points
    assert sorted(_cabi.SIGNATURES) == declared
    lib2 = _cabi.load()
    assert lib2.bqo_abi_version() == 1


def test_struct_layouts_match_header():
    from bayesian_quadrature_optimization_b200 import _cabi
    assert ctypes.sizeof(_cabi.KernelDesc) == 24
    assert ctypes.sizeof(_cabi.StageC) == 24 + 8 * 4 + 15 * 8
    assert ctypes.sizeof(_cabi.InnerOpt) == 48
    d = _cabi.KernelDesc(2, 3, 2, 4, 1, 4)
    assert _cabi.load().bqo_hyper_stride(ctypes.byref(d)) == 4 + 2 * 16 + 2 * 16


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from bayesian_quadrature_optimization_b200 import _cabi, engine
    with pytest.raises(_cabi.BqoLibraryError):
        engine.GPEngine(_cabi.KIND_MATERN52, 1)
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    # building the model is host-only work (defaults, priors, samplers); every number comes from
    # the device, so the first numerical call fails loudly
    gp = GPFittingGaussian(["Matern52"], {"points": [[0.0], [1.0]], "evaluations": [0.0, 1.0],
                                          "var_noise": []}, [1], bounds_domain=[[0, 1]])
    for call in (lambda: gp.log_likelihood(0.0, 0.0, np.array([1.0])),
                 lambda: gp.log_prob_parameters(np.array([0.0, 0.0, 1.0])),
                 lambda: gp.sample_parameters(1),
                 lambda: gp.compute_posterior_parameters(np.array([[0.5]])),
                 lambda: gp.evaluate_cov(np.array([[0.5]]), np.array([1.0]))):
        with pytest.raises(_cabi.BqoLibraryError):
            call()


def test_product_has_no_oracle_import():
    pkg = os.path.join(ROOT, "bayesian_quadrature_optimization_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src, "product code must not touch oracle/: " + f


def test_inner_optimiser_restatement_finds_lbfgsb_quality_maxima():
    """The projected-BFGS iteration (CPU restatement of the CUDA kernel) reaches maxima at least
    as good as scipy's L-BFGS-B under the reference's settings (maxiter=10, factr=1e12)."""
    from scipy.optimize import fmin_l_bfgs_b
    from oracle import inner_opt
    rng = np.random.RandomState(0)
    d = 3
    lo, up = np.zeros(d), np.ones(d)
    worse = 0
    trials = 40
    for _ in range(trials):
        c = rng.uniform(-0.2, 1.2, (6, d))
        w = rng.normal(0, 1, 6)
        s = rng.uniform(0.2, 0.6, 6)

        def F(x):
            r2 = ((x[None, :] - c) ** 2).sum(1) / s ** 2
            return float((w * np.exp(-r2)).sum())

        def G(x):
            r2 = ((x[None, :] - c) ** 2).sum(1) / s ** 2
            return ((w * np.exp(-r2) * (-2.0 / s ** 2))[:, None] * (x[None, :] - c)).sum(0)

        x0 = rng.uniform(0, 1, d)
        Fm, xm, ne = inner_opt.pbfgs_maximize(lambda x, e: (F(x), G(x)), x0, lo, up)
        opt = fmin_l_bfgs_b(lambda x: -F(x), x0, fprime=lambda x: -G(x),
                            bounds=list(zip(lo, up)), factr=1e12, maxiter=10)
        assert abs(F(xm) - Fm) < 1e-12
        assert np.all(xm >= lo) and np.all(xm <= up)
        assert Fm >= F(np.clip(x0, lo, up)) - 1e-15
        if Fm < -opt[1] - 1e-3 * max(1.0, abs(opt[1])):
            worse += 1
    assert worse <= trials // 5, worse


def test_shard_range_partitions():
    from bayesian_quadrature_optimization_b200.distributed import shard_range
    for n in (1, 7, 64, 100):
        for ws in (1, 2, 3, 8):
            got = []
            for r in range(ws):
                lo, hi = shard_range(n, r, ws)
                got += list(range(lo, hi))
            assert got == list(range(n))


def _gloo_worker(rank, world_size, port, out):
    import torch
    import torch.distributed as dist
    from bayesian_quadrature_optimization_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    rng = np.random.RandomState(0)
    values = rng.normal(0, 1, 11)
    values[3] = values[9] = values.max() + 1.0  # a tie across the two shards
    pts = rng.uniform(0, 1, (11, 2))
    lo, hi = D.shard_range(11, rank, world_size)
    loc = int(np.argmax(values[lo:hi]))
    v, gi, p = D.allgather_best(torch.tensor(values[lo + loc]), torch.tensor(lo + loc),
                                torch.tensor(pts[lo + loc]))
    part = torch.tensor(values[lo:hi].sum()).reshape(1)
    tot = D.allgather_ordered_sum(part)

    # hyper-sample partition: a stand-in engine whose per-hyper-sample scores are known
    thetas = np.arange(5 * 3, dtype=float).reshape(5, 3)

    class FakeBQ(object):
        def __init__(self, th):
            self.th = th

        def evaluate_candidates(self, cands, z, starts, max_mean=None, **kw):
            per = torch.tensor(self.th.sum(1))[:, None] * torch.tensor(cands.sum(1))[None, :]
            return {"evaluations": per.mean(0), "gradient": per.mean(0)[:, None] * torch.ones(1, 2)}

    cands = rng.uniform(0, 1, (4, 2))
    res = D.evaluate_candidates_hyper_sharded(lambda th: FakeBQ(th), thetas, cands, None, None)
    out.put((rank, v, gi, p.tolist(), float(tot[0]), res["evaluations"].tolist(),
             res["gradient"].tolist(), res["argmax"]))
    dist.destroy_process_group()


def test_world_size_two_gloo_selection():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    rng = np.random.RandomState(0)
    values = rng.normal(0, 1, 11)
    values[3] = values[9] = values.max() + 1.0
    pts = rng.uniform(0, 1, (11, 2))
    cands = rng.uniform(0, 1, (4, 2))
    thetas = np.arange(5 * 3, dtype=float).reshape(5, 3)
    want = (thetas.sum(1)[:, None] * cands.sum(1)[None, :]).mean(0)
    for rank, v, gi, p, tot, ev, gr, am in res:
        assert gi == 3 and v == values[3]          # np.argmax tie rule: first index
        lo0, hi0 = 0, 6
        assert tot == values[:6].sum() + values[6:].sum()
        # hyper-samples 3 + 2 over the two ranks: same totals, bit-identical on both ranks
        assert np.allclose(ev, want, rtol=1e-14) and am == int(np.argmax(want))
        assert np.allclose(np.array(gr)[:, 0], want, rtol=1e-14)
    assert res[0][5] == res[1][5] and res[0][6] == res[1][6]


def test_host_sgd_matches_oracle_and_handles_unavailable_gradients():
    """lib/stochastic_gradient_descent.SGD (a36): Adam path equals the oracle's restatement bit for
    bit; the object np.nan triggers the random perturbation; momentum SGD converges too."""
    from bayesian_quadrature_optimization_b200.lib.stochastic_gradient_descent import SGD
    from oracle import bqo_oracle as orc
    target = np.array([0.3, 0.8, 0.55])
    grad = lambda x: 2.0 * (x - target)
    bounds = [(0.0, 1.0)] * 3
    start = np.array([0.9, 0.1, 0.2])
    got = SGD(start, grad, 2, bounds=bounds, maxepoch=40)
    want = orc.sgd_adam(start, grad, 2, bounds=bounds, maxepoch=40)
    assert np.array_equal(got, want)
    assert np.array_equal(start, np.array([0.9, 0.1, 0.2]))  # inputs are never mutated
    calls = {"n": 0}

    def sometimes_unavailable(x, scale):
        calls["n"] += 1
        if calls["n"] in (1, 2, 7):
            return np.nan
        return scale * 2.0 * (x - target)

    np.random.seed(0)
    out = SGD(start, sometimes_unavailable, 1, args=(1.0,), bounds=bounds, maxepoch=200)
    assert np.all(out >= 0.0) and np.all(out <= 1.0) and np.linalg.norm(out - target) < 0.15
    out2 = SGD(start, grad, 1, bounds=bounds, adam=False, learning_rate=0.05, momentum=0.5,
               maxepoch=300)
    assert np.linalg.norm(out2 - target) < 0.05
    # projection: an optimum outside the box ends on the boundary
    out3 = SGD(start, lambda x: 2.0 * (x - np.array([2.0, -1.0, 0.5])), 1, bounds=bounds,
               maxepoch=250)
    assert out3[0] == 1.0 and out3[1] == 0.0


def test_bayesian_evaluations_mean_and_std():
    from bayesian_quadrature_optimization_b200.bayesian.bayesian_evaluations import \
        BayesianEvaluations

    class Model(object):
        samples_parameters = [np.array([0.1, 1.0, 2.0, 3.0]), np.array([0.2, 2.0, 4.0, 5.0]),
                              np.array([0.3, 3.0, 6.0, 7.0])]

    def f(point, extra, var_noise, mean, params):
        return float(point[0] * extra + var_noise + mean + params.sum())

    vals = [2.0 * 10 + 0.2 + 2.0 + 9.0, 2.0 * 10 + 0.3 + 3.0 + 13.0]
    m, s = BayesianEvaluations.evaluate(f, np.array([2.0]), Model(), 2, None, 10)
    assert m == np.mean(vals) and s == np.std(vals)
    g = lambda point, var_noise, mean, params: point * mean
    m2, s2 = BayesianEvaluations.evaluate(g, np.array([1.0, 2.0]), Model(), 3)
    assert np.allclose(m2, np.array([2.0, 4.0])) and m2.shape == (2,) and s2.shape == (2,)


def test_bench_accounting_helpers():
    """bench.py's flop count is SURVEY 8d's formula, and the issue-port model is plain arithmetic on
    the committed instruction counts (no GPU needed): a launch that takes exactly the modelled
    cycles reports a ratio of 1."""
    import bench
    d = bench.DX + bench.DW
    want = (bench.M_W * bench.N_TRAIN * (2 * d + 10) + bench.M_W * bench.N_TRAIN * (3 * bench.DX + 8)
            + 4 * bench.N_TRAIN * (1 + bench.DX))
    assert bench.f_eval_flops() == want == 901120
    cfg = bench.workload_config()
    assert cfg["units_per_step"] == cfg["candidates_per_step"] * cfg["hyper_samples"]
    evals = 1.0e7
    m = bench.issue_model(evals, 100.0, None, 34.6)
    assert "model_cycles_over_elapsed_cycles" not in m  # no clock sample, no cycle ratio
    cyc = m["issue_cycles_per_point"]
    assert cyc == 2.0 * m["fp64_instr_per_point"] + m["fp64_3src_instr_per_point"] + m["other_instr_per_point"]
    mhz = 1965.0
    ms = evals / 32.0 * bench.N_TRAIN * cyc / (148 * 4) / (mhz * 1e6) * 1e3
    m = bench.issue_model(evals, ms, {"sm_mhz": mhz}, 34.6)
    assert abs(m["model_cycles_over_elapsed_cycles"] - 1.0) < 1e-12
    assert 0.0 < m["executed_frac_of_dfma_issue_rate"] < 1.0


def _pp_decode(tsk, S, nb, with_w):
    """Python restatement of pp_decode (csrc/bqo_potrf_persistent.cuh): task index ->
    (kind, matrix, tile row, tile column); kind 0 = factor tile, 1 = inverse tile."""
    if tsk < 3 * S:
        w = tsk // S
        return (0, tsk % S, 0 if w == 0 else 1, 1 if w == 2 else 0)
    rem = tsk - 3 * S
    for jj in range(0, nb - 1):
        chain = jj + 2 <= nb - 1
        c1a = S if chain else 0
        if rem < c1a:
            return (0, rem, jj + 2, jj)
        rem -= c1a
        c2 = 2 * S if chain else 0
        if rem < c2:
            return (0, rem % S, jj + 2, jj + 1 if rem < S else jj + 2)
        rem -= c2
        c1b = S * (nb - jj - 3) if chain else 0
        if rem < c1b:
            return (0, rem % S, jj + 3 + rem // S, jj)
        rem -= c1b
        c3 = S * jj if with_w else 0
        if rem < c3:
            return (1, rem % S, jj, rem // S)
        rem -= c3
    return (1, rem % S, nb - 1, rem // S)


@pytest.mark.parametrize("nb,S", [(1, 1), (1, 3), (2, 2), (3, 1), (5, 3), (8, 2), (32, 1)])
@pytest.mark.parametrize("with_w", [False, True])
def test_persistent_cholesky_task_order_is_topological(nb, S, with_w):
    """The persistent kernel hands tasks out in list order and a task waits for its inputs: no
    deadlock needs every input of a task to sit EARLIER in the list, and every tile exactly once."""
    total = S * nb * nb if with_w else S * nb * (nb + 1) // 2
    seen = {}
    for t in range(total):
        kind, s, i, j = _pp_decode(t, S, nb, with_w)
        assert 0 <= j <= i < nb and 0 <= s < S
        assert (kind, s, i, j) not in seen
        if kind == 0:
            for k in range(j):          # left-looking product: rows i and j up to column j
                assert (0, s, i, k) in seen and (0, s, j, k) in seen
            if i > j:
                assert (0, s, j, j) in seen            # inverse of the diagonal block
        else:
            assert i > j and (0, s, i, i) in seen      # W(i,i) comes from the diagonal task
            for k in range(j + 1, i):
                assert (1, s, k, j) in seen            # W(k,j) above it in the same column
            for k in range(j, i):
                assert (0, s, i, k) in seen            # row i of L
        seen[(kind, s, i, j)] = t
    n_factor = sum(1 for k in seen if k[0] == 0)
    assert n_factor == S * nb * (nb + 1) // 2
    assert len(seen) - n_factor == (S * nb * (nb - 1) // 2 if with_w else 0)


def test_simplex_start_points_match_the_reference(golden):
    """DomainService.get_points_domain with a simplex domain (sbo/services/domain.py:115-170)."""
    from bayesian_quadrature_optimization_b200.services.domain import DomainService
    g = golden("host_misc")
    bounds = [list(b) for b in g["simplex_bounds"]]
    np.random.seed(17)
    pts = np.array(DomainService.get_points_domain(7, bounds, type_bounds=6 * [0],
                                                   simplex_domain=int(g["simplex_domain"])))
    assert np.array_equal(pts, g["simplex_points"])
    assert np.all(pts[:, 2:].sum(1) <= int(g["simplex_domain"]))


def test_subsampled_finite_expectation_matches_the_reference(golden):
    """`n_samples` of parameters_distribution (uniform_finite, sbo/lib/expectations.py:37-42): the
    tasks drawn for a seed and the quadrature weights they induce, weighted and uniform."""
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import \
        BayesianQuadrature
    from bayesian_quadrature_optimization_b200.lib import constant as K
    g = golden("host_misc")
    T = len(g["weights"])
    pts = np.array([[0.1 * i, i % T] for i in range(2 * T)], dtype=float)
    td = {"points": pts.tolist(), "evaluations": list(np.arange(2 * T, dtype=float)), "var_noise": []}
    gp = GPFittingGaussian([K.PRODUCT_KERNELS_SEPARABLE, K.MATERN52_NAME, K.TASKS_KERNEL_NAME], td,
                           [2, 1, T], bounds_domain=[[0, 2], list(range(T))], type_bounds=[0, 1],
                           define_samplers=False)
    np.random.seed(23)
    bq = BayesianQuadrature(gp, [0], K.WEIGHTED_UNIFORM_FINITE, parameters_distribution={
        "weights": g["weights"], "domain_random": np.arange(T).reshape(T, 1), "n_samples": 11})
    assert np.array_equal(bq.task_sample, g["sub_weighted_index"])
    # E f = sum_t w_t f(t): the reference's sub-sampled mean of f(x, t) = table[t] * x at x = 2
    assert abs(np.dot(bq.weights, g["table"]) * 2.0 - float(g["sub_weighted"].reshape(-1)[0])) < 1e-14
    np.random.seed(29)
    bq = BayesianQuadrature(gp, [0], K.UNIFORM_FINITE, parameters_distribution={
        K.TASKS: T, "n_samples": 9})
    assert np.array_equal(bq.task_sample, g["sub_uniform_index"])
    assert abs(np.dot(bq.weights, g["table"]) * 2.0 - float(g["sub_uniform"].reshape(-1)[0])) < 1e-14


def test_quadrature_models_build_on_the_host_for_every_distribution():
    """BayesianQuadrature's constructor is host-only work: every supported distribution builds
    without a device, anything else raises NameError like the reference's dictionary lookup."""
    from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
    from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import \
        BayesianQuadrature
    from bayesian_quadrature_optimization_b200.lib import constant as K
    T = 3
    pts = np.array([[0.1 * i, i % T] for i in range(9)], dtype=float)
    td = {"points": pts.tolist(), "evaluations": list(np.arange(9.0)), "var_noise": []}
    gp = GPFittingGaussian([K.PRODUCT_KERNELS_SEPARABLE, K.MATERN52_NAME, K.TASKS_KERNEL_NAME], td,
                           [2, 1, T], bounds_domain=[[0, 1], list(range(T))], type_bounds=[0, 1],
                           define_samplers=False)
    bq = BayesianQuadrature(gp, [0], K.UNIFORM_FINITE, parameters_distribution={K.TASKS: T})
    assert bq.n_tasks == T and bq.weights is None and bq.separate_tasks
    bq = BayesianQuadrature(gp, [0], K.WEIGHTED_UNIFORM_FINITE, parameters_distribution={
        "weights": [0.2, 0.3, 0.5], "domain_random": np.arange(T).reshape(T, 1)})
    assert bq.task_continue and np.allclose(bq.weights, [0.2, 0.3, 0.5])
    with pytest.raises(NameError):
        BayesianQuadrature(gp, [0], "beta", parameters_distribution={})
    td2 = {"points": np.random.RandomState(0).uniform(0, 1, (6, 3)).tolist(),
           "evaluations": list(np.arange(6.0)), "var_noise": []}
    gp2 = GPFittingGaussian([K.SCALED_KERNEL, K.MATERN52_NAME], td2, [3],
                            bounds_domain=[[0, 1]] * 3, define_samplers=False)
    np.random.seed(1)
    bq2 = BayesianQuadrature(gp2, [0, 1], K.GAMMA, parameters_distribution={"a": [2.0], "scale": [1.0]})
    assert bq2.w_sets.shape == (1, 10, 1) and bq2.w_double_sets.shape == (2, 100, 1)
