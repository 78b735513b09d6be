#!/usr/bin/env python
"""bench.py -- BQO acquisition-plus-gradient throughput on the north-star configuration.

Workload (SURVEY.md section 8d, config C3 = BASELINE.json configs[2], the one the metric is
quoted on): synthetic GP, Scaled(Matern52) over d = 6 columns (d_x = 4, d_w = 2), n = 2048
training points, S = 20 hyper-samples, Z = 100 standard-normal samples, M_W = 10 gamma W samples
per quadrature, 10 000 candidates processed in batches.

A "step" is one pass of the hot path over one batch of CAND_PER_STEP candidates: stage B
(gamma / grad gamma rows, triangular solves against the 20 resident factors, denominators),
stage C (inner maximisation of a(x) + Z b(x) for every (candidate, hyper-sample, Z), gradient
of b at the maximisers) and the Monte-Carlo reduction -> acquisition value and gradient per
candidate.  Unit = one (candidate, hyper-sample) pair including all Z samples.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle port (the reference is
pure Python and cannot travel to the GPU box) on the host cores with a process pool, like the
reference's own mp.Pool.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_TRAIN, DX, DW = 2048, 4, 2
N_HYPER, N_Z, M_W, N_WSETS = 20, 100, 10, 8
N_CAND_TOTAL = 10000
# 111 x 20 = 2220 units = 15 per SM; one CTA per unit (all 100 Z samples): 2220 CTAs = 15 per SM
# (sweeps of {37,74,111,148} x {20,25,34,50,100} in profiles/r01_tuning_sweeps.md)
CAND_PER_STEP = int(os.environ.get("BQO_CAND_PER_STEP", "111"))
N_RESTARTS_MC, MAXITER_MC, FACTR_MC = 5, 10, 1e12  # bgo() defaults, sbo/services/bgo.py:43-44
Z_PER_CTA = int(os.environ.get("BQO_Z_PER_CTA", "100"))
ZS_SEED = 3


def make_workload(seed_shift=0):
    """seeds fixed so CPU and GPU arms see identical arrays."""
    rng = np.random.RandomState(0)
    x = rng.uniform(0, 1, (N_TRAIN, DX))
    w = rng.gamma(2.0, 1.0, (N_TRAIN, DW))
    pts = np.concatenate([x, w], 1)
    y = np.sin(pts.sum(1)) + 0.01 * rng.normal(0, 1, N_TRAIN)
    r2 = np.random.RandomState(2)
    base = np.array([1e-4, 0.0] + DX * [0.3] + DW * [2.0] + [1.0])
    thetas = base[None, :] * (1.0 + 0.05 * r2.normal(0, 1, (N_HYPER, base.size)))
    thetas[:, 1] = 0.05 * r2.normal(0, 1, N_HYPER)
    # Z samples and `starts_api` exactly as the public API draws them after np.random.seed(ZS_SEED)
    # (SBO.generate_samples_starting_points_evaluate_mc: Z first, then one uniform vector per
    # coordinate, services/domain.py): the end-to-end leg re-seeds before every step
    r3 = np.random.RandomState(ZS_SEED)
    zs = r3.normal(0, 1, N_Z)
    starts_api = np.array([r3.uniform(0, 1, N_RESTARTS_MC + 1) for _ in range(DX)]).T.copy()
    # the device-resident leg keeps round 1's start points, so that `value` stays comparable
    # between rounds; `e2e.device_resident_same_inputs` times it on the API's draws as well
    starts = np.random.RandomState(6).uniform(0, 1, (N_RESTARTS_MC + 1, DX))
    wsets = np.random.RandomState(4).gamma(2.0, 1.0, (N_WSETS, M_W, DW))
    r5 = np.random.RandomState(5 + seed_shift)
    cands = np.concatenate([r5.uniform(0, 1, (N_CAND_TOTAL, DX)),
                            r5.gamma(2.0, 1.0, (N_CAND_TOTAL, DW))], 1)
    return dict(pts=pts, y=y, thetas=thetas, zs=zs, wsets=wsets, cands=cands, starts=starts,
                starts_api=starts_api, lower=np.zeros(DX), upper=np.ones(DX))


def f_eval_flops():
    """Algorithmic FP64 flops of one evaluation of (a, b, grad a, grad b), SURVEY 8d stage C:
    M_W n (2d+10) + M_W n (3 d_x + 8) + 4 n (1 + d_x)."""
    d = DX + DW
    return M_W * N_TRAIN * (2 * d + 10) + M_W * N_TRAIN * (3 * DX + 8) + 4 * N_TRAIN * (1 + DX)


def issue_model(evals_per_launch, kernel_ms, clocks, dfma_peak_tflops=None):
    """What the kernel EXECUTES, next to the algorithmic count above: SASS instruction counts of the
    inner loop (one warp iteration = one training point x 10 W samples for the 32 evaluations held
    by the lanes; profiles/r02_inner_maximize.md) priced with
    the measured issue costs (FP64 2 cycles, 3 with three register sources, others 1), against
    the cycles the launch took at the sampled SM clock."""
    fp64, fp64_3src, other = 182.0, 25.3, 101.9
    cyc = 2.0 * fp64 + fp64_3src + other
    out = {"fp64_instr_per_point": fp64, "other_instr_per_point": other,
           "issue_cycles_per_point": cyc, "fp64_3src_instr_per_point": fp64_3src,
           "source": "executed-instruction counts of the ncu captures in "
           "profiles/r02_inner_maximize.md (283.9 warp instructions per point, 182 of them FP64; "
           "three-source DFMAs: the 35.3 of profiles/r01_inner_maximize_final.md less the ten of "
           "the square root, which the Halley step does without); "
           "148 SMs x 4 sub-partitions"}
    mhz = (clocks or {}).get("sm_mhz")
    if mhz:
        need = evals_per_launch / 32.0 * N_TRAIN * cyc / (148 * 4)
        # a ratio near (or a little above) 1: the launch takes what its instruction mix costs
        out["model_cycles_over_elapsed_cycles"] = need / (kernel_ms * 1e-3 * mhz * 1e6)
    rate = evals_per_launch * N_TRAIN * fp64 / (kernel_ms * 1e-3) / 1e12
    out["executed_fp64_thread_instr_T_per_s"] = rate
    if dfma_peak_tflops:
        out["executed_frac_of_dfma_issue_rate"] = rate / (dfma_peak_tflops / 2.0)
    return out


# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(np.max(smax))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------
REFERENCE_ROOT = "/root/reference"


def _limit_blas_threads():
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        # one BLAS/OpenMP thread per worker (the pool already uses every core); the environment
        # variable alone does not reach libraries the parent process loaded before the fork
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass


def _cpu_unit_worker(args):
    """One (candidate, hyper-sample) unit on the CPU oracle, timed for each Z count in `z_counts`
    so that the per-unit fixed cost and the per-Z cost can be separated."""
    cand_idx, hyper_idx, z_counts = args
    _limit_blas_threads()
    from oracle import bqo_oracle as orc
    wl = _WL
    spec = orc.KernelSpec(orc.SCALED_MATERN52, DX + DW)
    gp = orc.OracleGP(spec, wl["pts"], wl["y"])
    bq = orc.OracleBQ(gp, list(range(DX)), orc.GAMMA)
    sbo = orc.OracleSBO(bq, [(0.0, 1.0)] * DX)
    th = wl["thetas"][hyper_idx]
    gp.chol_solve(th[0], th[1], th[2:])  # factor cached per theta (untimed), as in the reference
    times = []
    for zc in z_counts:
        t0 = time.time()
        sbo.evaluate_mc_bayesian_candidate_points_no_restarts(
            wl["cands"][cand_idx:cand_idx + 1], [th], wl["zs"][:zc], wl["starts"],
            w=wl["wsets"][0], compute_gradient=True, factr=FACTR_MC, maxiter=MAXITER_MC)
        times.append(time.time() - t0)
    return times


def _ref_unit_worker(args):
    """The same unit on the UNMODIFIED reference (imported through tests/golden/ref_shim.py; only
    where /root/reference exists, i.e. in the build container): its own GPFittingGaussian,
    BayesianQuadrature and SBO.evaluate_mc_bayesian_candidate_points_no_restarts
    (sbo/acquisition_functions/sbo.py:523-670) with the Pool replaced by its sequential
    dispatcher -- the pool here is over units, as in _cpu_unit_worker."""
    cand_idx, hyper_idx, z_counts = args
    _limit_blas_threads()
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import ref_shim
    ref_shim.install()
    import warnings
    warnings.filterwarnings("ignore")
    from stratified_bayesian_optimization.models.gp_fitting_gaussian import GPFittingGaussian
    from stratified_bayesian_optimization.numerical_tools.bayesian_quadrature import BayesianQuadrature
    from stratified_bayesian_optimization.acquisition_functions.sbo import SBO
    from stratified_bayesian_optimization.lib.constant import SCALED_KERNEL, MATERN52_NAME, GAMMA
    wl = _WL
    th = wl["thetas"][hyper_idx]
    td = {"points": wl["pts"].tolist(), "evaluations": list(wl["y"]), "var_noise": []}
    gp = GPFittingGaussian([SCALED_KERNEL, MATERN52_NAME], td, [DX + DW],
                           bounds_domain=[[0, 1]] * DX + [[0, 20]] * DW, noise=True,
                           kernel_values=list(th[2:]), mean_value=[th[1]], var_noise_value=[th[0]],
                           define_samplers=False)
    gp.samples_parameters = [np.array(th)]
    bq = BayesianQuadrature(gp, list(range(DX)), GAMMA, parameters_distribution={"a": [2.0], "scale": [1.0]})
    sbo = SBO(bq, None)
    gp._cholesky_solve_vectors_for_posterior(th[0], th[1], th[2:])  # factor cached (untimed)
    times = []
    np.random.seed(3)
    for zc in z_counts:
        t0 = time.time()
        sbo.evaluate_mc_bayesian_candidate_points_no_restarts(
            wl["cands"][cand_idx:cand_idx + 1], 1, zc, n_restarts=N_RESTARTS_MC, n_threads=0,
            compute_max_mean=False, compute_gradient=True, factr=FACTR_MC, maxiter=MAXITER_MC)
        times.append(time.time() - t0)
    return times


_WL = None


def _host_cores():
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    return max(1, min(cores, 32))


def cpu_baseline(wl, z_counts=(1, 3), cores=None, impl=None):
    """The reference's CPU path for the bench workload, timed on the host cores.  `cores`
    independent units run concurrently in a process pool (the reference's own parallelism is an
    mp.Pool over units, sbo/lib/parallel.py:42-47).  Every unit is timed at the Z counts of
    `z_counts`; time(Z) = F + c Z is fitted by least squares over all of them and evaluated at the
    workload's Z = 100, so the fixed per-unit cost is not multiplied by 100.
    impl: "reference-shim" (the unmodified reference, needs /root/reference) or "port" (oracle/);
    default: the reference when it is there."""
    import multiprocessing as mp
    global _WL
    _WL = wl
    if impl is None:
        impl = "reference-shim" if os.path.isdir(REFERENCE_ROOT) else "port"
    worker = _ref_unit_worker if impl == "reference-shim" else _cpu_unit_worker
    cores = cores or _host_cores()
    jobs = [(i % 4, i % N_HYPER, tuple(z_counts)) for i in range(cores)]
    t0 = time.time()
    if cores > 1:
        ctx = mp.get_context("fork")
        with ctx.Pool(processes=cores) as pool:
            res = pool.map(worker, jobs)
    else:
        res = [worker(j) for j in jobs]
    wall = time.time() - t0
    t_mean = np.mean(np.array(res, dtype=float), axis=0)          # mean over units, per Z count
    zc = np.array(z_counts, dtype=float)
    if len(zc) > 1:
        c, F = np.polyfit(zc, t_mean, 1)
    else:
        c, F = t_mean[0] / zc[0], 0.0
    c = max(float(c), 1e-9)
    F = max(float(F), 0.0)
    t_full = F + c * N_Z
    value = cores / t_full
    timed = ", ".join("Z=%d: %.2f s" % (z, t) for z, t in zip(z_counts, t_mean))
    return {
        "value": value, "unit": "units/s", "cores": cores, "kind": impl,
        "sample": "%d concurrent units (1 candidate x 1 hyper-sample each, factor cached), each timed "
                  "at %s (of %d Z samples), n=%d, R=%d starts + 16 vertices, scipy L-BFGS-B "
                  "maxiter=%d; per-unit time fitted as F + c Z = %.2f + %.3f Z s, evaluated at Z=%d "
                  "(%.1f s per unit); %.1f s wall"
                  % (cores, timed, N_Z, N_TRAIN, N_RESTARTS_MC + 1, MAXITER_MC, F, c, N_Z, t_full, wall),
        "z_counts": list(z_counts), "seconds_at_z": [float(t) for t in t_mean],
        "wall_s": wall,
    }


def cpu_loglik_baseline(wl, reps=1):
    """log-likelihood + gradient per second on the CPU path
    (gp.log_likelihood + gp.grad_log_likelihood, sbo/models/gp_fitting_gaussian.py:772-897: one
    n x n dpotrs and one n^3 product per parameter), cold cache, BLAS free to use every core."""
    from oracle import bqo_oracle as orc
    spec = orc.KernelSpec(orc.SCALED_MATERN52, DX + DW)
    gp = orc.OracleGP(spec, wl["pts"], wl["y"])
    ts = []
    for r in range(reps):
        th = wl["thetas"][r % N_HYPER]
        gp.clean_cache()
        t0 = time.time()
        gp.log_likelihood(th[0], th[1], th[2:])
        gp.grad_log_likelihood(th[0], th[1], th[2:])
        ts.append(time.time() - t0)
    t = float(np.min(ts))
    return {"value": 1.0 / t, "unit": "evals/s", "cores": _host_cores(), "kind": "port",
            "sample": "%d cold-cache evaluation(s) of log_likelihood + grad_log_likelihood at n=%d, "
                      "p=%d kernel parameters, multithreaded BLAS: %.2f s each" % (reps, N_TRAIN, DX + DW + 1, t)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores, K timed
    steps after W warm-up steps; a step = every core evaluates one unit at Z = 1 and Z = 3.  One
    extra untimed pass at Z = 10 checks that the per-unit time is linear in Z."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload()
    vals = []
    base = None
    for s in range(args.warmup + args.steps):
        base = cpu_baseline(wl)
        if s >= args.warmup:
            vals.append(base)
    check = cpu_baseline(wl, z_counts=(10,))
    F_c = np.polyfit(vals[-1]["z_counts"], vals[-1]["seconds_at_z"], 1)
    value = float(np.mean([v["value"] for v in vals]))
    ms = float(np.mean([v["wall_s"] for v in vals])) * 1e3
    cb = dict(vals[-1])
    cb["value"] = value
    cb.pop("wall_s", None)
    cb["linearity_check"] = {"z": 10, "measured_s": check["seconds_at_z"][0],
                             "predicted_s": float(F_c[1] + F_c[0] * 10)}
    line = {
        "impl": "reference", "metric": "bqo_acq_grad_units_per_s", "value": value, "unit": "units/s",
        "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": cb,
        "secondary_cpu_baseline": cpu_loglik_baseline(wl),
        "e2e": {"value": value, "unit": "units/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config():
    return {
        "workload": "C3 synthetic GP: Scaled(Matern52) d=6 (d_x=4, d_w=2, gamma W), n=2048, "
                    "S=20 hyper-samples, Z=100, M_W=10, 10k candidates in batches",
        "n_training": N_TRAIN, "d_x": DX, "d_w": DW, "hyper_samples": N_HYPER, "z_samples": N_Z,
        "w_samples": M_W, "candidates_per_step": CAND_PER_STEP,
        "units_per_step": CAND_PER_STEP * N_HYPER,
        "inner_maximiser": "R=%d starts + 16 vertices, maxiter=%d, factr=%g"
                           % (N_RESTARTS_MC + 1, MAXITER_MC, FACTR_MC),
        "l2": "256 MiB buffer rewritten between steps (L2 flush); per-step inputs (20 factors, "
              "640 MB) exceed L2 as well",
        "precomputed": "per (hyper-sample, W set): W part of the squared distances (26 MB table, built "
                       "once with the engine, outside the timed steps like the factors); roofline "
                       "flops keep SURVEY 8d's count, which includes that part",
        "z_per_cta": Z_PER_CTA,
        "e2e_inputs": "same candidate batches and Z samples as the device-resident steps; the API "
                      "draws Z and its start points itself (np.random re-seeded with %d before every "
                      "end-to-end step), so its start points differ from the fixed ones of `value`: "
                      "e2e.device_resident_same_inputs is the device-resident step on the API's own "
                      "draws" % ZS_SEED,
    }



# ------------------------------------------------------------------------------------------------
# C5 (BASELINE.json configs[4]): n = 8192, S = 64 hyper-samples, 256 restarts, strong
# scaling over 1/2/4/8 GPUs with the hyper-samples sharded first.
C5_N, C5_S, C5_RESTARTS = 8192, 64, 256


def make_workload_c5():
    rng = np.random.RandomState(3)
    pts = np.concatenate([rng.uniform(0, 1, (C5_N, DX)), rng.gamma(2.0, 1.0, (C5_N, DW))], 1)
    y = np.sin(pts.sum(1)) + 0.01 * rng.normal(0, 1, C5_N)
    r2 = np.random.RandomState(2)
    base = np.array([1e-4, 0.0] + DX * [0.3] + DW * [2.0] + [1.0])
    thetas = base[None, :] * (1.0 + 0.05 * r2.normal(0, 1, (C5_S, base.size)))
    thetas[:, 1] = 0.05 * r2.normal(0, 1, C5_S)
    r5 = np.random.RandomState(5)
    cands = np.concatenate([r5.uniform(0, 1, (C5_RESTARTS, DX)), r5.gamma(2.0, 1.0, (C5_RESTARTS, DW))], 1)
    return dict(pts=pts, y=y, thetas=thetas, cands=cands,
                zs=np.random.RandomState(3).normal(0, 1, N_Z),
                wsets=np.random.RandomState(4).gamma(2.0, 1.0, (1, M_W, DW)),
                starts=np.random.RandomState(6).uniform(0, 1, (N_RESTARTS_MC + 1, DX)),
                lower=np.zeros(DX), upper=np.ones(DX))


def run_c5(args, dev, rank, world, steps, warmup=1):
    """One acquisition evaluation (value + gradient of all 256 restart points) per step.  Rank r
    holds the factors of hyper-samples [r S/N, (r+1) S/N) only (32 GB of factors in total at
    n = 8192: what bounds memory) and scores every restart point against them; the
    per-candidate partial sums [C x (1 + d)] are combined INSIDE the timed step with one
    all_gather and a rank-ordered sum (bit-identical totals and argmax on every rank,
    distributed.allgather_ordered_sum).  Strong scaling: the total work per step is fixed."""
    import torch
    import torch.distributed as dist
    from bayesian_quadrature_optimization_b200 import _cabi, engine as E, distributed as D
    wl = make_workload_c5()
    eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, DX + DW, device=dev)
    eng.set_data(wl["pts"], wl["y"])
    lo, hi = D.shard_range(C5_S, rank, world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    hb = eng.hypers(wl["thetas"][lo:hi]).factorize()
    e1.record()
    torch.cuda.synchronize()
    factor_ms = e0.elapsed_time(e1)
    assert np.all(hb.info == 0)
    bq = E.BQEngine(hb, DX, wsets=wl["wsets"], lower=wl["lower"], upper=wl["upper"])
    cands = eng.to_device(wl["cands"])
    zs, starts = eng.to_device(wl["zs"]), eng.to_device(wl["starts"])
    n_loc = float(hi - lo)
    out = {}

    def step():
        res = bq.evaluate_candidates(cands, zs, starts, maxiter=MAXITER_MC, factr=FACTR_MC)
        part = torch.cat([res["evaluations"].reshape(-1, 1), res["gradient"]], dim=1) * n_loc
        tot = D.allgather_ordered_sum(part) / float(C5_S)
        out["argmax"] = torch.argmax(tot[:, 0])
        out["tot"] = tot

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    both = torch.tensor([ms, factor_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(both, op=dist.ReduceOp.MAX)
    ms, factor_ms = float(both[0].item()), float(both[1].item())
    # every rank must hold the same bits (rank-ordered sum)
    same = True
    if world > 1:
        bufs = [torch.empty_like(out["tot"]) for _ in range(world)]
        dist.all_gather(bufs, out["tot"])
        same = all(torch.equal(bufs[0], b) for b in bufs)
    units = C5_RESTARTS * C5_S
    res = {
        "workload": "C5 scaling sweep: Scaled(Matern52) d=6, n=%d, S=%d hyper-samples sharded over "
                    "the ranks (%d per GPU), %d restart points, Z=%d, M_W=%d"
                    % (C5_N, C5_S, hi - lo, C5_RESTARTS, N_Z, M_W),
        "metric": "bqo_acq_grad_units_per_s", "value": units * steps / (ms * 1e-3), "unit": "units/s",
        "scaling": "strong", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms / steps, "units_per_step": units,
        "factorize_ms_max_over_ranks": factor_ms,
        "factor_bytes_per_gpu": float(hi - lo) * C5_N * C5_N * 8.0,
        "exchange": "per step, inside the timed region: all_gather of the [%d x %d] FP64 partial sums "
                    "+ rank-ordered sum (%d bytes per rank)" % (C5_RESTARTS, 1 + DX + DW,
                                                               C5_RESTARTS * (1 + DX + DW) * 8),
        "identical_bits_on_all_ranks": bool(same), "argmax": int(out["argmax"].item()),
    }
    del bq, hb, eng
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import ctypes as C
    import torch
    import torch.distributed as dist

    from bayesian_quadrature_optimization_b200 import _cabi, engine as E

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL_DEBUG is left as the caller set it (the driver counts ranks in the INFO log); the
        # JSON line is the only line this program itself prints
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    wl = make_workload()
    # weak scaling: every rank owns its own slice of the candidate list and a full copy of the
    # training data and of the S factors; the only exchange is the final argmax (one all_gather).
    eng = E.GPEngine(_cabi.KIND_SCALED_MATERN52, DX + DW, device=dev)
    eng.set_data(wl["pts"], wl["y"])

    # FP64 peak of this GPU, measured now (MEASURED_PEAKS.json has no FP64 entry)
    lib = eng.lib
    ms, fl = C.c_double(), C.c_double()
    probe_ws = torch.empty(int(lib.bqo_fp64_peak_probe_workspace_bytes()), dtype=torch.uint8, device=dev)
    _cabi.check(lib.bqo_fp64_peak_probe(0, 100000, probe_ws.data_ptr(), C.byref(ms), C.byref(fl)), "probe")
    dfma_peak = fl.value / (ms.value * 1e-3) / 1e12
    _cabi.check(lib.bqo_fp64_peak_probe(1, 100000, probe_ws.data_ptr(), C.byref(ms), C.byref(fl)), "probe")
    dmma_peak = fl.value / (ms.value * 1e-3) / 1e12
    # independent anchor for the FP64 tensor peak: cuBLAS DGEMM 4096^3 through torch.matmul
    ga = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    gb = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    torch.matmul(ga, gb)
    torch.cuda.synchronize()
    dg = []
    for rep in range(4):
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        torch.matmul(ga, gb)
        g1.record()
        torch.cuda.synchronize()
        dg.append(g0.elapsed_time(g1))
    dgemm_peak = 2.0 * 4096.0 ** 3 / (min(dg) * 1e-3) / 1e12
    del ga, gb
    tensor_peak = max(dmma_peak, dgemm_peak)   # the larger of probe and anchor is the roofline

    # ---- stage A (once per BO iteration; secondary metric: log-likelihood + gradient per s) ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_a_ms = []
    for rep in range(2 if args.profile else 3):
        hb = eng.hypers(wl["thetas"])
        torch.cuda.synchronize()
        ev0.record()
        hb.factorize(with_inverse=True)   # log-likelihood AND gradient are asked for
        llh = hb.log_likelihood()
        gll = hb.grad_log_likelihood()
        ev1.record()
        torch.cuda.synchronize()
        stage_a_ms.append(ev0.elapsed_time(ev1))
    stage_a = float(np.min(stage_a_ms[1:]))
    assert np.all(hb.info == 0)
    # kernel-matrix assembly alone: the one HBM-bound kernel of stage A (S n^2 doubles written)
    asm_ms = []
    for rep in range(4):
        ev0.record()
        kmat = hb.cov(add_noise=True)
        ev1.record()
        torch.cuda.synchronize()
        asm_ms.append(ev0.elapsed_time(ev1))
    del kmat
    asm_ms = float(np.min(asm_ms[1:]))
    asm_bytes = float(N_HYPER) * N_TRAIN * N_TRAIN * 8.0
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    except Exception:
        hbm_peak, hbm_src = 6550.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
    n = N_TRAIN
    p = DX + DW + 1
    f_ll = n * n * (2 * (DX + DW) + 10) + n ** 3 / 3.0 + 2 * n * n + 2 * n ** 3 / 3.0 \
        + p * n * n * (2 * (DX + DW) + 12)

    bq = E.BQEngine(hb, DX, wsets=wl["wsets"], lower=wl["lower"], upper=wl["upper"])
    zs = eng.to_device(wl["zs"])
    starts = eng.to_device(wl["starts"])
    cands_all = eng.to_device(wl["cands"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    per_rank = N_CAND_TOTAL // world
    opt = dict(maxiter=MAXITER_MC, factr=FACTR_MC, z_per_cta=Z_PER_CTA)

    def batch(step):
        lo = rank * per_rank + (step * CAND_PER_STEP) % max(1, per_rank - CAND_PER_STEP)
        return cands_all[lo:lo + CAND_PER_STEP]

    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    evals_total = [0]
    best = [None]

    from bayesian_quadrature_optimization_b200 import distributed as D

    def step(i, timed):
        cb = batch(i)
        ub = bq.setup(cb)
        k0.record()
        out_max, out_arg, n_evals = bq.inner_maximize(ub, zs, starts, count_evals=True, **opt)
        k1.record()
        gvb, is_nan = bq.grad_vector_b(ub, out_arg)
        value, grad = bq.acq_reduce(ub, out_max, gvb, is_nan, zs)
        v, idx = torch.max(value, dim=0)
        if world > 1:
            # the one exchange of the path, once per step: every rank's best (value, global
            # candidate index, point) -> the same argmax on all ranks (distributed.allgather_best)
            lo_ = rank * per_rank + (i * CAND_PER_STEP) % max(1, per_rank - CAND_PER_STEP)
            best[0] = D.allgather_best(v, idx + lo_, cb[idx])
        else:
            best[0] = (v, idx + 0)
        flush.fill_(i & 0xFF)
        return n_evals

    for i in range(args.warmup):
        step(i, False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    launches0 = int(lib.bqo_launch_count())
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    counts = []
    kevs = []
    for i in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0, k1 = a, b
        counts.append(step(args.warmup + i, True))
        kevs.append((a, b))
    t1.record()
    torch.cuda.synchronize()
    gpu_launches = int(lib.bqo_launch_count()) - launches0   # counted by the library itself
    elapsed_ms = t0.elapsed_time(t1)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    kern_ms = [a.elapsed_time(b) for a, b in kevs]
    evals = float(sum(int(c.sum().item()) for c in counts))
    units_per_step = CAND_PER_STEP * N_HYPER
    value = world * units_per_step * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the public API (host numpy in, numpy out) ---------------------------
    e2e = None
    if not args.profile:
        from bayesian_quadrature_optimization_b200.models.gp_fitting_gaussian import GPFittingGaussian
        from bayesian_quadrature_optimization_b200.numerical_tools.bayesian_quadrature import \
            BayesianQuadrature
        from bayesian_quadrature_optimization_b200.acquisition_functions.sbo import SBO
        from bayesian_quadrature_optimization_b200.lib.constant import SCALED_KERNEL, MATERN52_NAME, GAMMA
        td = {"points": wl["pts"], "evaluations": wl["y"], "var_noise": []}
        bounds = [[0, 1]] * DX + [[0, 20]] * DW
        gp = GPFittingGaussian([SCALED_KERNEL, MATERN52_NAME], td, [DX + DW], bounds_domain=bounds,
                               noise=True, kernel_values=list(wl["thetas"][0][2:]),
                               mean_value=[0.0], var_noise_value=[1e-4], device=dev)
        gp.samples_parameters = [t.copy() for t in wl["thetas"]]
        quad = BayesianQuadrature(gp, list(range(DX)), GAMMA,
                                  parameters_distribution={"a": [2.0], "scale": [1.0]})
        quad.set_w_sets(wl["wsets"])
        sbo = SBO(quad)
        np.random.seed(7 + rank)
        host_cands = wl["cands"]

        def e2e_step(i):
            np.random.seed(ZS_SEED)   # the API draws Z and the starts itself: same draws as wl
            lo = rank * per_rank + (i * CAND_PER_STEP) % max(1, per_rank - CAND_PER_STEP)
            r = sbo.evaluate_mc_bayesian_candidate_points_no_restarts(
                host_cands[lo:lo + CAND_PER_STEP], N_HYPER, N_Z, n_restarts=N_RESTARTS_MC,
                compute_gradient=True, factr=FACTR_MC, maxiter=MAXITER_MC)
            return r

        for i in range(max(1, min(args.warmup, 2))):
            e2e_step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        w0 = time.perf_counter()
        n_e2e = max(1, args.steps)
        for i in range(n_e2e):
            r = e2e_step(args.warmup + i)   # the candidate batches of the device-resident steps
        torch.cuda.synchronize()
        dt = time.perf_counter() - w0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = CAND_PER_STEP * (DX + DW) * 8 + N_Z * 8 + (N_RESTARTS_MC + 1) * DX * 8
        d2h = CAND_PER_STEP * 8 + CAND_PER_STEP * (DX + DW) * 8
        e2e = {"value": world * units_per_step * n_e2e / dt, "unit": "units/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e}
        # the device-resident step on exactly the inputs the API drew (same Z, the API's start
        # points, same candidate batches): what the end-to-end number is to be compared with
        starts_dev = starts
        starts = eng.to_device(wl["starts_api"])
        step(args.warmup, False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        cnt2 = [step(args.warmup + i, True) for i in range(n_e2e)]
        s1.record()
        torch.cuda.synchronize()
        ms2 = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([ms2], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        starts = starts_dev
        e2e["device_resident_same_inputs"] = world * units_per_step * n_e2e / (ms2 * 1e-3)
        e2e["evals_per_unit_per_z_same_inputs"] = float(
            sum(int(c.sum().item()) for c in cnt2)) / n_e2e / (units_per_step * N_Z)

    c5 = None
    if not args.profile and not args.no_c5:
        del bq, hb
        torch.cuda.empty_cache()
        c5 = run_c5(args, dev, rank, world, steps=max(1, min(args.steps, args.c5_steps)))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    kern_avg_ms = float(np.mean(kern_ms))
    flops_per_launch = evals / args.steps * f_eval_flops()
    achieved = flops_per_launch / (kern_avg_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "inner_maximize_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    cpu = cpu_ll = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(wl)
        cpu.pop("wall_s", None)
        check = cpu_baseline(wl, z_counts=(10,))
        fit = np.polyfit(cpu["z_counts"], cpu["seconds_at_z"], 1)
        cpu["linearity_check"] = {"z": 10, "measured_s": check["seconds_at_z"][0],
                                  "predicted_s": float(fit[1] + fit[0] * 10)}
        cpu_ll = cpu_loglik_baseline(wl)
    line = {
        "metric": "bqo_acq_grad_units_per_s", "value": value, "unit": "units/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(),
        "roofline": {
            "bound": "fp64", "kernel": "inner_maximize_kernel<4,2>", "achieved": achieved,
            "peak": dfma_peak, "unit": "TFLOP/s", "frac": achieved / dfma_peak,
            "traffic": traffic,
            "peak_source": "FP64 DFMA (non-tensor) rate measured in this run by bqo_fp64_peak_probe "
                           "(MEASURED_PEAKS.json holds no FP64 figure).  Independent anchors measured "
                           "in the same run: DMMA m8n8k4 probe %.1f TF/s, cuBLAS DGEMM 4096^3 "
                           "(torch.matmul) %.1f TF/s; architectural rate 148 SM x 64 lanes x 2 x "
                           "clock" % (dmma_peak, dgemm_peak),
            "anchors": {"dfma_probe_tflops": dfma_peak, "dmma_probe_tflops": dmma_peak,
                        "cublas_dgemm_tflops": dgemm_peak},
            "flops_per_eval": f_eval_flops(), "evals_per_launch": evals / args.steps,
            "evals_per_unit_per_z": evals / args.steps / (units_per_step * N_Z),
            "kernel_ms": kern_avg_ms, "kernel_share_of_step": kern_avg_ms / (elapsed_ms / args.steps),
            "issue_model": issue_model(evals / args.steps, kern_avg_ms, clocks, dfma_peak),
        },
        "secondary": {
            "metric": "gp_loglik_grad_per_s", "value": N_HYPER / (stage_a * 1e-3), "unit": "evals/s",
            "ms_for_%d_hyper_samples" % N_HYPER: stage_a,
            "achieved_tflops": N_HYPER * f_ll / (stage_a * 1e-3) / 1e12, "peak_tflops": tensor_peak,
            "frac": N_HYPER * f_ll / (stage_a * 1e-3) / 1e12 / tensor_peak,
            "bound": "fp64 tensor (DMMA)", "flops_per_eval": f_ll,
            "peak_source": "max(DMMA m8n8k4 probe %.1f, cuBLAS DGEMM %.1f) TF/s, both measured in "
                           "this run" % (dmma_peak, dgemm_peak),
            "cpu_baseline": cpu_ll,
        },
        "assembly": {
            "metric": "kernel_matrix_assembly", "ms_for_%d_hyper_samples" % N_HYPER: asm_ms,
            "bound": "hbm", "achieved": asm_bytes / (asm_ms * 1e-3) / 1e9, "peak": hbm_peak,
            "unit": "GB/s", "frac": asm_bytes / (asm_ms * 1e-3) / 1e9 / hbm_peak,
            "bytes": asm_bytes, "peak_source": hbm_src,
        },
        "cpu_baseline": cpu,
        "e2e": e2e,
        "c5": c5,
        "gpu_launches": gpu_launches,
        "clocks": clocks,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    global CAND_PER_STEP
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the C5 strong-scaling object")
    ap.add_argument("--c5-steps", type=int, default=2,
                    help="timed steps of the C5 object (a step takes ~14 s on one GPU)")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: 8 candidates per step, no stage-A repeats, no e2e, no CPU leg")
    args = ap.parse_args()
    if args.profile:
        CAND_PER_STEP = 8
        args.no_cpu_baseline = True
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
