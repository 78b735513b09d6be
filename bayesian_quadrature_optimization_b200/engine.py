"""Device engine: owns the HBM-resident state of one GP model and drives libbqo_sm100.so.

PyTorch is used for buffer ownership, pinned host<->device copies and the CUDA stream only;
every numerical step is a call into the C ABI (include/bqo_sm100.h).  There is no CPU path:
constructing an engine without CUDA raises.

Layout in HBM (all FP64, row-major):
    X      [n][dim]           raw training inputs (x columns, then w / task column)
    y      [n]                evaluations
    theta  [S][2+p]           hyper-samples (var_noise, mean, kernel parameters)
    hyp    [S][stride]        prepared hyper blocks (ls, 1/ls, sigma2, task covariance tables)
    Xs     [S][dim_m][n]      X / ls per hyper-sample, structure of arrays
    L      [S][n][n]          Cholesky factors of Sigma_s (lower, strict upper zeroed)
    alpha  [S][n]             Sigma_s^-1 (y - mean_s)
    per candidate batch: G, R [S][C*(1+dim_m)][n]  (gamma | grad gamma rows, unsolved / solved)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from ._cabi import KernelDesc, StageC, InnerOpt, check, ptr

F64 = torch.float64
I32 = torch.int32


def _cabi_taskc_offset() -> int:
    """Offset (doubles) of the task covariance table inside a hyper block (BQO_H_TASKC)."""
    return 4 + 2 * _cabi.MAX_DIM


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def kernel_desc(kind: int, dim: int, n_tasks: int = 0, same_correlation: bool = False) -> KernelDesc:
    dim_m = dim - 1 if kind == _cabi.KIND_PRODUCT_TASKS else dim
    if kind == _cabi.KIND_MATERN52:
        n_params = dim
    elif kind == _cabi.KIND_SCALED_MATERN52:
        n_params = dim + 1
    else:
        nt = min(n_tasks, 2) if same_correlation else n_tasks * (n_tasks + 1) // 2
        n_params = dim_m + nt
    return KernelDesc(kind, dim, dim_m, n_tasks, int(bool(same_correlation)), n_params)


class GPEngine(object):
    """Training data of one GP on one GPU."""

    def __init__(self, kind, dim, n_tasks=0, same_correlation=False, device=None):
        if not torch.cuda.is_available():
            raise _cabi.BqoLibraryError(
                "CUDA device required: bayesian_quadrature_optimization_b200 has no CPU path")
        self.lib = _cabi.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        check(self.lib.bqo_set_device(self.device.index or 0), "bqo_set_device")
        self.desc = kernel_desc(kind, dim, n_tasks, same_correlation)
        self.hyp_stride = int(self.lib.bqo_hyper_stride(C.byref(self.desc)))
        self.n = 0
        self.X = self.y = self.noise_obs = None

    # -- data ---------------------------------------------------------------------------------
    def to_device(self, a, dtype=F64):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64 if dtype == F64 else np.int32))
        return torch.from_numpy(arr).to(self.device, non_blocking=False)

    def set_data(self, points, evaluations, var_noise_obs=None):
        self.X = self.to_device(points).reshape(-1, self.desc.dim)
        self.y = self.to_device(evaluations).reshape(-1)
        self.n = int(self.X.shape[0])
        self.noise_obs = None if var_noise_obs is None else self.to_device(var_noise_obs).reshape(-1)

    # Guard bands (debugging aid; compute-sanitizer is not available on the B200 pool): with
    # GPEngine.guard_doubles > 0 every buffer handed to the library sits between two bands of a
    # known bit pattern, and check_guards() reports any band a kernel wrote into.
    guard_doubles = 0
    _guards = []
    _GUARD_VALUE = -6.02214076e23

    def empty(self, *shape, dtype=F64):
        g = GPEngine.guard_doubles
        if g <= 0 or dtype != F64:
            return torch.empty(*shape, dtype=dtype, device=self.device)
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        numel = 1
        for d in shape:
            numel *= int(d)
        raw = torch.full((numel + 2 * g,), GPEngine._GUARD_VALUE, dtype=F64, device=self.device)
        GPEngine._guards.append((raw, g, numel, tuple(int(d) for d in shape)))
        return raw[g:g + numel].view(*shape)

    @staticmethod
    def check_guards(clear=True):
        """Number of guard bands that no longer hold the pattern (0 = no out-of-bounds write)."""
        bad = []
        for raw, g, numel, shape in GPEngine._guards:
            lo, hi = raw[:g], raw[g + numel:]
            if not (bool((lo == GPEngine._GUARD_VALUE).all()) and bool((hi == GPEngine._GUARD_VALUE).all())):
                bad.append(shape)
        if clear:
            GPEngine._guards = []
        return bad

    def hypers(self, thetas) -> "HyperBatch":
        return HyperBatch(self, thetas)


def potrf_batched(eng: GPEngine, A: torch.Tensor):
    """In-place lower Cholesky of A[S][n][n] (bqo_potrf_batched with its scratch) -> info[S]."""
    S, n = int(A.shape[0]), int(A.shape[1])
    info = torch.zeros(S, dtype=I32, device=A.device)
    work = torch.empty(int(eng.lib.bqo_potrf_workspace_bytes(n, S)) // 8 + 2, dtype=F64,
                       device=A.device)
    check(eng.lib.bqo_potrf_batched(ptr(A), n, S, ptr(info), ptr(work), _stream()),
          "bqo_potrf_batched")
    return info


class HyperBatch(object):
    """S hyper-samples resident on the device with their prepared blocks and factors."""

    def __init__(self, eng: GPEngine, thetas):
        self.eng = eng
        lib, d = eng.lib, eng.desc
        self.theta = eng.to_device(thetas).reshape(-1, 2 + d.n_params)
        self.S = int(self.theta.shape[0])
        self.hyp = eng.empty(self.S, eng.hyp_stride)
        check(lib.bqo_hyper_prepare(C.byref(d), ptr(self.theta), self.S, ptr(self.hyp), _stream()),
              "bqo_hyper_prepare")
        self.Xs = eng.empty(self.S, d.dim_m, eng.n)
        self.tid = eng.empty(eng.n, dtype=I32) if d.kind == _cabi.KIND_PRODUCT_TASKS else None
        check(lib.bqo_scale_points(C.byref(d), ptr(self.hyp), self.S, ptr(eng.X), eng.n, d.dim,
                                   ptr(self.Xs), ptr(self.tid), _stream()), "bqo_scale_points")
        self.L = None
        self.Winv = None
        self.alpha = None
        self._info = None
        self._jitter = None

    # -- kernel matrices ------------------------------------------------------------------------
    def cov(self, add_noise=False):
        eng = self.eng
        K = eng.empty(self.S, eng.n, eng.n)
        check(eng.lib.bqo_kernel_assemble_batched(
            C.byref(eng.desc), ptr(self.hyp), self.S, ptr(self.Xs), ptr(self.tid), eng.n,
            ptr(eng.noise_obs) if add_noise else None, int(add_noise), None, ptr(K), _stream()),
            "bqo_kernel_assemble_batched")
        return K

    def grad_cov(self):
        eng = self.eng
        dK = eng.empty(self.S, eng.desc.n_params, eng.n, eng.n)
        check(eng.lib.bqo_kernel_grad_params_batched(
            C.byref(eng.desc), ptr(self.hyp), self.S, ptr(self.Xs), ptr(self.tid), eng.n, ptr(dK),
            _stream()), "bqo_kernel_grad_params_batched")
        return dK

    def cross_cov(self, P, grad=False):
        """k(P, X) [S][m][n] (and d/dP [S][m][dim_m][n])."""
        eng = self.eng
        P = eng.to_device(P).reshape(-1, eng.desc.dim)
        m = int(P.shape[0])
        KP = eng.empty(self.S, m, eng.n)
        dKP = eng.empty(self.S, m, eng.desc.dim_m, eng.n) if grad else None
        check(eng.lib.bqo_cross_cov_batched(
            C.byref(eng.desc), ptr(self.hyp), self.S, ptr(self.Xs), ptr(self.tid), eng.n, ptr(P), m,
            0, ptr(KP), ptr(dKP), _stream()), "bqo_cross_cov_batched")
        return (KP, dKP) if grad else KP

    def cross_cov_hessian(self, P):
        """d2 k(P_q, X_i) / dP dP: [S][m][dim_m][dim_m][n] (a5)."""
        eng = self.eng
        P = eng.to_device(P).reshape(-1, eng.desc.dim)
        m = int(P.shape[0])
        dm = eng.desc.dim_m
        out = eng.empty(self.S, m, dm, dm, eng.n)
        check(eng.lib.bqo_cross_cov_hessian_batched(
            C.byref(eng.desc), ptr(self.hyp), self.S, ptr(self.Xs), ptr(self.tid), eng.n, ptr(P), m,
            ptr(out), _stream()), "bqo_cross_cov_hessian_batched")
        return out

    # -- factorisation ----------------------------------------------------------------------------
    def factorize(self, max_tries=7, with_inverse=False):
        """Sigma_s = K_s + noise; L_s (first attempt of the reference's jitter policy); alpha_s.
        Nothing is copied to the host here: `info` / `jitter` resolve lazily (one device-to-host
        read) and continue the policy for the matrices that failed.
        with_inverse: the same launch also leaves W_s = L_s^-1 (the log-likelihood gradient forms
        Sigma^-1 = W^T W from it), as extra tasks that fill the SMs the factorisation leaves idle."""
        if self.L is not None:
            return self
        eng = self.eng
        S = self.S
        self.L = eng.empty(S, eng.n, eng.n)
        self.Winv = eng.empty(S, eng.n, eng.n) if with_inverse else None
        wscratch = eng.empty(S, eng.n, eng.n) if (with_inverse and eng.n % 2 == 1) else None
        self._max_tries = int(max_tries)
        self._info_dev = torch.zeros(S, dtype=I32, device=eng.device)
        self._diag_mean = eng.empty(S)
        self._diag_nonpos = torch.zeros(S, dtype=I32, device=eng.device)
        nbytes = int(eng.lib.bqo_potrf_workspace_bytes(eng.n, S))
        work = torch.empty(nbytes // 8 + 2, dtype=F64, device=eng.device)
        check(eng.lib.bqo_chol_cov_batched(
            C.byref(eng.desc), ptr(self.hyp), S, ptr(self.Xs), ptr(self.tid), eng.n,
            ptr(eng.noise_obs), ptr(self.L), ptr(self._info_dev), ptr(self._diag_mean),
            ptr(self._diag_nonpos), ptr(self.Winv), ptr(wscratch), ptr(work), _stream()),
            "bqo_chol_cov_batched")
        self._info = self._jitter = None
        self._ill = None
        self._alpha_refined = False
        self._alpha()
        return self

    def _alpha(self):
        eng = self.eng
        if self.alpha is None:
            self.alpha = eng.empty(self.S, eng.n)
        if self.Winv is not None:
            # L^-1 is there: alpha = W^T (W r) in two passes over W instead of ~70 solve launches
            work = eng.empty(self.S, eng.n)
            check(eng.lib.bqo_alpha_from_inverse_batched(
                ptr(self.hyp), eng.hyp_stride, ptr(self.Winv), eng.n, self.S, ptr(eng.y),
                ptr(self.alpha), ptr(work), _stream()), "bqo_alpha_from_inverse_batched")
            return
        check(eng.lib.bqo_alpha_batched(ptr(self.hyp), eng.hyp_stride, ptr(self.L), eng.n, self.S,
                                        ptr(eng.y), ptr(self.alpha), _stream()), "bqo_alpha_batched")

    def _resolve(self, info_host=None):
        """Host copy of the factorisation codes; runs the rest of the reference's jitter policy
        (sbo/lib/la_functions.py:19-38: "not positive definite matrix" when a diagonal entry is
        <= 0, else Sigma + jitter I with jitter = mean(diag) 1e-6 10^t, t < max_tries) for the
        matrices whose first attempt failed.  -> info: 0 ok, -1, -2 as in the two LinAlgErrors."""
        if self._info is not None:
            return
        eng = self.eng
        S = self.S
        if info_host is None:
            info_host = self._info_dev.cpu().numpy()
        info = np.zeros(S, dtype=np.int32)
        jitter = np.zeros(S, dtype=np.float64)
        bad = np.nonzero(info_host)[0]
        if len(bad) > 0:
            if np.any(info_host[bad] == -3):
                raise _cabi.BqoLibraryError("persistent Cholesky: scheduling wait limit reached")
            mean = self._diag_mean.cpu().numpy()
            nonpos = self._diag_nonpos.cpu().numpy()
            work1 = torch.empty(int(eng.lib.bqo_potrf_workspace_bytes(eng.n, 1)) // 8 + 2, dtype=F64,
                                device=eng.device)
            jd = eng.empty(1)
            one = torch.zeros(1, dtype=I32, device=eng.device)
            for s in bad:
                if nonpos[s]:
                    info[s] = -1
                    continue
                jit = mean[s] * 1e-6
                ok = False
                for _ in range(self._max_tries):
                    if not np.isfinite(jit):
                        break
                    check(eng.lib.bqo_chol_cov_retry(
                        C.byref(eng.desc), ptr(self.hyp), int(s), ptr(self.Xs), ptr(self.tid), eng.n,
                        ptr(eng.noise_obs), float(jit), ptr(jd), ptr(self.L), ptr(one), ptr(work1),
                        _stream()), "bqo_chol_cov_retry")
                    if int(one.item()) == 0:
                        ok = True
                        break
                    jit *= 10.0
                if ok:
                    jitter[s] = jit
                else:
                    info[s] = -2
            self.Winv = None  # refactored matrices: the gradient rebuilds L^-1 from L
            self._ill = None
            self._alpha_refined = False
            self._alpha()
        self._info, self._jitter = info, jitter

    @property
    def info(self):
        if self.L is None:
            return None
        self._resolve()
        return self._info

    @info.setter
    def info(self, value):
        self._info = value

    @property
    def jitter(self):
        if self.L is None:
            return None
        self._resolve()
        return self._jitter

    @jitter.setter
    def jitter(self, value):
        self._jitter = value

    def log_likelihood_host(self):
        """np.array(S) of log-likelihoods with -inf where the factorisation failed: the values and
        the factorisation codes come back in ONE device-to-host copy (the slice sampler's probes)."""
        self.factorize()
        llh = self.log_likelihood()
        if self._info is None:
            both = torch.cat([llh, self._info_dev.to(F64)]).cpu().numpy()
            codes = both[self.S:].astype(np.int32)
            if not np.any(codes):
                self._resolve(info_host=codes)
                return both[:self.S]
            self._resolve(info_host=codes)
            llh = self.log_likelihood()
        out = llh.cpu().numpy()
        out[self._info != 0] = -np.inf
        return out

    def appended(self, eng2: GPEngine) -> "HyperBatch":
        """The same hyper-samples on `eng2`, whose training set is this engine's plus ONE point
        appended last: every factor is extended in O(n^2) (bqo_chol_append_batched) instead of
        refactorised in O(n^3); alpha is recomputed with two vector solves.  Falls back to a
        full factorisation when an extended matrix is not positive definite."""
        eng = self.eng
        if self.L is None or eng2.n != eng.n + 1:
            raise ValueError("appended() needs a factorised batch and exactly one new point")
        n, S = eng.n, self.S
        hb2 = HyperBatch(eng2, self.theta)
        if np.any(self.info != 0) or np.any(self.jitter > 0):
            # a failed factor is not a factor, and a jittered one was built for the old mean(diag):
            # the reference restarts the jitter policy from 0 on the new matrix (la_functions.py:8-38)
            return hb2.factorize()
        knew = hb2.cross_cov(eng2.X[n:n + 1]).reshape(S, n + 1).clone()
        extra = self.theta[:, 0] + eng.to_device(self.jitter)
        if eng2.noise_obs is not None:
            extra = extra + eng2.noise_obs[n]
        knew[:, n] += extra
        L_new = eng2.empty(S, n + 1, n + 1)
        info = torch.zeros(S, dtype=I32, device=eng.device)
        work = eng2.empty(S * n)
        check(eng.lib.bqo_chol_append_batched(ptr(self.L), n, S, ptr(knew), ptr(L_new), ptr(info),
                                              ptr(work), _stream()), "bqo_chol_append_batched")
        if bool(info.any().item()):
            return hb2.factorize()
        hb2.L = L_new
        hb2.Winv = None
        hb2._max_tries = self._max_tries
        hb2.info = np.zeros(S, dtype=np.int32)
        hb2.jitter = self.jitter.copy()
        hb2._alpha()
        return hb2

    # -- solves ----------------------------------------------------------------------------------
    # The blocked factorisation and the blocked triangular solves multiply by the explicitly
    # inverted 64 x 64 diagonal blocks, so their backward error grows with the condition number of
    # those blocks (LAPACK's substitution does not).  Harmless up to cond(Sigma) ~ 1e6; at the
    # reference's own test vector (theta = [1, 5, 50, 9.6, -3, -0.1], cond = 7e9) b(x), which
    # cancels to 1e-13 of its terms, moved by 2e-5.  Hyper-samples whose factor says
    # cond(Sigma) >= (max L_ii / min L_ii)^2 >= REFINE_COND therefore get classical iterative
    # refinement against the re-assembled Sigma (x += Sigma^-1 (b - Sigma x), FP64 residual):
    # it restores the accuracy of a backward-stable solve and costs nothing elsewhere.
    REFINE_COND = 1e6
    REFINE_STEPS = 2

    def ill_conditioned(self):
        """Indices of the hyper-samples that get refined solves (one small device-to-host read per
        factorisation; not on the slice sampler's log-likelihood path)."""
        self.factorize()
        ill = self.__dict__.get("_ill")
        if ill is None:
            dg = torch.diagonal(self.L, dim1=1, dim2=2)
            ratio = (dg.amax(dim=1) / dg.amin(dim=1)).cpu().numpy()
            ok = self.info == 0
            with np.errstate(invalid="ignore"):
                ill = np.nonzero(ok & (ratio * ratio >= self.REFINE_COND))[0]
            self._ill = ill
        return ill

    def _cov_one(self, s):
        """Sigma_s (noise and the factorisation's jitter on the diagonal): [n][n]."""
        eng = self.eng
        K = eng.empty(1, eng.n, eng.n)
        check(eng.lib.bqo_kernel_assemble_batched(
            C.byref(eng.desc), ptr(self.hyp[s:s + 1]), 1, ptr(self.Xs[s:s + 1]), ptr(self.tid), eng.n,
            ptr(eng.noise_obs), 1, None, ptr(K), _stream()), "bqo_kernel_assemble_batched")
        K = K[0]
        if self.jitter[s] > 0:
            K.diagonal().add_(float(self.jitter[s]))
        return K

    def _refine(self, s, rhs, x):
        """x[m][n] <- x + Sigma_s^-1 (rhs - x Sigma_s), REFINE_STEPS times (rows are vectors)."""
        eng = self.eng
        Sig = self._cov_one(s)
        m = int(x.shape[0])
        for _ in range(self.REFINE_STEPS):
            r = (rhs - torch.matmul(x, Sig)).contiguous()
            check(eng.lib.bqo_potrs_batched(ptr(self.L[s]), eng.n, 1, ptr(r), m, _stream()),
                  "bqo_potrs_batched")
            x.add_(r)

    def solve(self, Bt):
        """Sigma_s^-1 applied to Bt[S][m][n] in place."""
        eng = self.eng
        self.factorize()
        m = int(Bt.shape[1])
        ill = self.ill_conditioned()
        rhs = {int(s): Bt[s].clone() for s in ill}
        check(eng.lib.bqo_potrs_batched(ptr(self.L), eng.n, self.S, ptr(Bt), m, _stream()),
              "bqo_potrs_batched")
        for s, b0 in rhs.items():
            self._refine(s, b0, Bt[s])
        return Bt

    def refined_alpha(self):
        """alpha with the refinement applied where `ill_conditioned` says so (in place, once)."""
        self.factorize()
        if not self.__dict__.get("_alpha_refined"):
            eng = self.eng
            for s in self.ill_conditioned():
                s = int(s)
                rhs = (eng.y - self.theta[s, 1]).reshape(1, eng.n)
                self._refine(s, rhs, self.alpha[s].reshape(1, eng.n))
            self._alpha_refined = True
        return self.alpha

    def log_likelihood(self):
        eng = self.eng
        self.factorize()
        llh = eng.empty(self.S)
        check(eng.lib.bqo_loglik_batched(ptr(self.hyp), eng.hyp_stride, ptr(self.L), ptr(self.alpha),
                                         ptr(eng.y), eng.n, self.S, ptr(llh), _stream()),
              "bqo_loglik_batched")
        return llh

    def grad_log_likelihood(self):
        eng = self.eng
        self.factorize(with_inverse=True)
        nbytes = int(eng.lib.bqo_loglik_grad_workspace_bytes(eng.n, self.S))
        work = torch.empty(nbytes // 8 + 1, dtype=F64, device=eng.device)
        grad = eng.empty(self.S, 2 + eng.desc.n_params)

        def launch():
            check(eng.lib.bqo_loglik_grad_batched(
                C.byref(eng.desc), ptr(self.hyp), self.S, ptr(self.Xs), ptr(self.tid), eng.n,
                ptr(self.L), ptr(self.alpha), ptr(eng.y), ptr(self.Winv), ptr(grad), ptr(work),
                _stream()), "bqo_loglik_grad_batched")

        launch()
        if self._info is None:
            # the factorisation codes are read only now, behind the queued kernels (no pipeline
            # bubble); a matrix that needed jitter was refactored by _resolve: its gradient is
            # recomputed from the new factor (L^-1 rebuilt from L)
            self._resolve()
            if np.any(self._jitter > 0):
                launch()
        self.Winv = None  # S n^2 doubles: not kept alive in the caller's factor cache
        return grad

    # -- posterior / EI ------------------------------------------------------------------------------
    def posterior(self, P, var=True, grad=False):
        eng = self.eng
        self.factorize()
        self.refined_alpha()
        d = eng.desc
        P = eng.to_device(P).reshape(-1, d.dim)
        m = int(P.shape[0])
        mean = eng.empty(self.S, m)
        v = eng.empty(self.S, m) if var else None
        gm = eng.empty(self.S, m, d.dim_m) if grad else None
        gv = eng.empty(self.S, m, d.dim_m) if (grad and var) else None
        work = eng.empty(self.S * m * (2 + d.dim_m) * eng.n)
        check(eng.lib.bqo_posterior_batched(
            C.byref(d), ptr(self.hyp), self.S, ptr(self.Xs), ptr(self.tid), eng.n, ptr(self.L),
            ptr(self.alpha), ptr(P), m, ptr(mean), ptr(v), ptr(gm), ptr(gv), ptr(work), _stream()),
            "bqo_posterior_batched")
        return {"mean": mean, "var": v, "grad_mean": gm, "grad_var": gv}

    def ei(self, P, best, grad=False):
        eng = self.eng
        post = self.posterior(P, var=True, grad=grad)
        best = eng.to_device(best).reshape(-1)
        m = int(post["mean"].shape[1])
        ei = eng.empty(self.S, m)
        gei = eng.empty(self.S, m, eng.desc.dim_m) if grad else None
        check(eng.lib.bqo_ei_batched(ptr(post["mean"]), ptr(post["var"]), ptr(post["grad_mean"]),
                                     ptr(post["grad_var"]), ptr(best), self.S, m, eng.desc.dim_m,
                                     ptr(ei), ptr(gei), _stream()), "bqo_ei_batched")
        return ei, gei


class UnitBatch(object):
    """(candidate, hyper-sample) units of one candidate batch, ordered u = c * S + k."""

    def __init__(self):
        self.cand = self.G = self.R = self.den = self.beta4 = None
        self.unit_k = self.unit_c = None
        self.C = 0
        self.S_units = 0  # hyper-samples per candidate in the unit ordering u = c * S_units + k'


class BQEngine(object):
    """Bayesian-quadrature stage B/C on top of a factorised HyperBatch."""

    def __init__(self, hb: HyperBatch, dx: int, weights=None, wsets=None, lower=None, upper=None,
                 bq_var_noise=None, w_double=None):
        self.hb = hb.factorize()
        self.w_double = w_double  # [2][n_samples][d_w]: the two draws of the prior variance
        self.eng = eng = hb.eng
        d = eng.desc
        self.dx = int(dx)
        self.dw = d.dim_m - self.dx
        self.is_product = d.kind == _cabi.KIND_PRODUCT_TASKS
        if self.is_product and self.dw != 0:
            raise ValueError("finite-W models integrate the task column only")
        self.qt = None
        if self.is_product:
            self.weights = None if weights is None else eng.to_device(weights).reshape(-1)
            self.qt = eng.empty(hb.S, d.n_tasks)
            check(eng.lib.bqo_task_quadrature_weights(C.byref(d), ptr(hb.hyp), hb.S,
                                                      ptr(self.weights), ptr(self.qt), _stream()),
                  "bqo_task_quadrature_weights")
        self.alpha_q = eng.empty(hb.S, eng.n)
        hb.refined_alpha()
        check(eng.lib.bqo_weight_alpha(C.byref(d), ptr(hb.alpha), ptr(self.qt), ptr(hb.tid), hb.S,
                                       eng.n, ptr(self.alpha_q), _stream()), "bqo_weight_alpha")
        if self.dw > 0:
            if wsets is None:
                raise ValueError("Monte-Carlo W models need pre-drawn W sample sets")
            ws = eng.to_device(wsets)
            if ws.dim() == 2:
                ws = ws.unsqueeze(0)
            if ws.dim() != 3 or int(ws.shape[2]) != self.dw:
                raise ValueError("wsets must be [n_wsets][M][d_w]")
            self.wsets = ws.contiguous()
            self.n_wsets, self.M = int(ws.shape[0]), int(ws.shape[1])
            # the W part of every squared distance is independent of x: one table per
            # (hyper-sample, W set), read by every evaluation of the inner maximiser
            nblk = (eng.n + 31) // 32
            self.wdist = eng.empty(hb.S, self.n_wsets, nblk, self.M, 32)
            check(eng.lib.bqo_wdist_precompute(
                C.byref(d), ptr(hb.hyp), hb.S, ptr(hb.Xs), eng.n, self.dx, self.dw, ptr(self.wsets),
                self.n_wsets, self.M, ptr(self.wdist), _stream()), "bqo_wdist_precompute")
        else:
            self.wsets, self.n_wsets, self.M = None, 0, 1
            self.wdist = None
        self.lower = None if lower is None else eng.to_device(lower).reshape(-1)
        self.upper = None if upper is None else eng.to_device(upper).reshape(-1)
        # noise term of the denominator: mean observed noise if any, else var_noise_s
        if bq_var_noise is not None:
            self.add_noise = torch.full((hb.S,), float(bq_var_noise), dtype=F64, device=eng.device)
        else:
            self.add_noise = hb.theta[:, 0].contiguous()

    # -- posterior of G(x) = E_W F(x, W) ------------------------------------------------------------------
    def quadrature_rows(self, xpts, w_set=0):
        """B(x_p, X_j) for every hyper-sample: [S][P][n]
        (evaluate_quadrature_cross_cov, sbo/numerical_tools/bayesian_quadrature.py:308-334)."""
        eng, hb, d = self.eng, self.hb, self.eng.desc
        xpts = eng.to_device(xpts).reshape(-1, self.dx)
        P = int(xpts.shape[0])
        if self.is_product:
            # finite W: the task sum factorises, B = k_M(x, X_j.x) * q[task_j]; k_M comes from the
            # cross-covariance kernel with a dummy task column, divided by its task factor
            pts = torch.cat([xpts, torch.zeros(P, 1, dtype=F64, device=eng.device)], dim=1)
            KP = hb.cross_cov(pts)                                   # k_M * C[0][task_j]
            T = d.n_tasks
            C = hb.hyp[:, _cabi_taskc_offset():_cabi_taskc_offset() + T * T].reshape(hb.S, T, T)
            tid = hb.tid.long()
            c0 = C[:, 0, :][:, tid]                                  # [S][n]
            q = self.qt[:, tid]                                      # [S][n]
            return KP * (q / c0).unsqueeze(1)
        ws = self.wsets[w_set]                                       # [M][dw]
        M = int(ws.shape[0])
        pts = torch.cat([xpts.unsqueeze(1).expand(P, M, self.dx),
                         ws.unsqueeze(0).expand(P, M, self.dw)], dim=2).reshape(P * M, d.dim)
        KP = hb.cross_cov(pts.contiguous())                          # [S][P*M][n]
        return KP.reshape(hb.S, P, M, eng.n).mean(dim=2)

    def quadrature_rows_grad(self, xpts, w_set=0):
        """(B, grad_x B): [S][P][n] and [S][P][d_x][n]
        (evaluate_grad_quadrature_cross_cov, bayesian_quadrature.py:336-362)."""
        eng, hb, d = self.eng, self.hb, self.eng.desc
        xpts = eng.to_device(xpts).reshape(-1, self.dx)
        P = int(xpts.shape[0])
        if self.is_product:
            pts = torch.cat([xpts, torch.zeros(P, 1, dtype=F64, device=eng.device)], dim=1)
            KP, dKP = hb.cross_cov(pts, grad=True)                   # both carry C[0][task_j]
            T = d.n_tasks
            C = hb.hyp[:, _cabi_taskc_offset():_cabi_taskc_offset() + T * T].reshape(hb.S, T, T)
            tid = hb.tid.long()
            f = (self.qt[:, tid] / C[:, 0, :][:, tid])               # [S][n]
            return KP * f.unsqueeze(1), dKP[:, :, :self.dx, :] * f.unsqueeze(1).unsqueeze(2)
        ws = self.wsets[w_set]
        M = int(ws.shape[0])
        pts = torch.cat([xpts.unsqueeze(1).expand(P, M, self.dx),
                         ws.unsqueeze(0).expand(P, M, self.dw)], dim=2).reshape(P * M, d.dim)
        KP, dKP = hb.cross_cov(pts.contiguous(), grad=True)
        B = KP.reshape(hb.S, P, M, eng.n).mean(dim=2)
        dB = dKP.reshape(hb.S, P, M, d.dim_m, eng.n)[:, :, :, :self.dx, :].mean(dim=2)
        return B, dB

    def quadrature_posterior_grad(self, xpts, w_set=0):
        """mean [S][P], var [S][P], grad mean [S][P][d_x], grad var [S][P][d_x] of the posterior
        of G (gradient_posterior_parameters, bayesian_quadrature.py:581-643: the prior term
        E_W E_W' k is constant in x for these radial kernels)."""
        hb = self.hb
        B, dB = self.quadrature_rows_grad(xpts, w_set)
        mean = hb.theta[:, 1:2] + (B * hb.alpha.unsqueeze(1)).sum(dim=2)
        sol = hb.solve(B.clone().contiguous())
        var = self.quadrature_prior_var(xpts, w_set) - (B * sol).sum(dim=2)
        gmean = (dB * hb.alpha.unsqueeze(1).unsqueeze(2)).sum(dim=3)
        gvar = -2.0 * (dB * sol.unsqueeze(2)).sum(dim=3)
        return mean, var, gmean.contiguous(), gvar.contiguous()

    def quadrature_prior_var(self, xpts, w_set=0):
        """E_W E_W' k((x,W),(x,W')) for every hyper-sample: [S][P]
        (evaluate_quadrate_cov, bayesian_quadrature.py:282-306; exact double sum for finite W,
        the W set against itself for Monte-Carlo W)."""
        eng, hb, d = self.eng, self.hb, self.eng.desc
        xpts = eng.to_device(xpts).reshape(-1, self.dx)
        P = int(xpts.shape[0])
        if self.is_product:
            T = d.n_tasks
            C = hb.hyp[:, _cabi_taskc_offset():_cabi_taskc_offset() + T * T].reshape(hb.S, T, T)
            w = self.weights if self.weights is not None else torch.ones(T, dtype=F64, device=eng.device)
            w = w / w.sum()
            v = torch.einsum('a,sab,b->s', w, C, w)
            return v.unsqueeze(1).expand(hb.S, P).contiguous()
        # stationary kernel: k((x, w), (x, w')) only depends on w - w', so the double mean over the
        # two independent W draws (gamma_expect with double=True, sbo/lib/expectations.py:93-112;
        # `w_double`, else the W set against itself) is the same for every x: computed once per
        # engine and broadcast
        cache = self.__dict__.setdefault("_prior_var_cache", {})
        key = "double" if self.w_double is not None else int(w_set)
        v = cache.get(key)
        if v is None:
            if self.w_double is not None:
                wd = eng.to_device(self.w_double)
                w1, w2 = wd[0], wd[1]
            else:
                w1 = w2 = self.wsets[w_set]
            x0 = xpts[0:1]
            p1 = torch.cat([x0.expand(int(w1.shape[0]), self.dx), w1], dim=1).contiguous()
            p2 = torch.cat([x0.expand(int(w2.shape[0]), self.dx), w2], dim=1).contiguous()
            tmp = GPEngine(d.kind, d.dim, d.n_tasks, bool(d.same_corr), device=eng.device)
            tmp.set_data(p2, torch.zeros(int(p2.shape[0]), dtype=F64, device=eng.device))
            v = tmp.hypers(hb.theta).cross_cov(p1).mean(dim=(1, 2))
            cache[key] = v
        return v.unsqueeze(1).expand(hb.S, P).contiguous()

    def quadrature_posterior(self, xpts, w_set=0, var=True):
        """Posterior mean (and variance) of G at P points for every hyper-sample
        (BayesianQuadrature.compute_posterior_parameters, :449-537) -> mean [S][P], var [S][P]."""
        hb = self.hb
        B = self.quadrature_rows(xpts, w_set)
        mean = hb.theta[:, 1:2] + torch.einsum('spn,sn->sp', B, hb.alpha)
        if not var:
            return mean, None
        sol = hb.solve(B.clone().contiguous())
        v = self.quadrature_prior_var(xpts, w_set) - (B * sol).sum(dim=2)
        return mean, v

    # -- stage B ---------------------------------------------------------------------------------------
    def setup(self, cands) -> UnitBatch:
        eng, hb, d = self.eng, self.hb, self.eng.desc
        ub = UnitBatch()
        ub.cand = eng.to_device(cands).reshape(-1, d.dim)
        Cn = ub.C = int(ub.cand.shape[0])
        rows = 1 + d.dim_m
        ub.G = eng.empty(hb.S, Cn * rows, eng.n)
        check(eng.lib.bqo_cross_cov_batched(
            C.byref(d), ptr(hb.hyp), hb.S, ptr(hb.Xs), ptr(hb.tid), eng.n, ptr(ub.cand), Cn, 1,
            ptr(ub.G), None, _stream()), "bqo_cross_cov_batched")
        ub.R = hb.solve(ub.G.clone())
        ub.den = eng.empty(hb.S, Cn)
        ub.beta4 = eng.empty(hb.S, Cn, d.dim_m)
        check(eng.lib.bqo_candidate_setup_batched(
            C.byref(d), ptr(hb.hyp), hb.S, ptr(ub.G), ptr(ub.R), eng.n, Cn, ptr(ub.cand),
            ptr(self.add_noise), ptr(self.qt), ptr(hb.tid), ptr(ub.den), ptr(ub.beta4), _stream()),
            "bqo_candidate_setup_batched")
        u = torch.arange(Cn * hb.S, device=eng.device, dtype=torch.int64)
        ub.unit_c = (u // hb.S).to(I32).contiguous()
        ub.unit_k = (u % hb.S).to(I32).contiguous()
        ub.S_units = hb.S
        return ub

    def subset(self, ub: UnitBatch, hyper_indices) -> UnitBatch:
        """Same candidates, units restricted to the given hyper-samples (u = c * len(ks) + k')."""
        eng = self.eng
        ks = eng.to_device(np.asarray(hyper_indices, dtype=np.int32), I32).reshape(-1)
        nk = int(ks.shape[0])
        sub = UnitBatch()
        sub.cand, sub.G, sub.R, sub.den, sub.beta4, sub.C = ub.cand, ub.G, ub.R, ub.den, ub.beta4, ub.C
        u = torch.arange(ub.C * nk, device=eng.device, dtype=torch.int64)
        sub.unit_c = (u // nk).to(I32).contiguous()
        sub.unit_k = ks[(u % nk)].contiguous()
        sub.S_units = nk
        return sub

    def _stage_c(self, ub: UnitBatch) -> StageC:
        eng, hb = self.eng, self.hb
        pb = StageC()
        pb.desc = eng.desc
        pb.n, pb.S, pb.dx, pb.dw = eng.n, hb.S, self.dx, self.dw
        pb.n_units, pb.C = int(ub.unit_k.shape[0]), ub.C
        pb.M, pb.n_wsets = self.M, self.n_wsets
        pb.hyp, pb.Xs, pb.tid = ptr(hb.hyp), ptr(hb.Xs), ptr(hb.tid)
        pb.alpha, pb.qt = ptr(self.alpha_q), ptr(self.qt)
        pb.R, pb.den, pb.beta4, pb.cand = ptr(ub.R), ptr(ub.den), ptr(ub.beta4), ptr(ub.cand)
        pb.unit_k, pb.unit_c = ptr(ub.unit_k), ptr(ub.unit_c)
        pb.wsets, pb.lower, pb.upper = ptr(self.wsets), ptr(self.lower), ptr(self.upper)
        pb.wdist = ptr(self.wdist) if self.use_wdist else None
        return pb

    use_wdist = True  # class-level switch (tests compare against the on-the-fly distances)

    # -- stage C ---------------------------------------------------------------------------------------
    def sample_ab(self, ub, pts, pt_unit, pt_wset=None):
        """[P][2+2dx]: a, b, grad a, grad b at P points (unit and W set per point)."""
        eng = self.eng
        pts = eng.to_device(pts).reshape(-1, self.dx)
        P = int(pts.shape[0])
        pt_unit = eng.to_device(pt_unit, I32).reshape(-1)
        pt_wset = None if pt_wset is None else eng.to_device(pt_wset, I32).reshape(-1)
        out = eng.empty(P, 2 + 2 * self.dx)
        pb = self._stage_c(ub)
        check(eng.lib.bqo_sample_ab_batched(C.byref(pb), ptr(pts), ptr(pt_unit), ptr(pt_wset), P,
                                            ptr(out), _stream()), "bqo_sample_ab_batched")
        return out

    def sample_ab_hess(self, ub, pts, pt_unit, pt_wset=None):
        """[P][2][dx][dx]: Hessians of a and b at P points (unit and W set per point; a31)."""
        eng = self.eng
        pts = eng.to_device(pts).reshape(-1, self.dx)
        P = int(pts.shape[0])
        pt_unit = eng.to_device(pt_unit, I32).reshape(-1)
        pt_wset = None if pt_wset is None else eng.to_device(pt_wset, I32).reshape(-1)
        out = eng.empty(P, 2, self.dx, self.dx)
        pb = self._stage_c(ub)
        check(eng.lib.bqo_sample_ab_hess_batched(C.byref(pb), ptr(pts), ptr(pt_unit), ptr(pt_wset), P,
                                                 ptr(out), _stream()), "bqo_sample_ab_hess_batched")
        return out

    def inner_maximize(self, ub, z, starts, maxiter=10, factr=1e12, pgtol=1e-5, max_ls=20,
                       use_vertices=True, z_per_cta=None, count_evals=False):
        eng = self.eng
        z = eng.to_device(z).reshape(-1)
        starts = eng.to_device(starts)
        per_hyper = starts.dim() == 3
        n_starts = int(starts.shape[-2])
        n_units = int(ub.unit_k.shape[0])
        n_z = int(z.shape[0])
        if z_per_cta is None:
            z_per_cta = n_z
        opt = InnerOpt(n_z, n_starts, int(maxiter), int(max_ls), float(factr), float(pgtol),
                       int(bool(use_vertices)), int(z_per_cta), int(per_hyper))
        out_max = eng.empty(n_units, n_z)
        out_arg = eng.empty(n_units, n_z, self.dx)
        n_evals = torch.zeros(n_units, dtype=torch.int64, device=eng.device) if count_evals else None
        pb = self._stage_c(ub)
        check(eng.lib.bqo_inner_maximize_batched(C.byref(pb), C.byref(opt), ptr(z), ptr(starts),
                                                 ptr(out_max), ptr(out_arg), ptr(n_evals), _stream()),
              "bqo_inner_maximize_batched")
        return out_max, out_arg, n_evals

    def grad_vector_b(self, ub, pts, pt_wset=None):
        """pts [n_units][npts][dx] -> grad_c b [n_units][npts][dim_m], nan flags [n_units]."""
        eng = self.eng
        n_units = int(ub.unit_k.shape[0])
        pts = eng.to_device(pts).reshape(n_units, -1, self.dx)
        npts = int(pts.shape[1])
        pt_wset = None if pt_wset is None else eng.to_device(pt_wset, I32).reshape(-1)
        out = eng.empty(n_units, npts, eng.desc.dim_m)
        is_nan = torch.zeros(n_units, dtype=I32, device=eng.device)
        pb = self._stage_c(ub)
        check(eng.lib.bqo_grad_vector_b_batched(C.byref(pb), ptr(pts), npts, ptr(pt_wset), ptr(out),
                                                ptr(is_nan), _stream()), "bqo_grad_vector_b_batched")
        return out, is_nan

    def acq_reduce(self, ub, out_max, gvb, is_nan, z, max_mean=None):
        eng, hb = self.eng, self.hb
        z = eng.to_device(z).reshape(-1)
        n_z = int(z.shape[0])
        dim_g = eng.desc.dim_m if gvb is not None else 0
        value = eng.empty(ub.C)
        grad = eng.empty(ub.C, dim_g) if gvb is not None else None
        mm = None if max_mean is None else eng.to_device(max_mean).reshape(-1)
        check(eng.lib.bqo_acq_reduce(ptr(out_max), ptr(gvb), ptr(is_nan), ptr(z), ptr(mm), ub.C,
                                     ub.S_units, n_z, dim_g, ptr(value), ptr(grad), _stream()),
              "bqo_acq_reduce")
        return value, grad

    def evaluate_candidates(self, cands, z, starts, max_mean=None, compute_gradient=True,
                            hyper_indices=None, **opt):
        """One pass of the hot path over a batch of candidates: stage B, inner maximisation for
        every (candidate, hyper-sample, Z), gradient of b at the maximisers, MC reduction.
        `hyper_indices` restricts the average to a subset of the resident hyper-samples (the
        stochastic gradients of the SGD driver)."""
        ub = self.setup(cands)
        if hyper_indices is not None:
            ub = self.subset(ub, hyper_indices)
        out_max, out_arg, n_evals = self.inner_maximize(ub, z, starts, **opt)
        gvb = is_nan = None
        if compute_gradient:
            gvb, is_nan = self.grad_vector_b(ub, out_arg)
        value, grad = self.acq_reduce(ub, out_max, gvb, is_nan, z, max_mean)
        return {"evaluations": value, "gradient": grad, "max": out_max, "optimum": out_arg,
                "n_evals": n_evals, "units": ub}


def kg_discretized(eng: GPEngine, a, b):
    """KG values and envelope indices for B problems sharing `a` (discretised SBO)."""
    a = eng.to_device(a).reshape(-1)
    b = eng.to_device(b).reshape(-1, a.shape[0])
    m, B = int(a.shape[0]), int(b.shape[0])
    nbytes = int(eng.lib.bqo_kg_workspace_bytes(m, B))
    work = torch.empty(nbytes // 8 + 1, dtype=F64, device=eng.device)
    kg = eng.empty(B)
    cnt = torch.zeros(B, dtype=I32, device=eng.device)
    idx = torch.zeros(B, m, dtype=I32, device=eng.device)
    c_out = torch.zeros(B, m, dtype=F64, device=eng.device)
    check(eng.lib.bqo_kg_discretized_batched(ptr(a), ptr(b), m, B, ptr(kg), ptr(cnt), ptr(idx),
                                             ptr(c_out), ptr(work), _stream()),
          "bqo_kg_discretized_batched")
    return kg, cnt, idx, c_out
