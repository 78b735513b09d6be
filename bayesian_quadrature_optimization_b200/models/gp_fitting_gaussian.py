"""GPFittingGaussian: host-side mirror of sbo/models/gp_fitting_gaussian.py over the CUDA engine.

Same method names, argument meaning, return types (numpy float64 / Python floats / dicts) and
error behaviour as the reference for the GP hot path; every number is produced by
libbqo_sm100.so on the GPU.  Methods taking (var_noise, mean, parameters_kernel) evaluate ONE
hyper-sample like the reference; the `*_batch` methods are the batched extension (all
slice-sampled hyper-samples in one launch) the acquisition functions use.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np
from scipy import linalg

from .. import _cabi
from ..engine import GPEngine, kernel_desc
from ..lib.constant import (
    MATERN52_NAME, TASKS_KERNEL_NAME, PRODUCT_KERNELS_SEPARABLE, SCALED_KERNEL, SAME_CORRELATION,
    MEAN_NAME, VAR_NOISE_NAME,
)


from ..entities.parameter import ParameterEntity  # noqa: E402,F401


class _KernelParameters(object):
    """The part of the reference's kernel objects the GP model exposes."""

    def __init__(self, values, dimension):
        self._values = np.asarray(values, dtype=float)
        self.dimension = dimension
        self.dimension_parameters = len(self._values)

    @property
    def hypers_values_as_array(self):
        return self._values

    def update_value_parameters(self, params):
        self._values = np.asarray(params, dtype=float)


def kernel_kind(type_kernel, dimensions, **kernel_parameters):
    """(kind, dim, n_tasks, same_correlation) from the reference's (type_kernel, dimensions)
    convention (sbo/models/gp_fitting_gaussian.py:88-97; sbo/services/bgo.py:155-165)."""
    same_corr = bool(kernel_parameters.get(SAME_CORRELATION, False))
    if type_kernel[0] == PRODUCT_KERNELS_SEPARABLE:
        if list(type_kernel[1:]) != [MATERN52_NAME, TASKS_KERNEL_NAME]:
            raise NameError("only Matern52 x Tasks_Kernel products are on the accelerated path")
        return _cabi.KIND_PRODUCT_TASKS, int(dimensions[1]) + 1, int(dimensions[2]), same_corr
    if type_kernel[0] == SCALED_KERNEL:
        if type_kernel[1] != MATERN52_NAME:
            raise NameError("only a scaled Matern52 is on the accelerated path")
        return _cabi.KIND_SCALED_MATERN52, int(dimensions[0]), 0, False
    if type_kernel[0] == MATERN52_NAME:
        return _cabi.KIND_MATERN52, int(dimensions[0]), 0, False
    raise NameError(type_kernel[0] + " doesn't exist")


class _ProbeBatcher(object):
    """Rendezvous of the chains of sample_parameters_chains: every running chain hands in the
    parameter vectors it needs next; when all of them wait, ONE batched log-probability call
    serves everybody (the GPU work is issued by the chain that arrived last)."""

    def __init__(self, model, n_chains):
        import threading
        self.model = model
        self.cond = threading.Condition()
        self.active = set(range(n_chains))
        self.pending = {}
        self.ready = {}

    def _flush(self):
        order = sorted(self.pending)
        vectors = [v for c in order for v in self.pending[c]]
        try:
            import torch
            # worker threads start on device 0: make the model's device current for the launch
            with torch.cuda.device(self.model._engine().device):
                values = list(self.model.log_prob_parameters_batch(vectors))
        except BaseException as exc:
            values = exc
        k = 0
        for c in order:
            m = len(self.pending[c])
            self.ready[c] = values if isinstance(values, BaseException) else values[k:k + m]
            k += m
        self.pending = {}
        self.cond.notify_all()

    def request(self, cid, vectors):
        with self.cond:
            self.pending[cid] = [np.asarray(v, dtype=float) for v in vectors]
            if set(self.pending) >= self.active:
                self._flush()
            while cid not in self.ready:
                self.cond.wait()
            out = self.ready.pop(cid)
        if isinstance(out, BaseException):
            raise out
        return out

    def done(self, cid):
        with self.cond:
            self.active.discard(cid)
            if self.pending and set(self.pending) >= self.active:
                self._flush()


class GPFittingGaussian(object):

    _possible_kernels_ = [MATERN52_NAME, TASKS_KERNEL_NAME, PRODUCT_KERNELS_SEPARABLE]

    def __init__(self, type_kernel, training_data, dimensions=None, bounds_domain=None,
                 kernel_values=None, mean_value=None, var_noise_value=None, thinning=0, n_burning=0,
                 start_point_sampler=None, max_steps_out=1, data=None, random_seed=None,
                 type_bounds=None, training_name=None, problem_name=None,
                 name_model='gp_fitting_gaussian', samples_parameters=None, noise=False,
                 simplex_domain=None, define_samplers=True, device=None, n_chains=1,
                 **kernel_parameters):
        if random_seed is not None:
            np.random.seed(random_seed)
        if bounds_domain is None:
            raise TypeError("bounds_domain is required")  # reference: len(None) at :151
        if type_bounds is None:
            type_bounds = len(bounds_domain) * [0]
        self.name_model = name_model
        self.training_name = training_name
        self.problem_name = problem_name
        self.noise = noise
        self.type_kernel = type_kernel
        self.additional_kernel_parameters = kernel_parameters
        self.simplex_domain = simplex_domain
        self.bounds = bounds_domain
        self.training_data = training_data
        self.dimension_domain = len(bounds_domain)
        self.dimensions = dimensions
        self.type_bounds = type_bounds
        self.training_data_as_array = self.convert_from_list_to_numpy(training_data)
        self.data = self.convert_from_list_to_numpy(training_data if data is None else data)
        self.thinning = thinning
        self.max_steps_out = max_steps_out
        self.n_burning = n_burning
        self.samples_parameters = []
        if samples_parameters is not None:
            self.samples_parameters = [np.array(sample) for sample in samples_parameters]
        self.start_point_sampler = start_point_sampler

        self._kind, self._dim, self._n_tasks, self._same_corr = kernel_kind(
            type_kernel, dimensions, **kernel_parameters)
        if self.data['points'].shape[1] != self._dim:
            raise ValueError("points have %d columns, kernel expects %d"
                             % (self.data['points'].shape[1], self._dim))
        self._device = device
        # > 1: sample_parameters draws from that many independent chains advanced side by side
        # (sample_parameters_chains); 1 = the reference's single chain
        self.n_chains = int(n_chains)
        self._eng = None
        self._hb_cache = OrderedDict()
        self._hb_cache_size = 64
        self._hb_cache_bytes = int(float(os.environ.get("BQO_FACTOR_CACHE_GB", "24")) * 2 ** 30)
        self.best_solution = {}

        # model values (set_parameters_kernel, :355-422, and get_values_parameters_from_data,
        # :442-490): the data-driven defaults are computed from the arguments as given; only then
        # are mean and var_noise forced to 0 for noise-free models and models with observed noise
        n_params = kernel_desc(self._kind, self._dim, self._n_tasks, self._same_corr).n_params
        if mean_value is not None and len(mean_value) == 0:
            mean_value = None
        if var_noise_value is not None and len(var_noise_value) == 0:
            var_noise_value = None
        if kernel_values is not None and len(kernel_values) == 0:
            kernel_values = None
        ev = self.training_data_as_array['evaluations']
        prior_mean = [np.mean(ev)] if mean_value is None else mean_value
        if var_noise_value is None:
            prior_var_noise = [np.var(ev)] if len(ev) > 1 else [0.1]
        else:
            prior_var_noise = var_noise_value
        if kernel_values is None:
            kernel_values = self._default_kernel_values(prior_var_noise[0])
        if len(kernel_values) != n_params:
            raise ValueError("kernel_values must have %d entries" % n_params)
        if not self.noise or self.data.get('var_noise') is not None:
            mean_value = [0.0]
            var_noise_value = [0.0]
        if mean_value is None:
            mean_value = list(prior_mean)
        if var_noise_value is None:
            var_noise_value = list(prior_var_noise)
        self.kernel = _KernelParameters(kernel_values, self._dim)
        self.mean = ParameterEntity(MEAN_NAME, np.array(mean_value, dtype=float))
        self.var_noise = ParameterEntity(VAR_NOISE_NAME, np.array(var_noise_value, dtype=float))
        self.dimension_parameters = n_params + 2
        self.kernel_values = list(kernel_values)
        self.mean_value = list(mean_value)
        self.var_noise_value = list(var_noise_value)

        if type_bounds is not None and len(type_bounds) > 0 and type_bounds[-1] == 1:
            self.separate_tasks = True
            self.tasks = self.bounds[-1]
        else:
            self.separate_tasks = False
        self.model_only_x = False
        self._define_priors()
        self.slice_samplers = []
        if define_samplers:
            self.set_samplers()

    # -- data plumbing ----------------------------------------------------------------------------
    @staticmethod
    def convert_from_list_to_numpy(data_as_list):
        """:513-535."""
        data = {}
        data['points'] = np.array(data_as_list['points'], dtype=float)
        data['evaluations'] = np.array(data_as_list['evaluations'], dtype=float)
        vn = data_as_list.get('var_noise')
        data['var_noise'] = np.array(vn, dtype=float) if vn is not None and len(vn) > 0 else None
        return data

    def _default_kernel_values(self, sigma2):
        """define_prior_parameters_using_data (lib/util_gp_fitting.py:161-240) followed by
        get_default_values_kernel (lib/util.py:162-207): length scales = mean |x_i - x_j| over all
        ordered pairs of training points / 0.324 per coordinate (kernels/matern52.py:298-326; the
        O(n^2) list becomes a sorted-order sum), sigma2 of a scaled kernel = `sigma2`
        (kernels/scaled_kernel.py:332-350), task parameters from the empirical covariance between
        tasks (_default_task_values)."""
        pts = self.training_data_as_array['points']
        dm = self._dim - 1 if self._kind == _cabi.KIND_PRODUCT_TASKS else self._dim
        ls = []
        n = pts.shape[0]
        for l in range(dm):
            if n > 1:
                x = np.sort(pts[:, l])
                k = np.arange(n)
                mean_abs = 2.0 * np.sum((2 * k - n + 1) * x) / (n * n)
            else:
                mean_abs = 0.324
            ls.append(mean_abs / 0.324)
        if self._kind == _cabi.KIND_MATERN52:
            return ls
        if self._kind == _cabi.KIND_SCALED_MATERN52:
            return ls + [sigma2]
        return ls + self._default_task_values(sigma2)

    def _default_task_values(self, sigma2):
        """TasksKernel.define_prior_parameters (kernels/tasks_kernel.py:404-521): log-Cholesky
        entries (or the two same-correlation parameters) of the empirical covariance between the
        tasks' evaluations, paired in observation order and truncated to the shorter task; zeros
        when some task has fewer than two evaluations."""
        from ..lib.constant import SMALLEST_POSITIVE_NUMBER
        T = self._n_tasks
        if T == 1:
            return [sigma2]
        data = self.training_data_as_array
        ids = data['points'][:, self._dim - 1]
        groups = [data['evaluations'][ids == t] for t in range(T)]
        n_params = min(T, 2) if self._same_corr else T * (T + 1) // 2
        if min(len(g) for g in groups) < 2:
            return n_params * [0.0]
        centred = [g - np.mean(g) for g in groups]
        cov = np.zeros((T, T))
        for i in range(T):
            for j in range(i + 1):
                d = min(len(centred[i]), len(centred[j]))
                cov[i, j] = cov[j, i] = np.sum(centred[i][:d] * centred[j][:d]) / (d - 1.0)
        if self._same_corr:
            params = [np.log(max(np.mean(np.diag(cov)), 0.1))]
            off = cov[~np.eye(T, dtype=bool)]
            off = off[off != 0]
            if np.all(off == 0):
                return params
            return params + [np.log(max(np.mean(off), 0.1))]
        low = np.zeros((T, T))
        for j in range(T):
            low[j, j] = np.sqrt(max(cov[j, j] - np.sum(low[j, :j] ** 2), SMALLEST_POSITIVE_NUMBER))
            for i in range(j + 1, T):
                low[i, j] = (cov[i, j] - np.sum(low[i, :j] * low[j, :j])) / low[j, j]
        return [np.log(max(low[i, j], 0.0001)) for i in range(T) for j in range(i + 1)]

    def _engine(self) -> GPEngine:
        if self._eng is None:
            self._eng = GPEngine(self._kind, self._dim, self._n_tasks, self._same_corr,
                                 device=self._device)
            self._eng.set_data(self.data['points'], self.data['evaluations'], self.data.get('var_noise'))
        return self._eng

    def _temp_engine(self, points):
        eng = GPEngine(self._kind, self._dim, self._n_tasks, self._same_corr, device=self._device)
        points = np.asarray(points, dtype=float)
        eng.set_data(points, np.zeros(points.shape[0]))
        return eng

    def _defaults(self, var_noise, mean, parameters_kernel):
        if var_noise is None:
            var_noise = self.var_noise.value[0]
        if parameters_kernel is None:
            parameters_kernel = self.kernel.hypers_values_as_array
        if mean is None:
            mean = self.mean.value[0]
        return float(var_noise), float(mean), np.asarray(parameters_kernel, dtype=float)

    def hyper_batch(self, thetas):
        """Device-resident batch of hyper-samples (cached by value)."""
        thetas = np.ascontiguousarray(np.asarray(thetas, dtype=float).reshape(
            -1, self.dimension_parameters))
        key = thetas.tobytes()
        hb = self._hb_cache.get(key)
        if hb is None:
            hb = self._engine().hypers(thetas)
            self._hb_cache[key] = hb
            # least-recently-used eviction by entry count and by device bytes: a factorised batch
            # holds S n^2 doubles (plus scaled points), 10 GB for 20 hyper-samples at n = 8192
            n = self._engine().n

            def nbytes(h):
                return 8 * h.S * n * (n if h.L is not None else 0) + 8 * h.S * n * (self._dim + 2)

            while len(self._hb_cache) > 1 and (
                    len(self._hb_cache) > self._hb_cache_size or
                    sum(nbytes(h) for h in self._hb_cache.values()) > self._hb_cache_bytes):
                self._hb_cache.popitem(last=False)
        else:
            self._hb_cache.move_to_end(key)
        return hb

    def _hb1(self, var_noise, mean, parameters_kernel):
        return self.hyper_batch(np.concatenate([[var_noise, mean], parameters_kernel]))

    @staticmethod
    def _raise_info(info):
        if info == -1:
            raise linalg.LinAlgError("not positive definite matrix")
        if info != 0:
            raise linalg.LinAlgError("not positive definite, even with jitter.")

    # -- model parameters -------------------------------------------------------------------------
    @property
    def get_value_parameters_model(self):
        """:645-655."""
        return np.concatenate([self.var_noise.value, self.mean.value,
                               self.kernel.hypers_values_as_array])

    def update_value_parameters(self, vector):
        """:1040-1051."""
        self.var_noise.set_value(vector[0:1])
        self.mean.set_value(vector[1:2])
        self.kernel.update_value_parameters(vector[2:])

    def add_points_evaluations(self, point, evaluation, var_noise_eval=None):
        """:492-511: append data.  The reference drops its cached factors here (TODO at :506-508);
        when exactly one point arrives, the device-resident factors of every cached hyper-sample
        batch are extended in O(n^2) instead (HyperBatch.appended), so hyper-parameters that are
        used again after the update (fixed / MLE parameters, the current slice-sampling state)
        need no new O(n^3) factorisation."""
        point = np.asarray(point, dtype=float).reshape(-1, self._dim)
        self.data['points'] = np.append(self.data['points'], point, axis=0)
        self.data['evaluations'] = np.append(self.data['evaluations'], evaluation)
        if var_noise_eval is not None:
            self.data['var_noise'] = np.append(self.data['var_noise'], var_noise_eval)
        old_cache, old_eng = self._hb_cache, self._eng
        self.clean_cache()
        self._eng = None
        if point.shape[0] == 1 and old_eng is not None and self.incremental_updates:
            # only the batches that are asked for again after the update: the model's own
            # parameter values (fixed / MLE) and the state of the slice-sampling chain, which the
            # next start_new_chain / sample_parameters call evaluates first.  Probe batches of past
            # slice-sampling sweeps are dropped; the old factors are released one by one so that
            # old and new copies of the whole cache never coexist.
            wanted = [np.ascontiguousarray(self.get_value_parameters_model.reshape(1, -1)).tobytes()]
            if len(self.samples_parameters) > 0:
                last = np.asarray(self.samples_parameters[-1], dtype=float).reshape(1, -1)
                wanted.append(np.ascontiguousarray(last).tobytes())
            keep = [(k, old_cache[k]) for k in dict.fromkeys(wanted)
                    if k in old_cache and old_cache[k].L is not None]
            old_cache.clear()
            if keep:
                eng2 = self._engine()
                while keep:
                    key, hb = keep.pop()
                    self._hb_cache[key] = hb.appended(eng2)
                    del hb

    incremental_updates = True  # class-level switch (tests compare both paths)
    batched_slice_probes = True  # slice sampler sends independent probes in one batched call

    def clean_cache(self, keep_factors=False):
        """:1628-1637.  `keep_factors=True` (used by the acquisition functions' clean_cache) keeps
        the device-resident factors: they are keyed by parameter values and always belong to the
        current data (add_points_evaluations rebuilds or extends them), so they never go stale."""
        if not keep_factors:
            self._hb_cache = OrderedDict()
        self._log_prob_memo = OrderedDict()
        self.best_solution = {}

    # -- kernel matrices ---------------------------------------------------------------------------
    def evaluate_cov(self, points, parameters_kernel):
        """:701-731 -> np.array(n x n)."""
        eng = self._temp_engine(points)
        hb = eng.hypers(np.concatenate([[0.0, 0.0], np.asarray(parameters_kernel, dtype=float)]))
        return hb.cov().cpu().numpy()[0]

    def evaluate_grad_cov(self, parameters_kernel, points):
        """:800-829 -> {i: np.array(n x n)}."""
        eng = self._temp_engine(points)
        hb = eng.hypers(np.concatenate([[0.0, 0.0], np.asarray(parameters_kernel, dtype=float)]))
        dK = hb.grad_cov().cpu().numpy()[0]
        return {i: dK[i] for i in range(dK.shape[0])}

    def evaluate_cross_cov(self, points_1, points_2, parameters_kernel):
        """:1107-1143 -> np.array(n x m)."""
        eng = self._temp_engine(points_2)
        hb = eng.hypers(np.concatenate([[0.0, 0.0], np.asarray(parameters_kernel, dtype=float)]))
        return hb.cross_cov(points_1).cpu().numpy()[0]

    def evaluate_grad_cross_cov_respect_point(self, points_1, points_2, parameters_kernel):
        """:1145-1170 -> np.array(m x k) (zero column for a task coordinate)."""
        eng = self._temp_engine(points_2)
        hb = eng.hypers(np.concatenate([[0.0, 0.0], np.asarray(parameters_kernel, dtype=float)]))
        _, dKP = hb.cross_cov(points_1, grad=True)
        g = dKP.cpu().numpy()[0, 0].T  # m x dim_m
        if g.shape[1] < self._dim:
            g = np.concatenate([g, np.zeros((g.shape[0], self._dim - g.shape[1]))], axis=1)
        return g

    def evaluate_hessian_cross_cov_respect_point(self, points_1, points_2, parameters_kernel):
        """:1174-1200 -> np.array(m x k x k): Hessians of k(points_1, points_2[i]) with respect to
        the single point points_1 (zero rows / columns for a task coordinate)."""
        eng = self._temp_engine(points_2)
        hb = eng.hypers(np.concatenate([[0.0, 0.0], np.asarray(parameters_kernel, dtype=float)]))
        H = hb.cross_cov_hessian(np.asarray(points_1, dtype=float).reshape(1, -1))
        H = H.cpu().numpy()[0, 0].transpose(2, 0, 1)  # m x dim_m x dim_m
        if H.shape[1] < self._dim:
            full = np.zeros((H.shape[0], self._dim, self._dim))
            full[:, :H.shape[1], :H.shape[2]] = H
            H = full
        return H

    # -- factorisation, likelihood ------------------------------------------------------------------
    def _chol_cov_including_noise(self, var_noise, parameters_kernel, historical_points=None,
                                  cache=True, clear_cache=True):
        """:733-770 -> (chol, cov)."""
        hb = self._hb1(var_noise, 0.0 if self.mean is None else self.mean.value[0], parameters_kernel)
        cov = hb.cov(add_noise=True).cpu().numpy()[0]
        hb.factorize(max_tries=7)
        self._raise_info(int(hb.info[0]))
        return hb.L.cpu().numpy()[0], cov

    @staticmethod
    def _factorize_for_likelihood(hb):
        """Few large matrices (the slice sampler's probes): alpha from L^-1, which extra tasks of
        the factorisation launch produce on otherwise idle SMs, beats two chains of vector solves
        (1.72 -> 1.45 ms for one n = 2048 matrix, profiles/r02_stage_a.md).  The caller drops
        hb.Winv once alpha is there."""
        small = hb.S <= 4 and hb.eng.n >= 1024 and hb.eng.n % 2 == 0
        return hb.factorize(max_tries=7, with_inverse=small)

    def log_likelihood(self, var_noise, mean, parameters_kernel):
        """:772-798 -> float."""
        hb = self._hb1(var_noise, mean, parameters_kernel)
        self._factorize_for_likelihood(hb)
        self._raise_info(int(hb.info[0]))
        hb.Winv = None
        return float(hb.log_likelihood().cpu().numpy()[0])

    def grad_log_likelihood(self, var_noise, mean, parameters_kernel):
        """:876-897 -> np.array(2 + p): [d var_noise, d mean, d kernel params]."""
        hb = self._hb1(var_noise, mean, parameters_kernel)
        hb.factorize(max_tries=7, with_inverse=True)
        self._raise_info(int(hb.info[0]))
        return hb.grad_log_likelihood().cpu().numpy()[0]

    def grad_log_likelihood_dict(self, var_noise, mean, parameters_kernel):
        g = self.grad_log_likelihood(var_noise, mean, parameters_kernel)
        return {'var_noise': g[0], 'mean': g[1], 'kernel_params': g[2:]}

    def log_likelihood_batch(self, thetas):
        """Batched extension: log-likelihood of S hyper-samples -> np.array(S); -inf where the
        factorisation failed (objective_llh convention, :946-960)."""
        hb = self.hyper_batch(thetas)
        self._factorize_for_likelihood(hb)
        out = hb.log_likelihood_host()
        hb.Winv = None
        return out

    def grad_log_likelihood_batch(self, thetas):
        hb = self.hyper_batch(thetas)
        hb.factorize(max_tries=7, with_inverse=True)
        return hb.grad_log_likelihood().cpu().numpy()

    def _cholesky_solve_vectors_for_posterior(self, var_noise, mean, parameters_kernel,
                                              historical_points=None, historical_evaluations=None,
                                              cache=True, clear_cache=True):
        """:1202-1251 -> {'chol', 'solve'}."""
        hb = self._hb1(var_noise, mean, parameters_kernel)
        hb.factorize(max_tries=7)
        self._raise_info(int(hb.info[0]))
        return {'chol': hb.L.cpu().numpy()[0], 'solve': hb.alpha.cpu().numpy()[0]}

    # -- posterior ------------------------------------------------------------------------------------
    def compute_posterior_parameters(self, points, var_noise=None, mean=None,
                                     parameters_kernel=None, only_mean=False):
        """:1253-1311 -> {'mean': np.array(n), 'cov': np.array(n x n)}."""
        var_noise, mean, parameters_kernel = self._defaults(var_noise, mean, parameters_kernel)
        hb = self._hb1(var_noise, mean, parameters_kernel)
        hb.factorize(max_tries=7)
        self._raise_info(int(hb.info[0]))
        points = np.asarray(points, dtype=float)
        if only_mean:
            post = hb.posterior(points, var=False)
            return {'mean': post['mean'].cpu().numpy()[0], 'cov': None}
        if points.shape[0] == 1:
            post = hb.posterior(points, var=True)
            return {'mean': post['mean'].cpu().numpy()[0],
                    'cov': post['var'].cpu().numpy()[0].reshape(1, 1)}
        # full posterior covariance: k(P,P) - k* Sigma^-1 k*^T
        KP = hb.cross_cov(points)
        sol = hb.solve(KP.clone())
        mean_v = hb.posterior(points, var=False)['mean'].cpu().numpy()[0]
        kpp = self.evaluate_cov(points, parameters_kernel)
        cov = kpp - (KP[0] @ sol[0].T).cpu().numpy()
        return {'mean': mean_v, 'cov': cov}

    def gradient_posterior_parameters(self, point, var_noise=None, mean=None,
                                      parameters_kernel=None, parallel=True, only_mean=False):
        """:1313-1353 -> {'mean': np.array(k), 'cov': np.array(1 x k)}."""
        var_noise, mean, parameters_kernel = self._defaults(var_noise, mean, parameters_kernel)
        hb = self._hb1(var_noise, mean, parameters_kernel)
        hb.factorize(max_tries=7)
        self._raise_info(int(hb.info[0]))
        post = hb.posterior(np.asarray(point, dtype=float), var=True, grad=True)
        gm = post['grad_mean'].cpu().numpy()[0, 0]
        gv = post['grad_var'].cpu().numpy()[0, 0]
        pad = self._dim - gm.shape[0]
        if pad > 0:
            gm = np.concatenate([gm, np.zeros(pad)])
            gv = np.concatenate([gv, np.zeros(pad)])
        return {'mean': gm, 'cov': gv.reshape(1, -1)}

    def get_historical_best_solution(self, var_noise=None, mean=None, parameters_kernel=None,
                                     noisy_evaluations=False):
        """:1590-1626."""
        var_noise, mean, parameters_kernel = self._defaults(var_noise, mean, parameters_kernel)
        index = (var_noise, mean, tuple(parameters_kernel))
        if index in self.best_solution:
            return self.best_solution[index]
        if not noisy_evaluations:
            evaluations = self.data['evaluations']
        else:
            evaluations = self.compute_posterior_parameters(
                self.data['points'], var_noise, mean, parameters_kernel, only_mean=True)['mean']
        best = np.max(evaluations)
        self.best_solution[index] = best
        return best

    # -- hyper-posterior: priors (host) + GPU likelihood, slice sampling ----------------------------
    def _define_priors(self):
        """Priors and bounds of set_parameters_kernel (:392-408) and of the kernels'
        define_default_kernel (matern52.py:124-161, scaled_kernel.py:136-175,
        tasks_kernel.py:146-189); `_kernel_priors` / `_kernel_bounds` follow the parameter order."""
        from ..priors.priors import (UniformPrior, GaussianPrior, LogNormalSquare, HorseShoePrior,
                                     NonNegativePrior, MultivariateNormalPrior, Constant)
        from ..lib.constant import SMALLEST_POSITIVE_NUMBER, SMALLEST_NUMBER, LARGEST_NUMBER
        free = (SMALLEST_NUMBER, LARGEST_NUMBER)  # entities/parameter.py:48-58
        if self.noise and self.data.get('var_noise') is None:
            self.mean.prior = GaussianPrior(1, self.mean_value[0], 1.0)
            self.var_noise.prior = NonNegativePrior(1, HorseShoePrior(1, self.var_noise_value[0]))
            self.var_noise.bounds = [(SMALLEST_POSITIVE_NUMBER, None)]
        else:
            self.mean.set_value([0.0])
            self.var_noise.set_value([0.0])
            self.mean.prior = Constant(1, 0.0)
            self.var_noise.prior = Constant(1, 0.0)
            self.var_noise.bounds = [(0.0, 0.0)]
        self.mean.bounds = [free]
        dm = self._dim - 1 if self._kind == _cabi.KIND_PRODUCT_TASKS else self._dim
        diffs = [float(self.bounds[l][1] - self.bounds[l][0]) / 0.001 for l in range(dm)]
        self._kernel_priors = [(slice(0, dm), UniformPrior(dm, dm * [SMALLEST_POSITIVE_NUMBER], diffs))]
        self._kernel_bounds = [(SMALLEST_POSITIVE_NUMBER, diff) for diff in diffs]
        kv = np.asarray(self.kernel_values, dtype=float)
        if self._kind == _cabi.KIND_SCALED_MATERN52:
            self._kernel_priors.append((slice(dm, dm + 1), LogNormalSquare(1, 1.0, np.sqrt(kv[dm]))))
            self._kernel_bounds.append((SMALLEST_POSITIVE_NUMBER, None))
        elif self._kind == _cabi.KIND_PRODUCT_TASKS:
            tv = kv[dm:]
            task_bounds = len(tv) * [free]
            if np.all(tv == 0.0):
                prior = UniformPrior(len(tv), len(tv) * [-1.0], len(tv) * [1.0])
            elif self._n_tasks == 1:
                prior = LogNormalSquare(1, 1.0, np.sqrt(tv[0]))
                task_bounds = [(SMALLEST_POSITIVE_NUMBER, None)]
            else:
                prior = MultivariateNormalPrior(len(tv), tv, np.eye(len(tv)))
            self._kernel_priors.append((slice(dm, len(kv)), prior))
            self._kernel_bounds += task_bounds
        self.length_scale_indexes = list(range(2, 2 + dm))

    def _log_prior_parameters(self, parameters):
        lp = self.var_noise.log_prior(parameters[0:1]) + self.mean.log_prior(parameters[1:2])
        kp = parameters[2:]
        for sl, prior in self._kernel_priors:
            lp += prior.logprob(kp[sl])
        return lp

    def log_prob_parameters(self, parameters):
        """:1543-1567: sum of log-priors + GPU log-likelihood.  The last few results are memoised
        by value: the slice sampler re-evaluates its current point at the start of every
        direction (slice_sampling.py:262-264), which is the point it has just accepted."""
        parameters = np.ascontiguousarray(np.asarray(parameters, dtype=float))
        key = (len(self.data['evaluations']), parameters.tobytes())
        memo = self.__dict__.setdefault('_log_prob_memo', OrderedDict())
        if key in memo:
            return memo[key]
        lp = self._log_prior_parameters(parameters)
        if not np.isinf(lp):
            try:
                lp += self.log_likelihood(parameters[0], parameters[1], parameters[2:])
            except linalg.LinAlgError:
                lp = -np.inf
        memo[key] = lp
        while len(memo) > 64:
            memo.popitem(last=False)
        return lp

    def log_prob_parameters_batch(self, vectors):
        """Batched log_prob_parameters: priors on the host, ONE batched factorisation for all
        vectors with a finite prior.  Values are bit-identical to the
        one-at-a-time path (every matrix is reduced by its own CTAs in a fixed order)."""
        vectors = [np.ascontiguousarray(np.asarray(v, dtype=float)) for v in vectors]
        memo = self.__dict__.setdefault('_log_prob_memo', OrderedDict())
        n = len(self.data['evaluations'])
        out = [None] * len(vectors)
        todo = []
        for i, v in enumerate(vectors):
            key = (n, v.tobytes())
            if key in memo:
                out[i] = memo[key]
                continue
            lp = self._log_prior_parameters(v)
            if np.isinf(lp):
                out[i] = lp
                memo[key] = lp
            else:
                out[i] = lp
                todo.append(i)
        if todo:
            llh = self.log_likelihood_batch(np.array([vectors[i] for i in todo]))
            for i, l in zip(todo, llh):
                out[i] = out[i] + l if np.isfinite(l) else -np.inf
                memo[(n, vectors[i].tobytes())] = out[i]
        while len(memo) > 64:
            memo.popitem(last=False)
        return out

    def set_samplers(self):
        """:214-269: a random-direction slice sampler on the non-length-scale parameters and a
        component-wise one on the length scales; optional burn-in."""
        from ..samplers.hyper_slice_sampler import SliceSampling
        log_prob = lambda vector, model: model.log_prob_parameters(vector)
        log_prob_batch = None
        if self.batched_slice_probes:
            log_prob_batch = lambda vectors, model: model.log_prob_parameters_batch(vectors)
        self.slice_samplers = []
        indexes = [i for i in range(self.dimension_parameters) if i not in self.length_scale_indexes]
        ignore_index = None
        if not self.noise or self.data.get('var_noise') is not None:
            ignore_index = [0, 1]
        if ignore_index is None or len(indexes) != len(ignore_index):
            self.slice_samplers.append(SliceSampling(
                log_prob, indexes, ignore_index=ignore_index,
                max_steps_out=self.max_steps_out, component_wise=False,
                log_prob_batch=log_prob_batch))
        self.slice_samplers.append(SliceSampling(
            log_prob, self.length_scale_indexes, max_steps_out=self.max_steps_out,
            component_wise=True, log_prob_batch=log_prob_batch))
        self._two_blocks = len(self.slice_samplers) == 2
        if self.start_point_sampler is not None and len(self.start_point_sampler) > 0:
            if len(self.samples_parameters) == 0:
                self.samples_parameters.append(np.array(self.start_point_sampler))
        else:
            self.samples_parameters = [self.get_value_parameters_model]
            if self.n_burning > 0:
                parameters = self.sample_parameters(float(self.n_burning) / (self.thinning + 1))
                self.samples_parameters = [parameters[-1]]
                self.start_point_sampler = parameters[-1]
            else:
                self.start_point_sampler = self.get_value_parameters_model

    def start_new_chain(self, random_seed=None):
        """:271-287."""
        if random_seed is not None:
            np.random.seed(random_seed)
        if self.n_burning > 0:
            parameters = self.sample_parameters(float(self.n_burning) / (self.thinning + 1))
        else:
            parameters = [self.samples_parameters[-1]]
        self.samples_parameters = [parameters[-1]]
        self.start_point_sampler = parameters[-1]

    def sample_parameters(self, n_samples, start_point=None, random_seed=None):
        """:289-353: n_samples * (thinning + 1) sweeps, every (thinning + 1)-th kept."""
        from ..samplers.hyper_slice_sampler import separate_vector, combine_vectors
        if random_seed is not None:
            np.random.seed(random_seed)
        if start_point is None:
            start_point = self.samples_parameters[-1]
        if not self.slice_samplers:  # define_samplers=False: the reference's loop over no samplers
            kept = [np.array(start_point, dtype=float)] * len(range(0, int(n_samples * (self.thinning + 1)), self.thinning + 1))
            self.samples_parameters += kept
            return kept
        if self.n_chains > 1 and float(n_samples).is_integer() and n_samples >= self.n_chains:
            # every chain starts at the current state, runs (thinning + 1) sweeps before its first
            # kept point and keeps ceil(n / chains) points; seeds come from the global stream so that
            # np.random.seed still fixes the run
            per = -(-int(n_samples) // self.n_chains)
            seed = int(np.random.randint(0, 2 ** 31 - 1 - self.n_chains))
            n_before = len(self.samples_parameters)
            kept = self.sample_parameters_chains(per, self.n_chains, random_seed=seed,
                                                 start_points=[start_point] * self.n_chains,
                                                 burn_in_sweeps=self.thinning + 1)[:int(n_samples)]
            self.samples_parameters = self.samples_parameters[:n_before] + kept
            return kept
        n_sweeps = int(n_samples * (self.thinning + 1))
        samples = self._slice_sweeps(start_point, n_sweeps, self.slice_samplers)
        samples_return = samples[::self.thinning + 1]
        if len(self.slice_samplers) != 1:  # the reference's one-sampler branch only returns (:319-336)
            self.samples_parameters += samples_return
        return samples_return

    def _slice_sweeps(self, start_point, n_sweeps, samplers):
        """n_sweeps sweeps of the block slice samplers from start_point -> list of points."""
        from ..samplers.hyper_slice_sampler import separate_vector, combine_vectors
        samples = []
        for _ in range(n_sweeps):
            points = separate_vector(start_point, self.length_scale_indexes)  # [others, ls]
            if len(samplers) == 2:
                blocks = [(samplers[0], 0), (samplers[1], 1)]
            else:
                blocks = [(samplers[0], 1)]
            for sampler, which in blocks:
                new_point, n_try = None, 0
                while new_point is None and n_try < 10:
                    try:
                        new_point = sampler.slice_sample(points[which], points[1 - which], self)
                    except _cabi.BqoLibraryError:  # no device / no library: not a numerical failure
                        raise
                    except Exception:
                        n_try += 1
                if new_point is None:
                    raise RuntimeError('program failed to compute a sample of the parameters')
                points[which] = new_point
            start_point = combine_vectors(points[1], points[0], self.length_scale_indexes)
            samples.append(start_point)
        return samples

    def sample_parameters_chains(self, n_samples, n_chains=4, start_points=None, random_seed=None,
                                 burn_in_sweeps=None):
        """SURVEY 8f item 1, second half: `n_chains` INDEPENDENT slice-sampling chains advanced side
        by side, their log-probability probes merged into one batched factorisation per round (a
        batch of 8 matrices costs 1.3x one matrix at n = 2048, profiles/r01_stage_a.md).

        Every chain is the reference's sequential algorithm with a numpy RandomState of its own
        (seed = random_seed + chain index), so chain c reproduces exactly the chain a single
        sampler with that RandomState would draw.  Each chain burns in for `burn_in_sweeps`
        (default: n_burning / (thinning + 1), as start_new_chain does) and then keeps every
        (thinning + 1)-th point until it has n_samples.  Returns the list of the n_chains *
        n_samples kept points (chain-major) and appends them to samples_parameters.  This is an
        extension: the reference runs one chain."""
        import copy
        import threading
        if not self.slice_samplers:
            self.set_samplers()
        n_chains = int(n_chains)
        base_seed = int(random_seed) if random_seed is not None else int(np.random.randint(0, 2 ** 31 - 1))
        if start_points is None:
            start_points = [np.array(self.samples_parameters[-1], dtype=float)] * n_chains
        if burn_in_sweeps is None:
            burn_in_sweeps = int(float(self.n_burning) / (self.thinning + 1)) if self.n_burning > 0 else 0
        n_sweeps = int(n_samples * (self.thinning + 1))
        batcher = _ProbeBatcher(self, n_chains)
        results = [None] * n_chains
        errors = []

        def run(cid):
            try:
                rng = np.random.RandomState(base_seed + cid)
                samplers = []
                for sp in self.slice_samplers:
                    sc = copy.copy(sp)
                    sc.rng = rng
                    sc.log_prob = lambda vector, model, cid=cid: batcher.request(cid, [vector])[0]
                    sc.log_prob_batch = lambda vectors, model, cid=cid: batcher.request(cid, vectors)
                    samplers.append(sc)
                point = np.array(start_points[cid], dtype=float)
                if burn_in_sweeps > 0:
                    point = self._slice_sweeps(point, burn_in_sweeps, samplers)[-1]
                results[cid] = self._slice_sweeps(point, n_sweeps, samplers)[::self.thinning + 1]
            except BaseException as exc:  # a dead chain must not leave the others waiting
                errors.append(exc)
            finally:
                batcher.done(cid)

        threads = [threading.Thread(target=run, args=(c,)) for c in range(n_chains)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        kept = [p for chain in results for p in chain]
        self.samples_parameters += kept
        return kept

    def sample_parameters_posterior(self, n_samples, random_seed=None, start_point=None):
        """:917-931 -> np.array(n_samples, n_parameters)."""
        if random_seed is not None:
            np.random.seed(random_seed)
        samples = self.sample_parameters(n_samples, start_point=start_point)
        return np.concatenate(samples).reshape(len(samples), len(samples[0]))

    # -- maximum likelihood (a16) -------------------------------------------------------------------
    @property
    def get_bounds_parameters(self):
        """:933-944: [(low, up)] for [var_noise, mean, kernel parameters]."""
        return list(self.var_noise.bounds) + list(self.mean.bounds) + list(self._kernel_bounds)

    def objective_llh(self, params):
        """:946-960: log-likelihood, or -inf when the factorisation fails."""
        try:
            return self.log_likelihood(params[0], params[1], params[2:])
        except (linalg.LinAlgError, ZeroDivisionError, ValueError):
            return -np.inf

    def grad_llh(self, params):
        """:962-973: gradient clipped to the finite range."""
        from ..lib.constant import SMALLEST_NUMBER, LARGEST_NUMBER
        return np.clip(self.grad_log_likelihood(params[0], params[1], params[2:]),
                       SMALLEST_NUMBER, LARGEST_NUMBER)

    def mle_parameters(self, start=None, random_seed=None):
        """:975-1011 -> {'solution', 'optimal_value', 'gradient', 'warnflag', 'task', 'nit',
        'funcalls'}: L-BFGS-B (scipy, as Optimization.optimize, sbo/lib/optimization.py:87-155)
        on the GPU log-likelihood and its GPU gradient."""
        from scipy.optimize import fmin_l_bfgs_b
        if random_seed is not None:
            np.random.seed(random_seed)
        if start is None:
            start = self.sample_parameters_posterior(1)[0, :]
        opt = fmin_l_bfgs_b(lambda x: -1.0 * self.objective_llh(x), np.asarray(start, dtype=float),
                            fprime=lambda x: -1.0 * self.grad_llh(x),
                            bounds=self.get_bounds_parameters)
        return {'solution': opt[0], 'optimal_value': -1.0 * opt[1], 'gradient': -1.0 * opt[2]['grad'],
                'warnflag': opt[2]['warnflag'], 'task': opt[2]['task'], 'nit': opt[2]['nit'],
                'funcalls': opt[2]['funcalls']}

    def fit_gp_regression(self, start=None, random_seed=None):
        """:1013-1038: MLE, then the model takes the optimum as its parameter values."""
        if random_seed is not None:
            np.random.seed(random_seed)
        results = self.mle_parameters(start=start)
        sol = results['solution']
        self.update_value_parameters(sol)
        self.kernel_values = list(sol[2:])
        self.mean_value = list(sol[1:2])
        self.var_noise_value = list(sol[0:1])
        return self

    # -- (de)serialisation: the reference's JSON wire format (f4) --------------------------------------
    @staticmethod
    def convert_from_numpy_to_list(data_as_np):
        """:537-556."""
        data = {'points': [list(p) for p in np.asarray(data_as_np['points'])],
                'evaluations': list(np.asarray(data_as_np['evaluations']))}
        vn = data_as_np.get('var_noise')
        data['var_noise'] = list(vn) if vn is not None else []
        return data

    def serialize(self):
        """:578-620: same keys as the reference, so existing gp_models/*.json files load."""
        samples_parameters = [list(p) for p in self.samples_parameters]
        if self.start_point_sampler is None:
            self.start_point_sampler = []
        return {
            'type_kernel': self.type_kernel,
            'training_data': self.training_data,
            'dimensions': self.dimensions,
            'kernel_values': list(self.kernel_values),
            'mean_value': list(self.mean_value),
            'var_noise_value': list(self.var_noise_value),
            'thinning': self.thinning,
            'n_burning': self.n_burning,
            'max_steps_out': self.max_steps_out,
            'data': self.convert_from_numpy_to_list(self.data),
            'bounds_domain': self.bounds if self.bounds is not None else [],
            'type_bounds': self.type_bounds,
            'name_model': self.name_model,
            'training_name': self.training_name if self.training_name is not None else '',
            'problem_name': self.problem_name if self.problem_name is not None else '',
            'same_correlation': self.additional_kernel_parameters.get(SAME_CORRELATION, False),
            'start_point_sampler': list(self.start_point_sampler),
            'samples_parameters': samples_parameters,
        }

    @classmethod
    def deserialize(cls, s, use_only_training_points=True, **extra):
        """:622-635.  `noise` is not part of the reference's wire format (its constructor default
        applies); pass it through `extra` when the model was built with noise=True."""
        kwargs = dict(s)
        kwargs.update(extra)
        model = cls(**kwargs)
        if use_only_training_points:
            model.data = model.convert_from_list_to_numpy(model.training_data)
            model._eng = None
            model.clean_cache()
        return model

    # -- maximisation of the posterior mean (a38, GP version) ---------------------------------------------
    def optimize_posterior_mean(self, start=None, random_seed=None, minimize=False, n_restarts=100,
                                n_best_restarts=10, parallel=True, n_treads=0, var_noise=None,
                                mean=None, parameters_kernel=None, n_samples_parameters=0,
                                start_new_chain=False, method_opt=None, maxepoch=10,
                                candidate_solutions=None, candidate_values=None):
        """:1373-1541 -> {'solution': np.array(d), 'optimal_value': np.array([value])}.

        All restarts climb the posterior mean together on the device: with fixed
        hyper-parameters 100 Adam steps on mu_n(x) and its gradient (the reference runs L-BFGS-B
        per start), with n_samples_parameters > 0 `maxepoch` steps on the mean over the last
        DEFAULT_N_PARAMETERS hyper-samples.  Task coordinates (type_bounds == 1) are kept fixed
        at the start point's task, as L-BFGS-B does with their [t, t] style bounds."""
        import torch
        from ..lib.batched_ascent import batched_adam_ascent
        from ..lib.constant import DEFAULT_N_PARAMETERS
        from ..services.domain import DomainService
        if candidate_solutions is not None and len(candidate_solutions) == 0:
            candidate_solutions = None
        if random_seed is not None:
            np.random.seed(random_seed)
        if start_new_chain and n_samples_parameters > 0:
            self.start_new_chain()
            self.sample_parameters(DEFAULT_N_PARAMETERS)
        if n_samples_parameters == 0:
            vn, mu, pk = self._defaults(var_noise, mean, parameters_kernel)
            thetas = np.concatenate([[vn, mu], pk]).reshape(1, -1)
        else:
            thetas = np.array(self.samples_parameters[-DEFAULT_N_PARAMETERS:])
        hb = self.hyper_batch(thetas).factorize(max_tries=7)
        for s in range(hb.S):
            self._raise_info(int(hb.info[s]))
        dim_m = hb.eng.desc.dim_m

        def averaged(p):
            post = hb.posterior(p, var=False, grad=True)
            g = post['grad_mean'].mean(dim=0)
            if g.shape[1] < p.shape[1]:  # task coordinate: no gradient
                pad = g.new_zeros(g.shape[0], p.shape[1] - g.shape[1])
                g = torch.cat([g, pad], dim=1)
            return post['mean'].mean(dim=0), g

        if start is None:
            start = np.array(DomainService.get_points_domain(n_restarts, self.bounds,
                                                             type_bounds=self.type_bounds))
            if start.shape[0] > n_best_restarts > 0:
                values = averaged(hb.eng.to_device(start))[0].cpu().numpy()
                start = start[np.argsort(values, kind='stable')[-n_best_restarts:]]
        else:
            start = np.asarray(start, dtype=float).reshape(1, -1)
        lower = np.array([b[0] for b in self.bounds], dtype=float)
        upper = np.array([b[-1] for b in self.bounds], dtype=float)
        for l, tb in enumerate(self.type_bounds):
            if tb == 1:  # finite set: the coordinate does not move
                lower[l], upper[l] = -np.inf, np.inf
        sign = -1.0 if minimize else 1.0
        ends, values = batched_adam_ascent(
            hb.eng, averaged, start, lower, upper,
            maxepoch=maxepoch if n_samples_parameters > 0 else 100, sign=sign)
        ind = int(np.argmax(sign * values))
        result = {'solution': ends[ind], 'optimal_value': np.array([values[ind]])}
        if candidate_solutions is not None:
            cand = np.array([list(c) for c in candidate_solutions], dtype=float)
            cvals = averaged(hb.eng.to_device(cand))[0].cpu().numpy()
            j = int(np.argmax(cvals))
            if cvals[j] > np.max(sign * values):
                result = {'solution': cand[j], 'optimal_value': np.array([cvals[j]])}
        if not hasattr(self, 'optimization_results'):
            self.optimization_results = []
        self.optimization_results.append(result)
        return result
