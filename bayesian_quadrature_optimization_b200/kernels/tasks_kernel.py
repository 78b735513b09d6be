"""TasksKernel: sbo/kernels/tasks_kernel.py over the CUDA path.  The task covariance table
C = L L^T (or its same-correlation form) and its parameter derivatives are built on the device
(bqo_hyper_prepare / bqo_kernel_grad_params_batched); covariances between task ids are gathers."""
import numpy as np

from . import _device as D
from .abstract_kernel import AbstractKernel
from ..engine import _cabi_taskc_offset
from ..entities.parameter import ParameterEntity
from ..lib.constant import (TASKS_KERNEL_NAME, LOWER_TRIANG_NAME, SAME_CORRELATION,
                            SMALLEST_POSITIVE_NUMBER)
from ..priors.priors import UniformPrior, LogNormalSquare, MultivariateNormalPrior


def n_task_parameters(n_tasks, same_correlation=False):
    """sbo/lib/util.py:142-148."""
    return min(n_tasks, 2) if same_correlation else n_tasks * (n_tasks + 1) // 2


class TasksKernel(AbstractKernel):

    def __init__(self, n_tasks, lower_triang, same_correlation=False, **kernel_parameters):
        super(TasksKernel, self).__init__(TASKS_KERNEL_NAME, 1,
                                          n_task_parameters(n_tasks, same_correlation))
        self.same_correlation = same_correlation
        self.lower_triang = lower_triang
        self.n_tasks = n_tasks
        self.base_cov_matrix = None
        self.chol_base_cov_matrix = None

    # the product kernel with one dummy Matern coordinate (all points at 0, length scale 1): the
    # Matern factor is k(0) = 1, so the product is the task covariance alone
    def _hb(self, task_ids):
        ids = np.asarray(task_ids, dtype=float).reshape(-1, 1)
        pts = np.concatenate([np.zeros_like(ids), ids], axis=1)
        params = np.concatenate([[1.0], self.lower_triang.value])
        return D.hyper_batch(D.KIND_PRODUCT_TASKS, 2, params, pts, n_tasks=self.n_tasks,
                             same_correlation=self.same_correlation)

    @property
    def hypers(self):
        return {self.lower_triang.name: self.lower_triang}

    @property
    def hypers_as_list(self):
        return [self.lower_triang]

    @property
    def hypers_values_as_array(self):
        return self.lower_triang.value

    def sample_parameters(self, number_samples, random_seed=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        return self.lower_triang.sample_from_prior(number_samples)

    def get_bounds_parameters(self):
        return self.lower_triang.bounds

    @property
    def name_parameters_as_list(self):
        return [(self.lower_triang.name, [(i, None) for i in range(self.dimension_parameters)])]

    def set_parameters(self, lower_triang=None):
        self.base_cov_matrix = None
        if lower_triang is not None:
            self.lower_triang = lower_triang
            self.compute_cov_matrix()

    def update_value_parameters(self, params):
        self.base_cov_matrix = None
        self.lower_triang.set_value(params)
        self.compute_cov_matrix()

    @classmethod
    def define_kernel_from_array(cls, dimension, params, **kwargs):
        lower_triang = ParameterEntity(LOWER_TRIANG_NAME, np.asarray(params, dtype=float), None)
        return cls(dimension, lower_triang, same_correlation=kwargs.get(SAME_CORRELATION, False))

    @classmethod
    def define_default_kernel(cls, dimension, bounds=None, default_values=None,
                              parameters_priors=None, **kwargs):
        """:144-189."""
        same_correlation = kwargs.get(SAME_CORRELATION, False)
        n_params = n_task_parameters(dimension, same_correlation)
        if parameters_priors is None:
            parameters_priors = {}
        if default_values is None:
            default_values = np.array(parameters_priors.get(LOWER_TRIANG_NAME, n_params * [0.0]))
        kernel = cls.define_kernel_from_array(dimension, default_values, **kwargs)
        if np.all(np.asarray(default_values) == 0.0):
            kernel.lower_triang.prior = UniformPrior(n_params, n_params * [-1.0], n_params * [1.0])
        elif dimension == 1:
            kernel.lower_triang.prior = LogNormalSquare(1, 1.0, np.sqrt(default_values[0]))
            kernel.lower_triang.bounds = [(SMALLEST_POSITIVE_NUMBER, None)]
        else:
            kernel.lower_triang.prior = MultivariateNormalPrior(n_params, default_values,
                                                                np.eye(n_params))
        return kernel

    def compute_cov_matrix(self):
        """:198-239: C = L L^T with L_ij = exp(p), or the same-correlation form; read back from
        the hyper block the device prepares (hyper_prepare_kernel)."""
        if self.base_cov_matrix is not None:
            return
        hb = self._hb([0.0])
        T = self.n_tasks
        off = _cabi_taskc_offset()
        h = hb.hyp.cpu().numpy()[0]
        self.base_cov_matrix = h[off:off + T * T].reshape(T, T).copy()
        self.chol_base_cov_matrix = h[off + T * T:off + 2 * T * T].reshape(T, T).copy()

    def cov(self, inputs):
        return self.cross_cov(inputs, inputs)

    def cross_cov(self, inputs_1, inputs_2):
        """:241-257 -> np.array(n x m): C[s_i, t_j]."""
        self.compute_cov_matrix()
        s = np.asarray(inputs_1).reshape(-1).astype(int)
        t = np.asarray(inputs_2).reshape(-1).astype(int)
        return self.base_cov_matrix[np.ix_(s, t)]

    def gradient_respect_parameters(self, inputs):
        """:259-287 -> {'lower_triangular': {i: np.array(n x n)}}."""
        ids = np.asarray(inputs).reshape(-1).astype(int)
        dK = self._hb(ids.astype(float)).grad_cov().cpu().numpy()[0][1:]  # [0] is the dummy ls
        return {self.lower_triang.name: {i: dK[i] for i in range(self.lower_triang.dimension)}}

    def grad_respect_point(self, point, inputs):
        return np.zeros((np.asarray(inputs).shape[0], 1))

    def hessian_respect_point(self, point, inputs):
        return np.zeros((np.asarray(inputs).shape[0], 1, 1))

    @classmethod
    def evaluate_grad_respect_point(cls, params, point, inputs, dimension):
        return cls.define_kernel_from_array(dimension, params).grad_respect_point(point, inputs)

    @classmethod
    def evaluate_hessian_respect_point(cls, params, point, inputs, dimension):
        return cls.define_kernel_from_array(dimension, params).hessian_respect_point(point, inputs)

    @classmethod
    def evaluate_cov_defined_by_params(cls, params, inputs, dimension, **kwargs):
        return cls.define_kernel_from_array(dimension, params, **kwargs).cov(inputs)

    @classmethod
    def evaluate_cross_cov_defined_by_params(cls, params, inputs_1, inputs_2, dimension, **kwargs):
        return cls.define_kernel_from_array(dimension, params, **kwargs).cross_cov(inputs_1, inputs_2)

    @classmethod
    def evaluate_grad_defined_by_params_respect_params(cls, params, inputs, dimension, **kwargs):
        kernel = cls.define_kernel_from_array(dimension, params, **kwargs)
        return D.simple_dictionary(kernel.gradient_respect_parameters(inputs),
                                   kernel.name_parameters_as_list)

    @staticmethod
    def compare_kernels(kernel1, kernel2):
        return (kernel1.name == kernel2.name and kernel1.dimension == kernel2.dimension and
                kernel1.dimension_parameters == kernel2.dimension_parameters and
                kernel1.n_tasks == kernel2.n_tasks and
                kernel1.same_correlation == kernel2.same_correlation and
                bool(np.all(kernel1.lower_triang.value == kernel2.lower_triang.value)))

    @staticmethod
    def parameters_from_list_to_dict(params, **kwargs):
        return {LOWER_TRIANG_NAME: params}


class GradientTasksKernel(object):

    @staticmethod
    def gradient_respect_parameters(chol_base_cov_matrix, n_tasks, same_correlation=False):
        """:579-622 -> {param index: np.array(T x T)}: derivatives of the task covariance table.
        The log-parameters are recovered from the factor (exp'ed entries), the table comes from
        the device."""
        L = np.asarray(chol_base_cov_matrix, dtype=float)
        if same_correlation:
            if n_tasks > 1:
                off = L[0, 1]
                params = np.array([np.log(L[0, 0] - off * (n_tasks - 1)), np.log(off)])
            else:
                params = np.array([np.log(L[0, 0])])
        else:
            params = np.array([np.log(L[i, j]) for i in range(n_tasks) for j in range(i + 1)])
        kernel = TasksKernel.define_kernel_from_array(n_tasks, params,
                                                      **{SAME_CORRELATION: same_correlation})
        grad = kernel.gradient_respect_parameters(np.arange(n_tasks).reshape(-1, 1))
        return grad[LOWER_TRIANG_NAME]
