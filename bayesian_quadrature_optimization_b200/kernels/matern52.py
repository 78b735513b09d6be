"""Matern52: sbo/kernels/matern52.py over the CUDA path (same class surface; every matrix comes
from bqo_cross_cov_batched / bqo_kernel_grad_params_batched / bqo_cross_cov_hessian_batched)."""
import numpy as np

from . import _device as D
from .abstract_kernel import AbstractKernel
from ..entities.parameter import ParameterEntity
from ..lib.constant import (MATERN52_NAME, LENGTH_SCALE_NAME, SMALLEST_POSITIVE_NUMBER,
                            LARGEST_NUMBER)
from ..priors.priors import UniformPrior


class Matern52(AbstractKernel):

    def __init__(self, dimension, length_scale, **kernel_parameters):
        super(Matern52, self).__init__(MATERN52_NAME, dimension, dimension)
        self.length_scale = length_scale

    @property
    def hypers(self):
        return {self.length_scale.name: self.length_scale}

    @property
    def hypers_as_list(self):
        return [self.length_scale]

    @property
    def hypers_values_as_array(self):
        return np.concatenate([self.length_scale.value])

    def sample_parameters(self, number_samples, random_seed=None):
        """:62-75 -> np.array(number_samples x dimension)."""
        if random_seed is not None:
            np.random.seed(random_seed)
        samples = [p.sample_from_prior(number_samples) for p in self.hypers_as_list]
        return np.concatenate(samples, 1)

    def get_bounds_parameters(self):
        return list(self.length_scale.bounds)

    @property
    def name_parameters_as_list(self):
        return [(self.length_scale.name, [(i, None) for i in range(self.dimension)])]

    def set_parameters(self, length_scale=None):
        if length_scale is not None:
            self.length_scale = length_scale

    def update_value_parameters(self, params):
        self.length_scale.set_value(params[0:self.dimension])

    @classmethod
    def define_kernel_from_array(cls, dimension, params, **kernel_parameters):
        return cls(dimension, ParameterEntity(LENGTH_SCALE_NAME, np.asarray(params, dtype=float)[0:dimension], None))

    @classmethod
    def define_default_kernel(cls, dimension, bounds=None, default_values=None,
                              parameters_priors=None, **kernel_parameters):
        """:125-161."""
        if parameters_priors is None:
            parameters_priors = {}
        if default_values is None:
            default_values = parameters_priors.get(LENGTH_SCALE_NAME, dimension * [1.0])
        kernel = cls.define_kernel_from_array(dimension, default_values)
        if bounds is not None:
            diffs = [float(bound[1] - bound[0]) / 0.001 for bound in bounds]
        else:
            diffs = parameters_priors.get(LENGTH_SCALE_NAME, dimension * [LARGEST_NUMBER])
        kernel.length_scale.prior = UniformPrior(dimension, dimension * [SMALLEST_POSITIVE_NUMBER], diffs)
        kernel.length_scale.bounds = [(SMALLEST_POSITIVE_NUMBER, diff) for diff in diffs]
        return kernel

    def _params(self):
        return self.length_scale.value

    def cov(self, inputs):
        return self.cross_cov(inputs, inputs)

    def cross_cov(self, inputs_1, inputs_2):
        """:173-183 -> np.array(n x m)."""
        return D.cross_cov(D.KIND_MATERN52, self.dimension, self._params(), inputs_1, inputs_2)

    def gradient_respect_parameters(self, inputs):
        """:185-195 -> {'length_scale': {i: np.array(n x n)}}."""
        dK = D.grad_params(D.KIND_MATERN52, self.dimension, self._params(), inputs)
        return {self.length_scale.name: {i: dK[i] for i in range(self.dimension)}}

    def grad_respect_point(self, point, inputs):
        """:197-207 -> np.array(n x d)."""
        return D.grad_point(D.KIND_MATERN52, self.dimension, self._params(), point, inputs)

    def hessian_respect_point(self, point, inputs):
        """:209-218 -> np.array(n x d x d)."""
        return D.hessian_point(D.KIND_MATERN52, self.dimension, self._params(), point, inputs)

    @classmethod
    def evaluate_grad_respect_point(cls, params, point, inputs, dimension):
        return cls.define_kernel_from_array(dimension, params).grad_respect_point(point, inputs)

    @classmethod
    def evaluate_hessian_respect_point(cls, params, point, inputs, dimension):
        return cls.define_kernel_from_array(dimension, params).hessian_respect_point(point, inputs)

    @classmethod
    def evaluate_cov_defined_by_params(cls, params, inputs, dimension, **kwargs):
        return cls.define_kernel_from_array(dimension, params).cov(inputs)

    @classmethod
    def evaluate_grad_defined_by_params_respect_params(cls, params, inputs, dimension, **kwargs):
        kernel = cls.define_kernel_from_array(dimension, params)
        return D.simple_dictionary(kernel.gradient_respect_parameters(inputs),
                                   kernel.name_parameters_as_list)

    @classmethod
    def evaluate_cross_cov_defined_by_params(cls, params, inputs_1, inputs_2, dimension, **kwargs):
        return cls.define_kernel_from_array(dimension, params).cross_cov(inputs_1, inputs_2)

    @staticmethod
    def define_prior_parameters(data, dimension):
        """:297-326: mean |x_i - x_j| over all ordered pairs / 0.324 per coordinate."""
        pts = np.asarray(data['points'], dtype=float)
        n = pts.shape[0]
        out = []
        for l in range(dimension):
            if n > 1:
                x = np.sort(pts[:, l])
                k = np.arange(n)
                mean_abs = 2.0 * np.sum((2 * k - n + 1) * x) / (n * n)
            else:
                mean_abs = 0.324
            out.append(mean_abs / 0.324)
        return {LENGTH_SCALE_NAME: out}

    @staticmethod
    def compare_kernels(kernel1, kernel2):
        return (kernel1.name == kernel2.name and kernel1.dimension == kernel2.dimension and
                kernel1.dimension_parameters == kernel2.dimension_parameters and
                bool(np.all(kernel1.length_scale.value == kernel2.length_scale.value)))

    @staticmethod
    def parameters_from_list_to_dict(params, **kwargs):
        return {LENGTH_SCALE_NAME: params}


class GradientLSMatern52(object):
    """The static helpers of sbo/kernels/matern52.py:371-503 (ls: ParameterEntity)."""

    @classmethod
    def gradient_respect_parameters_ls(cls, inputs, ls):
        dim = len(ls.value)
        dK = D.grad_params(D.KIND_MATERN52, dim, ls.value, inputs)
        return {ls.name: {i: dK[i] for i in range(dim)}}

    @classmethod
    def grad_respect_point(cls, ls, point, inputs):
        return D.grad_point(D.KIND_MATERN52, len(ls.value), ls.value, point, inputs)

    @classmethod
    def hessian_respect_point(cls, ls, point, inputs):
        return D.hessian_point(D.KIND_MATERN52, len(ls.value), ls.value, point, inputs)
