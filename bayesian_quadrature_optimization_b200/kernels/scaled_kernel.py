"""ScaledKernel: sbo/kernels/scaled_kernel.py (sigma2 * Matern52) over the CUDA path."""
import numpy as np

from . import _device as D
from .abstract_kernel import AbstractKernel
from .matern52 import Matern52
from ..entities.parameter import ParameterEntity
from ..lib.constant import SCALED_KERNEL, SIGMA2_NAME, MATERN52_NAME, SMALLEST_POSITIVE_NUMBER
from ..priors.priors import LogNormalSquare


def find_kernel_constructor(name):
    if name == MATERN52_NAME:
        return Matern52
    raise NameError("only a scaled Matern52 is on the accelerated path")


class ScaledKernel(AbstractKernel):

    def __init__(self, dimension, kernel, sigma2):
        super(ScaledKernel, self).__init__(SCALED_KERNEL, dimension, kernel.dimension_parameters + 1)
        self.kernel = kernel
        self.sigma2 = sigma2

    @property
    def hypers(self):
        hypers = dict(self.kernel.hypers)
        hypers[self.sigma2.name] = self.sigma2
        return hypers

    @property
    def hypers_as_list(self):
        return self.kernel.hypers_as_list + [self.sigma2]

    @property
    def hypers_values_as_array(self):
        return np.concatenate([self.kernel.hypers_values_as_array, self.sigma2.value])

    def sample_parameters(self, number_samples, random_seed=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        samples = [p.sample_from_prior(number_samples) for p in self.hypers_as_list]
        return np.concatenate(samples, 1)

    def get_bounds_parameters(self):
        bounds = []
        for parameter in self.hypers_as_list:
            bounds += list(parameter.bounds)
        return bounds

    @property
    def name_parameters_as_list(self):
        return self.kernel.name_parameters_as_list + [(self.sigma2.name, None)]

    def set_parameters(self, kernel_parameters=None, sigma2=None):
        if kernel_parameters is not None:
            self.kernel.set_parameters(*kernel_parameters)
        if sigma2 is not None:
            self.sigma2 = sigma2

    def update_value_parameters(self, params):
        self.kernel.update_value_parameters(params[0:-1])
        self.sigma2.set_value(params[-1:])

    @classmethod
    def define_kernel_from_array(cls, dimension, params, *args):
        params = np.asarray(params, dtype=float)
        kernel = None
        for name in args[0]:
            kernel = find_kernel_constructor(name).define_kernel_from_array(dimension, params[0:-1])
        return cls(dimension, kernel, ParameterEntity(SIGMA2_NAME, params[-1:], None))

    @classmethod
    def define_default_kernel(cls, dimension, bounds=None, default_values=None,
                              parameters_priors=None, *args):
        """:136-175."""
        if parameters_priors is None:
            parameters_priors = {}
        default_values_kernel = None
        if default_values is not None:
            default_values_kernel = default_values[0:-1]
            default_value_sigma = default_values[-1:]
        else:
            default_value_sigma = [parameters_priors.get(SIGMA2_NAME, 1.0)]
        kernel = None
        for name in args[0]:
            kernel = find_kernel_constructor(name).define_default_kernel(
                dimension, bounds, default_values_kernel, parameters_priors)
        sigma2 = ParameterEntity(SIGMA2_NAME, default_value_sigma,
                                 LogNormalSquare(1, 1.0, np.sqrt(default_value_sigma)),
                                 bounds=[(SMALLEST_POSITIVE_NUMBER, None)])
        return cls(dimension, kernel, sigma2)

    def _params(self):
        return self.hypers_values_as_array

    def cov(self, inputs):
        return self.cross_cov(inputs, inputs)

    def cross_cov(self, inputs_1, inputs_2):
        """:186-194."""
        return D.cross_cov(D.KIND_SCALED_MATERN52, self.dimension, self._params(), inputs_1, inputs_2)

    def gradient_respect_parameters(self, inputs):
        """:196-214 -> {'length_scale': {i: n x n}, 'sigma2': n x n}."""
        dK = D.grad_params(D.KIND_SCALED_MATERN52, self.dimension, self._params(), inputs)
        d = self.dimension
        grad = {self.kernel.length_scale.name: {i: dK[i] for i in range(d)}}
        grad[self.sigma2.name] = dK[d]
        return grad

    def grad_respect_point(self, point, inputs):
        return D.grad_point(D.KIND_SCALED_MATERN52, self.dimension, self._params(), point, inputs)

    def hessian_respect_point(self, point, inputs):
        return D.hessian_point(D.KIND_SCALED_MATERN52, self.dimension, self._params(), point, inputs)

    @classmethod
    def evaluate_grad_respect_point(cls, params, point, inputs, dimension, *args):
        return cls.define_kernel_from_array(dimension, params, *args).grad_respect_point(point, inputs)

    @classmethod
    def evaluate_hessian_respect_point(cls, params, point, inputs, dimension, *args):
        return cls.define_kernel_from_array(dimension, params, *args).hessian_respect_point(point, inputs)

    @classmethod
    def evaluate_cov_defined_by_params(cls, params, inputs, dimension, *args):
        return cls.define_kernel_from_array(dimension, params, *args).cov(inputs)

    @classmethod
    def evaluate_grad_defined_by_params_respect_params(cls, params, inputs, dimension, *args):
        kernel = cls.define_kernel_from_array(dimension, params, *args)
        return D.simple_dictionary(kernel.gradient_respect_parameters(inputs),
                                   kernel.name_parameters_as_list)

    @classmethod
    def evaluate_cross_cov_defined_by_params(cls, params, inputs_1, inputs_2, dimension, *args):
        return cls.define_kernel_from_array(dimension, params, *args).cross_cov(inputs_1, inputs_2)

    @staticmethod
    def define_prior_parameters(data, dimension, var_evaluations=None):
        if var_evaluations is None:
            var_evaluations = np.var(data['evaluations'])
        return {SIGMA2_NAME: var_evaluations}

    @staticmethod
    def compare_kernels(kernel1, kernel2):
        return (kernel1.name == kernel2.name and kernel1.dimension == kernel2.dimension and
                kernel1.dimension_parameters == kernel2.dimension_parameters and
                bool(np.all(kernel1.sigma2.value == kernel2.sigma2.value)) and
                Matern52.compare_kernels(kernel1.kernel, kernel2.kernel))

    @staticmethod
    def parameters_from_list_to_dict(params, **kwargs):
        out = find_kernel_constructor(kwargs['kernels']).parameters_from_list_to_dict(params[0:-1])
        out[SIGMA2_NAME] = params[-1]
        return out
