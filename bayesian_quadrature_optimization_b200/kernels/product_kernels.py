"""ProductKernels: sbo/kernels/product_kernels.py for Matern52(x) * TasksKernel(task) over the
CUDA path (kind BQO_KIND_PRODUCT_TASKS)."""
import numpy as np

from . import _device as D
from .abstract_kernel import AbstractKernel
from .matern52 import Matern52
from .tasks_kernel import TasksKernel
from ..lib.constant import (PRODUCT_KERNELS_SEPARABLE, MATERN52_NAME, TASKS_KERNEL_NAME,
                            SAME_CORRELATION)


def find_kernel_constructor(name):
    if name == MATERN52_NAME:
        return Matern52
    if name == TASKS_KERNEL_NAME:
        return TasksKernel
    raise NameError(name + " doesn't exist")


class ProductKernels(AbstractKernel):

    _possible_kernels_ = [MATERN52_NAME, TASKS_KERNEL_NAME]

    def __init__(self, *kernels):
        name = PRODUCT_KERNELS_SEPARABLE + ':'
        dimension = 0
        dimension_parameters = 0
        for kernel in kernels:
            name += kernel.name + '_'
            dimension += kernel.dimension
            dimension_parameters += kernel.dimension_parameters
        super(ProductKernels, self).__init__(name, dimension, dimension_parameters)
        self.names = [kernel.name for kernel in kernels]
        if self.names != [MATERN52_NAME, TASKS_KERNEL_NAME]:
            raise NameError("only Matern52 x Tasks_Kernel products are on the accelerated path")
        self.kernels = {kernel.name: kernel for kernel in kernels}
        self.parameters = {kernel.name: kernel.hypers for kernel in kernels}

    @property
    def hypers(self):
        return self.parameters

    @property
    def hypers_as_list(self):
        out = []
        for name in self.names:
            out += self.kernels[name].hypers_as_list
        return out

    @property
    def hypers_values_as_array(self):
        return np.concatenate([self.kernels[name].hypers_values_as_array for name in self.names])

    def sample_parameters(self, number_samples, random_seed=None):
        if random_seed is not None:
            np.random.seed(random_seed)
        return np.concatenate([self.kernels[name].sample_parameters(number_samples)
                               for name in self.names], 1)

    def get_bounds_parameters(self):
        bounds = []
        for name in self.names:
            bounds += list(self.kernels[name].get_bounds_parameters())
        return bounds

    @property
    def name_parameters_as_list(self):
        out = []
        for name in self.names:
            out += self.kernels[name].name_parameters_as_list
        return out

    @classmethod
    def define_kernel_from_array(cls, dimension, params, *args, **kernel_parameters):
        """:127-146: dimension / params are lists, one entry per kernel of the product."""
        kernels = []
        for name, dim, param in zip(args[0], dimension, params):
            kernels.append(find_kernel_constructor(name).define_kernel_from_array(
                dim, param, **kernel_parameters))
        return cls(*kernels)

    @classmethod
    def define_default_kernel(cls, dimension, bounds=None, default_values=None,
                              parameters_priors=None, *args, **kernel_parameters):
        """:148-195."""
        if bounds is None:
            bounds = len(dimension) * [None]
        kernels = []
        for index, (name, dim) in enumerate(zip(args[0], dimension)):
            value = None if default_values is None else default_values[index]
            kernels.append(find_kernel_constructor(name).define_default_kernel(
                dim, bounds[index], value, parameters_priors, **kernel_parameters))
        return cls(*kernels)

    def set_parameters(self, parameters):
        for name in self.names:
            if name in parameters:
                self.kernels[name].set_parameters(**parameters[name])
                self.parameters[name] = self.kernels[name].hypers

    def update_value_parameters(self, params):
        index = 0
        for name in self.names:
            dim_param = self.kernels[name].dimension_parameters
            self.kernels[name].update_value_parameters(params[index:index + dim_param])
            index += dim_param

    # -- device evaluation --------------------------------------------------------------------------
    def _kw(self):
        tasks = self.kernels[TASKS_KERNEL_NAME]
        return dict(n_tasks=tasks.n_tasks, same_correlation=tasks.same_correlation)

    def inputs_from_array_to_dict(self, inputs):
        inputs = np.asarray(inputs, dtype=float)
        out, cont = {}, 0
        for name in self.names:
            out[name] = inputs[:, cont:cont + self.kernels[name].dimension]
            cont += self.kernels[name].dimension
        return out

    def _array(self, inputs):
        if isinstance(inputs, dict):
            return np.concatenate([np.asarray(inputs[name], dtype=float).reshape(
                -1, self.kernels[name].dimension) for name in self.names], 1)
        return np.asarray(inputs, dtype=float)

    def cov(self, inputs):
        return self.cross_cov(inputs, inputs)

    def cross_cov(self, inputs_1, inputs_2):
        """:225-270 (arrays or {kernel_name: array} dictionaries)."""
        return D.cross_cov(D.KIND_PRODUCT_TASKS, self.dimension, self.hypers_values_as_array,
                           self._array(inputs_1), self._array(inputs_2), **self._kw())

    def cov_dict(self, inputs):
        return self.cross_cov(inputs, inputs)

    def cross_cov_dict(self, inputs_1, inputs_2):
        return self.cross_cov(inputs_1, inputs_2)

    def gradient_respect_parameters(self, inputs):
        """:272-302 -> {kernel_name: {parameter_name: {i: np.array(n x n)}}}."""
        dK = D.grad_params(D.KIND_PRODUCT_TASKS, self.dimension, self.hypers_values_as_array,
                           self._array(inputs), **self._kw())
        out, index = {}, 0
        for name in self.names:
            kernel = self.kernels[name]
            pname = kernel.hypers_as_list[0].name
            n_p = kernel.dimension_parameters
            out[name] = {pname: {i: dK[index + i] for i in range(n_p)}}
            index += n_p
        return out

    def grad_respect_point_dict(self, point, inputs):
        """:415-439 -> {kernel_name: np.array(n x d)} (zero for the task coordinate)."""
        g = D.grad_point(D.KIND_PRODUCT_TASKS, self.dimension, self.hypers_values_as_array,
                         self._array(point), self._array(inputs), **self._kw())
        return {MATERN52_NAME: g, TASKS_KERNEL_NAME: np.zeros((g.shape[0], 1))}

    def grad_respect_point(self, point, inputs):
        """:304-325 -> np.array(n x d)."""
        grad = self.grad_respect_point_dict(point, inputs)
        return np.concatenate([grad[name] for name in self.names], 1)

    def hessian_respect_point(self, point, inputs):
        """:327-365 -> np.array(n x d x d): the Matern block times the task covariance; the rows
        and columns of the task coordinate vanish (its gradient and Hessian are zero)."""
        Hm = D.hessian_point(D.KIND_PRODUCT_TASKS, self.dimension, self.hypers_values_as_array,
                             self._array(point), self._array(inputs), **self._kw())
        H = np.zeros((Hm.shape[0], self.dimension, self.dimension))
        H[:, :Hm.shape[1], :Hm.shape[2]] = Hm
        return H

    @classmethod
    def evaluate_grad_respect_point(cls, params, point, inputs, dimension, *args, **kernel_parameters):
        kernel = cls.define_kernel_from_array(dimension, params, *args, **kernel_parameters)
        return kernel.grad_respect_point(point, inputs)

    @classmethod
    def evaluate_hessian_respect_point(cls, params, point, inputs, dimension, *args,
                                       **kernel_parameters):
        kernel = cls.define_kernel_from_array(dimension, params, *args, **kernel_parameters)
        return kernel.hessian_respect_point(point, inputs)

    @classmethod
    def evaluate_cov_defined_by_params(cls, params, inputs, dimension, *args, **kernel_parameters):
        kernel = cls.define_kernel_from_array(dimension, params, *args, **kernel_parameters)
        return kernel.cov_dict(inputs)

    @classmethod
    def evaluate_grad_defined_by_params_respect_params(cls, params, inputs, dimension, *args,
                                                       **kernel_parameters):
        """:506-529 -> {i: np.array(n x n)} in parameter order."""
        kernel = cls.define_kernel_from_array(dimension, params, *args, **kernel_parameters)
        gradient = kernel.gradient_respect_parameters(inputs)
        flat = {}
        for name in kernel.names:
            flat.update(gradient[name])
        return D.simple_dictionary(flat, kernel.name_parameters_as_list)

    @classmethod
    def evaluate_cross_cov_defined_by_params(cls, params, inputs_1, inputs_2, dimension, *args,
                                             **kernel_parameters):
        kernel = cls.define_kernel_from_array(dimension, params, *args, **kernel_parameters)
        return kernel.cross_cov_dict(inputs_1, inputs_2)

    @classmethod
    def evaluate_cross_cov_defined_by_params_array(cls, params, inputs_1, inputs_2, dimension,
                                                   *args, **kernel_parameters):
        kernel = cls.define_kernel_from_array(dimension, params, *args, **kernel_parameters)
        return kernel.cross_cov(inputs_1, inputs_2)

    @staticmethod
    def compare_kernels(kernel1, kernel2):
        if (kernel1.name != kernel2.name or kernel1.dimension != kernel2.dimension or
                kernel1.dimension_parameters != kernel2.dimension_parameters or
                kernel1.names != kernel2.names):
            return False
        return all(find_kernel_constructor(name).compare_kernels(kernel1.kernels[name],
                                                                 kernel2.kernels[name])
                   for name in kernel1.names)

    @staticmethod
    def parameters_from_list_to_dict(params, **kwargs):
        """:573-600: {parameter name: values} following the kernels of the product."""
        out, index = {}, 0
        for name, dim in zip(kwargs['kernels'], kwargs['dimensions']):
            ctor = find_kernel_constructor(name)
            if name == MATERN52_NAME:
                n_p = dim
            else:
                n_p = len(params) - index
            out.update(ctor.parameters_from_list_to_dict(params[index:index + n_p]))
            index += n_p
        return out
