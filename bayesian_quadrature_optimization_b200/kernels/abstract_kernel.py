"""AbstractKernel: the contract of sbo/kernels/abstract_kernel.py:6-108."""


class AbstractKernel(object):

    def __init__(self, name, dimension, dimension_parameters):
        self.name = name
        self.dimension = dimension
        self.dimension_parameters = dimension_parameters

    @property
    def hypers(self):
        raise NotImplementedError("Not implemented")

    @property
    def hypers_as_list(self):
        raise NotImplementedError("Not implemented")

    @property
    def hypers_values_as_array(self):
        raise NotImplementedError("Not implemented")

    @property
    def name_parameters_as_list(self):
        raise NotImplementedError("Not implemented")

    def set_parameters(self, *params):
        raise NotImplementedError("Not implemented")

    def update_value_parameters(self, params):
        raise NotImplementedError("Not implemented")

    @classmethod
    def define_kernel_from_array(cls, dimension, params):
        raise NotImplementedError("Not implemented")

    @classmethod
    def define_default_kernel(cls, dimension, bounds, default_values, parameters_priors):
        raise NotImplementedError("Not implemented")

    def cov(self, inputs):
        raise NotImplementedError("Not implemented")

    def cross_cov(self, inputs_1, inputs_2):
        raise NotImplementedError("Not implemented")

    def gradient_respect_parameters(self, inputs):
        raise NotImplementedError("Not implemented")

    def grad_respect_point(self, point, inputs):
        raise NotImplementedError("Not implemented")

    @classmethod
    def evaluate_cov_defined_by_params(cls, params, inputs, dimension):
        raise NotImplementedError("Not implemented")

    @classmethod
    def evaluate_grad_defined_by_params_respect_params(cls, params, inputs, dimension):
        raise NotImplementedError("Not implemented")

    @classmethod
    def evaluate_cross_cov_defined_by_params(cls, params, inputs_1, inputs_2, dimension):
        raise NotImplementedError("Not implemented")

    @classmethod
    def evaluate_grad_respect_point(cls, params, point, inputs, dimension):
        raise NotImplementedError("Not implemented")

    def sample_parameters(self, number_samples, random_seed):
        raise NotImplementedError("Not implemented")

    def get_bounds_parameters(self):
        raise NotImplementedError("Not implemented")

    @staticmethod
    def compare_kernels(kernel1, kernel2):
        raise NotImplementedError("Not implemented")

    @staticmethod
    def parameters_from_list_to_dict(params):
        raise NotImplementedError("Not implemented")
