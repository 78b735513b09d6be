"""Base class of the kernels/ package.

The reference's contract (sbo/kernels/abstract_kernel.py:6-108) is a set of method NAMES that
GPFittingGaussian, the priors and the MLE path call on any kernel.  The names live in one table
here; whatever a concrete kernel does not define raises NotImplementedError, as in the reference.
"""

# name -> how the reference declares it
CONTRACT = {
    "property": ("hypers", "hypers_as_list", "hypers_values_as_array", "name_parameters_as_list"),
    "method": ("set_parameters", "update_value_parameters", "cov", "cross_cov",
               "gradient_respect_parameters", "grad_respect_point", "sample_parameters",
               "get_bounds_parameters"),
    "classmethod": ("define_kernel_from_array", "define_default_kernel",
                    "evaluate_cov_defined_by_params",
                    "evaluate_grad_defined_by_params_respect_params",
                    "evaluate_cross_cov_defined_by_params", "evaluate_grad_respect_point"),
    "staticmethod": ("compare_kernels", "parameters_from_list_to_dict"),
}


def _missing(name):
    def stub(*args, **kwargs):
        raise NotImplementedError("Not implemented")
    stub.__name__ = name
    stub.__doc__ = "Part of the kernel contract; the concrete kernel provides it."
    return stub


class AbstractKernel(object):

    def __init__(self, name, dimension, dimension_parameters):
        self.name, self.dimension, self.dimension_parameters = name, dimension, dimension_parameters


_WRAP = {"property": property, "method": lambda f: f, "classmethod": classmethod,
         "staticmethod": staticmethod}
for _kind, _names in CONTRACT.items():
    for _name in _names:
        setattr(AbstractKernel, _name, _WRAP[_kind](_missing(_name)))
del _kind, _names, _name
