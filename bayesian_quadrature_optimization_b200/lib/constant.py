"""Names and constants of the reference API that the hot path uses (the strings are part of the
wire format of serialize() and of the arguments users pass; values as in sbo/lib/constant.py)."""
import sys

# kernels: type_kernel entries of GPFittingGaussian and their per-kernel options
MATERN52_NAME, SCALED_KERNEL, TASKS_KERNEL_NAME = 'Matern52', 'Scaled_kernel', 'Tasks_Kernel'
PRODUCT_KERNELS_SEPARABLE = 'Product_of_kernels_with_separable_domain'
SAME_CORRELATION = 'same_correlation'

# parameter names (ParameterEntity.name, keys of the prior dictionaries)
LENGTH_SCALE_NAME, SIGMA2_NAME, LOWER_TRIANG_NAME = 'length_scale', 'sigma2', 'lower_triangular'
MEAN_NAME, VAR_NOISE_NAME = 'mean', 'var_noise'

# distributions of W and the keys of `parameters_distribution`
UNIFORM_FINITE, WEIGHTED_UNIFORM_FINITE = 'uniform_finite', 'weighted_uniform_finite'
GAMMA, EXPONENTIAL = 'gamma', 'exponential'
TASKS, WEIGHTS = 'n_tasks', 'weights'

# methods and optimisers
SBO_METHOD, EI_METHOD, BAYESIAN_QUADRATURE = 'sbo', 'ei', 'bayesian_quadrature'
LBFGS_NAME, SGD_NAME = 'lbfgs', 'sgd'

# numerical limits (bounds of positive parameters, clipping of the MLE gradient)
SMALLEST_POSITIVE_NUMBER = 1e-10
LARGEST_NUMBER = sys.float_info.max
SMALLEST_NUMBER = -LARGEST_NUMBER

# sample counts: hyper-samples and Z samples of bgo() (sbo/lib/constant.py), W samples of
# gamma_expect (sbo/lib/expectations.py:7)
DEFAULT_N_PARAMETERS, DEFAULT_N_SAMPLES, N_SAMPLES_W = 20, 100, 10
