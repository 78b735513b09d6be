"""Public names and constants of the reference API (sbo/lib/constant.py) used on the hot path."""
import sys

MATERN52_NAME = 'Matern52'
TASKS_KERNEL_NAME = 'Tasks_Kernel'
SAME_CORRELATION = 'same_correlation'
PRODUCT_KERNELS_SEPARABLE = 'Product_of_kernels_with_separable_domain'
SCALED_KERNEL = 'Scaled_kernel'

LENGTH_SCALE_NAME = 'length_scale'
SIGMA2_NAME = 'sigma2'
LOWER_TRIANG_NAME = 'lower_triangular'
MEAN_NAME = 'mean'
VAR_NOISE_NAME = 'var_noise'

LBFGS_NAME = 'lbfgs'
SGD_NAME = 'sgd'

SMALLEST_POSITIVE_NUMBER = 1e-10
SMALLEST_NUMBER = -sys.float_info.max
LARGEST_NUMBER = sys.float_info.max

UNIFORM_FINITE = 'uniform_finite'
TASKS = 'n_tasks'
GAMMA = 'gamma'
EXPONENTIAL = 'exponential'
WEIGHTED_UNIFORM_FINITE = 'weighted_uniform_finite'
WEIGHTS = 'weights'

SBO_METHOD = 'sbo'
EI_METHOD = 'ei'
BAYESIAN_QUADRATURE = 'bayesian_quadrature'

DEFAULT_N_PARAMETERS = 20
DEFAULT_N_SAMPLES = 100
N_SAMPLES_W = 10  # sbo/lib/expectations.py:7
