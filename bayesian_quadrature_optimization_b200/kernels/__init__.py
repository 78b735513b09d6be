"""The kernels/ API of the reference (sbo/kernels/: abstract_kernel.py:6-108 and the four kernels on
the accelerated path) with every matrix evaluated by libbqo_sm100.so."""
from .abstract_kernel import AbstractKernel  # noqa: F401
from .matern52 import Matern52, GradientLSMatern52  # noqa: F401
from .scaled_kernel import ScaledKernel  # noqa: F401
from .tasks_kernel import TasksKernel, GradientTasksKernel  # noqa: F401
from .product_kernels import ProductKernels  # noqa: F401
