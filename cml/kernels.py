from aqml_b200.kernels import *  # noqa: F401,F403
from aqml_b200.kernels import (gaussian_kernel, laplacian_kernel, get_local_kernels_gaussian,  # noqa: F401
                               get_local_kernels_laplacian, gaussian_kernels, laplacian_kernels,
                               linear_kernel, sargan_kernel, matern_kernel)
