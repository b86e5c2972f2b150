"""f2py module name of coreml/cml/fkernels.f90."""
from aqml_b200.kernels import (fgaussian_kernel, flaplacian_kernel, calc_lk_gaussian, calc_lk_laplacian,  # noqa: F401
                               flinear_kernel, fsargan_kernel, fmatern_kernel_l2,
                               calc_vector_kernels_gaussian, calc_vector_kernels_laplacian)
