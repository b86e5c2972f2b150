from aqml_b200.algo.aqml import fobj, get_kernel_width, calc_ka  # noqa: F401
