from aqml_b200.algo.aqml import fobj, get_kernel_width  # noqa: F401
