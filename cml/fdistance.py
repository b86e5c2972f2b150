"""f2py module name of coreml/cml/fdistance.f90."""
from aqml_b200.distance import fl2_distance  # noqa: F401
