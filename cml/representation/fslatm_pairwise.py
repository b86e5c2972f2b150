"""Import path of coreml/cml/representation/fslatm_pairwise.f90 (the reference's xb.py / slatm_aux.py / lo/dmx.py import it as
`cml.fslatm`): atom + bond aSLATM and the molecular kernels built on it, on the GPU (aqml_b200/representation/pairwise.py)."""
from aqml_b200.representation.pairwise import (calc_all_local_spectrum, calc_dij_max, calc_local_spectrum, calc_mk1, calc_mk2,  # noqa: F401
                                                get_neighbors, get_slatm_mbtypes, grid_sizes)
