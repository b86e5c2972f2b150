from aqml_b200.representation.core import generate_slatm, get_slatm_mbtypes  # noqa: F401
