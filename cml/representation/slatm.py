from aqml_b200.representation.slatm import get_boa, get_sbop, get_sbot, NBody  # noqa: F401
