from .core import generate_slatm, generate_slatm_batch, get_slatm_mbtypes  # noqa: F401
