from aqml_b200.representation.slatm_x import MolSet, all_jobs, select_jobs, slatm  # noqa: F401
