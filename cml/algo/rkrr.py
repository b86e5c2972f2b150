from aqml_b200.algo.krr import e_base, multi_fidelity_predict  # noqa: F401
