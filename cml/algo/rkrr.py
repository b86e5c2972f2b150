from aqml_b200.algo.krr import rkrr, e_base, multi_fidelity_predict  # noqa: F401
