from aqml_b200.algo.krr import calc_ae_dressed, solve, learning_curve, e_base, multi_fidelity_predict, H2KC  # noqa: F401
