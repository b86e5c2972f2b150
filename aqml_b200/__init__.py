"""aqml_b200 -- B200-native SLATM/aSLATM generation and kernel-matrix assembly (AQML hot path).

Public surface mirrors the reference's ``cml`` package for this path:
``aqml_b200.representation.{generate_slatm,get_slatm_mbtypes}``, ``aqml_b200.kernels.*``,
``aqml_b200.distance.*``; the top-level ``cml`` package in this repository re-exports them under
the reference's import paths.  All compute runs in libaqml_b200.so (hand-written CUDA, sm_100a);
there is no CPU fallback.
"""
__version__ = "0.1.0"
