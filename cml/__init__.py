"""Drop-in for the reference's ``cml`` package on the SLATM / kernel hot path.

``import cml.kernels``, ``cml.distance``, ``cml.representation.core`` / ``.slatm``,
``cml.fkernels`` / ``cml.fdistance`` / ``cml.representation.fslatm`` (the f2py module names,
coreml/setup.py:55-63) resolve to the B200 implementation in ``aqml_b200``.
"""
