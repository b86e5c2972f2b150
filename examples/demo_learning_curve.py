#!/usr/bin/env python
"""The reference's demo (`aqml -train g7/ -test target/ -p b3lypvdz`, demo/example/README.md) through the
drop-in `cml` package of this repository, every numerical stage on the GPU:

    aSLATM (generate_slatm) -> per-element dmax (get_kernel_width) -> sigma -> local Gaussian kernels
    -> dressed-atom baseline + KRR learning curve (host, as upstream)

Inputs are the demo molecules stored in tests/golden/demo.npz (geometries + B3LYP/cc-pVDZ energies of the 16
amons and the target).  Prints the same table as the reference's README.

    python examples/demo_learning_curve.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cml.algo.aqml as aq  # noqa: E402
from cml.representation.core import generate_slatm  # noqa: E402
from aqml_b200.algo import krr  # noqa: E402
from aqml_b200.representation.slatm import mbtypes_from_counts  # noqa: E402

FREE_ATOM = {1: -0.497858961664, 6: -37.830625604890, 8: -75.039064184306}  # ORCA B3LYP/cc-pVDZ, Hartree


def main():
    d = np.load(os.path.join(ROOT, "tests", "golden", "demo.npz"))
    n1, nas, zs, coords = int(d["n1"]), d["nas"], d["zs"], d["coords"]
    off = np.concatenate(([0], np.cumsum(nas)))
    mols = [(zs[off[i]:off[i + 1]], coords[off[i]:off[i + 1]]) for i in range(len(nas))]
    zsu = np.unique(zs)
    nzs = np.array([[(z == zz).sum() for zz in zsu] for z, _ in mols])
    mbtypes = mbtypes_from_counts(zsu, nzs.max(axis=0))                       # NBody.get_slatm_mbtypes
    x = np.concatenate([np.array(generate_slatm(c, z, mbtypes, local=True, sigmas=[.05, .05], dgrids=[.04, .04],
                                                rcut=4.8)) for z, c in mols])  # aqml.py:713-734
    A1 = off[n1]
    dsmax = aq.get_kernel_width(nas, zs, x)                                   # aqml.py:744-749
    fun = aq.fobj('g')
    sigmas = np.array([fun._factor * dsmax * 1.0])                            # aqml.py:775, coeff = 1
    zmax = int(zs.max())
    ks1 = fun.k(x[:A1], x[:A1], zs[:A1], zs[:A1], nas[:n1], nas[:n1], False, zmax, sigmas)   # aqml.py:788
    ks2 = fun.k(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], False, zmax, sigmas)   # aqml.py:789
    ys = d["energies"] * krr.H2KC
    nheav = np.array([(z > 1).sum() for z, _ in mols[:n1]])
    free = [krr.H2KC * FREE_ATOM[int(z)] for z in zsu]
    print("    coeff= 1.0 llambda= 0.0001")
    for ni, n, mae, rmse, emax, ys2p in krr.learning_curve(ks1[0], ks2[0], nzs, ys, nheav, free, 1e-4):
        print("  %2d %6d %12.4f" % (ni, n, mae))


if __name__ == "__main__":
    main()
