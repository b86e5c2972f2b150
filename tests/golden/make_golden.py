"""Generate the committed golden fixtures (run in the dev container only).

Inputs come from the reference's demo fixtures (the only known-answer test in the
reference, ``demo/example/README.md:56-63``): ``demo/example/reference/g7/*.xyz`` +
``demo/example/reference/target/01.xyz``.  ``/root/reference`` does not exist on the
GPU box, so the (tiny) geometries/energies and the oracle's outputs are stored in
``tests/golden/demo.npz``; the README's printed learning curve is typed into
``README_CURVE`` below and stored alongside.

    python tests/golden/make_golden.py
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import np_oracle as o  # noqa: E402

REF = '/root/reference/demo/example/reference/'

# demo/example/README.md:57-63  (N_I, n_train, |err|, signed err, "DressedAtom mae") kcal/mol
README_CURVE = np.array([
    [1, 1, 1167.1923, 1167.1923, -1882.0703],
    [2, 3, 130.1885, 130.1885, 130.1885],
    [3, 5, 66.9106, 66.9106, 74.5169],
    [4, 6, 47.5368, 47.5368, 76.0269],
    [5, 10, 40.1071, 40.1071, 51.3559],
    [6, 11, 0.8773, -0.8773, 17.4143],
    [7, 16, 0.2251, 0.2251, 17.9935],
])


def main():
    f1 = sorted(glob.glob(REF + 'g7/*.xyz'))
    f2 = sorted(glob.glob(REF + 'target/*.xyz'))
    mols = [o.read_xyz(f) for f in f1 + f2]
    n1 = len(f1)
    zs_list = [m[0] for m in mols]
    nas = np.array([len(z) for z in zs_list])
    zs = np.concatenate(zs_list)
    coords = np.concatenate([m[1] for m in mols])
    energies = np.array([m[2] for m in mols])
    mbtypes = o.get_slatm_mbtypes(zs_list)
    kw = dict(sigmas=[.05, .05], dgrids=[.04, .04], rcut=4.8)
    x_local = np.concatenate([np.array(o.generate_slatm(m[1], m[0], mbtypes, local=True, **kw)) for m in mols])
    x_global = np.array([o.generate_slatm(m[1], m[0], mbtypes, local=False, **kw) for m in mols])
    dso = o.get_kernel_width(nas, zs, x_local)
    coeffs = np.array([0.5, 1.0, 2.0])
    sig = np.array([o.width_factor() * dso * c for c in coeffs])
    A1 = nas[:n1].sum()
    x1, x2 = x_local[:A1], x_local[A1:]
    zs1, zs2 = zs[:A1], zs[A1:]
    k1 = o.calc_lk_gaussian(x1, x1, zs1, zs1, nas[:n1], nas[:n1], 8, False, sig)
    k2 = o.calc_lk_gaussian(x2, x1, zs2, zs1, nas[n1:], nas[:n1], 8, False, sig)
    zsu = np.unique(zs)
    ngs = np.array([[(z == zz).sum() for zz in zsu] for z in zs_list])
    nheav = np.array([(z > 1).sum() for z in zs_list[:n1]])
    rows = o.learning_curve(k1[1], k2[1], ngs, energies * o.H2KC, nheav, zsu)
    curve = np.array([r[:5] for r in rows])
    mb = np.full((len(mbtypes), 3), -1, dtype=np.int64)
    for i, t in enumerate(mbtypes):
        mb[i, :len(t)] = t
    out = os.path.join(ROOT, 'tests', 'golden', 'demo.npz')
    np.savez_compressed(out, n1=n1, nas=nas, zs=zs, coords=coords, energies=energies, mbtypes=mb,
                        x_local=x_local, x_global=x_global, dso=dso, coeffs=coeffs, sigmas=sig,
                        k1=k1, k2=k2, curve=curve, readme_curve=README_CURVE)
    print('wrote', out, os.path.getsize(out), 'bytes')
    print(curve)


if __name__ == '__main__':
    main()
