"""Tiny end-to-end run for compute-sanitizer (memcheck): every kernel family once, small shapes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aqml_b200 import distance, fused, kernels, synth  # noqa: E402
from aqml_b200.algo import aqml as algo  # noqa: E402
from aqml_b200.representation import core as rcore  # noqa: E402

nas, zs, coords = synth.qm9_like(12, 3)
mb = rcore.get_slatm_mbtypes(synth.split_zs(nas, zs))
kw = dict(sigmas=[.05, .05], dgrids=[.08, .08], rcut=4.8)
xl = rcore.generate_slatm_batch(coords, zs, nas, mb, local=True, **kw)
xg = rcore.generate_slatm_batch(coords, zs, nas, mb, local=False, **kw)
xo = rcore.generate_slatm_batch(coords, zs, nas, mb, local=True, sigmas=[.05, .05], dgrids=[.3, .3], rcut=4.8)  # legacy path
n1 = 8
A1 = int(nas[:n1].sum())
sig = np.full((3, 9), 12.0) * np.array([[0.5], [1.0], [2.0]])
k = kernels.get_local_kernels_gaussian(xl[A1:], xl[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], False, 9, sig)
k = kernels.get_local_kernels_gaussian(xl[A1:], xl[:A1], zs[A1:], zs[:A1], nas[n1:], nas[:n1], True, 9, sig, center=False)
k = kernels.get_local_kernels_laplacian(xl[A1:], xl[:A1], nas[n1:], nas[:n1], [300., 700.])
g = kernels.gaussian_kernel(xg[:5], xg, 80.0)
g = kernels.laplacian_kernel(xg[:5], xg, 800.0)
d = distance.l2_distance(xg, xg[:3])
w = algo.get_kernel_width(nas, zs, xl)
r1 = fused.PackedRep(coords[:A1], zs[:A1], nas[:n1], mb, sigmas=(.05, .05), dgrids=(.08, .08), rcut=4.8)
r2 = fused.PackedRep(coords[A1:], zs[A1:], nas[n1:], mb, sigmas=(.05, .05), dgrids=(.08, .08), rcut=4.8)
kf = fused.local_gaussian(r2, r1, sig, 9).cpu().numpy()
ww = fused.kernel_width(r1, r1, 9)
print("sanitize_small: ok", float(k.sum()), float(kf.sum()), xo.shape)
