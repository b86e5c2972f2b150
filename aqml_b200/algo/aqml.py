"""Host mirror of the hot-path helpers of ``cml.algo.aqml`` (coreml/cml/algo/aqml.py:184-255)."""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from .. import distance as cmld
from .. import kernels as cmlk

T, F = True, False


class fobj(object):
    """aqml.py:184-201: distance / kernel function pair and the dmax -> sigma factor."""

    def __init__(self, kernel, local=T):
        self.kernel = kernel
        if self.kernel[0] == 'g':
            self.d = cmld.l2_distance
            self.k = cmlk.get_local_kernels_gaussian if local else cmlk.gaussian_kernels
            self._factor = 1. / np.sqrt(2.0 * np.log(2.0))
        elif self.kernel[0] == 'l':
            self.d = cmld.manhattan_distance
            self.k = cmlk.get_local_kernels_laplacian if local else cmlk.laplacian_kernels
            self._factor = 1. / np.log(2.0)
        else:
            raise Exception('#ERROR: not supported!')


def get_kernel_width(nas, zs, x, cab=F, kernel='g', itarget=F, debug=F):
    """aqml.py:204-255.  One max-distance reduction per element on the GPU (never stores the
    distance matrix); ``x`` NumPy (A, D) or CUDA tensor."""
    zs = np.asarray(zs)
    zsu = np.unique(zs)
    zmax = int(zsu[-1])
    p = 2 if kernel[0] == 'g' else 1
    import torch
    xt = x if L.is_tensor(x) else torch.as_tensor(L.f64(x), device="cuda")
    xt = L.dev_f64(xt)
    if len(zsu) == 1 or cab:
        if len(zsu) == 1:
            a = b = xt
        else:  # aqml.py:223-226: the two largest-Z elements against each other
            a = xt[torch.as_tensor(zs == zsu[-2], device=xt.device)]
            b = xt[torch.as_tensor(zs == zsu[-1], device=xt.device)]
        dmax = cmld.max_distance(a, b, p)
        dso = np.ones(zmax) * 1.e-9
        for zi in zsu:
            dso[zi - 1] = dmax
        return dso
    dso = np.zeros(zmax)
    with torch.cuda.device(xt.device):
        L.call("aqml_kernel_width_dev", p, L.ptr(xt), xt.stride(0), L.ptr(L.i32(zs)), xt.shape[0], xt.shape[1], zmax,
               L.ptr(dso), L.current_stream())
    return dso


def calc_ka(x2, x1, zs2, zs1, nas2, nas1, cab, zmax, sigmas, zaq=None, kernel='g'):
    """aqml.py:258-309: atomic kernel matrix.  Per element z the (atoms2_z x atoms1_z) Gaussian-on-distance
    block ``exp(-0.5 (d/sigma_z)^2)`` (d = L2 for 'g', L1 for 'l'), all sigmas from one device pass; with
    ``zaq=None`` the training-atom axis is summed per training molecule -> (nsigmas, natoms2, nmol1)."""
    import torch
    if cab:
        raise Exception('??')  # aqml.py:277-278
    sig = np.asarray(sigmas, dtype=np.float64)
    nsg = len(sig)
    zs1 = np.asarray(zs1)
    zs2 = np.asarray(zs2)
    dev = x1.device if L.is_tensor(x1) else torch.device("cuda")
    x1t = x1 if L.is_tensor(x1) else torch.as_tensor(L.f64(x1), device=dev)
    x2t = x2 if L.is_tensor(x2) else torch.as_tensor(L.f64(x2), device=dev)
    na1, na2, n1 = len(zs1), len(zs2), len(nas1)
    if zaq is None and kernel[0] == 'g':
        # every test atom as a one-atom "molecule": the element-matched local kernel IS the atomic kernel summed over the
        # atoms of each training molecule -- the pair kernel's segmented epilogue does the sum, nothing is scattered
        ks = cmlk.get_local_kernels_gaussian(x2t, x1t, zs2, zs1, np.ones(na2, dtype=np.int32), nas1, False, zmax, sig)
        return ks if L.is_tensor(x1) else ks.cpu().numpy()
    zsu = np.unique(zs2) if zaq is None else [zaq]
    blocks = {}
    for zi in zsu:
        i1 = np.nonzero(zs1 == zi)[0]
        i2 = np.nonzero(zs2 == zi)[0]
        if len(i1) == 0 or len(i2) == 0:
            blocks[zi] = (i1, i2, torch.zeros((nsg, len(i2), len(i1)), dtype=torch.float64, device=dev))
            continue
        a = x2t[torch.as_tensor(i2, device=dev)]
        b = x1t[torch.as_tensor(i1, device=dev)]
        s_z = sig[:, zi - 1]
        if kernel[0] == 'g':
            k = cmlk.global_kernels(a, b, s_z, "gaussian")
        else:  # exp(-0.5 (d1/sigma)^2) on the Manhattan distance: exp(-d1/s)-kernel with log -> square
            d1 = -torch.log(cmlk.global_kernels(a, b, [1.0], "laplacian")[0])
            k = torch.stack([torch.exp(-0.5 * (d1 / s) ** 2) for s in s_z])
        blocks[zi] = (i1, i2, k)
    if zaq is not None:
        out = blocks[zaq][2]
        return out if L.is_tensor(x1) else out.cpu().numpy()
    # sum over the atoms of each training molecule, element block by element block (the (nsigmas, A2, A1) matrix of the
    # reference, aqml.py:296-305, is never formed)
    mol1 = np.repeat(np.arange(n1), np.asarray(nas1))
    ks = torch.zeros((nsg, na2, n1), dtype=torch.float64, device=dev)
    for zi, (i1, i2, k) in blocks.items():
        if len(i1) and len(i2):
            part = torch.zeros((nsg, len(i2), n1), dtype=torch.float64, device=dev)
            part.index_add_(2, torch.as_tensor(mol1[i1], device=dev), k)
            ks[:, torch.as_tensor(i2, device=dev)] = part
    return ks if L.is_tensor(x1) else ks.cpu().numpy()
