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
