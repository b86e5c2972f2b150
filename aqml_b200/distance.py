"""Host mirror of ``cml.distance`` (coreml/cml/distance.py) on top of the CUDA library."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _dist(fn, A, B):
    A, B = np.asarray(A), np.asarray(B)
    if len(A.shape) != 2 or len(B.shape) != 2:
        raise ValueError('expected matrices of dimension=2')  # distance.py:57-58
    if B.shape[1] != A.shape[1]:
        raise ValueError('expected matrices containing vectors of same size')  # distance.py:60-61
    A, B = L.f64(A), L.f64(B)
    D = np.empty((A.shape[0], B.shape[0]), order='F')
    L.call(fn, L.ptr(A), L.ptr(B), A.shape[0], B.shape[0], A.shape[1], L.ptr(D))
    return D


def l2_distance(A, B):
    """distance.py:39-68 -> fdistance.f90:25 fl2_distance."""
    return _dist("aqml_fl2_distance", A, B)


def l1_distance(A, B):
    """distance.py:6-37 -> fdistance.f90:3 fl1_distance (the upstream wrapper passes a wrong
    argument list, distance.py:35; this one implements the documented behaviour)."""
    return _dist("aqml_fl1_distance", A, B)


manhattan_distance = l1_distance  # name used by algo/aqml.py:197


def max_distance(A, B, p=2):
    """max_ij |A_i - B_j|_p without materialising the matrix (CUDA tensors)."""
    import torch
    A, B = L.dev_f64(A), L.dev_f64(B)
    out = C.c_double(0.0)
    with torch.cuda.device(A.device):
        L.call("aqml_max_distance_dev", int(p), L.ptr(A), A.shape[0], A.stride(0), L.ptr(B), B.shape[0], B.stride(0),
               A.shape[1], C.c_void_p(C.addressof(out)), L.current_stream())
    return out.value


# f2py-level name (cml.fdistance)
def fl2_distance(a_t, b_t):
    return l2_distance(np.asarray(a_t).T, np.asarray(b_t).T)
