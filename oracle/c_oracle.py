"""ctypes front-end of oracle/c_oracle.c (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

``load(fast=False)`` returns the checker build (no FMA contraction, portable ISA);
``load(fast=True)`` the ``-march=native`` build used as the timed CPU baseline
(built on the machine that runs it).  Never imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def build(fast=False):
    target = 'fast' if fast else 'all'
    subprocess.check_call(['make', '-s', '-C', HERE, target], stdout=subprocess.DEVNULL)


def load(fast=False):
    if fast in _libs:
        return _libs[fast]
    name = 'libc_oracle_fast.so' if fast else 'libc_oracle.so'
    path = os.path.join(HERE, name)
    src = os.path.join(HERE, 'c_oracle.c')
    if fast or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        try:
            if fast and os.path.exists(path):
                os.remove(path)  # -march=native must match *this* host
            build(fast)
        except Exception:
            if not os.path.exists(path):
                if fast:
                    return load(False)
                raise
    lib = C.CDLL(path)
    _libs[fast] = lib
    return lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads(fast=False):
    return int(load(fast).oracle_num_threads())


def set_num_threads(n, fast=False):
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use every host core."""
    load(fast).oracle_set_num_threads(C.c_int(int(n)))


def mbtypes_array(mbtypes):
    mb = np.full((len(mbtypes), 3), -1, dtype=np.int32)
    for i, t in enumerate(mbtypes):
        mb[i, :len(t)] = t
    return mb


def slatm_dim(mbtypes, nx2, nx3):
    return sum(1 if len(t) == 1 else (nx2 if len(t) == 2 else nx3) for t in mbtypes)


def generate_slatm_batch(coords, zs, nas, mbtypes, local, sigmas=(0.05, 0.05), dgrids=(0.03, 0.03),
                         rcut=4.8, rpower=6.0, izeff=False, fast=False):
    """All molecules at once; returns (A,D) if local else (nmol,D)."""
    from . import np_oracle as o
    lib = load(fast)
    coords = _d(coords)
    zs = _i(zs)
    off = _i(np.concatenate(([0], np.cumsum(nas))))
    mb = mbtypes_array(mbtypes)
    nx2 = o.grid2(rcut, dgrids[0])
    nx3 = o.grid3(dgrids[1])
    D = slatm_dim(mbtypes, nx2, nx3)
    nrow = int(off[-1]) if local else len(nas)
    out = np.zeros((nrow, D))
    lib.oracle_generate_slatm_batch(_p(coords), _p(zs), _p(off), C.c_int(len(nas)), _p(mb), C.c_int(len(mb)),
                                    C.c_int(int(local)), C.c_int(int(izeff)),
                                    C.c_double(sigmas[0]), C.c_double(sigmas[1]), C.c_double(dgrids[0]),
                                    C.c_double(dgrids[1]), C.c_double(rcut), C.c_double(rpower),
                                    C.c_int(nx2), C.c_int(nx3), C.c_int(D), _p(out))
    return out


def gaussian_kernel(A, B, sigma, fast=False):
    A, B = _d(A), _d(B)
    K = np.empty((A.shape[0], B.shape[0]), order='F')
    load(fast).oracle_fgaussian_kernel(_p(A), C.c_int(A.shape[0]), _p(B), C.c_int(B.shape[0]),
                                       C.c_int(A.shape[1]), _p(K), C.c_double(sigma))
    return K


def laplacian_kernel(A, B, sigma, fast=False):
    A, B = _d(A), _d(B)
    K = np.empty((A.shape[0], B.shape[0]), order='F')
    load(fast).oracle_flaplacian_kernel(_p(A), C.c_int(A.shape[0]), _p(B), C.c_int(B.shape[0]),
                                        C.c_int(A.shape[1]), _p(K), C.c_double(sigma))
    return K


def l2_distance(A, B, fast=False):
    A, B = _d(A), _d(B)
    Dm = np.empty((A.shape[0], B.shape[0]), order='F')
    load(fast).oracle_fl2_distance(_p(A), _p(B), C.c_int(A.shape[0]), C.c_int(B.shape[0]), C.c_int(A.shape[1]), _p(Dm))
    return Dm


def l1_distance(A, B, fast=False):
    A, B = _d(A), _d(B)
    Dm = np.empty((A.shape[0], B.shape[0]), order='F')
    load(fast).oracle_fl1_distance(_p(A), _p(B), C.c_int(A.shape[0]), C.c_int(B.shape[0]), C.c_int(A.shape[1]), _p(Dm))
    return Dm


def calc_lk_gaussian(x1, x2, zs1, zs2, nas1, nas2, zmax, iab, sigmas, fast=False):
    x1, x2 = _d(x1), _d(x2)
    zs1, zs2, nas1, nas2 = _i(zs1), _i(zs2), _i(nas1), _i(nas2)
    sig = _d(sigmas).reshape(len(sigmas), -1)
    assert sig.shape[1] == zmax
    K = np.zeros((sig.shape[0], len(nas1), len(nas2)))
    load(fast).oracle_calc_lk_gaussian(_p(x1), _p(x2), _p(zs1), _p(zs2), _p(nas1), _p(nas2), C.c_int(zmax),
                                       C.c_int(int(bool(iab))), _p(sig), C.c_int(len(nas1)), C.c_int(len(nas2)),
                                       C.c_int(sig.shape[0]), C.c_int(x1.shape[1]), _p(K))
    return K


def calc_lk_laplacian(x1, x2, nas1, nas2, sigmas, fast=False):
    x1, x2 = _d(x1), _d(x2)
    nas1, nas2 = _i(nas1), _i(nas2)
    sig = _d(sigmas).ravel()
    K = np.zeros((len(sig), len(nas1), len(nas2)))
    load(fast).oracle_calc_lk_laplacian(_p(x1), _p(x2), _p(nas1), _p(nas2), _p(sig), C.c_int(len(nas1)),
                                        C.c_int(len(nas2)), C.c_int(len(sig)), C.c_int(x1.shape[1]), _p(K))
    return K


def get_kernel_width(nas, zs, x, kernel='g', fast=False):
    """aqml.py:204-255 (cab=False branch) on top of the C distance loops."""
    zs = np.asarray(zs)
    x = _d(x)
    zsu = np.unique(zs)
    dist = l2_distance if kernel[0] == 'g' else l1_distance
    dso = np.ones(int(zsu[-1])) * 1.e-9
    for z in zsu:
        xi = x[zs == z]
        dso[z - 1] = np.max(dist(xi, xi, fast=fast))
    return dso
