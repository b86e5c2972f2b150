"""ctypes binding of the C ABI in include/aqml_b200.h (libaqml_b200.so, built in-tree).

There is no CPU fallback: if the shared library is missing, or no CUDA device is visible,
every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libaqml_b200.so")

KERNEL_GAUSSIAN = 0
KERNEL_LAPLACIAN = 1


class AqmlError(RuntimeError):
    pass


class SlatmParams(C.Structure):
    _fields_ = [("nmb", C.c_int), ("mbtypes", C.c_void_p), ("izeff", C.c_int), ("nx2", C.c_int), ("nx3", C.c_int),
                ("sigma2", C.c_double), ("sigma3", C.c_double), ("dgrid2", C.c_double), ("dgrid3", C.c_double),
                ("coeff2", C.c_double), ("coeff3", C.c_double), ("rcut", C.c_double), ("rpower", C.c_double)]


_P, _I, _D, _L = C.c_void_p, C.c_int, C.c_double, C.c_int64

# name -> (restype, argtypes); must list every symbol include/aqml_b200.h declares
SIGNATURES = {
    "aqml_version": (_I, []),
    "aqml_last_error": (C.c_char_p, []),
    "aqml_device_count": (_I, []),
    "aqml_launch_count": (_L, []),
    "aqml_pair_stats": (_I, [_P, _P, _I]),
    "aqml_tile_shape": (None, [_P, _P, _P]),
    "aqml_measure_i8_peak": (_I, [_P, _P]),
    "aqml_pairwise_dims": (_I, [_P, _D, _I, _I, _I, _P, _P, _P]),
    "aqml_pairwise_local_spectrum_dev": (_I, [_P, _P, _P, _I, _P, _I, _P, _I, _P, _I, _P, _I, _P, _D, _D, _D, _D, _D, _D, _P, _P, _P]),
    "aqml_fgaussian_kernel": (_I, [_P, _I, _P, _I, _I, _P, _D]),
    "aqml_flaplacian_kernel": (_I, [_P, _I, _P, _I, _I, _P, _D]),
    "aqml_fl2_distance": (_I, [_P, _P, _I, _I, _I, _P]),
    "aqml_fl1_distance": (_I, [_P, _P, _I, _I, _I, _P]),
    "aqml_flinear_kernel": (_I, [_P, _I, _P, _I, _I, _P]),
    "aqml_distance_dev": (_I, [_I, _P, _I, _L, _P, _I, _L, _I, _P, _P]),
    "aqml_global_kernel_dev": (_I, [_I, _P, _I, _L, _P, _I, _L, _I, _P, _I, _I, _P, _P]),
    "aqml_max_distance_dev": (_I, [_I, _P, _I, _L, _P, _I, _L, _I, _P, _P]),
    "aqml_calc_lk_gaussian": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _P]),
    "aqml_calc_lk_gaussian_dev": (_I, [_P, _L, _P, _L, _P, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P]),
    "aqml_calc_lk_laplacian": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "aqml_calc_lk_laplacian_dev": (_I, [_P, _L, _P, _L, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "aqml_kernel_width_dev": (_I, [_I, _P, _L, _P, _I, _I, _I, _P, _P]),
    "aqml_slatm_dim": (_I, [C.POINTER(SlatmParams)]),
    "aqml_calc_sbop": (_I, [_P, _P, _I, _I, _I, _I, _I, _D, _I, _D, _D, _D, _D, _P]),
    "aqml_calc_sbot": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _D, _I, _D, _D, _D, _P]),
    "aqml_generate_slatm_batch": (_I, [_P, _P, _P, _I, C.POINTER(SlatmParams), _I, _P]),
    "aqml_generate_slatm_batch_dev": (_I, [_P, _P, _P, _I, C.POINTER(SlatmParams), _I, _P, _L, _P]),
    "aqml_rep_create": (_I, [_P, _I, _P, _P, _I, C.POINTER(SlatmParams), _P, C.POINTER(_P)]),
    "aqml_rep_create_ex": (_I, [_P, _I, _P, _P, _I, C.POINTER(SlatmParams), _I, _P, C.POINTER(_P)]),
    "aqml_rep_slab": (_I, [_P, _I, C.POINTER(_P), C.POINTER(_L)]),
    "aqml_rep_elements": (_I, [_P]),
    "aqml_rep_element": (_I, [_P, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_P)]),
    "aqml_rows_colsum_dev": (_I, [_P, _L, _I, _I, _P, _P]),
    "aqml_rows_dist_dev": (_I, [_I, _P, _L, _I, _P, _I, _P, _P]),
    "aqml_rep_destroy": (None, [_P]),
    "aqml_rep_bytes": (_L, [_P]),
    "aqml_rep_local_gaussian": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "aqml_rep_local_gaussian_sym": (_I, [_P, _P, _I, _I, _P, _P]),
    "aqml_rep_kernel_width": (_I, [_P, _P, _I, _P, _P]),
    "aqml_rep_kernel_width_multi": (_I, [_P, _I, _P, _I, _I, _P, _P]),
}

_lib = None


def load():
    """Load libaqml_b200.so (raises if it has not been built: run ``python -m aqml_b200.build``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AqmlError(f"{LIB_PATH} is missing: build it with `python -m aqml_b200.build` "
                        "(aqml_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().aqml_last_error().decode("utf-8", "replace")
        raise AqmlError(f"aqml_b200 error {rc}: {msg}")


def call(name, *args):
    check(getattr(load(), name)(*args))


# ------------------------------------------------------------------------------------------
# array plumbing
# ------------------------------------------------------------------------------------------
def is_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a):
    """Host pointer of a NumPy array / device pointer of a torch tensor."""
    if a is None:
        return None
    if is_tensor(a):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def dev_f64(t):
    """Contiguous float64 CUDA tensor (rows contiguous) from a tensor."""
    import torch
    if t.dtype != torch.float64 or not t.is_cuda:
        raise AqmlError("device inputs must be float64 CUDA tensors")
    return t if t.stride(-1) == 1 else t.contiguous()


def current_stream(device=None):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def slatm_params(mbtypes, izeff=False, sigmas=(0.05, 0.05), dgrids=(0.03, 0.03), rcut=4.8, rpower=6,
                 normalize=True):
    """Resolve generate_slatm's keyword surface into the C struct (slatm.py:60-65,116-122)."""
    mb = np.full((len(mbtypes), 3), -1, dtype=np.int32)
    for i, t in enumerate(mbtypes):
        t = [int(v) for v in t]
        if not 1 <= len(t) <= 3:
            raise AqmlError(f"many-body type {t} must have 1..3 elements")
        mb[i, :len(t)] = t
    r0 = 0.1
    nx2 = int((rcut - r0) / dgrids[0]) + 1
    d2r = np.pi / 180
    a0 = -20.0 * d2r
    a1 = np.pi + 20.0 * d2r
    nx3 = int((a1 - a0) / dgrids[1]) + 1
    c2 = 1 / np.sqrt(2 * sigmas[0] ** 2 * np.pi) if normalize else 1.0
    c3 = 1 / np.sqrt(2 * sigmas[1] ** 2 * np.pi) if normalize else 1.0
    p = SlatmParams(len(mb), mb.ctypes.data, int(bool(izeff)), nx2, nx3, float(sigmas[0]), float(sigmas[1]),
                    float(dgrids[0]), float(dgrids[1]), float(c2), float(c3), float(rcut), float(rpower))
    p._keepalive = mb
    return p
