"""Host mirror of ``cml.kernels`` (coreml/cml/kernels.py) on top of the CUDA library.

Same names, argument order and return shapes as the reference; NumPy in -> NumPy out
(synchronous, like an f2py call), CUDA torch tensor in -> CUDA torch tensor out
(stream-ordered, nothing leaves the device).  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def _global_np(fn, A, B, sigma):
    A, B = L.f64(A), L.f64(B)
    if A.ndim != 2 or B.ndim != 2 or A.shape[1] != B.shape[1]:
        raise ValueError("expected 2-D arrays with the same representation size")
    K = np.empty((A.shape[0], B.shape[0]), order="F")  # kernels.py:32,61
    L.call(fn, L.ptr(A), A.shape[0], L.ptr(B), B.shape[0], A.shape[1], L.ptr(K), float(sigma))
    return K


def global_kernels(A, B, sigmas, kernel="gaussian", center=True):
    """All sigmas in one pass: returns (nsigmas, N, M).  Device tensors stay on the device."""
    kind = L.KERNEL_GAUSSIAN if kernel[0] == "g" else L.KERNEL_LAPLACIAN
    sig = L.f64(np.atleast_1d(sigmas))
    if L.is_tensor(A):
        import torch
        A, B = L.dev_f64(A), L.dev_f64(B)
        K = torch.empty((len(sig), A.shape[0], B.shape[0]), dtype=torch.float64, device=A.device)
        with torch.cuda.device(A.device):
            L.call("aqml_global_kernel_dev", kind, L.ptr(A), A.shape[0], A.stride(0), L.ptr(B), B.shape[0],
                   B.stride(0), A.shape[1], L.ptr(sig), len(sig), int(bool(center)), L.ptr(K), L.current_stream())
        return K
    return np.stack([(gaussian_kernel if kind == 0 else laplacian_kernel)(A, B, s) for s in sig])


def gaussian_kernel(A, B, sigma):
    """kernels.py:40-67 -> fkernels.f90:327 fgaussian_kernel.  K (N, M), Fortran-ordered."""
    if L.is_tensor(A):
        return global_kernels(A, B, [sigma], "gaussian")[0]
    return _global_np("aqml_fgaussian_kernel", A, B, sigma)


def laplacian_kernel(A, B, sigma):
    """kernels.py:11-38 -> fkernels.f90:361 flaplacian_kernel."""
    if L.is_tensor(A):
        return global_kernels(A, B, [sigma], "laplacian")[0]
    return _global_np("aqml_flaplacian_kernel", A, B, sigma)


def _device_distance(A, B, p):
    """(N, M) CUDA tensor of |A_i - B_j|_p (p = 1, 2) or A_i . B_j (p = 0); NumPy inputs are uploaded."""
    import torch
    if not L.is_tensor(A):
        A = torch.as_tensor(L.f64(A), device="cuda")
        B = torch.as_tensor(L.f64(B), device="cuda")
    A, B = L.dev_f64(A), L.dev_f64(B)
    if A.dim() != 2 or B.dim() != 2 or A.shape[1] != B.shape[1]:
        raise ValueError("expected 2-D arrays with the same representation size")
    out = torch.empty((A.shape[0], B.shape[0]), dtype=torch.float64, device=A.device)
    with torch.cuda.device(A.device):
        L.call("aqml_distance_dev", int(p), L.ptr(A), A.shape[0], A.stride(0), L.ptr(B), B.shape[0], B.stride(0),
               A.shape[1], L.ptr(out), L.current_stream())
    return out


def _like_input(K, A):
    return K if L.is_tensor(A) else np.asfortranarray(K.cpu().numpy())


def linear_kernel(A, B):
    """kernels.py:69-95 -> fkernels.f90:390 flinear_kernel: K_ij = A_i . B_j (FP64 tensor-core Gram matrix)."""
    if L.is_tensor(A):
        return _device_distance(A, B, 0)
    A, B = L.f64(A), L.f64(B)
    K = np.empty((A.shape[0], B.shape[0]), order="F")
    L.call("aqml_flinear_kernel", L.ptr(A), A.shape[0], L.ptr(B), B.shape[0], A.shape[1], L.ptr(K))
    return K


def sargan_kernel(A, B, sigma, gammas):
    """kernels.py:97-131 -> fkernels.f90:478 fsargan_kernel:
    K = exp(-d/sigma) (1 + sum_k gamma_k (d/sigma)^k), d = |A_i - B_j|_1 (device), scalar part in torch."""
    import torch
    if len(gammas) == 0:
        return laplacian_kernel(A, B, sigma)
    d = _device_distance(A, B, 1)
    x = d / sigma
    poly = torch.ones_like(x)
    for m, g in enumerate(gammas, start=1):
        poly = poly + float(g) * x ** m
    return _like_input(torch.exp(-x) * poly, A)


def matern_kernel(A, B, sigma, order=0, metric="l1"):
    """kernels.py:133-191 -> fsargan_kernel (l1) / fkernels.f90:414 fmatern_kernel_l2 (l2)."""
    import torch
    if metric == "l1":
        if order == 0:
            gammas = []
        elif order == 1:
            gammas, sigma = [1], sigma / np.sqrt(3)
        elif order == 2:
            gammas, sigma = [1, 1 / 3.0], sigma / np.sqrt(5)
        else:
            print("Order:%d not implemented in Matern Kernel" % order)
            raise SystemExit
        return sargan_kernel(A, B, sigma, gammas)
    elif metric != "l2":
        print("Error: Unknown distance metric %s in Matern kernel" % str(metric))
        raise SystemExit
    d = _device_distance(A, B, 2)
    if order == 0:
        K = torch.exp(-d / sigma)
    elif order == 1:
        s = np.sqrt(3.0) / sigma
        K = torch.exp(-s * d) * (1.0 + s * d)
    else:
        s = np.sqrt(5.0) / sigma
        K = torch.exp(-s * d) * (1.0 + s * d + 5.0 / (3.0 * sigma * sigma) * d * d)
    return _like_input(K, A)


# algo/aqml.py:194,198 names the global multi-sigma routines like this (absent upstream)
def gaussian_kernels(A, B, sigmas):
    return global_kernels(A, B, sigmas, "gaussian")


def laplacian_kernels(A, B, sigmas):
    return global_kernels(A, B, sigmas, "laplacian")


def get_local_kernels_gaussian(x1, x2, zs1, zs2, nas1, nas2, iab, zmax, sigmas, center=True):
    """kernels.py:193-235 -> fkernels.f90:23 calc_lk_gaussian.  Returns (nsigmas, N1, N2)."""
    zs1, zs2, nas1, nas2 = L.i32(zs1), L.i32(zs2), L.i32(nas1), L.i32(nas2)
    assert np.sum(nas1) == x1.shape[0], "Error in A input"
    assert np.sum(nas2) == x2.shape[0], "Error in B input"
    assert x1.shape[1] == x2.shape[1], "Error in representation sizes"
    if len(zs1) != x1.shape[0] or len(zs2) != x2.shape[0]:   # raw pointers below: never let a short zs reach the device
        raise ValueError(f"zs1 / zs2 have {len(zs1)} / {len(zs2)} entries for {x1.shape[0]} / {x2.shape[0]} atoms")
    sig = L.f64(sigmas)
    nsig = len(sig)
    sig = sig.reshape(nsig, -1)
    if sig.shape[1] != zmax:
        raise ValueError(f"sigmas must have shape (nsigmas, zmax={zmax})")
    nm1, nm2 = len(nas1), len(nas2)
    if L.is_tensor(x1):
        import torch
        x1, x2 = L.dev_f64(x1), L.dev_f64(x2)
        K = torch.empty((nsig, nm1, nm2), dtype=torch.float64, device=x1.device)
        with torch.cuda.device(x1.device):
            L.call("aqml_calc_lk_gaussian_dev", L.ptr(x1), x1.stride(0), L.ptr(x2), x2.stride(0), L.ptr(zs1),
                   L.ptr(zs2), L.ptr(nas1), L.ptr(nas2), int(zmax), int(bool(iab)), L.ptr(sig), nm1, nm2, nsig,
                   x1.shape[1], int(bool(center)), L.ptr(K), L.current_stream())
        return K
    same = x1 is x2
    x1 = L.f64(x1)
    x2 = x1 if same else L.f64(x2)
    K = np.empty((nsig, nm1, nm2))
    L.call("aqml_calc_lk_gaussian", L.ptr(x1), L.ptr(x2), L.ptr(zs1), L.ptr(zs2), L.ptr(nas1), L.ptr(nas2),
           int(zmax), int(bool(iab)), L.ptr(sig), nm1, nm2, nsig, x1.shape[1], L.ptr(K))
    return K


def get_local_kernels_laplacian(A, B, na, nb, sigmas):
    """kernels.py:237-276 -> fkernels.f90:97 calc_lk_laplacian.  Returns (nsigmas, N1, N2)."""
    na, nb = L.i32(na), L.i32(nb)
    assert np.sum(na) == A.shape[0], "Error in A input"
    assert np.sum(nb) == B.shape[0], "Error in B input"
    assert A.shape[1] == B.shape[1], "Error in representation sizes"
    sig = L.f64(sigmas).ravel()
    nsig, nma, nmb = len(sig), len(na), len(nb)
    if L.is_tensor(A):
        import torch
        A, B = L.dev_f64(A), L.dev_f64(B)
        K = torch.empty((nsig, nma, nmb), dtype=torch.float64, device=A.device)
        with torch.cuda.device(A.device):
            L.call("aqml_calc_lk_laplacian_dev", L.ptr(A), A.stride(0), L.ptr(B), B.stride(0), L.ptr(na), L.ptr(nb),
                   L.ptr(sig), nma, nmb, nsig, A.shape[1], L.ptr(K), L.current_stream())
        return K
    same = A is B
    A = L.f64(A)
    B = A if same else L.f64(B)
    K = np.empty((nsig, nma, nmb))
    L.call("aqml_calc_lk_laplacian", L.ptr(A), L.ptr(B), L.ptr(na), L.ptr(nb), L.ptr(sig), nma, nmb, nsig,
           A.shape[1], L.ptr(K))
    return K


# f2py-level names (cml.fkernels), kept for callers that import them directly (kernels.py:2-8)
def fgaussian_kernel(a_t, na, b_t, nb, k, sigma):
    """fkernels.f90:327: a_t, b_t are the (D, n) transposed views; k (na, nb) F-ordered, filled in place."""
    k[...] = _global_np("aqml_fgaussian_kernel", np.asarray(a_t).T, np.asarray(b_t).T, sigma)


def flaplacian_kernel(a_t, na, b_t, nb, k, sigma):
    k[...] = _global_np("aqml_flaplacian_kernel", np.asarray(a_t).T, np.asarray(b_t).T, sigma)


def flinear_kernel(a_t, na, b_t, nb, k):
    k[...] = linear_kernel(np.asarray(a_t).T, np.asarray(b_t).T)


def fsargan_kernel(a_t, na, b_t, nb, k, sigma, gammas, ng=None):
    k[...] = sargan_kernel(np.asarray(a_t).T, np.asarray(b_t).T, sigma, gammas)


def fmatern_kernel_l2(a_t, na, b_t, nb, k, sigma, order):
    k[...] = matern_kernel(np.asarray(a_t).T, np.asarray(b_t).T, sigma, order, "l2")


def calc_lk_gaussian(x1_t, x2_t, zs1, zs2, nas1, nas2, zmax, iab, sigmas, nm1=None, nm2=None, nsigmas=None):
    return get_local_kernels_gaussian(np.asarray(x1_t).T, np.asarray(x2_t).T, zs1, zs2, nas1, nas2, iab, zmax, sigmas)


def calc_lk_laplacian(x1_t, x2_t, nas1, nas2, sigmas, nm1=None, nm2=None, nsigmas=None):
    return get_local_kernels_laplacian(np.asarray(x1_t).T, np.asarray(x2_t).T, nas1, nas2, sigmas)


def _unpad(x_t, nas):
    """Fortran x(D, maxatoms, nm) (f2py view of a C-ordered (nm, maxatoms, D) array) -> (sum nas, D) rows."""
    x = np.asarray(x_t)
    if x.ndim != 3:
        raise ValueError("expected a 3-D padded array")
    x = x.transpose(2, 1, 0) if x.flags.f_contiguous or x.shape[0] != len(nas) else x
    return np.concatenate([x[i, :int(n)] for i, n in enumerate(nas)])


def calc_vector_kernels_gaussian(x1_t, x2_t, nas1, nas2, sigmas, nm1=None, nm2=None, nsigmas=None):
    """fkernels.f90:255-324: local Gaussian kernel on zero-padded per-molecule arrays, every atom pair, one
    sigma set.  (No Python wrapper upstream; provided as the f2py-level name.)"""
    a, b = _unpad(x1_t, nas1), _unpad(x2_t, nas2)
    sig = L.f64(sigmas).reshape(-1, 1)
    return get_local_kernels_gaussian(a, b, np.ones(len(a), np.int32), np.ones(len(b), np.int32), nas1, nas2, True, 1, sig)


def calc_vector_kernels_laplacian(x1_t, x2_t, nas1, nas2, sigmas, nm1=None, nm2=None, nsigmas=None):
    """fkernels.f90:185-253."""
    return get_local_kernels_laplacian(_unpad(x1_t, nas1), _unpad(x2_t, nas2), nas1, nas2, sigmas)
