"""Device-resident fused path: molecules -> element-packed aSLATM -> local Gaussian kernel.

The reference materialises dense (atoms, D) arrays on the host between ``generate_slatm`` and
``get_local_kernels_gaussian`` (algo/aqml.py:711-789).  Here the aSLATM kernel writes each atom
directly into the operand layout the FP64 tensor-core contraction reads (rows grouped by element,
columns restricted to that element's support), so neither the dense matrix nor a host round trip
ever exists.  Results are identical to the two-step route (tests/test_gpu_fused.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class _DeviceArray:
    """CUDA array interface over a raw device allocation (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, shape, typestr="|u1"):
        self.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class PackedRep:
    """aSLATM of a set of molecules, packed by element on the current CUDA device.

    ``compute=False`` only lays out the storage (``coordinates`` may be None): the contents are then received
    from the rank that computed them (``slab_tensor`` + a broadcast, ``parallel.share_reps``); ``rows=False`` (with
    ``compute=False``) leaves the packed FP64 rows out -- enough for the int8 tensor engine, which reads the digit planes."""

    def __init__(self, coordinates, nuclear_charges, nas, mbtypes, izeff=False, sigmas=(0.05, 0.05),
                 dgrids=(0.03, 0.03), rcut=4.8, rpower=6, compute=True, rows=True):
        import torch
        self._h = C.c_void_p()
        self.nas = L.i32(nas)
        self.zs = L.i32(nuclear_charges)
        self.nmol = len(self.nas)
        if len(self.zs) != int(self.nas.sum()):
            raise ValueError(f"nuclear_charges has {len(self.zs)} entries, nas sums to {int(self.nas.sum())}")
        p = L.slatm_params(mbtypes, izeff, sigmas, dgrids, rcut, rpower)
        on_dev = L.is_tensor(coordinates)
        if coordinates is None:
            if compute:
                raise ValueError("coordinates are required unless compute=False")
            coords = None
        else:
            coords = L.dev_f64(coordinates).contiguous() if on_dev else L.f64(coordinates)
            if tuple(coords.shape) != (len(self.zs), 3):
                raise ValueError(f"coordinates must have shape ({len(self.zs)}, 3), got {tuple(coords.shape)}")
        self.device = coords.device if on_dev else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            L.call("aqml_rep_create_ex", L.ptr(coords), int(on_dev), L.ptr(self.zs), L.ptr(self.nas), self.nmol, p,
                   (0 if compute else 1) | (0 if rows else 2), L.current_stream(), C.byref(self._h))
        self.has_rows = bool(rows)

    def slab_tensor(self, which=0):
        """The representation's storage as a uint8 CUDA tensor (a view: valid while this object lives): which = 0 what the
        pair kernel reads (norms, molecule index, masks, digit planes, exponents), which = 1 the packed FP64 rows."""
        import torch
        ptr, nb = C.c_void_p(), C.c_int64()
        L.call("aqml_rep_slab", self._h, int(which), C.byref(ptr), C.byref(nb))
        if not ptr.value:
            raise L.AqmlError("this representation was built without its FP64 rows")
        return torch.as_tensor(_DeviceArray(ptr.value, (nb.value,)), device=self.device)

    def element_rows(self):
        """{Z: (n, ld) float64 CUDA view of the element's valid packed rows} (views: valid while this object lives; n may be 0)."""
        import torch
        out = {}
        for e in range(int(L.load().aqml_rep_elements(self._h))):
            label, n, ld, x = C.c_int(), C.c_int(), C.c_int(), C.c_void_p()
            L.call("aqml_rep_element", self._h, e, C.byref(label), C.byref(n), C.byref(ld), C.byref(x))
            if x.value:
                if n.value > 0:
                    out[int(label.value)] = torch.as_tensor(_DeviceArray(x.value, (n.value, ld.value), "<f8"), device=self.device)
                else:   # an element of the many-body types without an atom in this set: an empty view with the right pitch
                    out[int(label.value)] = torch.empty((0, ld.value), dtype=torch.float64, device=self.device)
        return out

    @property
    def nbytes(self):
        return int(L.load().aqml_rep_bytes(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.load().aqml_rep_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def local_gaussian(rep1, rep2, sigmas, zmax, row0=0, nrows=None, out=None, symmetric=False):
    """K[k, a, b] for a in rep1[row0:row0+nrows], b in rep2; sigmas (nsigmas, zmax).  CUDA tensor.
    ``symmetric=True`` (rep1 is rep2, all rows): contract only the upper triangle and mirror."""
    import torch
    sig = L.f64(sigmas)
    sig = sig.reshape(len(sig), -1)
    if sig.shape[1] != zmax:
        raise ValueError(f"sigmas must have shape (nsigmas, zmax={zmax})")
    if symmetric:
        if rep1 is not rep2 or row0 != 0 or nrows not in (None, rep1.nmol):
            raise ValueError("symmetric=True needs rep1 is rep2 and the full row range")
        K = out if out is not None else torch.empty((sig.shape[0], rep1.nmol, rep1.nmol), dtype=torch.float64,
                                                    device=rep1.device)
        with torch.cuda.device(rep1.device):
            L.call("aqml_rep_local_gaussian_sym", rep1._h, L.ptr(sig), sig.shape[0], int(zmax), L.ptr(K),
                   L.current_stream())
        return K
    nrows = rep1.nmol - row0 if nrows is None else nrows
    K = out if out is not None else torch.empty((sig.shape[0], nrows, rep2.nmol), dtype=torch.float64, device=rep1.device)
    with torch.cuda.device(rep1.device):
        L.call("aqml_rep_local_gaussian", rep1._h, rep2._h, L.ptr(sig), sig.shape[0], int(zmax), int(row0), int(nrows),
               L.ptr(K), L.current_stream())
    return K


def kernel_width(rep1, rep2, zmax):
    """Per-element max L2 distance between rep1 and rep2 atoms (get_kernel_width, aqml.py:204-255).
    ``rep1`` / ``rep2`` may be lists of PackedRep (a data set packed in blocks): the maximum runs over all of them."""
    import torch
    l1 = list(rep1) if isinstance(rep1, (list, tuple)) else [rep1]
    l2 = list(rep2) if isinstance(rep2, (list, tuple)) else [rep2]
    dso = np.zeros(zmax)
    h1 = (C.c_void_p * len(l1))(*[r._h for r in l1])
    h2 = h1 if (len(l1) == len(l2) and all(a is b for a, b in zip(l1, l2))) else (C.c_void_p * len(l2))(*[r._h for r in l2])
    with torch.cuda.device(l1[0].device):
        L.call("aqml_rep_kernel_width_multi", h1, len(l1), h2, len(l2), int(zmax), L.ptr(dso), L.current_stream())
    return dso


def local_gaussian_to_host(coordinates, nuclear_charges, nas, rep_train, mbtypes, kernel_sigmas, zmax, out_host, chunks=2,
                           **slatm_kw):
    """K[k, a, b] between HOST-resident query molecules (coordinates (A,3), nuclear_charges (A,), nas) and a packed
    training set, delivered into the pinned host tensor ``out_host`` (nsigmas, len(nas), rep_train.nmol).

    The queries are processed in ``chunks`` row blocks (an int, or a tuple of relative block sizes): block c is packed and contracted on the current stream while
    block c-1 travels device -> host on a second stream (two device buffers), so the 8 bytes per kernel element of the
    result leave the GPU behind the compute instead of after it, and K never has to fit in device memory (two block
    buffers).  Two unequal blocks, (0.75, 0.25), are the sweet spot on the bench workload: e2e 338 ms against 351 ms for
    one contraction + one copy (only the last block's copy is exposed; every further block adds packing and a kernel tail).  Returns after everything has
    been enqueued; the current stream is made to wait for the copies (synchronise it before reading ``out_host``)."""
    import torch
    nas = np.asarray(nas)
    zs = np.asarray(nuclear_charges)
    coords = np.asarray(coordinates, dtype=np.float64).reshape(-1, 3)
    sig = L.f64(kernel_sigmas)   # (nsigmas, zmax) kernel widths; **slatm_kw are PackedRep's arguments (sigmas = smearing widths)
    sig = sig.reshape(len(sig), -1)
    nq, nsig = len(nas), sig.shape[0]
    off = np.concatenate(([0], np.cumsum(nas)))
    if isinstance(chunks, (list, tuple)):  # relative block sizes, e.g. (0.7, 0.2, 0.1): only the last block's copy is exposed
        fr = np.cumsum(np.asarray(chunks, dtype=np.float64))
        bounds = np.unique(np.concatenate(([0], np.rint(fr / fr[-1] * nq).astype(int))))
    else:
        bounds = np.linspace(0, nq, min(max(int(chunks), 1), max(nq, 1)) + 1).astype(int)
    dev = rep_train.device
    with torch.cuda.device(dev):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        rows_max = int(np.max(np.diff(bounds))) if nq else 0
        bufs = [torch.empty((nsig, rows_max, rep_train.nmol), dtype=torch.float64, device=dev) for _ in range(2)]
        done = [None, None]
        for c in range(len(bounds) - 1):
            lo, hi = int(bounds[c]), int(bounds[c + 1])
            if hi == lo:
                continue
            b = c & 1
            if done[b] is not None:
                cur.wait_event(done[b])          # the copy that last read this buffer has finished
            rep = PackedRep(coords[off[lo]:off[hi]], zs[off[lo]:off[hi]], nas[lo:hi], mbtypes, **slatm_kw)
            Kc = bufs[b][:, :hi - lo]
            if rows_max != hi - lo:
                Kc = torch.empty((nsig, hi - lo, rep_train.nmol), dtype=torch.float64, device=dev)
            local_gaussian(rep, rep_train, sig, zmax, out=Kc)
            rep.close()
            ready = torch.cuda.Event()
            ready.record(cur)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                for s in range(nsig):
                    out_host[s, lo:hi].copy_(Kc[s], non_blocking=True)
                Kc.record_stream(side)
                done[b] = torch.cuda.Event()
                done[b].record(side)
        cur.wait_stream(side)
    return out_host
