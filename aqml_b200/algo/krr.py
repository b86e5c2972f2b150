"""KRR consumers of the kernel matrices, as plain functions (what coreml/cml/algo/aqml.py:129-181,806-897,
krr.py:294-299 and rkrr.py:171-177,203-228 compute).  They sit OUTSIDE the timed hot path (the reference solves with
NumPy LAPACK on the host) and exist so that the parity tests can check predictions (<= 1e-8 Ha), not to re-create the
reference's driver objects -- its ``krr`` / ``rkrr`` classes (label bookkeeping, index selection, printing) are out of
scope.  ``device=True`` solves with torch.linalg on the GPU."""
from __future__ import annotations

import numpy as np

H2KC = 627.5094738898777  # io2/__init__.py:45


def _indep_cols(M):
    """math/matrix.py:18-33."""
    mr = np.linalg.matrix_rank
    idxc = [0]
    for i in range(1, M.shape[1]):
        if mr(M[:, idxc + [i]]) > mr(M[:, idxc]):
            idxc.append(i)
    return idxc


def calc_ae_dressed(ngs, ys, idx1, idx2, free_atom_energies=None):
    """aqml.py:129-181 (``ref is None``): dressed-atom baseline by least squares on element counts;
    rank-deficient training sets fall back to the given free-atom energies (same units as ys)."""
    ns1, ns2 = ngs[idx1], ngs[idx2]
    ns = np.concatenate((ns1, ns2), axis=0)
    ys1, ys2 = ys[idx1], ys[idx2]
    mr = np.linalg.matrix_rank
    istat = True
    if mr(ns1) < ns1.shape[1]:
        if mr(ns1) == mr(ns) and _indep_cols(ns1) == _indep_cols(ns):
            cols = _indep_cols(ns)
            ns1, ns2 = ns1[:, cols], ns2[:, cols]
            esb = np.linalg.lstsq(ns1, ys1, rcond=None)[0]
        else:
            if free_atom_energies is None:
                raise ValueError("rank-deficient element counts and no free-atom energies given")
            esb = np.asarray(free_atom_energies, dtype=np.float64)
    else:
        esb = np.linalg.lstsq(ns1, ys1, rcond=None)[0]
    return istat, ys1 - np.dot(ns1, esb), ys2 - np.dot(ns2, esb), esb


def solve(k1, ys1, llambda, device=False):
    """alphas = (K + lambda I)^-1 y   (aqml.py:874-875, krr.py:294-296: LU solve)."""
    if device:
        import torch
        kt = torch.as_tensor(k1, dtype=torch.float64, device="cuda").clone()
        kt.diagonal().add_(llambda)
        return torch.linalg.solve(kt, torch.as_tensor(ys1, dtype=torch.float64, device="cuda")).cpu().numpy()
    k1 = np.array(k1, dtype=np.float64, copy=True)
    k1[np.diag_indices_from(k1)] += llambda
    return np.linalg.solve(k1, ys1)


def learning_curve(ks1, ks2, ngs, ys, nheav1, free_atom_energies=None, llambda=1e-4, device=False):
    """aqml.py:806-897 for one coeff / lambda.  ks1 (N1,N1), ks2 (N2,N1); ys (N1+N2,) train then test.
    Returns rows (N_I, n_train, mae, rmse, max|err|, predictions incl. the dressed-atom baseline removed)."""
    ks1 = np.asarray(ks1)
    ks2 = np.asarray(ks2)
    n1, n2 = ks1.shape[0], ks2.shape[0]
    ims1 = np.arange(n1)
    idx2 = np.arange(n1, n1 + n2)
    nheav1 = np.asarray(nheav1)
    rows = []
    for ni in range(1, 1 + int(nheav1.max())):
        idx1 = ims1[nheav1 <= ni]
        if len(idx1) == 0:
            continue
        _, ys1, ys2, esb = calc_ae_dressed(ngs, ys, idx1, idx2, free_atom_energies)
        alphas = solve(ks1[idx1][:, idx1], ys1, llambda, device)
        ys2p = np.dot(ks2[:, idx1], alphas)
        dys2 = ys2p - ys2
        rows.append((ni, len(idx1), float(np.mean(np.abs(dys2))), float(np.sqrt(np.mean(dys2 * dys2))),
                     float(np.max(np.abs(dys2))), ys2p))
    return rows


def e_base(nzs1, ys1, nzs2):
    """rkrr.py:171-177 calc_e_base: least-squares dressed-atom baseline on the element counts."""
    esb = np.linalg.lstsq(nzs1, ys1, rcond=None)[0]
    return ys1 - np.dot(nzs1, esb), np.dot(nzs2, esb)


def multi_fidelity_predict(ks1, ks2, ys, n1s, nzs1, nzs2, llambda=1e-10, device=False):
    """rkrr.py:203-228,236-265: recursive / multi-fidelity KRR on ONE kernel matrix.

    ks1 (N,N), ks2 (N2,N) -- N training molecules ordered so that every fidelity level uses a prefix;
    ys (N, nl): labels per level (level 0 = cheapest); n1s[l] = prefix length trained at level l
    (non-increasing); nzs1 (N, n_el) / nzs2 (N2, n_el) element counts for the dressed-atom baselines.
    Level 0 learns y_0, level l>0 learns y_l - y_{l-1}; the prediction is the sum over levels.  Only
    nested index subsets of the same K differ between levels -- no extra kernel work."""
    ks1 = np.asarray(ks1)
    ks2 = np.asarray(ks2)
    ys = np.asarray(ys, dtype=np.float64)
    pred = np.zeros(ks2.shape[0])
    for l, n in enumerate(n1s):
        ims = np.arange(int(n))
        tgt = ys[ims, l] if l == 0 else ys[ims, l] - ys[ims, l - 1]
        y1, base2 = e_base(nzs1[ims], tgt, nzs2)
        alphas = solve(ks1[ims][:, ims], y1, llambda, device)
        pred += np.dot(ks2[:, ims], alphas) + base2
    return pred
