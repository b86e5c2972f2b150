"""KRR consumers of the kernel matrices (coreml/cml/algo/aqml.py:129-181,806-897; krr.py:185-325;
rkrr.py:179-267).  These sit OUTSIDE the timed hot path (the reference solves with NumPy LAPACK on the
host); they are here so the reference's workflow runs end to end on top of the GPU kernels.  ``device=True``
solves with torch.linalg on the GPU."""
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


class krr(object):
    """The reference's KRR driver object (coreml/cml/algo/krr.py:21-325), host NumPy as upstream: holds the labels and
    the molecule bookkeeping, fixes train / test indices (`get_idx`) and runs the learning curve on precomputed kernel
    matrices (`run`).  After `get_idx` the instance has the fields the block driver reads
    (`cml.representation.slatm_x.slatm(obj, ...)`: zs, coords, nas, nas1, nas2, nzs), so
    ``obj = krr(ys); obj.init_m(mols); obj.get_idx(...); mk = slatm(obj, ck=True, ...); obj.run([mk.mk1[i], mk.mk2[i]])``
    works as in algo/qml.py:362-372.  ``mols`` needs zs, nas, nsheav, coords, nzs and (optionally) imcs."""

    def __init__(self, ys):
        _ys = np.array(ys)
        if len(_ys.shape) == 2:
            nr, nc = _ys.shape
            assert nr > nc, '#ERROR: each column of `ys should correspond to one property!!'
        assert np.logical_not(np.any(np.isnan(_ys))), '#ERROR: property is NaN, check your sdf file'
        self._ys = _ys
        self._nm = len(ys)
        self.n2 = 0

    def init_m(self, obj, nbody=1, icn=False, nproc=1):
        if nbody != 1 or icn:
            raise NotImplementedError('many-body / coordination-number baselines need cheminfo graphs (krr.py:153-161)')
        self.obj = obj
        self._zs = np.asarray(obj.zs)
        self._zsu = np.unique(self._zs)
        self._nas = np.asarray(obj.nas)
        self._nsheav = np.asarray(obj.nsheav)
        self._coords = np.asarray(obj.coords)
        self._ias2 = np.cumsum(self._nas)
        self._ias1 = np.concatenate(([0], self._ias2[:-1])).astype(int)
        self._imcs = np.asarray(getattr(obj, 'imcs', np.zeros(len(self._nas), dtype=bool)))
        self._nzs = np.asarray(obj.nzs)
        if self._nm != len(self._nas):
            raise Exception('#ERROR: len(ys)!= len(_nas)??')

    def get_idx(self, idx=1, n2=None, n1s=None, iaml=True, seed1=1, namin=1, namax=7):
        """krr.py:56-151.  iaml: the last |idx| molecules are the test set, training sets grow with the number of heavy
        atoms; otherwise a seeded random permutation."""
        self.iaml = iaml
        assert isinstance(idx, (int, np.integer)), '#ERROR: idx should be a int'
        idx = abs(int(idx))
        n1s = [] if n1s is None else list(n1s)   # (upstream's mutable default argument accumulates across calls)
        if iaml:
            tidxs = np.arange(self._nm)
            tidx1, idx2 = tidxs[:-idx], tidxs[-idx:]
        else:
            np.random.seed(seed1)
            _tidx = np.random.permutation(self._nm)
            tidx1 = _tidx[:-idx]
            idx2 = _tidx[-n2:] if n2 else _tidx[-idx:]
        self.nm = self._nm
        nsu_heav = np.unique(self._nsheav)
        self.nsu_heav = nsu_heav
        if len(n1s) == 0:
            tidx1_sorted, cnt = [], 0
            t = self._nsheav[tidx1]
            for na in range(namin, namax + 1):
                if np.any(na == nsu_heav):
                    i = tidx1[t == na]
                    tidx1_sorted += list(i)
                    cnt += len(i)
                    n1s.append(cnt)
                else:
                    n1s.append(np.nan if cnt == 0 else cnt)
            tidx1 = np.array(tidx1_sorted, dtype=int)
        self.n1s = n1s
        tidx = np.concatenate((tidx1, idx2)).astype(int)
        n1, n2 = len(tidx1), len(idx2)
        self.n2 = n2
        self.ys = self._ys[tidx]
        _nas = self._nas[tidx]
        self.coords = np.concatenate([self._coords[self._ias1[i]:self._ias2[i]] for i in tidx])
        self.zs = np.concatenate([self._zs[self._ias1[i]:self._ias2[i]] for i in tidx]).astype(int)
        self.nas1 = _nas[:n1]
        self.nas2 = _nas[n1:] if n2 else np.array([], dtype=int)
        self.nas = np.concatenate((self.nas1, self.nas2))
        self.nzs = self._nzs[tidx]
        self.imcs = self._imcs[tidx]
        self.nsheav = self._nsheav[tidx]
        self.nat1, self.nat2 = int(self.nas1.sum()), int(self.nas2.sum())
        self.zs1, self.zs2 = self.zs[:self.nat1], self.zs[self.nat1:]
        self.ias2 = np.cumsum(self.nas)
        self.ias1 = np.concatenate(([0], self.ias2[:-1])).astype(int)

    @staticmethod
    def calc_ae_dressed(imcs1, ns1, ys1, ns2, ys2):
        """krr.py:163-183: dressed-atom baseline from the molecules that are not complexes."""
        flt = np.logical_not(imcs1)
        esb, _, rank, _ = np.linalg.lstsq(ns1[flt], ys1[flt], rcond=None)
        ys2_base = np.zeros(np.array(ys2).shape)
        if rank < len(ns1[0]):
            return ys1, ys2, ys2_base
        ys2_base = np.dot(ns2, esb)
        return ys1 - np.dot(ns1, esb), ys2 - ys2_base, ys2_base

    def run(self, mks, exclude=None, usebl=True, llambda=1e-10, iprt=True, device=False):
        """krr.py:185-325 for a fixed test set (``usek1o=False``): one KRR model per training-set size in `n1s`.
        mks = [k1 (N1,N1), k2 (N2,N1)] or a .npz with keys k1, k2.  Sets maes, rmses, errsmax, dys, n1so."""
        if isinstance(mks, str):
            dic = np.load(mks)
            mks = [dic['k1'].copy(), dic['k2'].copy()]
        mk1, mk2 = np.array(mks[0], dtype=np.float64), np.array(mks[1], dtype=np.float64)
        assert mk1.shape[0] <= self.nm
        if self.n2 == 0:
            raise NotImplementedError('usek1o=T (random test sets out of k1) is not supported; call get_idx with a test set')
        n2 = self.n2
        tidx = np.arange(self.nm)
        idx1, idx2 = tidx[:-n2], tidx[-n2:]
        _ys1, _ys2 = self.ys[idx1], self.ys[idx2]
        n1so, maes, rmses, dys, errsmax = [], [], [], [], []
        for n1 in self.n1s:
            if n1 is None or (isinstance(n1, float) and np.isnan(n1)):
                if iprt: print('%6s' % 'None')
                maes.append(np.nan); rmses.append(np.nan)
                continue
            _idx1 = np.setdiff1d(np.arange(int(n1)), exclude) if exclude else np.arange(int(n1))
            if len(_idx1) == 0:
                continue
            ys1i, ys2i = _ys1[_idx1], _ys2
            if usebl:
                ns1 = self.nzs[idx1][_idx1]
                rank1 = np.linalg.matrix_rank(ns1)
                if rank1 < len(ns1[0]):
                    if iprt: print('%6s # rank(ns1)=%d less than len(nzu)=%d' % ('None', rank1, len(ns1[0])))
                    maes.append(np.nan); rmses.append(np.nan)
                    continue
                if not np.all([si > 1 for si in ns1.shape]):
                    continue
                ys1i, ys2i, _ = krr.calc_ae_dressed(self.imcs[idx1][_idx1], ns1, ys1i, self.nzs[idx2], ys2i)
            n1so.append(len(_idx1))
            alphas = solve(mk1[_idx1][:, _idx1], ys1i, llambda, device)
            dys2i = np.dot(mk2[:, _idx1], alphas) - ys2i
            dys.append(dys2i)
            a = np.abs(dys2i)
            errmax = np.unique(dys2i[np.max(a) == a])[0]
            if n2 == 1:
                mae = dys2i[0]; rmse = abs(mae)
            else:
                mae = np.sum(a) / n2
                rmse = np.sqrt(np.sum(dys2i * dys2i) / n2)
            if iprt: print('%6d %12.4f %12.4f %12.4f' % (len(_idx1), mae, rmse, errmax))
            maes.append(mae); rmses.append(rmse); errsmax.append(errmax)
        self.dys, self.errsmax = dys, errsmax
        self.maes, self.rmses = np.array(maes), np.array(rmses)
        self.n1so = np.array(n1so, dtype=int)


class rkrr(object):
    """The reference's recursive / multi-fidelity KRR driver object (coreml/cml/algo/rkrr.py:18-267).  ``ys`` (nm, nl): one
    column per level of theory, cheapest first.  `run` trains levels 0 .. nl-2 on the nested prefixes ``n1s[::-1]`` (level 0
    learns y_0, level l learns y_l - y_{l-1}) and the last level for every size in ``n1s`` (NaN labels skipped); the
    prediction is the sum over levels plus the dressed-atom baselines.  Host NumPy as upstream; kernels come from the GPU."""

    def __init__(self, ys):
        self._ys = np.array(ys, dtype=np.float64)
        self.nm = len(ys)
        self.nl = len(ys[0])
        self.n2 = 0

    def init_m(self, obj):
        self._zs = np.asarray(obj.zs)
        self._zsu = np.unique(self._zs)
        self._nas = np.asarray(obj.nas)
        self._nsheav = np.asarray(obj.nsheav)
        self._coords = np.asarray(obj.coords)
        self._ias2 = np.cumsum(self._nas)
        self._ias1 = np.concatenate(([0], self._ias2[:-1])).astype(int)
        # rkrr.py:38-42 get_nzs: element counts per molecule
        self._nzs = np.array([[int((self._zs[b:e] == z).sum()) for z in self._zsu] for b, e in zip(self._ias1, self._ias2)])

    def get_idx(self, idx=-1, n1s=None, n1max=-1, seed1=1, namin=1, namax=7):
        """rkrr.py:45-158: idx = -n (the last n molecules are the targets), idx > 0 (seeded random split), or a list of
        training indices; namax = 0 (one training set or the given n1s), < 0 (all amons with N_I <= -namax), > 0 (one
        training set per heavy-atom count)."""
        n1s = [] if n1s is None else list(n1s)
        tidxs = np.arange(self.nm)
        if isinstance(idx, (int, np.integer)):
            if idx == 0:
                _idx1, idx2 = np.arange(self.nm), np.array([], dtype=int)
            elif idx > 0:
                np.random.seed(seed1)
                tidxs = np.random.permutation(self.nm)
                _idx1, idx2 = tidxs[:-idx], tidxs[-idx:]
                if n1max > 0:
                    _idx1 = _idx1[:n1max]
            else:
                _idx1, idx2 = tidxs[:idx], tidxs[idx:]
        elif isinstance(idx, (list, tuple, np.ndarray)):
            _idx1 = np.array(idx, dtype=int)
            idx2 = np.setdiff1d(tidxs, _idx1)
        else:
            raise TypeError('#ERROR: unsupported type of `idx')
        self.nsu_heav = np.unique(self._nsheav)
        idx1 = _idx1
        aml = True
        if namax == 0:
            aml = False
            if len(n1s) == 0:
                n1s = [len(_idx1)]
        elif namax < 0:
            idx1 = _idx1[np.logical_and(self._nsheav[_idx1] >= namin, self._nsheav[_idx1] <= -namax)]
            idx2 = np.setdiff1d(tidxs, idx1)
            n1s = [len(idx1)]
        else:
            n1s, idx1_sorted, cnt = [], [], 0
            t = self._nsheav[_idx1]
            for na in range(namin, namax + 1):
                if np.any(na == self.nsu_heav):
                    idx_i = _idx1[t == na]
                    idx1_sorted += list(idx_i)
                    cnt += len(idx_i)
                    n1s.append(cnt)
                else:
                    n1s.append(np.nan if cnt == 0 else cnt)
            idx1 = np.array(idx1_sorted, dtype=int)
        self.aml, self.n1s = aml, n1s
        self.n2 = len(idx2)
        tidx = np.concatenate((idx1, idx2)).astype(int)
        n1 = len(idx1)
        self.ys = self._ys[tidx]
        _nas = self._nas[tidx]
        self.coords = np.concatenate([self._coords[self._ias1[i]:self._ias2[i]] for i in tidx])
        self.zs = np.concatenate([self._zs[self._ias1[i]:self._ias2[i]] for i in tidx]).astype(int)
        self.nas1, self.nas2, self.nas = _nas[:n1], _nas[n1:], _nas
        self.nzs = self._nzs[tidx]
        self.nsheav = self._nsheav[tidx]
        self.nat1, self.nat2 = int(self.nas1.sum()), int(self.nas2.sum())
        self.zs1, self.zs2 = self.zs[:self.nat1], self.zs[self.nat1:]

    @staticmethod
    def calc_e_base(nzs1, ys1, nzs2):
        return e_base(nzs1, ys1, nzs2)  # rkrr.py:171-177

    def run(self, mks, usebl=True, llambda=1e-10, iprt=True, device=False):
        """rkrr.py:179-267.  Sets maes (signed error of the first target per training-set size) and n1s."""
        if isinstance(mks, str):
            dic = np.load(mks)
            mks = [dic['k1'].copy(), dic['k2'].copy()]
        mk1, mk2 = np.array(mks[0], dtype=np.float64), np.array(mks[1], dtype=np.float64)
        n2, nl = self.n2, self.nl
        idxs2 = np.arange(self.nm)[-n2:]
        sizes = [n for n in self.n1s if not (n is None or (isinstance(n, float) and np.isnan(n)))]
        n1sr = sizes[::-1]
        ys2p = np.zeros(n2)
        ns_nested = []
        for l in range(nl - 1):
            ims1 = np.arange(int(n1sr[l]))
            ns_nested.append(len(ims1))
            tgt = self.ys[ims1, l] if l == 0 else self.ys[ims1, l] - self.ys[ims1, l - 1]
            ys1, ys2_base = e_base(self.nzs[ims1], tgt, self.nzs[idxs2])
            alphas = solve(mk1[ims1][:, ims1], ys1, llambda, device)
            ys2p += np.dot(mk2[:, ims1], alphas) + ys2_base
        if iprt: print(' ** ns_nested = ', ns_nested)
        l = nl - 1
        maes, n1s = [], []
        for _n1 in self.n1s:
            if _n1 is None or (isinstance(_n1, float) and np.isnan(_n1)):
                maes.append(np.nan); n1s.append(np.nan)
                continue
            _ims1 = np.arange(int(_n1))
            ims1 = _ims1[np.logical_not(np.isnan(self.ys[_ims1, l]))]
            nzs1, nzs2 = self.nzs[ims1], self.nzs[idxs2]
            if len(ims1) == 0 or np.linalg.matrix_rank(nzs1) < len(nzs1[0]):
                if iprt: print('%6s' % 'None')
                n1s.append(np.nan); maes.append(np.nan)
                continue
            tgt = self.ys[ims1, l] - self.ys[ims1, l - 1] if nl > 1 else self.ys[ims1, l]
            ys1, ys2_base = e_base(nzs1, tgt, nzs2)
            alphas = solve(mk1[ims1][:, ims1], ys1, llambda, device)
            dys = ys2p + np.dot(mk2[:, ims1], alphas) + ys2_base - self.ys[idxs2, l]
            if iprt: print('%6d %.2f' % (len(ims1), dys[0]))
            maes.append(dys[0]); n1s.append(len(ims1))
        self.maes, self.n1s_run = maes, n1s
        self.ys2p_low = ys2p
