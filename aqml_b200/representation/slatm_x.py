"""Block / chunk driver for (a)SLATM kernels with the reference's job ids and on-disk files
(coreml/cml/representation/slatm_x.py:25-579).

The reference class ``slatm`` cuts K(train, train) and K(test, train) into chunks of ``use_chunks[1:]``
molecules (job ids ``11iiIIIJJJ``, ``11ijIIIJJJ``, ``12ijIIIJJJ``, slatm_x.py:148-153), each chunk again into
``navks`` sub-blocks (``c_kernel``, :500-538), and stores ``x_*.npz{x}``, ``k_*.npz{k}``, ``*_sub_III_JJJ.npz{k}``
and ``*_dsmax_*.txt`` so that independent processes can fill the matrix and a later run can resume.  The class
below keeps the constructor, the attribute names (``mk1``, ``mk2``, ``dsmax``, ``mbtypes``) and the file names /
keys, and runs every representation and kernel block on the GPU through the same C ABI as the rest of the
package.  Upstream the file does not import (``cml.representations``, ``qk``, ``qd`` are undefined names,
:14,116,415); where it is ambiguous the intended behaviour noted next to each method is implemented, and the
parity tests compare against the oracle composed the same way ("parity unpinned" in SURVEY 8c terms).

``obj`` is anything with ``zs`` (A,), ``coords`` (A,3), ``nas`` (N1+N2,), ``nas1``, ``nas2`` (training molecules
first; krr.py:60-100 in the reference) -- `MolSet` is a minimal one."""
from __future__ import annotations

import os
from math import ceil

import numpy as np

from .. import kernels as K
from .. import distance as Dm
from .core import generate_slatm_batch
from .slatm import mbtypes_from_counts

T, F = True, False


class MolSet(object):
    """Training molecules followed by test molecules, the fields `slatm` reads from the reference's ``krr`` object."""

    def __init__(self, zs, coords, nas, n1):
        self.zs = np.asarray(zs, dtype=int)
        self.coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
        self.nas = np.asarray(nas, dtype=int)
        self.nas1, self.nas2 = self.nas[:n1], self.nas[n1:]
        zsu = np.unique(self.zs)
        off = np.concatenate(([0], np.cumsum(self.nas)))
        self.nzs = np.array([[int((self.zs[off[i]:off[i + 1]] == z).sum()) for z in zsu] for i in range(len(self.nas))])


def all_jobs(nblk1, nblk2):
    """slatm_x.py:147-153: every chunk job, in the reference's order."""
    jobs = []
    for i in range(nblk1):
        jobs.append('11ii%03d%03d' % (i + 1, i + 1))
        for j in range(i + 1, nblk1):
            jobs.append('11ij%03d%03d' % (i + 1, j + 1))
        for j in range(nblk2):
            jobs.append('12ij%03d%03d' % (i + 1, j + 1))
    return jobs


def select_jobs(subjobs, nblk1, nblk2):
    """slatm_x.py:154-184 -> (jobs, write_kernel_files_only)."""
    jobsa = all_jobs(nblk1, nblk2)
    if isinstance(subjobs, str):
        return jobsa, F
    if not isinstance(subjobs, (tuple, list)):
        raise Exception('invalid input for `subjobs')
    if subjobs[0] in ['a', 'all']:
        return jobsa, F
    assert subjobs[0][0] in ['0', '1'], "#ERROR: `subjobs should be like ['11ii',...]"
    jobs = []
    for sj in subjobs:
        jobs += [j for j in jobsa if j[:4] == sj] if len(sj) == 4 else [sj]
    return jobs, T


class slatm(object):

    def __init__(self, obj, subjobs='a', navks=[-1, -1], nblksmax=[-1, -1], saveblk=T,
                 wxo=F, ck=F, wd=None, label='', param={}, itarget=F, use_chunks=[F],
                 icg=F, izeff=F, forbid_w=F, cab=F, verbose=F, device=None):
        """Arguments as slatm_x.py:27-29.  ``icg`` (connectivity graphs) is not supported: the compiled
        Fortran ignores ``cg`` as well (slatm.py:53,107 always pass all-ones).  ``device``: CUDA device of the
        blocks (default: current)."""
        if icg:
            raise NotImplementedError('icg=T needs aqml.cheminfo graphs; the built reference ignores cg (slatm.py:53,107)')
        self.cab, self.izeff, self.icg, self.verbose = cab, izeff, icg, verbose
        self.zmax = int(max(obj.zs))
        wd = os.environ.get('PWD', '.') if wd is None else wd
        self.wd = wd[:-1] if wd.endswith('/') else wd
        self.obj, self.wxo, self.navks, self.forbid_w, self.saveblk = obj, wxo, list(navks), forbid_w, saveblk
        self.itarget = itarget
        self.device = device
        n1, n2 = len(obj.nas1), len(obj.nas2)
        self.n1, self.n2 = n1, n2
        self.ias2 = np.cumsum(obj.nas)
        self.ias1 = np.concatenate(([0], self.ias2[:-1]))
        ims1, ims2 = np.arange(n1), np.arange(n1, n1 + n2)

        _param = {'local': True, 'nbody': 3, 'dgrids': [0.04, 0.04], 'widths': [0.05, 0.05],
                  'rcut': 4.8, 'alchemy': False, 'iBoA': True, 'rpower2': 6, 'coeffs': [1.],
                  'isqrt_grid': T, 'rpower3': 3, 'ws': [1., 1., 1.], 'pbc': '000',
                  'kernel': 'gaussian', 'saves': [F, F, F], 'reuses': [F, F, F]}
        for key in param:
            if key not in _param:
                raise KeyError(key)
            _param[key] = param[key]
        for i, key in enumerate(['savex', 'saved', 'savek']):
            _param[key] = _param['saves'][i]
        for i, key in enumerate(['reusex', 'reused', 'reusek']):
            _param[key] = _param['reuses'][i]
        if _param['alchemy'] or _param['pbc'] != '000':
            raise NotImplementedError('alchemy / pbc chunks: use generate_slatm(..., pbc=) per atom')
        self.label = label if label == '' else label + '_'
        self.param = _param
        self.coeffs = _param['coeffs']
        self.ncoeff = len(self.coeffs)
        self.pbc = _param['pbc']
        self.kernel = _param['kernel']
        self.local = _param['local']
        if self.kernel in ['linear'] and self.local:
            raise AssertionError('#ERROR: for linear kernel, consider using global repr only!')
        self.srep = 'aslatm' if self.local else 'slatm'
        self.use_chunks = use_chunks
        if self.kernel in ['gaussian', 'g']:
            self.dfunc = Dm.l2_distance
            self._coeffs = np.array(self.coeffs) / np.sqrt(2.0 * np.log(2.0))
        elif self.kernel in ['l', 'laplacian']:
            self.dfunc = Dm.manhattan_distance
            self._coeffs = np.array(self.coeffs) / np.log(2.0)
        elif self.kernel in ['linear']:
            self._coeffs = np.array(self.coeffs)
        else:
            raise Exception('#ERROR: not supported!')

        if not ck:
            return
        self.nav1, self.nav2 = n1, n2
        if use_chunks[0]:
            nav1, nav2 = use_chunks[1:]
            nblk1, nblk2 = int(ceil(n1 / nav1)), int(ceil(n2 / nav2))
            self.nav1, self.nav2 = nav1, nav2
            ks1 = np.zeros((self.ncoeff, n1, n1))
            ks2 = np.zeros((self.ncoeff, n2, n1))
            nblk1max, nblk2max = nblksmax
            if nblk1max <= 0: nblk1max = nblk1
            if nblk2max <= 0: nblk2max = nblk2
            jobs, wko = select_jobs(subjobs, nblk1, nblk2)
            self.jobs = jobs
            for i in range(min(nblk1, nblk1max)):  # kernel(train_i, train_i)
                if '11ii%03d%03d' % (i + 1, i + 1) in jobs:
                    idx, k11 = self.get_kernel_block(ims1, ims1, i, i)
                    if not wko and k11 is not None:
                        ib1, ie1 = idx
                        ks1[:, ib1:ie1, ib1:ie1] = k11
            for i in range(min(nblk1, nblk1max)):  # kernel(train_i, train_j)
                for j in range(i + 1, min(nblk1, nblk1max)):
                    if '11ij%03d%03d' % (i + 1, j + 1) in jobs:
                        idx, k21 = self.get_kernel_block(ims1, ims1, i, j)
                        if not wko and k21 is not None:
                            ib1, ie1, ib2, ie2 = idx
                            ks1[:, ib2:ie2, ib1:ie1] = k21
                            ks1[:, ib1:ie1, ib2:ie2] = np.transpose(k21, (0, 2, 1))
            for i in range(min(nblk1, nblk1max)):  # kernel(test_j, train_i)
                for j in range(min(nblk2, nblk2max)):
                    if '12ij%03d%03d' % (i + 1, j + 1) in jobs:
                        idx, k21 = self.get_kernel_block(ims1, ims2, i, j)
                        if not wko and k21 is not None:
                            ib1, ie1, ib2, ie2 = idx
                            ks2[:, ib2 - n1:ie2 - n1, ib1:ie1] = k21
        else:
            _, ks1 = self.get_kernel_block(ims1, ims1, 0, 0)
            _, ks2 = self.get_kernel_block(ims1, ims2, 0, 0)
        self.mk1, self.mk2 = ks1, ks2

    # ------------------------------------------------------------------------------------------------
    def get_x_block(self):
        """slatm_x.py:227-241.  The widths come from the FIRST block that is computed (``hasattr`` test, :239):
        every later block reuses them, so all chunks of one run share one sigma set."""
        self._ims_blk, self._label_x, self._aux_x = self._ims1, self._label_x1, self._aux_x1
        self.x1, self.zs1 = self.get_x()
        self.x2 = None
        if self.dist in ['ij', 'ji']:
            self._ims_blk, self._label_x, self._aux_x = self._ims2, self._label_x2, self._aux_x2
            self.x2, self.zs2 = self.get_x()
        if not hasattr(self, 'dsmax'):
            self.get_dsmax('c11')

    def get_kernel_block(self, ims1, ims2, i, j):
        """kernel(block_i, block_j), slatm_x.py:244-317.  Returns (idx, k) with k (ncoeff, rows of block j,
        rows of block i) -- (ncoeff, n_i, n_i) for a diagonal chunk."""
        nm1, nm2 = len(ims1), len(ims2)
        nav1 = self.nav1
        if nm1 == nm2 and np.all(ims1 == ims2):
            dist, dbi, nav2 = 'ii', '11', self.nav1
        else:
            dist, dbi, nav2 = 'ij', '21', self.nav2
        self.dist, self.dbi = dist, dbi
        ib1 = ims1[0] + i * nav1; ie1 = min(ims1[0] + (i + 1) * nav1, ims1[-1] + 1)
        ib2 = ims2[0] + j * nav2; ie2 = min(ims2[0] + (j + 1) * nav2, ims2[-1] + 1)
        _ims1, _ims2 = np.arange(ib1, ie1), np.arange(ib2, ie2)
        idx = [ib1, ie1, ib2, ie2]
        if dbi == '11':
            if i == j:
                _ims2 = np.arange(0)
                idx = [ib1, ie1]
            else:
                self.dist = 'ij'
        self._navk1, self._navk2 = ([self.navks[0]] * 2) if dbi == '11' else self.navks
        self._ims1, self._ims2 = _ims1, _ims2
        self._nas1, self._nas2 = self.obj.nas[_ims1], self.obj.nas[_ims2]
        self._zs1 = np.concatenate([self.obj.zs[self.ias1[im]:self.ias2[im]] for im in _ims1]).astype(int)
        self._zs2 = (np.concatenate([self.obj.zs[self.ias1[im]:self.ias2[im]] for im in _ims2]).astype(int)
                     if len(_ims2) else np.zeros(0, dtype=int))
        self._aux_x1 = ''
        self._label_x1 = '_c1_b%03d' % (i + 1)
        if self.dist in ['ij', 'ji']:
            self._aux_x2 = '' if dbi[0] == '1' else self.label
            self._label_x2 = '_c%s_b%03d' % (dbi[0], j + 1)
        _label_k = '_c%s_blk_%03d_%03d' % (self.dbi, i + 1, j + 1)
        _label_db = '' if dbi == '11' else self.label
        k = None
        if not self.wxo:
            self.fk = self.wd + '/%sk_%s_%s%s.npz' % (_label_db, self.kernel, self.srep, _label_k)
            if not (os.path.exists(self.fk) and self.param['reusek']):
                self.get_x_block()
            k = self.get_kernel()
        else:
            self.get_x_block()  # write-x-only runs still have to produce (and save) the chunk's x
        self.x1 = self.x2 = None
        return idx, k

    def get_slatm_mbtypes(self):
        """slatm_x.py:320-349 (same enumeration as slatm.py:286-331; element order ascending)."""
        zsmax = np.unique(self.obj.zs)
        nzsmax = np.max(self.obj.nzs, axis=0)
        self.mbtypes = mbtypes_from_counts([int(z) for z in zsmax], nzsmax)
        nsx = np.array([sum(len(mb) == n for mb in self.mbtypes) for n in (1, 2, 3)], dtype=int)
        self.nsx = nsx
        self.ins2 = np.cumsum(nsx)
        self.ins1 = np.array([0, self.ins2[0], self.ins2[1]], dtype=int)

    def _cuda(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('aqml_b200 needs a CUDA device (no CPU fallback)')
        return torch.device('cuda', torch.cuda.current_device()) if self.device is None else torch.device(self.device)

    def get_x(self):
        """(a)SLATM of the current chunk on the device, slatm_x.py:351-406.  Columns after the one-body slots are
        multiplied by 1/sqrt(dgrid) for the Gaussian kernel (``isqrt_grid``, :396-402)."""
        import torch
        widths, dgrids, rcut, rpower2, savex, reusex = [self.param[key] for key in
                                                        ['widths', 'dgrids', 'rcut', 'rpower2', 'savex', 'reusex']]
        fx = self.wd + '/%sx_%s%s.npz' % (self._aux_x, self.srep, self._label_x)
        if not hasattr(self, 'mbtypes'):
            self.get_slatm_mbtypes()
        ims = self._ims_blk
        zs = np.concatenate([self.obj.zs[self.ias1[i]:self.ias2[i]] for i in ims]).astype(int)
        dev = self._cuda()
        if os.path.exists(fx) and reusex:
            x = torch.as_tensor(np.load(fx)['x'], device=dev)
        else:
            coords = np.concatenate([self.obj.coords[self.ias1[i]:self.ias2[i]] for i in ims])
            x = generate_slatm_batch(torch.as_tensor(coords, device=dev), zs, self.obj.nas[ims], self.mbtypes,
                                     izeff=self.izeff, local=self.local, sigmas=widths, dgrids=dgrids, rcut=rcut,
                                     rpower=rpower2)
            if self.param['isqrt_grid'] and self.kernel in ['g', 'gaussian']:
                x[:, int(self.nsx[0]):] *= 1. / np.sqrt(dgrids[0])
            if savex:
                if self.forbid_w:
                    raise Exception("#ERROR: if it's safe to write x, reset forbid_w=T")
                np.savez(fx, x=x.cpu().numpy())
        return x, zs

    def get_dsmax(self, _label=None):
        """Largest representation distance per element (slatm_x.py:419-493), from the atoms of the current first
        block (of the target block if ``itarget`` and no chunks, :431-436).  The maximum is reduced on the device
        (no distance matrix is formed)."""
        saved, reused = self.param['saved'], self.param['reused']
        label = self.label if _label is None else _label
        fdmax = self.wd + '/%s_dsmax_%s_%s.txt' % (label, self.kernel, self.srep)
        if os.path.exists(fdmax) and reused:
            dso = np.atleast_1d(np.loadtxt(fdmax))
        else:
            if (not self.use_chunks[0]) and self.itarget and self.x2 is not None:
                x, zs = self.x2, self._zs2
            else:
                x, zs = self.x1, self._zs1
            p = 2 if self.kernel in ['g', 'gaussian', 'linear'] else 1
            if self.local:
                assert len(x) == len(zs), '#ERROR: `x and `zs shape mismatch'
                zsu = np.unique(zs)
                if self.cab and len(zsu) > 1:  # :447-455: the two largest Z's
                    f1, f2 = self._mask(zs == zsu[-2], x), self._mask(zs == zsu[-1], x)
                    dsmax = np.array([Dm.max_distance(x[f1], x[f2], p)] * len(zsu))
                else:
                    dsmax = np.array([Dm.max_distance(x[self._mask(zs == z, x)], x[self._mask(zs == z, x)], p) for z in zsu])
                dso = np.ones(self.zmax) * 1.e-9
                for i, zi in enumerate(zsu):
                    dso[zi - 1] = dsmax[i]
            else:
                dso = np.array([Dm.max_distance(x, x, p)])
            if saved:
                np.savetxt(fdmax, dso, '%.8f')
        self.dsmax = dso

    @staticmethod
    def _mask(m, x):
        import torch
        return torch.as_tensor(m, device=x.device)

    def c_imap(self, nas):
        ias2 = np.cumsum(nas)
        return np.concatenate(([0], ias2[:-1])), ias2

    def kfunc(self, x1, x2, zs1, zs2, nas1, nas2, sigmas):
        if self.kernel in ['g', 'gaussian']:
            return K.get_local_kernels_gaussian(x1, x2, zs1, zs2, nas1, nas2, self.cab, self.zmax, sigmas)
        # the Laplacian local kernel has one width per coefficient and no element matching (fkernels.f90:149-173);
        # the reference passes the (ncoeff, zmax) table and would fail in f2py -- use the largest element's width
        return K.get_local_kernels_laplacian(x1, x2, nas1, nas2, np.asarray(sigmas).max(axis=1))

    def c_kernel(self, zs1, zs2, x1, x2, navk1, navk2, nas1, nas2, sigmas):
        """One chunk in navk1 x navk2 molecule sub-blocks with per-sub-block files (slatm_x.py:500-538)."""
        nm1, nm2 = len(nas1), len(nas2)
        if navk1 <= 0: navk1 = nm1
        if navk2 <= 0: navk2 = nm2
        nk1, nk2 = int(ceil(nm1 / navk1)), int(ceil(nm2 / navk2))
        k = np.zeros((self.ncoeff, nm1, nm2))
        iasb, iase = self.c_imap(nas1)
        jasb, jase = self.c_imap(nas2)
        for i in range(nk1):
            for j in range(nk2):
                imb, ime = i * navk1, min((i + 1) * navk1, nm1)
                jmb, jme = j * navk2, min((j + 1) * navk2, nm2)
                iab, iae = iasb[imb], iase[ime - 1]
                jab, jae = jasb[jmb], jase[jme - 1]
                fij = self.fk[:-4] + '_sub_%03d_%03d.npz' % (i + 1, j + 1)
                if os.path.exists(fij):
                    k[:, imb:ime, jmb:jme] = np.load(fij)['k']
                else:
                    ktmp = self.kfunc(x1[iab:iae], x2[jab:jae], zs1[iab:iae], zs2[jab:jae],
                                      nas1[imb:ime], nas2[jmb:jme], sigmas).cpu().numpy()
                    k[:, imb:ime, jmb:jme] = ktmp
                    if self.saveblk and not self.forbid_w:
                        np.savez(fij, k=ktmp)
        return k

    def get_kernel(self):
        """slatm_x.py:540-579."""
        savek, reusek = self.param['savek'], self.param['reusek']
        fk = self.fk
        if os.path.isfile(fk) and reusek:
            return np.load(fk)['k']
        x1, x2 = self.x1, self.x2
        if self.kernel in ['gaussian', 'g', 'laplacian', 'l']:
            sigmas = self.dsmax * self._coeffs[..., np.newaxis]
            if self.local:
                if self.dist == 'ii':
                    k = self.c_kernel(self.zs1, self.zs1, x1, x1, self._navk1, self._navk1, self._nas1, self._nas1, sigmas)
                else:
                    k = self.c_kernel(self.zs2, self.zs1, x2, x1, self._navk2, self._navk1, self._nas2, self._nas1, sigmas)
            else:
                kf = K.gaussian_kernels if self.kernel[0] == 'g' else K.laplacian_kernels
                k = kf(x1 if self.dist == 'ii' else x2, x1, sigmas[:, 0]).cpu().numpy()
        else:
            k = K.linear_kernel(x1 if self.dist == 'ii' else x2, x1).cpu().numpy()
        if savek:
            if self.forbid_w:
                raise Exception("#ERROR: if it's safe to write k, reset forbid_w=T")
            np.savez(fk, k=k)
        return k
