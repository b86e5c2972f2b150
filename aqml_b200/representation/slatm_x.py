"""Chunked kernel assembly with the reference's job ids and on-disk files (the CONTRACT of
coreml/cml/representation/slatm_x.py:129-223,242-317,500-579) on top of the device-resident path.

What is kept from the reference -- because other processes / later runs / reference-produced files depend on it:
  * job ids   ``11iiIIIIII`` (train chunk I x itself), ``11ijIIIJJJ`` (train I x train J, J > I), ``12ijIIIJJJ``
              (train I x test J), in that order (slatm_x.py:147-153); a list of ids selects work for one worker,
              which then only writes kernel files (:163-171);
  * files     ``[label_]x_<srep>_c<1|2>_bNNN.npz{x}``, ``[label_]k_<kernel>_<srep>_c<11|21>_blk_III_JJJ.npz{k}``,
              ``..._sub_III_JJJ.npz{k}`` per ``navks`` sub-block, ``<label>_dsmax_<kernel>_<srep>.txt`` (:308,357,425,522);
  * numbers   x is stored with the columns after the one-body slots times 1/sqrt(dgrid) (:396-402), the widths come from
              the first chunk that is computed (:239,431-436,465-476), K(test chunk, train chunk) is stored as
              (ncoeff, test rows, train rows).

How it is computed is not the reference's: a run is a list of `Job`s; every chunk of molecules is packed ONCE into the
element-grouped operand the tensor-core contraction reads (`fused.PackedRep`, cached per chunk), a job is one fused
contraction on the device, and the dense (atoms, D) matrix exists only when an ``x_*.npz`` file is asked for or read.
(The 1/sqrt(dgrid) scale multiplies every column that takes part in an element-matched distance, so K -- whose widths
are proportional to the largest such distance -- does not depend on it; only the stored x and dsmax do.)  Global SLATM,
Laplacian and linear kernels have no packed form and go through the dense device kernels.

Upstream the file does not import (``cml.representations``, ``qk``, ``qd`` are undefined names, :14,116,415): the
parity tests compare with the oracle composed as documented above ("parity unpinned" in SURVEY 8c terms).
"""
from __future__ import annotations

import collections
import os
from math import ceil

import numpy as np

from .. import distance as Dm
from .. import fused
from .. import kernels as K
from .core import generate_slatm_batch
from .slatm import mbtypes_from_counts

T, F = True, False

Job = collections.namedtuple("Job", "id kind i j")   # kind '11ii' | '11ij' | '12ij'; 0-based chunk numbers


class MolSet(object):
    """Training molecules followed by test molecules: zs (A,), coords (A,3), nas (N1+N2,), nas1 / nas2, nzs."""

    def __init__(self, zs, coords, nas, n1):
        self.zs = np.asarray(zs, dtype=int)
        self.coords = np.asarray(coords, dtype=np.float64).reshape(-1, 3)
        self.nas = np.asarray(nas, dtype=int)
        self.nas1, self.nas2 = self.nas[:n1], self.nas[n1:]
        zsu = np.unique(self.zs)
        off = np.concatenate(([0], np.cumsum(self.nas)))
        self.nzs = np.array([[int((self.zs[off[i]:off[i + 1]] == z).sum()) for z in zsu] for i in range(len(self.nas))])


def all_jobs(nblk1, nblk2):
    """Every chunk job id in the reference's order (slatm_x.py:147-153)."""
    return [j.id for j in plan_jobs(nblk1, nblk2)]


def plan_jobs(nblk1, nblk2):
    jobs = []
    for i in range(nblk1):
        jobs.append(Job('11ii%03d%03d' % (i + 1, i + 1), '11ii', i, i))
        jobs += [Job('11ij%03d%03d' % (i + 1, j + 1), '11ij', i, j) for j in range(i + 1, nblk1)]
        jobs += [Job('12ij%03d%03d' % (i + 1, j + 1), '12ij', i, j) for j in range(nblk2)]
    return jobs


def select_jobs(subjobs, nblk1, nblk2):
    """ids to run and "this is a worker: write kernel files only" (slatm_x.py:154-184).  'a' / ['all']: everything;
    otherwise a list of 4-character kinds ('11ii': all of that kind) and / or full ids."""
    ids = all_jobs(nblk1, nblk2)
    if isinstance(subjobs, str) or (isinstance(subjobs, (tuple, list)) and subjobs and subjobs[0] in ('a', 'all')):
        return ids, F
    if not isinstance(subjobs, (tuple, list)) or not subjobs:
        raise Exception('invalid input for `subjobs')
    assert subjobs[0][0] in ['0', '1'], "#ERROR: `subjobs should be like ['11ii',...]"
    picked = []
    for sj in subjobs:
        picked += [i for i in ids if i[:4] == sj] if len(sj) == 4 else [sj]
    return picked, T


DEFAULTS = {'local': True, 'nbody': 3, 'dgrids': [0.04, 0.04], 'widths': [0.05, 0.05], 'rcut': 4.8, 'alchemy': False,
            'iBoA': True, 'rpower2': 6, 'coeffs': [1.], 'isqrt_grid': T, 'rpower3': 3, 'ws': [1., 1., 1.], 'pbc': '000',
            'kernel': 'gaussian', 'saves': [F, F, F], 'reuses': [F, F, F]}


class slatm(object):
    """Constructor arguments as slatm_x.py:27-29 (``obj``: anything with zs, coords, nas, nas1, nas2, nzs -- `MolSet`).
    ``ck=True`` runs the selected jobs; results: ``mk1`` (ncoeff, N1, N1), ``mk2`` (ncoeff, N2, N1), ``dsmax``,
    ``mbtypes``, ``jobs``."""

    def __init__(self, obj, subjobs='a', navks=[-1, -1], nblksmax=[-1, -1], saveblk=T, wxo=F, ck=F, wd=None, label='',
                 param={}, itarget=F, use_chunks=[F], icg=F, izeff=F, forbid_w=F, cab=F, verbose=F, device=None):
        if icg:
            raise NotImplementedError('icg=T needs aqml.cheminfo graphs; the built reference ignores cg (slatm.py:53,107)')
        p = dict(DEFAULTS)
        for key in param:
            if key not in p:
                raise KeyError(key)
            p[key] = param[key]
        if p['alchemy'] or p['pbc'] != '000':
            raise NotImplementedError('alchemy / pbc chunks: use generate_slatm(..., pbc=) per atom')
        self.param, self.obj = p, obj
        self.save = dict(zip('xdk', p['saves']))
        self.reuse = dict(zip('xdk', p['reuses']))
        self.kernel = {'g': 'gaussian', 'l': 'laplacian'}.get(p['kernel'], p['kernel'])
        if self.kernel not in ('gaussian', 'laplacian', 'linear'):
            raise Exception('#ERROR: not supported!')
        self.local = bool(p['local'])
        if self.kernel == 'linear' and self.local:
            raise AssertionError('#ERROR: for linear kernel, consider using global repr only!')
        self.srep = 'aslatm' if self.local else 'slatm'
        self.coeffs = np.asarray(p['coeffs'], dtype=float)
        self.ncoeff = len(self.coeffs)
        self.width_factor = {'gaussian': 1 / np.sqrt(2 * np.log(2)), 'laplacian': 1 / np.log(2), 'linear': 1.0}[self.kernel]
        self.cab, self.izeff, self.itarget, self.device = cab, izeff, itarget, device
        self.wxo, self.saveblk, self.forbid_w = wxo, saveblk, forbid_w
        wd = os.environ.get('PWD', '.') if wd is None else wd
        self.wd = wd.rstrip('/') or '/'
        self.label = label + '_' if label else ''
        self.zmax = int(np.max(obj.zs))
        self.n1, self.n2 = len(obj.nas1), len(obj.nas2)
        self.off = np.concatenate(([0], np.cumsum(obj.nas)))
        chunked = bool(use_chunks[0])
        self.chunked = chunked
        self.nav1, self.nav2 = (int(use_chunks[1]), int(use_chunks[2])) if chunked else (self.n1, max(self.n2, 1))
        self.nblk1, self.nblk2 = int(ceil(self.n1 / self.nav1)), int(ceil(self.n2 / self.nav2))
        self.navks = [int(v) for v in navks]
        self._reps = collections.OrderedDict()   # (set, chunk) -> PackedRep, a few most recent
        zsu = [int(z) for z in np.unique(obj.zs)]
        self.mbtypes = mbtypes_from_counts(zsu, np.max(obj.nzs, axis=0))     # slatm_x.py:320-349 == slatm.py:286-331
        self.n1body = sum(len(m) == 1 for m in self.mbtypes)
        if not ck:
            return
        ids, worker = select_jobs(subjobs, self.nblk1, self.nblk2) if chunked else (all_jobs(1, 1 if self.n2 else 0), F)
        self.jobs = ids
        lim1 = self.nblk1 if nblksmax[0] <= 0 else min(self.nblk1, nblksmax[0])
        lim2 = self.nblk2 if nblksmax[1] <= 0 else min(self.nblk2, nblksmax[1])
        todo = [j for j in plan_jobs(self.nblk1, self.nblk2) if j.id in set(ids) and j.i < lim1
                and (j.j < lim2 if j.kind == '12ij' else j.j < lim1)]
        # the reference runs all diagonal jobs first, then 11ij, then 12ij (slatm_x.py:186-217): the widths come from the
        # first chunk that is touched
        todo.sort(key=lambda j: ('11ii', '11ij', '12ij').index(j.kind))
        self.mk1 = np.zeros((self.ncoeff, self.n1, self.n1))
        self.mk2 = np.zeros((self.ncoeff, self.n2, self.n1))
        try:
            for job in todo:
                k = self.run_job(job)
                if worker or k is None:
                    continue
                r1, r2 = self.rows(1, job.i), self.rows(1 if job.kind[1] == '1' else 2, job.j)
                if job.kind == '11ii':
                    self.mk1[:, r1, r1] = k
                elif job.kind == '11ij':
                    self.mk1[:, r2, r1] = k
                    self.mk1[:, r1, r2] = np.transpose(k, (0, 2, 1))
                else:
                    self.mk2[:, r2, r1] = k
        finally:
            self.close()

    # ---- chunks ---------------------------------------------------------------------------------------------------
    def rows(self, which, c):
        """Molecule rows of chunk c of the training (1) / test (2) set inside mk1 / mk2."""
        nav, n = (self.nav1, self.n1) if which == 1 else (self.nav2, self.n2)
        return slice(c * nav, min((c + 1) * nav, n))

    def molecules(self, which, c):
        """(nas, zs, coords) of a chunk."""
        r = self.rows(which, c)
        lo, hi = (r.start, r.stop) if which == 1 else (self.n1 + r.start, self.n1 + r.stop)
        a, b = int(self.off[lo]), int(self.off[hi])
        return self.obj.nas[lo:hi], np.asarray(self.obj.zs[a:b], dtype=int), self.obj.coords[a:b]

    def _cuda(self):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('aqml_b200 needs a CUDA device (no CPU fallback)')
        return torch.device('cuda', torch.cuda.current_device()) if self.device is None else torch.device(self.device)

    def _slatm_kw(self):
        p = self.param
        return dict(izeff=self.izeff, sigmas=list(p['widths']), dgrids=list(p['dgrids']), rcut=p['rcut'], rpower=p['rpower2'])

    def fused_path(self):
        """Local Gaussian kernels without `cab` are contracted from packed operands; everything else from dense rows."""
        return self.local and self.kernel == 'gaussian' and not self.cab

    def rep(self, which, c):
        key = (which, c)
        if key in self._reps:
            self._reps.move_to_end(key)
            return self._reps[key]
        import torch
        nas, zs, coords = self.molecules(which, c)
        with torch.cuda.device(self._cuda()):
            r = fused.PackedRep(torch.as_tensor(coords, device=self._cuda()), zs, nas, self.mbtypes, **self._slatm_kw())
        self._reps[key] = r
        while len(self._reps) > 4:
            self._reps.popitem(last=False)[1].close()
        return r

    def close(self):
        for r in self._reps.values():
            r.close()
        self._reps.clear()

    def x_file(self, which, c):
        return '%s/%sx_%s_c%d_b%03d.npz' % (self.wd, self.label if which == 2 else '', self.srep, which, c + 1)

    def dense_x(self, which, c):
        """The chunk's (a)SLATM rows as the reference stores them (device tensor), from its file when ``reusex``."""
        import torch
        fx = self.x_file(which, c)
        if os.path.exists(fx) and self.reuse['x']:
            return torch.as_tensor(np.load(fx)['x'], device=self._cuda())
        nas, zs, coords = self.molecules(which, c)
        x = generate_slatm_batch(torch.as_tensor(coords, device=self._cuda()), zs, nas, self.mbtypes, local=self.local,
                                 **self._slatm_kw())
        if self.param['isqrt_grid'] and self.kernel == 'gaussian':
            x[:, self.n1body:] *= 1. / np.sqrt(self.param['dgrids'][0])
        if self.save['x']:
            if self.forbid_w:
                raise Exception("#ERROR: if it's safe to write x, reset forbid_w=T")
            np.savez(fx, x=x.cpu().numpy())
        return x

    # ---- widths ---------------------------------------------------------------------------------------------------
    def widths(self, which, c, x=None):
        """dsmax: the largest representation distance per element among the atoms of ONE chunk -- the first one that is
        computed; every later job of the run reuses it (slatm_x.py:239,419-493).  Reduced on the device."""
        if hasattr(self, 'dsmax'):
            return self.dsmax
        fd = '%s/c11_dsmax_%s_%s.txt' % (self.wd, self.kernel, self.srep)
        if os.path.exists(fd) and self.reuse['d']:
            self.dsmax = np.atleast_1d(np.loadtxt(fd))
            return self.dsmax
        scale = 1. / np.sqrt(self.param['dgrids'][0]) if (self.param['isqrt_grid'] and self.kernel == 'gaussian') else 1.
        if self.fused_path() and x is None:
            r = self.rep(which, c)
            dso = fused.kernel_width(r, r, self.zmax)
            dso = np.where(dso > 1.e-9, dso * scale, dso)          # absent elements keep the 1e-9 placeholder
        else:
            x = self.dense_x(which, c) if x is None else x
            p = 1 if self.kernel == 'laplacian' else 2
            if self.local:
                import torch
                zs = self.molecules(which, c)[1]
                zsu = np.unique(zs)
                pick = lambda z: x[torch.as_tensor(zs == z, device=x.device)]  # noqa: E731
                if self.cab and len(zsu) > 1:                       # one width from the two largest Z's (:447-455)
                    vals = [Dm.max_distance(pick(zsu[-2]), pick(zsu[-1]), p)] * len(zsu)
                else:
                    vals = [Dm.max_distance(pick(z), pick(z), p) for z in zsu]
                dso = np.ones(self.zmax) * 1.e-9
                dso[zsu - 1] = vals
            else:
                dso = np.array([Dm.max_distance(x, x, p)])
        if self.save['d']:
            np.savetxt(fd, dso, '%.8f')
        self.dsmax = dso
        return dso

    # ---- one job --------------------------------------------------------------------------------------------------
    def k_file(self, job):
        c = '11' if job.kind[1] == '1' else '21'
        return '%s/%sk_%s_%s_c%s_blk_%03d_%03d.npz' % (self.wd, self.label if c == '21' else '', self.kernel, self.srep, c,
                                                      job.i + 1, job.j + 1)

    def run_job(self, job):
        """K(chunk j, chunk i) as (ncoeff, rows of j, rows of i); diagonal jobs (ncoeff, n_i, n_i).  None for a
        write-x-only run."""
        w2 = 1 if job.kind[1] == '1' else 2
        first = (2, job.j) if (not self.chunked and self.itarget and w2 == 2) else (1, job.i)   # :431-436
        if self.wxo:
            self.dense_x(1, job.i)
            if job.kind != '11ii':
                self.dense_x(w2, job.j)
            return None
        fk = self.k_file(job)
        if os.path.isfile(fk) and self.reuse['k']:
            return np.load(fk)['k']
        need_x = self.save['x'] or self.reuse['x'] or not self.fused_path()
        x1 = self.dense_x(1, job.i) if need_x else None
        x2 = x1 if job.kind == '11ii' else (self.dense_x(w2, job.j) if need_x else None)
        dso = self.widths(*first, x=(x2 if first[0] == 2 else x1) if not self.fused_path() else None)
        nas1, zs1, _ = self.molecules(1, job.i)
        nas2, zs2, _ = self.molecules(w2, job.j)
        subfile = lambda a, b: fk[:-4] + '_sub_%03d_%03d.npz' % (a + 1, b + 1)  # noqa: E731
        nk = [(1 if v <= 0 else int(ceil(n / v)), n if v <= 0 else v)
              for v, n in zip(([self.navks[0]] * 2 if w2 == 1 else self.navks[::-1]), (len(nas2), len(nas1)))]
        (nk2, av2), (nk1, av1) = nk
        subs = [(a, b) for a in range(nk2) for b in range(nk1)]
        if all(os.path.exists(subfile(a, b)) for a, b in subs):     # every sub-block is there (this or another worker wrote them)
            k = np.zeros((self.ncoeff, len(nas2), len(nas1)))
            for a, b in subs:
                k[:, a * av2:(a + 1) * av2, b * av1:(b + 1) * av1] = np.load(subfile(a, b))['k']
        else:
            k = self.contract(job, w2, dso, x1, x2, nas1, zs1, nas2, zs2)
            if self.saveblk and not self.forbid_w:
                for a, b in subs:
                    if not os.path.exists(subfile(a, b)):
                        np.savez(subfile(a, b), k=k[:, a * av2:(a + 1) * av2, b * av1:(b + 1) * av1])
        if self.save['k']:
            if self.forbid_w:
                raise Exception("#ERROR: if it's safe to write k, reset forbid_w=T")
            np.savez(fk, k=k)
        return k

    def contract(self, job, w2, dso, x1, x2, nas1, zs1, nas2, zs2):
        """One chunk pair on the device."""
        sig = self.coeffs[:, None] * self.width_factor * np.asarray(dso)[None, :]
        if self.fused_path():
            scale = np.sqrt(self.param['dgrids'][0]) if self.param['isqrt_grid'] else 1.
            sig_packed = np.where(sig > 1.e-8, sig * scale, sig)      # back to the unscaled operand's units
            r1 = self.rep(1, job.i)
            if job.kind == '11ii':
                return fused.local_gaussian(r1, r1, sig_packed, self.zmax, symmetric=True).cpu().numpy()
            return fused.local_gaussian(self.rep(w2, job.j), r1, sig_packed, self.zmax).cpu().numpy()
        if self.kernel == 'linear':
            return K.linear_kernel(x2, x1).cpu().numpy()[None]
        if not self.local:
            kf = K.gaussian_kernels if self.kernel == 'gaussian' else K.laplacian_kernels
            return kf(x2, x1, sig[:, 0]).cpu().numpy()
        if self.kernel == 'gaussian':
            return K.get_local_kernels_gaussian(x2, x1, zs2, zs1, nas2, nas1, self.cab, self.zmax, sig).cpu().numpy()
        # the Laplacian local kernel has one width per coefficient and no element matching (fkernels.f90:149-173); the
        # reference passes the (ncoeff, zmax) table and would fail in f2py -- the largest element's width is used
        return K.get_local_kernels_laplacian(x2, x1, nas2, nas1, sig.max(axis=1)).cpu().numpy()
