"""GPU: the block / chunk driver (slatm_x.py job ids, sub-blocks, files) against the oracle composed the same way."""
import os

import numpy as np
import pytest

from conftest import kernel_rel_err, rel_err
from aqml_b200 import synth
from aqml_b200.representation.slatm_x import MolSet, slatm
from oracle import c_oracle as co
from oracle import np_oracle as o

pytestmark = pytest.mark.gpu

N1, N2, NAV1, NAV2 = 22, 7, 8, 4
PARAM = {'dgrids': [0.08, 0.08], 'coeffs': [0.5, 1.0, 2.0]}


def _expected(nas, zs, coords, first_block=NAV1):
    """What slatm_x.py computes: x scaled by 1/sqrt(dgrid) after the one-body slots (:396-402), widths from the
    atoms of the first chunk (:239,431-436,465-476), K by calc_lk_gaussian."""
    mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
    x = co.generate_slatm_batch(coords, zs, nas, mb, True, sigmas=(.05, .05), dgrids=(.08, .08), rcut=4.8)
    n1body = sum(len(m) == 1 for m in mb)
    x[:, n1body:] *= 1.0 / np.sqrt(0.08)
    a = int(nas[:first_block].sum())
    zmax = int(zs.max())
    dso = np.full(zmax, 1e-9)
    for z in np.unique(zs[:a]):
        xs = x[:a][zs[:a] == z]
        dso[z - 1] = co.l2_distance(xs, xs).max()
    sig = np.array(PARAM['coeffs'])[:, None] / np.sqrt(2 * np.log(2)) * dso
    A1 = int(nas[:N1].sum())
    k1 = co.calc_lk_gaussian(x[:A1], x[:A1], zs[:A1], zs[:A1], nas[:N1], nas[:N1], zmax, False, sig)
    k2 = co.calc_lk_gaussian(x[A1:], x[:A1], zs[A1:], zs[:A1], nas[N1:], nas[:N1], zmax, False, sig)
    return x, dso, k1, k2


@pytest.fixture(scope="module")
def mols():
    nas, zs, coords = synth.qm9_like(N1 + N2, 70, max_heavy=5)  # seed: no element occurs exactly once in a first chunk (dmax 0)
    return nas, zs, coords


def test_chunked_driver_matches_oracle_and_writes_reference_files(mols, tmp_path):
    nas, zs, coords = mols
    x, dso, k1, k2 = _expected(nas, zs, coords)
    obj = MolSet(zs, coords, nas, N1)
    wd = str(tmp_path)
    run = slatm(obj, ck=True, wd=wd, label='t', param=dict(PARAM, saves=[True, True, True]), navks=[3, 2],
                use_chunks=[True, NAV1, NAV2])
    assert rel_err(run.dsmax, dso) < 1e-10
    assert run.mk1.shape == k1.shape and run.mk2.shape == k2.shape
    assert kernel_rel_err(run.mk1, k1) < 1e-10
    assert kernel_rel_err(run.mk2, k2) < 1e-10
    files = set(os.listdir(wd))
    # names and keys of slatm_x.py:308,357,425,522
    for f in ('x_aslatm_c1_b001.npz', 'x_aslatm_c1_b003.npz', 't_x_aslatm_c2_b002.npz', 'k_gaussian_aslatm_c11_blk_001_001.npz',
              'k_gaussian_aslatm_c11_blk_001_003.npz', 't_k_gaussian_aslatm_c21_blk_003_002.npz',
              'k_gaussian_aslatm_c11_blk_001_001_sub_003_003.npz', 't_k_gaussian_aslatm_c21_blk_001_001_sub_002_003.npz',
              'c11_dsmax_gaussian_aslatm.txt'):
        assert f in files, f
    a = int(nas[:NAV1].sum())
    assert rel_err(np.load(os.path.join(wd, 'x_aslatm_c1_b001.npz'))['x'], x[:a]) < 1e-10
    kb = np.load(os.path.join(wd, 't_k_gaussian_aslatm_c21_blk_003_002.npz'))['k']
    assert kernel_rel_err(kb, k2[:, NAV2:2 * NAV2, 2 * NAV1:N1]) < 1e-10
    # resume: every x / k / dsmax is read back, nothing is recomputed (poison one file to prove it is used)
    np.savez(os.path.join(wd, 'k_gaussian_aslatm_c11_blk_002_002.npz'), k=np.full((3, NAV1, NAV1), 7.0))
    again = slatm(obj, ck=True, wd=wd, label='t', param=dict(PARAM, reuses=[True, True, True]), navks=[3, 2],
                  use_chunks=[True, NAV1, NAV2])
    assert np.all(again.mk1[:, NAV1:2 * NAV1, NAV1:2 * NAV1] == 7.0)
    again.mk1[:, NAV1:2 * NAV1, NAV1:2 * NAV1] = run.mk1[:, NAV1:2 * NAV1, NAV1:2 * NAV1]
    assert np.array_equal(again.mk1, run.mk1) and np.array_equal(again.mk2, run.mk2)


def test_subjobs_fill_only_their_blocks(mols, tmp_path):
    nas, zs, coords = mols
    _, _, k1, k2 = _expected(nas, zs, coords)
    obj = MolSet(zs, coords, nas, N1)
    # a worker process that is given two job ids writes kernel files only (slatm_x.py:163-171)
    part = slatm(obj, subjobs=['11ij001002', '12ij002001'], ck=True, wd=str(tmp_path), label='t', saveblk=False,
                 param=dict(PARAM, saves=[False, True, True]), use_chunks=[True, NAV1, NAV2])
    assert not part.mk1.any() and not part.mk2.any()
    kb = np.load(os.path.join(str(tmp_path), 'k_gaussian_aslatm_c11_blk_001_002.npz'))['k']
    assert kernel_rel_err(kb, k1[:, NAV1:2 * NAV1, :NAV1]) < 1e-10
    # this worker derived the widths from ITS first chunk = chunk 1 here, so the numbers agree with the full run;
    kb = np.load(os.path.join(str(tmp_path), 't_k_gaussian_aslatm_c21_blk_002_001.npz'))['k']
    # (second job's first chunk is train chunk 2, but dsmax was already fixed by the first job)
    assert kernel_rel_err(kb, k2[:, :NAV2, NAV1:2 * NAV1]) < 1e-10


def test_unchunked_driver(mols, tmp_path):
    nas, zs, coords = mols
    _, dso, k1, k2 = _expected(nas, zs, coords, first_block=N1)
    run = slatm(MolSet(zs, coords, nas, N1), ck=True, wd=str(tmp_path), param=PARAM)
    assert rel_err(run.dsmax, dso) < 1e-10
    assert kernel_rel_err(run.mk1, k1) < 1e-10 and kernel_rel_err(run.mk2, k2) < 1e-10


def test_global_slatm_chunks(mols, tmp_path):
    nas, zs, coords = mols
    mb = o.get_slatm_mbtypes(synth.split_zs(nas, zs))
    x = co.generate_slatm_batch(coords, zs, nas, mb, False, sigmas=(.05, .05), dgrids=(.08, .08), rcut=4.8)
    x[:, sum(len(m) == 1 for m in mb):] *= 1.0 / np.sqrt(0.08)
    dmax = co.l2_distance(x[:NAV1], x[:NAV1]).max()
    run = slatm(MolSet(zs, coords, nas, N1), ck=True, wd=str(tmp_path), param=dict(PARAM, local=False),
                use_chunks=[True, NAV1, NAV2])
    assert rel_err(run.dsmax, np.array([dmax])) < 1e-10
    for k, c in enumerate(PARAM['coeffs']):
        s = c * dmax / np.sqrt(2 * np.log(2))
        assert kernel_rel_err(run.mk1[k], o.gaussian_kernel(x[:N1], x[:N1], s)) < 1e-10
        assert kernel_rel_err(run.mk2[k], o.gaussian_kernel(x[N1:], x[:N1], s)) < 1e-10
