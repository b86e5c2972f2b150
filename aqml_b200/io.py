"""On-disk interop with files the reference writes (SURVEY 5 checkpoint/resume, 8f row 3).

``aqml -savex -savek`` stores ``data/<train>-x1.npz {x1}``, ``data/<test>-x2.npz {x2}`` and
``data/<train>-kernels[-l].npz {ks1, ks2}`` (algo/aqml.py:626-665, 719, 736, 795); the block driver
stores ``x_*.npz {x}`` and ``k_*.npz {k}`` (representation/slatm_x.py:308,357,522-533).  These helpers
use exactly those file names and keys so intermediate results can be exchanged in both directions."""
from __future__ import annotations

import os

import numpy as np


def stem(folder, rcut=4.8, z=None):
    """'data/' + the folder path with '/' -> '-', plus the reference's suffixes (aqml.py:630-639)."""
    fn = folder[:-1] if folder.endswith('/') else folder
    fn = 'data/' + '-'.join(fn.split('/'))
    if rcut != 4.8:
        fn += '-rc%.1f' % rcut
    if z is not None:
        fn += '-z%d' % z
    return fn


def _host(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


def save_x(path, x, key):
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    np.savez(path, **{key: _host(x)})


def save_x1(train_folder, x1, rcut=4.8, z=None):
    save_x(stem(train_folder, rcut, z) + '-x1.npz', x1, 'x1')   # aqml.py:719


def save_x2(test_folder, x2, rcut=4.8, z=None):
    save_x(stem(test_folder, rcut, z) + '-x2.npz', x2, 'x2')    # aqml.py:736


def load_x1(train_folder, rcut=4.8, z=None):
    return np.load(stem(train_folder, rcut, z) + '-x1.npz')['x1']


def load_x2(test_folder, rcut=4.8, z=None):
    return np.load(stem(test_folder, rcut, z) + '-x2.npz')['x2']


def kernels_file(train_folder, kernel='g', rcut=4.8, z=None):
    return stem(train_folder, rcut, z) + '-kernels%s.npz' % ({'g': '', 'l': '-l'}[kernel[0]])  # aqml.py:663


def save_kernels(train_folder, ks1, ks2, kernel='g', rcut=4.8, z=None):
    path = kernels_file(train_folder, kernel, rcut, z)
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    np.savez(path, ks1=_host(ks1), ks2=_host(ks2))             # aqml.py:795


def load_kernels(train_folder, kernel='g', rcut=4.8, z=None, path=None):
    d = np.load(path or kernels_file(train_folder, kernel, rcut, z))
    return d['ks1'], d['ks2']                                   # aqml.py:699-700
