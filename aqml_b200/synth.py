"""Deterministic synthetic QM9-/amons-shaped molecules.

Pure NumPy, no reference code: a random heavy-atom tree (bond 1.25-1.55 A, no two
heavy atoms closer than 1.15 A) saturated with hydrogens (0.95-1.10 A, nothing closer
than 0.9 A).  Used by tests and bench.py so every number is reproducible from a seed.
"""
from __future__ import annotations

import numpy as np

QM9_HEAVY = ([6, 7, 8, 9], [0.72, 0.12, 0.15, 0.01])
BIG_HEAVY = ([6, 7, 8, 9, 15, 16, 17], [0.63, 0.12, 0.15, 0.01, 0.03, 0.04, 0.02])
VALENCE = {6: 4, 7: 3, 8: 2, 9: 1, 15: 3, 16: 2, 17: 1}


def _rand_dir(rng):
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def make_molecule(rng, n_heavy, heavy=QM9_HEAVY, max_atoms=29):
    zs = [int(rng.choice(heavy[0], p=heavy[1]))]
    xyz = [np.zeros(3)]
    deg = [0]
    guard = 0
    while len(zs) < n_heavy and guard < 10000:
        guard += 1
        cand = [i for i in range(len(zs)) if deg[i] < VALENCE[zs[i]]]
        if not cand:
            break
        p = int(rng.choice(cand))
        pos = xyz[p] + _rand_dir(rng) * rng.uniform(1.25, 1.55)
        d = np.linalg.norm(np.array(xyz) - pos, axis=1)
        d[p] = 10.0
        if d.min() < 1.15:
            continue
        z = int(rng.choice(heavy[0], p=heavy[1]))
        zs.append(z)
        xyz.append(pos)
        deg.append(1)
        deg[p] += 1
    nh = len(zs)
    for i in range(nh):
        free = VALENCE[zs[i]] - deg[i]
        for _ in range(max(free, 0)):
            if len(zs) >= max_atoms:
                break
            for _try in range(30):
                pos = xyz[i] + _rand_dir(rng) * rng.uniform(0.95, 1.10)
                d = np.linalg.norm(np.array(xyz) - pos, axis=1)
                d[i] = 10.0
                if d.min() >= 0.9:
                    zs.append(1)
                    xyz.append(pos)
                    break
    return np.array(zs, dtype=np.int32), np.array(xyz, dtype=np.float64)


def qm9_like(n_mol, seed, max_heavy=9):
    """Returns (nas[int32 n_mol], zs[int32 A], coords[float64 A,3])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sizes = np.arange(1, max_heavy + 1)
    if max_heavy >= 9:
        p = np.full(max_heavy, 0.07 / (max_heavy - 2))
        p[-1] = 0.80
        p[-2] = 0.13
    else:
        p = np.full(max_heavy, 1.0 / max_heavy)
    p = p / p.sum()
    nas, zs, coords = [], [], []
    for _ in range(n_mol):
        z, x = make_molecule(rng, int(rng.choice(sizes, p=p)))
        nas.append(len(z))
        zs.append(z)
        coords.append(x)
    return np.array(nas, dtype=np.int32), np.concatenate(zs), np.concatenate(coords)


def replicate(base, n_mol, seed, jitter=0.03):
    """Draw ``n_mol`` molecules from a pool ``base=(nas,zs,coords)``: each copy gets its own
    random rigid rotation and an independent Gaussian jitter (sigma ``jitter`` A) per atom,
    so every molecule is geometrically distinct.  Vectorised (bench-size sets in seconds)."""
    nas0, zs0, co0 = base
    rng = np.random.Generator(np.random.PCG64(seed))
    off0 = np.concatenate(([0], np.cumsum(nas0)))
    pick = rng.integers(0, len(nas0), size=n_mol)
    nas = nas0[pick].astype(np.int32)
    off = np.concatenate(([0], np.cumsum(nas)))
    A = int(off[-1])
    mol_of = np.repeat(np.arange(n_mol), nas)
    src = off0[pick][mol_of] + (np.arange(A) - off[:-1][mol_of])
    zs = zs0[src].astype(np.int32)
    q, r = np.linalg.qr(rng.normal(size=(n_mol, 3, 3)))
    q = q * np.sign(np.diagonal(r, axis1=1, axis2=2))[:, None, :]
    coords = np.einsum('aij,aj->ai', q[mol_of], co0[src]) + rng.normal(scale=jitter, size=(A, 3))
    return nas, zs, np.ascontiguousarray(coords)


def qm9_like_fast(n_mol, seed, pool=512, max_heavy=9):
    """QM9-shaped set of any size: ``pool`` tree-grown molecules, replicated with rotation+jitter."""
    if n_mol <= pool:
        return qm9_like(n_mol, seed, max_heavy)
    return replicate(qm9_like(pool, seed, max_heavy), n_mol, seed + 1)


def large_like(n_mol, seed, lo=100, hi=200):
    """Config-4 query molecules: 8 elements, 100-200 atoms."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nas, zs, coords = [], [], []
    for _ in range(n_mol):
        n_at = int(rng.integers(lo, hi + 1))
        z, x = make_molecule(rng, max(2, n_at // 2), heavy=BIG_HEAVY, max_atoms=n_at)
        nas.append(len(z))
        zs.append(z)
        coords.append(x)
    return np.array(nas, dtype=np.int32), np.concatenate(zs), np.concatenate(coords)


def split_zs(nas, zs):
    off = np.concatenate(([0], np.cumsum(nas)))
    return [zs[off[i]:off[i + 1]] for i in range(len(nas))]
