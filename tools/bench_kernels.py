#!/usr/bin/env python
"""Device-resident micro-benchmarks of every kernel on the path (not the headline bench; feeds DESIGN.md 4).

    python tools/bench_kernels.py > gpurun_out/kernels.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aqml_b200 import _lib as L, fused, kernels, synth  # noqa: E402
from aqml_b200.algo import aqml as algo  # noqa: E402
from aqml_b200.representation import core as rcore  # noqa: E402


def timeit(fn, warmup=2, reps=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return float(np.median(ts))


def main():
    out = {}
    dev = torch.device("cuda")
    kw = dict(sigmas=[.05, .05], dgrids=[.04, .04], rcut=4.8)
    nas, zs, coords = synth.qm9_like_fast(6000, 99, pool=512)
    mb = rcore.get_slatm_mbtypes(synth.split_zs(nas[:1500], zs[:int(nas[:1500].sum())]) +
                                 [np.array([1, 1, 1, 6, 6, 6, 7, 7, 7, 8, 8, 8, 9, 9, 9])])
    cd = torch.as_tensor(coords, device=dev)
    A = int(nas.sum())
    D = 8975
    # --- SLATM generation -----------------------------------------------------------------------------
    t = timeit(lambda: rcore.generate_slatm_batch(cd, zs, nas, mb, local=True, **kw))
    out["aslatm_dense"] = dict(atoms=A, D=D, seconds=t, atoms_per_s=A / t, GBps=8.0 * D * A / t / 1e9)
    t = timeit(lambda: rcore.generate_slatm_batch(cd, zs, nas, mb, local=False, **kw))
    out["slatm_global"] = dict(molecules=len(nas), atoms=A, seconds=t, molecules_per_s=len(nas) / t, atoms_per_s=A / t)
    t = timeit(lambda: fused.PackedRep(cd, zs, nas, mb, **kw).close())
    out["aslatm_packed"] = dict(atoms=A, seconds=t, atoms_per_s=A / t, GBps=8.0 * 2032 * A / t / 1e9)
    # --- global kernels on SLATM vectors ----------------------------------------------------------------
    xg = rcore.generate_slatm_batch(cd, zs, nas, mb, local=False, **kw)           # (6000, 8975)
    X = torch.cat([xg, xg * 1.01, xg * 0.99])[:16384]                               # 16384 rows
    sig = np.array([50., 100., 200., 400.])
    n1, n2 = 8192, 16384
    t = timeit(lambda: kernels.global_kernels(X[:n1], X[:n2], sig, "gaussian"))
    out["global_gaussian"] = dict(n1=n1, n2=n2, D=D, S=4, seconds=t, elements_per_s=4.0 * n1 * n2 / t,
                                  TFLOPs=2.0 * D * n1 * n2 / t / 1e12)
    n1l, n2l = 4096, 8192
    t = timeit(lambda: kernels.global_kernels(X[:n1l], X[:n2l], sig * 40, "laplacian"), warmup=1, reps=2)
    out["global_laplacian"] = dict(n1=n1l, n2=n2l, D=D, S=4, seconds=t, elements_per_s=4.0 * n1l * n2l / t,
                                   fp64_add_Tops=2.0 * D * n1l * n2l / t / 1e12)
    t = timeit(lambda: algo.cmld.max_distance(X[:8192], X[:8192], 2))
    out["max_distance_l2"] = dict(n=8192, D=D, seconds=t, TFLOPs_full_square=2.0 * D * 8192 * 8192 / t / 1e12)
    # --- local kernels through the dense (reference-shaped) API -----------------------------------------
    nm1, nm2 = 300, 1500
    a1 = int(nas[:nm1].sum())
    a2 = int(nas[nm1:nm1 + nm2].sum())
    xl = rcore.generate_slatm_batch(cd[:a1 + a2], zs[:a1 + a2], nas[:nm1 + nm2], mb, local=True, **kw)
    zmax = 9
    sg = np.array([c * np.full(zmax, 16.0) / np.sqrt(2 * np.log(2)) for c in (0.5, 1, 2, 4)])
    for center in (True, False):
        t = timeit(lambda: kernels.get_local_kernels_gaussian(xl[:a1], xl[a1:], zs[:a1], zs[a1:a1 + a2], nas[:nm1],
                                                              nas[nm1:nm1 + nm2], False, zmax, sg, center=center))
        out["local_gaussian_dense_api_center%d" % center] = dict(nm1=nm1, nm2=nm2, atoms=(a1, a2), seconds=t,
                                                                elements_per_s=4.0 * nm1 * nm2 / t)
    t = timeit(lambda: kernels.get_local_kernels_laplacian(xl[:a1], xl[a1:], nas[:nm1], nas[nm1:nm1 + nm2],
                                                           [400., 800., 1600., 3200.]), warmup=1, reps=2)
    out["local_laplacian"] = dict(nm1=nm1, nm2=nm2, atoms=(a1, a2), seconds=t, elements_per_s=4.0 * nm1 * nm2 / t,
                                  fp64_add_Tops=2.0 * D * a1 * a2 / t / 1e12)
    t = timeit(lambda: algo.get_kernel_width(nas[:nm1 + nm2], zs[:a1 + a2], xl))
    out["kernel_width"] = dict(atoms=a1 + a2, seconds=t)
    # --- peaks --------------------------------------------------------------------------------------
    n = 8192
    a = torch.randn((n, n), dtype=torch.float64, device=dev)
    b = torch.randn((n, n), dtype=torch.float64, device=dev)
    t = timeit(lambda: torch.matmul(a, b), warmup=2, reps=5)
    out["cublas_dgemm_8192"] = dict(seconds=t, TFLOPs=2.0 * n ** 3 / t / 1e12)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
