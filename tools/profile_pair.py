#!/usr/bin/env python
"""Small driver for ncu captures of the pair kernel and the aSLATM kernel (tools only; not part of the product).

  python tools/profile_pair.py [n_train] [n_test] [repeats]

Packs n_train + n_test QM9-shaped molecules, takes the kernel width, then runs K(test, train) `repeats` times and
K(train, train) (symmetric) once.  Prints the device time of each call.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from aqml_b200 import fused, synth
    from aqml_b200.representation import core as rcore
    n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 600
    rep = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    kw = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
    tr = synth.qm9_like_fast(n1, 20251021, pool=512)
    te = synth.qm9_like_fast(n2, 20252021, pool=256)
    zl = synth.split_zs(tr[0][:500], tr[1][:int(tr[0][:500].sum())])
    zl.append(np.array([1, 1, 1, 6, 6, 6, 7, 7, 7, 8, 8, 8, 9, 9, 9]))
    mb = rcore.get_slatm_mbtypes(zl)
    r1 = fused.PackedRep(tr[2], tr[1], tr[0], mb, **kw)
    r2 = fused.PackedRep(te[2], te[1], te[0], mb, **kw)
    dso = fused.kernel_width([r1, r2], [r1, r2], 9)
    coeffs = tuple(float(v) for v in os.environ.get("AQML_PROFILE_COEFFS", "0.5,1,2,4").split(","))
    sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in coeffs])
    K = torch.empty((len(coeffs), n2, n1), dtype=torch.float64, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    for i in range(rep):
        a, b = ev(), ev()
        a.record()
        fused.local_gaussian(r2, r1, sig, 9, out=K)
        b.record()
        torch.cuda.synchronize()
        print("K(test,train) %d x %d: %.3f ms" % (n2, n1, a.elapsed_time(b)))
    a, b = ev(), ev()
    a.record()
    K11 = fused.local_gaussian(r1, r1, sig, 9, symmetric=True)
    b.record()
    torch.cuda.synchronize()
    print("K(train,train) symmetric %d: %.3f ms, checksum %.6e" % (n1, a.elapsed_time(b), float(K11.sum())))


if __name__ == "__main__":
    main()
