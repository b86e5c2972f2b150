import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from aqml_b200 import fused, synth
from aqml_b200.representation import core as rcore
tr = synth.qm9_like_fast(10000, 20251021, pool=512)
zl = synth.split_zs(tr[0][:500], tr[1][:int(tr[0][:500].sum())]); zl.append(np.array([1,1,1,6,6,6,7,7,7,8,8,8,9,9,9]))
mb = rcore.get_slatm_mbtypes(zl)
kw = dict(sigmas=(.05,.05), dgrids=(.04,.04), rcut=4.8)
c = torch.as_tensor(tr[2], device='cuda')
for it in range(4):
    torch.cuda.synchronize(); t0=time.perf_counter()
    a,b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = fused.PackedRep(c, tr[1], tr[0], mb, **kw); b.record()
    t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter()
    print("events %.2f ms; host enqueue %.2f ms; wall %.2f ms"%(a.elapsed_time(b), (t1-t0)*1e3, (t2-t0)*1e3)); r.close()
