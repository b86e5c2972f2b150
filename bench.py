#!/usr/bin/env python
"""Benchmark of the AQML hot path on B200 (contract: see the task prompt / DESIGN.md section 6).

Workload (BASELINE.json configs[1]): 20k train x 2k test QM9-shaped molecules, local aSLATM element-matched
Gaussian kernel, FP64, 4 sigmas -- as the KRR feed the reference builds it for (algo/aqml.py:711-789, 876).
One *step* = the whole hot path for that ONE fixed problem, cut over the N ranks (aqml_b200/sharded.py):

    coordinates -> aSLATM (train blocks 1/N per rank + test)              phase aslatm
                -> every block's packed operands to every rank             phase allgather  (NCCL, N > 1)
                -> kernel width over all train + test atoms                phase width      (+ all-reduce MAX)
                -> ks1 = K(train, train): lower block triangle, 4 sigmas   phase k_train
                -> ks2 = K(test, train), sharded by training columns       phase k_test
                -> y = ks2 . alpha                                         phase predict    (+ all-reduce SUM)

`value`  : UNIQUE kernel elements delivered per second, whole job: 4 * (n_train (n_train + 1) / 2 + n_test n_train) per
           step, coordinates already resident in HBM (device timed with CUDA events, max over ranks).
`e2e`    : same metric through the public API with HOST buffers: coordinates H2D from pinned memory, every K tile and
           the predictions D2H into pinned memory, all inside the timed region.
N > 1    : strong scaling of the fixed problem; collectives inside the timed region (see phases_ms).
`--impl reference`: the restated reference loops (C/OpenMP, oracle/c_oracle.c -- the Fortran cannot be compiled in
           this image) on the host cores, the same phases on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20251019 + 2
COEFFS = (0.5, 1.0, 2.0, 4.0)
SLATM_KW = dict(sigmas=(.05, .05), dgrids=(.04, .04), rcut=4.8)
ZMAX = 9
METRIC = "local_gaussian_kernel_elements_per_sec"
WORKLOAD = ("config2 as KRR feed: 20k train x 2k test QM9-shaped molecules, local aSLATM element-matched Gaussian "
            "kernel, FP64, 4 sigmas; step = coords -> aSLATM(train+test) -> kernel width -> K(train,train) "
            "[lower block triangle] -> K(test,train) -> K(test,train).alpha")


def init_nccl(local_rank):
    """NCCL prints its version banner on stdout; keep stdout for the single JSON line."""
    import torch
    import torch.distributed as dist
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-train", type=int, default=20000)
    ap.add_argument("--n-test", type=int, default=2000)
    ap.add_argument("--cpu-train", type=int, default=96, help="CPU-baseline sample: training molecules")
    ap.add_argument("--cpu-test", type=int, default=24, help="CPU-baseline sample: test molecules")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fp64-engine", action="store_true", help="skip the extra FP64 DMMA engine measurement")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"],
                    help="config2 (default, the headline line) | config3: global SLATM Gaussian kernel, "
                         "row block of --rows-per-gpu molecules x --n-cols molecules per GPU, 4 sigmas")
    ap.add_argument("--n-cols", type=int, default=100000)
    ap.add_argument("--rows-per-gpu", type=int, default=12500)
    ap.add_argument("--kernel", default="gaussian", choices=["gaussian", "laplacian"])
    ap.add_argument("--square", action="store_true", help="config3: the n_cols x n_cols problem, rows split over the ranks")
    ap.add_argument("--no-config3", action="store_true", help="N >= 4: skip the embedded 100k x 100k global-kernel measurement")
    return ap.parse_args()


def make_sets(n_train, n_test):
    from aqml_b200 import synth
    train = synth.qm9_like_fast(n_train, SEED, pool=512)
    test = synth.qm9_like_fast(n_test, SEED + 1000, pool=256)
    return train, test


def mbtypes_for(train, test):
    from aqml_b200 import synth
    from aqml_b200.representation import core as rcore
    zl = synth.split_zs(train[0][:2000], train[1][:int(train[0][:2000].sum())]) + \
        synth.split_zs(test[0][:500], test[1][:int(test[0][:500].sum())])
    # all five QM9 elements with full multiplicity (n1/n2/n3 = 5/15/75, D = 8975 at dgrid 0.04)
    zl.append(np.array([1, 1, 1, 6, 6, 6, 7, 7, 7, 8, 8, 8, 9, 9, 9]))
    return rcore.get_slatm_mbtypes(zl)


def sigmas_from(dso):
    """aqml.py:195,775: sigma = coeff * dmax / sqrt(2 ln 2) per element (absent elements keep 1e-9)."""
    return np.array([c * np.asarray(dso) / np.sqrt(2 * np.log(2)) for c in COEFFS])


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, index):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def wait_first(self, timeout=8.0):
        """Block until nvidia-smi has delivered its first sample: its start-up (NVML initialisation over every GPU of the box)
        takes driver locks and must not fall into the timed region."""
        t0 = time.time()
        while self.proc is not None and not self.samples and time.time() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.samples:
            if not (t0 <= t <= t1 + 0.3):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[6]))
            except Exception:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "power_w_max": max(power) if power else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm (restated reference loops): the same phases on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_sample(args, steps=1, warmup=0):
    """molecules -> aSLATM -> kernel width -> K(train, train) -> K(test, train) on the host cores, as the reference does
    it: dense (atoms, D) rows, the width from full per-element distance matrices (aqml.py:220-237), ks1 as a full
    n x n call of calc_lk_gaussian (aqml.py:788; the reference does not use its symmetry)."""
    from oracle import c_oracle as co
    train, test = make_sets(args.cpu_train, args.cpu_test)
    mb = mbtypes_for(*make_sets(2000, 500))
    co.load(fast=True)
    co.set_num_threads(len(os.sched_getaffinity(0)), fast=True)  # torchrun exports OMP_NUM_THREADS=1
    cores = co.num_threads(fast=True)
    nas = np.concatenate((train[0], test[0]))
    zs = np.concatenate((train[1], test[1]))

    def one():
        t = [time.perf_counter()]
        x1 = co.generate_slatm_batch(train[2], train[1], train[0], mb, True, fast=True, **SLATM_KW)
        x2 = co.generate_slatm_batch(test[2], test[1], test[0], mb, True, fast=True, **SLATM_KW)
        t.append(time.perf_counter())
        dso = co.get_kernel_width(nas, zs, np.concatenate((x1, x2)), fast=True)
        dso = np.concatenate((dso, np.full(ZMAX - len(dso), 1e-9)))
        sig = sigmas_from(dso)
        t.append(time.perf_counter())
        co.calc_lk_gaussian(x1, x1, train[1], train[1], train[0], train[0], ZMAX, False, sig, fast=True)
        t.append(time.perf_counter())
        co.calc_lk_gaussian(x2, x1, test[1], train[1], test[0], train[0], ZMAX, False, sig, fast=True)
        t.append(time.perf_counter())
        return np.diff(t)

    for _ in range(warmup):
        one()
    ph = np.median(np.array([one() for _ in range(max(steps, 1))]), axis=0)
    total = float(ph.sum())
    S, n, nt = len(COEFFS), args.cpu_train, args.cpu_test
    elems = S * (n * n + nt * n)          # what the reference computes: the full ks1 and ks2
    atoms = int(nas.sum())
    return dict(value=elems / total, unit="kernel elements/s", cores=cores, kind="port",
                sample=f"{n} train + {nt} test QM9-shaped molecules ({atoms} atoms, D=8975, 4 sigmas): molecules -> aSLATM "
                       f"-> kernel width -> K(train,train) [full {n}x{n}, as aqml.py:788] -> K(test,train); restated "
                       f"reference loops (C/OpenMP -O3 -march=native), not gfortran; value counts every element "
                       f"computed ({elems}), {S * (n * (n + 1) // 2 + nt * n) / total:.4g} unique elements/s",
                aslatm_atoms_per_sec=atoms / float(ph[0]), seconds=total,
                phases_ms={k: float(v) * 1e3 for k, v in zip(("aslatm", "width", "k_train", "k_test"), ph)})


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    r = cpu_sample(args, steps=args.steps, warmup=min(args.warmup, 1))
    ms = r["seconds"] * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": r["value"],
            "unit": "kernel elements/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD + " (CPU arm: bounded sample of it)", "sample": r["sample"]},
            "phases_ms": r["phases_ms"],
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "aslatm_atoms_per_sec": r["aslatm_atoms_per_sec"],
            "e2e": {"value": r["value"], "unit": "kernel elements/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def fp64_peak_tflops(torch):
    """cuBLAS DGEMM 8192^3, best of 10 (MEASURED_PEAKS.json has no FP64 entry)."""
    n = 8192
    a = torch.randn((n, n), dtype=torch.float64, device="cuda")
    b = torch.randn((n, n), dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    best = 1e9
    for i in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def run_ours(args):
    import torch
    import torch.distributed as dist
    from aqml_b200 import _lib as L
    from aqml_b200 import sharded

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        init_nccl(local_rank)
    lib = L.load()

    train, test = make_sets(args.n_train, args.n_test)   # the same fixed problem on every rank
    mb = mbtypes_for(train, test)
    S = len(COEFFS)
    dev = torch.device("cuda", local_rank)
    sk = sharded.ShardedKernels((train[0], train[1]), (test[0], test[1]), mb, ZMAX, rank, world, **SLATM_KW)
    c1 = torch.as_tensor(train[2], device=dev)
    c2 = torch.as_tensor(test[2], device=dev)
    rng = np.random.Generator(np.random.PCG64(SEED))
    alpha = torch.as_tensor(rng.standard_normal((S, args.n_train)) / args.n_train, device=dev)  # stands in for the solver's output
    PH = ("aslatm", "width", "allgather", "k_train", "k_test", "predict")   # N > 1: the broadcasts are enqueued before the width
                                                                              # pass and waited for after it ("allgather" = the wait)
    out = {}

    def ev():
        return torch.cuda.Event(enable_timing=True)

    dbg = bool(os.environ.get("AQML_BENCH_DEBUG"))
    host_t = []

    def step(timing=None, coords=None, after=None):
        e = [ev() for _ in range(len(PH) + 1)] if timing is not None else None
        ht = []

        def mark(i):
            if e:
                e[i].record()
            if dbg:
                ht.append(time.perf_counter())
        mark(0)
        sk.pack(*(coords or (c1, c2)))
        if not out:
            out["kt"], out["ks"] = sk.alloc_out(S)
        mark(1)
        sk.share_begin()
        sig = sigmas_from(sk.width())
        mark(2)
        sk.share_end()
        mark(3)
        sk.k_train(sig, out["kt"], after)
        mark(4)
        sk.k_test(sig, out["ks"], after)
        mark(5)
        y = sk.predict(out["ks"], alpha)
        mark(6)
        if timing is not None:
            timing.append(e)
            host_t.append(ht)
        return y, sig

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warm = max(args.warmup, 3)
    # the clock sampler starts BEFORE the warm-up: spawning nvidia-smi (fork of this process + NVML initialisation, which takes
    # driver locks) stalled the first timed step of every rank by ~0.5 s at N = 4 when it was started right before the timed region
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.wait_first()
    for _ in range(warm):
        step([])      # with (throw-away) timing events, exactly like a timed step: nothing lazily initialised is left for step 1
    barrier()
    rep_bytes = sum(r.nbytes for r in sk.reps) + sum(r.nbytes for r in sk.tests)
    launches0 = lib.aqml_launch_count()
    timing = []
    w0 = time.time()
    e_start, e_end = ev(), ev()
    e_start.record()
    for _ in range(args.steps):
        y, sig = step(timing)
    e_end.record()
    barrier()
    w1 = time.time()
    launches = int(lib.aqml_launch_count() - launches0)
    t_total_ms = e_start.elapsed_time(e_end)
    phases = [float(np.mean([e[i].elapsed_time(e[i + 1]) for e in timing])) for i in range(len(PH))]
    if os.environ.get("AQML_BENCH_DEBUG"):
        for k, e in enumerate(timing):
            print("rank %d step %d phases_ms %s | host enqueue ms %s" % (
                rank, k, " ".join("%s=%.1f" % (PH[i], e[i].elapsed_time(e[i + 1])) for i in range(len(PH))),
                " ".join("%.1f" % ((host_t[k][i + 1] - host_t[k][i]) * 1e3) for i in range(len(PH)))), file=sys.stderr, flush=True)
    clocks = sampler.stop(w0, w1) if sampler else None
    y_host = y.cpu().numpy()

    # ---- untimed: k-tiles the contraction executed in one step (device counters; reading them synchronises) -----
    L.call("aqml_pair_stats", None, None, 1)
    sk.k_train(sig, out["kt"])
    sk.k_test(sig, out["ks"])
    kt_exec, kt_dense = C.c_int64(0), C.c_int64(0)
    L.call("aqml_pair_stats", C.c_void_p(C.addressof(kt_exec)), C.c_void_p(C.addressof(kt_dense)), 1)
    kt_exec, kt_dense = float(kt_exec.value), float(kt_dense.value)
    bm, bn, bk = C.c_int(0), C.c_int(0), C.c_int(0)
    lib.aqml_tile_shape(C.c_void_p(C.addressof(bm)), C.c_void_p(C.addressof(bn)), C.c_void_p(C.addressof(bk)))
    ktile_flops = 2.0 * bm.value * bn.value * bk.value

    # ---- untimed: the FP64 DMMA engine on the K(test, train) phase of the same inputs ---------------------------
    fp64_engine = None
    engine = "dmma" if os.environ.get("AQML_PAIR_ENGINE", "") in ("dmma", "fp64") else "i8"
    if rank == 0 and world == 1 and not args.no_fp64_engine and engine == "i8":
        os.environ["AQML_PAIR_ENGINE"] = "dmma"
        try:
            sk.k_test(sig, out["ks"])
            torch.cuda.synchronize()
            L.call("aqml_pair_stats", None, None, 1)
            a, b = ev(), ev()
            a.record()
            sk.k_test(sig, out["ks"])
            b.record()
            torch.cuda.synchronize()
            ke, kd = C.c_int64(0), C.c_int64(0)
            L.call("aqml_pair_stats", C.c_void_p(C.addressof(ke)), C.c_void_p(C.addressof(kd)), 1)
            fp64_engine = {"k_test_ms": a.elapsed_time(b), "ktiles_executed": float(ke.value)}
        finally:
            del os.environ["AQML_PAIR_ENGINE"]

    # ---- end to end through the public API with host buffers ----------------------------------------------------
    e2e_ms, h2d, d2h, e2e_phases = None, 0, 0, [0.0] * len(PH)
    if not args.no_e2e:
        pc1 = torch.as_tensor(train[2]).pin_memory().numpy()
        pc2 = torch.as_tensor(test[2]).pin_memory().numpy()
        host = {k: torch.empty(v.shape, dtype=torch.float64, pin_memory=True) for k, v in list(out["kt"].items()) + list(out["ks"].items())}
        yh = torch.empty((S, args.n_test), dtype=torch.float64, pin_memory=True)
        side = torch.cuda.Stream()
        cur = torch.cuda.current_stream()

        def to_host(key, t):  # D2H of a finished tile on a second stream, behind the next tile's contraction
            ready = torch.cuda.Event()
            ready.record(cur)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                host[key].copy_(t, non_blocking=True)

        e2e_timing = []

        def e2e_step(tm=None):
            cur.wait_stream(side)            # the previous step's copies have left the tile buffers
            y, _ = step(tm, coords=(pc1, pc2), after=to_host)
            yh.copy_(y, non_blocking=True)
            cur.wait_stream(side)
        for _ in range(2):
            e2e_step([])
        barrier()
        a, b = ev(), ev()
        n_e2e = max(2, min(args.steps, 3))
        a.record()
        for _ in range(n_e2e):
            e2e_step(e2e_timing)
        b.record()
        barrier()
        e2e_ms = a.elapsed_time(b) / n_e2e
        e2e_phases = [float(np.mean([e[i].elapsed_time(e[i + 1]) for e in e2e_timing])) for i in range(len(PH))]
        atoms_mine = sum(int(sk.nas1[lo:hi].sum()) for lo, hi in (sk.bounds[b] for b in sk.mine)) + int(sk.nas2[slice(*sk.tbounds[rank])].sum())
        h2d = atoms_mine * (24 + 4) + sig.nbytes
        d2h = sum(int(v.numel()) * 8 for v in host.values()) + yh.numel() * 8
        del host

    # ---- reduce over ranks ----------------------------------------------------------------------------------------
    vals = torch.tensor([t_total_ms, e2e_ms or 0.0] + phases + e2e_phases, dtype=torch.float64, device=dev)
    sums = torch.tensor([float(h2d), float(d2h), float(sk.unique_elements(S)), kt_exec, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    vals = [float(v) for v in vals.cpu()]
    t_total_ms, e2e_max, phases, e2e_phases = vals[0], vals[1], vals[2:2 + len(PH)], vals[2 + len(PH):]
    h2d_all, d2h_all, elems_all, kt_exec_all, launches_all = [float(v) for v in sums.cpu()]
    ms_per_step = t_total_ms / args.steps
    ph = dict(zip(PH, phases))
    atoms_all = int(train[0].sum() + test[0].sum())

    # ---- N >= 4: BASELINE configs[2] / the north_star's target, driver-measured: the 100k x 100k global SLATM Gaussian kernel,
    # 4 sigmas, rows split over the ranks (K of one rank: 100k / N x 100k x 4 x 8 B -- it fits from N = 4 on) ----------------
    config3 = None
    if world >= 4 and not args.no_config3:
        sk.close()
        out.clear()
        torch.cuda.empty_cache()
        config3 = config3_measure(args, rank, local_rank, world, steps=2, warmup=2, n_cols=100000, rows_per_gpu=0, kernel="gaussian",
                                  square=True)

    if rank == 0:
        p = L.slatm_params(mb, **{k: SLATM_KW[k] for k in SLATM_KW})
        nel = 5
        d_z = {z: nel * p.nx2 + (nel * (nel + 1) // 2) * p.nx3 for z in (1, 6, 7, 8, 9)}
        # algorithmic FP64 flops of the contraction: 2 sum_z D_z n1_z n2_z; symmetric K(train, train): half
        flops = 0.0
        for z in d_z:
            n1z, n2z = float((train[1] == z).sum()), float((test[1] == z).sum())
            flops += 2.0 * d_z[z] * (n1z * (n1z + 1) / 2 + n2z * n1z)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        t_k_ms = ph["k_train"] + ph["k_test"]     # the contraction phases (i8_pair_kernel + a memset / mirror kernel per tile)
        nprod = 26 if int(os.environ.get("AQML_I8_PRODUCTS", "22") or 22) >= 26 else 22
        exec_flops = kt_exec * ktile_flops        # rank 0: FP64-equivalent flops of the k-tiles actually contracted
        dgemm_tf = fp64_peak_tflops(torch)
        traffic = None
        try:  # dram__bytes_read+write of the step's i8_pair_kernel launches, from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["i8_pair_kernel"]
            if args.n_train == 20000 and args.n_test == 2000 and world == 1:
                traffic = float(tr["dram_bytes_read"] + tr["dram_bytes_write"])
        except Exception:
            pass
        if engine == "i8":
            pk = C.c_double(0.0)
            L.call("aqml_measure_i8_peak", C.c_void_p(C.addressof(pk)), L.current_stream())
            i8_peak = pk.value / 1e12
            ach = exec_flops * nprod / (t_k_ms * 1e-3) / 1e12
            roofline = {
                "bound": "tensor", "achieved": ach, "peak": i8_peak, "unit": "TOP/s (int8 multiply-adds x 2, executed, per GPU)",
                "frac": ach / i8_peak, "traffic": traffic,
                "traffic_unit": "bytes of DRAM per step over the i8_pair_kernel launches (profiles/r02_traffic.json)",
                "kernel": "i8_pair_kernel (tcgen05.mma kind::i8: six 8-bit digit planes, %d exact int32 plane products in TMEM "
                          "per FP64 product, issued as stacked-N UTCIMMA)" % nprod,
                "peak_source": "aqml_measure_i8_peak(): UTCIMMA M128 N256 K32, operands resident in shared memory, measured in "
                               "this run on this GPU (best of 3)",
                "kernel_ms_per_step": t_k_ms,
                "int8_ops_per_step": exec_flops * nprod,
                "ktiles_executed_over_dense": kt_exec / max(kt_dense, 1.0),
                "algorithmic_fp64_flops_per_step": flops, "executed_fp64_equiv_flops_per_step": exec_flops,
                "fp64_equiv_tflops": flops / world / (t_k_ms * 1e-3) / 1e12,
                "dgemm_peak_tflops": dgemm_tf,
                "fp64_equiv_over_dgemm_peak": flops / world / (t_k_ms * 1e-3) / 1e12 / dgemm_tf,
                "note": "frac = int8 operations the tensor pipe executed / the pipe's peak measured in this run (<= 1). "
                        "fp64_equiv_* is a speed-up figure, not a utilisation: SURVEY 8(d) flop count 2*sum_z D_z*n1_z*n2_z "
                        "(halved for the symmetric K(train,train)) per GPU over the same time, against cuBLAS DGEMM 8192^3 "
                        "measured in this run; all-zero 16-column k-tiles are skipped exactly (ktiles_executed_over_dense)"}
            if peaks.get("bf16_tflops"):
                roofline["frac_of_2x_measured_bf16_burst"] = ach / (2.0 * peaks["bf16_tflops"])
        else:
            ach = exec_flops / (t_k_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "achieved": ach, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": ach / dgemm_tf,
                        "traffic": None, "kernel": "pair_tile_kernel<MODE_DOT,EPI_LOCAL> (FP64 DMMA.8x8x4), executed flops",
                        "peak_source": "cuBLAS DGEMM 8192^3 best of 10 measured in this run",
                        "algorithmic_fp64_flops_per_step": flops, "kernel_ms_per_step": t_k_ms}
        # aSLATM phase: bytes the kernel has to write -- packed FP64 operand 8 * sum_z D_z(padded to 16) * A_z and,
        # for the int8 engine, 6 more bytes of digit planes per element; per GPU this is 1/N of the training set + the test set
        el_all = sum((-(-d_z[z] // 16) * 16) * float((test[1] == z).sum() + (train[1] == z).sum()) for z in d_z)
        el_rank = sum((-(-d_z[z] // 16) * 16) * float((test[1] == z).sum() + (train[1] == z).sum() / world) for z in d_z)
        slatm_bytes = 8.0 * el_rank
        line = {
            "metric": METRIC,
            "value": elems_all / (ms_per_step * 1e-3), "unit": "kernel elements/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "dtype_detail": ("FP64 in, FP64 out, <= 1e-10 vs the FP64 oracle; a.b from exact int8 digit-plane products (int32 "
                             "accumulators) on the tcgen05 tensor cores, norms / exp / sums in FP64; AQML_PAIR_ENGINE=dmma: "
                             "FP64 DMMA throughout") if engine == "i8" else "FP64 DMMA throughout",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_train": args.n_train, "n_test": args.n_test, "atoms": atoms_all,
                       "D": int(lib.aqml_slatm_dim(p)), "D_z": d_z[6], "nsigmas": S,
                       "elements_per_step": int(elems_all),
                       "elements_counted": "unique elements: lower triangle (incl. diagonal) of K(train,train) + K(test,train), "
                                           "x 4 sigmas; diagonal tiles are delivered mirrored",
                       "sharding": "2N training blocks, rank r owns blocks r and 2N-1-r (+ test block r of N): their aSLATM, its block "
                                   "rows of the lower triangle of K(train,train), its columns of K(test,train); NCCL: one "
                                   "broadcast per block (pair-kernel operands) overlapped with the distributed width pass (six small "
                                   "collectives), all-reduce SUM (prediction)",
                       "l2": "inputs larger than L2 (packed operands + digit planes %.1f GB per GPU and step)" % (rep_bytes / 1e9)},
            "aslatm_atoms_per_sec": atoms_all / (ph["aslatm"] * 1e-3),
            "phases_ms": ph,
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_aslatm": {"bound": "hbm", "kernel": "slatm_units_kernel (packed FP64 output) + digit-plane split",
                                "achieved": slatm_bytes / (ph["aslatm"] * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": slatm_bytes / (ph["aslatm"] * 1e-3) / 1e9 / hbm_peak,
                                "algorithmic_bytes_per_step": slatm_bytes, "digit_plane_bytes_per_step": 6.0 * el_rank,
                                "whole_problem_bytes": 8.0 * el_all,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
            "prediction_checksum": float(np.abs(y_host).sum()),
        }
        if fp64_engine is not None:
            ef = fp64_engine["ktiles_executed"] * ktile_flops
            ms = fp64_engine["k_test_ms"]
            line["roofline_fp64_engine"] = {
                "kernel": "pair_tile_kernel<MODE_DOT,EPI_LOCAL> (FP64 DMMA.8x8x4), AQML_PAIR_ENGINE=dmma, K(test,train) phase of the "
                          "same inputs", "k_test_ms": ms, "bound": "tensor", "achieved": ef / (ms * 1e-3) / 1e12, "peak": dgemm_tf,
                "unit": "TFLOP/s", "frac": ef / (ms * 1e-3) / 1e12 / dgemm_tf,
                "speedup_of_default_engine": ms / ph["k_test"],
                "note": "the north_star's literal design point (FP64 DMMA tiles), kept as the fallback engine"}
        if e2e_ms is not None:
            line["e2e"] = {"value": elems_all / (e2e_max * 1e-3), "unit": "kernel elements/s",
                           "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_max,
                           "phases_ms": dict(zip(PH, e2e_phases))}
        if config3 is not None:
            line["config3_100k_x_100k"] = config3
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_sample(args, steps=1, warmup=0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["aslatm_atoms_per_sec"] = r["aslatm_atoms_per_sec"]
            line["cpu_baseline"]["phases_ms"] = r["phases_ms"]
        print(json.dumps(line))
    sk.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def config3_measure(args, rank, local_rank, world, steps, warmup, n_cols, rows_per_gpu, kernel, square):
    """BASELINE.json configs[2]: global SLATM kernel with 4 sigmas, row-block sharded.  square: THE n_cols x n_cols problem
    (north_star: "100k x 100k ... row blocks of the training set across the 8xB200 box") -- rank r builds K[:, rows_r, :] for
    its 1/N of the molecules against all of them; the SLATM rows are generated 1/N per rank and all-gathered once over NCCL
    (the rank's own rows are a slice of the result).  Otherwise every rank has its own rows_per_gpu query molecules.
    Step = coordinates -> global SLATM -> K block resident in HBM.  Returns the JSON-able result (all ranks)."""
    import torch
    import torch.distributed as dist
    from aqml_b200 import _lib as L
    from aqml_b200 import distance, kernels, parallel, synth
    from aqml_b200.representation import core as rcore
    dev = torch.device("cuda", local_rank)
    cols = synth.qm9_like_fast(n_cols, SEED + 7, pool=512)
    rows = None if square else synth.qm9_like_fast(rows_per_gpu, SEED + 7 + 1000 * (rank + 1), pool=256)
    mb = mbtypes_for(cols, cols if square else rows)
    kw = dict(sigmas=[.05, .05], dgrids=[.04, .04], rcut=4.8)
    cc = torch.as_tensor(cols[2], device=dev)
    cr = None if square else torch.as_tensor(rows[2], device=dev)
    nsub = int(cols[0][:2000].sum())
    xs = rcore.generate_slatm_batch(cc[:nsub], cols[1][:nsub], cols[0][:2000], mb, local=False, **kw)
    dmax = distance.max_distance(xs, xs, 2 if kernel == "gaussian" else 1)
    del xs
    fac = 1.0 / np.sqrt(2 * np.log(2)) if kernel == "gaussian" else 1.0 / np.log(2)
    sig = np.array([c * dmax * fac for c in COEFFS])
    D = int(L.load().aqml_slatm_dim(L.slatm_params(mb, **SLATM_KW)))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    # SLATM of the column molecules is sharded by molecules over the ranks, then one all-gather
    shard = world > 1 and not os.environ.get("AQML_NO_SLATM_SHARD")
    parts = [parallel.shard_range(n_cols, r, world) for r in range(world)]
    clo, chi = parts[rank]
    if shard:
        coff = np.concatenate(([0], np.cumsum(cols[0])))
        cc_loc, zs_loc, nas_loc = cc[coff[clo]:coff[chi]], cols[1][coff[clo]:coff[chi]], cols[0][clo:chi]
        counts = [h - l for l, h in parts]
    nrows = (chi - clo) if square else rows_per_gpu

    def step(tm=None):
        e = [ev() for _ in range(3)]
        e[0].record()
        if shard:
            xc = parallel.allgather_rows(rcore.generate_slatm_batch(cc_loc, zs_loc, nas_loc, mb, local=False, **kw), counts)
        else:
            xc = rcore.generate_slatm_batch(cc, cols[1], cols[0], mb, local=False, **kw)
        xr = xc[clo:chi] if square else rcore.generate_slatm_batch(cr, rows[1], rows[0], mb, local=False, **kw)
        e[1].record()
        K = kernels.global_kernels(xr, xc, sig, kernel)
        e[2].record()
        if tm is not None:
            tm.append(e)
        return K

    for _ in range(max(warmup, 1)):
        K = step()
        del K
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tm = []
    a, b = ev(), ev()
    a.record()
    for _ in range(steps):
        K = step(tm)
        del K
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b), float(np.mean([e[0].elapsed_time(e[1]) for e in tm])),
                      float(np.mean([e[1].elapsed_time(e[2]) for e in tm]))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot, t_sl, t_k = [float(v) for v in t.cpu()]
    torch.cuda.empty_cache()
    ms = tot / steps
    S = len(COEFFS)
    total_rows = n_cols if square else world * rows_per_gpu
    return {
        "metric": "global_%s_kernel_elements_per_sec" % kernel,
        "value": S * total_rows * n_cols / (ms * 1e-3), "unit": "kernel elements/s",
        "n_gpus": world, "steps": steps, "warmup": max(warmup, 1), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if square else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config3: global SLATM %s kernel, 4 sigmas, %s, D=%d; K row-block sharded%s"
                               % (kernel, ("the %d x %d problem, %d rows per GPU" % (n_cols, n_cols, nrows)) if square
                                  else ("%d rows per GPU x %d columns" % (rows_per_gpu, n_cols)), D,
                                  "; SLATM generated 1/N per rank + one NCCL all-gather" if shard else ", no collective")},
        "phases_ms": {"slatm_incl_allgather": t_sl, "kernel_incl_pack": t_k},
        "slatm_molecules_per_sec": (n_cols if square else world * (n_cols + rows_per_gpu)) / (t_sl * 1e-3),
        "kernel_fp64_equiv_tflops_per_gpu": 2.0 * D * nrows * n_cols / (t_k * 1e-3) / 1e12}


def run_config3(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        init_nccl(local_rank)
    res = config3_measure(args, rank, local_rank, world, args.steps, max(args.warmup, 3), args.n_cols, args.rows_per_gpu, args.kernel,
                          square=args.square)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "config3":
        run_config3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
