#!/usr/bin/env python
"""Benchmark of the AQML hot path on B200 (contract: see the task prompt / DESIGN.md section 6).

Workload (BASELINE.json configs[1]): 20k train x 2k test QM9-shaped molecules, local aSLATM
element-matched Gaussian kernel, FP64, 4 sigmas.  One *step* = the whole hot path for one batch:

    coordinates -> aSLATM (train + test, written element-packed) -> K[4, n_test, n_train]

`value`  : kernel elements/s with coordinates already resident in HBM (device timed, CUDA events).
`e2e`    : same metric through the C ABI with HOST buffers (coords H2D, K D2H inside the timing).
N > 1    : weak scaling -- every rank owns its own 2k test molecules (a row block of K) against
           the same 20k training molecules; no data-path collective.
`--impl reference`: the restated reference loops (C/OpenMP, oracle/c_oracle.c -- the Fortran cannot
           be compiled in this image) on the host cores, bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-train", type=int, default=20000)
    ap.add_argument("--n-test", type=int, default=2000)
    ap.add_argument("--cpu-train", type=int, default=256, help="CPU-baseline sample: training molecules")
    ap.add_argument("--cpu-test", type=int, default=48, help="CPU-baseline sample: test molecules")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fp64-engine", action="store_true", help="skip the extra FP64 DMMA engine measurement")
    ap.add_argument("--workload", default="config2", choices=["config2", "config3"],
                    help="config2 (default, the headline line) | config3: global SLATM Gaussian kernel, "
                         "row block of --rows-per-gpu molecules x --n-cols molecules per GPU, 4 sigmas")
    ap.add_argument("--n-cols", type=int, default=100000)
    ap.add_argument("--rows-per-gpu", type=int, default=12500)
    ap.add_argument("--kernel", default="gaussian", choices=["gaussian", "laplacian"])
    return ap.parse_args()


def make_sets(n_train, n_test, rank):
    from aqml_b200 import synth
    train = synth.qm9_like_fast(n_train, SEED, pool=512)
    test = synth.qm9_like_fast(n_test, SEED + 1000 * (rank + 1), pool=256)
    return train, test


def mbtypes_for(train, test):
    from aqml_b200 import synth
    from aqml_b200.representation import core as rcore
    zl = synth.split_zs(train[0][:2000], train[1][:int(train[0][:2000].sum())]) + \
        synth.split_zs(test[0][:500], test[1][:int(test[0][:500].sum())])
    # all five QM9 elements with full multiplicity (n1/n2/n3 = 5/15/75, D = 8975 at dgrid 0.04)
    zl.append(np.array([1, 1, 1, 6, 6, 6, 7, 7, 7, 8, 8, 8, 9, 9, 9]))
    return rcore.get_slatm_mbtypes(zl)


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
# CPU arm (restated reference loops)
# ------------------------------------------------------------------------------------------------
def cpu_sample(args, steps=1, warmup=0):
    """Time molecules -> aSLATM -> local Gaussian kernel on the host cores for a bounded sample."""
    from aqml_b200 import synth
    from oracle import c_oracle as co
    train, test = make_sets(args.cpu_train, args.cpu_test, 0)
    mb = mbtypes_for(*make_sets(2000, 500, 0))
    sig = np.array([c * np.full(ZMAX, 16.0) / np.sqrt(2 * np.log(2)) for c in COEFFS])
    co.load(fast=True)
    co.set_num_threads(len(os.sched_getaffinity(0)), fast=True)  # torchrun exports OMP_NUM_THREADS=1
    cores = co.num_threads(fast=True)

    def one():
        t0 = time.perf_counter()
        x1 = co.generate_slatm_batch(train[2], train[1], train[0], mb, True, fast=True, **SLATM_KW)
        x2 = co.generate_slatm_batch(test[2], test[1], test[0], mb, True, fast=True, **SLATM_KW)
        t1 = time.perf_counter()
        K = co.calc_lk_gaussian(x2, x1, test[1], train[1], test[0], train[0], ZMAX, False, sig, fast=True)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, K

    for _ in range(warmup):
        one()
    ts, tk = [], []
    for _ in range(max(steps, 1)):
        a, b, K = one()
        ts.append(a)
        tk.append(b)
    t_slatm, t_kernel = float(np.median(ts)), float(np.median(tk))
    elems = len(COEFFS) * args.cpu_test * args.cpu_train
    atoms = int(train[0].sum() + test[0].sum())
    return dict(value=elems / (t_slatm + t_kernel), unit="kernel elements/s", cores=cores, kind="port",
                sample=f"{args.cpu_test} test x {args.cpu_train} train QM9-shaped molecules ({atoms} atoms, D=8975, "
                       f"4 sigmas), molecules->aSLATM->K; restated reference loops (C/OpenMP -O3 -march=native), "
                       f"not gfortran",
                aslatm_atoms_per_sec=atoms / t_slatm, seconds=t_slatm + t_kernel,
                seconds_slatm=t_slatm, seconds_kernel=t_kernel)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    r = cpu_sample(args, steps=args.steps, warmup=min(args.warmup, 1))
    ms = r["seconds"] * 1e3
    line = {"impl": "reference", "metric": "local_gaussian_kernel_elements_per_sec", "value": r["value"],
            "unit": "kernel elements/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "config2: 20k train x 2k test QM9-shaped, local aSLATM element-matched Gaussian "
                                   "kernel, 4 sigmas (CPU arm: bounded sample of it)", "sample": r["sample"]},
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
    from aqml_b200 import fused

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        init_nccl(local_rank)
    lib = L.load()

    train, test = make_sets(args.n_train, args.n_test, rank)
    mb = mbtypes_for(train, test)
    S = len(COEFFS)
    dev = torch.device("cuda", local_rank)
    c1 = torch.as_tensor(train[2], device=dev)
    c2 = torch.as_tensor(test[2], device=dev)
    K = torch.empty((S, args.n_test, args.n_train), dtype=torch.float64, device=dev)

    # sigmas from the per-element dmax of a training subsample (width heuristic, aqml.py:204-255,775)
    nsub = min(args.n_train, 400)
    asub = int(train[0][:nsub].sum())
    sub = fused.PackedRep(train[2][:asub], train[1][:asub], train[0][:nsub], mb, **SLATM_KW)
    dso = fused.kernel_width(sub, sub, ZMAX)
    sub.close()
    sig = np.array([c * dso / np.sqrt(2 * np.log(2)) for c in COEFFS])

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step(timing=None):
        e = [ev() for _ in range(3)] if timing is not None else None
        if e:
            e[0].record()
        r1 = fused.PackedRep(c1, train[1], train[0], mb, **SLATM_KW)
        r2 = fused.PackedRep(c2, test[1], test[0], mb, **SLATM_KW)
        if e:
            e[1].record()
        fused.local_gaussian(r2, r1, sig, ZMAX, out=K)
        if e:
            e[2].record()
            timing.append(e)
        nbytes = r1.nbytes + r2.nbytes
        r1.close()
        r2.close()
        return nbytes

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        rep_bytes = step()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    import ctypes as C
    L.call("aqml_pair_stats", None, None, 1)  # reset the executed / dense k-tile counters
    launches0 = lib.aqml_launch_count()
    timing = []
    w0 = time.time()
    e_start, e_end = ev(), ev()
    e_start.record()
    for _ in range(args.steps):
        step(timing)
    e_end.record()
    barrier()
    w1 = time.time()
    launches = int(lib.aqml_launch_count() - launches0)
    kt_exec, kt_dense = C.c_int64(0), C.c_int64(0)
    L.call("aqml_pair_stats", C.c_void_p(C.addressof(kt_exec)), C.c_void_p(C.addressof(kt_dense)), 1)
    kt_exec, kt_dense = kt_exec.value / args.steps, kt_dense.value / args.steps  # per launch
    bm, bn, bk = C.c_int(0), C.c_int(0), C.c_int(0)
    lib.aqml_tile_shape(C.c_void_p(C.addressof(bm)), C.c_void_p(C.addressof(bn)), C.c_void_p(C.addressof(bk)))
    ktile_flops = 2.0 * bm.value * bn.value * bk.value
    t_total_ms = e_start.elapsed_time(e_end)
    t_slatm_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in timing]))
    t_kernel_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in timing]))
    clocks = sampler.stop(w0, w1) if sampler else None

    # ---- the FP64 DMMA engine on the same inputs (reported next to the default int8 tensor engine) --------------
    fp64_engine = None
    if rank == 0 and world == 1 and not args.no_fp64_engine and os.environ.get("AQML_PAIR_ENGINE", "") not in ("dmma", "fp64"):
        os.environ["AQML_PAIR_ENGINE"] = "dmma"
        try:
            for it in range(3):  # last pass timed: aSLATM alone (no digit planes are built for the FP64 engine)
                if it:
                    r1.close()
                    r2.close()
                torch.cuda.synchronize()
                s0, s1 = ev(), ev()
                s0.record()
                r1 = fused.PackedRep(c1, train[1], train[0], mb, **SLATM_KW)
                r2 = fused.PackedRep(c2, test[1], test[0], mb, **SLATM_KW)
                s1.record()
                torch.cuda.synchronize()
                t_slatm_only = s0.elapsed_time(s1)
            fused.local_gaussian(r2, r1, sig, ZMAX, out=K)
            torch.cuda.synchronize()
            L.call("aqml_pair_stats", None, None, 1)
            a, b = ev(), ev()
            a.record()
            for _ in range(2):
                fused.local_gaussian(r2, r1, sig, ZMAX, out=K)
            b.record()
            torch.cuda.synchronize()
            ke, kd = C.c_int64(0), C.c_int64(0)
            L.call("aqml_pair_stats", C.c_void_p(C.addressof(ke)), C.c_void_p(C.addressof(kd)), 1)
            fp64_engine = {"kernel_ms": a.elapsed_time(b) / 2, "ktiles_executed": ke.value / 2, "aslatm_ms": t_slatm_only}
            r1.close()
            r2.close()
        finally:
            del os.environ["AQML_PAIR_ENGINE"]

    # ---- end-to-end through the C ABI with host buffers -----------------------------------------
    e2e_ms = None
    if not args.no_e2e:
        kh = torch.empty((S, args.n_test, args.n_train), dtype=torch.float64, pin_memory=True)
        pc1 = torch.as_tensor(train[2]).pin_memory().numpy()
        pc2 = torch.as_tensor(test[2]).pin_memory().numpy()

        e2e_chunks = os.environ.get("AQML_E2E_CHUNKS", "0.75,0.25")  # measured: 1 -> 350.5, 2 -> 345.1, 0.75,0.25 -> 337.6, 0.7,0.2,0.1 -> 340.3 ms
        e2e_chunks = tuple(float(v) for v in e2e_chunks.split(",")) if "," in e2e_chunks else int(e2e_chunks)

        def e2e_step():
            r1 = fused.PackedRep(pc1, train[1], train[0], mb, **SLATM_KW)   # coords H2D inside
            if e2e_chunks != 1:  # queries in row blocks, D2H of block c-1 behind the contraction of block c
                fused.local_gaussian_to_host(pc2, test[1], test[0], r1, mb, sig, ZMAX, kh, chunks=e2e_chunks, **SLATM_KW)
                r1.close()
                return
            r2 = fused.PackedRep(pc2, test[1], test[0], mb, **SLATM_KW)
            fused.local_gaussian(r2, r1, sig, ZMAX, out=K)
            kh.copy_(K, non_blocking=True)                                    # K D2H
            r1.close()
            r2.close()
        for _ in range(2):
            e2e_step()
        barrier()
        a, b = ev(), ev()
        a.record()
        n_e2e = max(2, min(args.steps, 3))
        for _ in range(n_e2e):
            e2e_step()
        b.record()
        barrier()
        e2e_ms = a.elapsed_time(b) / n_e2e
        h2d = int(train[2].nbytes + test[2].nbytes + 4 * (len(train[1]) + len(test[1])) + sig.nbytes)
        d2h = int(K.numel() * 8)

    # ---- reduce over ranks ---------------------------------------------------------------------
    vals = torch.tensor([t_total_ms, t_slatm_ms, t_kernel_ms, e2e_ms or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    t_total_ms, t_slatm_ms, t_kernel_ms, e2e_max = [float(v) for v in vals.cpu()]
    ms_per_step = t_total_ms / args.steps
    elems_rank = S * args.n_test * args.n_train
    atoms_rank = int(train[0].sum() + test[0].sum())

    if rank == 0:
        # algorithmic work of the dominant kernel (pair_tile_kernel<DOT,LOCAL>): 2 * sum_z D_z n1_z n2_z
        p = L.slatm_params(mb, **{k: SLATM_KW[k] for k in SLATM_KW})
        nel = 5
        d_z = {z: nel * p.nx2 + (nel * (nel + 1) // 2) * p.nx3 for z in (1, 6, 7, 8, 9)}
        flops = 0.0
        for z in d_z:
            flops += 2.0 * d_z[z] * float((test[1] == z).sum()) * float((train[1] == z).sum())
        peak_tf = fp64_peak_tflops(torch)
        ach_tf = flops / (t_kernel_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # algorithmic bytes of the aSLATM phase: the packed FP64 operand 8 * sum_z D_z(padded to 16) * A_z; the
        # int8 digit planes written next to it (7 B per element) are an implementation detail of the default engine
        slatm_bytes = 0.0
        for z in d_z:
            slatm_bytes += 8.0 * (-(-d_z[z] // 16) * 16) * float((test[1] == z).sum() + (train[1] == z).sum())
        engine = "dmma" if os.environ.get("AQML_PAIR_ENGINE", "") in ("dmma", "fp64") else "i8"
        exec_flops = kt_exec * ktile_flops  # FP64-equivalent flops of the k-tiles actually contracted
        traffic = None
        try:  # dram__bytes_read+write of one launch, from the committed ncu --set full capture of this command
            key = "pair_tile_kernel<MODE_DOT,EPI_LOCAL>" if engine == "dmma" else "i8_pair_kernel"
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[key]
            if args.n_train == 20000 and args.n_test == 2000:
                traffic = float(tr["dram_bytes_read"] + tr["dram_bytes_write"])
        except Exception:
            pass
        roofline = {"bound": "tensor", "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf,
                    "traffic": traffic, "traffic_unit": "bytes of DRAM per launch (profiles/r01_traffic.json)",
                    "algorithmic_flops_per_launch": flops,
                    # all-zero 16-column k-tiles are skipped exactly: the work actually issued,
                    # incl. row/column padding, and its share of the dense sweep
                    "executed_flops_per_launch": exec_flops,
                    "executed_tflops": exec_flops / (t_kernel_ms * 1e-3) / 1e12,
                    "executed_frac_of_peak": exec_flops / (t_kernel_ms * 1e-3) / 1e12 / peak_tf,
                    "ktiles_executed_over_dense": kt_exec / max(kt_dense, 1.0),
                    "peak_source": "cuBLAS DGEMM 8192^3 best of 10 measured in this run (MEASURED_PEAKS.json has no FP64 entry)"}
        if engine == "dmma":
            roofline["kernel"] = "pair_tile_kernel<MODE_DOT,EPI_LOCAL> (FP64 DMMA.8x8x4)"
            roofline["note"] = ("achieved/frac use the SURVEY 8(d) algorithmic flop count 2*sum_z D_z*n1_z*n2_z; frac > 1 because "
                                "16-column k-tiles that are all-zero in either 64-row operand block are skipped (exact, "
                                "bit-identical; counted on the device). Hardware utilisation = executed_frac_of_peak; the FP64 "
                                "DMMA/DFMA ceiling of the chip is 37.0 TF (profiles/r01_micro_dmma_dfma.txt)")
        else:
            int8_ops = exec_flops * 28.0  # 28 digit-plane products per FP64 product
            int8_peak = 4.13  # POP/s, tcgen05 kind::i8 at N >= 128 measured on this pool (profiles/r01_micro_i8_umma.txt)
            roofline["kernel"] = "i8_pair_kernel (tcgen05.mma kind::i8, 7 x 7-bit digit planes, 28 exact int32 products in TMEM)"
            roofline["int8_pops"] = int8_ops / (t_kernel_ms * 1e-3) / 1e15
            roofline["int8_peak_pops"] = int8_peak
            roofline["int8_frac_of_peak"] = roofline["int8_pops"] / int8_peak
            roofline["int8_frac_of_n64_rate"] = roofline["int8_pops"] / 2.9
            if peaks.get("bf16_tflops"):  # MEASURED_PEAKS.json: cuBLAS bf16 dense; the int8 pipe does twice that per clock
                roofline["measured_bf16_tflops"] = peaks["bf16_tflops"]
                roofline["int8_frac_of_2x_measured_bf16"] = roofline["int8_pops"] * 1e3 / (2.0 * peaks["bf16_tflops"])
            roofline["note"] = ("FP64-accurate contraction from exact int8 digit-plane products on the 5th-generation tensor "
                                "cores (DESIGN.md 4.1b): achieved/frac are FP64-EQUIVALENT flops (SURVEY 8(d) count "
                                "2*sum_z D_z*n1_z*n2_z) against the measured cuBLAS DGEMM peak, so frac > 1 is the point of the "
                                "design (plus exact skipping of all-zero 16-column k-tiles). Hardware utilisation = "
                                "int8_frac_of_peak (28 int8 MMAs per FP64 k-tile; an N=64 UTCIMMA is capped at 2.9 POP/s by "
                                "shared-memory operand bandwidth, int8_frac_of_n64_rate). Same results to <= 1e-10 as the "
                                "FP64 DMMA engine (AQML_PAIR_ENGINE=dmma), asserted in tests/")
        line = {
            "metric": "local_gaussian_kernel_elements_per_sec",
            "value": world * elems_rank / (ms_per_step * 1e-3), "unit": "kernel elements/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "dtype_detail": ("FP64 in, FP64 out, <= 1e-10 vs the FP64 oracle; default engine: a.b from exact int8 digit-plane products "
                             "(int32 accumulators) on the tcgen05 tensor cores, norms / exp / sums in FP64; AQML_PAIR_ENGINE=dmma: "
                             "FP64 DMMA throughout") if engine == "i8" else "FP64 DMMA throughout",
            "data": "synthetic",
            "config": {"workload": "config2: 20k train x 2k test QM9-shaped molecules, local aSLATM element-matched "
                                   "Gaussian kernel, FP64, 4 sigmas; step = coords -> aSLATM(train+test) -> K",
                       "n_train": args.n_train, "n_test_per_gpu": args.n_test, "atoms_per_gpu": atoms_rank,
                       "D": int(lib.aqml_slatm_dim(p)), "D_z": d_z[6], "nsigmas": S,
                       "sharding": "row blocks of K (test molecules) per GPU, no collective",
                       "l2": "inputs larger than L2 (packed operands + digit planes %.1f GB per step)" % (rep_bytes / 1e9)},
            "aslatm_atoms_per_sec": world * atoms_rank / (t_slatm_ms * 1e-3),
            "phases_ms": {"aslatm": t_slatm_ms, "kernel": t_kernel_ms},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_aslatm": {"bound": "hbm", "kernel": "slatm_units_kernel (packed output) + digit-plane split",
                                "achieved": slatm_bytes / (t_slatm_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                "frac": slatm_bytes / (t_slatm_ms * 1e-3) / 1e9 / hbm_peak,
                                "algorithmic_bytes_per_step": slatm_bytes,
                                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
        }
        if fp64_engine is not None:
            ef = fp64_engine["ktiles_executed"] * ktile_flops
            ms = fp64_engine["kernel_ms"]
            line["roofline_fp64_engine"] = {
                "kernel": "pair_tile_kernel<MODE_DOT,EPI_LOCAL> (FP64 DMMA.8x8x4), AQML_PAIR_ENGINE=dmma, same inputs, kernel only",
                "kernel_ms": ms, "bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": flops / (ms * 1e-3) / 1e12 / peak_tf, "executed_tflops": ef / (ms * 1e-3) / 1e12,
                "executed_frac_of_peak": ef / (ms * 1e-3) / 1e12 / peak_tf,
                "speedup_of_default_engine": ms / t_kernel_ms,
                "aslatm_ms_without_digit_planes": fp64_engine["aslatm_ms"],
                "aslatm_atoms_per_sec_without_digit_planes": atoms_rank / (fp64_engine["aslatm_ms"] * 1e-3),
                "note": "the north_star's literal design point (FP64 DMMA tiles): 88 % of the cuBLAS DGEMM peak, tensor pipe 84 % "
                        "active (profiles/r01_pair_tile_kernel_ncu.txt); kept as the fallback engine"}
        if e2e_ms is not None:
            line["e2e"] = {"value": world * elems_rank / (e2e_max * 1e-3), "unit": "kernel elements/s",
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_max}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_sample(args, steps=1, warmup=0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["cpu_baseline"]["aslatm_atoms_per_sec"] = r["aslatm_atoms_per_sec"]
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_config3(args):
    """BASELINE.json configs[2]: 100k x 100k global SLATM kernel with 4 sigmas, row-block sharded: every rank
    builds K[:, rows_r, :] for its own --rows-per-gpu molecules against all --n-cols molecules (no collective).
    Step = coordinates -> global SLATM (rows + columns) -> K block resident in HBM."""
    import torch
    import torch.distributed as dist
    from aqml_b200 import _lib as L
    from aqml_b200 import kernels, synth
    from aqml_b200.representation import core as rcore
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        init_nccl(local_rank)
    dev = torch.device("cuda", local_rank)
    cols = synth.qm9_like_fast(args.n_cols, SEED + 7, pool=512)
    rows = synth.qm9_like_fast(args.rows_per_gpu, SEED + 7 + 1000 * (rank + 1), pool=256)
    mb = mbtypes_for(cols, rows)
    kw = dict(sigmas=[.05, .05], dgrids=[.04, .04], rcut=4.8)
    cc, cr = torch.as_tensor(cols[2], device=dev), torch.as_tensor(rows[2], device=dev)
    xs = rcore.generate_slatm_batch(cc[:int(cols[0][:2000].sum())], cols[1][:int(cols[0][:2000].sum())], cols[0][:2000],
                                    mb, local=False, **kw)
    from aqml_b200 import distance
    dmax = distance.max_distance(xs, xs, 2 if args.kernel == "gaussian" else 1)
    del xs
    fac = 1.0 / np.sqrt(2 * np.log(2)) if args.kernel == "gaussian" else 1.0 / np.log(2)
    sig = np.array([c * dmax * fac for c in COEFFS])
    D = int(L.load().aqml_slatm_dim(L.slatm_params(mb, **SLATM_KW)))
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    # SLATM of the column molecules is sharded by molecules over the ranks, then one all-gather
    from aqml_b200 import parallel
    shard = world > 1 and not os.environ.get("AQML_NO_SLATM_SHARD")
    if shard:
        parts = [parallel.shard_range(args.n_cols, r, world) for r in range(world)]
        clo, chi = parts[rank]
        coff = np.concatenate(([0], np.cumsum(cols[0])))
        cc_loc, zs_loc, nas_loc = cc[coff[clo]:coff[chi]], cols[1][coff[clo]:coff[chi]], cols[0][clo:chi]
        counts = [h - l for l, h in parts]

    def step(tm=None):
        e = [ev() for _ in range(3)]
        e[0].record()
        if shard:
            xc = parallel.allgather_rows(rcore.generate_slatm_batch(cc_loc, zs_loc, nas_loc, mb, local=False, **kw), counts)
        else:
            xc = rcore.generate_slatm_batch(cc, cols[1], cols[0], mb, local=False, **kw)
        xr = rcore.generate_slatm_batch(cr, rows[1], rows[0], mb, local=False, **kw)
        e[1].record()
        K = kernels.global_kernels(xr, xc, sig, args.kernel)
        e[2].record()
        if tm is not None:
            tm.append(e)
        return K

    for _ in range(max(args.warmup, 3)):
        K = step()
        del K
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tm = []
    a, b = ev(), ev()
    a.record()
    for _ in range(args.steps):
        K = step(tm)
        del K
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b), float(np.mean([e[0].elapsed_time(e[1]) for e in tm])),
                      float(np.mean([e[1].elapsed_time(e[2]) for e in tm]))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot, t_sl, t_k = [float(v) for v in t.cpu()]
    if rank == 0:
        ms = tot / args.steps
        S = len(COEFFS)
        ops = 2.0 * D * args.rows_per_gpu * args.n_cols
        print(json.dumps({
            "metric": "global_%s_kernel_elements_per_sec" % args.kernel,
            "value": world * S * args.rows_per_gpu * args.n_cols / (ms * 1e-3), "unit": "kernel elements/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config3: global SLATM %s kernel, 4 sigmas, %d rows per GPU x %d columns, D=%d; "
                                   "K row-block sharded%s" % (args.kernel, args.rows_per_gpu, args.n_cols, D,
                                                           "; column SLATM generated 1/N per rank + one NCCL all-gather" if shard
                                                           else ", no collective")},
            "phases_ms": {"slatm": t_sl, "kernel_incl_pack": t_k},
            "slatm_molecules_per_sec": world * (args.n_cols + args.rows_per_gpu) / (t_sl * 1e-3),
            "kernel_tera_ops": ops / (t_k * 1e-3) / 1e12}))
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
