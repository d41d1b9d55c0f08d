#!/usr/bin/env python3
"""Benchmark of the J+K Fock build (BASELINE.json metric: Fock build time / iteration, shell quartets/s,
% of FP64 peak) -- see DESIGN.md "Measurement".

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload w32] [--impl ours|reference]

A step = one full RHF J+K Fock build (all surviving shell quartets of the workload) for one synthetic density.
N > 1 (torchrun): the bra shell-pair list is split cyclically over ranks (int2.F90:759-761), partial Fock
matrices are summed with ONE NCCL all-reduce (int2.F90:1396); strong scaling (total work fixed).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="w32")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cam", action="store_true", help="range-separated two-pass build (CAM-B3LYP: alpha 0.19, beta 0.46, mu 0.33)")
    ap.add_argument("--cutoff", type=float, default=5e-11, help="integral cutoff (MRSF workloads; the reference's response default is 1e-8, types.F90:185)")
    ap.add_argument("--nvec", type=int, default=12, help="MRSF workloads (c5): Davidson trial vectors per build (x 7 densities)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for n, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def host_threads():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must not
    inherit that (a 1-thread reference would inflate every N > 1 ratio), so the oracle is given an explicit count."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def workload_config(args, bs, W, sx, nvec=None):
    """The keys that identify the workload: identical in the `ours` and `reference` arms."""
    cfg = {"workload": W.WORKLOADS.get(args.workload, args.workload), "nshell": int(bs.nshell), "nbf": int(bs.nbf),
           "cutoff": 5e-11, "scale_exchange": sx,
           "density": ("synthetic general (non-symmetric) decaying, seed 7" if nvec else "synthetic decaying, seed 7"),
           "cam": dict(CAM, passes=2) if args.cam else None}
    if nvec:
        cfg.update({"nvec": nvec, "densities": nvec * 7})
    return cfg


def dgemm_peak_tflops(dev, n=6144, reps=3):
    """Independent cross-check of the FP64 roofline denominator: cuBLAS DGEMM through torch (informational; `roofline.peak`
    stays the FMA microbenchmark of the library, which this number should not exceed by much).  None when it cannot run."""
    try:
        import torch
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        torch.matmul(a, b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(reps):
            torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(dev)
        return 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    except Exception:
        return None


def dram_traffic(workload):
    """HBM bytes per build summed over the ERI launches, from the committed ncu pass
    (profiles/r02_dram_<workload>.json, written by tools/dram_traffic.py on the GPU box); None when not captured."""
    p = os.path.join(ROOT, "profiles", f"r02_dram_{workload}.json")
    try:
        return json.load(open(p))["dram_bytes_per_build"]
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


CAM = {"alpha": 0.19, "beta": 0.46, "mu": 0.33}  # CAM-B3LYP


def cpu_sample(bs, d_packed, sx, target_s, nthreads=0, cam=False):
    nthreads = nthreads or host_threads()
    """Oracle (C++/OpenMP restatement of int2_twoei, the reference's OpenMP CPU path) on a bounded,
    strided sample of the cost-sorted bra shell-pair list of the SAME workload."""
    from oracle.oracle import Oracle, max_threads, set_fast_rys
    set_fast_rys(True)  # roots from polynomial tables, as stock OpenQP does for nroots <= 5 (timed baseline only)
    o = Oracle(bs)
    o.set_screening()
    npair = bs.nshell * (bs.nshell + 1) // 2
    # probe with a sparse stride, then size the sample for ~target_s seconds
    stride = max(1, npair // 400)

    def build(stride):
        if cam:
            _, st = o.fock_cam(d_packed, CAM["alpha"], CAM["beta"], CAM["mu"], nthreads=nthreads, stride=stride, offset=1 % stride)
            return dict(st, nquartets=st["nquartets_both_passes"])
        return o.fock(d_packed, sx, 1.0, nthreads=nthreads, stride=stride, offset=1 % stride)[1]

    if cam:
        o.schwarz_attenuated(CAM["mu"])  # setup, not part of the timed builds
    t = time.perf_counter()
    st = build(stride)
    dt = time.perf_counter() - t
    est_full = dt * stride
    stride2 = max(1, int(round(est_full / target_s)))
    if stride2 < stride:
        t = time.perf_counter()
        st = build(stride2)
        dt = time.perf_counter() - t
        stride = stride2
    return {"quartets": st["nquartets"], "seconds": dt, "stride": stride, "cores": nthreads}


MRSF_WORKLOADS = ("c5",)

_SAVED_STDOUT = None


def quiet_stdout():
    """Multi-rank runs: libraries (NCCL prints its version line) must not write to stdout, which carries only the one JSON
    line -- point fd 1 at stderr for the duration of the run and keep the real stdout for emit()."""
    global _SAVED_STDOUT
    if _SAVED_STDOUT is None:
        sys.stdout.flush()
        _SAVED_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    txt = json.dumps(line) + "\n"
    if _SAVED_STDOUT is None:
        sys.stdout.write(txt)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_SAVED_STDOUT, txt.encode())


def mrsf_cpu_sample(bs, d3, sx, target_s, nthreads=0, cutoff=5e-11):
    nthreads = nthreads or host_threads()
    """Oracle int2_mrsf_data_t build (tdhf_mrsf_lib.F90:218-333) on a strided sample of the bra shell-pair list."""
    from oracle.oracle import Oracle, set_fast_rys
    set_fast_rys(True)  # see cpu_sample
    o = Oracle(bs, cutoff)
    o.set_screening()
    npair = bs.nshell * (bs.nshell + 1) // 2
    stride = max(1, npair // 200)
    t = time.perf_counter()
    _, st = o.mrsf(d3, sx, 1.0, nthreads=nthreads, stride=stride, offset=1 % stride)
    dt = time.perf_counter() - t
    stride2 = max(1, int(round(dt * stride / target_s)))
    if stride2 < stride:
        t = time.perf_counter()
        _, st = o.mrsf(d3, sx, 1.0, nthreads=nthreads, stride=stride2, offset=1 % stride2)
        dt = time.perf_counter() - t
        stride = stride2
    return {"quartets": st["nquartets"], "seconds": dt, "stride": stride, "cores": nthreads}


def main_mrsf(args):
    """config 5: one step = one int2_mrsf_data_t build, nvec x 7 general densities digested in one pass over the
    surviving quartets (batched multi-density J/K for the MRSF Davidson trial vectors)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from openqp_b200 import workloads as W
    from openqp_b200.int2 import Int2Compute, lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        quiet_stdout()
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    mol, bs = W.build(args.workload)
    nvec, ncomp, sx = args.nvec, 7, 0.5  # BHHLYP: 50 % exact exchange
    d3 = W.mrsf_densities(bs, nvec, ncomp)
    d3f = np.ascontiguousarray(np.transpose(d3, (3, 2, 1, 0)))  # Fortran d3(v, c, mu, nu), v fastest
    drv = Int2Compute(local).init(bs, args.cutoff)
    t0 = time.perf_counter()
    drv.set_screening()
    t_screen = time.perf_counter() - t0
    drv.set_partition(rank, world)
    stream = torch.cuda.current_stream(dev)
    drv.set_stream(stream.cuda_stream)
    fp64_peak = drv.fp64_peak_tflops()
    d_dev = torch.from_numpy(d3f).to(dev)
    f_dev = torch.zeros_like(d_dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def step_dev():
        drv.mrsf_dev(d_dev.data_ptr(), f_dev.data_ptr(), nvec, ncomp, scale_exchange=sx, scale_coulomb=1.0)
        if world > 1:
            dist.all_reduce(f_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot_ms, kernel_ms, flops, nq, launches = 0.0, 0.0, 0.0, 0, 0
    for _ in range(args.steps):
        flush.fill_(1.0)
        barrier()
        ev0.record(stream)
        step_dev()
        ev1.record(stream)
        barrier()
        tot_ms += ev0.elapsed_time(ev1)
        st = drv.last_stats()
        kernel_ms += st["kernel_ms"]; flops += st["flops"]; nq += st["nquartets"]; launches += st["launches"] + 3
    t = torch.tensor([tot_ms, float(nq), flops, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tot_ms, kernel_ms = float(tmax[0]), float(tmax[3])
        nq_all, flops_all = float(tsum[1]), float(tsum[2])
    else:
        nq_all, flops_all = float(nq), flops
    clocks = sampler.stop() if rank == 0 else None

    # e2e: the reference-facing call with HOST buffers (pinned): H2D of d3, build, D2H of f3 inside the timed region
    d_pin = torch.from_numpy(d3f).pin_memory()
    f_pin = torch.empty_like(d_pin)
    ns = C.c_longlong(0)

    def step_host():
        if world == 1:
            rc = lib().oqpb_jk_mrsf(drv._h, C.c_void_p(d_pin.data_ptr()), nvec, ncomp, C.c_double(sx), C.c_double(1.0),
                                    C.c_void_p(f_pin.data_ptr()), C.byref(ns))
            assert rc == 0
        else:
            d_dev.copy_(d_pin, non_blocking=True)
            step_dev()
            f_pin.copy_(f_dev, non_blocking=True)
            torch.cuda.synchronize(dev)

    step_host()
    barrier()
    e2e_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1.0)
        barrier()
        ev0.record(stream)
        step_host()
        ev1.record(stream)
        barrier()
        e2e_ms += ev0.elapsed_time(ev1)
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0])
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        ms_per_step = tot_ms / args.steps
        nq_step = nq_all / args.steps
        achieved = flops_all / world / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
        nbytes = d3f.nbytes
        line = {
            "metric": "shell_quartets_per_s", "value": nq_step / (ms_per_step * 1e-3), "unit": "quartets/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args, bs, W, sx, nvec), cutoff=args.cutoff),
            "detail": {"quartets_per_build": nq_step, "mrsf_builds_per_s": 1e3 / ms_per_step,
                       "l2": "flushed between steps (256 MB fill)", "schwarz_setup_s": t_screen,
                       "parallelism": f"bra shell pairs cyclic over {world} GPU(s), 1 NCCL all-reduce of f3"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None, "traffic": dram_traffic(args.workload),
                         "kernel": "eri_*_kernel family, MODE_GEN: Rys ERI + DMMA m8n8k4 multi-density digestion",
                         "peak_source": "FP64 FMA microbenchmark in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "algorithmic_flops_per_step": flops_all / args.steps, "kernel_ms_per_step": kernel_ms / args.steps,
                         "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_source": peak_kind},
            "e2e": {"value": nq_step / (e2e_ms / args.steps * 1e-3), "unit": "quartets/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes},
            "gpu_launches": launches, "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            r = mrsf_cpu_sample(bs, d3, sx, args.cpu_seconds, cutoff=args.cutoff)
            line["cpu_baseline"] = {"value": r["quartets"] / r["seconds"], "unit": "quartets/s", "cores": r["cores"], "kind": "port",
                                    "sample": f"every {r['stride']}-th bra shell pair, {r['quartets']} quartets in {r['seconds']:.1f} s"}
        emit(line)
    drv.clean()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference binary cannot be built here
    (Fortran + network-only externals), so this times the oracle port with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from openqp_b200 import workloads as W
    mol, bs = W.build(args.workload)
    sx = W.scale_exchange(args.workload)
    if args.workload in MRSF_WORKLOADS:
        sx = 0.5  # BHHLYP, as in main_mrsf
        dp = W.mrsf_densities(bs, args.nvec)
        sampler = lambda *a, **k: mrsf_cpu_sample(*a, cutoff=args.cutoff, **k)
    else:
        from openqp_b200.scf import pack
        dp = pack(W.synthetic_density(bs))
        sampler = cpu_sample
    times, quartets, stride, cores = [], 0, 1, 1
    for it in range(args.warmup + args.steps):
        kw = {"cam": True} if (args.cam and args.workload not in MRSF_WORKLOADS) else {}
        r = sampler(bs, dp, sx, args.cpu_seconds if it >= args.warmup else min(args.cpu_seconds, 3.0), **kw)
        if it >= args.warmup:
            times.append(r["seconds"]); quartets += r["quartets"]; stride = r["stride"]; cores = r["cores"]
    tot = sum(times)
    val = quartets / tot
    line = {"impl": "reference", "metric": "shell_quartets_per_s", "value": val, "unit": "quartets/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": (dict(workload_config(args, bs, W, sx, args.nvec), cutoff=args.cutoff) if args.workload in MRSF_WORKLOADS
                       else workload_config(args, bs, W, sx)),
            "cpu_baseline": {"value": val, "unit": "quartets/s", "cores": cores, "kind": "port",
                             "note": "Rys-only C++/OpenMP port of int2_twoei, Rys roots from polynomial tables; stock OpenQP (rotated-axis s/p/d code, libint for f) is faster",
                             "sample": f"every {stride}-th bra shell pair of the cost-sorted list (int2.F90:864-921), "
                                       f"{quartets} quartets per {args.steps} steps"},
            "e2e": {"value": val, "unit": "quartets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload in MRSF_WORKLOADS:
        return main_mrsf(args)
    import torch
    import torch.distributed as dist
    from openqp_b200 import workloads as W
    from openqp_b200.int2 import Int2Compute, Int2RhfData
    from openqp_b200.scf import pack

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        quiet_stdout()
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    mol, bs = W.build(args.workload)
    sx = W.scale_exchange(args.workload)
    d = pack(W.synthetic_density(bs))
    drv = Int2Compute(local).init(bs)
    t0 = time.perf_counter()
    drv.set_screening()
    t_screen = time.perf_counter() - t0
    drv.set_partition(rank, world)
    stream = torch.cuda.current_stream(dev)
    drv.set_stream(stream.cuda_stream)
    fp64_peak = drv.fp64_peak_tflops()

    d_dev = torch.from_numpy(d).to(dev)
    f_dev = torch.zeros_like(d_dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    if args.cam:
        drv.set_screening_cam(CAM["mu"])  # attenuated Schwarz matrix: once per geometry, like set_screening

    ar_events = []

    def step_dev(timed=False):
        if args.cam:
            drv.fock_cam_dev(d_dev.data_ptr(), f_dev.data_ptr(), 1, CAM["alpha"], CAM["beta"], CAM["mu"])
        else:
            drv.fock_dev(d_dev.data_ptr(), f_dev.data_ptr(), 1, scale_exchange=sx)
        if world > 1:
            if timed:
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(stream)
            dist.all_reduce(f_dev)
            if timed:
                a1.record(stream)
                ar_events.append((a0, a1))
        drv.fock_post_dev(f_dev.data_ptr(), 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, flops, nq, launches = 0.0, 0.0, 0, 0
    # ---- timed region: K steps, inputs resident in HBM, L2 flushed between steps
    barrier()
    tot_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1.0)
        barrier()
        ev0.record(stream)
        step_dev(timed=True)
        ev1.record(stream)
        barrier()
        tot_ms += ev0.elapsed_time(ev1)
        st = drv.last_stats()
        kernel_ms += st["kernel_ms"]; flops += st["flops"]; nq += st["nquartets"]; launches += st["launches"] + 8
    ar_ms = sum(a0.elapsed_time(a1) for a0, a1 in ar_events)
    t = torch.tensor([tot_ms, float(nq), flops, kernel_ms, ar_ms], dtype=torch.float64, device=dev)
    rank_stats = None
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        # per-rank ERI kernel time (imbalance) and the all-reduce as seen by the fastest rank (the slowest rank's
        # all-reduce time is the collective itself, the others' includes the wait for it)
        rank_stats = {"kernel_ms_min": float(tmin[3]) / args.steps, "kernel_ms_mean": float(tsum[3]) / world / args.steps,
                      "kernel_ms_max": float(tmax[3]) / args.steps, "allreduce_ms": float(tmin[4]) / args.steps,
                      "allreduce_plus_wait_ms_max": float(tmax[4]) / args.steps}
        tot_ms, kernel_ms = float(tmax[0]), float(tmax[3])
        nq_all, flops_all = float(tsum[1]), float(tsum[2])
    else:
        nq_all, flops_all = float(nq), flops
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the reference-facing call with HOST buffers (oqpb_fock: H2D of D, build, D2H of F) per step
    d_pin = torch.from_numpy(d).pin_memory()
    f_pin = torch.empty_like(d_pin)
    import ctypes as C
    from openqp_b200.int2 import lib
    ns = C.c_longlong(0)

    def step_host():
        if world == 1 and args.cam:
            rc = lib().oqpb_fock_cam(drv._h, 0, C.c_void_p(d_pin.data_ptr()), C.c_void_p(f_pin.data_ptr()), 1,
                                     C.c_double(CAM["alpha"]), C.c_double(CAM["beta"]), C.c_double(CAM["mu"]),
                                     C.c_double(1.0), C.c_double(0.0), 1, C.byref(ns))
            assert rc == 0
        elif world == 1:
            rc = lib().oqpb_fock(drv._h, 0, C.c_void_p(d_pin.data_ptr()), C.c_void_p(f_pin.data_ptr()), 1, C.c_double(sx),
                                 C.c_double(1.0), 1, C.byref(ns))
            assert rc == 0
        else:
            d_dev.copy_(d_pin, non_blocking=True)
            step_dev()
            f_pin.copy_(f_dev, non_blocking=True)
            torch.cuda.synchronize(dev)

    step_host()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1.0)
        barrier()
        e0.record(stream)
        step_host()
        e1.record(stream)
        barrier()
        e2e_ms += e0.elapsed_time(e1)
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0])

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        ms_per_step = tot_ms / args.steps
        nq_step = nq_all / args.steps
        value = nq_step / (ms_per_step * 1e-3)
        achieved = flops_all / world / (kernel_ms * 1e-3) / 1e12 if kernel_ms > 0 else 0.0
        ntri = bs.ntri
        line = {
            "metric": "shell_quartets_per_s", "value": value, "unit": "quartets/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, bs, W, sx),
            "detail": {"quartets_per_build": nq_step, "fock_builds_per_s": 1e3 / ms_per_step,
                       "l2": "flushed between steps (256 MB fill)", "schwarz_setup_s": t_screen, "ranks": rank_stats,
                       "parallelism": f"bra shell pairs cyclic over {world} GPU(s), 1 NCCL all-reduce of the packed Fock"},
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None, "traffic": dram_traffic(args.workload),
                         "traffic_note": "HBM bytes per build over all ERI launches (ncu dram__bytes_read+write, profiles/); "
                                         "algorithmic bytes = 16 ntri + pair table, the bound is FP64",
                         "kernel": "eri_kernel<la,lb,lc,ld> family (Rys ERI + fused J/K digestion), per-GPU average",
                         "peak_source": "FP64 FMA microbenchmark in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "peak_cublas_dgemm_tflops": dgemm_peak_tflops(dev) if rank == 0 else None,
                         "algorithmic_flops_per_step": flops_all / args.steps, "kernel_ms_per_step": kernel_ms / args.steps,
                         "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_source": peak_kind},
            "e2e": {"value": nq_step / (e2e_ms / args.steps * 1e-3), "unit": "quartets/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": 8 * ntri, "d2h_bytes_per_step": 8 * ntri},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            r = cpu_sample(bs, d, sx, args.cpu_seconds, cam=args.cam)
            line["cpu_baseline"] = {"value": r["quartets"] / r["seconds"], "unit": "quartets/s", "cores": r["cores"], "kind": "port",
                                    "sample": f"every {r['stride']}-th bra shell pair of the cost-sorted list, "
                                              f"{r['quartets']} quartets in {r['seconds']:.1f} s",
                                    "note": "Rys-only C++/OpenMP port of int2_twoei, Rys roots from polynomial tables; "
                                            "stock OpenQP uses rotated-axis s/p/d code and libint for f and is faster"}
        emit(line)
    drv.clean()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
