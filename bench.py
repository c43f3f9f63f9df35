#!/usr/bin/env python
"""bench.py -- ALM inner iterations per second on the BASELINE.json MaxCut configuration.

A *step* is one ALM inner iteration of the LoRADS hot path (L-BFGS direction, the fused A(sym(R D^T)) / A(D D^T)
evaluation with objective, exact quartic line search, variable update, gradient 2 (C + A^*(w)) R, L-BFGS history
update, A(R R^T) + primal infeasibility; reference src_semi/lorads_alg/lorads_alm.c:1073-1146) on the synthetic
MaxCut SDP of BASELINE.json configs[1] (n = m = 100 000, 500 000 edges, seed 3, rank 24), at fixed rho from the
reference's random start (srand(925)).

  python bench.py --gpus N --steps K --warmup W                 our CUDA path (N > 1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W  the reference's own CPU code on the host cores

`value`   : iterations/s with everything resident in HBM (CUDA events on the solver's stream, max over ranks); the
            K-step region is repeated (`repeats`) and the median region is reported
`e2e`     : the same metric through the host-buffer C-ABI call lb2_alm_run_host (R and lambda copied host->device,
            K iterations, R copied back; wall clock, median of the repeats)
`roofline`: the A(UV^T) gather pass of the step (north_star kernel): algorithmic bytes / CUDA-event time of a launch that
            starts from a flushed L2, vs MEASURED_PEAKS.json; the adjoint SpMM and the other kernels beside it
`cpu_baseline`: the compiled reference (oracle/_ref) timed on this box's host cores on a bounded sample
"""
from __future__ import annotations

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

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on (default, fits one GPU)
    "cfg2": dict(kind="maxcut", n=100_000, edges=500_000, seed=3, label="BASELINE.json configs[1]"),
    # BASELINE.json configs[4]: n = 1e6, factors (224 MB each) no longer fit the 126 MB L2 (extra evidence only)
    "cfg5": dict(kind="maxcut", n=1_000_000, edges=5_000_000, seed=5, label="BASELINE.json configs[4]"),
}
WORKLOAD = WORKLOADS["cfg2"]
FALLBACK_HBM_GBS = 6650.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_instance():
    from lorads_b200 import sdpa
    return sdpa.maxcut(WORKLOAD["n"], WORKLOAD["edges"], WORKLOAD["seed"])


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def mark(self):
        """Start of the timed region: nvidia-smi is already running (it needs ~0.2 s to deliver its first line)."""
        self.t0 = time.time()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.time()
        deadline = t1 + 1.5
        while not self.lines and time.time() < deadline:      # very short timed region: wait for the first sample
            time.sleep(0.02)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = getattr(self, "t0", 0.0)
        inside = [ln for (ts, ln) in self.lines if t0 <= ts <= t1 + 0.05]
        if not inside and self.lines:
            inside = [self.lines[-1][1]]      # region shorter than the sampling period: closest sample (still under load)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_label: str):
    """DRAM bytes per launch of a kernel from the committed ncu --set full capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    try:
        table = json.load(open(p))["dram_bytes_per_launch"]
    except Exception:
        return None
    key = {"vc_auv_kernel<TRI>": "vc_auv_kernel<4,", "vc_auv_kernel<DUAL>": "vc_auv_kernel<2,", "vc_auv_kernel<SAME>": "vc_auv_kernel<0,",
           "vc_spmm_kernel": "vc_spmm_kernel", "axpby_dot2_kernel": "axpby_dot2_kernel",
           "auv_items_kernel<TRI>": "auv_items_kernel<4,", "auv_items_kernel<DUAL>": "auv_items_kernel<2,",
           "auv_items_kernel<SAME>": "auv_items_kernel<0,", "spmm_sym_kernel": "spmm_sym_kernel", "wsum_kernel": "wsum_kernel"}
    for prefix, ncu_prefix in key.items():
        if kernel_label.startswith(prefix):
            for name, v in table.items():
                if name.startswith(ncu_prefix):
                    return v
    return None


def kernel_bytes(S) -> dict:
    """Algorithmic (compulsory) bytes per launch of each hot kernel: every array the kernel has to read or write,
    counted once (factors 8*n*ld each, index/value arrays of the layout, m-vectors); formulas in DESIGN.md section 4."""
    n, ld, m = S.dim(0), S.info(17), S.m
    itAC, itA, nact = S.info(15), S.info(16), S.info(8)
    npat, nadj, nnzA = S.info(4), S.info(14), S.info(12)
    N = S.info(18)
    tri = S.info(20) == 1
    nout = 3 if tri else 2
    if S.info(21) == 1:
        # vertex-centric layout: adjacency entry = neighbour (4) + value (8) [+ constraint tag (4) where the kernel reads it],
        # per row two pointers + the row order (12), diagonal-singleton pair (12), lower-triangular singleton entry (16)
        nu, ndyn, nl = S.info(22), S.info(23), S.info(24)
        nsta = nu - ndyn
        auv_idx = 12 * nsta + 12 * n + 12 * n + 16 * nl
        return {
            0: (f"vc_auv_kernel<{'TRI' if tri else 'DUAL'}> A(sym(RD^T)),A(DD^T){',A(RR^T)' if tri else ''}+obj",
                2 * 8 * n * ld + auv_idx + nout * 8 * (nact + 1)),
            1: ("vc_auv_kernel<SAME> A(RR^T)", 8 * n * ld + 12 * n + 12 * n + 16 * nl + 8 * nact),
            2: ("wsum_kernel (multi-entry constraints only)", 0),
            3: ("vc_spmm_kernel G=2(C+A^*(w))R", 16 * nu + 12 * n + 8 * nact + 2 * 8 * n * ld),
            4: ("axpby_dot2_kernel (L-BFGS pass)", 4 * 8 * N),
        }
    dual = (("auv_items_kernel<TRI> A(sym(RD^T)),A(DD^T),A(RR^T)+obj", 2 * 8 * n * ld + 16 * itAC + 4 * (nact + 2) + 3 * 8 * (nact + 1))
            if tri else
            ("auv_items_kernel<DUAL> A(sym(RD^T)),A(DD^T)+obj", 2 * 8 * n * ld + 16 * itAC + 4 * (nact + 2) + 2 * 8 * (nact + 1)))
    return {
        0: dual,
        1: ("auv_items_kernel<SAME> A(RR^T)", 8 * n * ld + 16 * itA + 4 * (nact + 1) + 8 * nact),
        2: ("wsum_kernel S=C+A^*(w)", npat * (8 + 8 + 4) + nnzA * (4 + 8 + 8)),
        3: ("spmm_sym_kernel G=2SR", nadj * 8 + npat * 8 + 4 * (n + 1) + 2 * 8 * n * ld),
        4: ("axpby_dot2_kernel (L-BFGS pass)", 4 * 8 * N),
    }


def shard_description(world, rows):
    if world <= 1:
        return ""
    if rows:
        return (f"row slabs of the factors over {world} GPUs (cone blocks / rows with their pattern entries): NCCL all-reduce of the "
                "m-vectors per A() evaluation, scalar all-reduce per dot table, all-gather of the new direction")
    return f"factor columns sharded over {world} GPUs, NCCL all-reduce per A() evaluation and per dot"


def workload_config(rank_r, world, mode):
    """The `config` object of the JSON line: identical for our arm and the reference arm."""
    return {"workload": f"MaxCut SDP n=m={WORKLOAD['n']} edges={WORKLOAD['edges']} seed={WORKLOAD['seed']} rank={rank_r} ({WORKLOAD['label']}); one step = one ALM inner iteration, lorads_alm.c:1073-1146",
            "l2": "step timing: no explicit flush, one step streams ~12 factor-sized vectors (19 MB each) plus 30 MB of index data (> 126 MB L2 in total); per-kernel roofline timing: L2 flushed (384 MB read through it) before every launch",
            "parallelism": "single GPU" if world == 1 else mode}


class c_stdout_to_stderr:
    """The reference prints progress with printf; keep our stdout to the single JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_reference_rate(inst, iters: int):
    """ALM inner iterations/s of the compiled reference (oracle/_ref, 64-bit build) on this host; 1 thread."""
    from lorads_b200 import sdpa
    from oracle import ref
    if not ref.available(64):
        return None
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    d = tempfile.mkdtemp(prefix="lorads_bench_")
    path = os.path.join(d, inst.name + ".dat-s")
    sdpa.write_dat_s(inst, path)
    with c_stdout_to_stderr():
        R = ref.RefSolver(path, 64)
        rho = R.dinfo(6)
        sec = R.time_alm_inner_iters(rho, iters)
        rank = R.rank()
    return iters / sec, sec, rank


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    inst = make_instance()
    # bounded sample: at most 60 timed and 10 warm-up iterations of the reference (0.33 s each on one core)
    steps = max(1, min(args.steps, 60))
    warm = min(max(3, args.warmup), 10)          # our arm always warms up at least 3 steps
    t0 = time.time()
    out = cpu_reference_rate(inst, warm + steps)
    if out is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libloradsref64.so missing (build with make -C oracle ref where /root/reference exists)"}))
        return 0
    rate, sec, r = out
    line = {
        "impl": "reference", "metric": "alm_inner_iterations_per_second", "value": rate, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 / rate, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the same config object as our arm prints at this N (the reference itself runs on the host cores of rank 0)
        "config": workload_config(r, args.gpus, shard_description(args.gpus, os.environ.get("LORADS_B200_SHARD") == "rows")),
        "cpu_baseline": {"value": rate, "unit": "iterations/s", "cores": 1, "kind": "reference",
                         "sample": f"{warm + steps} inner iterations of the untouched reference (oracle/_ref, -DMAC_INT64, OpenBLAS 1 thread) from the srand(925) start, {sec:.1f} s; host has {os.cpu_count()} cores, the reference is single-threaded"},
        "e2e": {"value": rate, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the LoRADS B200 kernel layer has no CPU fallback")
    torch.cuda.set_device(local)
    from lorads_b200.capi import Solver, load_library
    import ctypes as C
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def new_comm():
        """A fresh library communicator (NCCL id created on rank 0, broadcast through torch.distributed)."""
        if world == 1:
            return None
        lib = load_library()
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            rc = lib.lb2_comm_unique_id(buf)
            assert rc == 0, lib.lb2_last_error()
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        return (C.create_string_buffer(bytes(uid.cpu().numpy().tobytes()), 128), rank, world)

    comm = new_comm()
    inst = make_instance()
    t_setup = time.time()
    S = Solver(inst, device=local, comm=comm)
    t_setup = time.time() - t_setup
    rho = S.dinfo(6)
    n, r, m = S.dim(0), S.rank(0), S.m
    shard_desc = shard_description(world, S.info(28) == 1)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident timing ----------------
    # The K-step region is short (K = 20 is 5 ms): it is repeated and the MEDIAN region is reported.  Every repeat
    # restarts from the same point (ALG_START of the random start advanced by the warm-up), so all regions do the same work.
    sampler = ClockSampler(local)
    sampler.start()        # started before the warm-up so that samples are flowing when the timed region begins
    S.alm_prepare(rho)
    S.time_alm_inner_iters(rho, max(3, args.warmup))
    est = S.time_alm_inner_iters(rho, min(args.steps, 20))[0] / min(args.steps, 20)
    repeats = args.repeats if args.repeats > 0 else int(max(1, min(25, np.ceil(0.5 / max(est * args.steps, 1e-6)))))
    if dist is not None:
        t = torch.tensor([repeats], dtype=torch.int64, device="cuda")
        dist.broadcast(t, 0)
        repeats = int(t.item())
    regions, launches, done = [], 0, args.steps
    sampler.mark()
    for _ in range(repeats):
        S.alm_prepare(rho)     # same start for every region
        barrier()
        l0 = S.launches
        sec, done = S.time_alm_inner_iters(rho, args.steps)
        launches = S.launches - l0
        barrier()
        if dist is not None:
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        regions.append(sec)
    clocks = sampler.stop()
    sec = float(np.median(regions))
    if done != args.steps:
        log(f"warning: only {done} of {args.steps} iterations ran (line search reported no root)")
    value = done / sec

    # ---------------- end to end through host buffers ----------------
    Rh = torch.empty(n * r, dtype=torch.float64).pin_memory().numpy()
    Ro = torch.empty(n * r, dtype=torch.float64).pin_memory().numpy()
    lam = torch.zeros(m, dtype=torch.float64).pin_memory().numpy()
    # sharded runs: every rank holds host buffers of the full factor and moves only the columns it owns
    Rh[:] = np.ascontiguousarray(S.get_factor("R").T).ravel()
    S.alm_run_host(Rh, lam, rho, 3, Ro)        # warm
    e2e_regions = []
    for _ in range(repeats):
        barrier()
        t0 = time.perf_counter()
        done_e, _ = S.alm_run_host(Rh, lam, rho, args.steps, Ro)
        torch.cuda.synchronize()
        e2e_sec = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_sec = float(t.item())
        e2e_regions.append(e2e_sec)
    e2e_sec = float(np.median(e2e_regions))
    e2e = {"value": done_e / e2e_sec, "unit": "iterations/s",
           "h2d_bytes_per_step": (Rh.nbytes + world * lam.nbytes) / max(done_e, 1), "d2h_bytes_per_step": Ro.nbytes / max(done_e, 1),
           "note": "one lb2_alm_run_host call per rank: R and lambda host->device (pinned, transposed on the device), K iterations, R device->host; copies amortised over the K steps of the call; bytes summed over ranks; median of the repeats"}

    # ---------------- roofline of the hot kernels (cone 0, CUDA events) ----------------
    # every rank takes part: with sharding the A(UV^T) launches are followed by their all-reduce.
    # cold: the L2 is flushed (a 384 MB buffer read through it) before EVERY launch and each launch is timed on its own --
    #       the figure the roofline uses (the 38 MB of factors would otherwise be served from the 126 MB L2);
    # hot : 200 back-to-back launches (L2 resident), for reference.
    peak, peak_src = measured_peak_gbs()
    kb = kernel_bytes(S)
    kt, kt_hot = {}, {}
    # what the cold method itself costs per launch (flush, event, launch of an idle GPU, event): a one-thread kernel timed
    # the same way; it is subtracted from every cold figure and reported
    S.bench_kernel(99, 10, True)
    cold_overhead = S.bench_kernel(99, 40, True)
    for which in kb:
        if kb[which][1] == 0:
            kt[which] = kt_hot[which] = 0.0
            continue
        kt_hot[which] = S.bench_kernel(which, 200)
        kt[which] = max(S.bench_kernel(which, 40, True) - cold_overhead, 1e-6)
    comm_us = None
    if world > 1:
        barrier()
        comm_us = {"allgather_direction": S.bench_kernel(10, 50) * 1e3, "allreduce_m_vectors": S.bench_kernel(11, 50) * 1e3,
                   "allreduce_scalar_table": S.bench_kernel(12, 50) * 1e3}
    barrier()

    # ---------------- BASELINE.json configs[4] (MaxCut n = 1e6) at the same N: iterations/s + gather kernels, informational
    cfg5 = None
    if not args.no_cfg5 and WORKLOAD is WORKLOADS["cfg2"]:
        try:
            from lorads_b200 import sdpa as _sdpa
            w5 = WORKLOADS["cfg5"]
            t5 = time.time()
            inst5 = _sdpa.maxcut(w5["n"], w5["edges"], w5["seed"])
            S5 = Solver(inst5, device=local, comm=new_comm())
            setup5 = time.time() - t5
            rho5 = S5.dinfo(6)
            S5.alm_prepare(rho5)
            S5.time_alm_inner_iters(rho5, 5)
            S5.alm_prepare(rho5)
            barrier()
            sec5, done5 = S5.time_alm_inner_iters(rho5, 40)
            barrier()
            if dist is not None:
                t = torch.tensor([sec5], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec5 = float(t.item())
            kb5 = kernel_bytes(S5)
            k5 = {}
            for wk in (0, 3):
                cold5 = max(S5.bench_kernel(wk, 10, True) - cold_overhead, 1e-6)
                k5[kb5[wk][0]] = {"ms_cold": cold5, "alg_bytes": kb5[wk][1], "gbs": kb5[wk][1] / (cold5 * 1e-3) / 1e9,
                                  "frac_of_peak": kb5[wk][1] / (cold5 * 1e-3) / 1e9 / peak}
            comm5 = None
            if world > 1:
                barrier()
                comm5 = {"allgather_direction": S5.bench_kernel(10, 10) * 1e3, "allreduce_m_vectors": S5.bench_kernel(11, 10) * 1e3,
                         "allreduce_scalar_table": S5.bench_kernel(12, 10) * 1e3}
            cfg5 = {"workload": f"MaxCut SDP n=m={w5['n']} edges={w5['edges']} seed={w5['seed']} rank={S5.rank(0)} ({w5['label']})",
                    "collectives_us": comm5,
                    "alm_inner_iterations_per_second": done5 / sec5, "ms_per_step": 1e3 * sec5 / max(done5, 1), "steps": done5,
                    "n_gpus": world, "setup_seconds_incl_generation": setup5, "kernels_this_rank": k5}
            S5.close()
            del inst5
        except Exception as e:
            log("cfg5 secondary failed:", e)
    barrier()
    if rank != 0:
        S.close()
        if dist is not None:
            dist.destroy_process_group()
        return 0
    # per-iteration multiplicity: 1 fused A(UV^T) pass, 1 adjoint SpMM, ~6 BLAS-1 passes
    # (with the three-output gather pass the separate A(RR^T) launch is not part of the pipelined step)
    mult = {0: 1, 1: 0 if S.info(20) == 1 else 1, 2: 1, 3: 1, 4: 6}
    ms_step = 1e3 * sec / done
    kernels = {}
    for w in kb:
        if kb[w][1] == 0:
            continue
        gbs = kb[w][1] / (kt[w] * 1e-3) / 1e9
        kernels[kb[w][0]] = {"ms_cold": kt[w], "ms_hot": kt_hot[w], "alg_bytes": kb[w][1], "gbs": gbs,
                             "gbs_hot": kb[w][1] / (kt_hot[w] * 1e-3) / 1e9,
                             "ncu_dram_bytes": ncu_traffic(kb[w][0]) if WORKLOAD is WORKLOADS["cfg2"] else None,
                             "frac_of_peak": gbs / peak, "frac_of_peak_hot": kb[w][1] / (kt_hot[w] * 1e-3) / 1e9 / peak,
                             "share_of_step_hot": kt_hot[w] * mult[w] / ms_step}
    dom = 0                      # the north_star kernel: the fused A(UV^T) gather pass; the adjoint SpMM sits beside it
    achieved = kb[dom][1] / (kt[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": kb[dom][0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(kb[dom][0]) if WORKLOAD is WORKLOADS["cfg2"] else None,
                "traffic_source": "profiles/r02_ncu_traffic.json (ncu --set full, cold-cache replay, per launch)",
                "peak_source": peak_src, "alg_bytes_per_launch": kb[dom][1], "launch_ms": kt[dom],
                "timing": "CUDA events around single launches, L2 flushed (384 MB read through it, clean lines) before every launch, mean of 40, minus the same measurement of a one-thread kernel (cold_event_overhead_ms)",
                "cold_event_overhead_ms": cold_overhead,
                "collectives_us": comm_us,
                "adjoint": {"kernel": kb[3][0], "achieved": kb[3][1] / (kt[3] * 1e-3) / 1e9,
                            "frac": kb[3][1] / (kt[3] * 1e-3) / 1e9 / peak, "alg_bytes_per_launch": kb[3][1], "launch_ms": kt[3],
                            "traffic": ncu_traffic(kb[3][0]) if WORKLOAD is WORKLOADS["cfg2"] else None},
                "kernels": kernels}

    # ---------------- the other BASELINE.json metrics on the same workload (N=1, informational) ----------------
    secondary = {"cfg5": cfg5} if cfg5 is not None else None
    if world == 1 and not args.no_secondary:
        try:
            secondary = secondary or {}
            # CG iterations/s of the ADMM block solve (LORADSUpdateSDPVarOne): tolerance 0 forces exactly 60 iterations
            S.admm_init_constr()
            S.update_sdp_var_one("U", "V", rho, 0.0, 5)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            its = S.update_sdp_var_one("U", "V", rho, 0.0, 60)
            torch.cuda.synchronize()
            secondary["cg_iterations_per_second"] = its / (time.perf_counter() - t0)
            # the same step with the reference's two-loop recursion carried out pass by pass (the default is its
            # Gram-table form: same mathematics, one reduction instead of five, at 1 GPU and at N alike)
            os.environ["LORADS_B200_EXACT_LBFGS"] = "1"
            try:
                Sg = Solver(inst, device=local)
                Sg.alm_prepare(rho)
                Sg.time_alm_inner_iters(rho, 10)
                Sg.alm_prepare(rho)
                torch.cuda.synchronize()
                sg, dg = Sg.time_alm_inner_iters(rho, max(args.steps, 100))
                secondary["two_loop_lbfgs_iterations_per_second"] = dg / sg
                Sg.close()
            finally:
                os.environ.pop("LORADS_B200_EXACT_LBFGS", None)
            # wall time of a whole solve to DIMACS 1e-5 (default parameters, fresh solver, data already on the host)
            if WORKLOAD is WORKLOADS["cfg2"]:
                from lorads_b200.capi import default_params
                S2 = Solver(inst, device=local)
                with c_stdout_to_stderr():
                    res = S2.solve(default_params())
                S2.close()
                secondary["whole_solve"] = {"seconds_to_dimacs_1e-5": res["solveSeconds"], "alm_inner_iterations": res["almInnerIter"],
                                            "admm_iterations": res["admmIter"], "status": res["status"],
                                            "pInfeasL1": res["pInfeasL1"], "dInfeasL1": res["dInfeasL1"], "pdGap": res["pdGap"],
                                            "final_rank": res["finalRank0"]}
        except Exception as e:
            log("secondary metrics failed:", e)

    # ---------------- BASELINE.json configs[4] (MaxCut n = 1e6) at the same N: iterations/s, informational ----------------
    # (all ranks take part; placed after rank 0 gathered everything it needs from the cfg2 solver)
    # ---------------- CPU baseline (bounded sample) ----------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            out = cpu_reference_rate(inst, args.cpu_iters)
            if out is not None:
                rate, csec, _ = out
                cpu = {"value": rate, "unit": "iterations/s", "cores": 1, "kind": "reference",
                       "sample": f"{args.cpu_iters} inner iterations of the untouched reference (oracle/_ref, -DMAC_INT64, OpenBLAS 1 thread) on the same instance and start, {csec:.1f} s; host has {os.cpu_count()} cores, the reference is single-threaded"}
        except Exception as e:  # the baseline must never take the bench line down
            log("cpu baseline failed:", e)
    if cpu is None:
        cpu = {"value": None, "unit": "iterations/s", "cores": 1, "kind": "reference", "sample": "not run"}

    line = {
        "secondary": secondary,
        "metric": "alm_inner_iterations_per_second", "value": value, "unit": "iterations/s", "n_gpus": world,
        "steps": done, "warmup": max(3, args.warmup), "repeats": repeats, "ms_per_step": 1e3 * sec / done, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(r, world, shard_desc), "setup_seconds": t_setup,
        "region_seconds": {"median": sec, "min": float(min(regions)), "max": float(max(regions))},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    S.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--repeats", type=int, default=0, help="repeats of the timed K-step region (0: enough for ~0.5 s, at most 25)")
    ap.add_argument("--cpu-iters", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the MaxCut n=1e6 secondary measurement")
    ap.add_argument("--no-secondary", action="store_true", help="skip the CG / two-loop / whole-solve secondary measurements")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    global WORKLOAD
    WORKLOAD = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
