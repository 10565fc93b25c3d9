#!/usr/bin/env python
"""Benchmark of the BEM hot path: one step = BEMProblem::solve() = assemble_system + solve_system
(dense fp64 assembly of both matrices + preconditioned GMRES) on a synthetic tank + Wigley hull.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

N = 1: BASELINE.json configs[1] (~20k collocation nodes, one B200).  N > 1: configs[2] -- the SAME
~40k-node problem at every GPU count, matrices row-sharded (strong scaling); `--ladder` restores the
weak ladder N = 20k sqrt(P) of round 1; `--config 4` / `--config 5` select the 100k-node problem and
the IDA call pattern.  Launched by torchrun it runs one rank per GPU; WITHOUT torchrun `--gpus N`
drives all N GPUs from this ONE process through a single context (wbem_params.n_gpus).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BEM matrix entries/s (fp64) through assemble_system + GMRES solve_system"
UNIT = "entries/s"
FLOP_PER_EVAL = 34.0   # algorithmic flop per (node, quadrature point)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons during the timed regions.  NVML in-process (a query
    costs ~0.1 ms and does not disturb the run; spawning nvidia-smi every few ms measurably slows
    the HBM-bound mat-vec), nvidia-smi as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []      # (sm_mhz, sm_max_mhz, power_w, set of reasons)
        self.stop_flag = False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates all GPUs of the box; CUDA_VISIBLE_DEVICES may renumber them
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = gpu_index
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
            self.source = "nvml"
        except Exception:
            self.source = "nvidia-smi"

    def sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(h) / 1e3
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        self.samples.append((float(sm), float(mx), float(pw), {k for k, bit in self.REASONS if mask & bit}))

    def sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                              "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
        f = [x.strip() for x in out.strip().split(",")]
        if len(f) >= 8:
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            self.samples.append((float(f[1]), float(f[2]), float(f[3]),
                                 {k for k, v in zip(names, f[4:8]) if v.lower().startswith("active")}))

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml:
                    self.sample_nvml()
                else:
                    self.sample_smi()
            except Exception:
                pass
            time.sleep(0.02 if self.nvml else 0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [k for k, _ in self.REASONS if any(k in s[3] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": reasons,
                "power_w_max": max(s[2] for s in self.samples), "samples": len(self.samples), "source": self.source}


def host_threads():
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU
    arm sets its own thread count instead of inheriting that."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


# BASELINE.json configs (index as in the file, 1-based in the comments of SURVEY 8d):
#   2: ~20k nodes on 1 GPU            3: ~40k nodes, the SAME problem on 1/2/4/8 GPUs (strong scaling)
#   4: ~100k nodes on 8 GPUs          5: the IDA call pattern (2 residual solve() + 5 J.v solve_system() per time step)
CONFIG_NODES = {2: 20000, 3: 40000, 4: 100000, 5: 4000}


def pick_workload(args, world):
    cfg = args.config if args.config else (2 if world == 1 else 3)
    nodes = args.nodes if args.nodes else CONFIG_NODES[cfg]
    scaling = "strong"
    if args.ladder:      # round-1 weak ladder: constant matrix entries per GPU
        nodes = int(round((args.nodes or 20000) * math.sqrt(world)))
        scaling = "weak"
    return cfg, nodes, scaling


def build_case(n_target, froude=0.28):
    from wavebem_b200 import meshgen
    m = meshgen.wigley_tank_for_nodes(n_target)
    bc = meshgen.towing_tank_bc(m, froude=froude)
    return m, bc


# ------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline: the oracle (a port: the reference cannot be built here)
# ------------------------------------------------------------------------------------------
def cpu_sample(m, slab_rows, gmres_iters, threads):
    """One bounded sample of the step on the host cores: assemble `slab_rows` rows of both
    matrices and stream them through the (3 + 2k) dense mat-vec units of solve_system
    (alpha, 2 for the rhs, 2 per GMRES iteration; reference bem_problem.cc:609, 648-650, 702-706).
    Rows are independent, so entries/s of the slab is the throughput of the full step.
    threads = 1 is the reference as it is (single-threaded, main.cc:26; cell-outer / node-inner
    loop order of bem_problem.cc:190-214); threads > 1 gives every OpenMP thread a block of rows."""
    from oracle import oracle as orc
    n = m.n_nodes
    r0 = (n - slab_rows) // 2
    t0 = time.perf_counter()
    nm, dm = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, r0, r0 + slab_rows,
                               nthreads=threads)
    t_asm = time.perf_counter() - t0
    x = np.sin(0.37 * np.arange(n))
    t0 = time.perf_counter()
    units = 3 + 2 * gmres_iters
    for u in range(units):
        orc.fullmatrix_vmult(nm if u % 2 == 0 else dm, x, nthreads=threads)
    t_mv = time.perf_counter() - t0
    return t_asm, t_mv, 2.0 * slab_rows * n


# GMRES(98) iterations of the reference algorithm (band-100 preconditioner, tol 1e-10) on the
# tank + Wigley mesh family, measured with precond_kind=0 (scripts/conv_probe.py; the counts agree
# with the oracle's where the oracle is affordable, tests/test_gpu_parity.py)
BAND_ITERS = ((4052, 55), (8798, 61), (20073, 83), (27749, 110), (39778, 151), (57001, 154), (80745, 553))


def reference_band_iters(n):
    import bisect
    xs = [a for a, _ in BAND_ITERS]
    k = bisect.bisect_left(xs, n)
    if k == 0:
        return BAND_ITERS[0][1]
    if k == len(xs):
        return BAND_ITERS[-1][1]
    (x0, y0), (x1, y1) = BAND_ITERS[k - 1], BAND_ITERS[k]
    return int(round(y0 + (y1 - y0) * (n - x0) / (x1 - x0)))


def cpu_baseline_block(m, slab, iters, threads, slab_1t):
    """All-core and faithful single-thread numbers of the CPU port on bounded slabs."""
    ta, tm, ent = cpu_sample(m, slab, iters, threads)
    t1a, t1m, ent1 = cpu_sample(m, slab_1t, iters, 1)
    n = m.n_nodes
    return {
        "value": ent / (ta + tm), "unit": UNIT, "cores": threads, "kind": "port",
        "sample": f"{slab}-row slab of the N={n} step: assembly of both matrices + {3 + 2 * iters} dense mat-vec units "
                  f"(k={iters} GMRES iterations: what the reference's band-100 preconditioner needs on this system), "
                  "oracle/wbem_oracle.c, OpenMP over rows; entries/s of the slab = of the full step",
        "assembly_entries_per_s": ent / ta, "assembly_s": ta, "matvec_s": tm,
        "value_1thread": ent1 / (t1a + t1m), "assembly_entries_per_s_1thread": ent1 / t1a,
        "sample_1thread": f"{slab_1t}-row slab, 1 thread: the reference as it runs (single-threaded, main.cc:26), "
                          "loop order of bem_problem.cc:190-214",
    }, ta + tm


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    threads = host_threads()     # NOT omp_get_max_threads(): torchrun sets OMP_NUM_THREADS=1
    world = max(1, args.gpus)
    cfg, n_target, scaling = pick_workload(args, world)
    m, bc = build_case(n_target)
    slab = args.ref_slab_rows
    iters = args.ref_gmres_iters if args.ref_gmres_iters > 0 else reference_band_iters(m.n_nodes)
    for _ in range(args.warmup):
        cpu_sample(m, max(16, slab // 8), 2, threads)
    tot_t, tot_e, asm_t = 0.0, 0.0, 0.0
    for _ in range(args.steps):
        ta, tm, ent = cpu_sample(m, slab, iters, threads)
        tot_t += ta + tm
        asm_t += ta
        tot_e += ent
    t1a, t1m, ent1 = cpu_sample(m, max(16, slab // 16), iters, 1)
    value = tot_e / tot_t
    sample = (f"{slab}-row slab of the N={m.n_nodes} step per timed step: both matrices assembled + "
              f"{3 + 2 * iters} dense mat-vec units (k={iters} GMRES its with the reference's band-100 "
              "preconditioner); oracle/wbem_oracle.c with OpenMP over rows")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, m, scaling) + "; CPU sample", "baseline_config": cfg,
                   "nodes": m.n_nodes, "slab_rows": slab},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "assembly_entries_per_s": tot_e / asm_t,
                         "value_1thread": ent1 / (t1a + t1m),
                         "sample_1thread": f"{max(16, slab // 16)}-row slab on 1 thread: the reference as it runs "
                                           "(single-threaded, main.cc:26)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(cfg, m, scaling):
    names = {2: "BASELINE configs[1]: ~20k nodes on one GPU", 3: "BASELINE configs[2]: ~40k nodes, same problem at 1/2/4/8 GPUs",
             4: "BASELINE configs[3]: ~100k nodes, row-sharded", 5: "BASELINE configs[4]: IDA call pattern (emulated)"}
    w = f"tank+Wigley hull, N={m.n_nodes} nodes, C={m.n_cells} cells, Gauss 4x4 + QGaussOneOverR(5) ({names[cfg]}"
    return w + (", weak ladder N ~ 20k sqrt(P))" if scaling == "weak" else ")")


# ------------------------------------------------------------------------------------------
def kernel_source_hash():
    import hashlib
    h = hashlib.sha1()
    for f in ("assemble.cu", "operator.cu"):
        h.update(open(os.path.join(ROOT, "wavebem_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def profile_number(fname, key):
    """DRAM traffic per launch from the committed ncu capture -- only while it describes the kernels
    as they are now (the capture script stores a hash of the kernel sources beside the number)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", fname)))
        if d.get("kernel_source_sha1_16") != kernel_source_hash():
            return None
        return d.get(key)
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist

    import wavebem_b200 as wb
    from wavebem_b200 import dist as wd

    rank, world, local = wd.env_rank_world()
    under_torchrun = "WORLD_SIZE" in os.environ and world > 1
    single_process = (not under_torchrun) and args.gpus > 1   # ONE process drives all GPUs (wbem_params.n_gpus)
    n_gpus = args.gpus if single_process else world
    if under_torchrun and world != max(1, args.gpus) and rank == 0:
        print(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}", file=sys.stderr)
    torch.cuda.set_device(local)
    if under_torchrun:
        dist.init_process_group("nccl")
    cfg, n_target, scaling = pick_workload(args, n_gpus)
    m, bc = build_case(n_target)
    n = m.n_nodes
    kind = 1 if args.precond == "spai" else 0
    common = dict(gmres_tol=args.tol, gmres_max_steps=args.max_steps, precond_kind=kind, auto_constraints=1)
    t_topo = time.perf_counter()
    if single_process:
        ctx = wb.Context(n_gpus=n_gpus, **common)
    else:
        ctx = wb.Context(device=local, rank=rank, world_size=world, **common)
    ctx.set_topology(n, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    p2p = None
    if under_torchrun:
        wd.init_comm(ctx)
        p2p = (not args.no_p2p) and wd.init_peer_gather(ctx)
    set_topology_s = time.perf_counter() - t_topo
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    # no constraint lines are handed over: solve_system runs compute_constraints itself
    # (compute_normals + compute_surface_gradients + the double-node walk, bem_problem.cc:845)

    dev = torch.device("cuda", local)
    d_xyz = torch.from_numpy(m.xyz).to(dev)
    d_bc = torch.from_numpy(bc).to(dev)
    d_phi = torch.zeros(n, dtype=torch.float64, device=dev)
    d_dphi = torch.zeros(n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if under_torchrun:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if not under_torchrun:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if cfg == 5:
        return run_ida_pattern(args, ctx, m, bc, n_gpus, rank, under_torchrun, single_process, barrier, max_over_ranks,
                               scaling)

    def step_dev():
        rc, it, res = ctx.solve_dev(d_xyz.data_ptr(), d_phi.data_ptr(), d_dphi.data_ptr(), d_bc.data_ptr())
        return rc, it, res

    # ---- device-resident leg (value) ----
    t_first = time.perf_counter()
    step_dev()          # the first step also pays the once-per-mesh preconditioner pattern
    first_step_s = time.perf_counter() - t_first
    for _ in range(max(0, args.warmup - 1)):
        step_dev()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ctx.reset_counters()
    ctx.timer_start()
    acc = dict(asm=0.0, reg=0.0, sing=0.0, geo=0.0, alpha=0.0, solve=0.0, gmres=0.0, gemv=0.0, gemv_calls=0,
               precond_setup=0.0, precond_apply=0.0, allgather=0.0, rhs=0.0, constraints=0.0)
    iters = res = rc = 0
    for _ in range(args.steps):
        rc, iters, res = step_dev()
        t = ctx.timings()
        acc["asm"] += t["assemble_total_ms"]; acc["reg"] += t["assemble_regular_ms"]
        acc["sing"] += t["assemble_singular_ms"]; acc["geo"] += t["geometry_ms"]; acc["alpha"] += t["alpha_ms"]
        acc["solve"] += t["solve_system_total_ms"]; acc["gmres"] += t["gmres_ms"]; acc["gemv"] += t["gemv_ms_sum"]
        acc["gemv_calls"] += t["gemv_calls"]; acc["precond_setup"] += t["precond_setup_ms"]
        acc["precond_apply"] += t["precond_apply_ms_sum"]; acc["allgather"] += t["allgather_ms_sum"]
        acc["rhs"] += t["rhs_ms"]; acc["constraints"] += t["constraints_ms"]
    ms_total = ctx.timer_stop()
    barrier()
    launches = ctx.timings()["kernel_launches"]
    gemv_bytes = ctx.timings()["gemv_bytes_last"]
    ms_total = max_over_ranks(ms_total)
    K = args.steps
    entries = 2.0 * n * n
    value = entries * K / (ms_total * 1e-3)
    sol_spai = ctx.get_sol()

    # ---- end-to-end leg through the host-buffer C ABI (wbem_solve) ----
    h_xyz = torch.from_numpy(m.xyz).pin_memory().numpy()
    h_bc = torch.from_numpy(bc).pin_memory().numpy()
    h_phi = torch.zeros(n, dtype=torch.float64).pin_memory().numpy()
    h_dphi = torch.zeros(n, dtype=torch.float64).pin_memory().numpy()
    import ctypes as C
    it_c, res_c = C.c_int(0), C.c_double(0)

    def step_host():
        return wb.lib().wbem_solve(ctx._h, h_xyz.ctypes.data_as(C.c_void_p), h_phi.ctypes.data_as(C.c_void_p),
                                   h_dphi.ctypes.data_as(C.c_void_p), h_bc.ctypes.data_as(C.c_void_p),
                                   C.byref(it_c), C.byref(res_c))
    for _ in range(max(1, args.warmup // 2)):
        step_host()
    barrier()
    ctx.timer_start()
    for _ in range(K):
        step_host()
    e2e_ms = max_over_ranks(ctx.timer_stop())
    barrier()
    sampler.stop_flag = True   # clocks were sampled over both timed regions (device-resident and e2e)
    e2e_value = entries * K / (e2e_ms * 1e-3)
    # every process (or, in one process, every row block) uploads its own replica of the inputs
    h2d = 8 * n * (3 + 3) * n_gpus   # support points, phi, dphi_dn, tmp_rhs
    d2h = 8 * n * 2 * (world if under_torchrun else 1)   # phi, dphi_dn

    # ---- the reference's own preconditioner (band-100 LU, bem_problem.cc:1107-1149) on the same
    # assembled system, outside the timed regions: its iteration count is what the CPU arm runs, and
    # assembly + this solve is the like-for-like GPU step ----
    band_iters, band_solve_ms, band_rc, sol_band = iters, acc["solve"] / K, rc, sol_spai
    if kind == 1:
        ctx.set_precond_kind(0)
        for _ in range(2):
            band_rc, band_iters, _ = ctx.solve_system_dev(d_phi.data_ptr(), d_dphi.data_ptr(), d_bc.data_ptr())
        band_solve_ms = max_over_ranks(ctx.timings()["solve_system_total_ms"])
        sol_band = ctx.get_sol()
        ctx.set_precond_kind(1)

    # ---- parity, outside the timed regions, at THIS size and sharding: a slab of this rank's rows
    # against the CPU oracle (1e-11 of the row scale, DESIGN "Parity metric") and the
    # SPAI-preconditioned solution against the band-preconditioned one ----
    parity = None
    if not args.no_parity:
        r_lo, r_hi = (0, n) if single_process else (ctx.row0, ctx.row1)
        rows = min(args.parity_rows, r_hi - r_lo)
        # in one process take rows that straddle the first block boundary; else the middle of this rank's block
        mid = (-(-n // n_gpus)) if (single_process and n_gpus > 1) else (r_lo + r_hi) // 2
        a = max(r_lo, min(r_hi - rows, mid - rows // 2))
        gn, gd = ctx.get_rows(0, a, a + rows), ctx.get_rows(1, a, a + rows)
        galpha = ctx.get_alpha()
        if rank == 0:
            from oracle import oracle as orc
            orc.build()
            on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, a, a + rows,
                                       nthreads=host_threads())
            oalpha = orc.compute_alpha(on)
            sc_n = np.maximum(np.maximum(np.abs(on).max(axis=1, keepdims=True), np.abs(oalpha)[:, None]),
                              np.maximum(np.abs(on), np.abs(gn)))
            sc_d = np.maximum(np.abs(od).max(axis=1, keepdims=True), np.maximum(np.abs(od), np.abs(gd)))
            parity = {"rows_checked": int(rows), "first_row": int(a), "row_block": [int(r_lo), int(r_hi)],
                      "max_err_N": float((np.abs(gn - on) / sc_n).max()), "max_err_D": float((np.abs(gd - od) / sc_d).max()),
                      "max_err_alpha": float(np.abs(galpha[a:a + rows] - oalpha).max()),
                      "sol_vs_band_relerr": float(np.linalg.norm(sol_spai - sol_band) / np.linalg.norm(sol_band)),
                      "checker": "oracle/wbem_oracle.c on the same rows (entries: |a-b| / max(|a|,|b|,row scale)); "
                                 "solution: SPAI- vs band-preconditioned GMRES on the same system",
                      "tolerance": {"entries": 1e-11, "solution": 50 * args.tol}}
        barrier()

    # ---- strong-scaling companion (config 3 at P > 1): the same workload on ONE GPU, in this run ----
    n1_same = None
    if cfg == 3 and n_gpus > 1 and not args.ladder and not args.no_n1_companion:
        if rank == 0:
            c1 = wb.Context(device=local, **common)
            c1.set_topology(n, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
            c1.set_masks(m.surface_nodes, m.other_nodes)
            for _ in range(2):
                c1.solve_dev(d_xyz.data_ptr(), d_phi.data_ptr(), d_dphi.data_ptr(), d_bc.data_ptr())
            c1.timer_start()
            for _ in range(3):
                c1.solve_dev(d_xyz.data_ptr(), d_phi.data_ptr(), d_dphi.data_ptr(), d_bc.data_ptr())
            n1_same = {"n_gpus": 1, "nodes": n, "ms_per_step": c1.timer_stop() / 3,
                       "note": "same workload on one GPU of this box, measured in this run (rank 0, the others idle)"}
            c1.close()
        barrier()

    # ---- rooflines ----
    hbm_peak, peak_src = measured_peaks()
    gemv_ms_avg = acc["gemv"] / max(1, acc["gemv_calls"])
    nloc = (ctx.row1 - ctx.row0) if not single_process else -(-n // n_gpus)
    gemv_alg_bytes = gemv_bytes + 16.0 * n          # matrix chunks + x in + y out
    gemv_gbs = gemv_alg_bytes / (gemv_ms_avg * 1e-3) / 1e9 if gemv_ms_avg > 0 else 0.0
    fp64_peak = ctx.measure_fp64_peak()
    copy_bw = ctx.measure_copy_bw()
    evals = 16.0 * nloc * m.n_cells     # singular pairs (50-point rule) are < 0.1 % of F_A
    asm_tflops = FLOP_PER_EVAL * evals / (acc["reg"] / K * 1e-3) / 1e12 if acc["reg"] > 0 else 0.0
    traffic = profile_number("gemv_dram_bytes_per_launch.json", "bytes_per_launch") if n_gpus == 1 and cfg == 2 else None
    asm_traffic = profile_number("assemble_dram_bytes_per_step.json", "bytes_per_step") if n_gpus == 1 and cfg == 2 else None

    if rank == 0:
        asm_ms = acc["asm"] / K
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, m, scaling) + f", GMRES tol {args.tol:g}/max {args.max_steps}, "
                                   + ("local-inverse sparse approximate inverse preconditioner (spai.cu)" if kind == 1
                                      else "band-100 preconditioner"),
                       "baseline_config": cfg, "nodes": n, "cells": m.n_cells, "rows_per_gpu": nloc,
                       "parallelism": f"rows/{n_gpus}",
                       "process_model": ("one process, one thread calling; the library drives all GPUs (wbem_params.n_gpus)"
                                         if single_process else ("one process per GPU (torchrun)" if world > 1 else "one process")),
                       "gather": ("fused peer-to-peer stores in k_bem_gemv (peer access, one process)" if single_process else
                                  ("fused peer-to-peer stores in k_bem_gemv (CUDA IPC over NVLink)" if p2p else
                                   ("ncclAllGather" if world > 1 else "none"))),
                       "l2_policy": "inputs larger than L2 (both matrices, 16 N^2 / P bytes per GPU >> 126 MB)"},
            "gmres_iters": iters, "gmres_last_residual": res, "gmres_converged": rc == 0,
            "reference_preconditioner": {"kind": "band-100 block-cyclic-reduction LU (bem_problem.cc:1107-1149), same "
                                                 "assembled system, measured outside the timed region",
                                         "gmres_iters": band_iters, "gmres_solve_ms": band_solve_ms,
                                         "gmres_converged": band_rc == 0,
                                         "like_for_like_step_ms": asm_ms + band_solve_ms,
                                         "like_for_like_entries_per_s": entries / ((asm_ms + band_solve_ms) * 1e-3)},
            "parity": parity,
            "assembly_entries_per_s": entries / (asm_ms * 1e-3),
            "assemble_ms": asm_ms, "assemble_regular_ms": acc["reg"] / K,
            "assemble_singular_ms": acc["sing"] / K, "geometry_ms": acc["geo"] / K, "alpha_ms": acc["alpha"] / K,
            "gmres_solve_ms": acc["solve"] / K, "rhs_ms": acc["rhs"] / K,
            "compute_constraints_ms": acc["constraints"] / K, "precond_setup_ms": acc["precond_setup"] / K,
            "precond_apply_ms_per_call": acc["precond_apply"] / max(1, acc["gemv_calls"]),
            "gemv_ms_per_call": gemv_ms_avg, "gemv_calls_per_step": acc["gemv_calls"] / K,
            "non_gemv_us_per_gmres_iter": 1e3 * (acc["gmres"] - acc["gemv"]) / max(1, K * iters),
            "allgather_ms_per_step": acc["allgather"] / K,
            "once_per_mesh": {"set_topology_s": set_topology_s, "first_step_s": first_step_s,
                              "note": "reinit(): tiling plan, uploads, communicator; the first step adds the "
                                      "preconditioner's sparsity pattern and the mass-matrix structure"},
            "same_workload_one_gpu": n1_same,
            "roofline_gemv": {"bound": "hbm", "kernel": "k_bem_gemv", "achieved": gemv_gbs, "peak": hbm_peak,
                              "unit": "GB/s", "frac": gemv_gbs / hbm_peak, "traffic": traffic,
                              "peak_source": peak_src, "bytes_per_launch": gemv_alg_bytes,
                              "share_of_step": acc["gemv"] / (ms_total)},
            # the path's second bound is the FP64 pipe, not the tensor cores: "fp64" says so; peak is
            # the DFMA rate measured in this run (MEASURED_PEAKS.json has no FP64 entry)
            "roofline_assembly": {"bound": "fp64", "kernel": "k_assemble_rows (one persistent launch per assembly)",
                                  "achieved": asm_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                                  "frac": asm_tflops / fp64_peak if fp64_peak else None, "traffic": asm_traffic,
                                  "flop_per_eval": FLOP_PER_EVAL, "evals_per_step": evals,
                                  "peak_source": "DFMA loops measured in this run (wbem_measure_fp64_peak: best of "
                                                 "two kernels; nominal 148 SM x 64 FMA/clk x 1.965 GHz = 37.2)",
                                  "share_of_step": acc["reg"] / ms_total},
            "copy_bw_gbs_this_run": copy_bw,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
        }
        # "roofline" = the kernel with the larger share of this step
        dom = "roofline_gemv" if line["roofline_gemv"]["share_of_step"] >= line["roofline_assembly"]["share_of_step"] \
            else "roofline_assembly"
        line["roofline"] = dict(line[dom])
        if n_gpus == 1 and not args.no_cpu_baseline:
            from oracle import oracle as orc
            orc.build()
            line["cpu_baseline"], _ = cpu_baseline_block(m, args.cpu_slab_rows, band_iters, host_threads(),
                                                         max(16, args.cpu_slab_rows // 16))
        print(json.dumps(line), flush=True)
    ctx.close()
    if under_torchrun:
        dist.barrier()
        dist.destroy_process_group()


def run_ida_pattern(args, ctx, base, bc0, n_gpus, rank, under_torchrun, single_process, barrier, max_over_ranks, scaling):
    """BASELINE configs[4]: what a Wigley Froude-0.3 IDA/BDF run makes BEMProblem do -- emulated (SUNDIALS
    is not in this image, the integrator itself is NOT run).  One step = one TIME step = `--residuals`
    residual evaluations with the mesh moved (solve() = geometry upload + assemble_system + solve_system,
    free_surface.cc:5306-5307, 6099-6104) + `--jv` Jacobian-vector products (solve_system() on the
    unchanged matrices with new boundary data, free_surface.cc:4993)."""
    from wavebem_b200 import meshgen
    n = base.n_nodes
    L = meshgen.WIGLEY_L
    kw = {k: base.meta[k] for k in ("nxm", "nt", "nxu", "nxd", "nz", "nzh")}
    froude, g = 0.3, 9.81
    k_wave = g / (froude ** 2 * g * L)
    omega = np.sqrt(g * k_wave)
    bc = meshgen.towing_tank_bc(base, froude=froude)
    z = np.zeros(n)
    dt = 0.02
    # geometries and boundary data of all steps are prepared first: the timed region holds BEM calls only
    geos = [[meshgen.wigley_tank(**kw, renumber="hierarchical", wave_amp=0.01 * L, wave_k=k_wave,
                                 wave_phase=omega * (s * dt + 0.3 * dt * r)).xyz for r in range(args.residuals)]
            for s in range(args.warmup + args.steps)]
    dirs = [bc * np.cos(0.1 * (j + 1) * np.arange(n)) for j in range(args.jv)]
    its = []

    def time_step(s):
        for r in range(args.residuals):
            _, _, it, _ = ctx.solve(geos[s][r], z, z, bc)
            its.append(it)
        if args.jv_batched:
            ctx.solve_system_multi(z, z, np.stack(dirs))
        else:
            for v in dirs:
                ctx.solve_system(z, z, v)
    for s in range(args.warmup):
        time_step(s)
    sampler = ClockSampler(0)
    barrier()
    sampler.start()
    ctx.reset_counters()
    t0 = time.perf_counter()
    ctx.timer_start()
    for s in range(args.warmup, args.warmup + args.steps):
        time_step(s)
    ms = max_over_ranks(ctx.timer_stop())
    wall = time.perf_counter() - t0
    barrier()
    sampler.stop_flag = True
    K = args.steps
    entries = args.residuals * 2.0 * n * n      # matrix entries assembled per time step
    if rank == 0:
        line = {
            "metric": METRIC, "value": entries * K / (ms * 1e-3), "unit": UNIT, "n_gpus": n_gpus, "steps": K,
            "warmup": args.warmup, "ms_per_step": ms / K, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(5, base, scaling) + f": per time step {args.residuals} x solve() on the moved mesh "
                                   f"+ {args.jv} x solve_system() ({'one batched multi-RHS call' if args.jv_batched else 'one call each'}); "
                                   "IDA itself is not run", "baseline_config": 5, "nodes": n, "cells": base.n_cells,
                       "parallelism": f"rows/{n_gpus}", "l2_policy": "inputs larger than L2 at N >= 4k (16 N^2 bytes)"},
            "gmres_iters_mean": float(np.mean(its)), "wall_ms_per_step": 1e3 * wall / K,
            "e2e": {"value": entries * K / wall, "unit": UNIT, "ms_per_step": 1e3 * wall / K,
                    "h2d_bytes_per_step": 8 * n * (args.residuals * 6 + args.jv * 3) * n_gpus,
                    "d2h_bytes_per_step": 8 * n * 2 * (args.residuals + args.jv)},
            "gpu_launches": int(ctx.timings()["kernel_launches"]), "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if under_torchrun:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4, 5],
                    help="BASELINE.json configs (1-based): 2 = ~20k nodes (default at 1 GPU), 3 = ~40k nodes, the same "
                         "problem at every GPU count (default at > 1 GPU: strong scaling), 4 = ~100k nodes, 5 = IDA call pattern")
    ap.add_argument("--nodes", type=int, default=0, help="override the node count of the config")
    ap.add_argument("--ladder", action="store_true", help="weak ladder of round 1: N = nodes * sqrt(P), constant entries per GPU")
    ap.add_argument("--tol", type=float, default=1e-10)       # prm-files/default-2.prm:107-113
    ap.add_argument("--max-steps", type=int, default=1000)
    ap.add_argument("--cpu-slab-rows", type=int, default=2048)
    ap.add_argument("--ref-slab-rows", type=int, default=1024)
    ap.add_argument("--ref-gmres-iters", type=int, default=0,
                    help="GMRES iterations of the CPU arm; 0 = what the reference's band-100 preconditioner needs at "
                         "this size (table measured with precond_kind=0 on the GPU)")
    ap.add_argument("--precond", default="spai", choices=["spai", "band"],
                    help="spai = local-inverse sparse approximate inverse (default), band = the reference's band LU")
    ap.add_argument("--parity-rows", type=int, default=128)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-n1-companion", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-p2p", action="store_true", help="use ncclAllGather instead of the fused peer-to-peer gather")
    ap.add_argument("--residuals", type=int, default=2, help="config 5: residual evaluations per time step")
    ap.add_argument("--jv", type=int, default=5, help="config 5: Jacobian-vector products per time step")
    ap.add_argument("--jv-batched", action="store_true", help="config 5: the J.v solves as one multi-RHS call")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
