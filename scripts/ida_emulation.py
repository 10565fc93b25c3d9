"""BASELINE config 5: the call pattern of a Wigley Froude-0.3 IDA/BDF run, emulated.

SUNDIALS IDA is not available here, so the time integrator itself is NOT run; what is
reproduced is what it makes BEMProblem do:
  per time step:  r residual evaluations with "Keep BEM in sync with geometry" = true
                  -> geometry moves (comp_dom.update_mapping, free_surface.cc:5306-5307) and
                     bem.solve() = assemble_system + solve_system      (free_surface.cc:6099-6104)
                  m Jacobian-vector products -> bem.solve_system() on unchanged matrices with
                     new boundary data                                  (free_surface.cc:4993)
    python scripts/ida_emulation.py [--nodes 4000] [--steps 50] [--residuals 2] [--jv 5]
    torchrun --nproc-per-node 8 scripts/ida_emulation.py --nodes 40000
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wavebem_b200 as wb  # noqa: E402
from wavebem_b200 import dist as wd  # noqa: E402
from wavebem_b200 import meshgen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=4000)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--residuals", type=int, default=2)
    ap.add_argument("--jv", type=int, default=5)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--precond", default="spai", choices=["spai", "band"])
    args = ap.parse_args()
    rank, world, local = wd.env_rank_world()
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl")
    L = meshgen.WIGLEY_L
    base = meshgen.wigley_tank_for_nodes(args.nodes)
    kw = {k: base.meta[k] for k in ("nxm", "nt", "nxu", "nxd", "nz", "nzh")}
    n = base.n_nodes
    ctx = wb.Context(device=local, rank=rank, world_size=world, gmres_tol=args.tol, gmres_max_steps=1000,
                     precond_kind=1 if args.precond == "spai" else 0, auto_constraints=1)
    ctx.set_topology(n, base.cells, base.dir_flag, base.dn_ptr, base.dn_idx)
    wd.init_comm(ctx)
    wd.init_peer_gather(ctx)
    ctx.set_masks(base.surface_nodes, base.other_nodes)
    froude, g = 0.3, 9.81
    k_wave = g / (froude ** 2 * g * L)            # deep-water wave number of the ship wave
    omega = np.sqrt(g * k_wave)
    bc0 = meshgen.towing_tank_bc(base, froude=froude)
    z = np.zeros(n)
    t_solve, t_jv, its = [], [], []
    dt = 0.02

    def sync():
        # ranks build their meshes at different speeds: line them up before a timed call
        if world > 1:
            dist.barrier()

    for step in range(args.steps):
        t = step * dt
        for r in range(args.residuals):
            m = meshgen.wigley_tank(**kw, renumber="hierarchical", wave_amp=0.01 * L, wave_k=k_wave,
                                    wave_phase=omega * (t + 0.3 * dt * r))
            bc = bc0      # compute_constraints runs inside solve_system (auto_constraints)
            sync()
            t0 = time.perf_counter()
            phi, dphi, it, res = ctx.solve(m.xyz, z, z, bc)
            t_solve.append(time.perf_counter() - t0)
            its.append(it)
        for j in range(args.jv):
            v = bc0 * np.cos(0.1 * (j + 1) * np.arange(n))      # a Krylov direction's boundary data
            sync()
            t0 = time.perf_counter()
            ctx.solve_system(z, z, v)
            t_jv.append(time.perf_counter() - t0)
    if rank == 0:
        # the first time step pays the once-per-mesh host work (tiling plan, sparsity pattern)
        first = (sum(t_solve[:args.residuals]) + sum(t_jv[:args.jv]))
        if args.steps > 1:
            t_solve, t_jv, its = t_solve[args.residuals:], t_jv[args.jv:], its[args.residuals:]
        per_step = (sum(t_solve) + sum(t_jv)) / max(1, args.steps - 1 if args.steps > 1 else 1)
        print(json.dumps({"config": "IDA call pattern (emulated; IDA itself not run)", "nodes": n, "n_gpus": world,
                          "preconditioner": args.precond, "steps": args.steps, "residuals_per_step": args.residuals, "jv_per_step": args.jv,
                          "solve_ms_mean": 1e3 * float(np.mean(t_solve)), "solve_system_ms_mean": 1e3 * float(np.mean(t_jv)),
                          "bem_ms_per_time_step": 1e3 * per_step, "first_time_step_ms": 1e3 * first, "gmres_iters_mean": float(np.mean(its))}))
    ctx.close()


if __name__ == "__main__":
    main()
