"""GMRES convergence probe on one GPU (development helper): iterations / solve time of the
tank + Wigley case for several sizes, Krylov basis sizes and preconditioner bands.

    python scripts/conv_probe.py 20000,40000,80000 100,300 100,128
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import wavebem_b200 as wb
from bench import build_case

sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20000").split(",")]
bases = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "100").split(",")]
bands = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "100").split(",")]
tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-10
kinds = [int(x) for x in (sys.argv[5] if len(sys.argv) > 5 else "0").split(",")]
for n_target in sizes:
    t0 = time.time()
    m, bc, cl = build_case(n_target)
    print(f"N={m.n_nodes} C={m.n_cells} host setup {time.time() - t0:.1f}s", flush=True)
    for basis, band, kind in [(a, b, c) for a in bases for b in bands for c in kinds]:
        if True:
            ctx = wb.Context(gmres_tol=tol, gmres_max_steps=1500, gmres_n_tmp_vectors=basis,
                             preconditioner_band=band, precond_kind=kind)
            ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
            ctx.set_masks(m.surface_nodes, m.other_nodes)
            ctx.set_constraints(cl)
            ctx.set_geometry(m.xyz)
            ctx.assemble()
            phi = np.zeros(m.n_nodes)
            dphi = np.zeros(m.n_nodes)
            _, _, it, res = ctx.solve_system(phi, dphi, bc, raise_on_no_convergence=False)
            rc = int(res > tol)
            t = ctx.timings()
            t_pattern = t["precond_setup_ms"]
            if kind == 1:   # second solve: the sparsity pattern (host, once per mesh) is cached
                ctx.set_masks(m.surface_nodes, m.other_nodes)
                ctx.solve_system(phi, dphi, bc, raise_on_no_convergence=False)
                t = ctx.timings()
                print(f"    (first solve incl. host pattern build: precond setup {t_pattern:.1f} ms)")
            print(f"  kind {kind} basis {basis:4d} band {band:4d}: rc {rc} iters {it:5d} res {res:.3e} "
                  f"gmres {t['gmres_ms']:.1f} ms  precond setup {t['precond_setup_ms']:.2f} ms "
                  f"gemv/call {t['gemv_ms_sum'] / max(1, t['gemv_calls']):.3f} ms", flush=True)
            ctx.close()
            del ctx
