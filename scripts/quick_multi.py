"""Block solve timing (development helper): wbem_solve_system_multi vs single solves, IDA-like sequence."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wavebem_b200 as wb
from wavebem_b200 import meshgen
n_t = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
base = meshgen.wigley_tank_for_nodes(n_t)
kw = {k: base.meta[k] for k in ("nxm", "nt", "nxu", "nxd", "nz", "nzh")}
L = meshgen.WIGLEY_L
geos = [meshgen.wigley_tank(**kw, renumber="hierarchical", wave_amp=0.01 * L, wave_k=4.0, wave_phase=0.1 * s).xyz for s in range(3)]
m = base
bc = meshgen.towing_tank_bc(m, froude=0.3); n = m.n_nodes
ctx = wb.Context(gmres_tol=1e-10, gmres_max_steps=1000, auto_constraints=1, precond_kind=1)
ctx.set_topology(n, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
ctx.set_masks(m.surface_nodes, m.other_nodes)
z = np.zeros(n)
V = np.stack([bc * np.cos(0.1 * (j + 1) * np.arange(n)) for j in range(5)])
for s in range(3):
    t0 = time.perf_counter(); ctx.solve(geos[s], z, z, bc); t1 = time.perf_counter()
    ts = ctx.timings()
    t2 = time.perf_counter(); _, _, it, _ = ctx.solve_system_multi(z, z, V); t3 = time.perf_counter()
    t = ctx.timings()
    print(f"step {s}: solve wall {1e3*(t1-t0):.1f} ms (constraints {ts['constraints_ms']:.2f}); multi(5) wall {1e3*(t3-t2):.1f} ms device {t['solve_system_total_ms']:.1f} "
          f"gemv {t['gemv_ms_sum']:.1f}/{t['gemv_calls']} iters {list(it)}")
    t2 = time.perf_counter()
    for v in V: ctx.solve_system(z, z, v)
    t3 = time.perf_counter()
    print(f"         5 single solves wall {1e3*(t3-t2):.1f} ms, last: constraints {ctx.timings()['constraints_ms']:.2f} ms, total {ctx.timings()['solve_system_total_ms']:.2f}")
