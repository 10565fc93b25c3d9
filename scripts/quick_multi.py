"""Block solve timing (development helper): wbem_solve_system_multi vs single solves."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wavebem_b200 as wb
from wavebem_b200 import meshgen
n_t = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
m = meshgen.wigley_tank_for_nodes(n_t)
bc = meshgen.towing_tank_bc(m); n = m.n_nodes
ctx = wb.Context(gmres_tol=1e-10, gmres_max_steps=1000, auto_constraints=1, precond_kind=1)
ctx.set_topology(n, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx); ctx.set_geometry(m.xyz); ctx.assemble()
ctx.set_masks(m.surface_nodes, m.other_nodes)
z = np.zeros(n)
ctx.solve_system(z, z, bc)
t0 = time.perf_counter(); ctx.solve_system(z, z, bc); t1 = time.perf_counter()
t = ctx.timings()
print(f"single: wall {1e3*(t1-t0):.2f} ms, device {t['solve_system_total_ms']:.2f} ms, gemv {t['gemv_ms_sum']:.2f} ms / {t['gemv_calls']} calls")
for nv in (1, 2, 4, 5, 8):
    V = np.stack([bc * np.cos(0.1 * (j + 1) * np.arange(n)) for j in range(nv)])
    ctx.solve_system_multi(z, z, V)
    t0 = time.perf_counter(); _, _, it, _ = ctx.solve_system_multi(z, z, V); t1 = time.perf_counter()
    t = ctx.timings()
    print(f"nrhs {nv}: wall {1e3*(t1-t0):.2f} ms, device {t['solve_system_total_ms']:.2f} ms, gemv {t['gemv_ms_sum']:.2f} ms / {t['gemv_calls']} calls "
          f"= {t['gemv_ms_sum']/max(1,t['gemv_calls']):.3f} ms each, iters {list(it)}")
