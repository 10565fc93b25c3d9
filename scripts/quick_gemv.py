"""Time one rank's mat-vec of a row-sharded run on a single GPU (development helper).
    WBEM_DIAG_NO_COMM=1 python scripts/quick_gemv.py <nodes> <world>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import wavebem_b200 as wb
from wavebem_b200 import meshgen
from conftest import make_problem
n, world = int(sys.argv[1]), int(sys.argv[2])
m = meshgen.wigley_tank_for_nodes(n)
bc, nn, cl = make_problem(m)
for rank in ([0, world // 2, world - 1] if world > 1 else [0]):
    ctx = wb.Context(world_size=world, rank=rank)
    ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    ctx.set_geometry(m.xyz)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(cl)
    ms, by = ctx.time_operator(10, True)
    print(f"N={m.n_nodes} world={world} rank={rank} rows={ctx.row1 - ctx.row0}: operator {ms * 1e3:.1f} us, "
          f"{by / 1e9:.3f} GB -> {by / ms / 1e6:.0f} GB/s")
    ctx.close()
