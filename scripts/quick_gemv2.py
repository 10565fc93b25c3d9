import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import wavebem_b200 as wb
from wavebem_b200 import meshgen
from wavebem_b200.constraints import ConstraintLines
from conftest import make_problem
n, world = int(sys.argv[1]), int(sys.argv[2])
m = meshgen.wigley_tank_for_nodes(n)
bc, nn, cl = make_problem(m)
for rank in range(world):
    for con in (True, False):
        ctx = wb.Context(world_size=world, rank=rank)
        ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
        ctx.set_geometry(m.xyz)
        ctx.assemble()
        ctx.set_masks(m.surface_nodes, m.other_nodes)
        ctx.set_constraints(cl if con else ConstraintLines.empty(m.n_nodes))
        ms, by = ctx.time_operator(10, True)
        r0, r1 = ctx.row0, ctx.row1
        ncon = int(((cl.lines >= r0) & (cl.lines < r1)).sum())
        print(f"rank={rank} rows={r1 - r0} constrained_in_block={ncon} use_constraints={con}: {ms * 1e3:.1f} us, "
              f"{by / 1e9:.3f} GB -> {by / ms / 1e6:.0f} GB/s; surf rows {int(m.surface_nodes[r0:r1].sum())}")
        ctx.close()
