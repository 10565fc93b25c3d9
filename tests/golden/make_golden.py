"""Generates tests/golden/*.npz.

The reference (mathLab/WaveBEM) cannot be built or run in this image and ships no fixture for the
path, so these golden vectors come from the CPU oracle (oracle/wbem_oracle.c, pinned by the KATs
in tests/test_oracle_kat.py), generated once on this container's x86-64 and committed.  They
freeze the oracle against accidental change and give the GPU tests a fixture that does not
depend on the oracle being rebuilt.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from wavebem_b200 import meshgen  # noqa: E402
from wavebem_b200.constraints import compute_constraints  # noqa: E402


def case(name, mesh, surface_nodes, bc, tol=1e-12):
    n = mesh.n_nodes
    nn = meshgen.cell_normals_at_nodes(mesh)
    cl = compute_constraints(mesh.dn_ptr, mesh.dn_idx, surface_nodes, bc, nodes_normals=nn)
    con = orc.Constraints(n, cl.lines, cl.ptr, cl.col, cl.val, cl.inhom)
    o = 1.0 - surface_nodes
    nm, dm = orc.assemble_rows(mesh.xyz, mesh.cells, mesh.dir_flag, mesh.dn_ptr, mesh.dn_idx, nthreads=1)
    alpha = orc.compute_alpha(nm, nthreads=1)
    x = np.sin(0.37 * np.arange(n))
    sol = orc.solve_system(nm, dm, surface_nodes, o, bc, con, np.zeros(n), np.zeros(n), tol=tol, max_steps=400,
                           nthreads=1)
    assert sol["converged"]
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), xyz=mesh.xyz, cells=mesh.cells, dir_flag=mesh.dir_flag, dn_ptr=mesh.dn_ptr,
        dn_idx=mesh.dn_idx, surface_nodes=surface_nodes, other_nodes=o, bc=bc, con_lines=cl.lines, con_ptr=cl.ptr,
        con_col=cl.col, con_val=cl.val, con_inhom=cl.inhom, neumann=nm, dirichlet=dm, alpha=alpha, x=x,
        vmult=orc.vmult(nm, dm, alpha, surface_nodes, o, x, nthreads=1),
        rhs=orc.compute_rhs(nm, dm, alpha, surface_nodes, o, x, nthreads=1),
        cvmult=orc.constrained_vmult(nm, dm, alpha, surface_nodes, o, con, x, nthreads=1),
        system_rhs=sol["rhs"], sol=sol["sol"], iters=sol["iters"], tol=tol)
    print(name, "N", n, "lines", cl.n_lines, "GMRES its", sol["iters"])


def projections(name):
    """<name>_projections.npz: the L2 projections run by compute_constraints and the lines it produces
    (computational_domain.cc:1525-1620, bem_problem.cc:1153-1293, :990-1105) on a committed case."""
    g = np.load(os.path.join(HERE, name + ".npz"))
    s = g["surface_nodes"]
    nrm = orc.compute_normals(g["xyz"], g["cells"], g["dir_flag"])
    grd = orc.compute_surface_gradients(g["xyz"], g["cells"], g["dir_flag"], g["bc"], s)
    cl = compute_constraints(g["dn_ptr"], g["dn_idx"], s, g["bc"], nodes_normals=nrm, node_surface_gradients=grd)
    np.savez_compressed(os.path.join(HERE, name + "_projections.npz"), case=name, nodes_normals=nrm,
                        node_surface_gradients=grd, con_lines=cl.lines, con_ptr=cl.ptr, con_col=cl.col,
                        con_val=cl.val, con_inhom=cl.inhom)
    print(name, "projections: lines", cl.n_lines)


if __name__ == "__main__":
    if "--projections-only" in sys.argv:
        projections("cube3_mixed")
        projections("tank_small")
        sys.exit(0)
    m = meshgen.cube(3, renumber="random", seed=1, flip_every=4)
    top = (m.node_patch == m.patch_names.index("z1")).astype(float)
    x0 = np.array([1.7, 1.3, 2.1])
    d = m.xyz - x0
    r = np.linalg.norm(d, axis=1)
    nn = meshgen.cell_normals_at_nodes(m)
    case("cube3_mixed", m, top, np.where(top == 1, 1 / r, -(d * nn).sum(1) / r ** 3))
    t = meshgen.wigley_tank(nxm=8, nt=4, nxu=3, nxd=4, nz=2, nzh=3, renumber="hierarchical")
    case("tank_small", t, t.surface_nodes, meshgen.towing_tank_bc(t))
    projections("cube3_mixed")
    projections("tank_small")
