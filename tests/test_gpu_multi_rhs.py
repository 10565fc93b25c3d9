"""wbem_solve_system_multi: the J.v pattern of FreeSurface::jacobian (free_surface.cc:4918-4993) -- several
solve_system calls on unchanged matrices -- as one call whose GMRES iterations share each pass over the
matrices.  Every system must agree with its own single solve_system within the solver tolerance."""
import numpy as np
import pytest

from conftest import make_problem
from wavebem_b200 import meshgen

pytestmark = pytest.mark.gpu


def _directions(bc, n, k):
    return np.stack([bc * np.cos(0.1 * (j + 1) * np.arange(n)) + 0.05 * j for j in range(k)])


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("nrhs", [1, 3, 8, 11])
def test_block_solve_matches_single_solves(wb, kind, nrhs):
    m = meshgen.wigley_tank(nxm=14, nt=6, nxu=5, nxd=7, nz=3, nzh=4)
    bc = meshgen.towing_tank_bc(m)
    n = m.n_nodes
    tol = 1e-11
    ctx = wb.Context(gmres_tol=tol, gmres_max_steps=400, precond_kind=kind, auto_constraints=1)
    ctx.set_topology(n, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    ctx.set_geometry(m.xyz)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    V = _directions(bc, n, nrhs)
    z = np.zeros(n)
    singles = [ctx.solve_system(z, z, v) for v in V]
    phi, dphi, it, res = ctx.solve_system_multi(z, z, V)
    t_multi = ctx.timings()
    for b in range(nrhs):
        p1, d1, it1, _ = singles[b]
        scale = np.linalg.norm(p1) + np.linalg.norm(d1)
        assert np.linalg.norm(phi[b] - p1) + np.linalg.norm(dphi[b] - d1) <= 50 * tol * max(1.0, scale)
        assert abs(int(it[b]) - it1) <= 1 and res[b] <= tol
    # the matrices were streamed once per block iteration, not once per system and iteration
    blocks = -(-nrhs // 8)
    assert t_multi["gemv_calls"] <= blocks * (max(it) + 3)
    # in-out semantics: only the unknown half is overwritten (:869-879)
    s = m.surface_nodes == 1
    phi0 = np.where(s, 7.0, 0.0)
    p2, d2, _, _ = ctx.solve_system_multi(phi0, z, V[:2])
    assert np.all(p2[:, s] == 7.0) and np.all(d2[:, ~s] == 0.0)
    ctx.close()


def test_block_solve_with_installed_constraints_pure_neumann_and_groups(wb, orc):
    """caller-installed lines (shared inhomogeneities), the pure-Neumann shift (bem_problem.cc:667-668)
    inside the block mat-vec, and the block solve through a single-process group of 3 row blocks."""
    m = meshgen.wigley_tank(nxm=10, nt=5, nxu=4, nxd=5, nz=3, nzh=3)
    bc, nn, cl = make_problem(m)
    n = m.n_nodes
    kw = dict(gmres_tol=1e-11, gmres_max_steps=400)
    one = wb.Context(**kw)
    grp = wb.Context(n_gpus=3, devices=[0, 0, 0], **kw)
    outs = []
    for c in (one, grp):
        c.set_topology(n, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
        c.set_geometry(m.xyz)
        c.assemble()
        c.set_masks(m.surface_nodes, m.other_nodes)
        c.set_constraints(cl)
        outs.append(c.solve_system_multi(np.zeros(n), np.zeros(n), np.stack([bc, bc, bc])))
    p1, d1, it1, _ = one.solve_system(np.zeros(n), np.zeros(n), bc)
    for phi, dphi, it, res in outs:
        for b in range(3):
            assert np.linalg.norm(phi[b] - p1) + np.linalg.norm(dphi[b] - d1) <= 1e-9 * (np.linalg.norm(p1) + np.linalg.norm(d1))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])   # group = one block, bitwise
    # pure Neumann (bem_problem.cc:667-668): the shift by the norm is applied per vector of the block
    one.set_masks(np.zeros(n), np.ones(n))
    one.set_constraints(type(cl)(n, cl.lines[:0], np.zeros(1, np.uint32), cl.col[:0], cl.val[:0], cl.inhom[:0]))
    x = np.stack([np.sin(0.3 * np.arange(n)), np.cos(0.2 * np.arange(n)), np.ones(n)])
    y1 = np.stack([one.constrained_vmult(v) for v in x])
    assert np.array_equal(one.constrained_vmult_multi(x), y1)
    one.close()
    grp.close()
