"""Locally refined meshes with hanging nodes (SURVEY 8d "hanging-node variant"): the reference condenses
DoFTools::make_hanging_node_constraints into the two L2 projections of compute_constraints
(source/computational_domain.cc:1535-1538, source/bem_problem.cc:1167-1170) and into the scalar
ConstraintMatrix of the solve (source/bem_problem.cc:1000).  Checked against oracle/projections.py (normals,
surface gradients), the host walk (constraint lines) and the oracle's solve_system."""
import numpy as np
import pytest

from conftest import rel_err_rowscaled
from oracle import projections
from wavebem_b200 import meshgen
from wavebem_b200.constraints import compute_constraints

pytestmark = pytest.mark.gpu


def _refined_tank(wave=0.0):
    m = meshgen.wigley_tank(nxm=12, nt=6, nxu=4, nxd=6, nz=3, nzh=3, wave_amp=wave, wave_phase=0.3)
    cx = m.xyz[m.cells.astype(int)].mean(axis=1)
    hull = np.isin(m.cell_patch, m.meta["hull_patches"])
    fs = np.isin(m.cell_patch, [m.patch_names.index(k) for k in ("fs_mid_right", "fs_mid_left")])
    mask = (hull & (cx[:, 0] > 0.2)) | (fs & (np.abs(cx[:, 0]) < 0.6))     # part of the hull and of the free surface
    return meshgen.refine_cells(m, mask)


def _lines(cl):
    return {int(l): (sorted(zip(cl.col[cl.ptr[k]:cl.ptr[k + 1]].tolist(), cl.val[cl.ptr[k]:cl.ptr[k + 1]].tolist())),
                     float(cl.inhom[k])) for k, l in enumerate(cl.lines)}


@pytest.mark.parametrize("wave", [0.0, 0.02])
def test_projections_lines_and_solve_on_a_locally_refined_mesh(wb, orc, wave):
    r, hang = _refined_tank(wave)
    n = r.n_nodes
    assert len(hang) >= 8
    bc = meshgen.towing_tank_bc(r)
    ctx = wb.Context(gmres_tol=1e-11, gmres_max_steps=600)
    ctx.set_topology(n, r.cells, r.dir_flag, r.dn_ptr, r.dn_idx)
    ctx.set_geometry(r.xyz)
    ctx.set_masks(r.surface_nodes, r.other_nodes)
    ctx.set_hanging_constraints(hang)
    # the two projections with the hanging lines condensed into the mass systems
    nrm = ctx.compute_normals()
    ref_n = projections.l2_projection(0, r.xyz, r.cells, r.dir_flag, hanging=hang)
    assert np.abs(nrm - ref_n).max() < 1e-10
    f = np.cos(1.3 * r.xyz[:, 0]) + r.xyz[:, 1] * r.xyz[:, 2]
    grd = ctx.compute_surface_gradients(f)
    ref_g = projections.l2_projection(1, r.xyz, r.cells, r.dir_flag, f * r.surface_nodes, hanging=hang)
    assert np.abs(grd - ref_g).max() <= 1e-10 * max(1.0, np.abs(ref_g).max())
    # they differ from the unconstrained projections (the condensation is not a no-op) ...
    assert np.abs(ref_g - projections.l2_projection(1, r.xyz, r.cells, r.dir_flag, f * r.surface_nodes)).max() > 1e-7
    # ... and the hanging values are the mean of their masters (distribute())
    h0, ent = hang[0]
    assert np.abs(grd[h0] - sum(w * grd[m_] for m_, w in ent)).max() < 1e-12
    # compute_constraints: double-node lines merged with the hanging lines, closed
    got = ctx.compute_constraints(bc)
    ref_g_bc = projections.l2_projection(1, r.xyz, r.cells, r.dir_flag, bc * r.surface_nodes, hanging=hang)
    ref = compute_constraints(r.dn_ptr, r.dn_idx, r.surface_nodes, bc, nodes_normals=ref_n,
                              node_surface_gradients=ref_g_bc, hanging=hang)
    a, b = _lines(got), _lines(ref)
    assert a.keys() == b.keys() and all(h in a for h, _ in hang)
    for k in a:
        assert a[k][0] == b[k][0], k
        assert abs(a[k][1] - b[k][1]) <= 1e-9 * max(1.0, abs(b[k][1])), (k, a[k], b[k])
    # assembly on the non-conforming mesh and the constrained solve against the oracle
    ctx.assemble()
    on, od = orc.assemble_rows(r.xyz, r.cells, r.dir_flag, r.dn_ptr, r.dn_idx)
    oalpha = orc.compute_alpha(on)
    assert rel_err_rowscaled(ctx.get_rows(0), on, diag=oalpha) < 1e-11
    assert rel_err_rowscaled(ctx.get_rows(1), od) < 1e-11
    z = np.zeros(n)
    phi, dphi, it, res = ctx.solve_system(z, z, bc)
    con = orc.Constraints(n, got.lines, got.ptr, got.col, got.val, got.inhom)
    o = orc.solve_system(on, od, r.surface_nodes, r.other_nodes, bc, con, z, z, tol=1e-11, max_steps=600)
    assert np.linalg.norm(ctx.get_sol() - o["sol"]) <= 1e-8 * max(1.0, np.linalg.norm(o["sol"]))
    sol = ctx.get_sol()
    assert abs(sol[h0] - sum(w * sol[m_] for m_, w in ent)) < 1e-9          # the hanging constraint holds in the solution
    # the library's own compute_constraints inside solve_system (auto_constraints) gives the same solve
    auto = wb.Context(gmres_tol=1e-11, gmres_max_steps=600, auto_constraints=1, precond_kind=1)
    auto.set_topology(n, r.cells, r.dir_flag, r.dn_ptr, r.dn_idx)
    auto.set_masks(r.surface_nodes, r.other_nodes)
    auto.set_hanging_constraints(hang)
    p2, d2, _, _ = auto.solve(r.xyz, z, z, bc)
    assert np.linalg.norm(auto.get_sol() - o["sol"]) <= 1e-8 * max(1.0, np.linalg.norm(o["sol"]))
    ctx.close()
    auto.close()
