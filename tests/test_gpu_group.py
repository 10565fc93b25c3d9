"""Single-process multi-GPU context (wbem_params.n_gpus = P): ONE handle owns P row blocks.

The row blocks may share a device (devices=[0]*P), so these tests exercise every piece of the
sharded code -- row0 != 0, a ragged / empty last block, the per-block singular-pair maps, the
gathers of alpha / band rows / near-field rows, the row-sharded mat-vec and the replicated GMRES --
on the ONE GPU the driver's test box has, against the oracle and bitwise against a P = 1 context.
On a box with >= 2 GPUs the same checks run across devices with the fused peer-store mat-vec.
"""
import numpy as np
import pytest

from conftest import make_problem, rel_err_rowscaled
from wavebem_b200 import meshgen

pytestmark = pytest.mark.gpu


def _setup(ctx, m, cl=None):
    ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    ctx.set_geometry(m.xyz)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    if cl is not None:
        ctx.set_constraints(cl)
    return ctx


def _orc_con(orc, cl):
    return orc.Constraints(cl.n, cl.lines, cl.ptr, cl.col, cl.val, cl.inhom)


@pytest.fixture(scope="module")
def case(wb, orc):
    m = meshgen.wigley_tank(nxm=14, nt=6, nxu=5, nxd=7, nz=3, nzh=4)
    bc, nn, cl = make_problem(m)
    one = _setup(wb.Context(gmres_tol=1e-12, gmres_max_steps=400), m, cl)
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    z = np.zeros(m.n_nodes)
    ref = orc.solve_system(on, od, m.surface_nodes, m.other_nodes, bc, _orc_con(orc, cl), z, z, tol=1e-12, max_steps=400)
    yield dict(m=m, bc=bc, cl=cl, one=one, on=on, od=od, ref=ref)
    one.close()


@pytest.mark.parametrize("P,fused", [(2, 0), (3, 0), (8, 0), (2, 1)])
@pytest.mark.parametrize("kind", [0, 1])
def test_row_blocks_on_one_device_match_oracle_and_one_block(wb, orc, case, P, fused, kind):
    m, bc, cl, one = case["m"], case["bc"], case["cl"], case["one"]
    n = m.n_nodes
    g = _setup(wb.Context(n_gpus=P, devices=[0] * P, fused_gather_on_shared_device=fused, gmres_tol=1e-12,
                          gmres_max_steps=400, precond_kind=kind), m, cl)
    assert (g.row0, g.row1) == (0, n)
    # every row block's rows, against the oracle (1e-11) and bitwise against the one-block context
    gn, gd = g.get_rows(0), g.get_rows(1)
    oalpha = orc.compute_alpha(case["on"])
    assert rel_err_rowscaled(gn, case["on"], diag=oalpha) < 1e-11
    assert rel_err_rowscaled(gd, case["od"]) < 1e-11
    assert np.array_equal(gn, one.get_rows(0)) and np.array_equal(gd, one.get_rows(1))
    # a slab that straddles a block boundary
    chunk = -(-n // P)
    a, b = max(0, chunk - 5), min(n, chunk + 7)
    assert np.array_equal(g.get_rows(0, a, b), gn[a:b])
    # alpha: gathered slices
    assert np.array_equal(g.get_alpha(), one.get_alpha())
    assert np.abs(g.get_alpha() - oalpha).max() < 1e-12
    # operator applications: local rows of every block, gathered
    x = np.sin(0.37 * np.arange(n))
    con = _orc_con(orc, cl)
    ref_y = orc.constrained_vmult(case["on"], case["od"], oalpha, m.surface_nodes, m.other_nodes, con, x)
    y = g.constrained_vmult(x)
    assert np.abs(y - ref_y).max() < 1e-11 * max(1.0, np.abs(ref_y).max())
    assert np.array_equal(y, one.constrained_vmult(x))
    assert np.array_equal(g.compute_rhs(bc), one.compute_rhs(bc))
    assert np.array_equal(g.vmult(x), one.vmult(x))
    # solve_system: same iterations and bitwise the same solution as one block; oracle within tolerance
    z = np.zeros(n)
    one.set_precond_kind(kind)
    phi1, dphi1, it1, _ = one.solve_system(z, z, bc)
    phi, dphi, it, res = g.solve_system(z, z, bc)
    assert it == it1 and np.array_equal(phi, phi1) and np.array_equal(dphi, dphi1)
    sol = g.get_sol()
    assert np.linalg.norm(sol - case["ref"]["sol"]) <= 1e-9 * np.linalg.norm(case["ref"]["sol"])
    if kind == 0:
        assert it == case["ref"]["iters"]
        assert np.array_equal(g.get_band(), one.get_band())
    # residual of the converged pair
    s = m.surface_nodes == 1
    r = g.residual(np.where(s, bc, phi), np.where(s, dphi, bc))
    assert np.abs(r).max() < 1e-9
    assert np.array_equal(r, one.residual(np.where(s, bc, phi), np.where(s, dphi, bc)))
    g.close()


def test_group_solve_with_library_constraints_and_moving_mesh(wb, orc, case):
    """solve() = geometry upload + assemble_system + solve_system with compute_constraints inside the
    library (auto_constraints), re-run after the mesh moved, 3 blocks vs 1 block."""
    m, bc = case["m"], case["bc"]
    n = m.n_nodes
    kw = dict(gmres_tol=1e-11, gmres_max_steps=400, precond_kind=1, auto_constraints=1)
    one = wb.Context(**kw)
    g = wb.Context(n_gpus=3, devices=[0, 0, 0], **kw)
    for c in (one, g):
        c.set_topology(n, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
        c.set_masks(m.surface_nodes, m.other_nodes)
    z = np.zeros(n)
    xyz2 = m.xyz.copy()
    fs = m.surface_nodes == 1
    xyz2[fs, 2] += 0.01 * np.cos(2.0 * xyz2[fs, 0])
    for xyz in (m.xyz, xyz2):
        p1, d1, it1, _ = one.solve(xyz, z, z, bc)
        p3, d3, it3, _ = g.solve(xyz, z, z, bc)
        assert it1 == it3 and np.array_equal(p1, p3) and np.array_equal(d1, d3)
    assert np.array_equal(one.compute_normals(), g.compute_normals())
    assert g.timings()["kernel_launches"] > 2 * one.timings()["kernel_launches"]
    one.close()
    g.close()


def test_group_reports_errors_instead_of_hanging(wb, case):
    m = case["m"]
    g = wb.Context(n_gpus=2, devices=[0, 0])
    with pytest.raises(wb.WbemError):
        g.assemble()                      # before set_topology / set_geometry: every block refuses
    g.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    with pytest.raises(wb.WbemError):
        g.get_rows(0, 5, m.n_nodes + 3)   # out of range
    with pytest.raises(wb.WbemError):
        g.ipc_export()                    # a one-process context has no IPC exchange
    g.close()
    with pytest.raises(wb.WbemError):
        wb.Context(n_gpus=2, devices=[0, 99])


def test_two_quadrature_orders_alive_on_one_device(wb, orc):
    """The quadrature tables are per context (not __constant__): a second context with other orders
    must not change what the first one integrates with."""
    m = meshgen.wigley_tank(nxm=10, nt=5, nxu=4, nxd=5, nz=3, nzh=3)
    a = wb.Context(quad_order=4, sing_order=5)
    b = wb.Context(quad_order=3, sing_order=7)
    for c in (a, b):
        c.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
        c.set_geometry(m.xyz)
    a.assemble()
    b.assemble()
    a.assemble()  # after b's tables were uploaded
    for c, (q, s) in ((a, (4, 5)), (b, (3, 7))):
        on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, quad_order=q, sing_order=s)
        assert rel_err_rowscaled(c.get_rows(1), od) < 1e-11
        assert rel_err_rowscaled(c.get_rows(0), on, diag=orc.compute_alpha(on)) < 1e-11
        nrm = c.compute_normals()
        assert np.abs(nrm - orc.compute_normals(m.xyz, m.cells, m.dir_flag, quad_order=q)).max() < 1e-11
    a.close()
    b.close()


def test_group_across_devices(wb, orc, case):
    """>= 2 GPUs: the same single-process context across devices, fused peer-store mat-vec."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs (the shared-device variants above cover the logic on one)")
    m, bc, cl, one = case["m"], case["bc"], case["cl"], case["one"]
    z = np.zeros(m.n_nodes)
    for kind in (0, 1):
        one.set_precond_kind(kind)
        phi1, dphi1, it1, _ = one.solve_system(z, z, bc)
        for P in sorted({2, min(ndev, 4), ndev}):
            g = _setup(wb.Context(n_gpus=P, gmres_tol=1e-12, gmres_max_steps=400, precond_kind=kind), m, cl)
            assert np.array_equal(g.get_rows(0), one.get_rows(0))
            phi, dphi, it, _ = g.solve_system(z, z, bc)
            assert it == it1 and np.array_equal(phi, phi1) and np.array_equal(dphi, dphi1)
            g.close()
