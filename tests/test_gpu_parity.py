"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the
same inputs -- and, at BASELINE.json's full size, through size-independent properties.

Tolerances (north_star): matrix entries 1e-11 relative (row-scaled, see conftest.rel_err_rowscaled
and DESIGN.md "Parity metric"), GMRES solution within the solver tolerance, hull
drag/potential 1e-8 relative.
"""
import numpy as np
import pytest

from conftest import make_problem, rel_err_rowscaled
from wavebem_b200 import meshgen
from oracle.postproc import hull_pressure_force

pytestmark = pytest.mark.gpu

ENTRY_TOL = 1e-11


def _ctx(wb, mesh, **params):
    ctx = wb.Context(**params)
    ctx.set_topology(mesh.n_nodes, mesh.cells, mesh.dir_flag, mesh.dn_ptr, mesh.dn_idx)
    ctx.set_geometry(mesh.xyz)
    return ctx


def _orc_con(orc, cl):
    return orc.Constraints(cl.n, cl.lines, cl.ptr, cl.col, cl.val, cl.inhom)


MESHES = {
    "cube1": lambda: meshgen.cube(1),
    "cube5": lambda: meshgen.cube(5),
    "cube4_random_flipped": lambda: meshgen.cube(4, renumber="random", seed=5, flip_every=3),
    "sphere6": lambda: meshgen.sphere(6, radius=0.7, center=(0.1, -0.2, 0.3)),
    "tank": lambda: meshgen.wigley_tank(),
    "tank_random": lambda: meshgen.wigley_tank(nxm=12, nt=6, nxu=4, nxd=6, nz=3, nzh=3, renumber="random", seed=2),
    "tank_wave": lambda: meshgen.wigley_tank(nxm=12, nt=6, nxu=4, nxd=6, nz=3, nzh=3, wave_amp=0.025, wave_phase=0.7),
}


def test_fast_rsqrt_is_accurate(wb):
    ctx = wb.Context()
    rng = np.random.default_rng(0)
    x = np.concatenate([10.0 ** rng.uniform(-12, 12, 200000), rng.uniform(0.5, 2.0, 200000),
                        [1.0, 4.0, 0.25, 1e-300, 1e300]])
    y = ctx.selftest_rsqrt(x)
    ref = 1.0 / np.sqrt(x.astype(np.longdouble))
    rel = np.abs((y - ref) / ref).astype(np.float64)
    assert rel.max() < 2.3e-16, rel.max()  # ~1 ulp


@pytest.mark.parametrize("name", list(MESHES))
def test_assembly_matches_oracle(wb, orc, name):
    m = MESHES[name]()
    ctx = _ctx(wb, m)
    ctx.assemble()
    gn, gd = ctx.get_rows(0), ctx.get_rows(1)
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    oalpha = orc.compute_alpha(on)
    assert rel_err_rowscaled(gd, od) < ENTRY_TOL
    assert rel_err_rowscaled(gn, on, diag=oalpha) < ENTRY_TOL
    # against the long-double arbiter the GPU is as accurate as the reference arithmetic
    ln, ld = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, long_double=True)
    assert rel_err_rowscaled(gd, ld) < max(4 * rel_err_rowscaled(od, ld), 1e-13)
    assert np.abs(gn - ln).max() < max(4 * np.abs(on - ln).max(), 1e-14)
    # alpha (compute_alpha) = -row sums
    alpha = ctx.get_alpha()
    assert np.abs(alpha - oalpha).max() < 1e-12
    ctx.close()


@pytest.mark.parametrize("orders", [(4, 5), (3, 4), (5, 6), (2, 8), (8, 12)])
def test_simple_variant_and_other_quadrature_orders(wb, orc, orders):
    q, s = orders
    m = meshgen.wigley_tank(nxm=10, nt=5, nxu=4, nxd=5, nz=3, nzh=3)
    ctx = _ctx(wb, m, quad_order=q, sing_order=s, assemble_variant=1)
    ctx.assemble()
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, quad_order=q, sing_order=s)
    assert rel_err_rowscaled(ctx.get_rows(1), od) < ENTRY_TOL
    assert rel_err_rowscaled(ctx.get_rows(0), on, diag=orc.compute_alpha(on)) < ENTRY_TOL
    if q != 4:  # default variant falls back to the same generic kernel
        c2 = _ctx(wb, m, quad_order=q, sing_order=s)
        c2.assemble()
        assert rel_err_rowscaled(c2.get_rows(1), od) < ENTRY_TOL
        c2.close()
    ctx.close()


def test_tiled_and_simple_kernels_agree_and_tiled_is_deterministic(wb):
    m = meshgen.wigley_tank(renumber="random", seed=9)
    a = _ctx(wb, m)
    a.assemble()
    n1, d1 = a.get_rows(0), a.get_rows(1)
    a.set_geometry(m.xyz)
    a.assemble()
    assert np.array_equal(n1, a.get_rows(0)) and np.array_equal(d1, a.get_rows(1))  # bitwise
    b = _ctx(wb, m, assemble_variant=1)
    b.assemble()
    assert rel_err_rowscaled(n1, b.get_rows(0), diag=a.get_alpha()) < 1e-12
    assert rel_err_rowscaled(d1, b.get_rows(1)) < 1e-12
    a.close()
    b.close()


@pytest.mark.parametrize("make", [lambda: meshgen.wigley_tank(), lambda: meshgen.wigley_tank_for_nodes(6000),
                                  lambda: meshgen.wigley_tank(renumber="random", seed=4)])
def test_stream_kernel_order_independent_and_equal_to_colour_kernel(wb, orc, monkeypatch, make):
    """The single-launch kernel (persistent CTAs, ticket order, dependency flags between clusters sharing a
    column) gives bitwise the same matrices for every item order (row tiles grouped by 1, 2 or 5) and on a
    repeated assembly; the colour-per-launch kernel (assemble_variant = 2: per-point arithmetic, the fallback
    for caller-supplied FEValues) agrees to rounding; both match the oracle."""
    m = make()
    res = []
    for g in ("1", "2", "5"):
        monkeypatch.setenv("WBEM_ASM_GROUP", g)
        c = _ctx(wb, m)
        c.assemble()
        res.append((c.get_rows(0), c.get_rows(1), c.get_alpha()))
        if g == "2":
            c.assemble()
            assert np.array_equal(res[-1][0], c.get_rows(0)) and np.array_equal(res[-1][1], c.get_rows(1))
        c.close()
    monkeypatch.delenv("WBEM_ASM_GROUP")
    for r in res[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(res[0], r))
    c2 = _ctx(wb, m, assemble_variant=2)
    c2.assemble()
    assert rel_err_rowscaled(res[0][0], c2.get_rows(0), diag=res[0][2]) < 1e-13
    assert rel_err_rowscaled(res[0][1], c2.get_rows(1)) < 1e-13
    assert np.abs(res[0][2] - c2.get_alpha()).max() < 1e-13
    r0, r1 = m.n_nodes // 3, min(m.n_nodes, m.n_nodes // 3 + 96)
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, r0, r1)
    for got in (res[0], (c2.get_rows(0), c2.get_rows(1))):
        assert rel_err_rowscaled(got[1][r0:r1], od) < ENTRY_TOL
        assert rel_err_rowscaled(got[0][r0:r1], on, diag=orc.compute_alpha(on)) < ENTRY_TOL
    c2.close()


@pytest.fixture(scope="module")
def tank_case(wb, orc):
    m = meshgen.wigley_tank()
    bc, nn, cl = make_problem(m)
    ctx = _ctx(wb, m)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(cl)
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    alpha = orc.compute_alpha(on)
    yield dict(m=m, bc=bc, cl=cl, con=_orc_con(orc, cl), ctx=ctx, on=on, od=od, alpha=alpha)
    ctx.close()


def test_operator_applications_match_oracle(wb, orc, tank_case):
    t = tank_case
    m, ctx = t["m"], t["ctx"]
    s, o = m.surface_nodes, m.other_nodes
    for x in (np.sin(0.37 * np.arange(m.n_nodes)), t["bc"], np.ones(m.n_nodes)):
        for g, r in ((ctx.vmult(x), orc.vmult(t["on"], t["od"], t["alpha"], s, o, x)),
                     (ctx.compute_rhs(x), orc.compute_rhs(t["on"], t["od"], t["alpha"], s, o, x)),
                     (ctx.constrained_vmult(x), orc.constrained_vmult(t["on"], t["od"], t["alpha"], s, o, t["con"], x))):
            assert np.abs(g - r).max() <= 1e-12 * max(1.0, np.abs(r).max())
    rhs = ctx.compute_rhs(t["bc"])
    assert np.array_equal(ctx.distribute_rhs(rhs), orc.distribute_rhs(t["con"], rhs))
    # linearity (size-independent property)
    x, y = np.cos(0.1 * np.arange(m.n_nodes)), np.sin(0.3 * np.arange(m.n_nodes))
    lhs = ctx.constrained_vmult(2.5 * x - 0.75 * y)
    rhs2 = 2.5 * ctx.constrained_vmult(x) - 0.75 * ctx.constrained_vmult(y)
    assert np.abs(lhs - rhs2).max() < 1e-12


def test_arbitrary_masks_and_pure_neumann(wb, orc, tank_case):
    """Masks need not be 0/1 or complementary for vmult/compute_rhs to follow the reference
    formula; an all-Neumann mask triggers the -||dst|| shift (bem_problem.cc:667-668)."""
    t = tank_case
    m = t["m"]
    ctx = _ctx(wb, m)
    ctx.assemble()
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, m.n_nodes)
    for s, o in ((np.zeros(m.n_nodes), np.ones(m.n_nodes)),
                 (rng.integers(0, 2, m.n_nodes).astype(float), rng.uniform(0, 1, m.n_nodes)),
                 (np.ones(m.n_nodes), np.zeros(m.n_nodes))):
        ctx.set_masks(s, o)
        r = orc.vmult(t["on"], t["od"], t["alpha"], s, o, x)
        assert np.abs(ctx.vmult(x) - r).max() <= 1e-12 * max(1.0, np.abs(r).max())
        r = orc.compute_rhs(t["on"], t["od"], t["alpha"], s, o, x)
        assert np.abs(ctx.compute_rhs(x) - r).max() <= 1e-12 * max(1.0, np.abs(r).max())
    ctx.close()


@pytest.mark.parametrize("on_host", [0, 1])
def test_band_preconditioner_matches_oracle(wb, orc, tank_case, on_host):
    t = tank_case
    m = t["m"]
    ctx = _ctx(wb, m, precond_on_host=on_host)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(t["cl"])
    ctx.assemble_preconditioner()
    dense = orc.band_system_dense(t["on"], t["od"], t["alpha"], m.surface_nodes, t["con"], band=100)
    band = ctx.get_band()
    n = m.n_nodes
    for r in (0, 1, 49, 50, 51, n // 2, n - 51, n - 50, n - 1):
        for k in range(100):
            i = r - 50 + 1 + k
            ref = dense[r, i] if 0 <= i < n else 0.0
            assert abs(band[r, k] - ref) <= 1e-12 * max(1.0, abs(ref))
    v = np.sin(0.11 * np.arange(n))
    z = ctx.precond_vmult(v)
    zo = orc.precond_apply(t["on"], t["od"], t["alpha"], m.surface_nodes, t["con"], v, band=100)
    assert np.abs(z - zo).max() <= 1e-9 * np.abs(zo).max()
    assert np.abs(dense @ z - v).max() < 1e-10
    ctx.close()


@pytest.mark.parametrize("tol,max_steps", [(1e-10, 400), (1e-16, 200)])
def test_solve_system_matches_oracle(wb, orc, tank_case, tol, max_steps):
    t = tank_case
    m = t["m"]
    n = m.n_nodes
    ctx = _ctx(wb, m, gmres_tol=tol, gmres_max_steps=max_steps)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(t["cl"])
    phi0 = np.where(m.surface_nodes == 1, t["bc"], 0.0)
    dphi0 = np.where(m.surface_nodes == 1, 0.0, t["bc"])
    phi, dphi, iters, res = ctx.solve_system(phi0, dphi0, t["bc"])
    ref = orc.solve_system(t["on"], t["od"], m.surface_nodes, m.other_nodes, t["bc"], t["con"], phi0, dphi0,
                           tol=tol, max_steps=max_steps)
    assert ref["converged"] and res <= tol
    assert abs(iters - ref["iters"]) <= max(3, ref["iters"] // 10)
    assert np.abs(ctx.get_system_rhs() - ref["rhs"]).max() < 1e-12
    scale = np.linalg.norm(ref["sol"])
    # "within the solver tolerance": both are tol-accurate solutions of the same system
    assert np.linalg.norm(ctx.get_sol() - ref["sol"]) <= max(50 * tol, 1e-11) * max(scale, 1.0)
    # only the unknown half is overwritten (bem_problem.cc:869-879)
    s = m.surface_nodes == 1
    assert np.array_equal(phi[s], phi0[s]) and np.array_equal(dphi[~s], dphi0[~s])
    # hull drag / potential functional, 1e-8 relative
    vinf = np.array([0.28 * np.sqrt(9.81 * 2.5), 0, 0])
    fg, pg = hull_pressure_force(m, phi, dphi, vinf)
    fo, po = hull_pressure_force(m, np.where(s, phi0, ref["phi"]), np.where(s, ref["dphi_dn"], dphi0), vinf)
    assert abs(fg[0] - fo[0]) <= 1e-8 * abs(fo[0]) and abs(pg - po) <= 1e-8 * abs(po)
    # the BIE residual of the converged pair vanishes
    r = ctx.residual(phi, dphi)
    ro = orc.residual(t["on"], t["od"], m.surface_nodes, m.other_nodes, t["con"], phi, dphi)
    assert np.abs(r - ro).max() < 1e-11 and np.abs(r).max() < max(1e3 * tol, 1e-11)
    ctx.close()


def test_no_convergence_is_reported(wb, tank_case):
    t = tank_case
    m = t["m"]
    ctx = _ctx(wb, m, gmres_tol=1e-30, gmres_max_steps=9)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(t["cl"])
    z = np.zeros(m.n_nodes)
    with pytest.raises(wb.NoConvergence) as e:
        ctx.solve_system(z, z, t["bc"])
    assert e.value.last_step == 9
    ctx.close()


def test_calls_out_of_order_fail_loudly(wb):
    m = meshgen.cube(2)
    ctx = wb.Context()
    with pytest.raises(wb.WbemError):
        ctx._chk(wb.lib().wbem_assemble(ctx._h))
    ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    with pytest.raises(wb.WbemError, match="set_geometry"):
        ctx.assemble()
    ctx.set_geometry(m.xyz)
    with pytest.raises(wb.WbemError, match="assemble"):
        ctx.vmult(np.zeros(m.n_nodes))
    ctx.close()


def test_bem_problem_mirror_api(wb, orc):
    """The reference-shaped class: reinit / solve / solve_system / vmult / residual."""
    m = meshgen.wigley_tank(nxm=12, nt=6, nxu=4, nxd=6, nz=3, nzh=3)
    m.nodes_normals = meshgen.cell_normals_at_nodes(m)
    bc = meshgen.towing_tank_bc(m)
    bem = wb.BEMProblem(m, gmres_tol=1e-12, gmres_max_steps=300)
    bem.reinit()
    n = m.n_nodes
    phi, dphi = np.zeros(n), np.zeros(n)
    bem.solve(phi, dphi, bc)
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    con = _orc_con(orc, bem.constraints)
    ref = orc.solve_system(on, od, m.surface_nodes, m.other_nodes, bc, con, np.zeros(n), np.zeros(n),
                           tol=1e-12, max_steps=300)
    assert np.linalg.norm(bem.sol - ref["sol"]) < 1e-9 * np.linalg.norm(ref["sol"])
    assert np.abs(bem.alpha - ref["alpha"]).max() < 1e-12
    assert rel_err_rowscaled(bem.neumann_matrix(5, 9), on[5:9], diag=ref["alpha"][5:9]) < ENTRY_TOL
    # solve_system again with changed boundary data (the J.v pattern): matrices are reused
    phi2, dphi2 = np.zeros(n), np.zeros(n)
    bem.solve_system(phi2, dphi2, 2 * bc)
    assert np.linalg.norm(bem.sol - 2 * ref["sol"]) < 1e-8 * np.linalg.norm(ref["sol"])
    dst = np.zeros(n)
    bem.vmult(dst, ref["sol"])
    assert np.abs(dst - orc.vmult(on, od, ref["alpha"], m.surface_nodes, m.other_nodes, ref["sol"])).max() < 1e-12
    res = np.zeros(n)
    bem.residual(res, np.where(m.surface_nodes == 1, bc, phi), np.where(m.surface_nodes == 1, dphi, bc))
    assert np.abs(res).max() < 1e-9
    bem2 = wb.BEMProblem(m, gmres_tol=1e-30, gmres_max_steps=5)
    bem2.reinit()
    with pytest.raises(wb.NoConvergence):
        bem2.solve(np.zeros(n), np.zeros(n), bc)
    # a domain without its own nodes_normals: compute_constraints runs inside the library
    # (auto_constraints, bem_problem.cc:845) and produces the same lines here (flat free surface)
    lines_host = _lines_dict(bem.constraints)
    del m.nodes_normals
    bem3 = wb.BEMProblem(m, gmres_tol=1e-12, gmres_max_steps=300)
    bem3.reinit()
    phi3, dphi3 = np.zeros(n), np.zeros(n)
    bem3.solve(phi3, dphi3, bc)
    assert np.array_equal(phi3, phi) and np.array_equal(dphi3, dphi)
    got = _lines_dict(bem3.constraints)
    assert got.keys() == lines_host.keys()
    nrm = bem3.compute_normals()
    assert np.abs(nrm - orc.compute_normals(m.xyz, m.cells, m.dir_flag)).max() < 1e-11


def test_hanging_node_lines(wb, orc):
    """Constraint lines with two masters and weight 1/2 (make_hanging_node_constraints)."""
    from wavebem_b200.constraints import compute_constraints
    m = meshgen.cube(4)
    n = m.n_nodes
    top = m.node_patch == m.patch_names.index("z1")
    s, o = top.astype(float), 1.0 - top
    interior = np.nonzero(~m.node_on_patch_boundary & ~top)[0]
    hang = [(int(interior[3]), [(int(interior[2]), 0.5), (int(interior[4]), 0.5)])]
    bc = np.sin(np.arange(n) * 0.2)
    nn = meshgen.cell_normals_at_nodes(m)
    cl = compute_constraints(m.dn_ptr, m.dn_idx, s, bc, nodes_normals=nn, hanging=hang)
    ctx = _ctx(wb, m, gmres_tol=1e-12, gmres_max_steps=400)
    ctx.assemble()
    ctx.set_masks(s, o)
    ctx.set_constraints(cl)
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    con = _orc_con(orc, cl)
    alpha = orc.compute_alpha(on)
    x = np.cos(np.arange(n) * 0.3)
    assert np.abs(ctx.constrained_vmult(x) - orc.constrained_vmult(on, od, alpha, s, o, con, x)).max() < 1e-12
    z = np.zeros(n)
    _, _, it, _ = ctx.solve_system(z, z, bc)
    ref = orc.solve_system(on, od, s, o, bc, con, z, z, tol=1e-12, max_steps=400)
    assert np.linalg.norm(ctx.get_sol() - ref["sol"]) < 1e-9 * np.linalg.norm(ref["sol"])
    h = hang[0]
    assert abs(ctx.get_sol()[h[0]] - 0.5 * (ctx.get_sol()[h[1][0][0]] + ctx.get_sol()[h[1][1][0]])) < 1e-10
    ctx.close()


def test_full_size_20k_properties_and_row_slab_parity(wb, orc):
    """BASELINE configs[1] size (N ~ 20k): entries of a row slab against the oracle, and
    size-independent properties of the whole path."""
    m = meshgen.wigley_tank_for_nodes(20000)
    bc, nn, cl = make_problem(m)
    n = m.n_nodes
    ctx = _ctx(wb, m, gmres_tol=1e-10, gmres_max_steps=400)
    ctx.assemble()
    for r0 in (0, n // 2 - 64, n - 128):
        on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, r0, r0 + 128)
        assert rel_err_rowscaled(ctx.get_rows(0, r0, r0 + 128), on, diag=orc.compute_alpha(on)) < ENTRY_TOL
        assert rel_err_rowscaled(ctx.get_rows(1, r0, r0 + 128), od) < ENTRY_TOL
    alpha = ctx.get_alpha()
    flat = np.isin(m.node_patch, [m.patch_names.index(k) for k in ("bottom", "fs_up", "fs_down")]) & \
        ~m.node_on_patch_boundary
    assert np.abs(alpha[flat] - 0.5).max() < 1e-3   # solid angle of a smooth point
    assert alpha.min() > 0 and alpha.max() <= 1.0 + 1e-4   # ~1 at the thin ends of the keel
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(cl)
    x, y = np.cos(0.1 * np.arange(n)), np.sin(0.3 * np.arange(n))
    assert np.abs(ctx.constrained_vmult(x + 2 * y) - (ctx.constrained_vmult(x) + 2 * ctx.constrained_vmult(y))).max() < 1e-11
    z = np.zeros(n)
    phi, dphi, it, res = ctx.solve_system(z, z, bc)
    assert res <= 1e-10 and it < 400
    s = m.surface_nodes == 1
    r = ctx.residual(np.where(s, bc, phi), np.where(s, dphi, bc))
    assert np.abs(r).max() < 1e-8
    # re-solving with the same data reproduces the result (cached preconditioner path)
    phi2, dphi2, it2, _ = ctx.solve_system(z, z, bc)
    assert it2 == it and np.array_equal(phi2, phi)
    ctx.close()


def test_pure_neumann_with_constraints(wb, orc):
    """All-Neumann cube: the -||dst|| shift uses the norm over ALL rows of the raw product
    (bem_problem.cc:667-668) even though constrained rows are overwritten afterwards."""
    from wavebem_b200.constraints import compute_constraints
    m = meshgen.cube(5)
    n = m.n_nodes
    s, o = np.zeros(n), np.ones(n)
    bc = np.sin(0.3 * np.arange(n))
    cl = compute_constraints(m.dn_ptr, m.dn_idx, s, bc)
    ctx = _ctx(wb, m)
    ctx.assemble()
    ctx.set_masks(s, o)
    ctx.set_constraints(cl)
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    alpha = orc.compute_alpha(on)
    x = np.cos(0.2 * np.arange(n))
    ref = orc.constrained_vmult(on, od, alpha, s, o, _orc_con(orc, cl), x)
    assert np.abs(ctx.constrained_vmult(x) - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
    ctx.close()


@pytest.mark.parametrize("n_tmp,band", [(12, 100), (100, 0), (30, 40), (400, 0)])
def test_gmres_restart_and_preconditioner_options(wb, orc, tank_case, n_tmp, band):
    """Short restart cycles (AdditionalData(n_tmp)) and other band widths, incl. no preconditioner."""
    t = tank_case
    m = t["m"]
    n = m.n_nodes
    tol, steps = 1e-9, 3000
    ctx = _ctx(wb, m, gmres_tol=tol, gmres_max_steps=steps, gmres_n_tmp_vectors=n_tmp, preconditioner_band=band)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(t["cl"])
    z = np.zeros(n)
    try:
        phi, dphi, it, res = ctx.solve_system(z, z, t["bc"])
        conv = True
    except wb.NoConvergence:
        conv = False
    ref = orc.solve_system(t["on"], t["od"], m.surface_nodes, m.other_nodes, t["bc"], t["con"], z, z, tol=tol,
                           max_steps=steps, n_tmp_vectors=n_tmp, band=max(band, 2), use_precond=band > 0)
    if conv != ref["converged"]:
        # GMRES(10) crawls on this system (~3000 iterations): rounding decides on which side of the step limit
        # it ends.  The side that converged must have needed (almost) all the steps.
        assert (it if conv else ref["iters"]) >= 0.9 * steps
        ctx.close()
        return
    if conv:
        assert abs(it - ref["iters"]) <= max(5, ref["iters"] // 5)
        assert np.linalg.norm(ctx.get_sol() - ref["sol"]) <= 1e-6 * np.linalg.norm(ref["sol"])
    ctx.close()


def test_tiny_and_ragged_sizes(wb, orc):
    """N far below one tile / one 64-block, N not a multiple of anything, a single cell."""
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0.3]], dtype=float)
    cells = np.array([[0, 1, 2, 3]], dtype=np.uint32)
    ptr, idx = np.arange(5, dtype=np.uint32), np.arange(4, dtype=np.uint32)
    ctx = wb.Context()
    ctx.set_topology(4, cells, np.ones(1, np.uint8), ptr, idx)
    ctx.set_geometry(xyz)
    ctx.assemble()
    on, od = orc.assemble_rows(xyz, cells, np.ones(1, np.uint8), ptr, idx)
    assert rel_err_rowscaled(ctx.get_rows(0), on, diag=np.ones(4)) < ENTRY_TOL
    assert rel_err_rowscaled(ctx.get_rows(1), od) < ENTRY_TOL
    ctx.close()
    for make in (lambda: meshgen.cube(2), lambda: meshgen.sphere(3), lambda: meshgen.wigley_tank(nxm=7, nt=3, nxu=2, nxd=3, nz=2, nzh=2)):
        m = make()
        bc, nn, cl = make_problem(m, bc=np.cos(np.arange(m.n_nodes) * 0.4)) if m.surface_nodes is not None else (None, None, None)
        s = m.surface_nodes if m.surface_nodes is not None else (m.node_patch == 0).astype(float)
        o = 1.0 - s
        if cl is None:
            from wavebem_b200.constraints import compute_constraints
            bc = np.cos(np.arange(m.n_nodes) * 0.4)
            cl = compute_constraints(m.dn_ptr, m.dn_idx, s, bc, nodes_normals=meshgen.cell_normals_at_nodes(m))
        ctx = _ctx(wb, m, gmres_tol=1e-11, gmres_max_steps=500)
        ctx.assemble()
        ctx.set_masks(s, o)
        ctx.set_constraints(cl)
        on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
        z = np.zeros(m.n_nodes)
        _, _, it, _ = ctx.solve_system(z, z, bc)
        ref = orc.solve_system(on, od, s, o, bc, _orc_con(orc, cl), z, z, tol=1e-11, max_steps=500)
        assert np.linalg.norm(ctx.get_sol() - ref["sol"]) <= 1e-8 * max(1.0, np.linalg.norm(ref["sol"]))
        ctx.close()


def test_reinit_with_another_mesh_and_moving_geometry(wb, orc):
    """reinit() after a remesh (free_surface.cc:1674) and re-assembly after the nodes moved
    (free_surface.cc:5306-5307) on the same context."""
    ctx = wb.Context()
    for m in (meshgen.cube(3), meshgen.wigley_tank(nxm=10, nt=5, nxu=4, nxd=5, nz=3, nzh=3)):
        ctx.set_topology(m.n_nodes, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
        for amp in (0.0, 0.03):
            xyz = m.xyz.copy()
            xyz[:, 2] += amp * np.sin(xyz[:, 0]) * np.cos(xyz[:, 1])
            ctx.set_geometry(xyz)
            ctx.assemble()
            on, od = orc.assemble_rows(xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
            assert rel_err_rowscaled(ctx.get_rows(0), on, diag=orc.compute_alpha(on)) < ENTRY_TOL
            assert rel_err_rowscaled(ctx.get_rows(1), od) < ENTRY_TOL
    ctx.close()


def _dense_constrained_operator(t):
    """The matrix GMRES sees: merged operator (bem_problem.cc:620-670 with 0/1 masks) with the
    constrained rows of ConstrainedOperator::vmult (constrained_matrix.h:73-86)."""
    m, cl = t["m"], t["cl"]
    s, o = m.surface_nodes, m.other_nodes
    n = m.n_nodes
    A = t["on"] * o[None, :] - t["od"] * s[None, :]
    A[np.arange(n), np.arange(n)] += t["alpha"] * o
    for k, line in enumerate(cl.lines):
        A[line, :] = 0.0
        A[line, line] = 1.0
        for c, v in zip(cl.col[cl.ptr[k]:cl.ptr[k + 1]], cl.val[cl.ptr[k]:cl.ptr[k + 1]]):
            A[line, c] -= v
    return A


def test_spai_preconditioner_rows_and_solve(wb, orc, tank_case):
    """precond_kind = 1 (spai.cu): every row of M solves its local system A[S,S]^T m = e_i; the
    GMRES solution agrees with the oracle's (band-preconditioned) within the solver tolerance
    and needs far fewer iterations."""
    t = tank_case
    m = t["m"]
    n = m.n_nodes
    tol = 1e-10
    ctx = _ctx(wb, m, precond_kind=1, gmres_tol=tol, gmres_max_steps=400)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(t["cl"])
    ctx.assemble_preconditioner()
    nbr, val, n_sing = ctx.get_spai()
    assert n_sing == 0 and nbr.shape == (n, 32)
    A = _dense_constrained_operator(t)
    M = np.zeros((n, n))
    for i in range(n):
        S = nbr[i][nbr[i] != 0xFFFFFFFF].astype(np.int64)
        assert i in S and np.all(np.diff(S) > 0)
        ref = np.linalg.solve(A[np.ix_(S, S)].T, (S == i).astype(float))
        got = val[i][: len(S)]
        assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max(), i
        M[i, S] = got
    # row i of M A is e_i on S_i
    i = n // 2
    S = nbr[i].astype(np.int64)
    assert np.abs((M[i] @ A)[S] - (S == i)).max() < 1e-10
    v = np.sin(0.11 * np.arange(n))
    assert np.abs(ctx.precond_vmult(v) - M @ v).max() <= 1e-12 * np.abs(M @ v).max()
    phi0 = np.where(m.surface_nodes == 1, t["bc"], 0.0)
    dphi0 = np.where(m.surface_nodes == 1, 0.0, t["bc"])
    phi, dphi, iters, res = ctx.solve_system(phi0, dphi0, t["bc"])
    ref = orc.solve_system(t["on"], t["od"], m.surface_nodes, m.other_nodes, t["bc"], t["con"], phi0, dphi0,
                           tol=tol, max_steps=400)
    assert ref["converged"] and res <= tol
    assert iters <= ref["iters"] // 2, (iters, ref["iters"])
    scale = np.linalg.norm(ref["sol"])
    assert np.linalg.norm(ctx.get_sol() - ref["sol"]) <= 50 * tol * max(scale, 1.0)
    vinf = np.array([0.28 * np.sqrt(9.81 * 2.5), 0, 0])
    s = m.surface_nodes == 1
    fg, pg = hull_pressure_force(m, phi, dphi, vinf)
    fo, po = hull_pressure_force(m, np.where(s, phi0, ref["phi"]), np.where(s, ref["dphi_dn"], dphi0), vinf)
    assert abs(fg[0] - fo[0]) <= 1e-8 * abs(fo[0]) and abs(pg - po) <= 1e-8 * abs(po)
    # a second solve on the unchanged operator reuses the preconditioner and reproduces the result
    phi2, dphi2, iters2, _ = ctx.solve_system(phi0, dphi0, t["bc"])
    assert iters2 == iters and np.array_equal(phi2, phi) and np.array_equal(dphi2, dphi)
    ctx.close()


def test_spai_gmres_batches_restarts_and_step_limit(wb, orc, tank_case):
    """With the sparse approximate inverse the host enqueues 4 iterations ahead and the Arnoldi step's small
    algebra runs on the device: restart cycles shorter than, equal to and not a multiple of the batch, a step
    limit that falls inside a batch (NoConvergence with exactly max_steps iterations), and the same solution as
    one long cycle."""
    t = tank_case
    m = t["m"]
    n = m.n_nodes
    z = np.zeros(n)
    tol = 1e-10

    def run(n_tmp, steps):
        ctx = _ctx(wb, m, precond_kind=1, gmres_tol=tol, gmres_max_steps=steps, gmres_n_tmp_vectors=n_tmp)
        ctx.assemble()
        ctx.set_masks(m.surface_nodes, m.other_nodes)
        ctx.set_constraints(t["cl"])
        try:
            _, _, it, res = ctx.solve_system(z, z, t["bc"])
            out = (True, it, res, ctx.get_sol())
        except wb.NoConvergence:
            out = (False, int(ctx.timings()["gmres_iters"]), None, ctx.get_sol())
        ctx.close()
        return out

    ok, it_long, res, x_long = run(100, 400)
    assert ok and res <= tol and it_long < 40
    for n_tmp in (5, 6, 9, 12):  # cycles of 3, 4, 7, 10 inner iterations
        ok, it, res, x = run(n_tmp, 2000)
        assert ok and res <= tol and it >= it_long
        assert np.linalg.norm(x - x_long) <= 1e-7 * np.linalg.norm(x_long)
    for steps in (it_long - 1, it_long - 2, 5, 1):
        ok, it, _, _ = run(100, steps)
        assert not ok and it == steps
    ok, it, _, _ = run(100, it_long)
    assert ok and it == it_long


@pytest.mark.parametrize("mesh", ["cube1", "cube4_random_flipped"])
def test_spai_on_tiny_and_randomly_numbered_meshes(wb, orc, mesh):
    """Fewer dofs than K (padding slots) and a random numbering (the band preconditioner depends
    on the numbering, the local inverse does not)."""
    m = MESHES[mesh]()
    n = m.n_nodes
    from wavebem_b200.constraints import compute_constraints
    s = (m.node_patch == 1).astype(float)            # one Dirichlet face, five Neumann faces
    m.surface_nodes, m.other_nodes = s, 1.0 - s
    bc = np.cos(np.arange(n) * 0.4)
    cl = compute_constraints(m.dn_ptr, m.dn_idx, s, bc, nodes_normals=meshgen.cell_normals_at_nodes(m))
    sols = {}
    for kind in (0, 1):
        ctx = _ctx(wb, m, precond_kind=kind, gmres_tol=1e-12, gmres_max_steps=300)
        ctx.assemble()
        ctx.set_masks(m.surface_nodes, m.other_nodes)
        ctx.set_constraints(cl)
        z = np.zeros(n)
        _, _, it, res = ctx.solve_system(z, z, bc)
        sols[kind] = (ctx.get_sol(), it)
        if kind == 1:
            nbr, val, n_sing = ctx.get_spai()
            assert n_sing == 0
            assert np.all((nbr != 0xFFFFFFFF).sum(axis=1) == min(32, n))
        ctx.close()
    assert np.linalg.norm(sols[0][0] - sols[1][0]) <= 1e-9 * max(1.0, np.linalg.norm(sols[0][0]))


@pytest.mark.parametrize("name", ["tank", "sphere6", "cube4_random_flipped", "tank_wave"])
def test_normals_and_surface_gradients_match_oracle(wb, orc, name):
    """wbem_compute_normals / wbem_compute_surface_gradients (constraints.cu: node-gathered mass
    matrix + CG in one cooperative kernel) against the oracle's Cholesky solve of the reference's
    two L2 projections (computational_domain.cc:1525-1620, bem_problem.cc:1153-1293)."""
    m = MESHES[name]()
    n = m.n_nodes
    s = m.surface_nodes if m.surface_nodes is not None else (m.node_patch <= 1).astype(float)
    ctx = _ctx(wb, m)
    ctx.set_masks(s, 1.0 - s)
    gn = ctx.compute_normals()
    on = orc.compute_normals(m.xyz, m.cells, m.dir_flag)
    assert np.abs(gn - on).max() < 1e-11
    assert 0 < ctx.mass_cg_iterations() < 200   # (a flat patch with constant normals needs one step)
    f = np.cos(1.3 * m.xyz[:, 0]) + m.xyz[:, 1] * m.xyz[:, 2]
    gg = ctx.compute_surface_gradients(f)
    og = orc.compute_surface_gradients(m.xyz, m.cells, m.dir_flag, f, s)
    assert np.abs(gg - og).max() <= 1e-11 * max(1.0, np.abs(og).max())
    # linear in tmp_rhs, zero for a zero field
    assert np.abs(ctx.compute_surface_gradients(np.zeros(n))).max() == 0.0
    assert np.abs(ctx.compute_surface_gradients(2 * f) - 2 * gg).max() <= 1e-11 * max(1.0, np.abs(og).max())
    ctx.close()


def _lines_dict(cl):
    return {int(l): (sorted(zip(cl.col[cl.ptr[k]:cl.ptr[k + 1]].tolist(), cl.val[cl.ptr[k]:cl.ptr[k + 1]].tolist())),
                     float(cl.inhom[k])) for k, l in enumerate(cl.lines)}


def test_compute_constraints_inside_the_library(wb, orc):
    """wbem_compute_constraints restates bem_problem.cc:990-1105: same lines as the host
    restatement fed with the oracle's normals / surface gradients -- flat Dirichlet-Dirichlet
    edges (tank free surface), sharp ones (two Dirichlet faces of a cube: inhomogeneities from the
    surface gradients), Dirichlet-Neumann and all-Neumann sets, plus caller-owned hanging lines."""
    from wavebem_b200.constraints import compute_constraints
    cases = []
    t = meshgen.wigley_tank(nxm=12, nt=6, nxu=4, nxd=6, nz=3, nzh=3)
    cases.append((t, t.surface_nodes, meshgen.towing_tank_bc(t), None))
    c = meshgen.cube(4, renumber="random", seed=7)
    sc = (c.node_patch <= 2).astype(float)      # three Dirichlet faces meeting in sharp edges
    cases.append((c, sc, np.sin(2.0 * c.xyz[:, 0]) + c.xyz[:, 1] ** 2 - c.xyz[:, 2], None))
    interior = np.nonzero(~c.node_on_patch_boundary)[0]
    hanging = [(int(interior[3]), [(int(interior[0]), 0.5), (int(interior[1]), 0.5)])]
    cases.append((c, sc, np.cos(c.xyz[:, 1]), hanging))
    # chains, resolved like ConstraintMatrix::close() (:1104): (1) a triple node with one FLAT and one SHARP
    # Dirichlet edge -- two coplanar free-surface patches and a Dirichlet side wall: the flat double refers
    # to a dof the sharp edge turns into an inhomogeneous line; (2) a hanging node whose masters are a
    # double-node pair, one of which is itself constrained
    s2 = np.isin(t.node_patch, [t.patch_names.index(k) for k in
                                ("fs_up", "fs_down", "fs_mid_right", "fs_mid_left", "side_left")]).astype(float)
    cases.append((t, s2, np.sin(t.xyz[:, 0]) + 0.3 * t.xyz[:, 2] + 0.1 * t.xyz[:, 1] ** 2, None))
    pair = next((int(i), int(t_)) for i in range(c.n_nodes) for t_ in c.dn_idx[c.dn_ptr[i]:c.dn_ptr[i + 1]]
                if t_ != i and sc[i] == 0 and sc[t_] == 0)
    cases.append((c, sc, np.cos(c.xyz[:, 1]), [(int(interior[5]), [(pair[0], 0.5), (pair[1], 0.5)])]))
    n_chain_cases = 0
    # the two chain cases really contain chains before close(): the walk alone (resolve_chains disabled)
    # leaves an entry that refers to a constrained dof
    import wavebem_b200.constraints as _con
    keep = _con.resolve_chains
    try:
        _con.resolve_chains = lambda lines: None
        for m, s, bc, hanging in cases[-2:]:
            nrm = orc.compute_normals(m.xyz, m.cells, m.dir_flag)
            grd = orc.compute_surface_gradients(m.xyz, m.cells, m.dir_flag, bc, s)
            raw = _lines_dict(compute_constraints(m.dn_ptr, m.dn_idx, s, bc, nodes_normals=nrm,
                                                  node_surface_gradients=grd, hanging=hanging))
            assert any(col in raw for e, _ in raw.values() for col, _ in e)
    finally:
        _con.resolve_chains = keep
    for m, s, bc, hanging in cases:
        ctx = _ctx(wb, m, gmres_tol=1e-12, gmres_max_steps=400)
        ctx.set_masks(s, 1.0 - s)
        if hanging:
            ctx.set_hanging_constraints(hanging)
        got = ctx.compute_constraints(bc)
        nrm = orc.compute_normals(m.xyz, m.cells, m.dir_flag)
        grd = orc.compute_surface_gradients(m.xyz, m.cells, m.dir_flag, bc, s)
        if hanging:   # the hanging lines are condensed into both projections (computational_domain.cc:1535-1538)
            from oracle import projections
            nrm = projections.l2_projection(0, m.xyz, m.cells, m.dir_flag, hanging=hanging)
            grd = projections.l2_projection(1, m.xyz, m.cells, m.dir_flag, bc * s, hanging=hanging)
        ref = compute_constraints(m.dn_ptr, m.dn_idx, s, bc, nodes_normals=nrm, node_surface_gradients=grd,
                                  hanging=hanging)
        a, b = _lines_dict(got), _lines_dict(ref)
        assert a.keys() == b.keys() and len(a) > 0
        for k in a:
            assert a[k][0] == b[k][0], k
            assert abs(a[k][1] - b[k][1]) <= 1e-10 * max(1.0, abs(b[k][1])), (k, a[k], b[k])
        if m is c and not hanging:
            assert any(not e and ih != 0.0 for e, ih in a.values())      # sharp-edge inhomogeneities present
        assert not any(col in a for e, _ in a.values() for col, _ in e)   # closed: no entry refers to a constrained dof
        n_chain_cases += 1
        # auto_constraints = 1: solve_system computes the lines itself (reference :845) -- same answer
        ctx.assemble()
        z = np.zeros(m.n_nodes)
        phi1, dphi1, it1, _ = ctx.solve_system(z, z, bc)
        auto = _ctx(wb, m, gmres_tol=1e-12, gmres_max_steps=400, auto_constraints=1)
        auto.set_masks(s, 1.0 - s)
        if hanging:
            auto.set_hanging_constraints(hanging)
        auto.assemble()
        phi2, dphi2, it2, _ = auto.solve_system(z, z, bc)
        assert it1 == it2 and np.array_equal(phi1, phi2) and np.array_equal(dphi1, dphi2)
        assert auto.timings()["constraints_ms"] > 0.0
        r = auto.residual(np.where(s == 1, bc, phi2), np.where(s == 1, dphi2, bc))
        assert np.abs(r).max() < 1e-9
        ctx.close()
        auto.close()


def test_gmres_entry_point_and_literal_fevalues_input(wb, orc, tank_case):
    """wbem_gmres = solver.solve(cc, sol, system_rhs, preconditioner) alone (bem_problem.cc:853);
    wbem_set_fevalues = the reference's own FEValues output (:192-196) as the regular-pair input."""
    t = tank_case
    m = t["m"]
    n = m.n_nodes
    ctx = _ctx(wb, m, gmres_tol=1e-10, gmres_max_steps=400)
    ctx.assemble()
    ctx.set_masks(m.surface_nodes, m.other_nodes)
    ctx.set_constraints(t["cl"])
    rhs = ctx.distribute_rhs(ctx.compute_rhs(t["bc"]))
    sol, it, res = ctx.gmres(rhs)
    z = np.zeros(n)
    _, _, it2, _ = ctx.solve_system(z, z, t["bc"])
    assert it == it2 and res <= 1e-10 and np.array_equal(sol, ctx.get_sol())
    assert np.abs(ctx.constrained_vmult(sol) - rhs).max() < 1e-8
    # linearity of the solve in the right-hand side
    sol2, _, _ = ctx.gmres(-2.0 * rhs)
    assert np.linalg.norm(sol2 + 2.0 * sol) <= 1e-7 * np.linalg.norm(sol)
    rows_n, rows_d = ctx.get_rows(0), ctx.get_rows(1)
    uv, w = orc.qgauss2(4)
    nc = m.n_cells
    qp, nr, jw = np.zeros((nc, 16, 3)), np.zeros((nc, 16, 3)), np.zeros((nc, 16))
    for c in range(nc):
        qp[c], nr[c], jw[c], _ = orc.fe_values(m.xyz[m.cells[c].astype(np.int64)], m.dir_flag[c], uv, w)
    ctx.set_fevalues(qp, nr, jw)
    ctx.assemble()
    assert rel_err_rowscaled(ctx.get_rows(1), rows_d) < 1e-13
    assert rel_err_rowscaled(ctx.get_rows(0), rows_n, diag=t["alpha"]) < 1e-13
    assert rel_err_rowscaled(ctx.get_rows(1), t["od"]) < ENTRY_TOL
    assert rel_err_rowscaled(ctx.get_rows(0), t["on"], diag=t["alpha"]) < ENTRY_TOL
    # scaled JxW scales the regular part of D: the values really are taken from the caller
    ctx.set_fevalues(qp, nr, 2.0 * jw)
    ctx.assemble()
    d2 = ctx.get_rows(1)
    single = np.nonzero(np.diff(m.dn_ptr.astype(np.int64)) == 1)[0][::37]   # rows without double nodes
    cells = m.cells.astype(np.int64)
    for i in single:
        regular = np.ones(n, dtype=bool)                                   # columns fed by regular pairs only:
        regular[np.unique(cells[(cells == i).any(axis=1)])] = False        # not a dof of a cell that holds i
        sel = regular & (np.abs(rows_d[i]) > 1e-12)
        assert sel.any() and np.abs(d2[i][sel] / rows_d[i][sel] - 2.0).max() < 1e-9
    # the next set_geometry goes back to the values recomputed from the support points
    ctx.set_geometry(m.xyz)
    ctx.assemble()
    assert np.array_equal(ctx.get_rows(1), rows_d) and np.array_equal(ctx.get_rows(0), rows_n)
    ctx.close()
