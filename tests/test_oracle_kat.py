"""Known-answer tests that pin the CPU oracle from first principles.

The reference (mathLab/WaveBEM) has no test for this path and cannot be built here, so these
KATs are what anchors the oracle: quadrature identities, flat-panel integrals, solid angles of
closed surfaces, manufactured harmonic fields, and linear-algebra cross checks with numpy.
"""
import numpy as np
import pytest

from wavebem_b200 import meshgen
from wavebem_b200.constraints import compute_constraints


def test_gauss_legendre_exactness(orc):
    for n in (1, 2, 4, 5, 8):
        x, w = orc.gauss01(n)
        assert abs(w.sum() - 1.0) < 1e-15
        for k in range(2 * n):  # exact for degree <= 2n-1
            assert abs((w * x ** k).sum() - 1.0 / (k + 1)) < 2e-15
    # published 4-point nodes on [-1,1]
    x, w = orc.gauss01(4)
    ref = np.array([-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526])
    assert np.abs((2 * x - 1) - ref).max() < 1e-15
    assert np.abs(2 * w - np.array([0.3478548451374538, 0.6521451548625461, 0.6521451548625461, 0.3478548451374538])).max() < 1e-15


def test_qgauss2_order_x_fastest(orc):
    uv, w = orc.qgauss2(4)
    x, w1 = orc.gauss01(4)
    assert np.allclose(uv[:4, 0], x) and np.allclose(uv[:4, 1], x[0])
    assert np.allclose(uv[4, :], [x[0], x[1]])
    assert abs(w.sum() - 1) < 1e-15


def test_qgauss_one_over_r_identities(orc):
    """SURVEY 8c KAT (1): sum w = 1 (area) to 2.3e-7 and sum w/R = 2 ln(1+sqrt 2) to 4.6e-8 at n=5;
    both to <= 5e-15 at n=12."""
    exact = 2 * np.log(1 + np.sqrt(2))
    corners = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=float)
    for v in range(4):
        uv, w = orc.qgauss_one_over_r(5, v)
        assert len(w) == 50
        assert (uv >= -1e-15).all() and (uv <= 1 + 1e-15).all()
        r = np.linalg.norm(uv - corners[v], axis=1)
        assert abs(w.sum() - 1) < 2.4e-7
        assert abs((w / r).sum() - exact) < 4.7e-8
        uv, w = orc.qgauss_one_over_r(12, v)
        r = np.linalg.norm(uv - corners[v], axis=1)
        assert abs(w.sum() - 1) < 5e-15
        assert abs((w / r).sum() - exact) < 5e-15


def test_fe_values_flat_and_flags(orc):
    X = np.array([[0, 0, 0], [2, 0, 0], [0, 3, 0], [2, 3, 0]], dtype=float)
    uv, w = orc.qgauss2(4)
    qp, nr, jw, sh = orc.fe_values(X, 1, uv, w)
    assert np.allclose(nr, [0, 0, 1]) and abs(jw.sum() - 6) < 1e-14
    assert np.allclose(qp[:, 0], 2 * uv[:, 0]) and np.allclose(qp[:, 1], 3 * uv[:, 1])
    assert np.allclose(sh.sum(axis=0), 1)
    _, nr2, _, _ = orc.fe_values(X, 0, uv, w)
    assert np.allclose(nr2, [0, 0, -1])


def test_flat_panel_corner(orc):
    """KAT (2): unit square panel, node at a corner: single-layer row sum = 2 ln(1+sqrt2)/(4 pi),
    double layer = 0."""
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], dtype=float)
    cells = np.array([[0, 1, 2, 3]], dtype=np.uint32)
    ptr = np.arange(5, dtype=np.uint32)
    idx = np.arange(4, dtype=np.uint32)
    nm, dm = orc.assemble_rows(xyz, cells, np.ones(1, np.uint8), ptr, idx)
    exact = 2 * np.log(1 + np.sqrt(2)) / (4 * np.pi)
    assert np.abs(dm.sum(axis=1) - exact).max() < 1e-8
    assert np.abs(nm).max() < 1e-16


@pytest.mark.parametrize("n", [3, 6])
def test_cube_solid_angles(orc, n):
    """KAT (3): alpha = -sum_j N_ij -> 1/2 (face), 1/4 (edge), 1/8 (corner)."""
    m = meshgen.cube(n)
    nm, _ = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    alpha = orc.compute_alpha(nm)
    sz = np.diff(m.dn_ptr)
    assert np.abs(alpha[sz == 1] - 0.5).max() < 1e-5
    assert np.abs(alpha[sz == 2] - 0.25).max() < 1e-5
    assert np.abs(alpha[sz == 3] - 0.125).max() < 1e-8


def test_renumbering_and_direction_flags_are_equivalent(orc):
    m0 = meshgen.cube(3)
    m1 = meshgen.cube(3, flip_every=2)
    n0, d0 = orc.assemble_rows(m0.xyz, m0.cells, m0.dir_flag, m0.dn_ptr, m0.dn_idx)
    n1, d1 = orc.assemble_rows(m1.xyz, m1.cells, m1.dir_flag, m1.dn_ptr, m1.dn_idx)
    assert np.abs(n0 - n1).max() < 1e-14 and np.abs(d0 - d1).max() < 1e-14


def test_sphere_alpha_and_harmonic_residual(orc):
    """KAT (3)+(4): sphere alpha -> 1/2; phi = 1/|x - x0| (x0 outside) satisfies
    alpha phi + N phi - D dphi/dn = 0 up to O(h^2)."""
    errs = []
    for n in (4, 8):
        m = meshgen.sphere(n)
        nm, dm = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
        alpha = orc.compute_alpha(nm)
        x0 = np.array([2.5, 0.3, -0.4])
        d = m.xyz - x0
        r = np.linalg.norm(d, axis=1)
        phi = 1 / r
        normal = m.xyz / np.linalg.norm(m.xyz, axis=1, keepdims=True)
        dphi = -(d * normal).sum(axis=1) / r ** 3
        res = alpha * phi + nm @ phi - dm @ dphi
        errs.append(np.abs(res).max())
        assert np.abs(alpha - 0.5).max() < 0.65 / n  # faceted sphere: O(h) solid-angle defect
    assert errs[1] < errs[0] / 2.5 and errs[1] < 5e-3


def test_cube_linear_field_exact_geometry(orc):
    """phi = x on the cube (flat faces: geometry exact, phi in the Q1 space)."""
    m = meshgen.cube(6)
    nm, dm = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    alpha = orc.compute_alpha(nm)
    nn = meshgen.cell_normals_at_nodes(m)
    res = alpha * m.xyz[:, 0] + nm @ m.xyz[:, 0] - dm @ nn[:, 0]
    assert np.abs(res).max() < 2e-3


def test_long_double_variant_agrees(orc):
    m = meshgen.wigley_tank(nxm=8, nt=4, nxu=3, nxd=4, nz=3, nzh=3)
    n0, d0 = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    n1, d1 = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, long_double=True)
    from conftest import rel_err_rowscaled
    # single layer: positive integrand, pure rounding
    assert rel_err_rowscaled(d0, d1) < 1e-13
    # double layer: (R.n)/r^3 between coplanar neighbours is a cancellation to ~1e-16 * |R|/r^3,
    # an ABSOLUTE noise floor (~1e-16) that rows with small entries (flat free surface, ~1e-4)
    # see as ~1e-12 relative.  It is a property of the reference arithmetic itself.
    assert rel_err_rowscaled(n0, n1) < 2e-11
    assert np.abs(n0 - n1).max() < 2e-13  # worst: singular pairs on flat patches (exact value 0)


def test_row_slab_equals_full(orc):
    m = meshgen.cube(4)
    n0, d0 = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, nthreads=1)
    n1, d1 = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx, 37, 91, nthreads=3)
    assert np.array_equal(n0[37:91], n1) and np.array_equal(d0[37:91], d1)


def _tank_problem(orc, **kw):
    m = meshgen.wigley_tank(nxm=10, nt=5, nxu=4, nxd=5, nz=3, nzh=3, **kw)
    bc = meshgen.towing_tank_bc(m)
    nn = meshgen.cell_normals_at_nodes(m)
    cl = compute_constraints(m.dn_ptr, m.dn_idx, m.surface_nodes, bc, nodes_normals=nn)
    con = orc.Constraints(m.n_nodes, cl.lines, cl.ptr, cl.col, cl.val, cl.inhom)
    nm, dm = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    return m, bc, con, nm, dm


def test_operator_algebra_matches_numpy(orc):
    """vmult / compute_rhs / constrained rows against direct numpy formulas (Appendix A.5)."""
    m, bc, con, nm, dm = _tank_problem(orc)
    s, o = m.surface_nodes, m.other_nodes
    alpha = orc.compute_alpha(nm)
    assert np.allclose(alpha, -nm.sum(axis=1), rtol=0, atol=1e-13)
    x = np.sin(0.37 * np.arange(m.n_nodes))
    y = orc.vmult(nm, dm, alpha, s, o, x)
    assert np.allclose(y, -dm @ (s * x) + nm @ (o * x) + alpha * o * x, rtol=0, atol=1e-12)
    b = orc.compute_rhs(nm, dm, alpha, s, o, x)
    assert np.allclose(b, -(nm @ (s * x) + alpha * s * x) + dm @ (o * x), rtol=0, atol=1e-12)
    yc = orc.constrained_vmult(nm, dm, alpha, s, o, con, x)
    free = con.line_of < 0
    assert np.array_equal(yc[free], y[free])
    for k, i in enumerate(con.lines):
        e = slice(con.ptr[k], con.ptr[k + 1])
        assert abs(yc[i] - (x[i] - (con.val[e] * x[con.col[e]]).sum())) < 1e-14
    r = orc.distribute_rhs(con, b)
    assert np.array_equal(r[con.lines], con.inhom) and np.array_equal(r[free], b[free])


def test_pure_neumann_shift(orc):
    m, bc, con, nm, dm = _tank_problem(orc)
    z = np.zeros(m.n_nodes)
    one = np.ones(m.n_nodes)
    alpha = orc.compute_alpha(nm)
    x = np.cos(0.2 * np.arange(m.n_nodes))
    y = orc.vmult(nm, dm, alpha, z, one, x)
    y0 = nm @ x + alpha * x
    assert np.allclose(y, y0 - np.linalg.norm(y0), rtol=0, atol=1e-11)


def test_band_preconditioner_matches_numpy(orc):
    m, bc, con, nm, dm = _tank_problem(orc)
    s = m.surface_nodes
    alpha = orc.compute_alpha(nm)
    n = m.n_nodes
    B = orc.band_system_dense(nm, dm, alpha, s, con, band=100)
    # literal restatement of bem_problem.cc:1126-1145
    ref = np.zeros((n, n))
    for i in range(n):
        if con.line_of[i] >= 0:
            ref[i, i] = 1
        for j in range(max(i - 50, 0), min(i + 50, n)):
            if con.line_of[j] < 0:
                if s[i] == 0:
                    ref[j, i] = nm[j, i] + (alpha[i] if i == j else 0.0)
                else:
                    ref[j, i] = -dm[j, i]
    assert np.array_equal(B, ref)
    v = np.sin(np.arange(n) * 0.11)
    z = orc.precond_apply(nm, dm, alpha, s, con, v, band=100)
    assert np.allclose(ref @ z, v, rtol=0, atol=1e-11)


def test_gmres_solution_matches_direct_solve(orc):
    """KAT (6): solve_system vs a dense LU of the same constrained operator."""
    m, bc, con, nm, dm = _tank_problem(orc)
    s, o = m.surface_nodes, m.other_nodes
    n = m.n_nodes
    r = orc.solve_system(nm, dm, s, o, bc, con, np.zeros(n), np.zeros(n), tol=1e-12, max_steps=300)
    assert r["converged"]
    alpha = r["alpha"]
    A = nm * o[None, :] + np.diag(alpha * o) - dm * s[None, :]
    for k, i in enumerate(con.lines):
        A[i, :] = 0
        A[i, i] = 1
        e = slice(con.ptr[k], con.ptr[k + 1])
        A[i, con.col[e]] -= con.val[e]
    x = np.linalg.solve(A, r["rhs"])
    assert np.linalg.norm(r["sol"] - x) / np.linalg.norm(x) < 1e-9
    assert np.array_equal(r["phi"][s == 0], r["sol"][s == 0])
    assert np.array_equal(r["dphi_dn"][s == 1], r["sol"][s == 1])
    # unpreconditioned GMRES reaches the same solution
    r2 = orc.solve_system(nm, dm, s, o, bc, con, np.zeros(n), np.zeros(n), tol=1e-11, max_steps=2000,
                          use_precond=False)
    assert r2["converged"] and np.linalg.norm(r2["sol"] - x) / np.linalg.norm(x) < 1e-8
    # residual() of the converged pair vanishes
    phi = np.where(s == 1, bc, r["phi"])
    dphi = np.where(s == 1, r["dphi_dn"], bc)
    res = orc.residual(nm, dm, s, o, con, phi, dphi)
    assert np.abs(res).max() < 1e-10


def test_gmres_no_convergence_reported(orc):
    m, bc, con, nm, dm = _tank_problem(orc)
    n = m.n_nodes
    r = orc.solve_system(nm, dm, m.surface_nodes, m.other_nodes, bc, con, np.zeros(n), np.zeros(n),
                         tol=1e-30, max_steps=7)
    assert not r["converged"] and r["iters"] == 7


def test_mixed_problem_recovers_trace(orc):
    """KAT (4): cube, top face Dirichlet, rest Neumann, exact field phi = 1/|x-x0|: the
    solve recovers the missing trace to discretisation accuracy."""
    errs = []
    for n in (4, 8):
        m = meshgen.cube(n)
        x0 = np.array([1.9, 1.4, 2.2])
        d = m.xyz - x0
        r = np.linalg.norm(d, axis=1)
        phi_ex = 1 / r
        nn = meshgen.cell_normals_at_nodes(m)
        dphi_ex = -(d * nn).sum(axis=1) / r ** 3
        top = m.node_patch == m.patch_names.index("z1")
        s = top.astype(float)
        o = 1 - s
        bc = np.where(top, phi_ex, dphi_ex)
        cl = compute_constraints(m.dn_ptr, m.dn_idx, s, bc, nodes_normals=nn)
        con = orc.Constraints(m.n_nodes, cl.lines, cl.ptr, cl.col, cl.val, cl.inhom)
        nm, dm = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
        out = orc.solve_system(nm, dm, s, o, bc, con, np.zeros(m.n_nodes), np.zeros(m.n_nodes), tol=1e-12,
                               max_steps=400)
        assert out["converged"]
        errs.append(np.abs(out["phi"][~top] - phi_ex[~top]).max() / np.abs(phi_ex).max())
    assert errs[1] < errs[0] and errs[1] < 2e-2


def test_l2_projected_normals_and_surface_gradients(orc):
    """compute_normals (computational_domain.cc:1525-1620) and compute_surface_gradients
    (bem_problem.cc:1153-1293), first principles: flat faces give exact answers, on a sphere
    the projected normal converges to the radial direction."""
    from wavebem_b200 import meshgen
    m = meshgen.cube(4, renumber="random", seed=3, flip_every=3)
    nrm = orc.compute_normals(m.xyz, m.cells, m.dir_flag)
    axis = {"z0": (0, 0, -1), "z1": (0, 0, 1), "y0": (0, -1, 0), "y1": (0, 1, 0), "x0": (-1, 0, 0), "x1": (1, 0, 0)}
    exact_n = np.array([axis[m.patch_names[p]] for p in m.node_patch], dtype=float)
    assert np.abs(nrm - exact_n).max() < 1e-13            # patch-wise dofs: exact outward face normals
    grad = np.array([1.0, 2.0, -0.5])
    g = orc.compute_surface_gradients(m.xyz, m.cells, m.dir_flag, m.xyz @ grad, np.ones(m.n_nodes))
    assert np.abs(g - (grad - (exact_n @ grad)[:, None] * exact_n)).max() < 1e-13
    # only the Dirichlet part of tmp_rhs enters (phi = tmp_rhs o surface_nodes, :1157-1158)
    s = (m.node_patch == 1).astype(float)
    g1 = orc.compute_surface_gradients(m.xyz, m.cells, m.dir_flag, m.xyz @ grad, s)
    assert np.abs(g1[m.node_patch == 1] - g[m.node_patch == 1]).max() < 1e-13
    assert np.abs(g1[m.node_patch != 1]).max() == 0.0
    errs = []
    for n in (6, 12):
        sp = meshgen.sphere(n, radius=0.7, center=(0.1, -0.2, 0.3))
        nr = orc.compute_normals(sp.xyz, sp.cells, sp.dir_flag)
        rad = (sp.xyz - np.array([0.1, -0.2, 0.3])) / 0.7
        assert np.abs(np.linalg.norm(nr, axis=1) - 1).max() < 1e-14
        errs.append(np.sqrt(np.mean(np.sum((nr - rad) ** 2, axis=1))))
    assert errs[1] < 0.6 * errs[0] and errs[1] < 0.03


def test_compute_constraints_sharp_dirichlet_edges_recover_the_normal_derivative(orc):
    """compute_constraints on an edge between two Dirichlet faces (bem_problem.cc:1054-1075): the
    imposed normal derivatives follow from the two surface gradients.  For a linear potential on a
    cube (orthogonal faces) the formula must return grad(phi).n exactly on both sides; flat
    Dirichlet-Dirichlet, Dirichlet-Neumann and Neumann-Neumann double nodes give the other three
    kinds of lines."""
    from wavebem_b200 import meshgen
    from wavebem_b200.constraints import compute_constraints
    m = meshgen.cube(3)
    grad = np.array([1.0, 2.0, -0.5])
    phi = m.xyz @ grad
    s = (m.node_patch <= 2).astype(float)          # faces z0, z1, y0 Dirichlet; y1, x0, x1 Neumann
    nrm = orc.compute_normals(m.xyz, m.cells, m.dir_flag)
    grd = orc.compute_surface_gradients(m.xyz, m.cells, m.dir_flag, phi, s)
    cl = compute_constraints(m.dn_ptr, m.dn_idx, s, phi, nodes_normals=nrm, node_surface_gradients=grd)
    kinds = {"sharp": 0, "dirichlet_neumann": 0, "neumann_neumann": 0}
    for k, line in enumerate(cl.lines):
        ent = list(zip(cl.col[cl.ptr[k]:cl.ptr[k + 1]], cl.val[cl.ptr[k]:cl.ptr[k + 1]]))
        doubles = set(int(j) for j in m.double_nodes_set(line))
        if s[line] == 1 and not ent:
            # a Dirichlet dof on a sharp Dirichlet-Dirichlet edge: its unknown dphi/dn is imposed
            others = [j for j in doubles if j != line and s[j] == 1]
            assert others and abs(cl.inhom[k] - grad @ nrm[line]) < 1e-12
            kinds["sharp"] += 1
        elif s[line] == 0 and not ent:
            # a Neumann double of a Dirichlet node: its potential is the imposed one
            first = min(j for j in doubles if s[j] == 1)
            assert cl.inhom[k] == phi[first]
            kinds["dirichlet_neumann"] += 1
        else:
            assert len(ent) == 1 and ent[0][1] == 1.0 and int(ent[0][0]) in doubles and cl.inhom[k] == 0.0
            assert s[line] == 0 and s[int(ent[0][0])] == 0
            kinds["neumann_neumann"] += 1
    assert all(v > 0 for v in kinds.values()), kinds
    # every double node except the first of its set is constrained (first Dirichlet ones on sharp edges too)
    n_sets = len({tuple(m.double_nodes_set(i)) for i in range(m.n_nodes) if len(m.double_nodes_set(i)) > 1})
    n_in_sets = sum(1 for i in range(m.n_nodes) if len(m.double_nodes_set(i)) > 1)
    assert n_in_sets - n_sets <= cl.n_lines <= n_in_sets


def test_singular_rule_on_a_warped_cell_against_adaptive_quadrature(orc):
    """The 50-point QGaussOneOverR rule at each of the four vertices of a NON-planar bilinear cell
    (the reference's singular branch, bem_problem.cc:261-525), against an adaptive polar
    integration centred at the vertex (integrand * rho is smooth).  Pins the vertex <-> rule
    orientation and the Q1 mapping on a cell without any symmetry."""
    from scipy import integrate
    X = np.array([[0.0, 0.0, 0.0], [1.3, 0.1, 0.2], [-0.1, 0.9, -0.15], [1.1, 1.2, 0.35]])
    cells = np.array([[0, 1, 2, 3]], dtype=np.uint32)
    ptr, idx = np.arange(5, dtype=np.uint32), np.arange(4, dtype=np.uint32)
    on, od = orc.assemble_rows(X, cells, np.ones(1, np.uint8), ptr, idx)   # every (node, cell) pair is singular

    def shape(u, v):
        return np.array([(1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v])

    def geom(u, v):
        tu = (1 - v) * (X[1] - X[0]) + v * (X[3] - X[2])
        tv = (1 - u) * (X[2] - X[0]) + u * (X[3] - X[1])
        cr = np.cross(tu, tv)
        return shape(u, v) @ X, cr

    ref_d, ref_n = np.zeros((4, 4)), np.zeros((4, 4))
    for k, (u0, v0) in enumerate([(0, 0), (1, 0), (0, 1), (1, 1)]):
        su, sv = (1 if u0 == 0 else -1), (1 if v0 == 0 else -1)

        def integrand(rho, th, j, which):
            u, v = u0 + su * rho * np.cos(th), v0 + sv * rho * np.sin(th)
            y, cr = geom(u, v)
            R = y - X[k]
            r = np.linalg.norm(R)
            f = shape(u, v)[j]
            if which == 0:
                return f * np.linalg.norm(cr) / (4 * np.pi * r) * rho
            return f * (R @ cr) / (-4 * np.pi * r ** 3) * rho

        for j in range(4):
            for which, out in ((0, ref_d), (1, ref_n)):
                val = 0.0
                for a, b, rmax in ((0, np.pi / 4, lambda th: 1 / np.cos(th)), (np.pi / 4, np.pi / 2, lambda th: 1 / np.sin(th))):
                    v, _ = integrate.dblquad(lambda rho, th: integrand(rho, th, j, which), a, b, 1e-14, rmax,
                                             epsabs=1e-12, epsrel=1e-10)
                    val += v
                out[k, j] = val
    assert np.abs(od - ref_d).max() < 2e-6 * np.abs(ref_d).max()
    assert np.abs(on - ref_n).max() < 2e-5 * max(np.abs(ref_n).max(), 1e-3)


def test_regular_rule_on_a_warped_cell_against_adaptive_quadrature(orc):
    """Regular (node, cell) pairs, bem_problem.cc:241-260: the 4x4 Gauss integrals of both kernels
    over a non-planar bilinear cell, seen from nodes a few cell sizes away, against scipy's adaptive
    quadrature of the same integrands (normal orientation, JxW and shape functions included)."""
    from scipy import integrate
    X = np.array([[0.0, 0.0, 0.0], [1.3, 0.1, 0.2], [-0.1, 0.9, -0.15], [1.1, 1.2, 0.35],
                  [3.0, 2.0, 1.5], [3.5, 2.0, 1.5], [3.0, 2.5, 1.5], [3.5, 2.5, 1.6]])
    cells = np.array([[0, 1, 2, 3], [4, 5, 6, 7]], dtype=np.uint32)
    ptr, idx = np.arange(9, dtype=np.uint32), np.arange(8, dtype=np.uint32)
    for flag in (1, 0):
        on, od = orc.assemble_rows(X, cells, np.array([flag, 1], np.uint8), ptr, idx)

        def f(v, u, j, which, i):
            sh = np.array([(1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v])
            tu = (1 - v) * (X[1] - X[0]) + v * (X[3] - X[2])
            tv = (1 - u) * (X[2] - X[0]) + u * (X[3] - X[1])
            cr = np.cross(tu, tv) * (1.0 if flag else -1.0)       # cell->direction_flag()
            R = sh @ X[:4] - X[i]
            r = np.linalg.norm(R)
            return sh[j] * (np.linalg.norm(cr) / (4 * np.pi * r) if which == 0 else (R @ cr) / (-4 * np.pi * r ** 3))

        for i in (4, 7):
            for j in range(4):
                d, _ = integrate.dblquad(f, 0, 1, 0, 1, args=(j, 0, i), epsabs=1e-14, epsrel=1e-12)
                nn, _ = integrate.dblquad(f, 0, 1, 0, 1, args=(j, 1, i), epsabs=1e-14, epsrel=1e-12)
                assert abs(od[i, j] / d - 1) < 5e-7 and abs(on[i, j] / nn - 1) < 5e-6


# ---------------------------------------------------------------------------------------------
# post-processing integrals (oracle/postproc.py, the checkers of csrc/postproc.cu)
# ---------------------------------------------------------------------------------------------
def test_internal_velocities_are_the_gradient_of_the_representation_formula():
    """compute_internal_velocities (free_surface.cc:10426-10537) differentiates G and dG/dn with Sacado;
    the restatement writes the gradients out: check them against central differences of the potential."""
    from oracle import postproc
    from wavebem_b200 import meshgen
    m = meshgen.cube(3)
    rng = np.random.default_rng(3)
    phi, dphi = rng.normal(size=m.n_nodes), rng.normal(size=m.n_nodes)
    pts = np.array([[0.6, 0.45, 0.7], [0.2, 0.75, 0.4]])      # inside the unit cube
    v = postproc.internal_velocities(m, phi, dphi, pts)
    h = 1e-5
    fd = np.zeros_like(v)
    for d in range(3):
        e = np.zeros(3)
        e[d] = h
        fd[:, d] = (postproc.potential_at(m, phi, dphi, pts + e) - postproc.potential_at(m, phi, dphi, pts - e)) / (2 * h)
    assert np.abs(v - fd).max() < 1e-8 * max(1.0, np.abs(v).max())


def test_internal_velocity_of_a_linear_potential_inside_a_cube():
    """phi = x is harmonic: Green's representation returns phi and grad phi = (1,0,0) at interior points
    (outward normals; exact up to the quadrature of the 1/r kernels over the flat faces)."""
    from oracle import postproc
    from wavebem_b200 import meshgen
    m = meshgen.cube(8)
    nn = meshgen.cell_normals_at_nodes(m)     # face normals (double nodes on the edges keep them per face)
    phi = m.xyz[:, 0].copy()
    dphi = nn[:, 0].copy()
    pts = np.array([[0.0, 0.0, 0.0], [0.1, -0.15, 0.05]]) + m.xyz.mean(axis=0)
    v = postproc.internal_velocities(m, phi, dphi, pts)
    p = postproc.potential_at(m, phi, dphi, pts)
    assert np.abs(v - np.array([1.0, 0.0, 0.0])).max() < 2e-3
    assert np.abs(p - pts[:, 0]).max() < 2e-3


def test_sparse_projection_restatement_matches_the_c_oracle_and_condenses_hanging_nodes(orc):
    """oracle/projections.py (scipy.sparse, with hanging-node condensation) against orc_l2_projection on a
    conforming mesh; on a locally refined FLAT patch the condensed projection of a linear field's gradient
    is exact, and the hanging nodes end up on the mean of their masters (distribute())."""
    from oracle import projections
    from wavebem_b200 import meshgen
    m = meshgen.wigley_tank(nxm=10, nt=5, nxu=4, nxd=5, nz=3, nzh=3, wave_amp=0.02)
    f = np.cos(1.3 * m.xyz[:, 0]) + m.xyz[:, 1] * m.xyz[:, 2]
    assert np.abs(projections.l2_projection(0, m.xyz, m.cells, m.dir_flag) -
                  orc.compute_normals(m.xyz, m.cells, m.dir_flag)).max() < 1e-11
    s = m.surface_nodes
    assert np.abs(projections.l2_projection(1, m.xyz, m.cells, m.dir_flag, f * s) -
                  orc.compute_surface_gradients(m.xyz, m.cells, m.dir_flag, f, s)).max() < 1e-10
    # refine part of the (flat) bottom patch: hanging nodes on the interface
    bottom = m.cell_patch == m.patch_names.index("bottom")
    cx = m.xyz[m.cells.astype(int)].mean(axis=1)[:, 0]
    r, hang = meshgen.refine_cells(m, bottom & (cx > np.median(cx[bottom])))
    assert len(hang) > 0
    lin = 0.7 * r.xyz[:, 0] - 0.2 * r.xyz[:, 1]            # linear in the plane of the bottom
    g = projections.l2_projection(1, r.xyz, r.cells, r.dir_flag, lin, hanging=hang)
    on_bottom = r.node_patch == r.patch_names.index("bottom")
    assert np.abs(g[on_bottom] - np.array([0.7, -0.2, 0.0])).max() < 1e-10
    nrm = projections.l2_projection(0, r.xyz, r.cells, r.dir_flag, hanging=hang)
    assert np.abs(nrm[on_bottom] - nrm[on_bottom][0]).max() < 1e-12
    # a curved patch (the hull): the constrained solution differs from the unconstrained one, and the
    # hanging values are the mean of their masters before normalisation
    hull = np.isin(m.cell_patch, m.meta["hull_patches"])
    r2, hang2 = meshgen.refine_cells(m, hull & (cx > 0.2))
    f2 = np.sin(2.0 * r2.xyz[:, 0]) + r2.xyz[:, 2]
    gc = projections.l2_projection(1, r2.xyz, r2.cells, r2.dir_flag, f2, hanging=hang2)
    gu = projections.l2_projection(1, r2.xyz, r2.cells, r2.dir_flag, f2)
    h0, ent = hang2[0]
    assert np.abs(gc[h0] - sum(w * gc[mm] for mm, w in ent)).max() < 1e-12
    assert np.abs(gc - gu).max() > 1e-6


def test_line_records_reproduce_the_pointwise_integrand_on_warped_cells(orc):
    """The identities behind k_assemble_rows' line records (assemble.cu): on a bilinear cell, along a Gauss line
    v = v_j, y(u) = A + u B and d_u y x d_v y = P + u Q with B.P = B.Q = 0, hence for D = A - x
        |y - x|^2 = D.D + u (2 B.D) + u^2 B.B        and        (y - x).n JxW = w_u w_v (D.P + u D.Q) sign.
    A numpy restatement of the records (k_cell_geometry) against the oracle's FEValues (quadrature points,
    normals, JxW of deal.II's MappingQ1) on warped, non-planar cells of both orientations, for collocation points
    near and far: integrands and the four shape-function sums of both kernels agree to rounding."""
    rng = np.random.default_rng(5)
    x1, w1 = orc.gauss01(4)
    uv, w = orc.qgauss2(4)            # point q = 4 j + i  <->  (u_i, v_j)
    assert np.allclose(uv[:, 0], np.tile(x1, 4)) and np.allclose(uv[:, 1], np.repeat(x1, 4))
    for trial in range(20):
        X = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0]], float) * rng.uniform(0.2, 3.0, 3)
        X += rng.normal(0, 0.15, X.shape) + rng.uniform(-5, 5, 3)   # warped (non-planar), anywhere in space
        flag = trial % 2
        sgn = 1.0 if flag else -1.0
        qp, nr, jw, sh = orc.fe_values(X, flag, uv, w)
        cc = X[2] - X[0]
        dd = (X[3] - X[2]) - (X[1] - X[0])
        for x in (X.mean(0) + rng.normal(0, 0.7, 3), X[0] + rng.normal(0, 30.0, 3), X[3] + 0.3 * (X[3] - X[0])):
            r_ref = qp - x
            r2_ref = (r_ref ** 2).sum(1)
            rn_ref = (r_ref * nr).sum(1) * jw
            sums_n = np.zeros(4)
            sums_d = np.zeros(4)
            for j in range(4):
                v = x1[j]
                A = X[0] + v * cc
                B = (X[1] - X[0]) + v * dd
                P = np.cross(B, cc) * sgn * w1[j]
                Q = np.cross(B, dd) * sgn * w1[j]
                assert abs(B @ np.cross(B, cc)) < 1e-12 and abs(B @ np.cross(B, dd)) < 1e-12
                D = A - x
                c0, c1, c2, e0, e1 = D @ D, 2 * (B @ D), B @ B, D @ P, D @ Q
                for i in range(4):
                    u = x1[i]
                    q = 4 * j + i
                    r2 = (c2 * u + c1) * u + c0
                    rn = w1[i] * (e1 * u + e0)
                    assert abs(r2 - r2_ref[q]) <= 1e-13 * max(r2_ref[q], (np.abs(X - x) ** 2).sum(1).max())
                    assert abs(rn - rn_ref[q]) <= 1e-13 * np.abs(r_ref[q]).max() * jw[q] + 1e-15
                    ri = 1.0 / np.sqrt(r2)
                    kn, kd = rn * ri ** 3 / (-4 * np.pi), jw[q] * ri / (4 * np.pi)
                    phi = np.array([(1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v])
                    sums_n += kn * phi
                    sums_d += kd * phi
            # the reference's pointwise sums (bem_problem.cc:241-260 with LaplaceKernel, laplace_kernel.h:47-57)
            r = np.sqrt(r2_ref)
            ref_n = ((r_ref * nr).sum(1) / (-4 * np.pi * r ** 3) * jw) @ sh.T
            ref_d = (1.0 / (4 * np.pi * r) * jw) @ sh.T
            assert np.abs(sums_n - ref_n).max() <= 1e-13 * max(np.abs(ref_n).max(), 1e-300) + 1e-16
            assert np.abs(sums_d - ref_d).max() <= 1e-13 * np.abs(ref_d).max()
