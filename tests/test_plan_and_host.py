"""Host logic that runs without a GPU: the assembly tiling plan, the mesh generators'
double-node sets, and the compute_constraints restatement."""
import ctypes as C

import numpy as np
import pytest

from wavebem_b200 import meshgen
from wavebem_b200.constraints import compute_constraints


def _plan_check(wb, mesh, w=48, mc=64):
    st = np.zeros(9)
    rc = wb.lib().wbem_plan_check(C.c_uint32(mesh.n_nodes), C.c_uint32(mesh.n_cells),
                                  mesh.cells.ctypes.data_as(C.c_void_p), C.c_uint32(w), C.c_uint32(mc),
                                  st.ctypes.data_as(C.c_void_p))
    return rc, st


@pytest.mark.parametrize("make", [
    lambda: meshgen.cube(1), lambda: meshgen.cube(5), lambda: meshgen.cube(5, renumber="random", seed=3),
    lambda: meshgen.sphere(7, flip_every=3), lambda: meshgen.wigley_tank(),
    lambda: meshgen.wigley_tank(renumber="random", seed=11), lambda: meshgen.wigley_tank_for_nodes(20000),
])
def test_plan_invariants(wb, make):
    m = make()
    rc, st = _plan_check(wb, m)
    assert rc == 0, f"plan invariant {rc} violated"
    assert st[2] <= 64 and st[3] <= 48 and st[6] == m.n_nodes


def test_stream_kernel_item_order(wb):
    """Ticket order of the single-launch assembly kernel (assemble.cu, decode_ticket_order): a bijection onto
    (row tile, cluster), predecessors always drawn earlier, for every group size and ragged row counts; with
    G tiles per group an item and its predecessors are ~G x (clusters of a colour) tickets apart."""
    m = meshgen.wigley_tank_for_nodes(6000)
    cells = np.ascontiguousarray(m.cells, dtype=np.uint32)
    f = wb.lib().wbem_stream_order_check
    mean = {}
    for n_rows in (m.n_nodes, m.n_nodes // 3 + 5, 1, 129):
        for g in (1, 2, 3, 7, 1000):
            st = np.zeros(4)
            rc = f(C.c_uint32(m.n_nodes), C.c_uint32(m.n_cells), cells.ctypes.data_as(C.c_void_p), C.c_uint32(n_rows),
                   C.c_uint32(g), st.ctypes.data_as(C.c_void_p))
            assert rc == 0, (n_rows, g, rc)
            assert st[3] == (n_rows + 127) // 128 and st[1] >= 1
            if n_rows == m.n_nodes:
                mean[g] = st[2]
    assert mean[2] > 1.8 * mean[1] and mean[3] > 2.7 * mean[1]
    m2 = meshgen.cube(4, renumber="random", seed=3)
    st = np.zeros(4)
    rc = f(C.c_uint32(m2.n_nodes), C.c_uint32(m2.n_cells), np.ascontiguousarray(m2.cells, dtype=np.uint32).ctypes.data_as(C.c_void_p),
           C.c_uint32(m2.n_nodes), C.c_uint32(2), st.ctypes.data_as(C.c_void_p))
    assert rc in (0, 101)


def test_plan_small_tiles_and_degenerate_cells(wb):
    m = meshgen.cube(4)
    for w, mc in ((4, 1), (6, 2), (9, 4), (16, 64)):
        rc, st = _plan_check(wb, m, w, mc)
        assert rc == 0 and st[3] <= w and st[2] <= mc
    # a cell with a repeated dof (collapsed quad) and a dof no cell uses
    cells = np.array([[0, 1, 2, 2], [1, 3, 2, 4]], dtype=np.uint32)
    st = np.zeros(9)
    rc = wb.lib().wbem_plan_check(C.c_uint32(6), C.c_uint32(2), cells.ctypes.data_as(C.c_void_p),
                                  C.c_uint32(48), C.c_uint32(36), st.ctypes.data_as(C.c_void_p))
    assert rc == 0 and st[6] == 5
    bad = np.array([[0, 1, 2, 9]], dtype=np.uint32)
    rc = wb.lib().wbem_plan_check(C.c_uint32(4), C.c_uint32(1), bad.ctypes.data_as(C.c_void_p),
                                  C.c_uint32(48), C.c_uint32(36), None)
    assert rc != 0


def test_double_nodes_sets():
    m = meshgen.cube(3)
    sz = np.diff(m.dn_ptr)
    assert (sz == 3).sum() == 24 and (sz == 2).sum() == 12 * 2 * 2 and (sz == 1).sum() == 6 * 4
    for i in range(m.n_nodes):
        s = m.double_nodes_set(i)
        assert i in s and (np.diff(s) > 0).all()
        for j in s:  # symmetric
            assert i in m.double_nodes_set(j)
            assert np.linalg.norm(m.xyz[i] - m.xyz[j]) < 1e-8
    t = meshgen.wigley_tank()
    assert (np.diff(t.dn_ptr)[~t.node_on_patch_boundary] == 1).all()
    assert set(np.unique(t.surface_nodes + t.other_nodes)) == {1.0}


def test_tank_node_count_targets():
    for target in (4000, 20000):
        m = meshgen.wigley_tank_for_nodes(target)
        assert abs(m.n_nodes - target) / target < 0.02


def test_compute_constraints_branches():
    m = meshgen.cube(2)
    n = m.n_nodes
    nn = meshgen.cell_normals_at_nodes(m)
    top = m.node_patch == m.patch_names.index("z1")
    bc = np.arange(n, dtype=float)
    # all Neumann: every double is tied to the smallest member
    cl = compute_constraints(m.dn_ptr, m.dn_idx, np.zeros(n), bc)
    assert (np.diff(cl.lines.astype(int)) > 0).all()
    for k, d in enumerate(cl.lines):
        s = m.double_nodes_set(d)
        assert d != s[0] and list(cl.col[cl.ptr[k]:cl.ptr[k + 1]]) == [s[0]] and cl.inhom[k] == 0
    assert cl.n_lines == sum(len(m.double_nodes_set(i)) - 1 for i in range(n) if m.double_nodes_set(i)[0] == i)
    # top Dirichlet: Neumann doubles of a Dirichlet node take its boundary value
    cl = compute_constraints(m.dn_ptr, m.dn_idx, top.astype(float), bc, nodes_normals=nn)
    for k, d in enumerate(cl.lines):
        s = list(m.double_nodes_set(d))
        dir_members = [j for j in s if top[j]]
        if dir_members and not top[d]:
            assert cl.ptr[k] == cl.ptr[k + 1] and cl.inhom[k] == bc[dir_members[0]]
    # two Dirichlet faces meeting at an edge need the surface gradients
    two = top | (m.node_patch == m.patch_names.index("x1"))
    with pytest.raises(ValueError):
        compute_constraints(m.dn_ptr, m.dn_idx, two.astype(float), bc, nodes_normals=nn)
    g = np.zeros((n, 3))
    cl = compute_constraints(m.dn_ptr, m.dn_idx, two.astype(float), bc, nodes_normals=nn,
                             node_surface_gradients=g)
    assert cl.n_lines > 0
    # hanging-node lines pass through
    cl = compute_constraints(m.dn_ptr, m.dn_idx, np.zeros(n), bc, hanging=[(4, [(0, 0.5), (8, 0.5)])])
    k = list(cl.lines).index(4)
    assert list(cl.val[cl.ptr[k]:cl.ptr[k + 1]]) == [0.5, 0.5]


def _spai_pattern(wb, mesh):
    import ctypes as C
    nbr = np.empty((mesh.n_nodes, 32), dtype=np.uint32)
    st = (C.c_double * 4)()
    xyz = np.ascontiguousarray(mesh.xyz, dtype=np.float64)
    rc = wb.lib().wbem_spai_pattern_check(C.c_uint32(mesh.n_nodes), C.c_uint32(mesh.n_cells),
                                          mesh.cells.ctypes.data_as(C.c_void_p), mesh.dn_ptr.ctypes.data_as(C.c_void_p),
                                          mesh.dn_idx.ctypes.data_as(C.c_void_p), xyz.ctypes.data_as(C.c_void_p),
                                          nbr.ctypes.data_as(C.c_void_p), st)
    return rc, nbr, list(st)


def test_spai_pattern_invariants_and_locality(wb):
    """Sparsity pattern of the local-inverse preconditioner (spai.cu): every row sorted,
    duplicate-free, holds its own dof and its double nodes; on a mesh with more than K dofs the
    rows are full and made of geometrically near dofs."""
    from wavebem_b200 import meshgen
    for m in (meshgen.cube(1), meshgen.cube(4, renumber="random", seed=1), meshgen.wigley_tank()):
        rc, nbr, st = _spai_pattern(wb, m)
        assert rc == 0 and st[0] == 32
        n = m.n_nodes
        for i in range(n):
            row = nbr[i][nbr[i] != 0xFFFFFFFF]
            assert i in row and len(row) == min(32, n) or len(row) >= min(32, len(row))
            for d in m.double_nodes_set(i):
                assert d in row or len(row) == 32
        if n > 64:
            assert st[3] == 0                      # every row has K entries
            tree_d = np.linalg.norm(m.xyz[nbr.astype(np.int64)] - m.xyz[:, None, :], axis=2)
            # the chosen dofs are near: compare with the true 32nd nearest neighbour (rows differ
            # from the exact kNN only where patches face each other without sharing an edge)
            from scipy.spatial import cKDTree
            d32 = cKDTree(m.xyz).query(m.xyz, k=32)[0][:, -1]
            ratio = tree_d.max(axis=1) / d32
            assert np.median(ratio) < 1.2 and np.percentile(ratio, 99) < 3.0


def test_bench_reference_iteration_table():
    """bench.py --impl reference runs the GMRES iteration count the reference's band preconditioner
    needs at the benchmark size (table measured with precond_kind = 0)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.reference_band_iters(20073) == 83 and bench.reference_band_iters(57001) == 154
    assert bench.reference_band_iters(100) == bench.BAND_ITERS[0][1]
    assert bench.reference_band_iters(10 ** 6) == bench.BAND_ITERS[-1][1]
    xs = [bench.reference_band_iters(n) for n in range(4000, 81000, 1000)]
    assert all(b >= a for a, b in zip(xs, xs[1:]))
    peak, src = bench.measured_peaks()
    assert peak > 1000 and isinstance(src, str)


def test_spai_pattern_gives_a_better_preconditioner_than_the_band(wb, orc):
    """The ALGORITHM of spai.cu restated in numpy on the oracle's matrices (no GPU): with the
    library's sparsity pattern, rows m_i = e_i^T A[S,S]^-1 precondition GMRES better than the
    reference's band-100 LU on the same system (iterations of a plain numpy GMRES vs the oracle's)."""
    import scipy.linalg as sla
    from conftest import make_problem
    from wavebem_b200 import meshgen
    m = meshgen.wigley_tank(nxm=12, nt=6, nxu=4, nxd=6, nz=3, nzh=3)
    bc, nn, cl = make_problem(m)
    n = m.n_nodes
    on, od = orc.assemble_rows(m.xyz, m.cells, m.dir_flag, m.dn_ptr, m.dn_idx)
    alpha = orc.compute_alpha(on)
    s, o = m.surface_nodes, m.other_nodes
    A = on * o[None, :] - od * s[None, :]
    A[np.arange(n), np.arange(n)] += alpha * o
    b = orc.compute_rhs(on, od, alpha, s, o, bc)
    for k, line in enumerate(cl.lines):
        A[line, :] = 0.0
        A[line, line] = 1.0
        for c, v in zip(cl.col[cl.ptr[k]:cl.ptr[k + 1]], cl.val[cl.ptr[k]:cl.ptr[k + 1]]):
            A[line, c] -= v
        b[line] = cl.inhom[k]
    rc, nbr, st = _spai_pattern(wb, m)
    assert rc == 0
    M = np.zeros((n, n))
    for i in range(n):
        S = nbr[i][nbr[i] != 0xFFFFFFFF].astype(np.int64)
        M[i, S] = np.linalg.solve(A[np.ix_(S, S)].T, (S == i).astype(float))

    def gmres_iters(prec, tol=1e-10, restart=98, maxit=400):
        x, it = np.zeros(n), 0
        while it < maxit:
            r = prec(b - A @ x)
            beta = np.linalg.norm(r)
            V = np.zeros((restart + 1, n))
            H = np.zeros((restart + 1, restart))
            V[0] = r / beta
            for k in range(restart):
                w = prec(A @ V[k])
                for _ in range(2):
                    h = V[:k + 1] @ w
                    w -= h @ V[:k + 1]
                    H[:k + 1, k] += h
                H[k + 1, k] = np.linalg.norm(w)
                V[k + 1] = w / H[k + 1, k]
                it += 1
                e1 = np.zeros(k + 2)
                e1[0] = beta
                y, res, _, _ = np.linalg.lstsq(H[:k + 2, :k + 1], e1, rcond=None)
                if np.linalg.norm(H[:k + 2, :k + 1] @ y - e1) <= tol or it >= maxit:
                    return it
            x += y @ V[:restart]
        return it

    it_spai = gmres_iters(lambda v: M @ v)
    con = orc.Constraints(cl.n, cl.lines, cl.ptr, cl.col, cl.val, cl.inhom)
    ref = orc.solve_system(on, od, s, o, bc, con, np.zeros(n), np.zeros(n), tol=1e-10, max_steps=400)
    assert ref["converged"] and it_spai <= 0.6 * ref["iters"], (it_spai, ref["iters"])


def test_generate_double_nodes_set_matches_the_reference_loop(wb):
    """wbem_generate_double_nodes_set (grid search) against a literal numpy restatement of the
    reference's all-pairs loop (computational_domain.cc:279-303) and against the mesh generator's
    KD-tree version, including triple nodes, interior coincidences that must NOT be searched from
    (non-boundary dofs) but are found as doubles of boundary ones, and points within / beyond tol."""
    from wavebem_b200 import meshgen
    for m in (meshgen.cube(3, renumber="random", seed=2), meshgen.wigley_tank(nxm=10, nt=5, nxu=4, nxd=5, nz=3, nzh=3)):
        ptr, idx = wb.generate_double_nodes_set(m.xyz, m.node_on_patch_boundary)
        assert np.array_equal(ptr, m.dn_ptr) and np.array_equal(idx, m.dn_idx)
    rng = np.random.default_rng(4)
    x = rng.uniform(-3, 3, (400, 3))
    x[50] = x[10] + np.array([3e-9, -4e-9, 0.0])          # 5e-9 apart: inside tol
    x[51] = x[10] + np.array([0.0, 0.0, 9e-9])             # 9e-9: inside
    x[52] = x[10] + np.array([8e-9, 8e-9, 0.0])            # 1.13e-8: outside
    x[60] = x[20]                                          # exact duplicate of a non-boundary dof
    bnd = np.zeros(400, dtype=bool)
    bnd[[10, 50, 60]] = True
    ptr, idx = wb.generate_double_nodes_set(x, bnd, tol=1e-8)
    for i in range(400):
        got = idx[ptr[i]:ptr[i + 1]].tolist()
        if bnd[i]:
            want = sorted(set([i]) | set(np.nonzero(np.linalg.norm(x - x[i], axis=1) < 1e-8)[0].tolist()))
        else:
            want = [i]
        assert got == want, (i, got, want)
    assert idx[ptr[10]:ptr[11]].tolist() == [10, 50, 51] and idx[ptr[60]:ptr[61]].tolist() == [20, 60]
    assert idx[ptr[20]:ptr[21]].tolist() == [20]           # 20 is not a boundary dof: no search from it
    # boundary_dofs = None: every dof is tested
    ptr2, idx2 = wb.generate_double_nodes_set(x, None)
    assert idx2[ptr2[20]:ptr2[21]].tolist() == [20, 60]
