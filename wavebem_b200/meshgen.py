"""Synthetic multi-patch quad surface meshes in the flattened form the BEM path consumes.

The reference builds its meshes from IGES hulls through OpenCASCADE
(source/numerical_towing_tank.cc:290-1790), which is not available here, so benchmark and
test meshes are synthesised with the same ingredients:

* towing tank 12L x 4L x 2L around the hull (numerical_towing_tank.cc:1272-1274),
* analytic Wigley hull L=2.5, B=L/10, T=B/1.6 (source/boat_surface.cc:52-54, 93),
* surface split in patches whose edge nodes are duplicated ("double nodes",
  source/computational_domain.cc:258-307, tol 1e-8 at :271),
* free-surface nodes flagged Dirichlet (surface_nodes=1), everything else Neumann
  (other_nodes=1) (numerical_towing_tank.cc:1799-1807).

Cells list their 4 dofs in deal.II lexicographic vertex order; `dir_flag` is
cell->direction_flag() (normal = d_u x d_v, flipped when 0).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

WIGLEY_L = 2.5
WIGLEY_B = WIGLEY_L / 10.0
WIGLEY_T = WIGLEY_B / 1.6


@dataclass
class SurfaceMesh:
    xyz: np.ndarray            # (N,3) support_points
    cells: np.ndarray          # (C,4) uint32
    dir_flag: np.ndarray       # (C,) uint8
    cell_patch: np.ndarray     # (C,) int32
    node_patch: np.ndarray     # (N,) int32
    node_on_patch_boundary: np.ndarray  # (N,) bool
    patch_names: list
    dn_ptr: np.ndarray = None  # CSR of double_nodes_set (each set sorted, holds i itself)
    dn_idx: np.ndarray = None
    surface_nodes: np.ndarray = None   # (N,) 1.0 where phi is imposed (Dirichlet)
    other_nodes: np.ndarray = None     # (N,) 1.0 where dphi_dn is imposed (Neumann)
    meta: dict = field(default_factory=dict)

    @property
    def n_nodes(self):
        return self.xyz.shape[0]

    @property
    def n_cells(self):
        return self.cells.shape[0]

    def double_nodes_set(self, i):
        return self.dn_idx[self.dn_ptr[i]:self.dn_ptr[i + 1]]


# ----------------------------------------------------------------------------------------
# structured patches
# ----------------------------------------------------------------------------------------
def _patch(points_grid, outward_is_uxv=True):
    """points_grid: (nv+1, nu+1, 3), u fastest.  Returns xyz, cells, dir_flag, boundary mask."""
    nv1, nu1, _ = points_grid.shape
    xyz = points_grid.reshape(-1, 3)
    idx = np.arange(nu1 * nv1).reshape(nv1, nu1)
    c0 = idx[:-1, :-1].ravel()
    c1 = idx[:-1, 1:].ravel()
    c2 = idx[1:, :-1].ravel()
    c3 = idx[1:, 1:].ravel()
    cells = np.stack([c0, c1, c2, c3], axis=1)
    bnd = np.zeros((nv1, nu1), dtype=bool)
    bnd[0, :] = bnd[-1, :] = True
    bnd[:, 0] = bnd[:, -1] = True
    dirf = np.full(cells.shape[0], 1 if outward_is_uxv else 0, dtype=np.uint8)
    return xyz, cells, dirf, bnd.ravel(), (nu1 - 1, nv1 - 1)


def _morton(i, j):
    """Interleave the bits of two index arrays (z-order)."""
    code = np.zeros_like(i, dtype=np.int64)
    for b in range(20):
        code |= ((i >> b) & 1) << (2 * b)
        code |= ((j >> b) & 1) << (2 * b + 1)
    return code


def _finish(patches, names, renumber=None, seed=0, flip_every=0):
    """Concatenate patches, optionally renumber dofs, build double-node sets."""
    xyz_l, cells_l, dir_l, bnd_l, cp_l, np_l = [], [], [], [], [], []
    off = 0
    morton_l = []
    for p, (xyz, cells, dirf, bnd, (nu, nv)) in enumerate(patches):
        k = np.arange(cells.shape[0])
        morton_l.append(_morton(k % nu, k // nu))
        xyz_l.append(xyz)
        cells_l.append(cells + off)
        dir_l.append(dirf)
        bnd_l.append(bnd)
        cp_l.append(np.full(cells.shape[0], p, dtype=np.int32))
        np_l.append(np.full(xyz.shape[0], p, dtype=np.int32))
        off += xyz.shape[0]
    xyz = np.ascontiguousarray(np.concatenate(xyz_l), dtype=np.float64)
    cells = np.concatenate(cells_l).astype(np.int64)
    dirf = np.concatenate(dir_l)
    bnd = np.concatenate(bnd_l)
    cpatch = np.concatenate(cp_l)
    npatch = np.concatenate(np_l)
    n = xyz.shape[0]

    if flip_every:
        # reverse the local orientation of every k-th cell and clear its direction flag:
        # same geometry and normal, exercises the direction_flag path
        sel = np.arange(cells.shape[0]) % flip_every == 0
        cells[sel] = cells[sel][:, [1, 0, 3, 2]]
        dirf = dirf.copy()
        dirf[sel] = 1 - dirf[sel]

    if renumber == "hierarchical":
        # deal.II-like numbering: cells in z-order inside each patch (the order hierarchical
        # refinement produces), dofs numbered at their first appearance in that cell sequence.
        # A band |i-j| < 50 in this numbering covers a compact 2-D neighbourhood, as in the
        # reference, instead of a strip of a lexicographic grid.
        order = np.lexsort((np.concatenate(morton_l), cpatch))
        cells, dirf, cpatch = cells[order], dirf[order], cpatch[order]
        flat = cells.ravel()
        _, first = np.unique(flat, return_index=True)
        seq = flat[np.sort(first)]                  # old dof ids in order of first appearance
        perm = np.empty(n, dtype=np.int64)           # old -> new
        perm[seq] = np.arange(n)
        inv = seq
        xyz = xyz[inv]
        bnd = bnd[inv]
        npatch = npatch[inv]
        cells = perm[cells]
    elif renumber == "random":
        rng = np.random.default_rng(seed)
        perm = rng.permutation(n)            # old -> new
        inv = np.empty(n, dtype=np.int64)
        inv[perm] = np.arange(n)
        xyz = xyz[inv]
        bnd = bnd[inv]
        npatch = npatch[inv]
        cells = perm[cells]
        cperm = rng.permutation(cells.shape[0])
        cells, dirf, cpatch = cells[cperm], dirf[cperm], cpatch[cperm]

    mesh = SurfaceMesh(xyz=xyz, cells=np.ascontiguousarray(cells, dtype=np.uint32),
                       dir_flag=np.ascontiguousarray(dirf, dtype=np.uint8), cell_patch=cpatch,
                       node_patch=npatch, node_on_patch_boundary=bnd, patch_names=list(names))
    mesh.dn_ptr, mesh.dn_idx = generate_double_nodes_set(xyz, bnd)
    return mesh


def generate_double_nodes_set(xyz, boundary_dofs, tol=1e-8):
    """ComputationalDomain::generate_double_nodes_set (source/computational_domain.cc:258-307):
    set[i] = {i} U {j : |x_i - x_j| < tol} for boundary dofs i, {i} otherwise.  CSR, sorted."""
    from scipy.spatial import cKDTree

    n = xyz.shape[0]
    sets = [[i] for i in range(n)]
    b = np.nonzero(boundary_dofs)[0]
    if len(b):
        tree = cKDTree(xyz)
        nb = tree.query_ball_point(xyz[b], r=tol * (1 - 1e-9))
        for i, lst in zip(b, nb):
            s = set(lst)
            s.add(int(i))
            sets[i] = sorted(s)
    ptr = np.zeros(n + 1, dtype=np.uint32)
    ptr[1:] = np.cumsum([len(s) for s in sets])
    idx = np.fromiter((j for s in sets for j in s), dtype=np.uint32, count=int(ptr[-1]))
    return ptr, idx


def _graded(a, b, n, first_frac):
    """n cells from a to b, geometric sizes, first cell = first_frac * uniform size."""
    if n <= 0:
        return np.array([a, b], dtype=np.float64)[:1]
    if n == 1 or abs(first_frac - 1.0) < 1e-12:
        return np.linspace(a, b, n + 1)
    # solve sum_{k<n} r^k = n / first_frac
    target = n / first_frac
    lo, hi = (1.0 + 1e-12, 50.0) if first_frac < 1 else (1e-3, 1.0 - 1e-12)
    for _ in range(200):
        r = 0.5 * (lo + hi)
        s = (r ** n - 1) / (r - 1)
        if (s < target) == (first_frac < 1):
            lo = r
        else:
            hi = r
    r = 0.5 * (lo + hi)
    sizes = r ** np.arange(n)
    t = np.concatenate([[0.0], np.cumsum(sizes)])
    t /= t[-1]
    return a + (b - a) * t


# ----------------------------------------------------------------------------------------
# closed unit cube / sphere (known-answer tests)
# ----------------------------------------------------------------------------------------
def _cube_faces(n, side):
    g = np.linspace(0.0, side, n + 1)
    U, V = np.meshgrid(g, g, indexing="xy")     # U varies fastest (axis 1)
    Z0 = np.zeros_like(U)
    S = np.full_like(U, side)
    # (points, outward normal is +d_u x d_v)
    return {
        "z0": (np.stack([U, V, Z0], -1), False),   # x cross y = +z ; outward is -z
        "z1": (np.stack([U, V, S], -1), True),
        "y0": (np.stack([U, Z0, V], -1), True),    # x cross z = -y ; outward at y=0
        "y1": (np.stack([U, S, V], -1), False),
        "x0": (np.stack([Z0, U, V], -1), False),   # y cross z = +x ; outward is -x
        "x1": (np.stack([S, U, V], -1), True),
    }


def cube(n=4, side=1.0, renumber=None, seed=0, flip_every=0):
    """Closed cube [0,side]^3, 6 patches of n x n cells, outward normals, patch-wise double
    nodes along the 12 edges (triple at the 8 corners)."""
    faces = _cube_faces(n, side)
    patches = [_patch(pts, out) for pts, out in faces.values()]
    m = _finish(patches, list(faces.keys()), renumber, seed, flip_every)
    m.meta = dict(kind="cube", n=n, side=side)
    return m


def sphere(n=4, radius=1.0, center=(0.0, 0.0, 0.0), renumber=None, seed=0, flip_every=0):
    """Cube-sphere: the 6 cube patches projected radially onto a sphere."""
    faces = _cube_faces(n, 2.0)
    patches = []
    for pts, out in faces.values():
        p = pts - 1.0
        p = p / np.linalg.norm(p, axis=-1, keepdims=True) * radius + np.asarray(center)
        patches.append(_patch(p, out))
    m = _finish(patches, list(faces.keys()), renumber, seed, flip_every)
    m.meta = dict(kind="sphere", n=n, radius=radius, center=tuple(center))
    return m


# ----------------------------------------------------------------------------------------
# towing tank + Wigley hull
# ----------------------------------------------------------------------------------------
def wigley_y(x, z, L=WIGLEY_L, B=WIGLEY_B, T=WIGLEY_T):
    """BoatSurface::HullFunction, y >= 0 side (source/boat_surface.cc:93)."""
    return 0.5 * B * (1.0 - (2.0 * x / L) ** 2) * (1.0 - (z / T) ** 2)


def wigley_tank(nxm=24, nt=10, nxu=8, nxd=12, nz=6, nzh=6, grade=0.25, renumber=None, seed=0,
                flip_every=0, wave_amp=0.0, wave_k=2.0, wave_phase=0.0):
    """Tank [-6L,6L] x [-2L,2L] x [-2L,0] with the Wigley hull piercing the free surface.

    nxm cells along the hull, nt across each half free surface, nxu/nxd up/downstream,
    nz over the tank depth, nzh over the hull draught.  `grade` < 1 clusters cells towards
    the hull.  wave_amp > 0 displaces free-surface nodes by a travelling wave (used to
    emulate the geometry motion between IDA residual evaluations)."""
    L, B, T = WIGLEY_L, WIGLEY_B, WIGLEY_T
    Lx, Ly, Lz = 12 * L, 4 * L, 2 * L
    xu = _graded(-L / 2, -Lx / 2, nxu, grade)[::-1]       # inflow -> bow
    xm = -0.5 * L * np.cos(np.pi * np.arange(nxm + 1) / nxm) if nxm > 1 else np.array([-L / 2, L / 2])
    xm = 0.5 * (xm + np.linspace(-L / 2, L / 2, nxm + 1))   # mild end clustering
    xd = _graded(L / 2, Lx / 2, nxd, grade)
    gt = _graded(0.0, 1.0, nt, grade)                      # across the half width, from hull
    zt = _graded(0.0, -Lz, nz, grade)                      # tank depth from the surface
    zh = _graded(0.0, -T, nzh, 1.0)                        # hull draught
    yfull = np.concatenate([-(Ly / 2) * gt[::-1], (Ly / 2) * gt[1:]])
    xall = np.concatenate([xu, xm[1:], xd[1:]])

    def eta(x, y):
        if wave_amp == 0.0:
            return np.zeros_like(x)
        # the wave is switched off near the hull (so waterline nodes stay on the hull patch)
        d = np.sqrt(np.maximum(np.abs(x) - L / 2, 0.0) ** 2 + y ** 2)
        rho = np.clip((d - 0.15 * L) / (0.35 * L), 0.0, 1.0)
        return wave_amp * np.cos(wave_k * x - wave_phase) * rho

    patches, names = [], []

    def fs_rect(xs, name):
        X, Y = np.meshgrid(xs, yfull, indexing="xy")
        patches.append(_patch(np.stack([X, Y, eta(X, Y)], -1), True))
        names.append(name)

    fs_rect(xu, "fs_up")
    fs_rect(xd, "fs_down")
    yw = wigley_y(xm, 0.0)
    for sgn, name in ((1.0, "fs_mid_right"), (-1.0, "fs_mid_left")):
        X = np.tile(xm, (nt + 1, 1))
        Y = sgn * (yw[None, :] + gt[:, None] * (Ly / 2 - yw[None, :]))
        Z = eta(X, Y)
        # x cross y = +z for sgn>0 ; for the mirrored side orientation reverses
        patches.append(_patch(np.stack([X, Y, Z], -1), sgn > 0))
        names.append(name)
    for sgn, name in ((1.0, "hull_right"), (-1.0, "hull_left")):
        X = np.tile(xm, (nzh + 1, 1))
        Zg = np.tile(zh[:, None], (1, nxm + 1))
        Y = sgn * wigley_y(X, Zg)
        # d_x cross d_z(downwards): for the y>0 side the water lies at larger y; the normal
        # pointing out of the water points towards the hull centre plane (-y).
        patches.append(_patch(np.stack([X, Y, Zg], -1), sgn < 0))
        names.append(name)
    for xs, name, out_is_uxv in ((-Lx / 2, "inflow", True), (Lx / 2, "outflow", False)):
        Yg, Zg = np.meshgrid(yfull, zt, indexing="xy")
        # d_y x d_z(down) = -x: outward at the inflow wall
        patches.append(_patch(np.stack([np.full_like(Yg, xs), Yg, Zg], -1), out_is_uxv))
        names.append(name)
    for ys, name, out_is_uxv in ((-Ly / 2, "side_left", False), (Ly / 2, "side_right", True)):
        Xg, Zg = np.meshgrid(xall, zt, indexing="xy")
        # d_x x d_z(down) = +y
        patches.append(_patch(np.stack([Xg, np.full_like(Xg, ys), Zg], -1), out_is_uxv))
        names.append(name)
    Xg, Yg = np.meshgrid(xall, yfull, indexing="xy")
    patches.append(_patch(np.stack([Xg, Yg, np.full_like(Xg, -Lz)], -1), False))
    names.append("bottom")

    m = _finish(patches, names, renumber, seed, flip_every)
    fs = np.isin(m.node_patch, [names.index(k) for k in names if k.startswith("fs_")])
    m.surface_nodes = fs.astype(np.float64)
    m.other_nodes = 1.0 - m.surface_nodes
    m.meta = dict(kind="wigley_tank", nxm=nxm, nt=nt, nxu=nxu, nxd=nxd, nz=nz, nzh=nzh, grade=grade,
                  L=L, B=B, T=T, hull_patches=[names.index("hull_right"), names.index("hull_left")])
    return m


def wigley_tank_for_nodes(n_target, **kw):
    kw.setdefault("renumber", "hierarchical")
    """Pick the resolution whose node count is closest to n_target (keeps the aspect of the
    default resolution: most nodes near the hull and on the free surface)."""
    best = None
    for nxm in range(6, 2000):
        nt = max(3, int(round(nxm * 0.42)))
        nxu = max(2, int(round(nxm * 0.33)))
        nxd = max(3, int(round(nxm * 0.5)))
        nz = max(2, int(round(nxm * 0.25)))
        nzh = max(2, int(round(nxm * 0.25)))
        nx = nxu + nxm + nxd
        n = ((nxu + 1) + (nxd + 1)) * (2 * nt + 1) + 2 * (nxm + 1) * (nt + 1) + 2 * (nxm + 1) * (nzh + 1) \
            + 2 * (2 * nt + 1) * (nz + 1) + 2 * (nx + 1) * (nz + 1) + (nx + 1) * (2 * nt + 1)
        if best is None or abs(n - n_target) < abs(best[0] - n_target):
            best = (n, dict(nxm=nxm, nt=nt, nxu=nxu, nxd=nxd, nz=nz, nzh=nzh))
        if n > n_target * 1.2:
            break
    return wigley_tank(**best[1], **kw)


# ----------------------------------------------------------------------------------------
# boundary data
# ----------------------------------------------------------------------------------------
def cell_normals_at_nodes(mesh: SurfaceMesh):
    """Area-weighted average of the adjacent cells' normals at every dof (a lumped stand-in
    for ComputationalDomain::compute_normals, source/computational_domain.cc:1525-1620)."""
    X = mesh.xyz[mesh.cells.astype(np.int64)]
    du = 0.5 * ((X[:, 1] - X[:, 0]) + (X[:, 3] - X[:, 2]))
    dv = 0.5 * ((X[:, 2] - X[:, 0]) + (X[:, 3] - X[:, 1]))
    c = np.cross(du, dv) * np.where(mesh.dir_flag[:, None] > 0, 1.0, -1.0)
    nn = np.zeros_like(mesh.xyz)
    for k in range(4):
        np.add.at(nn, mesh.cells[:, k].astype(np.int64), c)
    nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    return nn


def towing_tank_bc(mesh: SurfaceMesh, froude=0.28, g=9.81):
    """tmp_rhs for the steady double-body-like problem of SURVEY 8(d) config 1: phi = 0 on the
    free surface, dphi_dn = -n . Vinf on the hull, 0 on the tank walls
    (pattern of source/free_surface.cc:8798-8849)."""
    vinf = np.array([froude * np.sqrt(g * mesh.meta["L"]), 0.0, 0.0])
    nn = cell_normals_at_nodes(mesh)
    bc = np.zeros(mesh.n_nodes)
    hull = np.isin(mesh.node_patch, mesh.meta["hull_patches"])
    bc[hull] = -(nn[hull] @ vinf)
    return bc


# ----------------------------------------------------------------------------------------
# local refinement with hanging nodes
# ----------------------------------------------------------------------------------------
def refine_cells(mesh: SurfaceMesh, cell_mask):
    """Split the marked cells into four children each (one level of deal.II's isotropic refinement of a
    quad).  Edge midpoints shared by two marked cells become ordinary dofs; a midpoint on an edge whose
    other cell stays coarse is a HANGING node, constrained to the mean of the edge's two vertices
    (DoFTools::make_hanging_node_constraints for FE_Q(1), reference source/bem_problem.cc:1000 and
    source/computational_domain.cc:1535-1538).  New nodes sit on the parent's bilinear map.
    Returns (refined mesh, hanging), hanging = [(dof, [(master, 0.5), (master, 0.5)]), ...]."""
    cell_mask = np.asarray(cell_mask, dtype=bool)
    cells = mesh.cells.astype(np.int64)
    n0 = mesh.n_nodes
    edges_of = ((0, 1), (0, 2), (1, 3), (2, 3))          # local vertex pairs, lexicographic vertex order
    edge_cells = {}
    for c, dofs in enumerate(cells):
        for a, b in edges_of:
            key = (min(dofs[a], dofs[b]), max(dofs[a], dofs[b]))
            edge_cells.setdefault(key, []).append(c)
    xyz = [mesh.xyz]
    node_patch = list(mesh.node_patch)
    on_bnd = list(mesh.node_on_patch_boundary)
    surf = list(mesh.surface_nodes) if mesh.surface_nodes is not None else None
    mid = {}
    hanging = []
    new_cells, new_dir, new_patch = [], [], []
    n = n0

    def new_node(pos, patch, bnd, s):
        nonlocal n
        xyz.append(np.asarray(pos, dtype=np.float64)[None, :])
        node_patch.append(patch)
        on_bnd.append(bnd)
        if surf is not None:
            surf.append(s)
        n += 1
        return n - 1

    for c, dofs in enumerate(cells):
        if not cell_mask[c]:
            new_cells.append(dofs)
            new_dir.append(mesh.dir_flag[c])
            new_patch.append(mesh.cell_patch[c])
            continue
        X = mesh.xyz[dofs]
        patch = mesh.cell_patch[c]
        sval = mesh.surface_nodes[dofs[0]] if surf is not None else 0.0
        m = []
        for a, b in edges_of:
            key = (min(dofs[a], dofs[b]), max(dofs[a], dofs[b]))
            if key not in mid:
                owners = edge_cells[key]
                is_bnd = len(owners) == 1                      # an edge of the patch boundary
                mid[key] = new_node(0.5 * (X[a] + X[b]), patch, is_bnd, sval)
                if not is_bnd and not all(cell_mask[o] for o in owners):
                    hanging.append((mid[key], [(int(key[0]), 0.5), (int(key[1]), 0.5)]))
            m.append(mid[key])
        e01, e02, e13, e23 = m
        ctr = new_node(0.25 * X.sum(axis=0), patch, False, sval)
        for child in ((dofs[0], e01, e02, ctr), (e01, dofs[1], ctr, e13), (e02, ctr, dofs[2], e23), (ctr, e13, e23, dofs[3])):
            new_cells.append(np.array(child, dtype=np.int64))
            new_dir.append(mesh.dir_flag[c])
            new_patch.append(patch)
    xyz = np.ascontiguousarray(np.concatenate(xyz), dtype=np.float64)
    out = SurfaceMesh(xyz=xyz, cells=np.ascontiguousarray(np.array(new_cells), dtype=np.uint32),
                      dir_flag=np.ascontiguousarray(new_dir, dtype=np.uint8), cell_patch=np.array(new_patch, dtype=np.int32),
                      node_patch=np.array(node_patch, dtype=np.int32), node_on_patch_boundary=np.array(on_bnd, dtype=bool),
                      patch_names=list(mesh.patch_names), meta=dict(mesh.meta))
    out.dn_ptr, out.dn_idx = generate_double_nodes_set(xyz, out.node_on_patch_boundary)
    if surf is not None:
        out.surface_nodes = np.array(surf, dtype=np.float64)
        out.other_nodes = 1.0 - out.surface_nodes
    return out, sorted(hanging)
