"""TEST INFRASTRUCTURE (oracle): the two L2 projections of compute_constraints WITH hanging-node
constraints, restated with scipy.sparse -- the checker of csrc/constraints.cu on locally refined meshes.

    ComputationalDomain<3>::compute_normals        source/computational_domain.cc:1525-1620
    BEMProblem<3>::compute_surface_gradients       source/bem_problem.cc:1153-1293

Both assemble the mass matrix of FESystem(FE_Q(1),3) (component-wise the scalar Q1 mass matrix) and a
right-hand side through vector_constraints.distribute_local_to_global (hanging-node lines of vector_dh,
:1535-1538 / :1167-1170), solve with SparseDirectUMFPACK and distribute().  That is the reduced system
    (C^T M C) x_free = C^T b ,   x = C x_free
with C the N x N_free matrix of the constraint lines.  Without hanging lines it reproduces
oracle/wbem_oracle.c's orc_l2_projection (checked in tests/test_oracle_kat.py).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

_G4 = np.array([-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526]) * 0.5 + 0.5
_W4 = np.array([0.3478548451374538, 0.6521451548625461, 0.6521451548625461, 0.3478548451374538]) * 0.5


def _assemble(xyz, cells, dir_flag, which, phi):
    cells = np.asarray(cells, dtype=np.int64)
    n = xyz.shape[0]
    X = xyz[cells]                                   # (C,4,3)
    sgn = np.where(np.asarray(dir_flag) > 0, 1.0, -1.0)
    lm = np.zeros((cells.shape[0], 4, 4))
    lr = np.zeros((cells.shape[0], 4, 3))
    P = None if phi is None else np.asarray(phi)[cells]
    for iv, v in enumerate(_G4):
        for iu, u in enumerate(_G4):
            w = _W4[iu] * _W4[iv]
            sh = np.array([(1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v])
            du = np.array([-(1 - v), (1 - v), -v, v])
            dv = np.array([-(1 - u), -u, (1 - u), u])
            tu = np.einsum("k,ckd->cd", du, X)
            tv = np.einsum("k,ckd->cd", dv, X)
            g00, g01, g11 = (tu * tu).sum(1), (tu * tv).sum(1), (tv * tv).sum(1)
            det = g00 * g11 - g01 * g01
            jxw = np.sqrt(det) * w
            if which == 0:
                cr = np.cross(tu, tv)
                vec = sgn[:, None] * cr / np.linalg.norm(cr, axis=1)[:, None]
            else:
                pu, pv = P @ du, P @ dv
                a = (g11 * pu - g01 * pv) / det
                b = (g00 * pv - g01 * pu) / det
                vec = a[:, None] * tu + b[:, None] * tv
            lm += jxw[:, None, None] * (sh[:, None] * sh[None, :])[None]
            lr += jxw[:, None, None] * sh[None, :, None] * vec[:, None, :]
    rows = np.repeat(cells, 4, axis=1).ravel()
    cols = np.tile(cells, (1, 4)).ravel()
    M = sp.coo_matrix((lm.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    B = np.zeros((n, 3))
    for k in range(4):
        np.add.at(B, cells[:, k], lr[:, k, :])
    return M, B


def l2_projection(which, xyz, cells, dir_flag, phi=None, hanging=None):
    """which = 0: node normals (normalised, :1613); 1: surface gradients of the nodal field phi."""
    xyz = np.asarray(xyz, dtype=np.float64)
    n = xyz.shape[0]
    M, B = _assemble(xyz, cells, dir_flag, which, phi)
    touched = np.zeros(n, dtype=bool)
    touched[np.asarray(cells, dtype=np.int64).ravel()] = True
    hanging = hanging or []
    is_h = np.zeros(n, dtype=bool)
    for h, _ in hanging:
        is_h[h] = True
    free = np.nonzero(~is_h & touched)[0]
    col_of = -np.ones(n, dtype=np.int64)
    col_of[free] = np.arange(len(free))
    r, c, v = list(free), list(col_of[free]), [1.0] * len(free)
    for h, ent in hanging:
        for m, w in ent:
            assert not is_h[m], "hanging node constrained to a hanging node"
            r.append(h)
            c.append(col_of[m])
            v.append(w)
    Cm = sp.coo_matrix((v, (r, c)), shape=(n, len(free))).tocsc()
    A = (Cm.T @ M @ Cm).tocsc()
    xf = spla.splu(A).solve(Cm.T @ B)
    x = Cm @ xf
    if which == 0:
        nrm = np.linalg.norm(x, axis=1)
        x = np.where(nrm[:, None] > 0, x / np.where(nrm > 0, nrm, 1.0)[:, None], x)
    return x
