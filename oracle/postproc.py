"""TEST INFRASTRUCTURE (oracle): numpy restatements of the two post-processing integrals, the
checkers of wbem_pressure_force / wbem_internal_velocities (csrc/postproc.cu).  Only tests/,
__graft_entry__.smoke() and bench.py's checker legs may import this.

hull_pressure_force restates the steady terms of the pressure integration in the reference
(source/free_surface.cc:9534-9598):

    gradient = n * dphi_dn + grad_s(phi)
    press    = rho |Vinf|^2 / 2 - rho |gradient + Vinf|^2 / 2 - rho g z
    force    = sum over hull cells, q:  press * n * JxW            (press_force_test_1)

with Q1 shape functions on the Gauss 4x4 rule.  O(cells) numpy work, not a kernel.
"""
from __future__ import annotations

import numpy as np

_G4 = np.array([-0.8611363115940526, -0.3399810435848563, 0.3399810435848563, 0.8611363115940526]) * 0.5 + 0.5
_W4 = np.array([0.3478548451374538, 0.6521451548625461, 0.6521451548625461, 0.3478548451374538]) * 0.5


def hull_pressure_force(mesh, phi, dphi_dn, vinf, rho=1025.1, g=9.81, hull_patches=None, full=False,
                        baricenter=(0.0, 0.0, 0.0)):
    """Returns (force[3], mean hull potential); drag = force[0].  full=True: the 11 integrals of
    wbem_pressure_force (press_force_test_1[3], press_force_test_2[3], press_moment[3], area, int phi)."""
    if hull_patches is None:
        hull_patches = mesh.meta["hull_patches"]
    sel = np.isin(mesh.cell_patch, hull_patches)
    cells = mesh.cells[sel].astype(np.int64)
    sgn = np.where(mesh.dir_flag[sel] > 0, 1.0, -1.0)
    X = mesh.xyz[cells]                      # (C,4,3)
    P = np.asarray(phi)[cells]               # (C,4)
    Q = np.asarray(dphi_dn)[cells]
    vinf = np.asarray(vinf, dtype=np.float64)
    force = np.zeros(3)
    force2 = np.zeros(3)
    moment = np.zeros(3)
    bar = np.asarray(baricenter, dtype=np.float64)
    area = 0.0
    phi_int = 0.0
    for iv, v in enumerate(_G4):
        for iu, u in enumerate(_G4):
            w = _W4[iu] * _W4[iv]
            sh = np.array([(1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v])
            du = np.array([-(1 - v), (1 - v), -v, v])
            dv = np.array([-(1 - u), -u, (1 - u), u])
            tu = np.einsum("k,ckd->cd", du, X)
            tv = np.einsum("k,ckd->cd", dv, X)
            cr = np.cross(tu, tv)
            cn = np.linalg.norm(cr, axis=1)
            nrm = sgn[:, None] * cr / cn[:, None]
            jxw = cn * w
            # surface gradient: J (J^T J)^-1 [dphi/du, dphi/dv]
            g00 = (tu * tu).sum(1)
            g01 = (tu * tv).sum(1)
            g11 = (tv * tv).sum(1)
            det = g00 * g11 - g01 * g01
            pu = P @ du
            pv = P @ dv
            a = (g11 * pu - g01 * pv) / det
            b = (-g01 * pu + g00 * pv) / det
            grad_s = a[:, None] * tu + b[:, None] * tv
            grad = nrm * (Q @ sh)[:, None] + grad_s
            z = np.einsum("k,ck->c", sh, X[:, :, 2])
            tot = grad + vinf[None, :]
            press = rho * (vinf @ vinf) / 2 - rho * (tot * tot).sum(1) / 2 - rho * g * z
            force += ((press * jxw)[:, None] * nrm).sum(0)
            press2 = -rho * (grad @ vinf) - rho * (grad * grad).sum(1) / 2 - rho * g * z
            force2 += ((press2 * jxw)[:, None] * nrm).sum(0)
            y = np.einsum("k,ckd->cd", sh, X)
            moment += ((press * jxw)[:, None] * np.cross(y - bar[None, :], nrm)).sum(0)
            area += jxw.sum()
            phi_int += ((P @ sh) * jxw).sum()
    if full:
        return np.concatenate([force, force2, moment, [area, phi_int]])
    return force, phi_int / area


def internal_velocities(mesh, phi, dphi_dn, points):
    """FreeSurface<3>::compute_internal_velocities (source/free_surface.cc:10426-10537):
        v(x) = sum_cells sum_q [ dphi_dn(q) grad_x G - phi(q) grad_x dG/dn ] JxW(q)
    G = 1/(4 pi |r|), dG/dn = -(r.n)/(4 pi |r|^3), r = y_q - x.  The reference differentiates with
    Sacado; here the gradients are written out (tests/test_oracle_kat.py checks them against
    finite differences of the same potential)."""
    cells = mesh.cells.astype(np.int64)
    sgn = np.where(mesh.dir_flag > 0, 1.0, -1.0)
    X = mesh.xyz[cells]
    P = np.asarray(phi)[cells]
    Q = np.asarray(dphi_dn)[cells]
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    vel = np.zeros_like(pts)
    for iv, v in enumerate(_G4):
        for iu, u in enumerate(_G4):
            w = _W4[iu] * _W4[iv]
            sh = np.array([(1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v])
            du = np.array([-(1 - v), (1 - v), -v, v])
            dv = np.array([-(1 - u), -u, (1 - u), u])
            tu = np.einsum("k,ckd->cd", du, X)
            tv = np.einsum("k,ckd->cd", dv, X)
            cr = np.cross(tu, tv)
            cn = np.linalg.norm(cr, axis=1)
            nrm = sgn[:, None] * cr / cn[:, None]
            jxw = cn * w
            y = np.einsum("k,ckd->cd", sh, X)
            qphi, qdphi = P @ sh, Q @ sh
            for i, x in enumerate(pts):
                r = y - x[None, :]
                rn = np.linalg.norm(r, axis=1)
                grad_g = r / (4 * np.pi * rn[:, None] ** 3)
                rdotn = (r * nrm).sum(1)
                grad_dgdn = (nrm / rn[:, None] ** 3 - 3 * (rdotn / rn ** 5)[:, None] * r) / (4 * np.pi)
                vel[i] += ((qdphi * jxw)[:, None] * grad_g - (qphi * jxw)[:, None] * grad_dgdn).sum(0)
    return vel


def potential_at(mesh, phi, dphi_dn, points):
    """The representation formula itself, phi(x) = int (G dphi_dn - phi dG/dn) dS: what the velocities
    above are the gradient of (used by the finite-difference KAT)."""
    cells = mesh.cells.astype(np.int64)
    sgn = np.where(mesh.dir_flag > 0, 1.0, -1.0)
    X = mesh.xyz[cells]
    P = np.asarray(phi)[cells]
    Q = np.asarray(dphi_dn)[cells]
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    out = np.zeros(len(pts))
    for iv, v in enumerate(_G4):
        for iu, u in enumerate(_G4):
            w = _W4[iu] * _W4[iv]
            sh = np.array([(1 - u) * (1 - v), u * (1 - v), (1 - u) * v, u * v])
            du = np.array([-(1 - v), (1 - v), -v, v])
            dv = np.array([-(1 - u), -u, (1 - u), u])
            cr = np.cross(np.einsum("k,ckd->cd", du, X), np.einsum("k,ckd->cd", dv, X))
            cn = np.linalg.norm(cr, axis=1)
            nrm = sgn[:, None] * cr / cn[:, None]
            jxw = cn * w
            y = np.einsum("k,ckd->cd", sh, X)
            for i, x in enumerate(pts):
                r = y - x[None, :]
                rn = np.linalg.norm(r, axis=1)
                G = 1.0 / (4 * np.pi * rn)
                dGdn = -(r * nrm).sum(1) / (4 * np.pi * rn ** 3)
                out[i] += (((Q @ sh) * G - (P @ sh) * dGdn) * jxw).sum()
    return out
