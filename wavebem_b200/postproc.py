"""Host-side hull force functional used for the "hull drag / potential" parity criterion.

Restates the steady terms of the pressure integration in the reference
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


def hull_pressure_force(mesh, phi, dphi_dn, vinf, rho=1025.1, g=9.81, hull_patches=None):
    """Returns (force[3], mean hull potential).  drag = force[0]."""
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
            area += jxw.sum()
            phi_int += ((P @ sh) * jxw).sum()
    return force, phi_int / area
