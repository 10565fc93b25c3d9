"""ctypes front-end of the CPU oracle (oracle/wbem_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never from wavebem_b200/.

PARITY UNPINNED: the reference (mathLab/WaveBEM) cannot be compiled here and ships no
known-answer test for this path; see the header of wbem_oracle.c and DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compile oracle/wbem_oracle.c -> oracle/_build/liboracle.so (make)."""
    src = os.path.join(_HERE, "wbem_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _opt(a, dtype):
    """optional array -> ctypes pointer or NULL"""
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a.ctypes.data_as(C.c_void_p), a


def gauss01(n, long_double=False):
    x = np.zeros(n)
    w = np.zeros(n)
    f = lib().orc_gauss01_ld if long_double else lib().orc_gauss01
    f.argtypes = [C.c_int, _dp, _dp]
    assert f(n, x, w) == 0
    return x, w


def qgauss2(n):
    uv = np.zeros((n * n, 2))
    w = np.zeros(n * n)
    f = lib().orc_qgauss2
    f.argtypes = [C.c_int, _dp, _dp]
    assert f(n, uv, w) == 0
    return uv, w


def qgauss_one_over_r(n, vertex, factor_out=True):
    uv = np.zeros((2 * n * n, 2))
    w = np.zeros(2 * n * n)
    f = lib().orc_qgauss_one_over_r
    f.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp]
    assert f(n, vertex, int(factor_out), uv, w) == 0
    return uv, w


def fe_values(X, dir_flag, uv, w, long_double=False):
    """FEValues<2,3> on one Q1 cell: returns q_points, normals, JxW, shape[4][nq]."""
    X = np.ascontiguousarray(X, dtype=np.float64).reshape(12)
    uv = np.ascontiguousarray(uv, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    nq = len(w)
    qp = np.zeros((nq, 3))
    nr = np.zeros((nq, 3))
    jw = np.zeros(nq)
    sh = np.zeros((4, nq))
    f = lib().orc_fe_values_ld if long_double else lib().orc_fe_values
    f.argtypes = [_dp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp]
    f.restype = None
    f(X, int(dir_flag), nq, uv, w, qp, nr, jw, sh)
    return qp, nr, jw, sh


def assemble_rows(xyz, cells, dir_flag, dn_ptr, dn_idx, row0=0, row1=None, quad_order=4,
                  sing_order=5, nthreads=0, long_double=False):
    """BEMProblem<3>::assemble_system restricted to rows [row0,row1) -> (N_mat, D_mat)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    cells = np.ascontiguousarray(cells, dtype=np.uint32)
    dir_flag = np.ascontiguousarray(dir_flag, dtype=np.uint8)
    dn_ptr = np.ascontiguousarray(dn_ptr, dtype=np.uint32)
    dn_idx = np.ascontiguousarray(dn_idx, dtype=np.uint32)
    n = xyz.shape[0]
    c = cells.shape[0]
    if row1 is None:
        row1 = n
    nm = np.empty((row1 - row0, n))
    dm = np.empty((row1 - row0, n))
    f = lib().orc_assemble_rows_ld if long_double else lib().orc_assemble_rows
    f.argtypes = [C.c_int, C.c_int, _dp, _u32p, _u8p, _u32p, _u32p, C.c_int, C.c_int, C.c_int,
                  C.c_int, _dp, _dp, C.c_int]
    rc = f(n, c, xyz, cells, dir_flag, dn_ptr, dn_idx, quad_order, sing_order, row0, row1, nm, dm,
           nthreads)
    assert rc == 0
    return nm, dm


def compute_alpha(nm, nthreads=0):
    nrows, n = nm.shape
    alpha = np.zeros(nrows)
    f = lib().orc_compute_alpha
    f.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int]
    f.restype = None
    f(nrows, n, np.ascontiguousarray(nm), alpha, nthreads)
    return alpha


def fullmatrix_vmult(a, v, nthreads=0):
    nrows, n = a.shape
    w = np.zeros(nrows)
    f = lib().orc_fullmatrix_vmult
    f.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, C.c_int]
    f.restype = None
    f(nrows, n, np.ascontiguousarray(a), w, np.ascontiguousarray(v, dtype=np.float64), 0, nthreads)
    return w


def vmult(nm, dm, alpha, surface_nodes, other_nodes, src, nthreads=0):
    n = nm.shape[0]
    dst = np.zeros(n)
    f = lib().orc_vmult
    f.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int]
    f.restype = None
    f(n, nm, dm, np.ascontiguousarray(alpha), np.ascontiguousarray(surface_nodes),
      np.ascontiguousarray(other_nodes), dst, np.ascontiguousarray(src, dtype=np.float64), nthreads)
    return dst


def compute_rhs(nm, dm, alpha, surface_nodes, other_nodes, src, nthreads=0):
    n = nm.shape[0]
    dst = np.zeros(n)
    f = lib().orc_compute_rhs
    f.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int]
    f.restype = None
    f(n, nm, dm, np.ascontiguousarray(alpha), np.ascontiguousarray(surface_nodes),
      np.ascontiguousarray(other_nodes), dst, np.ascontiguousarray(src, dtype=np.float64), nthreads)
    return dst


class Constraints:
    """Flattened deal.II ConstraintMatrix: lines[k] is the constrained dof of line k,
    entries CSR (ptr, col, val), inhom[k]."""

    def __init__(self, n, lines=(), ptr=(0,), col=(), val=(), inhom=()):
        self.n = n
        self.lines = np.ascontiguousarray(lines, dtype=np.uint32)
        self.ptr = np.ascontiguousarray(ptr, dtype=np.uint32)
        self.col = np.ascontiguousarray(col, dtype=np.uint32)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.inhom = np.ascontiguousarray(inhom, dtype=np.float64)
        self.line_of = np.full(n, -1, dtype=np.int32)
        self.line_of[self.lines] = np.arange(len(self.lines), dtype=np.int32)

    def _args(self):
        def p(a):
            return a.ctypes.data_as(C.c_void_p) if a.size else None
        return (self.line_of.ctypes.data_as(C.c_void_p), self.ptr.ctypes.data_as(C.c_void_p),
                p(self.col), p(self.val), p(self.inhom))


def constrained_vmult(nm, dm, alpha, surface_nodes, other_nodes, con: Constraints, src, nthreads=0):
    n = nm.shape[0]
    dst = np.zeros(n)
    f = lib().orc_constrained_vmult
    f.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p, _dp, _dp, C.c_int]
    f.restype = None
    a = con._args()
    f(n, nm, dm, np.ascontiguousarray(alpha), np.ascontiguousarray(surface_nodes),
      np.ascontiguousarray(other_nodes), a[0], a[1], a[2], a[3], dst,
      np.ascontiguousarray(src, dtype=np.float64), nthreads)
    return dst


def distribute_rhs(con: Constraints, rhs):
    rhs = np.array(rhs, dtype=np.float64)
    f = lib().orc_distribute_rhs
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, _dp]
    f.restype = None
    a = con._args()
    f(con.n, a[0], a[4], rhs)
    return rhs


def band_system_dense(nm, dm, alpha, surface_nodes, con: Constraints, band=100):
    n = nm.shape[0]
    out = np.zeros((n, n))
    f = lib().orc_band_system_dense
    f.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_void_p, C.c_int, _dp]
    f.restype = None
    f(n, nm, dm, np.ascontiguousarray(alpha), np.ascontiguousarray(surface_nodes), con._args()[0],
      band, out)
    return out


def precond_apply(nm, dm, alpha, surface_nodes, con: Constraints, vec, band=100):
    n = nm.shape[0]
    out = np.zeros(n)
    f = lib().orc_precond_apply
    f.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_void_p, C.c_int, _dp, _dp]
    rc = f(n, nm, dm, np.ascontiguousarray(alpha), np.ascontiguousarray(surface_nodes),
           con._args()[0], band, np.ascontiguousarray(vec, dtype=np.float64), out)
    assert rc == 0
    return out


def solve_system(nm, dm, surface_nodes, other_nodes, tmp_rhs, con: Constraints, phi, dphi_dn,
                 tol=1e-16, max_steps=200, n_tmp_vectors=100, band=100, use_precond=True,
                 nthreads=0):
    """BEMProblem<3>::solve_system.  Returns dict(phi, dphi_dn, alpha, rhs, sol, iters,
    last_res, res_hist, converged)."""
    n = nm.shape[0]
    phi = np.array(phi, dtype=np.float64)
    dphi_dn = np.array(dphi_dn, dtype=np.float64)
    alpha = np.zeros(n)
    rhs = np.zeros(n)
    sol = np.zeros(n)
    iters = C.c_int(0)
    last = C.c_double(0)
    hist = np.full(max_steps + 2, np.nan)
    f = lib().orc_solve_system
    f.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, _dp,
                  _dp, _dp, _dp, _dp, C.POINTER(C.c_int), C.POINTER(C.c_double), _dp, C.c_int,
                  C.c_int]
    a = con._args()
    rc = f(n, nm, dm, np.ascontiguousarray(surface_nodes), np.ascontiguousarray(other_nodes),
           np.ascontiguousarray(tmp_rhs, dtype=np.float64), a[0], a[1], a[2], a[3], a[4], tol,
           max_steps, n_tmp_vectors, band, int(use_precond), phi, dphi_dn, alpha, rhs, sol,
           C.byref(iters), C.byref(last), hist, len(hist), nthreads)
    if rc < 0:
        raise RuntimeError("oracle: singular band preconditioner")
    return dict(phi=phi, dphi_dn=dphi_dn, alpha=alpha, rhs=rhs, sol=sol, iters=iters.value,
                last_res=last.value, res_hist=hist[: iters.value + 1], converged=(rc == 0))


def residual(nm, dm, surface_nodes, other_nodes, con: Constraints, phi, dphi_dn, nthreads=0):
    n = nm.shape[0]
    res = np.zeros(n)
    f = lib().orc_residual
    f.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_void_p, _dp, _dp, _dp, C.c_int]
    f.restype = None
    a = con._args()
    f(n, nm, dm, np.ascontiguousarray(surface_nodes), np.ascontiguousarray(other_nodes), a[0],
      a[1], a[2], a[3], a[4], np.ascontiguousarray(phi, dtype=np.float64),
      np.ascontiguousarray(dphi_dn, dtype=np.float64), res, nthreads)
    return res


def l2_projection(which, xyz, cells, dir_flag, phi=None, quad_order=4):
    """which = 0: ComputationalDomain::compute_normals (computational_domain.cc:1525-1620);
    which = 1: BEMProblem::compute_surface_gradients of the nodal field phi (bem_problem.cc:1153-1293)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    cells = np.ascontiguousarray(cells, dtype=np.uint32)
    dir_flag = np.ascontiguousarray(dir_flag, dtype=np.uint8)
    n = xyz.shape[0]
    phi = np.zeros(n) if phi is None else np.ascontiguousarray(phi, dtype=np.float64)
    out = np.zeros((n, 3))
    f = lib().orc_l2_projection
    f.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _u32p, _u8p, C.c_int, _dp, _dp]
    rc = f(int(which), n, cells.shape[0], xyz, cells, dir_flag, int(quad_order), phi, out)
    if rc:
        raise RuntimeError(f"oracle: L2 projection failed ({rc})")
    return out


def compute_normals(xyz, cells, dir_flag, quad_order=4):
    return l2_projection(0, xyz, cells, dir_flag, None, quad_order)


def compute_surface_gradients(xyz, cells, dir_flag, tmp_rhs, surface_nodes, quad_order=4):
    return l2_projection(1, xyz, cells, dir_flag, np.asarray(tmp_rhs) * np.asarray(surface_nodes), quad_order)


def max_threads():
    f = lib().orc_max_threads
    f.restype = C.c_int
    return f()
